"""Homogenised matrices from the minimisers — drop-in for ffthompy/postprocess.py.

`assembly_matrix` evaluates all D*D entries <A e_i, e_j> in ONE pass over the coefficient array on the device
(csrc/fh_pointwise.cu: fh_assemble_AH; the reference makes D*D matrix-vector passes, postprocess.py:66-69);
`postprocess` is the reference's bound-evaluation driver (postprocess.py:6-51): for every requested variant it takes
the coefficients from the material object — GaNi nodal values, or the exactly integrated (Ga) coefficients on the
double grid, onto which GaNi minimisers are interpolated spectrally — assembles A_H and, for the dual formulation,
inverts it, so that `mat_primal` / `mat_dual` hold guaranteed upper / lower bounds."""
import ctypes as C
import itertools

import numpy as np

from .general.base import Timer


def _one_pass(Afun, sol):
    """fh_assemble_AH if the operands have the plain solve-loop shape, else None"""
    from .tensors.objects import Tensor
    from . import device as dev, _lib as L
    dim = len(sol)
    if not isinstance(Afun, Tensor) or getattr(Afun, 'lazy', False) or Afun.Fourier or Afun.order != 2:
        return None
    if Afun.multype not in (21, '21') or Afun._is_complex() or tuple(Afun.shape) != (dim, dim) or dim not in (2, 3, 6):
        return None
    for s in sol:
        if not isinstance(s, Tensor) or s.Fourier or s.order != 1 or tuple(s.shape) != (dim,) or s._is_complex():
            return None
        if tuple(s.N) != tuple(Afun.N):
            return None
    n = int(np.prod(Afun.N))
    bufs = [s._dev() for s in sol]
    ptrs = (C.c_void_p*dim)(*[b.data_ptr() for b in bufs])
    out = (C.c_double*(dim*dim))()
    L.check(dev.lib().fh_assemble_AH(dim, dim, n, dev.ptr(Afun._dev()), ptrs, out))
    return np.array(out[:]).reshape(dim, dim)/float(n)


def assembly_matrix(Afun, solutions):
    """A_H[i, j] = <A e_i, e_j> (postprocess.py:53-70); minimisers living on another grid than the coefficients
    (GaNi solutions evaluated with Ga coefficients) are first interpolated spectrally (`Tensor.project`)."""
    dim = len(solutions)
    if np.allclose(Afun.N, solutions[0].N):
        sol = solutions
    else:
        sol = [s.project(Afun.N) for s in solutions]
    AH = _one_pass(Afun, sol)
    if AH is None:
        images = [Afun(s) for s in sol]
        AH = np.zeros([dim, dim])
        for ii, jj in itertools.product(range(dim), repeat=2):
            AH[ii, jj] = images[ii]*sol[jj]
    return AH


def _variant(pp, pb, mat, A, primaldual):
    """(name suffix, coefficients) of one entry of pb.postprocess (postprocess.py:13-37)"""
    kind = pp['kind']
    if kind in ('GaNi', 'gani'):
        if A.name != 'A_GaNi':
            A = mat.get_A_GaNi(pb.solve['N'], primaldual)
        return '', A
    if kind in ('Ga', 'ga'):
        if 'order' not in pp:
            return '', A                      # the coefficients of the solve
        Nbar = tuple(2*np.array(pb.solve['N'])-1)
        if pp['order'] is None:               # exact integration of the inclusion shapes
            return '', mat.get_A_Ga(Nbar=Nbar, primaldual=primaldual, order=None)
        tag = '_o%s_P%d' % (str(pp['order']), np.mean(pp['P']))
        return tag, mat.get_A_Ga(Nbar=Nbar, primaldual=primaldual, order=pp['order'], P=pp['P'])
    raise ValueError('postprocess kind %r' % (kind,))


def postprocess(pb, A, mat, solutions, results, primaldual):
    """evaluate every requested A_H variant of the minimisers and store sol_/res_/mat_<primaldual> in pb.output"""
    watch = Timer(name='postprocessing')
    print('\npostprocessing')
    matrices = {}
    for pp in pb.postprocess:
        tag, App = _variant(pp, pb, mat, A, primaldual)
        A = App                                # as in the reference, later variants see the last coefficients
        name = 'AH_%s%s_%s' % (pp['kind'], tag, primaldual)
        print('calculating: '+name)
        AH = assembly_matrix(App, solutions)
        matrices[name] = AH if primaldual == 'primal' else np.linalg.inv(AH)
    watch.measure()
    pb.output.update({'sol_'+primaldual: solutions, 'res_'+primaldual: results, 'mat_'+primaldual: matrices})


def add_macro2minimizer(X, E):
    """minimiser with the macroscopic value E as its mean, from a zero-mean or E-mean field (postprocess.py:73-86)"""
    mean = X.mean()
    if np.allclose(mean, E):
        return X
    if not np.allclose(mean, np.zeros_like(E)):
        raise ValueError("Field is neither zero-mean nor E-mean.")
    EN = X.zeros_like(name='EN')
    EN.set_mean(E)
    return X+EN
