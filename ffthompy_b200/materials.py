"""Material coefficients on the device — drop-in for ffthompy/materials.py (`Material`): the producer of the solve
loop's largest input.

Same configuration dictionaries, checks, method names and results as the reference:

* `get_A_GaNi(N, primaldual)` (materials.py:116-124): coefficients at the nodal points.  The characteristic functions
  of the inclusions ('cube'/'square', 'ball'/'circle', 'pyramid', 'otherwise', 'all') are evaluated per voxel by one
  kernel (fh_topologies) with the reference's own comparisons — the 27-fold periodically tiled coordinate arrays of
  materials.py:228-247 are never built — and combined with the phase matrices in the reference's summation order
  (fh_combine_phases), so the array is bit-identical to the reference's.
* `get_A_Ga(Nbar, primaldual, order, P)` (materials.py:54-114): exactly integrated coefficients on the doubled grid.
  order None: Fourier coefficients of the inclusion shapes (sinc / Bessel weights times the shift factor) -> centred
  inverse FFT on the device; order 0 / 1: nodal values on the grid P -> centred FFT, periodic extension or truncation
  to Nbar, times the weights of the piecewise constant / bilinear basis -> centred inverse FFT, all components batched.
  The 1-D weight vectors (and the 2-D Bessel weights of a ball, which need scipy.special.jn) are host metadata.
* the dual formulation inverts per voxel on the device (fh_inv_dxd), `.shift()` is the device roll kernel.
"""
import ctypes as C
import itertools

import numpy as np

from . import _lib as L
from . import device as dev
from . import ops
from .tensors import Tensor
from .trigpol import Grid, mean_index

inclusion_keys = {'ball': ['ball', 'circle'],
                  'cube': ['cube', 'square'],
                  'pyramid': ['bilinear_pyramid', 'pyramid']}
_KIND = {'cube': 0, 'ball': 1, 'pyramid': 2, 'otherwise': 3, 'all': 4}


def _kind_of(name):
    for k, names in inclusion_keys.items():
        if name in names:
            return _KIND[k]
    if name in ('otherwise', 'all'):
        return _KIND[name]
    raise NotImplementedError("Inclusion (%s) is not implemented." % (name,))


class Material(object):

    def __init__(self, material_conf):
        self.conf = material_conf
        if 'Y' not in self.conf:
            raise ValueError("The definition of PUC size (Y) is missing!")
        self.Y = material_conf['Y']
        if 'fun' in self.conf:
            return
        if 'inclusions' not in self.conf:
            raise NotImplementedError("Improper material definition!")
        n_incl = len(self.conf['inclusions'])
        for key in ('inclusions', 'positions', 'params', 'vals'):
            if key in self.conf and len(self.conf[key]) != n_incl:
                raise ValueError("Improper no. of values in material for (%s)!" % key)
        for ii, incl in enumerate(self.conf['inclusions']):
            if incl in ['all', 'otherwise']:
                continue
            try:
                if any(np.greater(self.conf['params'][ii], self.Y)):
                    raise ValueError("Improper parameters of inclusion!")
                self.conf['positions'][ii] = self.conf['positions'][ii] % self.Y
            except Exception:
                raise ValueError("Improper material definition!")

    # ------------------------------------------------------------------ nodal values
    def get_A_GaNi(self, N, primaldual='primal', tensor=True):
        """coefficients for the scheme with numerical integration (materials.py:116-124)"""
        if not tensor:
            raise NotImplementedError('the legacy Matrix form of get_A_GaNi raises in the reference too (SURVEY D.10)')
        A = self.evaluate_on(N)
        if primaldual == 'dual':
            A = A.inv()
        return A.shift()

    def evaluate_on(self, N):
        """`evaluate(Grid.get_coordinates(N, Y))` without building the coordinate arrays"""
        N = tuple(int(n) for n in np.array(N).ravel())
        if 'fun' in self.conf:
            return self.evaluate(Grid.get_coordinates(np.array(N), self.Y))
        chars = self._topologies(N)
        vals = [np.asarray(v, dtype=float) for v in self.conf['vals']]
        shp = vals[0].shape
        val = _combine(vals, chars, int(np.prod(N)))
        return Tensor(name='A_GaNi', val=val.reshape(shp+N), order=len(shp), N=N, Y=self.Y,
                      multype={2: 21, 4: 42}[len(shp)], Fourier=False, origin='c')

    def evaluate(self, coord, tensor=True):
        """coefficients at given coordinates (materials.py:172-214); inclusion-based materials are evaluated on the
        regular grid the coordinates span (they are produced by Grid.get_coordinates)"""
        if not tensor:
            raise NotImplementedError('legacy Matrix output')
        N = tuple(int(n) for n in coord.shape[1:])
        if 'fun' in self.conf:
            A_val = np.asarray(self.conf['fun'](coord), dtype=float)
            shp = A_val.shape[:A_val.ndim-len(N)]
            return Tensor(name='A_GaNi', val=A_val, order=len(shp), N=N, Y=self.Y, multype={2: 21, 4: 42}[len(shp)],
                          Fourier=False, origin='c')
        ref = Grid.get_coordinates(np.array(N), self.Y)
        if ref.shape != coord.shape or not np.array_equal(ref, coord):
            raise NotImplementedError('inclusion materials are evaluated on regular grids (Grid.get_coordinates)')
        return self.evaluate_on(N)

    def get_topologies(self, coord):
        """characteristic functions of the inclusions at the nodes of a regular grid (materials.py:216-308), as host
        arrays like the reference's"""
        N = tuple(int(n) for n in coord.shape[1:])
        chars = self._topologies(N)
        return [dev.download(chars[i]).reshape(N) for i in range(chars.shape[0])]

    def _topologies(self, N):
        d = len(N)
        Y = np.array(self.Y, dtype=float)
        ZN = Grid.get_ZNl(np.array(N), fft_form='c')
        xs = [Y[i]*ZN[i]/N[i] for i in range(d)]            # Grid.get_coordinates, one axis each (trigpol.py:55-71)
        coords = dev.upload(np.concatenate(xs).astype(float))
        incl = self.conf['inclusions']
        ninc = len(incl)
        kinds = [_kind_of(k) for k in incl]
        pos, par = np.zeros((ninc, 3)), np.zeros((ninc, 3))
        for ii, k in enumerate(kinds):
            if k in (_KIND['otherwise'], _KIND['all']):
                continue
            pos[ii, :d] = np.array(self.conf['positions'][ii], dtype=np.float64)
            par[ii, :d] = np.array(self.conf['params'][ii], dtype=np.float64)*np.ones(d)
        n = int(np.prod(N))
        out = dev.empty((ninc, n))
        flag = C.c_int()
        L.check(dev.lib().fh_topologies(d, L.i64arr(N), dev.ptr(coords), L.dblarr(np.concatenate([Y, np.ones(3-d)])),
                                        ninc, L.intarr(kinds), L.dblarr(pos.ravel()), L.dblarr(par.ravel()),
                                        dev.ptr(out), C.byref(flag)))
        if flag.value:
            raise NotImplementedError("Overlapping inclusions!")
        return out

    # ------------------------------------------------------------------ exact integration
    def get_A_Ga(self, Nbar, primaldual='primal', order=-1, P=None):
        """coefficients for the scheme with exact integration (materials.py:54-114)"""
        if order == -1:
            if 'order' in self.conf:
                order = self.conf['order']
            else:
                raise ValueError('The material order is undefined!')
        elif order not in [None, 'exact', 0, 1]:
            raise ValueError('Wrong material order (%s)!' % str(order))
        Nbar = tuple(int(n) for n in np.array(Nbar).ravel())
        n = int(np.prod(Nbar))
        if order in [None, 'exact']:     # inclusion-based composite
            chars = self.get_shape_functions(Nbar, on_device=True)
            vals = [np.asarray(v, dtype=float) for v in self.conf['vals']]
            if primaldual == 'dual':
                vals = [np.linalg.inv(v) for v in vals]
            shp = vals[0].shape
            val = _combine(vals, chars, n).reshape(shp+Nbar)
            name = 'A_Ga'
        else:                            # grid-based composite
            if P is None and 'P' in self.conf:
                P = self.conf['P']
            P = tuple(int(p) for p in np.array(P).ravel())
            vals = self.evaluate_on(P)
            if primaldual == 'dual':
                vals = vals.inv()
            h = np.array(self.Y, dtype=float)/np.array(P)
            W = _weights_1d(h, Nbar, self.Y, power=1 if order in [0, 'constant'] else 2)
            shp = tuple(vals.shape)
            ncomp = int(np.prod(shp))
            hAM0 = _cfftnc(vals._dev().reshape((ncomp,)+P), P, ncomp, scale=float(np.prod(P))/float(np.prod(P)))
            # np.prod(P)*cfftnc(...): the normalised centred transform times prod(P) = the plain centred DFT
            if np.allclose(P, Nbar):
                hAM = hAM0
            elif np.all(np.greater_equal(P, Nbar)) or np.all(np.less(P, Nbar)):
                hAM = _periodic_crop(hAM0, P, Nbar, ncomp)
            else:
                raise NotImplementedError("This combination of double N (%s) and P (%s) is not implemented."
                                          % (str(Nbar), str(P)))
            spec = _sep_product([w.astype(complex) for w in W], Nbar, hAM, ncomp)
            val = _icfftnc_real(spec, Nbar, ncomp).reshape(shp+Nbar)
            name = 'A_Ga_o{0}_P{1}'.format(order, np.array(P).max())
        return Tensor(name=name, val=val, N=Nbar, order=2, Y=self.Y, multype=21, Fourier=False, origin='c').shift()

    def get_shape_functions(self, N2, on_device=False):
        """exactly integrated characteristic functions of the inclusions on the grid N2 (materials.py:126-170)"""
        N2 = tuple(int(n) for n in np.array(N2).ravel())
        d = len(N2)
        n = int(np.prod(N2))
        incl = self.conf['inclusions']
        chars = dev.zeros((len(incl), n))
        Y = np.array(self.Y, dtype=float)
        for ii, kind in enumerate(incl):
            k = _kind_of(kind)
            if k in (_KIND['cube'], _KIND['ball'], _KIND['pyramid']):
                S = _shift_1d(N2, self.conf['positions'][ii], Y)
                if k == _KIND['ball']:
                    r = self.conf['params'][ii]/2
                    Wfull = np.zeros(N2) if r == 0 else get_weights_circ(r, np.array(N2), Y)
                    spec = _sep_product(S, N2, dev.upload(Wfull.astype(complex)).reshape((1,)+N2), 1)
                else:
                    par = np.array(self.conf['params'][ii], dtype=float)*np.ones(d)
                    W = _weights_1d(par if k == _KIND['cube'] else par/2., N2, Y, power=1 if k == _KIND['cube'] else 2)
                    spec = _sep_product([s*w for s, w in zip(S, W)], N2, None, 1)
                chars[ii] = _icfftnc_real(spec, N2, 1).reshape(n)
            elif k == _KIND['all']:
                chars[ii] = 1.
            else:                    # 'otherwise': the complement of everything listed before it
                chars[ii] = 1.
                for jj in range(len(incl)-1):
                    chars[ii] = ops.axpby(1., chars[ii], -1., chars[jj])
        if on_device:
            return chars
        return [dev.download(chars[i]).reshape(N2) for i in range(len(incl))]


# ----------------------------------------------------------------------------- device helpers
def _combine(vals, chars, n):
    """val[c] = sum_ii vals[ii][c] * chars[ii], accumulated in the order of the reference's `+=` loop"""
    ncomp = int(np.prod(vals[0].shape))
    nph = len(vals)
    coef = np.stack([np.asarray(v, dtype=float).ravel() for v in vals], axis=1)    # [ncomp][nphase]
    out = dev.empty((ncomp, n))
    L.check(dev.lib().fh_combine_phases(ncomp, nph, n, L.dblarr(coef.ravel()), dev.ptr(chars), dev.ptr(out)))
    return out


def _weights_1d(h, Nbar, Y, power):
    """per-axis factors of get_weights_con (power 1) / get_weights_lin (power 2) (materials.py:333-390): the product
    over the axes, divided by |Y| on the first one, is the reference's Wphi"""
    d = len(Nbar)
    ZN2l = Grid.get_ZNl(np.array(Nbar), fft_form='c')
    out = []
    for ii in range(d):
        f = h[ii]*np.sinc(h[ii]*ZN2l[ii]/Y[ii])**power
        out.append(f/np.prod(Y) if ii == 0 else f)
    return out


def _shift_1d(N, pos, Y):
    """per-axis factors of get_shift_inclusion (materials.py:318-330); the reference indexes them with the
    FFT-ordered frequencies of Grid.get_ZNl(N) while the weights are centred — reproduced as is"""
    ZN = Grid.get_ZNl(np.array(N))
    return [np.exp(-2*np.pi*1j*(pos[ii]*ZN[ii]/Y[ii])) for ii in range(len(N))]


def _sep_product(factors, N, data, batch):
    """out[b][k] = (data[b][k] or 1) * prod_a factors[a][k_a]   (complex, device)"""
    N = tuple(int(n) for n in N)
    f = dev.upload(np.concatenate([np.asarray(x, dtype=complex) for x in factors]))
    out = dev.empty((batch,)+N, complex_=True)
    L.check(dev.lib().fh_sep_product(len(N), L.i64arr(N), dev.ptr(f), int(batch),
                                     dev.ptr(data) if data is not None else None, dev.ptr(out)))
    return out


def _periodic_crop(X, P, M, batch):
    """decrease(tile(X, 2*ceil(M/2/P)+1), M) of materials.py:95-102 / plain decrease(X, M): the centre block of the
    periodically extended centred array, out[j] = X[(start + j) mod P]"""
    P, M = np.array(P), np.array(M)
    if np.all(P >= M):
        T = P
    else:
        T = P*(2*np.ceil(M.astype(np.float64)/2/P).astype(int)+1)
    start = np.fix((T-M+M % 2)/2).astype(int)           # trigpol.decrease (trigpol.py:191-214)
    out = dev.empty((batch,)+tuple(int(m) for m in M), complex_=True)
    L.check(dev.lib().fh_gather_periodic(len(P), L.i64arr(P), L.i64arr(M), L.i64arr(start), int(batch),
                                         dev.ptr(X), dev.ptr(out)))
    return out


def _cfftnc(x, N, batch, scale):
    """prod(N) * cfftnc(x, N) of tensors/fft.py:4-9 for `batch` real fields: ifftshift, plain DFT, fftshift"""
    xd = ops.roll(x, N, [-(n//2) for n in N], batch)
    X = ops.rfftn(xd, N, batch)
    return ops.spec_remap(X, N, 'r', N, 'c', batch, 1.)


def _icfftnc_real(X, N, batch):
    """real(icfftnc(X, N)) of tensors/fft.py:11-16 for `batch` centred spectra"""
    H = ops.spec_remap(X, N, 'c', N, 'r', batch, 1., flags=2)
    x = ops.irfftn(H, N, batch, 1.)
    return ops.roll(x, N, [n//2 for n in N], batch)


# ----------------------------------------------------------------------------- the reference's weight functions (host)
def get_shift_inclusion(N, h, Y):
    N = np.array(N, dtype=int)
    S = _shift_1d(N, h, np.array(Y, dtype=float))
    out = np.ones(tuple(N), dtype=np.complex128)
    for ii, s in enumerate(S):
        shape = np.ones(N.size, dtype=int)
        shape[ii] = N[ii]
        out = out*np.reshape(s, shape)
    return out


def _outer(factors, N):
    out = np.ones(tuple(int(n) for n in N))
    for ii, f in enumerate(factors):
        shape = np.ones(len(N), dtype=int)
        shape[ii] = N[ii]
        out = out*np.reshape(f, shape)
    return out


def get_weights_con(h, Nbar, Y):
    """integral weights of a constant rectangular inclusion of size h (materials.py:333-360)"""
    return _outer(_weights_1d(h, Nbar, Y, 1), Nbar)


def get_weights_lin(h, Nbar, Y):
    """integral weights of a bilinear inclusion with half-support h (materials.py:363-390)"""
    return _outer(_weights_1d(h, Nbar, Y, 2), Nbar)


def get_weights_circ(r, Nbar, Y):
    """integral weights of a disc of radius r (materials.py:393-425; the 2-D disc transform in any dimension,
    SURVEY D.9)"""
    import scipy.special as sp
    d = np.size(Y)
    ZN2l = Grid.get_ZNl(Nbar, fft_form='c')
    circ = 0
    for m in range(d):
        shape = np.ones(d, dtype=int)
        shape[m] = Nbar[m]
        circ = circ+np.reshape((ZN2l[m]/Y[m])**2, shape)
    circ = circ*np.ones(tuple(int(n) for n in Nbar))
    circ = circ**0.5
    ind = mean_index(Nbar, fft_form='c')
    circ[ind] = 1.
    Wphi = r**2*sp.jn(1, 2*np.pi*circ*r)/(circ*r)
    Wphi[ind] = np.pi*r**2
    return Wphi/np.prod(Y)
