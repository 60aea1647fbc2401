"""Operators on Tensors — drop-in for ffthompy/tensors/operators.py (DFT, Operator, grad, div,
laplace, symgrad, potential, grad_tensor, div_tensor, outer).

`Operator.__call__` recognises the solve-loop pattern
    Operator([[ Operator([[FiN, G^, FN]]), A ]])          (applications.py:33-58)
and runs it as one fused device pipeline (ffthompy_b200/fused.py); any other composition is
evaluated factor by factor, each factor being a device kernel.
"""
import itertools
from copy import copy

import numpy as np

from .. import ops
from ..trigpol import Grid, fft_form_default
from .objects import Tensor, TensorFuns


class DFT(TensorFuns):
    """(inverse) Discrete Fourier Transform (tensors/operators.py:14-106).

    fft_form: 0 = numpy.fft.fftn order, normalised by 1/prod(N); 'c' = the same, centred;
    'r' = real-input transform (half spectrum, un-normalised forward)."""

    def __init__(self, inverse=False, N=None, fft_form=fft_form_default, **kwargs):
        self.__dict__.update(kwargs)
        if 'name' not in list(kwargs.keys()):
            if inverse:
                self.name = 'iDFT'
            else:
                self.name = 'DFT'
        self.N = np.array(N, dtype=np.int32)
        self.inverse = inverse
        self._set_fft(fft_form)

    def __mul__(self, x):
        return self.__call__(x)

    def __call__(self, x):
        if isinstance(x, Tensor):
            assert(x.Fourier == self.inverse)
            assert(np.all(np.array(x.N) == self.N))
            assert(x.fft_form == self.fft_form)
            if self.inverse:
                return x.copy(name='iF({0})'.format(x.name[:10]), val=x._ifft_val(), Fourier=not x.Fourier)
            else:
                return x.copy(name='F({0})'.format(x.name[:10]), val=x._fft_val(), Fourier=not x.Fourier)
        elif (isinstance(x, Operator) or isinstance(x, DFT)):
            return Operator(mat=[[self, x]])
        else:
            raise ValueError('DFT.__call__')

    def matrix(self, shape=None):
        """Dense matrix of the (i)DFT — a test utility (tensors/operators.py:64-95); assembled on the
        host exactly as in the reference, it touches no field data."""
        N = self.N
        prodN = np.prod(N)
        if shape is not None:
            dim = np.prod(np.array(shape))
        else:
            raise ValueError('Missing shape of the DFT.')
        proddN = int(dim*prodN)
        ZN_input = Grid.get_ZNl(N, fft_form=0)
        ZN_output = Grid.get_ZNl(N, fft_form='c')
        if self.inverse:
            DFTcoef = lambda k, l, N: np.exp(2*np.pi*1j*np.sum(k*l/N))  # noqa: E731
        else:
            DFTcoef = lambda k, l, N: np.exp(-2*np.pi*1j*np.sum(k*l/N))/np.prod(N)  # noqa: E731
        DTM = np.zeros([self.pN(), self.pN()], dtype=np.complex128)
        for ii, kk in enumerate(itertools.product(*tuple(ZN_output))):
            for jj, ll in enumerate(itertools.product(*tuple(ZN_input))):
                DTM[ii, jj] = DFTcoef(np.array(kk, dtype=float), np.array(ll), N)
        DTMd = np.zeros([proddN, proddN], dtype=np.complex128)
        for ii in range(int(dim)):
            DTMd[prodN*ii:prodN*(ii+1), prodN*ii:prodN*(ii+1)] = DTM
        return np.asmatrix(DTMd)

    def __repr__(self):
        keys = ['name', 'inverse', 'fft_form', 'N']
        return self._repr(keys)

    def transpose(self):
        kwargs = copy(self.__dict__)
        kwargs.update(dict(inverse=not self.inverse))
        for k in ('N_fft', 'fft_coef'):
            kwargs.pop(k, None)
        return DFT(**kwargs)


class Operator():
    """Linear operator composed of tensors / transforms / operators: a sum of products applied
    right to left (tensors/operators.py:108-225)."""

    def __init__(self, name='Operator', mat_rev=None, mat=None, operand=None):
        self.name = name
        if mat_rev is not None:
            self.mat_rev = mat_rev
        elif mat is not None:
            self.mat_rev = []
            for summand in mat:
                no_oper = len(summand)
                summand_rev = []
                for m in np.arange(no_oper):
                    summand_rev.append(summand[no_oper-1-m])
                self.mat_rev.append(summand_rev)
        self.no_summands = len(self.mat_rev)
        self._fused = None
        if operand is not None:
            self.define_operand(operand)

    def fused(self):
        """The fused G·A pipeline if this operator has the solve-loop shape, else None."""
        from .. import fused
        self._fused = fused.get_fused(self, self._fused)
        return self._fused

    def __call__(self, x):
        f = self.fused() if isinstance(x, Tensor) else None
        if f is not None and f.accepts(x):
            res = x.copy(val=f.apply(x._dev()))
        else:
            res = None
            for summand in self.mat_rev:
                prod = x
                for matrix in summand:
                    prod = matrix(prod)
                res = prod if res is None else prod+res
            if res is x:
                res = x.copy()
        res.name = '{0}({1})'.format(self.name[:6], x.name[:10])
        return res

    def __repr__(self):
        s = 'Class : {0}\n    name : {1}\n    expression : '.format(self.__class__.__name__, self.name)
        flag_sum = False
        no_sum = len(self.mat_rev)
        for isum in np.arange(no_sum):
            if flag_sum:
                s += ' + '
            no_oper = len(self.mat_rev[isum])
            flag_mul = False
            for m in np.arange(no_oper):
                matrix = self.mat_rev[isum][no_oper-1-m]
                if flag_mul:
                    s += '*'
                s += matrix.name
                flag_mul = True
            flag_sum = True
        return s

    def define_operand(self, X):
        """tensors/operators.py:165-185"""
        if isinstance(X, Tensor):
            Y = self(X)
            self.matshape = (int(np.prod(Y._vshape())), int(np.prod(X._vshape())))
            self.X_reshape = X._vshape()
            self.X_order = X.order
            self.X_N = X.N
            self.Y_reshape = Y._vshape()
            self.Y_order = Y.order
        else:
            print('LinOper : This operand is not implemented!')

    def matvec(self, x):
        """__call__ for an operand recast into a one-dimensional numpy vector (SciPy bridge)."""
        X = Tensor(val=self.revec(x), order=self.X_order, N=self.X_N)
        AX = self.__call__(X)
        return AX.vec()

    def vec(self, X):
        return np.reshape(X, self.shape[1])

    def revec(self, x):
        return np.reshape(np.asarray(x), self.Y_reshape)

    def transpose(self):
        """Transpose (adjoint) of the linear operator."""
        mat = []
        for m in np.arange(self.no_summands):
            summand = []
            for n in np.arange(len(self.mat_rev[m])):
                summand.append(self.mat_rev[m][n].transpose())
            mat.append(summand)
        name = '({0}).T'.format(self.name[:10])
        return Operator(name=name, mat=mat)

    def __eq__(self, other):
        return self is other

    __hash__ = object.__hash__


def _fourier_of(X):
    return X if X.Fourier else DFT(N=X.N, fft_form=X.fft_form)(X)


def grad(X):
    """Gradient by Fourier multipliers 2*pi*i*xi (tensors/operators.py:227-259)."""
    if X.shape == (1,):
        shape = (X.dim,)
    else:
        shape = tuple(X.shape)+(X.dim,)
    FX = _fourier_of(X)
    nf = FX._ngrid
    val = ops.grad(ops.convert(FX._dev(), True), X.N, X.Y, X.fft_form, FX._ncomp, nf)
    gX = Tensor(name='grad({0})'.format(X.name[:10]), val=val.reshape(shape+tuple(FX.N_fft)), order=len(shape),
                N=X.N, Y=X.Y, Fourier=True, fft_form=X.fft_form)
    if not X.Fourier:
        iF = DFT(N=X.N, inverse=True, fft_form=gX.fft_form)
        gX = iF(gX)
    gX.name = 'grad({0})'.format(X.name[:10])
    return gX


def div(X):
    """Divergence (tensors/operators.py:261-288)."""
    if X.shape == (1,):
        shape = ()
    else:
        shape = tuple(X.shape[:-1])
    assert(X.shape[-1] == X.dim)
    assert(X.order == 1)
    FX = _fourier_of(X)
    nf = FX._ngrid
    val = ops.div(ops.convert(FX._dev(), True), X.N, X.Y, FX.fft_form, 1, nf)
    dX = Tensor(val=val.reshape(shape+tuple(FX.N_fft)), order=len(shape), N=X.N, Y=X.Y, Fourier=True,
                fft_form=X.fft_form)
    if not X.Fourier:
        iF = DFT(N=X.N, inverse=True, fft_form=dX.fft_form)
        dX = iF(dX)
    dX.name = 'div({0})'.format(X.name[:10])
    return dX


def laplace(X):
    return div(grad(X))


def symgrad(X):
    gX = grad(X)
    return 0.5*(gX+gX.transpose())


def _potential_dev(Fval, N, Y, fft_form, ncomp, nf):
    """potential_scalar for `ncomp` vector fields stored as (ncomp, dim) + N_fft
    (tensors/operators.py:296-309): u(k) = g_a(k)/(2 pi i xi_a), a = first axis with k_a != 0."""
    return ops.potential(Fval, N, Y, fft_form, ncomp, nf)


def potential(X, small_strain=False):
    """Potential of a curl-free / compatible field (tensors/operators.py:311-354)."""
    FX = _fourier_of(X)
    nf = FX._ngrid
    Nf = tuple(FX.N_fft)
    if X.order == 1:
        assert(X.dim == X.shape[0])
        val = _potential_dev(ops.convert(FX._dev(), True), X.N, X.Y, FX.fft_form, 1, nf)
        iX = Tensor(name='potential({0})'.format(X.name[:10]), val=val.reshape((1,)+Nf), order=1, N=X.N, Y=X.Y,
                    Fourier=True, fft_form=FX.fft_form)
    elif X.order == 2:
        assert(X.dim == X.shape[0])
        assert(X.dim == X.shape[1])
        if not small_strain:
            val = _potential_dev(ops.convert(FX._dev(), True), X.N, X.Y, FX.fft_form, X.dim, nf)
            iX = Tensor(name='potential({0})'.format(X.name[:10]), val=val.reshape((X.dim,)+Nf), order=1, N=X.N,
                        Y=X.Y, Fourier=True, fft_form=FX.fft_form)
        else:
            assert((X-X.transpose()).norm() < 1e-14)  # symmetricity
            d = X.dim
            grad_ep = grad(FX)  # gradient of strain, shape (d, d, d) + N_fft, index (i, j, k) = d_k eps_ij
            idx = np.arange(d**3).reshape((d, d, d))
            # gomeg_ijk = d_j eps_ik - d_i eps_jk   ('ikj->ijk' minus 'jki->ijk', operators.py:342)
            p1 = list(np.einsum('ikj->ijk', idx).ravel())
            p2 = list(np.einsum('jki->ijk', idx).ravel())
            g = grad_ep._dev()
            gom = ops.axpby(1., ops.gather_comps(g, p1, d**3), -1., ops.gather_comps(g, p2, d**3))
            # potential of every (i, j) vector field gomeg_ij.
            omeg = _potential_dev(gom, X.N, X.Y, FX.fft_form, d*d, nf)
            gradu = FX.copy(val=ops.axpby(1., ops.convert(FX._dev(), True).reshape(-1), 1., omeg).reshape((d, d)+Nf))
            iX = potential(gradu, small_strain=False)
    else:
        raise NotImplementedError()
    if X.Fourier:
        return iX
    else:
        iF = DFT(N=X.N, inverse=True, fft_form=FX.fft_form)
        return iF(iX)


def matrix2tensor(M):
    return Tensor(name=M.name, val=M.val, order=2, multype=21, Fourier=M.Fourier, fft_form=fft_form_default)


def vector2tensor(V):
    return Tensor(name=V.name, val=V.val, order=1, Fourier=V.Fourier)


def grad_div_tensor(N, Y=None, grad=True, div=True, fft_form=fft_form_default):
    if grad and div:
        return grad_tensor(N, Y, fft_form=fft_form), div_tensor(N, Y, fft_form=fft_form)
    elif grad:
        return grad_tensor(N, Y, fft_form=fft_form)
    elif div:
        return div_tensor(N, Y, fft_form=fft_form)


def grad_tensor(N, Y=None, fft_form=fft_form_default):
    """Materialised gradient multiplier 2 pi i xi (tensors/operators.py:363-382): built on the device by
    applying the gradient kernel to a field of ones."""
    if Y is None:
        Y = np.ones_like(N)
    N = np.array(N, dtype=int)
    one = Tensor(name='1', shape=(1,), N=N, Y=Y, Fourier=True, fft_form=fft_form)
    one.val = np.ones(one._vshape(), dtype=complex)
    nf = one._ngrid
    val = ops.grad(one._dev(), tuple(N), np.array(Y, dtype=float), fft_form, 1, nf)
    return Tensor(name='hgrad', val=val.reshape((N.size,)+tuple(one.N_fft)), order=1, N=N, multype='grad',
                  Fourier=True, fft_form=fft_form)


def div_tensor(N, Y=None, fft_form=fft_form_default):
    if Y is None:
        Y = np.ones_like(N)
    hGrad = grad_tensor(N, Y=Y, fft_form=fft_form)
    hGrad.multype = 'div'
    return hGrad


def outer(X, Y):
    """Point-wise outer product (tensors/operators.py:395-402)."""
    assert(np.allclose(X.N, Y.N))
    n = X._ngrid
    cx, cy = X._ncomp, Y._ncomp
    val = ops.hadamard(X._dev(), Y._dev(), n, cx*cy, cy, cx, 1, cy)
    return X.copy(name='outer({},{})'.format(X.name, Y.name), order=X.order+Y.order,
                  val=val.reshape(tuple(X.shape)+tuple(Y.shape)+X._vshape()[X.order:]))
