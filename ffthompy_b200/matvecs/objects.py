"""Legacy trigonometric-polynomial classes — drop-in for the part of ffthompy/matvecs/objects.py
that still works in the reference: `VecTri`, `Matrix`, the doubly centred,
normalised `DFT` and `LinOper`.  Same device kernels as ffthompy_b200.tensors; only the
conventions differ (centred real space AND centred Fourier space, forward transform divided
by prod(N)).  The Multi*/Scipy* wrappers and ShiftMatrix of the reference are unused by its
applications and tests and are not provided.
"""
import itertools
from warnings import warn

import numpy as np

from .. import device as dev
from .. import ops
from ..trigpol import Grid, mean_index


class Scalar():
    """Scalar value that multiplies VecTri or Matrix (matvecs/objects.py:60-80)."""

    def __init__(self, val=None, name='c'):
        self.val = 1. if val is None else val
        self.name = name

    def __call__(self, x):
        return self*x

    def __mul__(self, x):
        return x.__rmul__(self)

    def __repr__(self):
        return "Class : {0}\n    val = {1}".format(self.__class__.__name__, self.val)

    def transpose(self):
        return self


def get_name(x_name, oper, y_name):
    name = x_name+oper+y_name
    if len(name) > 20:
        name = 'oper({})'.format(oper)
    return name


def _is_dev(v):
    return type(v).__module__.startswith('torch') and hasattr(v, 'data_ptr')


class FieldFun():
    """Storage protocol shared with tensors.Tensor: `.val` is a NumPy view of device data; reading
    it makes the host copy authoritative."""
    _h = None
    _d = None

    @property
    def val(self):
        if self._h is None:
            self._h = dev.download(self._d)
        self._d = None
        return self._h

    @val.setter
    def val(self, v):
        if _is_dev(v):
            self._d, self._h = v, None
        else:
            self._h, self._d = np.asarray(v), None

    def _dev(self):
        if self._d is None:
            self._d = dev.upload(self._h)
        return self._d

    def _vshape(self):
        return tuple((self._d if self._d is not None else self._h).shape)

    def dN(self):
        return np.hstack([self.d, self.N])

    def ddN(self, M=None):
        if M is None:
            M = self.N
        return np.hstack([self.d, self.d, M])

    def pN(self):
        return np.prod(self.N)

    def pdN(self):
        return np.prod(self.dN())

    def mean_index(self):
        return mean_index(self.N, fft_form='c')

    def __getitem__(self, i):
        return self.val[i]

    def __repr__(self, full=False):
        ss = "Class : %s\n    name : %s\n" % (self.__class__.__name__, self.name)
        ss += '    Fourier = %s \n' % (self.Fourier)
        ss += '    dimension d = %g \n' % (self.d)
        ss += '    size N = %s \n' % str(self.N)
        ss += '    val.shape  = %s \n' % str(self._vshape())
        ss += '    norm = %s\n' % str(self.norm())
        ss += '    mean = %s\n' % str(self.mean())
        if full:
            ss += 'val = \n'+str(self.val)
        return ss


class VecTri(FieldFun, Grid):
    """Vector-valued trigonometric polynomial given by grid values or (centred) Fourier
    coefficients (matvecs/objects.py:83-409)."""

    def __init__(self, name='?', N=None, d=None, Fourier=False, valtype=None, **kwargs):
        self.Fourier = Fourier
        warn("The class {} will be depreciated. Use ffthompy.tensors.".format(self.__class__.__name__))
        if 'val' in kwargs:
            self.val = kwargs['val']
            self.N = np.array(self._vshape()[1:])
            self.d = self._vshape()[0]
        else:
            if N is None:
                raise ValueError("Parameter N is required!")
            self.N = np.array(N, dtype=np.int32)
            self.d = self.N.size if d is None else d
            if 'macroval' in kwargs:
                name = 'macroval' if name is None else name
                self.d = np.size(kwargs['macroval'])
                val = np.zeros(self.dN())
                for m in np.arange(self.d):
                    val[m] = kwargs['macroval'][m]
                self.val = val
            elif valtype == 'ones':
                self.val = np.ones(self.dN())
            elif valtype in ['random', 'rand']:
                self.val = np.random.random(self.dN())
            else:
                self.val = dev.zeros(tuple(self.dN()), complex_=bool(self.Fourier))
        if 'Y' in kwargs:
            self.Y = np.array(kwargs['Y'])
        self.name = name if name is not None else '?'
        self.valshape = self._vshape()
        self.size = int(np.prod(self._vshape()))

    def _like(self, val, name=None, Fourier=None):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            return VecTri(name=self.name if name is None else name, val=val,
                          Fourier=self.Fourier if Fourier is None else Fourier)

    def __mul__(self, x):
        if isinstance(x, VecTri):
            a, b = ops.promote(self._dev(), x._dev())
            scal = ops.dot(a, b)
            if not self.Fourier:
                scal = scal/np.prod(self.N)
            return scal
        elif np.size(x) == 1 and not hasattr(x, 'val'):
            self.val = ops.axpby(float(np.asarray(x).ravel()[0]), self._dev())  # in place, as the reference
            return self
        raise ValueError("The shape of vectors are not appropriate.")

    def __rmul__(self, x):
        if isinstance(x, Scalar):
            return self._like(ops.axpby(float(x.val), self._dev()), name=get_name('c', '*', self.name))
        elif np.size(x) == 1:
            return self._like(ops.axpby(float(np.asarray(x).ravel()[0]), self._dev()),
                              name=get_name('c', '*', self.name))
        raise ValueError()

    def __add__(self, x):
        if isinstance(x, VecTri):
            if self.Fourier != x.Fourier:
                raise ValueError("Mismatch in Fourier/shape coefficients!")
            a, b = ops.promote(self._dev(), x._dev())
            return self._like(ops.axpby(1., a, 1., b), name=get_name(self.name, '+', x.name))
        if np.size(x) == 1:
            return self._like(ops.add_scalar(self._dev(), float(np.asarray(x).ravel()[0])))
        full = np.broadcast_to(np.asarray(x), self._vshape())
        a, b = ops.promote(self._dev(), dev.upload(full))
        return self._like(ops.axpby(1., a, 1., b))

    def __radd__(self, x):
        return self+x

    def __neg__(self):
        return self._like(ops.axpby(-1., self._dev()), name='-'+self.name)

    def __sub__(self, x):
        return self.__add__(-x)

    def norm(self, ntype='L2'):
        if ntype in ['L2', 2]:
            return (self*self)**0.5
        elif ntype == 1:
            return ops.asum(self._dev())
        elif ntype == 'inf':
            return ops.amax(self._dev())
        raise NotImplementedError("The norm (%s) of VecTri is not implemented!" % ntype)

    def mean(self):
        mean = np.zeros(self.d)
        d = self._dev()
        n = int(np.prod(self.N))
        if self.Fourier:
            off = int(np.ravel_multi_index(mean_index(self.N, fft_form='c'), tuple(self.N)))
            e = 2 if dev.is_complex(d) else 1
            for di in range(self.d):
                mean[di] = ops.peek(d, (di*n+off)*e, 1)[0]
        else:
            mean[:] = ops.sum_comp(d, self.d)/n
        return mean

    def __call__(self):
        return self.val

    def vec(self):
        return np.matrix(self.val.ravel()).transpose()

    def __eq__(self, x):
        if isinstance(x, VecTri):
            return (self-x).norm()
        elif np.shape(x) == self._vshape():
            return np.linalg.norm(self.val-x)
        return False

    __hash__ = object.__hash__

    def project(self, M):
        """matvecs/objects.py:258-289: centred zero padding / truncation of the Fourier coefficients
        (trigpol.enlarge/decrease semantics — no Nyquist splitting, no rescaling)."""
        M = np.array(M, dtype=int)
        if np.allclose(self.N, M):
            return self
        if not (np.all(np.greater(M, self.N)) or np.all(np.less(M, self.N))):
            raise NotImplementedError()
        F = self if self.Fourier else DFT(inverse=False, N=self.N)(self)
        out = ops.spec_remap(F._dev(), tuple(self.N), 'c', tuple(M), 'c', self.d, 1.0, flags=1)
        R = self._like(out.reshape((self.d,)+tuple(M)), Fourier=True)
        if not self.Fourier:
            R = DFT(inverse=True, N=M)(R)
        R.name = self.name
        return R

    def transpose(self):
        return self

    @property
    def T(self):
        return self

    def fourier_transform(self):
        return DFT(inverse=self.Fourier, N=self.N)(self)

    def copy(self, name='copied'):
        return self._like(ops.clone(self._dev()), name=name)

    def zeros_like(self, name='zeros like '):
        return self._like(dev.zeros(self._vshape(), complex_=dev.is_complex(self._dev())), name=name+self.name)

    def empty_like(self, name='zeros like '):
        return self.zeros_like(name)


class Matrix(FieldFun):
    """(d, d) matrix field: material coefficients or a projection kernel in centred Fourier space
    (matvecs/objects.py:417-635)."""

    def __init__(self, name='?', Fourier=False, valtype='val', **kwargs):
        self.Fourier = Fourier
        self.name = name
        self.valtype = valtype
        val = kwargs.pop('val', None)
        self.__dict__.update(kwargs)
        if valtype in ['val']:
            self.val = val if _is_dev(val) else np.array(val)
            self.N = np.array(self._vshape()[2:])
            self.d = self._vshape()[0]
            if self._vshape()[1] != self.d:
                raise ValueError("Improper dimension of values %s." % str(self._vshape()))
        else:
            if not hasattr(self, 'N'):
                raise ValueError("Argument 'N' has to be defined!")
            if not hasattr(self, 'd'):
                self.d = np.size(self.N)
            dtype = np.complex128 if self.Fourier else np.float64
            if valtype in ['Id', 'id', 'identity']:
                v = np.zeros(self.ddN(), dtype=dtype)
                for m in np.arange(self.d):
                    v[m][m] = 1.
                self.val = v
            elif valtype in ['random']:
                self.val = np.random.random(self.ddN())
            elif valtype in ['homog']:
                v = np.zeros(self.ddN(), dtype=dtype)
                for m in np.arange(self.d):
                    for n in np.arange(self.d):
                        v[m, n] = np.array(val[m, n])
                self.val = v

    def __mul__(self, x):
        n = int(np.prod(self.N))
        if isinstance(x, VecTri):  # Matrix by VecTri multiplication
            y = ops.mul21(self._dev(), x._dev(), self.d, n, 1)
            return x._like(y.reshape(x._vshape()), name=get_name(self.name, '*', x.name), Fourier=x.Fourier)
        elif isinstance(x, Matrix):  # Matrix by Matrix multiplication
            y = ops.mul21(self._dev(), x._dev(), self.d, n, self.d)
            return Matrix(name=get_name(self.name, '*', x.name), val=y.reshape(self._vshape()))
        elif isinstance(x, LinOper) or isinstance(x, DFT):
            return LinOper(name=get_name(self.name, '*', x.name), mat=[[self, x]])
        elif isinstance(x, Scalar):
            return Matrix(name=get_name(self.name, '*', 'c'), val=ops.axpby(float(x.val), self._dev()),
                          Fourier=self.Fourier)
        elif np.size(x) == 1:  # Matrix by constant multiplication
            return Matrix(name=get_name(self.name, '*', 'c'),
                          val=ops.axpby(float(np.asarray(x).ravel()[0]), self._dev()), Fourier=self.Fourier)
        raise ValueError('Matrix.__mul__: unsupported operand')

    def __rmul__(self, x):
        return self*x

    def __call__(self, x):
        return self*x

    def norm(self):
        d = self._dev()
        return ops.dot(d, d)**0.5

    def mean(self):
        res = np.zeros([self.d, self.d])
        d = self._dev()
        n = int(np.prod(self.N))
        if self.Fourier:
            off = int(np.ravel_multi_index(mean_index(self.N, fft_form='c'), tuple(self.N)))
            e = 2 if dev.is_complex(d) else 1
            for c in range(self.d*self.d):
                res[c//self.d, c % self.d] = ops.peek(d, (c*n+off)*e, 1)[0]
        else:
            res[:] = (ops.sum_comp(d, self.d*self.d)/n).reshape(self.d, self.d)
        return res

    def __add__(self, x):
        if isinstance(x, Matrix):
            a, b = ops.promote(self._dev(), x._dev())
            return Matrix(name=get_name(self.name, '+', x.name), val=ops.axpby(1., a, 1., b), Fourier=self.Fourier)
        return Matrix(val=ops.add_scalar(self._dev(), float(x)), Fourier=self.Fourier)

    def __neg__(self):
        return Matrix(val=ops.axpby(-1., self._dev()))

    def __sub__(self, x):
        if isinstance(x, Matrix):
            return -x+self
        return 'this type of operation is not supported'

    def T(self):
        return self.transpose()

    def transpose(self):
        perm = [j*self.d+i for i in range(self.d) for j in range(self.d)]
        out = ops.gather_comps(self._dev(), perm, self.d*self.d)
        return Matrix(name=self.name, val=out.reshape(self._vshape()), Fourier=self.Fourier)

    def inv(self):
        if self.Fourier is False:
            return Matrix(name='inv(%s)' % (self.name),
                          val=ops.inv_dxd(self._dev(), self.d, int(np.prod(self.N))), Fourier=False)
        raise NotImplementedError("The inverse for Fourier coefficients!")

    def __eq__(self, x):
        if isinstance(x, Matrix) and self._vshape() == x._vshape():
            return (self-x).norm()
        return False

    __hash__ = object.__hash__

    def enlarge(self, M):
        """Centred zero padding of a Fourier-space kernel (matvecs/objects.py:599-609, odd N)."""
        M = tuple(int(m) for m in np.array(M).ravel())
        if not self.Fourier:
            raise NotImplementedError('Matrix.enlarge of grid values')
        was_real = not dev.is_complex(self._dev())
        out = ops.spec_remap(self._dev(), tuple(self.N), 'c', M, 'c', self.d*self.d, 1.0, flags=1)
        if was_real:
            out = ops.convert(out, False)
        return Matrix(name=self.name, val=out.reshape((self.d, self.d)+M), Fourier=True)


class DFT(FieldFun):
    """Doubly centred (inverse) DFT, forward normalised by prod(N)
    (matvecs/objects.py:690-800): F x = fftshift(fftn(ifftshift(x)))/prod(N)."""

    def __init__(self, inverse=False, N=None, normalized=True, **kwargs):
        self.__dict__.update(kwargs)
        if 'name' not in list(kwargs.keys()):
            self.name = 'iDFT' if inverse else 'DFT'
        self.N = np.array(N, dtype=np.int32)
        self.inverse = inverse
        self.norm_coef = np.prod(self.N) if normalized else 1.

    def __mul__(self, x):
        return self.__call__(x)

    def __call__(self, x):
        if isinstance(x, VecTri):
            if not self.inverse:
                return x._like(self.fftnc_dev(x._dev(), self.N), name=get_name('F', '*', x.name),
                               Fourier=not x.Fourier)
            return x._like(self.ifftnc_dev(x._dev(), self.N), name=get_name('Fi', '*', x.name),
                           Fourier=not x.Fourier)
        elif isinstance(x, (LinOper, Matrix, DFT)):
            return LinOper(mat=[[self, x]])
        raise ValueError('DFT.__call__: operand must be a VecTri or an operator')

    @staticmethod
    def fftnc_dev(x, N):
        N = tuple(int(n) for n in N)
        batch = int(x.numel())//int(np.prod(N))
        xs = ops.roll(ops.convert(x, False), N, [-(n//2) for n in N], batch)  # ifftshift
        X = ops.rfftn(xs, N, batch)
        out = ops.spec_remap(X, N, 'r', N, 'c', batch, 1./float(np.prod(N)))
        return out.reshape(tuple(x.shape[:x.dim()-len(N)])+N)

    @staticmethod
    def ifftnc_dev(X, N):
        N = tuple(int(n) for n in N)
        batch = int(X.numel())//int(np.prod(N))
        H = ops.spec_remap(ops.convert(X, True), N, 'c', N, 'r', batch, 1., flags=2)
        x = ops.irfftn(H, N, batch, 1.)
        out = ops.roll(x, N, [n//2 for n in N], batch)  # fftshift
        return out.reshape(tuple(X.shape[:X.dim()-len(N)])+N)

    @staticmethod
    def fftnc(x, N):
        """centred n-dimensional FFT of a host array (matvecs/objects.py:784-790), on the device"""
        return dev.download(DFT.fftnc_dev(dev.upload(np.asarray(x, dtype=float)), N))

    @staticmethod
    def ifftnc(Fx, N):
        return dev.download(DFT.ifftnc_dev(dev.upload(np.asarray(Fx, dtype=complex)), N))

    def matrix(self):
        """dense (i)DFT matrix, test utility assembled on the host (matvecs/objects.py:749-772)"""
        N = self.N
        prodN = int(np.prod(N))
        proddN = self.d*prodN
        ZNl = Grid.get_ZNl(N, fft_form='c')
        if self.inverse:
            DFTcoef = lambda k, l, N: np.exp(2*np.pi*1j*np.sum(k*l/N))  # noqa: E731
        else:
            DFTcoef = lambda k, l, N: np.exp(-2*np.pi*1j*np.sum(k*l/N))/np.prod(N)  # noqa: E731
        DTM = np.zeros([prodN, prodN], dtype=np.complex128)
        for ii, kk in enumerate(itertools.product(*tuple(ZNl))):
            for jj, ll in enumerate(itertools.product(*tuple(ZNl))):
                DTM[ii, jj] = DFTcoef(np.array(kk, dtype=float), np.array(ll), N)
        DTMd = np.zeros([proddN, proddN], dtype=np.complex128)
        for ii in range(self.d):
            DTMd[prodN*ii:prodN*(ii+1), prodN*ii:prodN*(ii+1)] = DTM
        return np.asmatrix(DTMd)

    def __repr__(self):
        ss = "Class : %s\n" % (self.__class__.__name__,)
        ss += '    name : %s\n' % self.name
        ss += '    inverse = %s\n' % self.inverse
        ss += '    size N = %s\n' % str(self.N)
        return ss

    def transpose(self):
        return DFT(name=self.name+'^T', inverse=not(self.inverse), N=self.N)


class LinOper():
    """Sum of products of operators applied right to left (matvecs/objects.py:802-935)."""

    def __init__(self, name='LinOper', dtype=None, X=None, **kwargs):
        self.name = name
        if 'mat_rev' in list(kwargs.keys()):
            self.mat_rev = kwargs['mat_rev']
        elif 'mat' in list(kwargs.keys()):
            self.mat_rev = [list(reversed(summand)) for summand in kwargs['mat']]
        self.no_summands = len(self.mat_rev)
        if X is not None:
            self.define_operand(X)
        self.dtype = np.float64 if dtype is None else dtype

    def __mul__(self, x):
        if isinstance(x, VecTri):
            return self(x)
        elif isinstance(x, (Matrix, LinOper, DFT)):
            return LinOper(name=self.name+'*'+x.name, mat=[[self, x]])

    def __add__(self, x):
        if isinstance(x, (Matrix, LinOper)):
            return LinOper(name=self.name+'+'+x.name, mat=[[self], [x]])
        return 'This operation is not supported!'

    def __call__(self, x):
        res = None
        for summand in self.mat_rev:
            prod = x
            for matrix in summand:
                prod = matrix(prod)
            res = prod if res is None else prod+res
        return res

    def __repr__(self):
        s = 'Class : %s\nname : %s\nexpression : ' % (self.__class__.__name__, self.name)
        s += ' + '.join('*'.join(m.name for m in reversed(summand)) for summand in self.mat_rev)
        return s

    def define_operand(self, X):
        if isinstance(X, VecTri):
            Y = self(X)
            self.matshape = (Y.size, X.size)
            self.X_reshape = X._vshape()
            self.Y_reshape = Y._vshape()
        else:
            print('LinOper : This operand is not implemented!')

    def matvec(self, x):
        X = VecTri(val=self.revec(x))
        return self.__call__(X).vec()

    def revec(self, x):
        return np.reshape(np.asarray(x), self.Y_reshape)

    def transpose(self):
        mat = [[m.transpose() for m in summand] for summand in self.mat_rev]
        return LinOper(name='(%s)^T' % self.name, mat=mat)


class MultiVector():
    """Block vector of mixed formulations: an ordered list of sub-vectors (VecTri) that adds, negates, scales and
    contracts block by block (matvecs/objects.py:938-1053).  The blocks keep their device storage; only `vec()`
    assembles a host array."""

    def __init__(self, name='MultiVector', val=None):
        self.name = name
        self.val = list(val)
        self.dim = len(self.val)
        self._iter = np.arange(self.dim)
        self.ltype = [type(b).__name__ for b in self.val]
        self.lshape, self.ldtype = [], []
        self.lsize = np.zeros(self.dim, dtype=np.int64)
        for m, b in enumerate(self.val):
            if isinstance(b, VecTri):
                self.lshape.append(b.valshape)
                self.ldtype.append(getattr(b, 'dtype', np.float64))
                self.lsize[m] = b.size
        self.size = int(np.sum(self.lsize))

    def _map(self, fun, name=None):
        return MultiVector(name=self.name if name is None else name, val=[fun(b) for b in self.val])

    def __mul__(self, x):
        if isinstance(x, MultiVector):          # block-wise scalar product, summed
            return sum((a*b for a, b in zip(self.val[1:], x.val[1:])), self.val[0]*x.val[0])
        if isinstance(x, Scalar):
            return self._map(lambda b: x.val*b)
        if np.size(x) == 1:
            c = float(np.asarray(x).ravel()[0])
            return self._map(lambda b: c*b)
        raise NotImplementedError()

    __rmul__ = __mul__
    __call__ = __mul__

    def __add__(self, x):
        return MultiVector(val=[a+b for a, b in zip(self.val, x.val)])

    def __neg__(self):
        return self._map(lambda b: -b)

    def __sub__(self, x):
        return self+(-x)

    def __getitem__(self, m):
        return self.val[m]

    def vec(self):
        return np.vstack([np.asarray(b.vec()) for b in self.val])

    def __eq__(self, x):
        same = [type(a).__name__ == type(b).__name__ for a, b in zip(self.val, x.val)]
        vals = [(a == b) if t else False for a, b, t in zip(self.val, x.val, same)]
        return 'subvector types : %s; subvector equality : %s' % (str(same), str(vals))

    __hash__ = object.__hash__

    def __repr__(self):
        s = 'Class : %s\n    name : %s\n' % (self.__class__.__name__, self.name)
        s += '    dim = %d ; size = %d\n' % (self.dim, self.size)
        s += '    blocks : [ %s ]\n' % ' , '.join('%s(%s)' % (getattr(b, 'name', '?'), t)
                                                  for b, t in zip(self.val, self.ltype))
        return s+''.join(str(b) for b in self.val)


class MultiOper():
    """Block operator acting on a MultiVector: row m of the result is sum_n val[m][n]*x[n]
    (matvecs/objects.py:1056-1101)."""

    def __init__(self, name='MultiOper', val=None):
        self.name = name
        self.val = [list(row) for row in val]
        self.no_row, self.no_col = len(self.val), len(self.val[0])
        self.shape = (self.no_row, self.no_col)

    def __call__(self, x):
        if not isinstance(x, MultiVector):
            raise NotImplementedError('MultiOper acts on a MultiVector')
        rows = []
        for row in self.val:
            acc = row[0]*x[0]
            for op, xb in zip(row[1:], x.val[1:]):
                acc = acc+op*xb
            rows.append(acc)
        return MultiVector(val=rows)

    __mul__ = __call__

    def transpose(self):
        return MultiOper(name='(%s)^T' % self.name,
                         val=[[self.val[n][m].transpose() for n in range(self.no_row)] for m in range(self.no_col)])

    def __repr__(self):
        s = 'Class : %s\n    name : %s\n    expression :\n' % (self.__class__.__name__, self.name)
        for row in self.val:
            s += '        [ %s ]\n' % ' , '.join(op.name for op in row)
        return s


class ScipyOper():
    """matvec / rmatvec of a (block) operator over flat host vectors, the shape scipy.sparse.linalg expects
    (matvecs/objects.py:1104-1165); `X` fixes the block structure of the operand."""

    def __init__(self, name='ScipyLinOper', A=None, X=None, AT=None, dtype=None):
        self.name = name
        self.A = A
        self.dtype = np.float64 if dtype is None else dtype
        if AT is not None:
            self.AT = AT
        Y = A(X)
        self.shape = (Y.size, X.size)
        self.X, self.Y = X, Y

    def revec(self, x):
        x = np.asarray(x).ravel()
        if isinstance(self.X, VecTri):
            return VecTri(val=np.reshape(x, self.X._vshape()))
        blocks, end = [], 0
        for m in self.X._iter:
            beg, end = end, end+int(self.X.lsize[m])
            if self.X.ltype[m] != 'VecTri':
                raise NotImplementedError('ScipyOper: block type %s' % self.X.ltype[m])
            blocks.append(VecTri(val=np.reshape(x[beg:end], self.X.lshape[m])))
        return MultiVector(val=blocks)

    revecD = revec

    def matvec(self, x):
        return self.A(self.revec(x)).vec()

    def rmatvec(self, x):
        return self.AT(self.revecD(x)).vec()

    def __repr__(self):
        return 'Class : %s\n    name : %s\n    shape = %s\n    A : %s\n' % (self.__class__.__name__, self.name,
                                                                             str(self.shape), self.A.name)


def enlargeF(xN, M):
    """Spectral interpolation of grid values to the grid M (matvecs/objects.py:1158-1178)."""
    N = tuple(int(n) for n in np.shape(xN))
    M = tuple(int(m) for m in np.array(M).ravel())
    F = DFT.fftnc_dev(dev.upload(np.asarray(xN, dtype=float)), N)
    FM = ops.spec_remap(F, N, 'c', M, 'c', 1, 1.0, flags=1)
    return dev.download(DFT.ifftnc_dev(FM.reshape(M), M))
