"""Tensors of trigonometric polynomials on the B200 — drop-in for
ffthompy/tensors/objects.py (class Tensor, Scalar, einsum, scalar_product, norm_fun).

`Tensor` keeps the reference's constructor, attributes and methods; the values live in
HBM (a torch CUDA buffer) and every operation is a kernel of libffthom_b200.so.

Host/device coherence of `.val`
    `.val` hands out a NumPy array, exactly as the reference does.  Reading it
    downloads the data (if the device copy is newer) and makes the HOST copy
    authoritative, because the caller may mutate the array in place (`X.val[...] = ...`,
    `X.val /= ...` are common in the reference's callers).  The next device operation
    uploads it again.  Objects that are only ever combined through the operator algebra
    (the whole CG loop) never leave the device.
"""
import itertools
from copy import copy

import numpy as np

from .. import device as dev
from .. import ops
from ..general.base import Representation
from ..trigpol import mean_index, fft_form_default

# Tensor.enlarge/decrease leave the operand converted to the 'c' form in the reference
# (tensors/objects.py:438,479).  Reproduced by default; see INTEGRATION.md.
REFERENCE_QUIRKS = True


def _is_dev(v):
    return type(v).__module__.startswith('torch') and hasattr(v, 'data_ptr')


class TensorFuns(Representation):

    def mean_index(self):
        return mean_index(self.N, self.fft_form)

    def __getitem__(self, ii):
        return self.val[ii]

    def pN(self):
        return np.prod(self.N)

    def point(self, ii):
        val = np.empty(self.shape)
        for ind in np.ndindex(*self.shape):
            val[ind] = self.val[ind][ii]
        return val

    def sub(self, ii):
        self.val[ii]

    def update(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    def _copy(self, keys, **kwargs):
        data = {}
        for k in keys:
            if k in kwargs:
                continue
            if k == 'val':
                data[k] = self._val_copy()
            else:
                data[k] = copy(getattr(self, k))
        data.update(kwargs)
        return self.__class__(**data)

    def copy(self, **kwargs):
        return self._copy(self.keys, **kwargs)

    def _set_fft(self, fft_form):
        assert(fft_form in ['c', 'r', 0])
        if fft_form in ['r']:
            self.N_fft = self.get_N_real(self.N)
            self.fft_coef = np.prod(self.N)
        else:
            self.N_fft = tuple(self.N)
            self.fft_coef = 1.
        self.fft_form = fft_form

    def __repr__(self, full=False, detailed=False):
        keys = ['order', 'name', 'Y', 'shape', 'N', 'Fourier', 'fft_form', 'origin', 'norm']
        ss = self._repr(keys)
        skip = 4*' '
        if np.prod(np.array(self.shape)) <= 36 or detailed:
            ss += '{0}norm component-wise =\n{1}\n'.format(skip, str(self.norm(componentwise=True)))
            ss += '{0}mean = \n{1}\n'.format(skip, str(self.mean()))
        if full:
            ss += '{0}val = \n{1}'.format(skip, str(self.val))
        return ss

    @staticmethod
    def get_N_real(N):
        N_rfft = np.copy(N)
        N_rfft[-1] = int(np.fix(N[-1]/2)+1)
        return tuple(int(n) for n in N_rfft)

    @staticmethod
    def get_N(N_rfft):
        N = np.copy(N_rfft)
        N[-1] = N_rfft[-1]*2-1
        return tuple(N)


class Tensor(TensorFuns):
    keys = ('name', 'val', 'order', 'Y', 'N', 'multype', 'Fourier', 'fft_form', 'origin')  # default keys

    def __init__(self, name='', val=None, order=None, shape=None, N=None, Y=None,
                 multype='scal', Fourier=False, fft_form=fft_form_default, origin=0):
        self.name = name
        self.Fourier = Fourier
        self.origin = origin
        self._h = None  # host copy (numpy) or None
        self._d = None  # device copy (torch CUDA tensor) or None
        self._version = 0  # bumped by every in-place change of the device buffer (caches key on it)

        if isinstance(val, np.ndarray) or _is_dev(val):  # define: val + order
            self.val = val
            self.order = int(order)
            vshape = tuple(int(s) for s in val.shape)
            self.shape = vshape[:self.order]
            if fft_form in ['r'] and Fourier:
                self.N = tuple(int(n) for n in np.array(N, dtype=int))
            else:
                self.N = vshape[self.order:]
            self._set_fft(fft_form)

        elif shape is not None and N is not None:  # define: shape + N
            self.N = tuple(int(n) for n in np.array(N, dtype=int))
            self._set_fft(fft_form)
            self.shape = tuple(int(s) for s in np.array(shape, dtype=int))
            self.order = len(self.shape)
            if not self.Fourier:
                self.val = dev.zeros(self.shape+self.N)
            else:
                self.val = dev.zeros(self.shape+self.N_fft, complex_=True)
        else:
            raise ValueError('Initialization of Tensor.')

        self.dim = len(self.N)
        if Y is None:
            self.Y = np.ones(self.dim, dtype=float)
        else:
            self.Y = np.array(Y, dtype=float)

        # definition of __mul__ operation
        self.multype = multype

    # ------------------------------------------------------------------ storage
    @property
    def val(self):
        if self._h is None:
            self._h = dev.download(self._d)
        self._d = None  # the caller may mutate the array: the host copy becomes authoritative
        return self._h

    @val.setter
    def val(self, v):
        if _is_dev(v):
            self._d, self._h = v, None
        else:
            self._h, self._d = np.asarray(v), None

    def _dev(self):
        """device buffer (uploading the host copy if it is the newer one)"""
        if self._d is None:
            self._d = dev.upload(self._h)
        return self._d

    def _val_copy(self):
        if self._d is not None:
            return ops.clone(self._d)
        return np.copy(self._h)

    def _vshape(self):
        return tuple((self._d if self._d is not None else self._h).shape)

    def _is_complex(self):
        if self._d is not None:
            return dev.is_complex(self._d)
        return np.iscomplexobj(self._h)

    @property
    def _ngrid(self):
        """points per component in the stored array"""
        return int(np.prod(self._vshape()[self.order:]))

    @property
    def _ncomp(self):
        return int(np.prod(self.shape)) if len(self.shape) else 1

    # ------------------------------------------------------------------ fft forms
    def set_fft_form(self, fft_form=fft_form_default, copy=False):
        """tensors/objects.py:135-167"""
        R = self.copy() if copy else self
        if self.fft_form == fft_form:
            return R
        if R.Fourier:
            pN = float(np.prod(R.N))
            scale = 1.
            if R.fft_form == 'r':
                scale = 1./pN
            elif fft_form == 'r':
                scale = pN
            was_real = not R._is_complex()
            new = ops.spec_remap(R._dev(), R.N, R.fft_form, R.N, fft_form, R._ncomp, scale)
            if was_real:
                new = ops.convert(new, False)
            shp = R.get_N_real(R.N) if fft_form == 'r' else tuple(R.N)
            R.val = new.reshape(R.shape+shp)
        R._set_fft(fft_form)
        return R

    def shift(self, origin=None):
        """Shift the origin in the real domain (tensors/objects.py:169-186)."""
        assert(not self.Fourier)
        if origin == self.origin:
            return self
        elif origin is None:
            if self.origin in [0]:
                sh = [n//2 for n in self.N]  # fftshift
                self.val = ops.roll(self._dev(), self.N, sh, self._ncomp)
                self.origin = 'c'
            elif self.origin in ['c']:
                sh = [-(n//2) for n in self.N]  # ifftshift
                self.val = ops.roll(self._dev(), self.N, sh, self._ncomp)
                self.origin = 0
            return self
        else:
            raise ValueError()

    def randomize(self):
        shp = self._vshape()
        val = np.random.random(shp)
        if self.Fourier:
            val = val+1j*np.random.random(shp)
        self.val = val
        return self

    # ------------------------------------------------------------------ algebra
    def __neg__(self):
        return self.copy(name='-'+self.name[:10], val=ops.axpby(-1., self._dev()))

    def __add__(self, x):
        if isinstance(x, Tensor):
            assert(self.Fourier == x.Fourier)
            assert(self._vshape() == x._vshape())
            name = '({0}+{1})'.format(self.name[:10], x.name[:10])
            a, b = ops.promote(self._dev(), x._dev())
            return self.copy(name=name, val=ops.axpby(1., a, 1., b))
        elif isinstance(x, float) or (isinstance(x, (int, np.floating, np.integer)) and not isinstance(x, bool)):
            if float(x) == 0.:
                return self.copy()
            return self.copy(val=ops.add_scalar(self._dev(), x))
        elif isinstance(x, np.ndarray):
            if x.size == 1 and not np.iscomplexobj(x):
                return self.copy(val=ops.add_scalar(self._dev(), float(x.ravel()[0])))
            full = np.broadcast_to(x, np.broadcast_shapes(x.shape, self._vshape()))
            assert(full.shape == self._vshape())
            a, b = ops.promote(self._dev(), dev.upload(full))
            return self.copy(val=ops.axpby(1., a, 1., b))
        else:
            raise ValueError('Tensor.__add__')

    def __sub__(self, x):
        return self.__add__(-x)

    def __rmul__(self, x):
        if isinstance(x, Scalar):
            return self.copy(val=ops.axpby(x.val, self._dev()))
        elif np.size(x) == 1:
            xv = np.asarray(x).ravel()[0]
            if np.iscomplexobj(xv) and xv.imag != 0:
                raise ValueError('complex scalar factors are not supported')
            return self.copy(val=ops.axpby(float(np.real(xv)), self._dev()))
        else:
            raise ValueError()

    def __call__(self, *args, **kwargs):
        return self.__mul__(*args, **kwargs)

    def __mul__(self, Y, multype=None, *args, **kwargs):
        """tensors/objects.py:223-245"""
        if multype is None:
            multype = self.multype
        X = self
        assert(X.Fourier == Y.Fourier)
        assert(X.fft_form == Y.fft_form)
        if multype in ['scal', 'scalar']:
            return scalar_product(X, Y)
        elif multype in [21, '21']:
            return einsum('ij...,j...->i...', X, Y)
        elif multype in [42, '42']:
            return einsum('ijkl...,kl...->ij...', X, Y)
        elif multype in [00, 'elementwise', 'hadamard']:
            return einsum('...,...->...', X, Y)
        elif multype in ['grad']:
            return einsum('i...,...->i...', X, Y)
        elif multype in ['div']:
            return einsum('i...,i...->...', X, Y)
        else:
            try:
                return einsum(multype, X, Y)
            except Exception:
                raise ValueError()

    def inv(self):
        """point-wise matrix inverse (tensors/objects.py:247-251, trigpol.py:120-159)"""
        assert(self.Fourier is False)
        assert(self.order == 2)
        assert(self.shape[0] == self.shape[1])
        val = ops.inv_dxd(self._dev(), self.shape[0], self._ngrid)
        return self.copy(name='inv({})'.format(self.name), val=val)

    def norm(self, ntype='L2', componentwise=False):
        if componentwise:
            scal = np.empty(self.shape)
            d = self._dev()
            for ind in np.ndindex(*self.shape):
                obj = Tensor(name='aux', val=d[ind], order=0, N=self.N, Y=self.Y, multype=self.multype,
                             Fourier=self.Fourier, fft_form=self.fft_form, origin=self.origin)
                scal[ind] = norm_fun(obj, ntype=ntype)
            return scal
        else:
            return norm_fun(self, ntype=ntype)

    def mean(self):
        """Mean of the trigonometric polynomial (tensors/objects.py:263-275)."""
        mean = np.zeros(self.shape)
        if self.Fourier:
            off = int(np.ravel_multi_index(self.mean_index(), self._vshape()[self.order:]))
            d = self._dev()
            e = 2 if dev.is_complex(d) else 1
            for c, di in enumerate(np.ndindex(*self.shape)):
                mean[di] = ops.peek(d, (c*self._ngrid+off)*e, 1)[0]/self.fft_coef
        else:
            sums = ops.sum_comp(self._dev(), self._ncomp)
            mean[...] = (sums/self._ngrid).reshape(self.shape)
        return mean

    def add_mean(self, mean):
        """tensors/objects.py:277-287 (the Fourier branch SETS the zero frequency, as the reference does)"""
        mean = np.asarray(mean, dtype=float)
        assert(self.shape == mean.shape)
        d = self._dev()
        if self.Fourier:
            off = int(np.ravel_multi_index(self.mean_index(), self._vshape()[self.order:]))
            cplx = dev.is_complex(d)
            for c, di in enumerate(np.ndindex(*self.shape)):
                v = float(mean[di]*self.fft_coef)
                ops.poke(d, (c*self._ngrid+off)*(2 if cplx else 1), [v, 0.] if cplx else [v])
        else:
            ops.add_comp(d, [float(v) for v in mean.ravel()])
        self._h = None  # device copy is now the newer one
        self._version += 1
        return self

    def set_mean(self, mean):
        mean = np.asarray(mean, dtype=float)
        assert(self.shape == mean.shape)
        self.add_mean(-self.mean())  # set mean to zero
        return self.add_mean(mean)

    def __eq__(self, Y, full=True, tol=1e-13):
        """Equality check up to `tol` (tensors/objects.py:302-317)."""
        X = self
        _bool = False
        res = np.inf
        squeezed = lambda shp: tuple(s for s in shp if s != 1)  # noqa: E731
        if (isinstance(Y, Tensor) and X.fft_form == Y.fft_form and
                squeezed(X._vshape()) == squeezed(Y._vshape()) and X.Fourier == Y.Fourier):
            a, b = ops.promote(X._dev(), Y._dev())
            diff = ops.axpby(1., a, -1., b.reshape(a.shape))
            res = ops.dot(diff, diff)**0.5
            if res < tol:
                _bool = True
        if full:
            return _bool, res
        else:
            return _bool

    __hash__ = object.__hash__

    def set_shape(self):
        shape_size = len(self._vshape())-len(self.N)
        self.shape = np.array(self._vshape()[:shape_size])
        return self.shape

    def _permute(self, perm, name):
        out = ops.gather_comps(self._dev(), perm, self._ncomp)
        return self.copy(name=name, val=out.reshape(self._vshape()))

    def transpose(self):
        """tensors/objects.py:324-331"""
        s = self.shape
        if self.order == 2:
            assert(s[0] == s[1])
            perm = [j*s[1]+i for i in range(s[0]) for j in range(s[1])]
        elif self.order == 4:
            idx = np.arange(int(np.prod(s))).reshape(s)
            perm = list(np.einsum('ijkl->klij', idx).ravel())
        else:
            raise NotImplementedError()
        return self._permute(perm, self.name[:10]+'.T')

    def transpose_left(self):
        assert(self.order == 4)
        idx = np.arange(int(np.prod(self.shape))).reshape(self.shape)
        return self._permute(list(np.einsum('ijkl->jikl', idx).ravel()), self.name[:10]+'.T')

    def transpose_right(self):
        assert(self.order == 4)
        idx = np.arange(int(np.prod(self.shape))).reshape(self.shape)
        return self._permute(list(np.einsum('ijkl->ijlk', idx).ravel()), self.name[:10]+'.T')

    def identity(self):
        """tensors/objects.py:345-349"""
        assert(self.order % 2 == 0)
        val = np.zeros(self._vshape(), dtype=complex if self._is_complex() else float)
        for ii in itertools.product(*tuple([list(range(n)) for n in self.shape[:int(self.order/2)]])):
            val[ii+ii] = 1.
        self.val = val

    def vec(self):
        return np.matrix(self.val.ravel()).transpose()

    def zeros_like(self, name=None):
        if name is None:
            name = 'zeros({})'.format(self.name[:10])
        return self.copy(name=name, val=dev.zeros(self._vshape(), complex_=self._is_complex()))

    def empty_like(self, name=None):
        if name is None:
            name = 'empty({})'.format(self.name[:10])
        return self.copy(name=name, val=dev.empty(self._vshape(), complex_=self._is_complex()))

    def calc_eigs(self, sort=True, symmetric=False, mandel=False):
        raise NotImplementedError('calc_eigs is host-side diagnostics of the reference '
                                  '(tensors/objects.py:368-400); not part of the solve loop')

    @property
    def axes(self):  # axes for Fourier transform
        return tuple(range(self.order, self.order+self.dim))

    # ------------------------------------------------------------------ transforms
    def _fft_val(self):
        """forward transform of the stored real values in the convention of self.fft_form"""
        X = ops.rfftn(ops.convert(self._dev(), False), self.N, self._ncomp)
        if self.fft_form != 'r':
            X = ops.spec_remap(X, self.N, 'r', self.N, self.fft_form, self._ncomp, 1./float(np.prod(self.N)))
        return X.reshape(self.shape+self.N_fft)

    def _ifft_val(self):
        """inverse transform (real part), tensors/fft.py:25-43"""
        X = ops.convert(self._dev(), True)
        pN = float(np.prod(self.N))
        if self.fft_form == 'r':
            x = ops.irfftn(X, self.N, self._ncomp, 1./pN)
        else:
            # ifftn(X).real * prod(N) == un-normalised inverse of the Hermitian part of X
            H = ops.spec_remap(X, self.N, self.fft_form, self.N, 'r', self._ncomp, 1., flags=2)
            x = ops.irfftn(H, self.N, self._ncomp, 1.)
        return x.reshape(self.shape+tuple(self.N))

    def fourier(self, Fourier=None, copy=False):
        """tensors/objects.py:406-426"""
        assert(self.origin == 0)
        if self.Fourier == Fourier:
            if copy:
                return self.copy()
            else:
                return self
        new = self._ifft_val() if self.Fourier else self._fft_val()
        if copy:
            return self.copy(val=new, Fourier=not self.Fourier)
        else:
            self.val = new
            self.Fourier = not self.Fourier
            return self

    # ------------------------------------------------------------------ resampling
    def _resample(self, M):
        M = tuple(int(m) for m in np.array(M).ravel())
        scale = float(np.prod(M))/float(np.prod(self.N)) if self.fft_form == 'r' else 1.
        was_real = not self._is_complex()
        new = ops.spec_remap(self._dev(), self.N, self.fft_form, M, self.fft_form, self._ncomp, scale)
        if was_real:
            new = ops.convert(new, False)
        shp = self.get_N_real(M) if self.fft_form == 'r' else M
        R = self.copy(val=new.reshape(self.shape+tuple(shp)), N=M)
        if REFERENCE_QUIRKS:
            self.set_fft_form('c')  # tensors/objects.py:438,479 leave the operand in 'c' form
        return R

    def enlarge(self, M):
        """Zero-pad the Fourier coefficients to the grid M (tensors/objects.py:428-467).  Even axes
        get their Nyquist plane split; 'r'-form values are rescaled by prod(M)/prod(N)."""
        assert(self.Fourier)
        if np.allclose(self.N, M):
            return self
        return self._resample(M)

    def decrease(self, M):
        """Drop the high frequencies (tensors/objects.py:469-486)."""
        assert(self.Fourier)
        if np.allclose(self.N, M):
            return self
        return self._resample(M)

    def project(self, M):
        """tensors/objects.py:488-511"""
        if np.allclose(self.N, M):
            return self
        Fourier = self.Fourier
        if Fourier:
            Y = self.copy()
        else:
            Y = self.fourier(copy=True)
        if np.all(np.greater(M, self.N)):
            Y = Y.enlarge(M)
        elif np.all(np.less(M, self.N)):
            Y = Y.decrease(M)
        else:
            raise NotImplementedError()
        if not Fourier:
            Y = Y.fourier()
        return Y

    def subfield(self, Y=None, M=None):
        """tensors/objects.py:513-534"""
        N = np.array(self.N)
        if Y is None and M is None:
            raise ValueError('Either Y or M has to be specified.')
        elif Y is not None:
            M = np.ceil(Y/self.Y*N).astype(int)
        elif M is not None:
            M = np.ceil(M).astype(int)
        ind = [slice(None) for i in range(len(self.shape))]
        beg = np.round((N-M)/2).astype(int)
        ind = tuple(ind+[slice(beg[i], beg[i]+M[i]) for i in range(self.dim)])
        return self.copy(val=np.ascontiguousarray(self.val[ind]))


class Scalar():
    """Scalar value that multiplies Tensors (tensors/objects.py:576-596)."""

    def __init__(self, val=None, name='c'):
        if val is not None:
            self.val = val
        else:
            self.val = 1.
        self.name = name

    def __call__(self, x):
        return self*x

    def __mul__(self, x):
        return x.__rmul__(self)

    def __repr__(self):
        ss = "Class : {0}\n".format(self.__class__.__name__)
        ss += "    val = {0}".format(self.val)
        return ss

    def transpose(self):
        return self


def einsum(str_operator, x, y):
    """Point-wise contractions behind Tensor.__mul__ (tensors/objects.py:599-604): the patterns the
    reference uses are dispatched to device kernels; anything else raises."""
    assert(x.Fourier == y.Fourier)
    assert(np.all(np.array(x.N) == np.array(y.N)))
    name = '{0}({1})'.format(x.name, y.name)
    n = y._ngrid
    if hasattr(x, '_lazy_apply') and str_operator == 'ij...,j...->i...':
        res = x._lazy_apply(y)
        if res is not None:
            return y.copy(name=name, val=res, order=y.order)
    xs, ys = x._vshape(), y._vshape()
    if str_operator in ('ij...,j...->i...', 'ijkl...,kl...->ij...'):
        half = 1 if str_operator.startswith('ij.') else 2
        D = int(np.prod(x.shape[:half]))
        assert(tuple(x.shape[half:2*half]) == tuple(y.shape[:half]))
        assert(xs[2*half:] == ys[y.order:])
        K = int(np.prod(y.shape[half:])) if len(y.shape) > half else 1
        val = ops.mul21(x._dev(), y._dev(), D, n, K)
        oshape = tuple(x.shape[:half])+tuple(y.shape[half:])
        return y.copy(name=name, val=val.reshape(oshape+ys[y.order:]), order=len(oshape))
    if str_operator == '...,...->...':
        grid = ys[y.order:]
        assert(xs[x.order:] == grid)
        ca, cb = x._ncomp, y._ncomp
        if tuple(x.shape) == tuple(y.shape):
            oshape, nc, adiv, bdiv = tuple(y.shape), cb, 1, 1
        elif len(x.shape) <= len(y.shape) and tuple(y.shape[len(y.shape)-len(x.shape):]) == tuple(x.shape):
            oshape, nc, adiv, bdiv = tuple(y.shape), cb, 1, 1  # x broadcast over leading components of y
        elif tuple(x.shape[len(x.shape)-len(y.shape):]) == tuple(y.shape):
            oshape, nc, adiv, bdiv = tuple(x.shape), ca, 1, 1
        else:
            raise ValueError('hadamard: shapes %s and %s do not broadcast' % (x.shape, y.shape))
        val = ops.hadamard(x._dev(), y._dev(), n, nc, adiv, ca, bdiv, cb)
        return y.copy(name=name, val=val.reshape(oshape+grid), order=len(oshape))
    if str_operator == 'i...,...->i...':
        grid = ys[y.order:]
        d, R = int(x.shape[0]), y._ncomp
        assert(x.order == 1 and xs[1:] == grid)
        val = ops.hadamard(x._dev(), y._dev(), n, d*R, R, d, 1, R)
        oshape = (d,)+tuple(y.shape)
        return y.copy(name=name, val=val.reshape(oshape+grid), order=len(oshape))
    if str_operator == 'i...,i...->...':
        grid = ys[y.order:]
        d = int(x.shape[0])
        assert(x.order == 1 and int(y.shape[0]) == d and xs[1:] == grid)
        R = y._ncomp//d
        val = ops.contract_first(x._dev(), y._dev(), n, d, R)
        oshape = tuple(y.shape[1:])
        return y.copy(name=name, val=val.reshape(oshape+grid), order=len(oshape))
    raise NotImplementedError('einsum pattern %r has no device kernel' % (str_operator,))


def norm_fun(X, ntype):
    """tensors/objects.py:606-616"""
    if ntype in ['L2', 2]:
        scal = (scalar_product(X, X))**0.5
    elif ntype == 1:
        scal = ops.asum(X._dev())
    elif ntype == 'inf':
        scal = ops.amax(X._dev())
    else:
        msg = "This type ({}) of norm is not implemented!".format(ntype)
        raise NotImplementedError(msg)
    return scal


def scalar_product(y, x):
    """tensors/objects.py:618-636"""
    assert(isinstance(x, Tensor))
    assert(y._vshape() == x._vshape())
    assert(y.fft_form == x.fft_form)
    a, b = ops.promote(y._dev(), x._dev())
    if y.Fourier:
        if x.fft_form in ['r']:
            batch = int(np.prod(y._vshape()[:len(y._vshape())-y.dim]))
            scal = ops.dot_rspec(y.N, batch, a, b)/np.prod(y.N)**2
        else:
            scal = ops.dot(a, b)
    else:
        scal = ops.dot(a, b)/np.prod(y.N)
    return scal
