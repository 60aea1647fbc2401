"""Green / projection operators in Fourier space — drop-in for ffthompy/projections.py.

`scalar(N, Y, ...)` and `elasticity(N, Y, ...)` return `GreenTensor`s: Tensors whose values
are known in closed form in xi (csrc/fh_green.cuh).  They behave like the reference's
materialised (D, D) + N_fft arrays — `.val`, `+`, scalar `*`, `enlarge`, `P*Q`, `norm` all
work — but as long as nobody asks for `.val` they stay a 7-coefficient descriptor, and
`Operator([[FiN, G, FN]])` applies them inside the FFT pipeline without ever building the
array (2.4 GB per multiplier at 256^3 elasticity in the reference).
"""
import numpy as np

from . import _lib as L
from . import device as dev
from .tensors.objects import Tensor, REFERENCE_QUIRKS  # noqa: F401
from .trigpol import get_Nodd, fft_form_default

KIND_SCALAR, KIND_ELASTIC = 0, 1
_COEFS = ('c0', 'cI', 'cS', 'cH', 'cL', 'cW')


class GreenTensor(Tensor):
    """Lazy projection multiplier.  `green` = {'kind', 'band', 'coef': {c0,cI,cS,cH,cL,cW}} with all
    scale factors folded into the coefficients; None once the values have been materialised
    and exposed to the host."""
    keys = Tensor.keys+('green',)

    def __init__(self, name='', val=None, order=None, shape=None, N=None, Y=None, multype=21,
                 Fourier=True, fft_form=fft_form_default, origin=0, green=None):
        self.green = green
        if green is not None and val is None:
            # lazy: no storage yet
            self.name, self.Fourier, self.origin = name, Fourier, origin
            self._h = self._d = None
            self.N = tuple(int(n) for n in np.array(N, dtype=int))
            self._set_fft(fft_form)
            self.shape = tuple(int(s) for s in shape)
            self.order = len(self.shape)
            self.dim = len(self.N)
            self.Y = np.ones(self.dim) if Y is None else np.array(Y, dtype=float)
            self.multype = multype
        else:
            if val is not None:
                shape = None
            Tensor.__init__(self, name=name, val=val, order=order, shape=shape, N=N, Y=Y, multype=multype,
                            Fourier=Fourier, fft_form=fft_form, origin=origin)

    # ---------------------------------------------------------------- storage
    @property
    def lazy(self):
        return self.green is not None

    def descriptor(self):
        """C struct fh_green of the current closed form."""
        g = L.fh_green()
        g.kind = self.green['kind']
        g.dim = self.dim
        for a in range(self.dim):
            g.N[a] = self.N[a]
            g.band[a] = self.green['band'][a]
            g.Y[a] = self.Y[a]
        for k in _COEFS:
            setattr(g, k, float(self.green['coef'][k]))
        g.scale = 1.0
        return g

    def _materialize(self):
        out = dev.empty(self.shape+self.N_fft)
        g = self.descriptor()
        L.check(dev.lib().fh_green_materialize(g, dev.form_code(self.fft_form), dev.ptr(out)))
        return out

    def _dev(self):
        if self._d is None and self._h is None and self.lazy:
            self._d = self._materialize()
        return Tensor._dev(self)

    @property
    def val(self):
        if self._h is None and self._d is None and self.lazy:
            self._d = self._materialize()
        self.green = None  # values are now exposed to (and may be changed by) the host
        return Tensor.val.fget(self)

    @val.setter
    def val(self, v):
        self.green = None
        Tensor.val.fset(self, v)

    def _vshape(self):
        if self._d is None and self._h is None:
            return tuple(self.shape)+tuple(self.N_fft)
        return Tensor._vshape(self)

    def _is_complex(self):
        if self._d is None and self._h is None:
            return False
        return Tensor._is_complex(self)

    def _copy(self, keys, **kwargs):
        if self.lazy and 'val' not in kwargs:
            args = dict(name=self.name, shape=self.shape, N=self.N, Y=self.Y, multype=self.multype,
                        Fourier=self.Fourier, fft_form=self.fft_form, origin=self.origin,
                        green={'kind': self.green['kind'], 'band': tuple(self.green['band']),
                               'coef': dict(self.green['coef'])})
            args.update({k: v for k, v in kwargs.items() if k in args})
            return GreenTensor(**args)
        # materialised data (or an explicit result array): an ordinary Tensor
        data = {k: getattr(self, k) for k in Tensor.keys if k != 'val' and k not in kwargs}
        if 'val' not in kwargs:
            data['val'] = Tensor._val_copy(self)
        data.update(kwargs)
        data.pop('green', None)
        return Tensor(**data)

    def _with_coef(self, coef, name=None, **kw):
        g = {'kind': self.green['kind'], 'band': tuple(self.green['band']), 'coef': coef}
        return GreenTensor(name=self.name if name is None else name, shape=self.shape,
                           N=kw.get('N', self.N), Y=self.Y, multype=self.multype, Fourier=True,
                           fft_form=kw.get('fft_form', self.fft_form), origin=self.origin, green=g)

    def _compatible(self, x):
        return (isinstance(x, GreenTensor) and self.lazy and x.lazy and self.green['kind'] == x.green['kind']
                and tuple(self.green['band']) == tuple(x.green['band']) and tuple(self.N) == tuple(x.N)
                and self.fft_form == x.fft_form and np.allclose(self.Y, x.Y))

    # ---------------------------------------------------------------- algebra that stays lazy
    def __neg__(self):
        if not self.lazy:
            return Tensor.__neg__(self)
        return self._with_coef({k: -v for k, v in self.green['coef'].items()}, name='-'+self.name[:10])

    def __add__(self, x):
        if self._compatible(x):
            coef = {k: self.green['coef'][k]+x.green['coef'][k] for k in _COEFS}
            return self._with_coef(coef, name='({0}+{1})'.format(self.name[:10], x.name[:10]))
        return Tensor.__add__(self, x)

    def __rmul__(self, x):
        if self.lazy and not hasattr(x, 'val') and np.size(x) == 1 and not np.iscomplexobj(x):
            s = float(np.asarray(x).ravel()[0])
            return self._with_coef({k: s*v for k, v in self.green['coef'].items()})
        return Tensor.__rmul__(self, x)

    def transpose(self):
        if self.lazy:  # all closed forms are symmetric matrices
            return self._with_coef(dict(self.green['coef']), name=self.name[:10]+'.T')
        return Tensor.transpose(self)

    def set_fft_form(self, fft_form=fft_form_default, copy=False):
        if not self.lazy:
            return Tensor.set_fft_form(self, fft_form, copy)
        R = self._with_coef(dict(self.green['coef'])) if copy else self
        if self.fft_form == fft_form:
            return R
        pN = float(np.prod(self.N))
        s = 1.
        if self.fft_form == 'r':  # tensors/objects.py:154
            s = 1./pN
        elif fft_form == 'r':     # tensors/objects.py:160,165
            s = pN
        R.green['coef'] = {k: s*v for k, v in R.green['coef'].items()}
        R._d = None  # drop the cached materialisation
        R._set_fft(fft_form)
        return R

    def _nyquist_free(self):
        return all(n % 2 == 1 or 2*b < n for n, b in zip(self.N, self.green['band']))

    def enlarge(self, M):
        """Zero padding to the grid M.  On an 'r'-form multiplier the reference's round trip through
        the 'c' form multiplies the values by prod(M)/prod(N) (tensors/objects.py:144-166,428-467;
        relied upon by applications.py:28-31) — reproduced here on the coefficients."""
        assert(self.Fourier)
        if np.allclose(self.N, M):
            return self
        if not (self.lazy and self._nyquist_free()):
            return Tensor.enlarge(self, M)
        M = tuple(int(m) for m in np.array(M).ravel())
        s = float(np.prod(M))/float(np.prod(self.N)) if self.fft_form == 'r' else 1.
        R = self._with_coef({k: s*v for k, v in self.green['coef'].items()}, N=M)
        if REFERENCE_QUIRKS:
            self.set_fft_form('c')
        return R

    def _lazy_apply(self, y):
        """G(y) for a spectrum y (shape (D,[K...]) + N_fft, complex) without materialising G."""
        if not self.lazy:
            return None
        if not (y.Fourier and y._is_complex() and y.order >= 1 and int(y.shape[0]) == int(self.shape[0])):
            return None
        K = y._ncomp//int(y.shape[0])
        out = dev.empty(y._vshape(), complex_=True)
        g = self.descriptor()
        L.check(dev.lib().fh_green_apply(g, dev.form_code(self.fft_form), int(K), dev.ptr(y._dev()), dev.ptr(out)))
        return out


def _green(name, kind, N, Y, band, fft_form, **coef):
    N = tuple(int(n) for n in N)
    d = len(N)
    D = d if kind == KIND_SCALAR else d*(d+1)//2
    c = {k: 0. for k in _COEFS}
    c.update(coef)
    return GreenTensor(name=name, shape=(D, D), N=N, Y=Y, multype=21, Fourier=True, fft_form=fft_form,
                       green={'kind': kind, 'band': tuple(int(b) for b in band), 'coef': c})


def _band(N, NyqNul):
    N = np.array(N, dtype=int)
    if NyqNul:
        return (get_Nodd(N)-1)//2
    return N//2


def scalar(N, Y, NyqNul=True, tensor=True, fft_form=fft_form_default):
    """Projections for scalar elliptic problems (ffthompy/projections.py:9-112).

    Returns (G0, G1, G2): the mean projection, the projection on curl-free zero-mean fields
    (xi (x) xi / |xi|^2) and on divergence-free zero-mean fields (I - G1), with the Nyquist
    frequencies of even grids zeroed when `NyqNul`."""
    if not tensor:
        raise NotImplementedError("tensor=False returned the deprecated matvecs.Matrix class; in the "
                                  "reference this path already fails for the default fft_form "
                                  "(projections.py:107-110). Use ffthompy_b200.matvecs for the legacy API.")
    N = np.array(N, dtype=int)
    Y = np.array(Y, dtype=float)
    band = _band(N, NyqNul)
    G0 = _green('hG0', KIND_SCALAR, N, Y, band, fft_form, c0=1.)
    G1 = _green('hG1', KIND_SCALAR, N, Y, band, fft_form, cH=1.)
    G2 = _green('hG2', KIND_SCALAR, N, Y, band, fft_form, cI=1., cH=-1.)
    return G0, G1, G2


def elasticity(N, Y, NyqNul=True, tensor=True, fft_form=fft_form_default):
    """Projections on admissible strain/stress fields in Mandel notation
    (ffthompy/projections.py:114-267).  Returns (G0, G1h, G1s, G2h, G2s).

    Even N with NyqNul — which raises in the shipped reference (projections.py:131,148-152) —
    is defined as SURVEY App. D.1 / projections.scalar do: assembled on the odd grid Nred and
    zero-padded, i.e. the band |k_i| <= (Nred_i-1)/2."""
    if not tensor:
        raise NotImplementedError("tensor=False (matvecs.Matrix) is not supported; see scalar().")
    N = np.array(N, dtype=int)
    Y = np.array(Y, dtype=float)
    d = N.size
    band = _band(N, NyqNul)
    G0 = _green('hG0', KIND_ELASTIC, N, Y, band, fft_form, c0=1.)
    G1h = _green('hG1h', KIND_ELASTIC, N, Y, band, fft_form, cH=1.)
    G1s = _green('hG1s', KIND_ELASTIC, N, Y, band, fft_form, cS=1., cH=-2.)
    G2h = _green('hG2h', KIND_ELASTIC, N, Y, band, fft_form, cH=1./(d-1), cL=float(d)/(d-1), cW=-1./(d-1))
    # G2s = IS0 - G1h - G1s - G2h (projections.py:240)
    G2s = _green('hG2s', KIND_ELASTIC, N, Y, band, fft_form, cI=1., cS=-1., cH=-1.+2.-1./(d-1),
                 cL=-float(d)/(d-1), cW=1./(d-1))
    return G0, G1h, G1s, G2h, G2s
