"""Thin typed wrappers over the C ABI working on device buffers (torch CUDA tensors).
Nothing here computes on the host."""
import ctypes as C

import numpy as np

from . import _lib as L
from . import device as dev


def _n(t):
    return dev.ndoubles(t)


def axpby(a, x, b=0.0, y=None):
    """a*x + b*y (same dtype/shape); y None -> a*x."""
    out = dev.empty(x.shape, complex_=dev.is_complex(x))
    L.check(dev.lib().fh_axpby(_n(x), float(a), dev.ptr(x), float(b), dev.ptr(y), dev.ptr(out)))
    return out


def add_scalar(x, s):
    out = dev.empty(x.shape, complex_=dev.is_complex(x))
    L.check(dev.lib().fh_add_scalar(_n(x), dev.ptr(x), float(s), int(dev.is_complex(x)), dev.ptr(out)))
    return out


def clone(x):
    out = dev.empty(x.shape, complex_=dev.is_complex(x))
    L.check(dev.lib().fh_copy(dev.ptr(out), dev.ptr(x), _n(x)))
    return out


def convert(x, to_complex):
    if dev.is_complex(x) == bool(to_complex):
        return x
    out = dev.empty(x.shape, complex_=bool(to_complex))
    L.check(dev.lib().fh_convert(int(x.numel()), dev.ptr(x), int(dev.is_complex(x)), dev.ptr(out), int(to_complex)))
    return out


def promote(x, y):
    c = dev.is_complex(x) or dev.is_complex(y)
    return convert(x, c), convert(y, c)


def dot(x, y):
    """sum over all doubles of x*y (for complex: Re sum x conj(y))."""
    r = C.c_double()
    L.check(dev.lib().fh_dot(_n(x), dev.ptr(x), dev.ptr(y), C.byref(r)))
    return r.value


def dot_rspec(N, batch, x, y):
    r = C.c_double()
    L.check(dev.lib().fh_dot_rspec(dev.plan(N), int(batch), int(dev.is_complex(x)), dev.ptr(x), dev.ptr(y),
                                   C.byref(r)))
    return r.value


def asum(x):
    r = C.c_double()
    L.check(dev.lib().fh_asum(int(x.numel()), dev.ptr(x), int(dev.is_complex(x)), C.byref(r)))
    return r.value


def amax(x):
    r = C.c_double()
    L.check(dev.lib().fh_amax(int(x.numel()), dev.ptr(x), int(dev.is_complex(x)), C.byref(r)))
    return r.value


def sum_comp(x, ncomp):
    """per-component sums of a real array viewed as (ncomp, n)."""
    n = int(x.numel())//max(int(ncomp), 1)
    out = np.zeros(int(ncomp))
    for c0 in range(0, int(ncomp), 64):
        c1 = min(int(ncomp), c0+64)
        buf = (C.c_double*(c1-c0))()
        sub = x.reshape(int(ncomp), n)[c0:c1]
        L.check(dev.lib().fh_sum_comp(c1-c0, n, dev.ptr(sub), buf))
        out[c0:c1] = buf[:]
    return out


def add_comp(x, vals):
    ncomp = len(vals)
    n = int(x.numel())//max(ncomp, 1)
    L.check(dev.lib().fh_add_comp(ncomp, n, dev.ptr(x), L.dblarr(vals)))


def peek(x, offset, count):
    buf = (C.c_double*count)()
    L.check(dev.lib().fh_peek(dev.ptr(x), int(offset), buf, int(count)))
    return np.array(buf[:])


def poke(x, offset, vals):
    L.check(dev.lib().fh_poke(dev.ptr(x), int(offset), L.dblarr(vals), len(vals)))


def gather_comps(x, perm, ncomp_in):
    """out[c] = x[perm[c]] over component planes; x viewed as (ncomp_in, n)."""
    cplx = dev.is_complex(x)
    n = _n(x)//int(ncomp_in)
    out = dev.empty((len(perm), n//(2 if cplx else 1)), complex_=cplx)
    L.check(dev.lib().fh_gather_comps(n, len(perm), L.intarr(perm), dev.ptr(x), dev.ptr(out)))
    return out


def mul21(A, x, D, n, K):
    """y[i,k] = sum_j A[i,j] x[j,k] per point; A: (D*D*n), x: (D*K*n)."""
    ac, xc = dev.is_complex(A), dev.is_complex(x)
    y = dev.empty((D*K*n,), complex_=ac or xc)
    L.check(dev.lib().fh_mul21(int(D), int(n), int(K), dev.ptr(A), int(ac), dev.ptr(x), int(xc), dev.ptr(y)))
    return y


def hadamard(a, b, n, nc, adiv, ca, bdiv, cb):
    ac, bc = dev.is_complex(a), dev.is_complex(b)
    out = dev.empty((nc*n,), complex_=ac or bc)
    L.check(dev.lib().fh_hadamard(int(n), int(nc), int(adiv), int(ca), int(bdiv), int(cb), dev.ptr(a), int(ac),
                                  dev.ptr(b), int(bc), dev.ptr(out)))
    return out


def contract_first(a, b, n, d, K):
    a = convert(a, True)
    b = convert(b, True)
    out = dev.empty((K*n,), complex_=True)
    L.check(dev.lib().fh_contract_first(int(n), int(d), int(K), dev.ptr(a), dev.ptr(b), dev.ptr(out)))
    return out


def inv_dxd(A, D, n):
    out = dev.empty(A.shape)
    L.check(dev.lib().fh_inv_dxd(int(D), int(n), dev.ptr(A), dev.ptr(out)))
    return out


def rfftn(x, N, batch):
    N = tuple(int(v) for v in N)
    X = dev.empty((int(batch),)+N[:-1]+(N[-1]//2+1,), complex_=True)
    L.check(dev.lib().fh_rfftn(dev.plan(N), dev.ptr(x), dev.ptr(X), int(batch)))
    return X


def irfftn(X, N, batch, scale):
    """X is left intact (a spectrum-sized scratch buffer is used for dim > 1)."""
    N = tuple(int(v) for v in N)
    x = dev.empty((int(batch),)+N)
    work = dev.empty(X.shape, complex_=True) if len(N) > 1 else None
    L.check(dev.lib().fh_irfftn(dev.plan(N), dev.ptr(X), dev.ptr(x), int(batch), float(scale), dev.ptr(work)))
    return x


def spec_remap(X, N, form_in, M, form_out, batch, scale, flags=0):
    """Fourier coefficients on grid N in `form_in` -> grid M in `form_out` (see fh_spec_remap)."""
    N = tuple(int(v) for v in N)
    M = tuple(int(v) for v in M)
    Xc = convert(X, True)
    shp = M[:-1]+(M[-1]//2+1,) if form_out == 'r' else M
    out = dev.empty((int(batch),)+shp, complex_=True)
    L.check(dev.lib().fh_spec_remap(len(N), L.i64arr(N), dev.form_code(form_in), L.i64arr(M),
                                    dev.form_code(form_out), int(batch), float(scale), int(flags),
                                    dev.ptr(Xc), dev.ptr(out)))
    return out


def roll(x, N, shift, batch):
    out = dev.empty(x.shape, complex_=dev.is_complex(x))
    L.check(dev.lib().fh_roll(len(N), L.i64arr(N), L.i64arr(shift), 2 if dev.is_complex(x) else 1, int(batch),
                              dev.ptr(x), dev.ptr(out)))
    return out


def _freq_args(N, Y, fft_form):
    return len(N), L.i64arr(N), L.dblarr(Y), dev.form_code(fft_form)


def grad(X, N, Y, fft_form, ncomp, nf):
    out = dev.empty((ncomp*len(N)*nf,), complex_=True)
    L.check(dev.lib().fh_grad(*_freq_args(N, Y, fft_form), int(ncomp), dev.ptr(X), dev.ptr(out)))
    return out


def div(X, N, Y, fft_form, ncomp, nf):
    out = dev.empty((ncomp*nf,), complex_=True)
    L.check(dev.lib().fh_div(*_freq_args(N, Y, fft_form), int(ncomp), dev.ptr(X), dev.ptr(out)))
    return out


def potential(X, N, Y, fft_form, ncomp, nf):
    out = dev.empty((ncomp*nf,), complex_=True)
    L.check(dev.lib().fh_potential(*_freq_args(N, Y, fft_form), int(ncomp), dev.ptr(X), dev.ptr(out)))
    return out
