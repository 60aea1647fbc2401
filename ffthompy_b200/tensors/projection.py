"""4th-order Green tensors — drop-in for ffthompy/tensors/projection.py.  The reference fills
these arrays with per-frequency Python loops (projection.py:13-15,42-48,62-66); here one kernel
evaluates the closed forms over all frequencies."""
import numpy as np

from .. import _lib as L
from .. import device as dev
from .. import projections as _proj
from ..trigpol import fft_form_default
from .objects import Tensor


def scalar(N, Y, fft_form=fft_form_default):
    """(G0, G1, G2) without Nyquist zeroing (tensors/projection.py:6-31); G2 = I - G1 - G0."""
    G0, G1, G2 = _proj.scalar(N, Y, NyqNul=False, tensor=True, fft_form=fft_form)
    for G in (G0, G1, G2):
        G.name = 'G1'
    return G0, G1, G2


def _green4(kind, name, N, Y, fft_form):
    N = np.array(N, dtype=int)
    dim = N.size
    assert(dim == 3)
    T = Tensor(name=name, shape=(dim,)*4, N=N, Y=Y, multype=42, Fourier=True, fft_form=fft_form)
    out = dev.empty((dim,)*4+tuple(T.N_fft))
    L.check(dev.lib().fh_green4_materialize(kind, dim, L.i64arr(N), L.dblarr(np.array(Y, dtype=float)),
                                            dev.form_code(fft_form), dev.ptr(out)))
    T.val = out
    return T


def elasticity_small_strain(N, Y, fft_form=fft_form_default):
    """G_ijkl = -n_i n_j n_k n_l + (d_ik n_j n_l + d_il n_j n_k + d_jk n_i n_l + d_jl n_i n_k)/2
    (tensors/projection.py:33-51)"""
    return _green4(0, 'Ghat', N, Y, fft_form)


def elasticity_large_deformation(N, Y, fft_form=fft_form_default):
    """G_ijkl = d_ik n_j n_l (tensors/projection.py:53-70)"""
    return _green4(1, 'Ghat', N, Y, fft_form)
