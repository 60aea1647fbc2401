"""The transform conventions of ffthompy/tensors/fft.py as functions on NumPy arrays (the reference's
`Material` calls them on host arrays): the data is uploaded, transformed by the device kernels and
downloaded.  Real input for the forward transforms; the inverse ones return the real part, as the
reference does."""
import numpy as np

from .. import device as dev
from .. import ops


def _batch(x, N):
    N = tuple(int(n) for n in N)
    return N, int(np.prod(np.shape(x)[:np.ndim(x)-len(N)])) if np.ndim(x) > len(N) else 1


def _fwd(x, N, form, centred_real=False):
    N, batch = _batch(x, N)
    xd = dev.upload(np.asarray(x, dtype=float))
    if centred_real:
        xd = ops.roll(xd, N, [-(n//2) for n in N], batch)  # ifftshift
    X = ops.rfftn(xd, N, batch)
    if form != 'r':
        X = ops.spec_remap(X, N, 'r', N, form, batch, 1./float(np.prod(N)))
    shp = N[:-1]+(N[-1]//2+1,) if form == 'r' else N
    return dev.download(X).reshape(np.shape(x)[:np.ndim(x)-len(N)]+shp)


def _inv(X, N, form, centred_real=False, scale_r=None):
    N, batch = _batch(X, N)
    Xd = dev.upload(np.asarray(X, dtype=complex))
    pN = float(np.prod(N))
    if form == 'r':
        x = ops.irfftn(Xd, N, batch, 1./pN)
    else:
        H = ops.spec_remap(Xd, N, form, N, 'r', batch, 1., flags=2)
        x = ops.irfftn(H, N, batch, 1.)
    if centred_real:
        x = ops.roll(x, N, [n//2 for n in N], batch)  # fftshift
    return dev.download(x).reshape(np.shape(X)[:np.ndim(X)-len(N)]+N)


def cfftnc(x, N):
    """real and Fourier centered n-dimensional FFT (tensors/fft.py:4-9)"""
    return _fwd(x, N, 'c', centred_real=True)


def icfftnc(Fx, N):
    """real and Fourier centered n-dimensional inverse FFT (tensors/fft.py:11-16)"""
    return _inv(Fx, N, 'c', centred_real=True)


def fftnc(x, N):
    """Fourier centered FFT (tensors/fft.py:18-23)"""
    return _fwd(x, N, 'c')


def icfftn(Fx, N):
    """Fourier centered inverse FFT (tensors/fft.py:25-30)"""
    return _inv(Fx, N, 'c')


def fftn(x, N):  # normalised FFT (tensors/fft.py:33-34)
    return _fwd(x, N, 0)


def ifftn(x, N):  # normalised FFT (tensors/fft.py:36-37)
    return _inv(x, N, 0)


def rfftn(x, N):  # real-valued FFT (tensors/fft.py:39-40)
    return _fwd(x, N, 'r')


def irfftn(x, N):  # real-valued FFT (tensors/fft.py:42-43)
    return _inv(x, N, 'r')
