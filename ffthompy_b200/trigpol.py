"""Grids and resampling of trigonometric polynomials — drop-in for ffthompy/trigpol.py.

Index/frequency vectors are host-side metadata (1-D, tiny); everything that touches a
field (`enlarge`, `decrease`, `get_inverse`) runs on the B200 through the C ABI.
"""
import numpy as np

from . import device as dev
from . import _lib as L

fft_form_default = 'r'  # real input data (trigpol.py:6)


class Grid():
    @staticmethod
    def get_ZNl(N, fft_form=fft_form_default):
        """Integer frequencies -N/2 <= k < N/2 (trigpol.py:11-23)."""
        ZNl = []
        N = np.atleast_1d(np.array(N, dtype=int))
        for m in range(N.size):
            ZNl.append(np.arange(np.fix(-N[m]/2.), np.fix(N[m]/2.+0.5), dtype=int))
        if fft_form in ['r', 0]:
            return [np.fft.ifftshift(val) for val in ZNl]
        return ZNl

    @staticmethod
    def get_xil(N, Y, fft_form=fft_form_default):
        """Discrete frequencies xi = k/Y (trigpol.py:26-40)."""
        xil = []
        for m in np.arange(np.size(N)):
            xil.append(np.arange(np.fix(-N[m]/2.), np.fix(N[m]/2.+0.5))/Y[m])
        if fft_form in ['r']:
            xil = [np.fft.ifftshift(xi) for xi in xil]
            xil[-1] = xil[-1][:int(np.fix(N[-1]/2)+1)]
        elif fft_form in [0]:
            xil = [np.fft.ifftshift(xi) for xi in xil]
        return xil

    @staticmethod
    def get_freq(N, Y, fft_form=fft_form_default):
        return Grid.get_xil(N, Y, fft_form=fft_form)

    @staticmethod
    def get_product(xi):
        xis = np.atleast_2d(xi[0])
        for ii in range(1, len(xi)):
            xis_new = np.tile(xi[ii], xis.shape[1])
            xis_old = np.repeat(xis, xi[ii].size, axis=1)
            xis = np.vstack([xis_old, xis_new])
        return xis

    @staticmethod
    def get_coordinates(N, Y):
        """Coordinates of the nodal points, Coord[i][j] = x_N^{(i,j)} (trigpol.py:55-71)."""
        d = np.size(N)
        ZNl = Grid.get_ZNl(N, fft_form='c')
        coord = np.zeros(np.hstack([d, N]))
        for ii in np.arange(d):
            x = Y[ii]*ZNl[ii]/N[ii]
            Nshape = np.ones(d, dtype=int)
            Nshape[ii] = N[ii]
            Nrep = np.copy(N)
            Nrep[ii] = 1
            coord[ii] = np.tile(np.reshape(x, Nshape), Nrep)
        return coord


def get_Nodd(N):
    """trigpol.py:216-218"""
    Nodd = N - ((N + 1) % 2)
    return Nodd


def mean_index(N, fft_form=fft_form_default):
    """trigpol.py:220-224"""
    if fft_form in [0, 'r']:
        return tuple(np.zeros_like(N, dtype=int))
    elif fft_form in ['c']:
        return tuple(np.array(np.fix(np.array(N)/2), dtype=int))


def _remap_centred(xN, M, pure_pad):
    xN = np.asarray(xN)
    N = xN.shape
    M = tuple(int(m) for m in np.array(M).ravel())
    src = dev.upload(xN.astype(np.complex128))
    out = dev.empty(M, complex_=True)
    L.check(dev.lib().fh_spec_remap(len(N), L.i64arr(N), 2, L.i64arr(M), 2, 1, 1.0, pure_pad,
                                    dev.ptr(src), dev.ptr(out)))
    res = dev.download(out)
    return res if np.iscomplexobj(xN) else res.real.astype(xN.dtype)


def enlarge(xN, M):
    """Enlarge a centred array of Fourier coefficients by zeros (trigpol.py:162-189)."""
    if np.allclose(np.array(M, dtype=float), np.array(np.shape(xN), dtype=float)):
        return xN
    return _remap_centred(xN, M, 1)


def decrease(xN, M):
    """Drop the highest frequencies of a centred array (trigpol.py:191-214)."""
    return _remap_centred(xN, M, 0)


def get_inverse(A):
    """Inverse of the coefficient matrices at all grid points, Gauss-Jordan without
    pivoting per voxel (trigpol.py:120-159) — on the device."""
    A = np.asarray(A)
    if A.shape[0] != A.shape[1]:
        raise NotImplementedError("Non-square matrix!")
    d = A.shape[0]
    n = int(np.prod(A.shape[2:]))
    src = dev.upload(A)
    out = dev.empty(A.shape)
    L.check(dev.lib().fh_inv_dxd(d, n, dev.ptr(src), dev.ptr(out)))
    return dev.download(out)
