"""Host-side restatement of the index arithmetic of the compile-time odd-length kernels (csrc/fh_odd.cu, N = 255 = 15 x 17)
and of the three-pass last-axis kernels (csrc/fh_reg3.cuh, N = 512 = 8 x 8 x 8): the same row / frequency maps in NumPy,
checked against numpy.fft.  No GPU: this pins the algebra the kernels implement (which butterfly reads which rows, where a
frequency lands, how two real lines share one complex transform for an odd length)."""
import numpy as np
import pytest


def dft(x, axis=0, inverse=False):
    n = x.shape[axis]
    k = np.arange(n)
    W = np.exp((2j if inverse else -2j)*np.pi*np.outer(k, k)/n)
    return np.tensordot(W, x, axes=([1], [axis])) if axis == 0 else np.moveaxis(np.tensordot(W, x, axes=([1], [axis])), 0, axis)


@pytest.mark.parametrize('N,R1,R2', [(255, 15, 17), (256, 16, 16), (128, 8, 16)])
def test_two_pass_stockham_maps(N, R1, R2):
    """k_c2c_fast: pass 1 a[j R1 + q] = DFT_R1(x[j + r R2]),  pass 2 X[j' + s R1] = DFT_R2(a[j' + r R1] w_N^(r j'))"""
    rng = np.random.default_rng(N)
    x = rng.standard_normal(N)+1j*rng.standard_normal(N)
    a = np.zeros(N, complex)
    for j in range(R2):
        a[j*R1:(j+1)*R1] = dft(x[j+R2*np.arange(R1)])
    w = np.exp(-2j*np.pi*np.arange(N)/N)
    X = np.zeros(N, complex)
    for jp in range(R1):
        r = np.arange(R2)
        X[jp+R1*np.arange(R2)] = dft(a[jp+r*R1]*w[(r*jp) % N])
    assert np.abs(X-np.fft.fft(x)).max() < 1e-10


@pytest.mark.parametrize('N,Ra,Rb', [(255, 15, 17), (256, 16, 16)])
def test_axis0_in_place_dif_and_its_mirror(N, Ra, Rb):
    """k_mid_green_odd / k_mid_green_pipe: F1 on rows {j + r Rb} (twiddle w_N^(jq)), F2 on rows {q Rb + s}; row q Rb + s
    then holds frequency q + Ra s; I2 / I1 are the mirrored network and give N x the input back"""
    rng = np.random.default_rng(N+1)
    x = rng.standard_normal(N)+1j*rng.standard_normal(N)
    w = np.exp(-2j*np.pi*np.arange(N)/N)
    buf = x.copy()
    for j in range(Rb):                                  # F1, in place
        rows = j+Rb*np.arange(Ra)
        y = dft(buf[rows])*w[(np.arange(Ra)*j) % N]
        buf[j+np.arange(Ra)*Rb] = y
    for q in range(Ra):                                  # F2, in place
        rows = q*Rb+np.arange(Rb)
        buf[rows] = dft(buf[rows])
    F = np.fft.fft(x)
    for row in range(N):
        q, s = divmod(row, Rb)
        assert abs(buf[row]-F[q+Ra*s]) < 1e-9
    for q in range(Ra):                                  # I2
        rows = q*Rb+np.arange(Rb)
        buf[rows] = dft(buf[rows], inverse=True)
    for j in range(Rb):                                  # I1
        rows = j+np.arange(Ra)*Rb
        buf[j+Rb*np.arange(Ra)] = dft(buf[rows]*np.conj(w[(np.arange(Ra)*j) % N]), inverse=True)
    assert np.abs(buf/N-x).max() < 1e-10


@pytest.mark.parametrize('N', [255, 15, 256])
def test_two_real_lines_per_complex_transform(N):
    """S1 / S5: Z = FFT(a + i b); X_a[k] = (Z[k] + conj Z[N-k]) / 2, X_b[k] = (Z[k] - conj Z[N-k]) / (2i) for k < nh, and back
    (the Nyquist special case exists for even N only)"""
    rng = np.random.default_rng(N+2)
    a, b = rng.standard_normal(N), rng.standard_normal(N)
    Z = np.fft.fft(a+1j*b)
    nh = N//2+1
    k = np.arange(nh)
    km = (N-k) % N
    Xa = np.empty(nh, complex)
    Xb = np.empty(nh, complex)
    Xa.real, Xa.imag = 0.5*(Z[k].real+Z[km].real), 0.5*(Z[k].imag-Z[km].imag)        # as k_fwd_last_* stores them
    Xb.real, Xb.imag = 0.5*(Z[k].imag+Z[km].imag), -0.5*(Z[k].real-Z[km].real)
    assert np.abs(Xa-np.fft.rfft(a)).max() < 1e-10 and np.abs(Xb-np.fft.rfft(b)).max() < 1e-10
    # S5: Hermitian completion of Z from the two half spectra
    Zr = np.zeros(N, complex)
    for kk in range(nh):
        av, bv = Xa[kk], Xb[kk]
        if kk == 0 or 2*kk == N:
            av, bv = av.real+0j, bv.real+0j
        Zr[kk] = complex(av.real-bv.imag, av.imag+bv.real)
        if kk > 0 and 2*kk != N:
            Zr[N-kk] = complex(av.real+bv.imag, -av.imag+bv.real)
    z = np.fft.ifft(Zr)
    assert np.abs(z.real-a).max() < 1e-10 and np.abs(z.imag-b).max() < 1e-10


def test_three_pass_positions_512():
    """fh_reg3.cuh: after the three in-place passes position q M + q2 R3 + k holds frequency q + R1 q2 + R1 R2 k"""
    N, R1, R2, R3 = 512, 8, 8, 8
    M = N//R1
    rng = np.random.default_rng(3)
    x = rng.standard_normal(N)+1j*rng.standard_normal(N)
    w = np.exp(-2j*np.pi*np.arange(N)/N)
    buf = x.copy()
    for u in range(M):                                   # pass 1
        v = dft(buf[u+M*np.arange(R1)])*w[(np.arange(R1)*u) % N]
        buf[np.arange(R1)*M+u] = v
    for q in range(R1):                                  # pass 2
        for jp in range(R3):
            b = q*M+jp
            v = dft(buf[b+R3*np.arange(R2)])*w[(R1*jp*np.arange(R2)) % N]
            buf[b+R3*np.arange(R2)] = v
    for q in range(R1):                                  # pass 3
        for q2 in range(R2):
            b = q*M+q2*R3
            buf[b:b+R3] = dft(buf[b:b+R3])
    F = np.fft.fft(x)
    for q in range(R1):
        for q2 in range(R2):
            for k in range(R3):
                assert abs(buf[q*M+q2*R3+k]-F[q+R1*q2+R1*R2*k]) < 1e-9
