"""CPU oracle: a NumPy restatement of FFTHomPy's Fourier-Galerkin solve loop.

TEST INFRASTRUCTURE ONLY.  Nothing under ffthompy_b200/ imports this module; it is the
checker used by tests/, by __graft_entry__.smoke() and by bench.py's CPU-baseline /
`--impl reference` legs.  Every function cites the reference routine (path:line under
the FFTHomPy tree) whose algorithm it restates; the arithmetic is NumPy's (pocketfft,
einsum), as in the reference.

Parity pin: tests/test_oracle_golden.py checks this module against fixtures generated
from the unmodified reference by oracle/make_golden.py (tests/golden/*.npz), including
the reference's own 12 golden example problems (test_results/python3/*).
"""
import itertools

import numpy as np

# ----------------------------------------------------------------------------- grids


def get_ZNl(N, fft_form='r'):
    """trigpol.py:11-23 — integer frequencies, FFT order for 'r'/0, centred for 'c'."""
    out = []
    for n in np.atleast_1d(N):
        z = np.arange(np.fix(-n/2.), np.fix(n/2.+0.5), dtype=int)
        out.append(z if fft_form == 'c' else np.fft.ifftshift(z))
    return out


def get_xil(N, Y, fft_form='r'):
    """trigpol.py:26-40 — xi = k/Y; the 'r' form keeps N[-1]//2+1 entries of the last axis."""
    xil = [z/Y[m] for m, z in enumerate(get_ZNl(N, fft_form))]
    if fft_form == 'r':
        xil[-1] = xil[-1][:int(N[-1])//2+1]
    return xil


def get_Nodd(N):
    """trigpol.py:216-218"""
    N = np.array(N, dtype=int)
    return N-((N+1) % 2)


def N_fft(N, fft_form='r'):
    """tensors/objects.py:79-83"""
    N = tuple(int(n) for n in N)
    return N[:-1]+(N[-1]//2+1,) if fft_form == 'r' else N


def mean_index(N, fft_form='r'):
    """trigpol.py:220-224"""
    if fft_form == 'c':
        return tuple(int(n)//2 for n in N)
    return tuple(0 for _ in N)


def get_coordinates(N, Y):
    """trigpol.py:55-71"""
    z = get_ZNl(N, 'c')
    return np.array(np.meshgrid(*[Y[i]*z[i]/N[i] for i in range(len(N))], indexing='ij'))


# ----------------------------------------------------------------------------- transforms


def _axes(x, N):
    return tuple(range(x.ndim-len(N), x.ndim))


def fftn(x, N, fft_form='r'):
    """tensors/fft.py:18-43 — forward transform in the convention of `fft_form`."""
    ax = _axes(x, N)
    N = tuple(int(n) for n in N)
    if fft_form == 'r':
        return np.fft.rfftn(x, s=N, axes=ax)
    X = np.fft.fftn(x, s=N, axes=ax)/np.prod(N)
    return np.fft.fftshift(X, ax) if fft_form == 'c' else X


def ifftn(X, N, fft_form='r'):
    """tensors/fft.py:25-43 — inverse transform (real part)."""
    ax = _axes(X, N)
    N = tuple(int(n) for n in N)
    if fft_form == 'r':
        return np.fft.irfftn(X, s=N, axes=ax)
    if fft_form == 'c':
        X = np.fft.ifftshift(X, ax)
    return np.fft.ifftn(X, s=N, axes=ax).real*np.prod(N)


def cfftnc(x, N):
    """tensors/fft.py:4-9 and matvecs/objects.py:784-790 — doubly centred, normalised."""
    ax = _axes(x, N)
    return np.fft.fftshift(np.fft.fftn(np.fft.ifftshift(x, ax), s=tuple(N), axes=ax), ax)/np.prod(N)


def icfftnc(X, N):
    """tensors/fft.py:11-16 and matvecs/objects.py:792-800"""
    ax = _axes(X, N)
    return np.fft.fftshift(np.fft.ifftn(np.fft.ifftshift(X, ax), s=tuple(N), axes=ax), ax).real*np.prod(N)


# ----------------------------------------------------------------------------- scalar products


def scalar_product(y, x, N, Fourier=False, fft_form='r'):
    """tensors/objects.py:618-636"""
    pN = np.prod(N)
    if Fourier:
        if fft_form == 'r':
            if N[-1] % 2 == 1:
                return (np.sum(y[..., 0]*np.conj(x[..., 0])).real
                        + 2*np.sum(y[..., 1:]*np.conj(x[..., 1:])).real)/pN**2
            return (np.sum(y[..., 0]*np.conj(x[..., 0])).real + np.sum(y[..., -1]*np.conj(x[..., -1])).real
                    + 2*np.sum(y[..., 1:-1]*np.conj(x[..., 1:-1])).real)/pN**2
        return np.sum(y*np.conj(x)).real
    return np.sum(y*x)/pN


def norm(x, N, Fourier=False, fft_form='r'):
    """tensors/objects.py:606-616 (L2)"""
    return scalar_product(x, x, N, Fourier, fft_form)**0.5


# ----------------------------------------------------------------------------- fft forms, resampling


def set_fft_form(val, N, form_in, form_out):
    """tensors/objects.py:135-167 — convert Fourier coefficients between 'r', 0, 'c'."""
    N = tuple(int(n) for n in N)
    if form_in == form_out:
        return val
    ax = _axes(val, N)
    if form_in == 'r':
        nval = np.flip(val[..., 1:].conj(), axis=-1)
        for a in ax[:-1]:
            first, rest = np.split(nval, [1], axis=a)
            nval = np.concatenate((first, np.flip(rest, axis=a)), axis=a)
        if N[-1] % 2 == 0:
            nval = nval[..., 1:]
        out = np.concatenate((val, nval), axis=-1)/np.prod(N)
        return np.fft.fftshift(out, axes=ax) if form_out == 'c' else out
    if form_in == 'c':
        out = np.fft.ifftshift(val, axes=ax)
        return out[..., :N[-1]//2+1]*np.prod(N) if form_out == 'r' else out
    # form_in == 0
    if form_out == 'c':
        return np.fft.fftshift(val, axes=ax)
    return val[..., :N[-1]//2+1]*np.prod(N)


def trigpol_enlarge(xN, M):
    """trigpol.py:162-189 — centred zero padding (no scaling)."""
    M = np.array(M, dtype=float)
    N = np.array(xN.shape, dtype=float)
    if np.allclose(M, N):
        return xN
    ibeg = np.ceil((M-N)/2).astype(int)
    iend = np.ceil((M+N)/2).astype(int)
    xM = np.zeros(M.astype(int), dtype=xN.dtype)
    xM[tuple(slice(ibeg[i], iend[i]) for i in range(N.size))] = xN
    return xM


def trigpol_decrease(xN, M):
    """trigpol.py:191-214 — centred truncation."""
    M = np.array(M, dtype=float)
    N = np.array(xN.shape, dtype=float)
    ibeg = np.fix((N-M+(M % 2))/2).astype(int)
    iend = np.fix((N+M+(M % 2))/2).astype(int)
    return xN[tuple(slice(ibeg[i], iend[i]) for i in range(N.size))]


def enlarge(val, N, M, fft_form='r'):
    """Tensor.enlarge, tensors/objects.py:428-467: zero-pad the spectrum from N to M.  Even axes get
    their Nyquist hyper-plane halved and mirrored.  For the 'r' form the round trip through 'c'
    rescales by prod(M)/prod(N) (fields keep nodal values; multipliers get scaled)."""
    N = tuple(int(n) for n in N)
    M = tuple(int(m) for m in M)
    if np.allclose(N, M):
        return val
    order = val.ndim-len(N)
    v = set_fft_form(val, N, fft_form, 'c')
    ax = tuple(range(order, v.ndim))
    for ii, a in enumerate(ax):
        if N[ii] % 2 == 0:
            n0, c = np.split(v, [1], axis=a)
            n2 = np.copy(n0)
            for jj, ac in enumerate(ax):
                if a == ac:
                    continue
                if n2.shape[ac] % 2 == 0:
                    n20, n2c = np.split(n2, [1], axis=ac)
                    n2 = np.concatenate((n20, np.flip(n2c, axis=ac)), axis=ac)
                else:
                    n2 = np.flip(n2, axis=ac)
            v = np.concatenate((0.5*n0, c, 0.5*n2.conj()), axis=a)
    Mf = np.array(M, dtype=float)
    Nc = np.array(v.shape[order:], dtype=float)
    ibeg = np.ceil((Mf-Nc)/2).astype(int)
    iend = np.ceil((Mf+Nc)/2).astype(int)
    new = np.zeros(val.shape[:order]+M, dtype=v.dtype)
    new[(slice(None),)*order+tuple(slice(ibeg[i], iend[i]) for i in range(len(N)))] = v
    return set_fft_form(new, M, 'c', fft_form)


def decrease(val, N, M, fft_form='r'):
    """Tensor.decrease, tensors/objects.py:469-486"""
    N = tuple(int(n) for n in N)
    M = tuple(int(m) for m in M)
    if np.allclose(N, M):
        return val
    order = val.ndim-len(N)
    v = set_fft_form(val, N, fft_form, 'c')
    new = np.zeros(val.shape[:order]+M, dtype=v.dtype)
    for di in np.ndindex(*val.shape[:order]):
        new[di] = trigpol_decrease(v[di], M)
    return set_fft_form(new, M, 'c', fft_form)


def project(val, N, M, Fourier=False, fft_form='r'):
    """Tensor.project, tensors/objects.py:488-511"""
    if np.allclose(N, M):
        return val
    X = val if Fourier else fftn(val, N, fft_form)
    if np.all(np.greater(M, N)):
        X = enlarge(X, N, M, fft_form)
    elif np.all(np.less(M, N)):
        X = decrease(X, N, M, fft_form)
    else:
        raise NotImplementedError()
    return X if Fourier else ifftn(X, M, fft_form)


# ----------------------------------------------------------------------------- Green operators


def _pad_form0(G, Nred, N):
    """zero-pad form-0 multipliers from the odd grid Nred to N (Tensor.enlarge on fft_form 0:
    shift, centred pad, unshift; no scaling) — projections.py:102-105,254-259."""
    if np.allclose(Nred, N):
        return G
    ax = tuple(range(G.ndim-len(N), G.ndim))
    c = np.fft.fftshift(G, axes=ax)
    out = np.zeros(G.shape[:G.ndim-len(N)]+tuple(int(n) for n in N), dtype=G.dtype)
    Mf, Nf = np.array(N, dtype=float), np.array(Nred, dtype=float)
    ibeg = np.ceil((Mf-Nf)/2).astype(int)
    iend = np.ceil((Mf+Nf)/2).astype(int)
    out[(Ellipsis,)+tuple(slice(ibeg[i], iend[i]) for i in range(len(N)))] = c
    return np.fft.ifftshift(out, axes=ax)


def _to_form(G, N, fft_form):
    """projections.py:107-110,262-265: arrays are assembled in form 0; 'r' = truncated last axis
    (the x prod(N) of set_fft_form is undone), 'c' = fftshift."""
    ax = tuple(range(G.ndim-len(N), G.ndim))
    if fft_form == 'r':
        return G[..., :int(N[-1])//2+1].copy()
    if fft_form == 'c':
        return np.fft.fftshift(G, axes=ax)
    return G


def proj_scalar(N, Y, NyqNul=True, fft_form='r'):
    """projections.py:9-112 -> (G0, G1, G2), each (d, d) + N_fft, real."""
    N = np.array(N, dtype=int)
    d = N.size
    Nred = get_Nodd(N) if NyqNul else N
    xi = get_xil(Nred, Y, 0)
    XI = np.meshgrid(*xi, indexing='ij')
    denom = sum(x**2 for x in XI)
    ic = mean_index(Nred, 0)
    denom[ic] = 1.
    G0 = np.zeros((d, d)+tuple(Nred))
    G1 = np.zeros((d, d)+tuple(Nred))
    G2 = np.zeros((d, d)+tuple(Nred))
    for m in range(d):
        G0[m, m][ic] = 1
        for n in range(d):
            G1[m, n] = XI[m]*XI[n]/denom
            G2[m, n] = (m == n)*np.ones(tuple(Nred))-G1[m, n]
            G2[m, n][ic] = 0
    out = []
    for G in (G0, G1, G2):
        if NyqNul:
            G = _pad_form0(G, Nred, N)
        out.append(_to_form(G, N, fft_form))
    return tuple(out)


def mandel_pairs(d):
    """mechanics/matcoef.py:175-186 — Mandel ordering (11,22,33,23,13,12) / (11,22,12)."""
    return [(0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1)] if d == 3 else [(0, 0), (1, 1), (0, 1)]


def proj_elasticity(N, Y, NyqNul=True, fft_form='r'):
    """projections.py:114-267 -> (G0, G1h, G1s, G2h, G2s), each (D, D) + N_fft, Mandel notation.

    The shipped routine raises for even N with NyqNul (it builds xi from N but reshapes to Nred,
    projections.py:131,148-152).  As SURVEY App. D.1 defines — and as projections.scalar does —
    the arrays are assembled on Nred = get_Nodd(N) and zero-padded to N; for odd N this is the
    shipped result."""
    N = np.array(N, dtype=int)
    d = N.size
    D = d*(d+1)//2
    Nred = get_Nodd(N) if NyqNul else N
    xi = get_xil(Nred, Y, 0)
    XI = np.meshgrid(*xi, indexing='ij')
    norm2 = sum(x**2 for x in XI)
    ic = mean_index(Nred, 0)
    nz = np.ones(tuple(Nred))
    nz[ic] = 0.  # IS0 / Lamh support: everything but the mean (projections.py:194-199)
    norm2[ic] = 1.
    norm4 = norm2**2
    pairs = mandel_pairs(d)
    w = [1. if i == j else 2**.5 for (i, j) in pairs]
    num = [[XI[i]*XI[j] for j in range(d)] for i in range(d)]
    # v = Mandel(xi (x) xi)/|xi|^2 ; G1h = v v^T (projections.py:187,216-229)
    G1h = np.zeros((D, D)+tuple(Nred))
    S = np.zeros_like(G1h)
    W = np.zeros_like(G1h)
    Lamh = np.zeros_like(G1h)
    IS0 = np.zeros_like(G1h)
    mean = np.zeros_like(G1h)
    for m, (i, j) in enumerate(pairs):
        IS0[m, m] = nz
        mean[m, m][ic] = 1
        for q, (k, l) in enumerate(pairs):
            G1h[m, q] = w[m]*w[q]*num[i][j]*num[k][l]/norm4
            # S_ijkl = (d_ik n_j n_l + d_il n_j n_k + d_jk n_i n_l + d_jl n_i n_k)/2 (projections.py:185,211-223)
            S[m, q] = w[m]*w[q]*0.5*((i == k)*num[j][l]+(i == l)*num[j][k]+(j == k)*num[i][l]+(j == l)*num[i][k])/norm2
            if m < d and q < d:
                Lamh[m, q] = nz/d
            if q < d:
                W[m, q] = w[m]*num[i][j]/norm2  # projections.py:207-208,222-223
    S = S*nz
    G1s = S-2*G1h
    G2h = 1./(d-1)*(d*Lamh+G1h-W-np.swapaxes(W, 0, 1))
    G2s = IS0-G1h-G1s-G2h
    out = []
    for G in (mean, G1h, G1s, G2h, G2s):
        if NyqNul:
            G = _pad_form0(G, Nred, N)
        out.append(_to_form(G, N, fft_form))
    return tuple(out)


def enlarge_multiplier(G, N, M):
    """applications.py:28-31,113-116: hG.enlarge(Nbar) on an 'r'-form multiplier = zero padding
    times prod(M)/prod(N)."""
    return enlarge(G, N, M, 'r').real


def green4(N, Y, kind='small_strain', fft_form='r'):
    """tensors/projection.py:33-70 — 4th-order Green tensors (d,d,d,d)+N_fft, no Nyquist zeroing."""
    N = np.array(N, dtype=int)
    d = N.size
    freq = get_xil(N, Y, fft_form)
    Q = np.meshgrid(*freq, indexing='ij')
    qq = sum(q*q for q in Q)
    qq1 = np.where(qq == 0, 1., qq)
    G = np.zeros((d,)*4+qq.shape)
    delta = lambda a, b: float(a == b)  # noqa: E731
    for i, j, k, l in itertools.product(range(d), repeat=4):
        if kind == 'small_strain':
            v = -Q[i]*Q[j]*Q[k]*Q[l]/qq1**2+.5*(delta(i, k)*Q[j]*Q[l]+delta(i, l)*Q[j]*Q[k]
                                               + delta(j, k)*Q[i]*Q[l]+delta(j, l)*Q[i]*Q[k])/qq1
        else:
            v = delta(i, k)*Q[j]*Q[l]/qq1
        G[i, j, k, l] = np.where(qq == 0, 0., v)
    return G


# ----------------------------------------------------------------------------- pointwise


def get_inverse(A):
    """trigpol.py:120-159 — Gauss-Jordan without pivoting, vectorised over voxels."""
    B = np.copy(A)
    d = A.shape[0]
    inv = np.zeros_like(A)
    for i in range(d):
        inv[i, i] = 1.
    for m in range(d):
        diag = np.copy(B[m, m])
        B[m, m] = 1.
        for n in range(m+1, d):
            B[m, n] = B[m, n]/diag
        for n in range(d):
            inv[m, n] = inv[m, n]/diag
        for k in range(m+1, d):
            f = np.copy(B[k, m])
            for l in range(d):
                B[k, l] = B[k, l]-B[m, l]*f
                inv[k, l] = inv[k, l]-inv[m, l]*f
    for m in range(d-1, -1, -1):
        for k in range(m-1, -1, -1):
            f = np.copy(B[k, m])
            for l in range(d):
                B[k, l] = B[k, l]-B[m, l]*f
                inv[k, l] = inv[k, l]-inv[m, l]*f
    return inv


def mul21(A, x):
    """tensors/objects.py:231-232,599-604"""
    return np.einsum('ij...,j...->i...', A, x)


# ----------------------------------------------------------------------------- differential operators


def grad(X, N, Y, fft_form='r'):
    """tensors/operators.py:227-259 on Fourier coefficients X (shape + N_fft) -> shape + (d,) + N_fft."""
    d = len(N)
    freq = get_xil(N, Y, fft_form)
    F = np.meshgrid(*freq, indexing='ij')
    order = X.ndim-d
    out = np.stack([2j*np.pi*F[i]*X for i in range(d)], axis=order)
    return out


def div(X, N, Y, fft_form='r'):
    """tensors/operators.py:261-288 — X: (d,) + N_fft."""
    d = len(N)
    F = np.meshgrid(*get_xil(N, Y, fft_form), indexing='ij')
    return sum(2j*np.pi*F[i]*X[i] for i in range(d))


def potential_scalar(x, N, Y, fft_form='r'):
    """tensors/operators.py:296-309 — x: (d,) + N_fft -> N_fft."""
    d = len(N)
    freq = get_xil(N, Y, fft_form)
    K = np.meshgrid(*freq, indexing='ij')
    out = np.zeros(x.shape[1:], dtype=complex)
    done = np.zeros(x.shape[1:], dtype=bool)
    for a in range(d):
        sel = (~done) & (K[a] != 0)
        out[sel] = x[a][sel]/(2j*np.pi*K[a][sel])
        done |= sel
    return out


# ----------------------------------------------------------------------------- operator and solvers


class GA(object):
    """Afun = Operator([[Operator([[FiN, G, FN]]), A]]) (applications.py:33-58, operators.py:136-144)
    with a materialised multiplier G, exactly the reference's arithmetic."""

    def __init__(self, A, G, N):
        self.A, self.G, self.N = A, G, tuple(int(n) for n in N)
        self.ax = tuple(range(1, len(self.N)+1))

    def __call__(self, x):
        s = np.einsum('ij...,j...->i...', self.A, x)
        S = np.fft.rfftn(s, s=self.N, axes=self.ax)
        S = np.einsum('ij...,j...->i...', self.G, S)
        return np.fft.irfftn(S, s=self.N, axes=self.ax)


def cg(Afun, B, x0, tol=1e-6, maxiter=1000, N=None):
    """general/solver.py:80-139 with scal = Tensor scalar product (objects.py:635).
    Returns x, info{'kit','norm_res','hist'}."""
    pN = float(np.prod(N if N is not None else B.shape[1:]))
    scal = lambda X, Y: np.sum(X*Y)/pN  # noqa: E731
    x = x0
    R = B-Afun(x0)
    P = R
    rr = scal(R, R)
    kit = 0
    norm_res = np.double(rr)**0.5
    hist = [norm_res]
    while norm_res > tol and kit < maxiter:
        kit += 1
        AP = Afun(P)
        alp = float(rr/scal(P, AP))
        x = x+alp*P
        R = R-alp*AP
        rrnext = scal(R, R)
        bet = rrnext/rr
        rr = rrnext
        P = R+bet*P
        norm_res = np.double(rr)**0.5
        hist.append(norm_res)
    if kit == 0:
        norm_res = 0
    return x, {'kit': kit, 'norm_res': norm_res, 'hist': np.array(hist)}


def richardson(Afun, B, x0, alpha, tol=1e-6, maxiter=1000, N=None):
    """general/solver.py:63-77"""
    pN = float(np.prod(N if N is not None else B.shape[1:]))
    omega = 1./alpha
    x = x0
    norm_res, kit = 1e15, 0
    while norm_res > tol and kit < maxiter:
        kit += 1
        res = B-Afun(x)
        x = x+omega*res
        norm_res = (np.sum(res*res)/pN)**0.5
    return x, {'kit': kit, 'norm_res': norm_res}


def assembly_matrix(A, sols, N):
    """postprocess.py:53-70 (solutions already on the grid of A)."""
    D = len(sols)
    pN = float(np.prod(N))
    AH = np.zeros((D, D))
    for i, j in itertools.product(range(D), repeat=2):
        AH[i, j] = np.sum(mul21(A, sols[i])*sols[j])/pN
    return AH


def homogenize(A, G, N, tol=1e-6, maxiter=1000, loads=None):
    """applications.py:60-90 + postprocess.py:41: loop over unit loads, CG, assemble A_H.
    Returns AH, list of infos, list of solutions (incl. the macroscopic load)."""
    N = tuple(int(n) for n in N)
    D = A.shape[0]
    Afun = GA(A, G, N)
    sols, infos = [], []
    for iL in (range(D) if loads is None else loads):
        E = np.zeros((D,)+N)
        E[iL] = 1.
        B = Afun(-E)
        X, info = cg(Afun, B, np.zeros_like(B), tol, maxiter, N)
        sols.append(X+E)
        infos.append(info)
    return assembly_matrix(A, sols, N), infos, sols


# ----------------------------------------------------------------------------- synthetic materials


def elastic_mandel(bulk, mu, dim=3, plane=None):
    """mechanics/matcoef.py:12-71 — isotropic stiffness in Mandel notation
    (3-D: 6x6; plane strain: the 3x3 sub-block 11,22,12)."""
    I = np.zeros((6, 6))
    I[:3, :3] = 1.
    IS = np.eye(6)
    C = 3*bulk*(I/3.)+2*mu*(IS-I/3.)
    if dim == 3:
        return C
    if plane == 'strain':
        idx = [0, 1, 5]
        return C[np.ix_(idx, idx)]
    raise NotImplementedError(plane)


def two_phase(N, seed, frac, C_matrix, C_incl):
    """SURVEY §8(d) C3/C5 generator: A = Cm*(1-phase)+Ci*phase, phase = rng.random(N) < frac."""
    rng = np.random.default_rng(seed)
    phase = (rng.random(tuple(N)) < frac).astype(float)
    return (np.einsum('ij,...->ij...', C_matrix, 1-phase)+np.einsum('ij,...->ij...', C_incl, phase)), phase
