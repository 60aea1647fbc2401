"""Low-level GPU check of the C ABI against plain numpy (run on the GPU box:
`python tests/lowlevel_check.py`).  Prints one line per check; exit code 1 on failure.
This is a development probe — the parity tests proper are tests/test_gpu_*.py."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from ffthompy_b200 import _lib as L  # noqa: E402

lib = L.load()
L.check(lib.fh_init(0))
dev = torch.device('cuda:0')
FAIL = []


_KEEP = []


def dv(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    _KEEP.append(t)  # keep temporaries alive: the caching allocator would recycle them mid-call
    if len(_KEEP) > 64:
        torch.cuda.synchronize()
        del _KEEP[:32]
    return t


def ptr(t):
    return C.c_void_p(t.data_ptr())


def report(name, err, tol):
    ok = bool(err <= tol)
    print('%-58s err=%.3e tol=%.1e %s' % (name, err, tol, 'ok' if ok else 'FAIL'), flush=True)
    if not ok:
        FAIL.append(name)


def plan(N):
    p = C.c_void_p()
    L.check(lib.fh_plan_create(C.byref(p), len(N), L.i64arr(N)))
    return p


rng = np.random.default_rng(1)

# ---------------------------------------------------------------- FFT
for N in [(4,), (5,), (16,), (4, 4), (5, 5), (6, 4), (5, 4, 6), (9, 9, 9), (8, 8, 8), (16, 12, 10), (31, 31), (61, 61),
          (15, 15, 15), (32, 32, 32), (17, 34, 7), (127, 3, 5), (64, 64, 64), (128, 128, 128), (255, 15, 255),
          (2, 2, 512), (2, 1024, 2), (73, 2, 22)]:
    for batch in (1, 3):
        x = rng.standard_normal((batch,)+N)
        p = plan(N)
        nh = N[-1]//2+1
        X = torch.zeros((batch,)+N[:-1]+(nh,), dtype=torch.complex128, device=dev)
        xd = dv(x)
        L.check(lib.fh_rfftn(p, ptr(xd), ptr(X), batch))
        ref = np.fft.rfftn(x, s=N, axes=tuple(range(1, len(N)+1)))
        err = np.abs(X.cpu().numpy()-ref).max()/max(1.0, np.abs(ref).max())
        report('rfftn N=%s batch=%d' % (N, batch), err, 2e-15*max(4, np.log2(np.prod(N))))
        # inverse (with workspace; X must be intact afterwards)
        Xc = X.clone()
        W = torch.zeros_like(X)
        y = torch.zeros((batch,)+N, dtype=torch.float64, device=dev)
        L.check(lib.fh_irfftn(p, ptr(X), ptr(y), batch, 1.0/np.prod(N), ptr(W)))
        err = np.abs(y.cpu().numpy()-x).max()
        report('irfftn(rfftn) roundtrip N=%s batch=%d' % (N, batch), err, 4e-15*max(4, np.log2(np.prod(N))))
        report('  irfftn keeps X intact', float((X-Xc).abs().max()), 0.0)
        # non-Hermitian-consistent input: numpy semantics
        Z = rng.standard_normal(ref.shape)+1j*rng.standard_normal(ref.shape)
        Zd = dv(Z)
        L.check(lib.fh_irfftn(p, ptr(Zd), ptr(y), batch, 1.0/np.prod(N), None))
        refi = np.fft.irfftn(Z, s=N, axes=tuple(range(1, len(N)+1)))
        err = np.abs(y.cpu().numpy()-refi).max()/np.abs(refi).max()
        report('irfftn generic input N=%s batch=%d' % (N, batch), err, 4e-15*max(4, np.log2(np.prod(N))))
        lib.fh_plan_destroy(p)

# ---------------------------------------------------------------- BLAS-1 / reductions
n = 1000003
a = rng.standard_normal(n)
b = rng.standard_normal(n)
ad, bd = dv(a), dv(b)
out = torch.zeros(n, dtype=torch.float64, device=dev)
L.check(lib.fh_axpby(n, 2.5, ptr(ad), -0.5, ptr(bd), ptr(out)))
report('axpby', np.abs(out.cpu().numpy()-(2.5*a-0.5*b)).max(), 4e-15)
res = C.c_double()
L.check(lib.fh_dot(n, ptr(ad), ptr(bd), C.byref(res)))
report('dot', abs(res.value-np.dot(a, b))/np.sqrt(n), 1e-13)
L.check(lib.fh_asum(n, ptr(ad), 0, C.byref(res)))
report('asum', abs(res.value-np.abs(a).sum())/n, 1e-14)
L.check(lib.fh_amax(n, ptr(ad), 0, C.byref(res)))
report('amax', abs(res.value-np.abs(a).max()), 0.0)
sums = (C.c_double*3)()
a3 = rng.standard_normal((3, 5001))
L.check(lib.fh_sum_comp(3, 5001, ptr(dv(a3)), sums))
report('sum_comp', np.abs(np.array(sums[:])-a3.sum(axis=1)).max(), 1e-12)

# ---------------------------------------------------------------- mul21, inverse
for D in (2, 3, 6, 5):
    npt = 777
    A = rng.standard_normal((D, D, npt))
    x = rng.standard_normal((D, npt))
    y = torch.zeros((D, npt), dtype=torch.float64, device=dev)
    L.check(lib.fh_mul21(D, npt, 1, ptr(dv(A)), 0, ptr(dv(x)), 0, ptr(y)))
    report('mul21 real D=%d' % D, np.abs(y.cpu().numpy()-np.einsum('ij...,j...->i...', A, x)).max(), 1e-14)
    xc = x+1j*rng.standard_normal((D, npt))
    yc = torch.zeros((D, npt), dtype=torch.complex128, device=dev)
    L.check(lib.fh_mul21(D, npt, 1, ptr(dv(A)), 0, ptr(dv(xc)), 1, ptr(yc)))
    report('mul21 real x complex D=%d' % D, np.abs(yc.cpu().numpy()-np.einsum('ij...,j...->i...', A, xc)).max(), 1e-14)
    B = rng.standard_normal((D, D, npt))
    Cm = torch.zeros((D, D, npt), dtype=torch.float64, device=dev)
    L.check(lib.fh_mul21(D, npt, D, ptr(dv(A)), 0, ptr(dv(B)), 0, ptr(Cm)))
    report('pointwise matmul D=%d' % D, np.abs(Cm.cpu().numpy()-np.einsum('ij...,jk...->ik...', A, B)).max(), 1e-14)
for D in (2, 3, 6):
    npt = 500
    M = rng.standard_normal((D, D, npt))
    A = np.einsum('ij...,kj...->ik...', M, M)+D*np.eye(D)[:, :, None]
    Ai = torch.zeros((D, D, npt), dtype=torch.float64, device=dev)
    L.check(lib.fh_inv_dxd(D, npt, ptr(dv(A)), ptr(Ai)))
    ref = np.moveaxis(np.linalg.inv(np.moveaxis(A, -1, 0)), 0, -1)
    report('inv_dxd D=%d' % D, np.abs(Ai.cpu().numpy()-ref).max(), 1e-12)

# ---------------------------------------------------------------- fused operator vs numpy
def green_arrays(N, Y, kind, coef, band=None, scale=1.0):
    """closed form evaluated in numpy (same formulas as fh_green.cuh) -> (D,D,N_fft)"""
    d = len(N)
    ks = [np.fft.fftfreq(n, 1.0/n).round().astype(int) for n in N]
    for a in range(d):
        if N[a] % 2 == 0:
            ks[a][N[a]//2] = -N[a]//2
    ks[-1] = ks[-1][:N[-1]//2+1]
    K = np.meshgrid(*ks, indexing='ij')
    xi = [K[a]/Y[a] for a in range(d)]
    s = sum(x*x for x in xi)
    zero = s == 0
    inb = np.ones_like(zero)
    if band is not None:
        for a in range(d):
            inb &= np.abs(K[a]) <= band[a]
    s1 = np.where(zero, 1.0, s)
    n_ = [x/np.sqrt(s1) for x in xi]
    c0, cI, cS, cH, cL, cW = coef
    if kind == 0:
        D = d
        G = np.zeros((D, D)+s.shape)
        for i in range(D):
            for j in range(D):
                G[i, j] = cI*(i == j)+cH*n_[i]*n_[j]
    else:
        D = d*(d+1)//2
        pairs = [(0, 0), (1, 1), (2, 2), (1, 2), (0, 2), (0, 1)] if d == 3 else [(0, 0), (1, 1), (0, 1)]
        w = [1.0 if i == j else np.sqrt(2.0) for i, j in pairs]
        v = [w[m]*n_[i]*n_[j] for m, (i, j) in enumerate(pairs)]
        G = np.zeros((D, D)+s.shape)
        for m, (i, j) in enumerate(pairs):
            for q, (k, l) in enumerate(pairs):
                S = 0.5*((i == k)*n_[j]*n_[l]+(i == l)*n_[j]*n_[k]+(j == k)*n_[i]*n_[l]+(j == l)*n_[i]*n_[k])
                S = S*w[m]*w[q]
                lam = (1.0/d if (m < d and q < d) else 0.0)
                Wm = (v[m] if q < d else 0.0)+(v[q] if m < d else 0.0)
                G[m, q] = cI*(m == q)+cS*S+cH*v[m]*v[q]+cL*lam+cW*Wm
    G = G*scale*(inb & ~zero)
    for i in range(D):
        G[i, i][zero] = c0
    return G


def make_green(N, Y, kind, coef, band=None, scale=1.0):
    g = L.fh_green()
    g.kind = kind
    g.dim = len(N)
    for a in range(len(N)):
        g.N[a] = N[a]
        g.Y[a] = Y[a]
        g.band[a] = band[a] if band is not None else 1 << 30
    g.c0, g.cI, g.cS, g.cH, g.cL, g.cW = coef
    g.scale = scale
    return g


def np_matvec(A, G, x, N):
    ax = tuple(range(1, len(N)+1))
    s = np.einsum('ij...,j...->i...', A, x)
    S = np.fft.rfftn(s, s=N, axes=ax)
    S = np.einsum('ij...,j...->i...', G, S)
    return np.fft.irfftn(S, s=N, axes=ax)


cases = [((6, 5), (1.0, 2.0), 0, (0, 0, 0, 1, 0, 0)), ((5, 5), (1.0, 1.0), 1, (0, 0, 1, -1, 0, 0)),
         ((6, 4), (1.0, 0.5), 1, (0, 1, -1, 1, 0, 0)),
         ((6, 5, 4), (1.0, 2.0, 0.5), 0, (0, 1, 0, -1, 0, 0)), ((8, 8, 8), (1, 1, 1), 1, (0, 0, 1, -1, 0, 0)),
         ((9, 9, 9), (1, 1, 1), 1, (0, 1, -1, 1, 0, 0)), ((6, 6, 6), (1, 1, 1), 1, (0.3, 0.2, 0.5, 0.7, 1.5, -0.5)),
         ((16, 16, 16), (1, 1, 1), 1, (0, 0, 1, -1, 0, 0)), ((32, 32, 32), (1, 1, 1), 0, (0, 0, 0, 1, 0, 0)),
         ((64, 64, 64), (1, 1, 1), 1, (0, 0, 1, -1, 0, 0)), ((64, 128, 64), (1, 2, 1), 0, (0, 1, 0, -1, 0, 0)),
         ((128, 64, 256), (1, 1, 1), 1, (0, 1, -1, 1, 0, 0)), ((64, 64), (1, 1), 1, (0, 0, 1, -1, 0, 0)),
         ((128, 256), (1, 1), 0, (0, 0, 0, 1, 0, 0)), ((64, 20, 128), (1, 1, 1), 0, (0, 0, 0, 1, 0, 0)),
         ((12, 64, 64), (1, 1, 1), 1, (0, 0, 1, -1, 0, 0)),
         ((512, 16, 32), (1, 1, 1), 1, (0, 0, 1, -1, 0, 0)), ((16, 512, 64), (1, 1, 1), 0, (0, 0, 0, 1, 0, 0)),
         ((32, 32, 1024), (1, 1, 2), 1, (0, 1, -1, 1, 0, 0)), ((2048, 16, 16), (1, 1, 1), 0, (0, 1, 0, -1, 0, 0)),
         ((512, 512), (1, 1), 1, (0, 0, 1, -1, 0, 0)), ((1024, 32), (1, 1), 0, (0, 0, 0, 1, 0, 0)),
         ((32, 16), (1, 1), 1, (0, 0, 1, -1, 0, 0)),
         ((16, 32, 512), (1, 1, 1), 1, (0, 0, 1, -1, 0, 0)), ((32, 512), (1, 2), 0, (0, 0, 0, 1, 0, 0)),
         ((512, 8, 512), (1, 1, 1), 0, (0, 0, 0, 1, 0, 0)),
         # run-time-length kernels (odd / mixed radices), incl. an axis that falls back
         ((15, 15, 15), (1, 1, 1), 1, (0, 0, 1, -1, 0, 0)), ((45, 35, 63), (1, 2, 1), 0, (0, 1, 0, -1, 0, 0)),
         ((255, 15, 51), (1, 1, 1), 1, (0, 1, -1, 1, 0, 0)), ((6, 10, 12), (1, 1, 1), 1, (0, 0, 1, -1, 0, 0)),
         ((27, 25, 49), (1, 1, 1), 0, (0, 0, 0, 1, 0, 0)), ((255, 255), (1, 1), 1, (0, 0, 1, -1, 0, 0)),
         ((99, 143), (1, 1), 0, (0, 0, 0, 1, 0, 0)), ((23, 29, 31), (1, 1, 1), 1, (0, 0, 1, -1, 0, 0)),
         ((85, 64, 57), (1, 1, 1), 1, (0, 0, 1, -1, 0, 0)), ((3, 5, 7), (1, 1, 1), 0, (0, 0, 0, 1, 0, 0))]
for N, Y, kind, coef in cases:
    d = len(N)
    D = d if kind == 0 else d*(d+1)//2
    band = [(n-((n+1) % 2)-1)//2 for n in N]
    for use_band, scale in ((True, 1.0), (False, 2.5)):
        G = green_arrays(N, Y, kind, coef, band if use_band else None, scale)
        g = make_green(N, Y, kind, coef, band if use_band else None, scale)
        Gm = torch.zeros(G.shape, dtype=torch.float64, device=dev)
        L.check(lib.fh_green_materialize(C.byref(g), 1, ptr(Gm)))
        report('green materialize N=%s kind=%d band=%s' % (N, kind, use_band), np.abs(Gm.cpu().numpy()-G).max(), 5e-15)
        M = rng.standard_normal((D, D)+N)
        A = np.einsum('ij...,kj...->ik...', M, M)+np.eye(D).reshape((D, D)+(1,)*d)
        x = rng.standard_normal((D,)+N)
        p = plan(N)
        nwork = lib.fh_ga_work_doubles(p, D)
        work = torch.zeros(nwork, dtype=torch.float64, device=dev)
        Ad, xd = dv(A), dv(x)
        op = C.c_void_p()
        L.check(lib.fh_ga_create(C.byref(op), p, D, ptr(Ad), 0, C.byref(g), ptr(work)))
        y = torch.zeros((D,)+N, dtype=torch.float64, device=dev)
        L.check(lib.fh_ga_apply(op, ptr(xd), ptr(y)))
        ref = np_matvec(A, G, x, N)
        report('ga_apply N=%s kind=%d band=%s' % (N, kind, use_band), np.abs(y.cpu().numpy()-ref).max()/np.abs(ref).max(),
               1e-13)
        # unfused green apply on a spectrum
        S = np.fft.rfftn(x, s=N, axes=tuple(range(1, d+1)))
        So = torch.zeros(S.shape, dtype=torch.complex128, device=dev)
        L.check(lib.fh_green_apply(C.byref(g), 1, 1, ptr(dv(S)), ptr(So)))
        refS = np.einsum('ij...,j...->i...', G, S)
        report('green_apply N=%s kind=%d' % (N, kind), np.abs(So.cpu().numpy()-refS).max()/np.abs(refS).max(), 1e-14)
        if use_band and coef[0] == 0:
            # CG on G A x = -G A E, compare with a numpy CG (solver.py:80-139)
            E = np.zeros((D,)+N)
            E[0] = 1.0
            B = np_matvec(A, G, -E, N)
            pn = np.prod(N)
            xx = np.zeros_like(B)
            R = B-np_matvec(A, G, xx, N)
            P = R
            rr = np.sum(R*R)/pn
            kit = 0
            hist = [rr**0.5]
            tol = 1e-8
            while rr**0.5 > tol and kit < 200:
                kit += 1
                AP = np_matvec(A, G, P, N)
                alp = rr/(np.sum(P*AP)/pn)
                xx = xx+alp*P
                R = R-alp*AP
                rrn = np.sum(R*R)/pn
                P = R+(rrn/rr)*P
                rr = rrn
                hist.append(rr**0.5)
            Bd = dv(B)
            xs = torch.zeros((D,)+N, dtype=torch.float64, device=dev)
            vecs = torch.zeros(3*D*int(pn), dtype=torch.float64, device=dev)
            kk = C.c_int64()
            nr = C.c_double()
            hh = (C.c_double*256)()
            L.check(lib.fh_cg(op, ptr(Bd), ptr(xs), tol, 200, ptr(vecs), C.byref(kk), C.byref(nr), hh, 256))
            report('cg kit N=%s kind=%d (ref %d, got %d)' % (N, kind, kit, kk.value), abs(kit-kk.value), 0)
            # same iteration count; the iterates of an ill-conditioned random problem agree to ~cond * eps
            report('cg solution', np.abs(xs.cpu().numpy()-xx).max()/max(np.abs(xx).max(), 1e-300), 1e-6)
            m = min(kit, kk.value)+1
            report('cg residual history', np.max(np.abs(np.array(hh[:m])-np.array(hist[:m]))/hist[0]), 1e-8)  # the pytest bar for residual histories (tests/test_gpu_reference_suite.py)
        lib.fh_ga_destroy(op)
        lib.fh_plan_destroy(p)

# ---------------------------------------------------------------- timing at 256^3, D = 6
def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps


if '--time' in sys.argv:
    for n in (128, 256):
        N = (n, n, n)
        D = 6
        p = plan(N)
        nreal = n**3
        x = torch.randn((D,)+N, dtype=torch.float64, device=dev)
        X = torch.zeros((D, n, n, n//2+1), dtype=torch.complex128, device=dev)
        y = torch.zeros_like(x)
        t = timeit(lambda: L.check(lib.fh_rfftn(p, ptr(x), ptr(X), D)))
        F = 8*D*nreal
        print('rfftn %d^3 D=6: %.3f ms  (3 passes, %.0f GB/s algorithmic over 6F)' % (n, t, 6*F/t/1e6))
        t = timeit(lambda: L.check(lib.fh_irfftn(p, ptr(X), ptr(y), D, 1.0, None)))
        print('irfftn %d^3 D=6: %.3f ms' % (n, t))
        t = timeit(lambda: torch.fft.rfftn(x, dim=(1, 2, 3)))
        print('cuFFT rfftn (torch.fft) %d^3 D=6: %.3f ms' % (n, t))
        A = torch.randn((D, D)+N, dtype=torch.float64, device=dev)
        g = make_green(N, (1, 1, 1), 1, (0, 0, 1, -1, 0, 0), [(n-((n+1) % 2)-1)//2]*3)
        work = torch.zeros(lib.fh_ga_work_doubles(p, D), dtype=torch.float64, device=dev)
        op = C.c_void_p()
        L.check(lib.fh_ga_create(C.byref(op), p, D, ptr(A), 0, C.byref(g), ptr(work)))
        fl, pi, mt = C.c_int(), C.c_int(), C.c_int()
        L.check(lib.fh_ga_config(op, C.byref(fl), C.byref(pi), C.byref(mt)))
        print('ga config: fast flags=%d pitch=%d mid_T=%d' % (fl.value, pi.value, mt.value))
        t = timeit(lambda: L.check(lib.fh_ga_apply(op, ptr(x), ptr(y))))
        print('ga_apply %d^3 D=6: %.3f ms' % (n, t))
        Fs = 16*D*n*n*pi.value
        CA = 8*D*D*nreal
        alg = {1: F+CA+Fs, 2: 2*Fs, 3: 2*Fs, 4: 2*Fs, 5: Fs+2*F}
        for st in range(1, 6):
            t = timeit(lambda: L.check(lib.fh_ga_stage(op, st, ptr(x), ptr(y))), reps=10)
            print('  stage %d: %.3f ms -> %.0f GB/s' % (st, t, alg[st]/t/1e6))
        B = torch.randn((D,)+N, dtype=torch.float64, device=dev)
        xs = torch.zeros((D,)+N, dtype=torch.float64, device=dev)
        vecs = torch.zeros(3*D*nreal, dtype=torch.float64, device=dev)
        nr = C.c_double()
        L.check(lib.fh_cg_begin(op, ptr(B), ptr(xs), ptr(vecs), C.byref(nr)))
        done = C.c_int64()
        L.check(lib.fh_cg_steps(op, ptr(xs), ptr(vecs), 0.0, 3, C.byref(done), C.byref(nr), None))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        L.check(lib.fh_cg_steps(op, ptr(xs), ptr(vecs), 0.0, 20, C.byref(done), C.byref(nr), None))
        torch.cuda.synchronize()
        t = (time.perf_counter()-t0)/20*1e3
        print('CG iteration %d^3 D=6: %.3f ms -> %.1f it/s, %.3e DOF/s, %.0f GB/s of the 15F+C_A(sym) model'
              % (n, t, 1e3/t, D*nreal*1e3/t, (15*F+8*21*nreal)/t/1e6))
        del B, xs, vecs
        out = torch.zeros_like(x)
        t = timeit(lambda: L.check(lib.fh_axpby(D*nreal, 1.0, ptr(x), 2.0, ptr(y), ptr(out))))
        print('axpby %d^3 D=6: %.3f ms -> %.0f GB/s' % (n, t, 3*F/t/1e6))
        t = timeit(lambda: out.copy_(x))
        print('torch copy: %.3f ms -> %.0f GB/s' % (t, 2*F/t/1e6))
        lib.fh_ga_destroy(op)
        lib.fh_plan_destroy(p)
        del x, X, y, A, work, out
        torch.cuda.empty_cache()

print('FAILED: %d' % len(FAIL))
for f in FAIL:
    print('  ', f)
sys.exit(1 if FAIL else 0)
