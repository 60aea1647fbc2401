import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from ffthompy_b200 import _lib as L
lib = L.load(); L.check(lib.fh_init(0)); dev = torch.device('cuda:0')
def ptr(t): return C.c_void_p(t.data_ptr())
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps
n = int(os.environ.get('BN', '256')); N = tuple(int(os.environ.get('BN%d' % a, n)) for a in range(3)); D = int(os.environ.get('BD', '6'))
p = C.c_void_p(); L.check(lib.fh_plan_create(C.byref(p), 3, L.i64arr(N)))
x = torch.randn((D,)+N, dtype=torch.float64, device=dev); y = torch.zeros_like(x)
mode = os.environ.get('BA', 'phase')
if mode == 'rand':
    A = torch.randn((D, D)+N, dtype=torch.float64, device=dev)
elif mode == 'sym':
    A = torch.randn((D, D)+N, dtype=torch.float64, device=dev)
    A = (A+A.transpose(0, 1)).contiguous()
else:
    ph = (torch.rand(N, device=dev) < 0.3).to(torch.float64)
    C0 = torch.eye(D, dtype=torch.float64, device=dev)*2+1
    C1 = torch.eye(D, dtype=torch.float64, device=dev)*10+5
    A = (C0[:, :, None, None, None]*(1-ph)+C1[:, :, None, None, None]*ph).contiguous()
g = L.fh_green(); g.kind = 1 if D == 6 else 0; g.dim = 3
for a in range(3): g.N[a] = N[a]; g.Y[a] = 1.0; g.band[a] = (N[a]-((N[a]+1) % 2)-1)//2
g.cS, g.cH, g.scale = 1.0, -1.0, 1.0
work = torch.zeros(lib.fh_ga_work_doubles(p, D), dtype=torch.float64, device=dev)
op = C.c_void_p(); L.check(lib.fh_ga_create(C.byref(op), p, D, ptr(A), 0, C.byref(g), ptr(work)))
fl, pi, mt = C.c_int(), C.c_int(), C.c_int(); L.check(lib.fh_ga_config(op, C.byref(fl), C.byref(pi), C.byref(mt)))
nreal = N[0]*N[1]*N[2]; F = 8*D*nreal; Fs = 16*D*N[0]*N[1]*pi.value; CA = 8*D*D*nreal
CA = {'rand': CA, 'sym': 8*(D*(D+1)//2)*nreal}.get(mode, nreal)
alg = {1: F+CA+Fs, 2: 2*Fs, 3: 2*Fs, 4: 2*Fs, 5: Fs+2*F}
out = ['N=%s A=%s cfg flags=%d pitch=%d midT=%d env=%s' % (N, mode, fl.value, pi.value, mt.value, {k: v for k, v in os.environ.items() if k.startswith('FH_')})]
for st in range(1, 6):
    t = timeit(lambda: L.check(lib.fh_ga_stage(op, st, ptr(x), ptr(y))))
    out.append('S%d %.3f ms %.0f GB/s' % (st, t, alg[st]/t/1e6))
if lib.fh_ga_stage(op, 6, ptr(x), ptr(y)) == 0:
    t = timeit(lambda: L.check(lib.fh_ga_stage(op, 6, ptr(x), ptr(y))))
    out.append('S2-4 chunked %.3f ms' % t)
t = timeit(lambda: L.check(lib.fh_ga_apply(op, ptr(x), ptr(y))))
out.append('apply %.3f ms' % t)
B = torch.randn((D,)+N, dtype=torch.float64, device=dev); xs = torch.zeros_like(B)
vecs = torch.zeros(3*D*nreal, dtype=torch.float64, device=dev); nr = C.c_double(); done = C.c_int64()
L.check(lib.fh_cg_begin(op, ptr(B), ptr(xs), ptr(vecs), C.byref(nr)))
L.check(lib.fh_cg_steps(op, ptr(xs), ptr(vecs), 0.0, 3, C.byref(done), C.byref(nr), None))
torch.cuda.synchronize(); t0 = time.perf_counter()
L.check(lib.fh_cg_steps(op, ptr(xs), ptr(vecs), 0.0, 20, C.byref(done), C.byref(nr), None))
torch.cuda.synchronize(); t = (time.perf_counter()-t0)/20*1e3
out.append('CG %.3f ms/it %.1f it/s' % (t, 1e3/t))
if lib.fh_ga_can_defer_x(op):
    r_, p_ = vecs[:D*nreal], vecs[D*nreal:2*D*nreal]
    L.check(lib.fh_ga_set_xacc(op, ptr(xs)))
    t = timeit(lambda: L.check(lib.fh_ga_slab_stage(op, 1, 0, ptr(p_), ptr(r_), 1, ptr(y))))
    L.check(lib.fh_ga_set_xacc(op, None))
    out.append('S1cg %.3f ms %.0f GB/s' % (t, (5*F+CA+Fs)/t/1e6))
    t = timeit(lambda: L.check(lib.fh_cgd_update_r(op, ptr(vecs))))
    out.append('U %.3f ms %.0f GB/s' % (t, 3*F/t/1e6))
print(' | '.join(out))
