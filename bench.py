#!/usr/bin/env python
"""bench.py — Fourier-Galerkin CG throughput on B200 (BASELINE.json metric: CG iterations/s and
voxel-DOF/s at 256^3 elasticity; % of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 256] [--impl ours|reference]

A "step" is ONE CG iteration of the hot path (operator application G·A·p, two dot products, three
vector updates) on the BASELINE config-3 workload: 3-D linear elasticity (D = 6, Mandel), random
two-phase microstructure (seed 20240901, 30 % inclusions, K/G = 1/1 | 10/5), GaNi, n^3 grid.
N > 1 GPUs (torchrun, one rank per GPU) run ONE slab-decomposed solve of BASELINE config 4 — the same
generator at 512^3, x-planes split over the ranks, FFT transposes over NVLink (ffthompy_b200/slab.py),
CG scalars by all-reduce; `--mode replicas` runs N independent 256^3 solves instead.

`--impl reference` times the reference's own CPU implementation of the path on the host cores: the UNMODIFIED
reference tree when it is present under baseline/_ref/ (a git-ignored copy made by __graft_entry__.build() where
/root/reference exists; kind "reference"), else the oracle restatement (oracle/, kind "port") -- on a bounded
sample (128^3) of the same generator; the line's config.grid is the grid that was actually timed.

Extra keys of the N = 1 line (task 4): `coefficient_modes` (the same CG step with the coefficients
streamed as a symmetric array and as a full DxD array instead of the 1-byte phase table), `config2_255` (BASELINE
config 2 shape: 3-D scalar, Ga grid 255^3), `strong_base` (the 512^3 solve on ONE GPU: the strong-scaling base of
the N > 1 lines, with kit and A_H[0,0] for the 1 <-> N agreement check), `solution` (kit, A_H[0,0] of the timed grid).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20240901
CPU_SAMPLE_N = 128          # the CPU path is timed at 128^3 (3.8 s / iteration, 5 GB)
REF_TREE = os.path.join(ROOT, 'baseline', '_ref')


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs"""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu='+self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def elastic_mandel(bulk, mu):
    I = np.zeros((6, 6))
    I[:3, :3] = 1.
    return 3*bulk*(I/3.)+2*mu*(np.eye(6)-I/3.)


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_cg_rate(n, iters, warm=1):
    """seconds per CG iteration of the NumPy path (materialised G^, einsum, np.fft — the reference's
    arithmetic, oracle/ffthom_oracle.py) on the same generator at n^3"""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import ffthom_oracle as O
    N = (n, n, n)
    A, _ = O.two_phase(N, SEED, 0.3, elastic_mandel(1, 1), elastic_mandel(10, 5))
    G = O.proj_elasticity(N, np.ones(3))
    G1 = G[1]+G[2]
    del G
    Afun = O.GA(A, G1, N)
    E = np.zeros((6,)+N)
    E[0] = 1.
    B = Afun(-E)
    pN = float(n**3)
    x = np.zeros_like(B)
    R = B-Afun(x)
    P = R
    rr = np.sum(R*R)/pN
    times = []
    for it in range(warm+iters):
        t0 = time.perf_counter()
        AP = Afun(P)
        alp = rr/(np.sum(P*AP)/pN)
        x = x+alp*P
        R = R-alp*AP
        rrn = np.sum(R*R)/pN
        P = R+(rrn/rr)*P
        rr = rrn
        times.append(time.perf_counter()-t0)
    return float(np.mean(times[warm:])), times[warm:]


def reference_tree():
    """the unmodified reference under baseline/_ref (copied there by __graft_entry__.build()), or None"""
    return REF_TREE if os.path.isdir(os.path.join(REF_TREE, 'ffthompy')) else None


def ref_cg_rate(n, iters, warm=1):
    """seconds per CG iteration of the UNMODIFIED reference (ffthompy Tensor / DFT / Operator from baseline/_ref,
    SURVEY App. C recipe, even-N projection per App. D.1 built from reference functions only) at n^3"""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    os.environ['FFTHOMPY_REFERENCE'] = REF_TREE
    import _refshim
    _refshim.REFERENCE = REF_TREE
    _refshim.install()
    import ffthompy.projections as rproj
    from ffthompy.tensors import Tensor as RTensor, DFT as RDFT, Operator as ROperator
    from ffthompy.trigpol import get_Nodd
    N = np.array([n, n, n])
    rng = np.random.default_rng(SEED)
    phase = (rng.random((n, n, n)) < 0.3).astype(float)
    Cm, Ci = elastic_mandel(1, 1), elastic_mandel(10, 5)
    A = RTensor(name='A', val=np.einsum('ij,...->ij...', Cm, 1-phase)+np.einsum('ij,...->ij...', Ci, phase),
                order=2, N=N, multype=21)
    if n % 2:
        Gs = rproj.elasticity(N, np.ones(3), NyqNul=True, tensor=True)
    else:
        Gs = rproj.elasticity(get_Nodd(N), np.ones(3), NyqNul=False, tensor=True, fft_form=0)[:3]
        Gs = [G.enlarge(N) for G in Gs]
        for G in Gs:
            G.set_fft_form('r')
            G.val = G.val/np.prod(G.N)
    G1 = Gs[1]+Gs[2]
    del Gs
    Afun = ROperator(name='FiGFA', mat=[[ROperator(name='G', mat=[[RDFT(inverse=True, N=N), G1, RDFT(inverse=False, N=N)]]), A]])
    EN = RTensor(name='EN', N=N, shape=(6,), Fourier=False)
    EN.set_mean(np.eye(6)[0])
    B = Afun(-EN)
    # the loop body of ffthompy/general/solver.py:123-136 with the reference's own Tensor algebra
    x = EN.zeros_like()
    R = B-Afun(x)
    P = R
    rr = R*R
    times = []
    for it in range(warm+iters):
        t0 = time.perf_counter()
        AP = Afun(P)
        alp = float(rr/(P*AP))
        x = x+alp*P
        R = R-alp*AP
        rrn = R*R
        P = R+(rrn/rr)*P
        rr = rrn
        times.append(time.perf_counter()-t0)
    return float(np.mean(times[warm:])), times[warm:]


def host_ram_gb():
    try:
        with open('/proc/meminfo') as f:
            for line in f:
                if line.startswith('MemTotal'):
                    return float(line.split()[1])/1e6
    except Exception:
        pass
    return None


def cpu_baseline(iters, warm=1):
    """(seconds per iteration, cpu_baseline dict) of the CPU path at CPU_SAMPLE_N^3: the reference itself when its tree
    travelled with the repo, else the oracle port"""
    n = CPU_SAMPLE_N
    if reference_tree():
        t_iter, _ = ref_cg_rate(n, iters, warm)
        kind, what = 'reference', ('UNMODIFIED reference (baseline/_ref: ffthompy Tensor/DFT/Operator, CG loop of '
                                   'general/solver.py:123-136)')
    else:
        t_iter, _ = cpu_cg_rate(n, iters, warm)
        kind, what = 'port', 'oracle (NumPy restatement of the reference path: rfftn / einsum / materialised G^)'
    return t_iter, {'value': 6*n**3/t_iter, 'unit': 'voxel-DOF/s', 'cores': 1, 'kind': kind,
                    'sample': '%s: %d CG iterations at %d^3 of the same generator (seed %d); numpy.fft and einsum are '
                              'single-threaded (host: %d cores, %.0f GB RAM)'
                              % (what, iters, n, SEED, os.cpu_count() or 0, host_ram_gb() or 0),
                    'ms_per_iteration': t_iter*1e3, 'grid': [n, n, n]}


REF_MAX_STEPS = 12   # 128^3 costs ~4 s per iteration on one core: the reference arm times at most this many


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = CPU_SAMPLE_N
    timed = max(1, min(args.steps, REF_MAX_STEPS))
    t_iter, cb = cpu_baseline(timed, warm=max(1, min(args.warmup, 2)))
    val = 6*n**3/t_iter
    line = {
        'impl': 'reference', 'metric': 'cg_voxel_dof_per_s', 'value': val, 'unit': 'voxel-DOF/s',
        'cg_iterations_per_s': 1./t_iter, 'n_gpus': args.gpus, 'steps': args.steps, 'steps_timed': timed,
        'warmup': args.warmup, 'ms_per_step': t_iter*1e3, 'higher_is_better': True,
        'scaling': 'strong' if (args.gpus > 1 and args.mode == 'slab') else 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': ref_config(args),
        'cpu_baseline': cb,
        'e2e': {'value': val, 'unit': 'voxel-DOF/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n, ngpu, where, slab=None):
    if slab is not None:
        return {'workload': '3-D linear elasticity (D=6 Mandel), random two-phase microstructure seed %d, 30%% '
                            'inclusions, GaNi %d^3, Galerkin CG, slab-decomposed over %d GPUs (BASELINE config 4)'
                            % (SEED, n, ngpu),
                'grid': [n, n, n], 'D': 6, 'step': 'one CG iteration',
                'l2': 'working set per iteration and GPU (%.1f GB) exceeds the 126 MB L2; no flush needed'
                      % (51*8*n**3/ngpu/1e9),
                'parallelism': 'slab x%d: rank g owns x-planes [g*n/%d, (g+1)*n/%d); exchange mode %s'
                               % (ngpu, ngpu, ngpu, slab)}
    return {'workload': '3-D linear elasticity (D=6 Mandel), random two-phase microstructure seed %d, 30%% inclusions, '
                        'GaNi %d^3, Galerkin CG (BASELINE config 3)' % (SEED, n),
            'grid': [n, n, n], 'D': 6, 'step': 'one CG iteration',
            'l2': 'working set per iteration (%.1f GB) exceeds the 126 MB L2; no flush needed' % (51*8*n**3/1e9),
            'parallelism': 'replicas x%d (independent solves, no collective)' % ngpu if ngpu > 1 else 'single GPU'}


def ref_config(args):
    """the reference arm states the grid it actually times (`grid`, a bounded sample) next to the grid of the GPU arm
    at the same N (`gpu_arm_grid`: 256^3 single GPU; the 512^3 solve for N > 1)"""
    target = args.slab_n if (args.gpus > 1 and args.mode == 'slab') else args.n
    cfg = workload_config(CPU_SAMPLE_N, 1, 'cpu')
    cfg['workload'] += ' -- CPU sample of the %d^3 GPU-arm workload (size-normalised metric, voxel-DOF/s)' % target
    cfg['gpu_arm_grid'] = [target]*3
    cfg['parallelism'] = 'reference arm: NumPy path on one host core of rank 0 (no decomposition)'
    return cfg


# ----------------------------------------------------------------------------- GPU arm
def two_phase_A(n, planes, plane0, torch, device, seed=SEED):
    """BASELINE config 3 coefficients for the x-planes [plane0, plane0+planes) of the n^3 grid, built on the device:
    A = Cm where phase == 0, Ci where phase == 1 (what `Cm*(1-phase) + Ci*phase` of SURVEY App. C evaluates to,
    entry for entry); the draw is the one-shot array of default_rng(SEED) (PCG64.advance: bit-identical)"""
    bg = np.random.PCG64(seed)
    bg.advance(plane0*n*n)
    ph = torch.from_numpy(np.random.Generator(bg).random((planes, n, n)) < 0.3).to(device)
    Cm, Ci = elastic_mandel(1, 1), elastic_mandel(10, 5)
    A = torch.empty((6, 6, planes, n, n), dtype=torch.float64, device=device)
    for i in range(6):
        for j in range(6):
            A[i, j] = torch.where(ph, float(Ci[i, j]), float(Cm[i, j]))
    return A


def build_problem(Ad, N, kind='elasticity', Nsolve=None):
    """the solve-loop operator of applications.py:58-69 over a device-resident coefficient array"""
    from ffthompy_b200.tensors import Tensor, DFT, Operator
    import ffthompy_b200.projections as proj
    N = np.array(N)
    D = int(Ad.shape[0])
    A = Tensor(name='A', val=Ad, order=2, N=N, multype=21)
    if kind == 'elasticity':
        _, G1h, G1s, _, _ = proj.elasticity(N, np.ones(3), NyqNul=True, tensor=True)
        G = G1h+G1s
    else:   # exact-integration (Ga) scalar problem: projection of the solve grid Nsolve, enlarged to N = 2 Nsolve - 1
        _, G, _ = proj.scalar(np.array(Nsolve), np.ones(3), NyqNul=True, tensor=True)
        G = G.enlarge(N)
    GN = Operator(name='G1', mat=[[DFT(name='FiN', inverse=True, N=N), G, DFT(name='FN', inverse=False, N=N)]])
    Afun = Operator(name='FiGFA', mat=[[GN, A]])
    EN = Tensor(name='EN', N=N, shape=(D,), Fourier=False)
    EN.set_mean(np.eye(D)[0])
    return A, Afun, EN


def timed_cg(f, Bd, D, N, K, W, world=1):
    """W untimed + exactly K timed iterations of the device-resident CG (fh_cg_begin / fh_cg_steps); returns
    (ms per iteration = max over ranks, launches, x, vecs)"""
    import torch
    import torch.distributed as dist
    from ffthompy_b200 import device as dev, _lib as L
    lib = dev.lib()
    nvox = int(np.prod(N))
    x = dev.zeros((D,)+tuple(N))
    vecs = dev.empty((3*D*nvox,))
    nres, done = C.c_double(), C.c_int64()
    L.check(lib.fh_cg_begin(f.handle, dev.ptr(Bd), dev.ptr(x), dev.ptr(vecs), C.byref(nres)))
    L.check(lib.fh_cg_steps(f.handle, dev.ptr(x), dev.ptr(vecs), 0.0, max(W, 3), C.byref(done), C.byref(nres), None))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = dev.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib.fh_cg_steps(f.handle, dev.ptr(x), dev.ptr(vecs), 0.0, K, C.byref(done), C.byref(nres), None))
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    assert done.value == K
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev.device())
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())/K, dev.launch_count()-launches0, x, vecs


def solve_load0(A, Afun, EN, tol=1e-6):
    """load case E = e_0 to `tol` through linear_solver (the reference's call, applications.py:75-81):
    (kit, A_H[0,0] = <A e, e>, seconds)"""
    import torch
    from ffthompy_b200.general.solver import linear_solver
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    X, info = linear_solver(solver='CG', Afun=Afun, B=Afun(-EN), x0=EN.zeros_like(),
                            par={'tol': tol, 'maxiter': 1000}, callback=None)
    e = X+EN
    ah = float(A(e)*e)
    torch.cuda.synchronize()
    return int(info['kit']), ah, time.perf_counter()-t0


def time_fn(fn, reps):
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)/reps


COEF_BYTES = {'phase': lambda D, nv: 1.*nv, 'symmetric': lambda D, nv: 8.*(D*(D+1)//2)*nv, 'full': lambda D, nv: 8.*D*D*nv}


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from ffthompy_b200 import device as dev, _lib as L
    from ffthompy_b200.tensors import Tensor
    dev.init(local)
    lib = dev.lib()
    n = args.n
    N = np.array([n, n, n])
    D = 6
    nvox = n**3
    K, W = args.steps, args.warmup
    peak, peak_src = measured_peaks()

    # ---- workload, resident in HBM (setup, untimed)
    # (replicas, N > 1: every rank its own microstructure of the same generator family)
    Ad = two_phase_A(n, n, 0, torch, dev.device(), seed=SEED+rank)
    A, Afun, EN = build_problem(Ad, N)
    B = Afun(-EN)
    f = Afun.fused()
    assert f is not None, 'the solve-loop operator was not fused'
    cfg = f.config()

    # ---- device-resident CG: warm-up, then exactly K iterations, clocks sampled during the timed region
    sampler = ClockSampler(local)
    sampler.start()
    ms_step, launches, x, vecs = timed_cg(f, B._dev(), D, N, K, W, world)
    clocks = sampler.stop()
    value = world*D*nvox/(ms_step*1e-3)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- per-stage timing of the operator pipeline (CUDA events on the launch stream) -> roofline
    F = 8.*D*nvox
    Fs = 16.*D*n*n*cfg['pitch']
    CA = COEF_BYTES[cfg['coefficients']](D, nvox)  # bytes of coefficient data S1 actually streams
    alg = {1: F+CA+Fs, 2: 2*Fs, 3: 2*Fs, 4: 2*Fs, 5: Fs+2*F}
    names = {1: 'S1 A.p + R2C (last axis)', 2: 'S2 C2C axis 1', 3: 'S3 C2C axis 0 + Green + inverse axis 0',
             4: 'S4 inverse C2C axis 1', 5: 'S5 C2R (last axis) + <p,Ap>'}
    xin = dev.empty((D,)+tuple(N))
    xin.copy_(EN._dev())
    y = dev.empty((D,)+tuple(N))
    stage_ms = {}
    for st in range(1, 6):
        stage_ms[st] = time_fn(lambda: L.check(lib.fh_ga_stage(f.handle, st, dev.ptr(xin), dev.ptr(y))), K)
    # the two kernels that exist only in their CG form: S1 with the folded vector updates (p = r + beta p and the
    # deferred x += alpha p) exactly as fh_cg_steps launches it, and the residual update with <r,r>
    r_, p_ = vecs[:D*nvox], vecs[D*nvox:2*D*nvox]
    if lib.fh_ga_can_defer_x(f.handle):
        L.check(lib.fh_ga_set_xacc(f.handle, dev.ptr(x)))
        stage_ms[6] = time_fn(lambda: L.check(lib.fh_ga_slab_stage(f.handle, 1, 0, dev.ptr(p_), dev.ptr(r_), 1, dev.ptr(y))), K)
        L.check(lib.fh_ga_set_xacc(f.handle, None))
        stage_ms[7] = time_fn(lambda: L.check(lib.fh_cgd_update_r(f.handle, dev.ptr(vecs))), K)
        alg[6] = 5*F+CA+Fs
        alg[7] = 3*F
        names[6] = 'S1 in its CG form: x += alpha p, p = r + beta p, A.p, R2C (last axis)'
        names[7] = 'U: r -= alpha Ap, <r,r>'
    cg_stages = [s_ for s_ in stage_ms if s_ != 1] if 6 in stage_ms else list(stage_ms)
    dom = max(cg_stages, key=lambda s_: stage_ms[s_])   # largest kernel of the timed CG iteration
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (profiles/), if it is the same kernel
    traffic = None
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fjs:
            tj = json.load(fjs)['dram_bytes_per_launch']
        tag = {1: 'k_fwd_last_fast', 2: 'k_c2c_fast<256, 8, 0>', 3: 'k_mid2', 4: 'k_c2c_fast<256, 8, 1>',
               5: 'k_inv_last_fast', 6: 'k_fwd_last_fast', 7: 'k_cg_update_r'}[dom]
        if n == 256:
            traffic = [v for k, v in tj.items() if tag in k][0]
    except Exception:
        traffic = None
    achieved = alg[dom]/(stage_ms[dom]*1e-3)/1e9
    B_iter = 15*F+8.*21*nvox  # 15F + C_A (symmetric-packed A) = 888 n bytes
    roofline = {'bound': 'hbm', 'kernel': names[dom], 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved/peak, 'traffic': traffic, 'peak_source': peak_src,
                'algorithmic_bytes_per_launch': alg[dom],
                'stages': {names[s]: {'ms': stage_ms[s], 'GB/s': alg[s]/(stage_ms[s]*1e-3)/1e9,
                                      'share_of_step': (stage_ms[s]/ms_step if s in cg_stages else None)}
                           for s in stage_ms},
                'stage_note': 'one CG iteration = S1 (CG form) + S2 + S3 + S4 + S5 + U + 2 scalar kernels; the plain S1 '
                              '(operator application outside CG) is timed for reference and is not part of the step',
                'iteration': {'algorithmic_bytes': B_iter, 'achieved_GB/s': B_iter/(ms_step*1e-3)/1e9,
                              'frac_of_peak': B_iter/(ms_step*1e-3)/1e9/peak,
                              'real_traffic_bytes': 10*F+8*Fs+CA,
                              'real_traffic_note': 'five spectrum passes (8 Fs) + ten field passes + coefficients as streamed'}}

    # ---- cuFFT comparator in the same run (torch.fft = cuFFT D2Z + Z2D, no coefficient / Green work)
    cufft_ms = time_fn(lambda: torch.fft.irfftn(torch.fft.rfftn(xin, dim=(1, 2, 3)), s=tuple(N), dim=(1, 2, 3)), 5)
    ga_ms = time_fn(lambda: L.check(lib.fh_ga_apply(f.handle, dev.ptr(xin), dev.ptr(y))), 5)
    del xin, y, vecs, x, r_, p_
    torch.cuda.empty_cache()

    # ---- the same CG step in the other two coefficient modes: the very same array streamed as
    # a symmetric DxD field (21 of 36 entries read) and as a full DxD field, instead of the 1-byte phase table
    modes = {cfg['coefficients']: {'ms_per_iteration': ms_step, 'cg_iterations_per_s': 1e3/ms_step,
                                   'frac_of_peak_iteration': B_iter/(ms_step*1e-3)/1e9/peak,
                                   'S1_cg_form_ms': stage_ms.get(6), 'coefficient_bytes_per_voxel': CA/nvox}}
    if not args.no_extras:
        for mode, code in (('symmetric', '1'), ('full', '0')):
            if mode in modes:
                continue
            os.environ['FH_AMODE'] = code
            try:
                A_m, Afun_m, EN_m = build_problem(Ad, N)
                fm = Afun_m.fused()
                assert fm.config()['coefficients'] == mode, fm.config()
                ms_m, _, xm, vm = timed_cg(fm, B._dev(), D, N, K, W)
                CAm = COEF_BYTES[mode](D, nvox)
                L.check(lib.fh_ga_set_xacc(fm.handle, dev.ptr(xm)))
                ym = dev.empty((D,)+tuple(N))
                s1 = time_fn(lambda: L.check(lib.fh_ga_slab_stage(fm.handle, 1, 0, dev.ptr(vm[D*nvox:2*D*nvox]),
                                                                  dev.ptr(vm[:D*nvox]), 1, dev.ptr(ym))), K)
                L.check(lib.fh_ga_set_xacc(fm.handle, None))
                modes[mode] = {'ms_per_iteration': ms_m, 'cg_iterations_per_s': 1e3/ms_m,
                               'frac_of_peak_iteration': B_iter/(ms_m*1e-3)/1e9/peak, 'S1_cg_form_ms': s1,
                               'S1_cg_form_GB/s': (5*F+CAm+Fs)/(s1*1e-3)/1e9, 'coefficient_bytes_per_voxel': CAm/nvox,
                               'real_traffic_frac_of_peak': (10*F+8*Fs+CAm)/(ms_m*1e-3)/1e9/peak}
                del A_m, Afun_m, EN_m, fm, xm, vm, ym
            finally:
                os.environ.pop('FH_AMODE', None)
            torch.cuda.empty_cache()

    # ---- solution check of the timed grid: load e_0 to 1e-6 (kit, A_H[0,0]); device-resident inputs
    kit0, ah0, _ = solve_load0(A, Afun, EN)
    solution = {'grid': [n]*3, 'kit': kit0, 'A_H00': ah0, 'tol': 1e-6}

    # ---- end to end through the operator API with HOST buffers: upload A (pinned), solve to 1e-6, download x
    from ffthompy_b200.general.solver import linear_solver
    A_host = torch.empty(Ad.shape, dtype=torch.float64, pin_memory=True)
    A_host.copy_(Ad)
    del A, Afun, f, Ad, B, EN
    torch.cuda.empty_cache()
    A_np = A_host.numpy()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    A2, Afun2, EN2 = build_problem(A_np, N)
    A2._dev()                                    # host -> device copy of the coefficients (first use would do it)
    torch.cuda.synchronize()
    t_up = time.perf_counter()
    B2 = Afun2(-EN2)                             # operator set-up (coefficient analysis, workspace) + right-hand side
    torch.cuda.synchronize()
    t_set = time.perf_counter()
    X, info = linear_solver(solver='CG', Afun=Afun2, B=B2, x0=EN2.zeros_like(),
                            par={'tol': 1e-6, 'maxiter': 1000}, callback=None)
    torch.cuda.synchronize()
    t_sol = time.perf_counter()
    x_host = X.val
    t_e2e = time.perf_counter()-t0
    kit = info['kit']
    e2e_parts = {'upload_A_s': t_up-t0, 'operator_setup_and_rhs_s': t_set-t_up, 'solve_s': t_sol-t_set,
                 'download_x_s': t0+t_e2e-t_sol}
    e2e = {'value': D*nvox*kit/t_e2e, 'unit': 'voxel-DOF/s', 'h2d_bytes_per_step': int(A_np.nbytes/kit),
           'd2h_bytes_per_step': int(x_host.nbytes/kit), 'cg_iterations': kit, 'seconds': t_e2e, 'breakdown': e2e_parts,
           'what': 'linear_solver(CG, tol 1e-6) through ffthompy_b200 Tensor/Operator API: pinned-host A uploaded, '
                   'solution downloaded, all inside the timed region'}
    del A2, Afun2, EN2, X, A_host, A_np, B2
    torch.cuda.empty_cache()

    extras = {}
    if not args.no_extras and world == 1:
        extras['config1_2d_31'] = bench_config1(torch, dev, K, W)
        extras['config2_255'] = bench_config2(torch, dev, L, lib, K, W, peak)
        torch.cuda.empty_cache()
        extras['strong_base'] = bench_512_single(torch, dev, K, W, peak)
        torch.cuda.empty_cache()

    # ---- CPU baseline on a bounded sample (rank 0 only)
    _, cpu = cpu_baseline(3, warm=1)

    line = {'metric': 'cg_voxel_dof_per_s', 'value': value, 'unit': 'voxel-DOF/s',
            'cg_iterations_per_s': world*1e3/ms_step, 'n_gpus': world, 'steps': K, 'warmup': max(W, 3),
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(n, world, 'gpu'), 'clocks': clocks, 'e2e': e2e,
            'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu,
            'solution': solution, 'coefficient_modes': modes,
            'comparators': {'cufft_rfftn_irfftn_ms': cufft_ms, 'fused_operator_ms': ga_ms,
                            'note': 'cuFFT (torch.fft) forward+inverse of the same (6,n,n,n) field, no A.p / Green / '
                                    'dot work, vs the whole fused operator G.A.p'},
            'kernels': {'axis_kernels': {k: cfg[k] for k in ('last', 'mid1', 'mid0')}, 'spectrum_pitch': cfg['pitch'],
                        'coefficient_mode': cfg['coefficients'], 'phases': cfg['nphase']}}
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def bench_config1(torch, dev, K, W):
    """BASELINE config 1 in shape (tutorials/02_homogenisation.py: 2-D scalar, square inclusion, N = 31 x 31, GaNi):
    microseconds per CG iteration of the device loop.  At this size the iteration is launch-latency bound (8 kernels
    and one 8-byte read-back), not bandwidth bound; the number is reported for completeness, not against a roofline."""
    from ffthompy_b200.tensors import Tensor, DFT, Operator
    import ffthompy_b200.projections as proj
    n, D = 31, 2
    N = np.array([n, n])
    xs = (np.arange(n)-n//2)/n
    incl = (np.abs(xs)[:, None] < 0.3) & (np.abs(xs)[None, :] < 0.3)
    Aval = np.einsum('ij,...->ij...', np.eye(2), 1.+10.*incl)
    A = Tensor(name='A', val=Aval, order=2, N=N, multype=21)
    _, G, _ = proj.scalar(N, np.ones(2), NyqNul=True, tensor=True)
    GN = Operator(name='G1', mat=[[DFT(name='FiN', inverse=True, N=N), G, DFT(name='FN', inverse=False, N=N)]])
    Afun = Operator(name='FiGFA', mat=[[GN, A]])
    EN = Tensor(name='EN', N=N, shape=(D,), Fourier=False)
    EN.set_mean(np.eye(D)[0])
    B = Afun(-EN)
    f = Afun.fused()
    ms, launches, x, vecs = timed_cg(f, B._dev(), D, (n, n), max(K, 50), W)
    del x, vecs
    return {'workload': '2-D scalar (D=2), square inclusion, GaNi 31 x 31 (tutorial 02 shape)', 'grid': [n, n],
            'us_per_iteration': ms*1e3, 'cg_iterations_per_s': 1e3/ms, 'launches_per_iteration': launches/max(K, 50),
            'note': 'launch-latency bound: kernels + one 8-byte read-back per iteration'}


def bench_config2(torch, dev, L, lib, K, W, peak):
    """BASELINE config 2 in shape: 3-D scalar (D = 3), N = 128^3 solved on the exact-integration grid Nbar = 255^3
    (= 15*17: the compile-time odd-length kernels of csrc/fh_odd.cu), Ga projection with the reference's prod(Nbar)/prod(N) scale.  The
    coefficient field is a synthetic non-piecewise-constant isotropic field (every voxel its own value, as the
    exactly integrated coefficients of Material.get_A_Ga are), so S1 streams it as a symmetric array."""
    n, ns, D = 255, 128, 3
    N = (n, n, n)
    nvox = n**3
    g = torch.Generator(device=dev.device())
    g.manual_seed(SEED)
    a = 1.+10.*torch.rand((n, n, n), dtype=torch.float64, device=dev.device(), generator=g)
    Ad = torch.zeros((3, 3, n, n, n), dtype=torch.float64, device=dev.device())
    for i in range(3):
        Ad[i, i] = a
    del a
    A, Afun, EN = build_problem(Ad, N, kind='scalar_Ga', Nsolve=(ns, ns, ns))
    B = Afun(-EN)
    f = Afun.fused()
    cfg = f.config()
    ms, _, x, vecs = timed_cg(f, B._dev(), D, N, K, W)
    B_iter = 408.*nvox
    del x, vecs
    return {'workload': '3-D scalar (D=3), N=128^3 on the Ga grid 255^3, synthetic non-constant isotropic coefficients',
            'grid': [n]*3, 'ms_per_iteration': ms, 'cg_iterations_per_s': 1e3/ms, 'voxel_dof_per_s': D*nvox*1e3/ms,
            'algorithmic_bytes': B_iter, 'frac_of_peak_iteration': B_iter/(ms*1e-3)/1e9/peak,
            'axis_kernels': {k: cfg[k] for k in ('last', 'mid1', 'mid0')}, 'coefficient_mode': cfg['coefficients']}


def bench_512_single(torch, dev, K, W, peak):
    """BASELINE config 4's grid on ONE GPU: the strong-scaling base of the N > 1 lines and the 1 <-> N agreement check
    (kit, A_H[0,0] of load e_0 at tol 1e-6)"""
    n, D = 512, 6
    free, _ = torch.cuda.mem_get_info()
    if free < 110e9:
        return {'skipped': 'needs ~100 GB of free HBM, %.0f GB available' % (free/1e9)}
    N = (n, n, n)
    nvox = float(n)**3
    Ad = two_phase_A(n, n, 0, torch, dev.device())
    A, Afun, EN = build_problem(Ad, N)
    B = Afun(-EN)
    f = Afun.fused()
    cfg = f.config()
    k = max(5, min(K, 20))
    ms, _, x, vecs = timed_cg(f, B._dev(), D, N, k, W)
    del x, vecs, B
    torch.cuda.empty_cache()
    kit, ah, sec = solve_load0(A, Afun, EN)
    B_iter = 888.*nvox
    return {'grid': [n]*3, 'ms_per_iteration': ms, 'cg_iterations_per_s': 1e3/ms, 'voxel_dof_per_s': D*nvox*1e3/ms,
            'frac_of_peak_iteration': B_iter/(ms*1e-3)/1e9/peak, 'steps': k, 'kit': kit, 'A_H00': ah, 'tol': 1e-6,
            'solve_seconds': sec, 'axis_kernels': {k_: cfg[k_] for k_ in ('last', 'mid1', 'mid0')},
            'coefficient_mode': cfg['coefficients']}


# ----------------------------------------------------------------------------- GPU arm, N > 1: slab-decomposed solve
def run_slab(args):
    """BASELINE config 4: ONE elasticity solve at slab_n^3, x-planes split over the ranks.  `value` = voxel-DOFs
    of the whole grid x CG iterations / max-over-ranks device time."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from ffthompy_b200 import device as dev
    import ffthompy_b200.projections as proj
    from ffthompy_b200.slab import SlabGA, SlabLayout, best_exchange
    dev.init(local)
    n = args.slab_n
    N = (n, n, n)
    D = 6
    nvox = float(n)**3
    K, W = args.steps, max(args.warmup, 3)
    lay = SlabLayout(N, world, rank)

    # every rank draws only its own x-planes of the one-shot array (PCG64.advance: bit-identical)
    Ad = two_phase_A(n, lay.n0l, lay.n0_off, torch, dev.device())
    _, G1h, G1s, _, _ = proj.elasticity(np.array(N), np.ones(3), NyqNul=True, tensor=True)
    G = G1h+G1s
    exchange, tuned = args.exchange, None
    if exchange is None:
        # set-up, untimed: the faster of the two NVLink exchange schemes for this grid and rank count
        exchange, tuned = best_exchange(Ad, G, N)
    op = SlabGA(Ad, G, N, exchange=exchange)
    shape = (D, lay.n0l, n, n)
    E = dev.zeros(shape)
    E[0] = -1.
    B = op.apply(E)
    del E
    x0 = dev.zeros(shape)
    st = op.cg_begin(B, x0)
    op.cg_steps(st, 0.0, W)
    sampler = ClockSampler(local)
    op.exchanged_bytes = 0
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = dev.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done = op.cg_steps(st, 0.0, K)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    assert done == K
    launches = dev.launch_count()-launches0
    clocks = sampler.stop()
    t = torch.tensor([e0.elapsed_time(e1), -float(launches)], dtype=torch.float64, device=dev.device())
    dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)          # device time, max over ranks
    dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
    ms_step = float(t[0].item())/K
    launches_all = int(-t[1].item())
    value = D*nvox*K/(float(t[0].item())*1e-3)
    sent = op.exchanged_bytes/K
    peak, peak_src = measured_peaks()

    # ---- per-stage device time, every rank inside the same stage (barrier in front of each launch)
    P = op.pitch
    F = 8.*D*nvox/world
    Fs = 16.*D*lay.n0l*n*P
    alg = {1: F+1.*nvox/world+Fs, 2: 2*Fs, 3: 2*Fs, 4: 2*Fs, 5: Fs+2*F}
    names = {1: 'S1 A.p + R2C (last axis)',
             2: 'S2 C2C axis 1, rows stored into the owners y-slab spectra over NVLink' if op.mode == 'push' else 'S2 C2C axis 1',
             3: {'peer': 'S3 C2C axis 0 + Green + inverse axis 0, fused with the exchange (remote loads/stores over NVLink)',
                 'push': 'S3 C2C axis 0 + Green + inverse axis 0, rows stored into the owners x-slab spectra over NVLink'
                 }.get(op.mode, 'S3 C2C axis 0 + Green + inverse axis 0'),
             4: 'S4 inverse C2C axis 1', 5: 'S5 C2R (last axis) + <p,Ap>'}
    stage_ms = {}
    if op.mode in ('peer', 'push', 'packed'):
        xin = dev.zeros(shape)
        xin.normal_()
        y = dev.zeros(shape)
        for stg in (1, 2, 3, 4, 5):
            ts = []
            for rep in range(4):
                dist.barrier()
                op._barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                op._stage(stg, 0, xin, None, 0, y)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            tt = torch.tensor([float(np.mean(ts[1:]))], dtype=torch.float64, device=dev.device())
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            stage_ms[stg] = float(tt.item())
        del xin, y
    B_iter = 15*8.*D*nvox+8.*21*nvox
    roofline = {'bound': 'hbm', 'peak': peak, 'unit': 'GB/s', 'peak_source': peak_src, 'traffic': None,
                'iteration': {'algorithmic_bytes': B_iter, 'achieved_GB/s_aggregate': B_iter/(ms_step*1e-3)/1e9,
                              'frac_of_aggregate_peak': B_iter/(ms_step*1e-3)/1e9/(world*peak)},
                'nvlink': {'GB_sent_per_gpu_per_iteration': sent/1e9, 'achieved_GB/s_per_gpu': sent/(ms_step*1e-3)/1e9,
                           'peak_GB/s_per_direction': 900.0, 'frac': sent/(ms_step*1e-3)/900e9}}
    if stage_ms:
        dom = max(stage_ms, key=lambda k: stage_ms[k])
        ach = alg[dom]/(stage_ms[dom]*1e-3)/1e9
        roofline.update({'kernel': names[dom], 'achieved': ach, 'frac': ach/peak, 'algorithmic_bytes_per_launch': alg[dom],
                         'stages': {names[k]: {'ms': stage_ms[k], 'GB/s_per_gpu': alg[k]/(stage_ms[k]*1e-3)/1e9,
                                               'share_of_step': stage_ms[k]/ms_step} for k in stage_ms}})
    else:
        roofline.update({'kernel': 'whole iteration (chunk-pipelined exchange: stages overlap)',
                         'achieved': B_iter/world/(ms_step*1e-3)/1e9, 'frac': B_iter/world/(ms_step*1e-3)/1e9/peak})
    mode, nchunk = op.mode, op.nchunk

    # ---- end to end with HOST buffers: pinned A slab -> device, operator set-up, solve to 1e-6, solution slab -> host
    del st, B, x0
    e2e = None
    ok = torch.ones(1, device=dev.device())
    try:
        A_host = torch.empty(Ad.shape, dtype=torch.float64, pin_memory=True)
        A_host.copy_(Ad)
    except Exception:
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    del op, Ad
    torch.cuda.empty_cache()
    if ok.item() > 0:
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        A2 = A_host.to(dev.device(), non_blocking=True)
        _, G1h, G1s, _, _ = proj.elasticity(np.array(N), np.ones(3), NyqNul=True, tensor=True)
        op2 = SlabGA(A2, G1h+G1s, N, exchange=exchange)
        E2 = dev.zeros(shape)
        E2[0] = -1.
        X, info = op2.cg(op2.apply(E2), dev.zeros(shape), tol=1e-6, maxiter=1000)
        x_host = dev.download(X)                  # fresh pageable host array, as Tensor.val hands out (csrc/fh_host.cu)
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter()-t0], dtype=torch.float64, device=dev.device())
        # A_H[0,0] = <A e, e>, e = X + e_0 (postprocess.py:53-70), for the 1 <-> N GPU agreement check
        from ffthompy_b200 import ops
        e_ = X.clone()
        e_[0] += 1.
        ah00 = op2.dot(ops.mul21(A2, e_, D, lay.n0l*n*n, 1), e_)
        del e_
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
        kit = info['kit']
        e2e = {'value': D*nvox*kit/t_e2e, 'unit': 'voxel-DOF/s',
               'h2d_bytes_per_step': int(world*A_host.numel()*8/kit), 'd2h_bytes_per_step': int(world*x_host.size*8/kit),
               'cg_iterations': kit, 'seconds': t_e2e, 'A_H00': ah00,
               'what': 'SlabGA(A_slab).cg(tol 1e-6) on every rank: pinned-host coefficient slab uploaded, operator '
                       'set-up (symmetric-memory rendezvous included), solve, solution slab downloaded; wall clock, '
                       'max over ranks'}
        del op2, A2, X
    if rank == 0:
        line = {'metric': 'cg_voxel_dof_per_s', 'value': value, 'unit': 'voxel-DOF/s',
                'cg_iterations_per_s': 1e3/ms_step, 'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms_step,
                'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': workload_config(n, world, 'gpu', slab='%s (chunks %d)' % (mode, nchunk)),
                'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches_all, 'roofline': roofline,
                'exchange_autotune_ms_per_apply': tuned,
                'solution': ({'grid': [n]*3, 'kit': e2e['cg_iterations'], 'A_H00': e2e['A_H00'], 'tol': 1e-6,
                              'compare_with': 'strong_base.kit / strong_base.A_H00 of the N = 1 line (same grid on one GPU)'}
                             if e2e else None),
                'note': 'strong scaling over N = 2, 4, 8 at the fixed 512^3 grid; the N = 1 line is the 256^3 '
                        'single-GPU workload (BASELINE config 3). value is size-normalised (voxel-DOF/s).'}
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--n', type=int, default=256)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mode', default='slab', choices=['slab', 'replicas'],
                    help='N > 1: one slab-decomposed solve (default) or N independent replicas')
    ap.add_argument('--slab-n', type=int, default=512, help='grid size of the slab-decomposed solve')
    ap.add_argument('--exchange', default=None, help='slab exchange mode (default: best available)')
    ap.add_argument('--no-extras', action='store_true', help='N = 1: skip coefficient modes, config 2 and the 512^3 base')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    elif int(os.environ.get('WORLD_SIZE', '1')) > 1 and args.mode == 'slab':
        run_slab(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
