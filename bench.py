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

`--impl reference` times the CPU restatement of the reference's NumPy path (oracle/, the reference
itself cannot travel to the GPU box) on a bounded sample of the same workload.
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
CPU_SAMPLE_N = 96


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


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    n = CPU_SAMPLE_N
    t_iter, times = cpu_cg_rate(n, max(1, args.steps), warm=max(1, min(args.warmup, 2)))
    dof = 6*n**3
    val = dof/t_iter
    line = {
        'impl': 'reference', 'metric': 'cg_voxel_dof_per_s', 'value': val, 'unit': 'voxel-DOF/s',
        'cg_iterations_per_s': 1./t_iter, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': t_iter*1e3, 'higher_is_better': True,
        'scaling': 'strong' if (args.gpus > 1 and args.mode == 'slab') else 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': ref_config(args),
        'cpu_baseline': {'value': val, 'unit': 'voxel-DOF/s', 'cores': 1, 'kind': 'port',
                         'sample': 'oracle CG iterations (NumPy rfftn/einsum, materialised G^) at %d^3 of the same '
                                   'generator; numpy.fft and einsum are single-threaded (host has %d cores)'
                                   % (n, os.cpu_count() or 0)},
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
    """the reference arm reports the config of the GPU arm at the same N (256^3 single GPU; the 512^3 solve for N > 1)"""
    if args.gpus > 1 and args.mode == 'slab':
        cfg = workload_config(args.slab_n, args.gpus, 'cpu', slab='n/a')
        cfg['parallelism'] = 'reference arm: NumPy path on the host cores of rank 0 (no decomposition)'
        return cfg
    return workload_config(args.n, 1, 'cpu')


# ----------------------------------------------------------------------------- GPU arm
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
    from ffthompy_b200.tensors import Tensor, DFT, Operator
    import ffthompy_b200.projections as proj
    from ffthompy_b200.general.solver import linear_solver
    dev.init(local)
    lib = dev.lib()
    n = args.n
    N = np.array([n, n, n])
    D = 6
    nvox = n**3
    K, W = args.steps, args.warmup

    # ---- workload, resident in HBM (setup, untimed): A = Cm (1-phase) + Ci phase
    rng = np.random.default_rng(SEED+rank)
    phase = torch.from_numpy((rng.random((n, n, n)) < 0.3)).to(dev.device()).to(torch.float64)
    Cm = torch.from_numpy(elastic_mandel(1, 1)).to(dev.device())
    Ci = torch.from_numpy(elastic_mandel(10, 5)).to(dev.device())
    Ad = (Cm[:, :, None, None, None]*(1-phase)+Ci[:, :, None, None, None]*phase).contiguous()
    del phase
    A = Tensor(name='A', val=Ad, order=2, N=N, multype=21)
    _, G1h, G1s, _, _ = proj.elasticity(N, np.ones(3), NyqNul=True, tensor=True)
    GN = Operator(name='G1', mat=[[DFT(name='FiN', inverse=True, N=N), G1h+G1s, DFT(name='FN', inverse=False, N=N)]])
    Afun = Operator(name='FiGFA', mat=[[GN, A]])
    EN = Tensor(name='EN', N=N, shape=(D,), Fourier=False)
    EN.set_mean(np.eye(D)[0])
    B = Afun(-EN)
    f = Afun.fused()
    assert f is not None, 'the solve-loop operator was not fused'
    Bd = B._dev()
    x = dev.zeros((D,)+tuple(N))
    vecs = dev.empty((3*D*nvox,))
    nres = C.c_double()
    done = C.c_int64()

    # ---- device-resident CG: warm-up, then exactly K iterations
    L.check(lib.fh_cg_begin(f.handle, dev.ptr(Bd), dev.ptr(x), dev.ptr(vecs), C.byref(nres)))
    L.check(lib.fh_cg_steps(f.handle, dev.ptr(x), dev.ptr(vecs), 0.0, max(W, 3), C.byref(done), C.byref(nres), None))
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    launches0 = dev.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib.fh_cg_steps(f.handle, dev.ptr(x), dev.ptr(vecs), 0.0, K, C.byref(done), C.byref(nres), None))
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = dev.launch_count()-launches0
    clocks = sampler.stop()
    assert done.value == K
    t = torch.tensor([ms], dtype=torch.float64, device=dev.device())
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ms_step = ms_max/K
    value = world*D*nvox*K/(ms_max*1e-3)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- per-stage timing of the operator pipeline (CUDA events on the launch stream) -> roofline
    peak, peak_src = measured_peaks()
    flags, pitch, midT = C.c_int(), C.c_int(), C.c_int()
    L.check(lib.fh_ga_config(f.handle, C.byref(flags), C.byref(pitch), C.byref(midT)))
    F = 8.*D*nvox
    Fs = 16.*D*n*n*pitch.value
    amode = (flags.value >> 4) & 3
    CA = {0: 8.*D*D*nvox, 1: 8.*21*nvox, 2: 1.*nvox}[amode]  # bytes of coefficient data S1 actually streams
    alg = {1: F+CA+Fs, 2: 2*Fs, 3: 2*Fs, 4: 2*Fs, 5: Fs+2*F}
    names = {1: 'S1 A.p + R2C (last axis)', 2: 'S2 C2C axis 1', 3: 'S3 C2C axis 0 + Green + inverse axis 0',
             4: 'S4 inverse C2C axis 1', 5: 'S5 C2R (last axis) + <p,Ap>'}
    xin = dev.empty((D,)+tuple(N))
    xin.copy_(EN._dev())
    y = dev.empty((D,)+tuple(N))
    stage_ms = {}

    def time_stage(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)/K

    for st in range(1, 6):
        stage_ms[st] = time_stage(lambda: L.check(lib.fh_ga_stage(f.handle, st, dev.ptr(xin), dev.ptr(y))))
    # the two kernels that exist only in their CG form: S1 with the folded vector updates (p = r + beta p and the
    # deferred x += alpha p) exactly as fh_cg_steps launches it, and the residual update with <r,r>
    r_, p_ = vecs[:D*nvox], vecs[D*nvox:2*D*nvox]
    if lib.fh_ga_can_defer_x(f.handle):
        L.check(lib.fh_ga_set_xacc(f.handle, dev.ptr(x)))
        stage_ms[6] = time_stage(lambda: L.check(lib.fh_ga_slab_stage(f.handle, 1, 0, dev.ptr(p_), dev.ptr(r_), 1, dev.ptr(y))))
        L.check(lib.fh_ga_set_xacc(f.handle, None))
        stage_ms[7] = time_stage(lambda: L.check(lib.fh_cgd_update_r(f.handle, dev.ptr(vecs))))
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
        tag = {1: 'k_fwd_last_fast', 2: 'k_c2c_fast<256, 8, 0>', 3: 'k_mid_green_pipe', 4: 'k_c2c_fast<256, 8, 1>',
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
                              'frac_of_peak': B_iter/(ms_step*1e-3)/1e9/peak}}

    # ---- cuFFT comparator in the same run (torch.fft = cuFFT D2Z + Z2D, no coefficient / Green work)
    for _ in range(2):
        torch.fft.irfftn(torch.fft.rfftn(xin, dim=(1, 2, 3)), s=tuple(N), dim=(1, 2, 3))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        torch.fft.irfftn(torch.fft.rfftn(xin, dim=(1, 2, 3)), s=tuple(N), dim=(1, 2, 3))
    e1.record()
    torch.cuda.synchronize()
    cufft_ms = e0.elapsed_time(e1)/5
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        L.check(lib.fh_ga_apply(f.handle, dev.ptr(xin), dev.ptr(y)))
    e1.record()
    torch.cuda.synchronize()
    ga_ms = e0.elapsed_time(e1)/5
    del xin, y, vecs, x

    # ---- end to end through the operator API with HOST buffers: upload A (pinned), solve to 1e-6, download x
    A_host = torch.empty(Ad.shape, dtype=torch.float64, pin_memory=True)
    A_host.copy_(Ad)
    del A, Afun, GN, f, Ad, B, Bd
    torch.cuda.empty_cache()
    A_np = A_host.numpy()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    A2 = Tensor(name='A', val=A_np, order=2, N=N, multype=21)
    _, G1h, G1s, _, _ = proj.elasticity(N, np.ones(3), NyqNul=True, tensor=True)
    GN2 = Operator(name='G1', mat=[[DFT(name='FiN', inverse=True, N=N), G1h+G1s, DFT(name='FN', inverse=False, N=N)]])
    Afun2 = Operator(name='FiGFA', mat=[[GN2, A2]])
    EN2 = Tensor(name='EN', N=N, shape=(D,), Fourier=False)
    EN2.set_mean(np.eye(D)[0])
    X, info = linear_solver(solver='CG', Afun=Afun2, B=Afun2(-EN2), x0=EN2.zeros_like(),
                            par={'tol': 1e-6, 'maxiter': 1000}, callback=None)
    x_host = X.val
    t_e2e = time.perf_counter()-t0
    kit = info['kit']
    e2e = {'value': D*nvox*kit/t_e2e, 'unit': 'voxel-DOF/s', 'h2d_bytes_per_step': int(A_np.nbytes/kit),
           'd2h_bytes_per_step': int(x_host.nbytes/kit), 'cg_iterations': kit, 'seconds': t_e2e,
           'what': 'linear_solver(CG, tol 1e-6) through ffthompy_b200 Tensor/Operator API: pinned-host A uploaded, '
                   'solution downloaded, all inside the timed region'}
    del A2, Afun2, GN2, X

    # ---- CPU baseline on a bounded sample (rank 0 only)
    t_cpu, _ = cpu_cg_rate(CPU_SAMPLE_N, 3, warm=1)
    cpu = {'value': D*CPU_SAMPLE_N**3/t_cpu, 'unit': 'voxel-DOF/s', 'cores': 1, 'kind': 'port',
           'sample': 'oracle (NumPy restatement of the reference path) CG iterations at %d^3, same generator; '
                     'numpy.fft/einsum are single-threaded (host has %d cores)' % (CPU_SAMPLE_N, os.cpu_count() or 0),
           'ms_per_iteration': t_cpu*1e3}

    line = {'metric': 'cg_voxel_dof_per_s', 'value': value, 'unit': 'voxel-DOF/s',
            'cg_iterations_per_s': world*1e3/ms_step, 'n_gpus': world, 'steps': K, 'warmup': max(W, 3),
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(n, world, 'gpu'), 'clocks': clocks, 'e2e': e2e,
            'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': cpu,
            'comparators': {'cufft_rfftn_irfftn_ms': cufft_ms, 'fused_operator_ms': ga_ms,
                            'note': 'cuFFT (torch.fft) forward+inverse of the same (6,n,n,n) field, no A.p / Green / '
                                    'dot work, vs the whole fused operator G.A.p'},
            'kernels': {'fast_axes_mask': flags.value & 7, 'spectrum_pitch': pitch.value, 'mid_T': midT.value,
                        'coefficient_mode': {0: 'full DxD array', 1: 'symmetric (upper triangle read)',
                                             2: 'phase table (1 byte/voxel)'}[(flags.value >> 4) & 3],
                        'phases': flags.value >> 8}}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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

    def make_A():
        # every rank draws only its own x-planes of the one-shot array (PCG64.advance: bit-identical)
        bg = np.random.PCG64(SEED)
        bg.advance(lay.n0_off*n*n)
        ph = np.random.Generator(bg).random((lay.n0l, n, n)) < 0.3
        phase = torch.from_numpy(ph).to(dev.device()).to(torch.float64)
        Cm = torch.from_numpy(elastic_mandel(1, 1)).to(dev.device())
        Ci = torch.from_numpy(elastic_mandel(10, 5)).to(dev.device())
        return (Cm[:, :, None, None, None]*(1-phase)+Ci[:, :, None, None, None]*phase).contiguous()

    Ad = make_A()
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
    names = {1: 'S1 A.p + R2C (last axis)', 2: 'S2 C2C axis 1',
             3: 'S3 C2C axis 0 + Green + inverse axis 0, fused with the exchange (remote loads/stores over NVLink)'
                if op.mode == 'peer' else 'S3 C2C axis 0 + Green + inverse axis 0',
             4: 'S4 inverse C2C axis 1', 5: 'S5 C2R (last axis) + <p,Ap>'}
    stage_ms = {}
    if op.mode in ('peer', 'packed'):
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
        x_host = X.cpu()
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter()-t0], dtype=torch.float64, device=dev.device())
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_e2e = float(tt.item())
        kit = info['kit']
        e2e = {'value': D*nvox*kit/t_e2e, 'unit': 'voxel-DOF/s',
               'h2d_bytes_per_step': int(world*A_host.numel()*8/kit), 'd2h_bytes_per_step': int(world*x_host.numel()*8/kit),
               'cg_iterations': kit, 'seconds': t_e2e,
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
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    elif int(os.environ.get('WORLD_SIZE', '1')) > 1 and args.mode == 'slab':
        run_slab(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
