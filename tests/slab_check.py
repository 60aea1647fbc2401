"""Multi-GPU check of the slab-decomposed operator / CG (run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/slab_check.py [--time 256]

Compares with the CPU oracle on the same seeded inputs (operator 1e-12, CG iteration counts equal)
and optionally times CG iterations at a large grid.  A development/validation script; the pytest
entry is tests/test_gpu_slab.py."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))


def main():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import ffthom_oracle as O
    from ffthompy_b200 import device as dev
    import ffthompy_b200.projections as proj
    from ffthompy_b200.slab import SlabGA, SlabLayout
    dev.init(local)
    ok = True
    for N in ([] if '--notest' in sys.argv else [(16, 16, 12), (64, 64, 64), (64, 128, 64), (32, 24, 20), (32, 16, 15), (128, 256, 32), (128, 16, 24), (256, 32, 16), (512, 16, 16),
                                                     (16, 512, 16), (16, 16, 512), (16, 16, 1024)]):
        if N[0] % world or N[1] % world or (world > 2 and np.prod(N) > 300000):
            continue
        D = 6
        lay = SlabLayout(N, world, rank)
        Aval, _ = O.two_phase(N, 20240901, 0.3, O.elastic_mandel(1, 1), O.elastic_mandel(10, 5))
        Go = O.proj_elasticity(N, np.ones(3))
        Afo = O.GA(Aval, Go[1]+Go[2], N)
        sl = slice(lay.n0_off, lay.n0_off+lay.n0l)
        _, G1h, G1s, _, _ = proj.elasticity(np.array(N), np.ones(3))
        rng = np.random.default_rng(3)
        x = rng.standard_normal((D,)+N)
        ref = Afo(x)
        E = np.zeros((D,)+N)
        E[0] = 1.
        B = Afo(-E)
        xo, io = O.cg(Afo, B, np.zeros_like(B), 1e-6, 1000, N)
        for mode in ('packed', 'direct', 'direct1', 'peer', 'push', 'kblock') + (('p2p',) if world > 1 else ()):
            try:
                op = SlabGA(dev.upload(Aval[:, :, sl]), G1h+G1s, N, exchange=mode.rstrip('1'),
                            nchunk=(1 if mode == 'direct1' else None))
            except Exception as e:
                if 'not in the' not in str(e) and 'cannot run' not in str(e) and 'push variant' not in str(e) and 'column-block variant' not in str(e):
                    raise
                if rank == 0:
                    print('N=%s world=%d %s: exchange kernels do not cover this grid' % (N, world, mode))
                continue
            y = op.apply(dev.upload(x[:, sl]))
            err = np.abs(y.cpu().numpy()-ref[:, sl]).max()/np.abs(ref).max()
            xs, info = op.cg(dev.upload(B[:, sl]), dev.zeros((D, lay.n0l)+N[1:]), tol=1e-6, maxiter=1000)
            errx = np.abs(xs.cpu().numpy()-xo[:, sl]).max()
            good = err < 1e-12 and info['kit'] == io['kit'] and errx < 1e-9
            ok = ok and good
            if rank == 0:
                print('N=%s world=%d %s(J=%d): operator err %.2e, CG kit %d (oracle %d), solution err %.2e %s'
                      % (N, world, mode, op.nchunk, err, info['kit'], io['kit'], errx, 'ok' if good else 'FAIL'),
                      flush=True)
            del op
    def time_case(n, kind, mode):
        """CG iterations of the BASELINE generator (Bernoulli(0.3) two-phase medium, seed 20240901) at n^3;
        every rank draws only its own slab (PCG64.advance: bit-identical to the one-shot array)"""
        N = (n, n, n)
        D = 6 if kind == 'elastic' else 3
        lay = SlabLayout(N, world, rank)
        bg = np.random.PCG64(20240901)
        bg.advance(lay.n0_off*N[1]*N[2])
        full = np.random.Generator(bg).random((lay.n0l, N[1], N[2])) < 0.3
        phase = torch.from_numpy(full).to(dev.device()).to(torch.float64)
        del full
        if kind == 'elastic':
            Cm = torch.from_numpy(O.elastic_mandel(1, 1)).to(dev.device())
            Ci = torch.from_numpy(O.elastic_mandel(10, 5)).to(dev.device())
            _, G1h, G1s, _, _ = proj.elasticity(np.array(N), np.ones(3))
            G = G1h+G1s
        else:
            Cm = torch.eye(3, dtype=torch.float64, device=dev.device())
            Ci = 11.*torch.eye(3, dtype=torch.float64, device=dev.device())
            G = proj.scalar(np.array(N), np.ones(3))[1]
        Ad = (Cm[:, :, None, None, None]*(1-phase)+Ci[:, :, None, None, None]*phase).contiguous()
        del phase
        op = SlabGA(Ad, G, N, exchange=mode or None,
                    nchunk=(int(os.environ['SLAB_J']) if 'SLAB_J' in os.environ else None))
        E = dev.zeros((D, lay.n0l)+N[1:])
        E[0] = -1.
        B = op.apply(E)
        del E
        x0 = dev.zeros((D, lay.n0l)+N[1:])
        op.cg(B, x0, tol=0., maxiter=3)
        op.exchanged_bytes = 0
        K = 20
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        xs, info = op.cg(B, x0, tol=0., maxiter=K)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)*1e-3], dtype=torch.float64, device=dev.device())
        dist.all_reduce(t, op=dist.ReduceOp.MAX)       # device time, max over ranks
        if rank == 0:
            # K iterations + the initial residual = K+1 operator applications
            per = t.item()/(K+1)
            nvox = float(n)**3
            b_iter = (15*8*D+8*D*(D+1)//2)*nvox           # 15F + C_A(sym)
            nv = op.exchanged_bytes/(K+1)
            rec = {'workload': '%s %d^3' % (kind, n), 'n_gpus': world, 'exchange': op.mode, 'chunks': op.nchunk,
                   'ms_per_iteration': round(per*1e3, 3), 'it_per_s': round(1./per, 2),
                   'voxel_dof_per_s': D*nvox/per, 'hbm_roofline_frac': b_iter/per/(world*6556.8e9),
                   'nvlink_GB_sent_per_gpu_per_iteration': round(nv/1e9, 3),
                   'nvlink_frac_of_900GBps': nv/per/900e9}
            import json
            print('SLAB ' + json.dumps(rec), flush=True)
            print('mode %s J=%d: slab CG %s %d^3 on %d GPUs: %.3f ms per iteration-equivalent -> %.1f it/s, '
                  '%.3e voxel-DOF/s, %.1f%% of the aggregate HBM roofline; NVLink %.2f GB sent per GPU per iteration'
                  % (op.mode, op.nchunk, kind, n, world, per*1e3, 1./per, D*nvox/per,
                     100*rec['hbm_roofline_frac'], nv/1e9), flush=True)
        if '--profile' in sys.argv and op.mode in ('peer', 'push', 'kblock', 'p2p', 'direct'):
            xp = dev.zeros((D, lay.n0l)+N[1:])
            xp.normal_()
            marks = op.profile_apply(xp)
            if rank == 0:
                print('profile [ms] (%s, J=%d): ' % (op.mode, op.nchunk)+' | '.join('%s %.3f' % kv for kv in marks.items()),
                      flush=True)
            del xp
        if '--stages' in sys.argv and op.mode in ('peer', 'push', 'packed'):
            # per-stage device time with every rank inside the same stage (barrier in front of each launch)
            x = dev.zeros((D, lay.n0l)+N[1:])
            x.normal_()
            y = dev.zeros((D, lay.n0l)+N[1:])
            out = []
            for st in (1, 2, 3, 4, 5):
                ts = []
                for rep in range(4):
                    dist.barrier()
                    op._barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    op._stage(st, 0, x, None, 0, y)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                out.append('S%d %.3f' % (st, min(ts[1:])))
            ts = []
            for rep in range(4):
                dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                op.apply(x, y)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            out.append('apply %.3f' % min(ts[1:]))
            if rank == 0:
                print('stages [ms] (%s): ' % op.mode+' | '.join(out), flush=True)
        del op, Ad, B, x0, xs
        torch.cuda.empty_cache()

    modes = (os.environ.get('SLAB_X') or '').split(',')
    for flag, kind in (('--time', 'elastic'), ('--time-scalar', 'scalar')):
        if flag in sys.argv:
            n = int(sys.argv[sys.argv.index(flag)+1])
            for mode in modes:
                try:
                    time_case(n, kind, mode)
                except Exception as e:   # e.g. an exchange mode that does not cover this grid
                    if rank == 0:
                        print('mode %s %s %d^3: %s' % (mode, kind, n, str(e)[:200]), flush=True)
                    torch.cuda.empty_cache()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
