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
    for N in ([] if '--notest' in sys.argv else [(16, 16, 12), (64, 64, 64), (64, 128, 64), (32, 24, 20), (32, 16, 15), (128, 256, 32)]):
        if N[0] % world or N[1] % world:
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
        for mode in ('packed', 'direct', 'direct1'):
            op = SlabGA(dev.upload(Aval[:, :, sl]), G1h+G1s, N, direct=(False if mode == 'packed' else None),
                        nchunk=(1 if mode == 'direct1' else None))
            if mode != 'packed' and not op.direct:
                if rank == 0:
                    print('N=%s world=%d %s: exchange kernels do not cover this grid (packed path used)' % (N, world, mode))
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
    if '--time' in sys.argv:
        n = int(sys.argv[sys.argv.index('--time')+1])
        N = (n, n, n)
        D = 6
        lay = SlabLayout(N, world, rank)
        rng = np.random.default_rng(20240901)
        full = rng.random(N) < 0.3   # same global microstructure on every rank, each keeps its slab
        phase = torch.from_numpy(full[lay.n0_off:lay.n0_off+lay.n0l]).to(dev.device()).to(torch.float64)
        del full
        Cm = torch.from_numpy(O.elastic_mandel(1, 1)).to(dev.device())
        Ci = torch.from_numpy(O.elastic_mandel(10, 5)).to(dev.device())
        Ad = (Cm[:, :, None, None, None]*(1-phase)+Ci[:, :, None, None, None]*phase).contiguous()
        _, G1h, G1s, _, _ = proj.elasticity(np.array(N), np.ones(3))
        op = SlabGA(Ad, G1h+G1s, N, direct=(False if '--packed' in sys.argv else None),
                    nchunk=(int(os.environ['SLAB_J']) if 'SLAB_J' in os.environ else None))
        E = dev.zeros((D, lay.n0l)+N[1:])
        E[0] = -1.
        B = op.apply(E)
        x0 = dev.zeros((D, lay.n0l)+N[1:])
        op.cg(B, x0, tol=0., maxiter=3)
        op.exchanged_bytes = 0
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        K = 20
        xs, info = op.cg(B, x0, tol=0., maxiter=K)
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter()-t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev.device())
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # K iterations + the initial residual = K+1 operator applications
        if rank == 0:
            per = t.item()/(K+1)
            print('mode %s J=%d: ' % ('direct' if op.direct else 'packed', op.nchunk), end='')
            print('slab CG %d^3 on %d GPUs: %.3f ms per iteration-equivalent -> %.1f it/s, %.3e voxel-DOF/s; '
                  'NVLink: %.2f GB sent per GPU per operator application'
                  % (n, world, per*1e3, 1./per, D*n**3/per, op.exchanged_bytes/(K+1)/1e9), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
