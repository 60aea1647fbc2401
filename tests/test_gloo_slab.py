"""World-size-2/4 CPU tests (gloo) of the slab decomposition used for multi-GPU solves
(ffthompy_b200/slab.py): the exchange helpers (pack, all_to_all_single, unpack) and the layout /
global-frequency bookkeeping are exercised with NumPy FFTs standing in for the CUDA stages, and the
result is compared with the oracle's global operator.  The same helpers run over NCCL on the GPUs."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..'))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, N, D, kind, q):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import ffthom_oracle as O
        from ffthompy_b200.slab import (SlabLayout, exchange_fwd, exchange_bwd, allreduce_sum, direct_offsets,
                                        kblock_offsets, push_offsets)
        dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
        lay = SlabLayout(N, world, rank)
        rng = np.random.default_rng(11)
        d = 3
        # global problem (same on every rank), then this rank's slab
        if kind == 'elastic':
            G = O.proj_elasticity(N, np.ones(3))
            G = G[1]+G[2]
        else:
            G = O.proj_scalar(N, np.ones(3))[1]
        M = rng.standard_normal((D, D)+N)
        A = np.einsum('ij...,kj...->ik...', M, M)+np.eye(D).reshape((D, D, 1, 1, 1))
        x = rng.standard_normal((D,)+N)
        ref = O.GA(A, G, N)(x)
        sl = slice(lay.n0_off, lay.n0_off+lay.n0l)
        A_loc, x_loc = A[:, :, sl], x[:, sl]
        # S1 + S2 on the slab: sigma = A x, rfft along axis 2, fft along axis 1
        s = np.einsum('ij...,j...->i...', A_loc, x_loc)
        spec = np.fft.fft(np.fft.rfft(s, axis=3), axis=2)                         # [D][n0l][N1][nh]
        P = (lay.nh+7)//8*8                                                       # padded rows, as on the device
        specp = np.zeros((D, lay.n0l, N[1], P), dtype=complex)
        specp[..., :lay.nh] = spec
        specT = exchange_fwd(torch.from_numpy(specp), lay).numpy()                # [D][N0][n1l][P]
        assert specT.shape == (D, N[0], lay.n1l, P)
        # S3: fft along axis 0, global Green multiplier for this rank's k1 range, inverse fft
        Y = np.fft.fft(specT[..., :lay.nh], axis=1)
        Gloc = G[:, :, :, lay.n1_off:lay.n1_off+lay.n1l, :]
        Y = np.einsum('ij...,j...->i...', Gloc, Y)
        Y = np.fft.ifft(Y, axis=1)
        backp = np.zeros_like(specT)
        backp[..., :lay.nh] = Y
        back = exchange_bwd(torch.from_numpy(backp), lay).numpy()                 # [D][n0l][N1][P]
        y_loc = np.fft.irfft(np.fft.ifft(back[..., :lay.nh], axis=2), n=N[2], axis=3)
        err = np.abs(y_loc-ref[:, sl]).max()/np.abs(ref).max()
        # round trip of the exchange and the scalar all-reduce
        rt = exchange_bwd(exchange_fwd(torch.from_numpy(specp), lay), lay).numpy()
        err_rt = np.abs(rt-specp).max()
        # zero-copy chunk-major exchange blocks (fh_ga_slab_direct): scatter as S2 writes them, one
        # all_to_all_single per chunk, gather as S3 reads them == the packed exchange; and back
        for J in [j for j in (1, 2) if lay.n0l % j == 0]:
            off1, off0, cs1, is1, cs0, ce = direct_offsets(lay, D, P, J)
            n0c = lay.n0l//J
            bufA = np.zeros(J*ce, dtype=complex)
            cc, jj, ii, kk, tt = np.meshgrid(np.arange(D), np.arange(J), np.arange(n0c), np.arange(N[1]), np.arange(P),
                                             indexing='ij')
            idxA = jj*ce+cc*cs1+ii*is1+off1[kk]+tt
            assert np.unique(idxA).size == idxA.size == bufA.size
            bufA[idxA] = specp[cc, jj*n0c+ii, kk, tt]
            tA, tB = torch.from_numpy(bufA).view(J, ce), torch.zeros(J, ce, dtype=torch.complex128)
            for j in range(J):
                dist.all_to_all_single(tB[j], tA[j])
            c3, i3, k3, t3 = np.meshgrid(np.arange(D), np.arange(N[0]), np.arange(lay.n1l), np.arange(P), indexing='ij')
            idxB = off0[i3]+c3*cs0+k3*P+t3
            assert np.unique(idxB).size == idxB.size == bufA.size
            assert np.array_equal(tB.numpy().reshape(-1)[idxB], specT)
            tA.zero_()
            for j in range(J):
                dist.all_to_all_single(tA[j], tB[j])
            assert np.array_equal(tA.numpy().reshape(-1)[idxA], specp[cc, jj*n0c+ii, kk, tt])
        # k2-block exchange buffers (fh_ga_slab_kblock): scatter as S2 writes block b, one all_to_all_single per block,
        # gather as S3 reads it == the columns of block b of the packed exchange; and back
        for nblk in [j for j in (1, 2, 3) if j <= P//8]:
            n0l, n1l = lay.n0l, lay.n1l
            for c0, w, base, per, off1, off0 in kblock_offsets(lay, D, P, nblk):
                bufA = np.zeros(world*per, dtype=complex)
                cc, ii, kk, tt = np.meshgrid(np.arange(D), np.arange(n0l), np.arange(N[1]), np.arange(w), indexing='ij')
                idxA = cc*n0l*n1l*w+ii*n1l*w+off1[kk]+tt
                assert np.unique(idxA).size == idxA.size == bufA.size
                bufA[idxA] = specp[cc, ii, kk, c0+tt]
                tA, tB = torch.from_numpy(bufA), torch.zeros(world*per, dtype=torch.complex128)
                dist.all_to_all_single(tB, tA)                  # piece g of `per` elements -> slot `rank` on peer g
                c3, i3, k3, t3 = np.meshgrid(np.arange(D), np.arange(N[0]), np.arange(n1l), np.arange(w), indexing='ij')
                idxB = off0[i3]+c3*n0l*n1l*w+k3*w+t3
                assert np.unique(idxB).size == idxB.size == bufA.size
                assert np.array_equal(tB.numpy()[idxB], specT[..., c0:c0+w])
                dist.all_to_all_single(tA, tB)
                assert np.array_equal(tA.numpy()[idxA], specp[cc, ii, kk, c0+tt])
            assert base == D*n0l*N[1]*c0 and c0+w == P
        # push exchange (fh_ga_slab_push): every rank STORES its S2 rows into the owners' y-slab spectra.  Emulated with
        # an all_gather of the x-slab spectra: what the stores of all ranks leave in this rank's y-slab spectrum
        allspec = [torch.zeros_like(torch.from_numpy(specp)) for _ in range(world)]
        dist.all_gather(allspec, torch.from_numpy(specp))
        mineT = np.zeros(D*N[0]*lay.n1l*P, dtype=complex)
        for src in range(world):
            lsrc = SlabLayout(N, world, src)
            off1, _ = push_offsets(lsrc, P)
            rows = np.arange(lay.n1_off, lay.n1_off+lay.n1l)               # the k1 rows this rank owns
            cc, ii, kk, tt = np.meshgrid(np.arange(D), np.arange(lsrc.n0l), rows, np.arange(P), indexing='ij')
            dst = cc*N[0]*lay.n1l*P+ii*lay.n1l*P+off1[kk]+tt
            mineT[dst] = allspec[src].numpy()[cc, ii, kk, tt]
        assert np.array_equal(mineT.reshape(D, N[0], lay.n1l, P), specT)
        # ... and S3's stores back into the owners' x-slab spectra
        allT = [torch.zeros_like(torch.from_numpy(backp)) for _ in range(world)]
        dist.all_gather(allT, torch.from_numpy(backp))
        mine = np.zeros(D*lay.n0l*N[1]*P, dtype=complex)
        for src in range(world):
            lsrc = SlabLayout(N, world, src)
            _, off0 = push_offsets(lsrc, P)
            planes = np.arange(lay.n0_off, lay.n0_off+lay.n0l)
            cc, ii, kk, tt = np.meshgrid(np.arange(D), planes, np.arange(lsrc.n1l), np.arange(P), indexing='ij')
            dst = cc*lay.n0l*N[1]*P+off0[ii]+kk*P+tt
            mine[dst] = allT[src].numpy()[cc, ii, kk, tt]
        assert np.array_equal(mine.reshape(D, lay.n0l, N[1], P), back)
        tot = allreduce_sum(np.sum(x_loc*ref[:, sl]), torch.device('cpu'))
        err_dot = abs(tot-np.sum(x*ref))/abs(np.sum(x*ref))
        q.put((rank, float(err), float(err_rt), float(err_dot)))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        q.put((rank, 'error: %r' % (e,), 0, 0))


@pytest.mark.parametrize('world,N,D,kind', [(2, (8, 6, 5), 3, 'scalar'), (2, (6, 4, 8), 6, 'elastic'),
                                            (4, (8, 8, 6), 6, 'elastic')])
def test_slab_exchange_and_operator(world, N, D, kind):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, D, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, err, err_rt, err_dot in res:
        assert not isinstance(err, str), err
        assert err < 1e-13, (rank, err)        # slab pipeline == global operator
        assert err_rt == 0.0                   # exchange round trip is exact
        assert err_dot < 1e-13


def test_layout_rejects_indivisible_grids():
    from ffthompy_b200.slab import SlabLayout
    with pytest.raises(ValueError):
        SlabLayout((9, 8, 8), 2, 0)
    lay = SlabLayout((8, 12, 6), 4, 3)
    assert (lay.n0l, lay.n0_off, lay.n1l, lay.n1_off, lay.nh) == (2, 6, 3, 9, 4)
