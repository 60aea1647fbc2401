"""Slab-decomposed Fourier-Galerkin operator and CG across GPUs: one process per GPU,
`torch.distributed` (NCCL over NVLink) for the exchange steps.

Rank g owns the x-planes [g*N0/G, (g+1)*N0/G) of every real-space field (A, x, r, p, Ap).  One
operator application is
    S1, S2  local (A.p, R2C last axis, C2C axis 1)        on [D][n0l][N1][pitch]
    all-to-all: re-partition the half spectrum along axis 1 -> [D][N0][n1l][pitch]
    S3      local (C2C axis 0, G^ with global (k0,k1,k2), inverse axis 0)
    all-to-all back
    S4, S5  local
and CG reduces its two scalars per iteration with all-reduce (device tensors, no host round trip).

Exchange modes:
* push (default on NVLink-connected GPUs for N0 in {128, 256, 512}; csrc/fh_slab2.cu): both spectrum layouts live in
  torch symmetric memory; S2 STORES every output row k1 straight into the y-slab spectrum of the rank that owns k1 and
  S3 stores every output row i0 into the x-slab spectrum of the owner of plane i0.  No exchange buffer, no copy
  engine or NCCL traffic, no remote loads (stores are fire-and-forget, so the NVLink transfer overlaps the transforms
  of the producing kernel), six kernels and two device barriers per operator application.
* kblock (N0 in {128, 256, 512}; csrc/fh_slab2.cu): the half-spectrum columns are split into blocks of whole
  8-column tiles.  S2, exchange, S3, exchange back and S4 are independent across blocks, so block b travels on the
  copy engines (pushes into the peers' symmetric exchange buffers) while blocks b+-1 are transformed on the SMs; only
  S1 and S5 need whole rows.  One block's transfer per direction is exposed instead of the whole exchange.
* peer (N0 a power of two 16..2048): the x-slab spectra live in
  torch symmetric memory (peer-mapped over NVLink); S3 of each rank gathers its rows straight from the
  owners' spectra with remote loads, applies G^ and scatters the result back with remote stores
  (fh_ga_slab_peer).  The exchange is fused into the axis-0 kernel — no all-to-all, no exchange buffer,
  two device barriers per operator application.
* direct (N0, N1 powers of two 16..2048): the exchange buffers are chunk-major blocks
  [J][G][D][n0l/J][n1l][pitch] that the axis-1 / axis-0 kernels address in place (fh_ga_slab_direct), so
  an exchange is ONE all_to_all_single per chunk with no pack/unpack pass, issued asynchronously:
  chunk j travels over NVLink while S1+S2 of chunk j+1 (forward) or S4+S5 of chunk j-1 (backward) run.
  `direct_offsets` states the block layout; tests/test_gloo_slab.py checks it on CPU with gloo.
* packed (any sizes): pack, all_to_all_single, unpack around the natural-layout kernels.
The transforms themselves are the CUDA kernels of the C ABI (fh_ga_create_slab / fh_ga_slab_stage).
"""
import ctypes as C

import numpy as np


class SlabLayout(object):
    """Even slab partition of a 3-D grid over `world` ranks (axis 0 in real space, axis 1 for the
    axis-0 pass)."""

    def __init__(self, N, world, rank):
        self.N = tuple(int(n) for n in N)
        assert len(self.N) == 3, 'slab decomposition is for 3-D grids'
        if self.N[0] % world or self.N[1] % world:
            raise ValueError('N0=%d and N1=%d must be divisible by the number of ranks (%d)'
                             % (self.N[0], self.N[1], world))
        self.world, self.rank = int(world), int(rank)
        self.n0l, self.n1l = self.N[0]//world, self.N[1]//world
        self.n0_off, self.n1_off = rank*self.n0l, rank*self.n1l
        self.nh = self.N[2]//2+1


def exchange_fwd(spec, layout, group=None, out=None):
    """x-slabs -> y-slabs: spec [D][n0l][N1][P] (this rank's planes, all k1) ->
    [D][N0][n1l][P] (all k0, this rank's k1 range).  One pack, one all_to_all_single, one unpack
    (written straight into `out` when given)."""
    import torch
    import torch.distributed as dist
    D, n0l, N1, P = spec.shape
    G, n1l = layout.world, layout.n1l
    send = spec.reshape(D, n0l, G, n1l, P).permute(2, 0, 1, 3, 4).contiguous()   # [G][D][n0l][n1l][P]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    if out is None:
        return recv.permute(1, 0, 2, 3, 4).reshape(D, G*n0l, n1l, P)              # [D][N0][n1l][P]
    out.view(D, G, n0l, n1l, P).copy_(recv.permute(1, 0, 2, 3, 4))
    return out


def exchange_bwd(specT, layout, group=None, out=None):
    """y-slabs -> x-slabs (inverse of exchange_fwd)."""
    import torch
    import torch.distributed as dist
    D, N0, n1l, P = specT.shape
    G, n0l = layout.world, layout.n0l
    send = specT.reshape(D, G, n0l, n1l, P).permute(1, 0, 2, 3, 4).contiguous()   # [G][D][n0l][n1l][P]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=group)
    if out is None:
        return recv.permute(1, 2, 0, 3, 4).reshape(D, n0l, G*n1l, P)              # [D][n0l][N1][P]
    out.view(D, n0l, G, n1l, P).copy_(recv.permute(1, 2, 0, 3, 4))
    return out


def allreduce_sum(value, device, group=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def direct_offsets(layout, D, P, nchunk):
    """Element offsets of the chunk-major exchange blocks (mirrors fh_ga_slab_direct in csrc/fh_fused.cu).

    Returns (off1, off0, cstride1, istride1, cstride0, chunk_elems):
      x-slab side, chunk j, panel (c, i0c), row k1, column t:
          j*chunk_elems + c*cstride1 + i0c*istride1 + off1[k1] + t
      y-slab side, component c, global plane i0, inner index ii in [0, n1l*P):
          off0[i0] + c*cstride0 + ii
    """
    G, n0l, n1l = layout.world, layout.n0l, layout.n1l
    assert n0l % nchunk == 0
    n0c = n0l//nchunk
    inner = n1l*P
    k1 = np.arange(layout.N[1])
    off1 = (k1//n1l)*D*n0c*inner+(k1 % n1l)*P
    i0 = np.arange(layout.N[0])
    g, rem = i0//n0l, i0 % n0l
    j, i0c = rem//n0c, rem % n0c
    off0 = ((j*G+g)*D)*n0c*inner+i0c*inner
    return off1, off0, n0c*inner, inner, n0c*inner, G*D*n0c*inner


def kblock_offsets(layout, D, P, nblk):
    """Layout of the k2-block exchange buffers (mirrors fh_ga_slab_kblock in csrc/fh_slab2.cu): the pitch/8 column
    tiles are dealt to `nblk` blocks; block b holds [G][D][n0l][n1l][w_b] elements at offset base[b].
    Returns a list of (col0, width, base, per_peer, off1[N1], off0[N0]) per block:
      x-slab side, panel (c, i0l), row k1, block column t:   base + c*n0l*n1l*w + i0l*n1l*w + off1[k1] + t
      y-slab side, component c, global plane i0, ii = k1l*w + t:   base + off0[i0] + c*n0l*n1l*w + ii"""
    G, n0l, n1l = layout.world, layout.n0l, layout.n1l
    ntile = P//8
    out, c0 = [], 0
    for b in range(nblk):
        w = 8*(ntile//nblk+(1 if b < ntile % nblk else 0))
        inner = n1l*w
        k1 = np.arange(layout.N[1])
        i0 = np.arange(layout.N[0])
        out.append((c0, w, D*n0l*layout.N[1]*c0, D*n0l*inner, (k1//n1l)*D*n0l*inner+(k1 % n1l)*w,
                    ((i0//n0l)*D*n0l+(i0 % n0l))*inner))
        c0 += w
    return out


def push_offsets(layout, P):
    """Row offsets of the push exchange (mirrors fh_ga_slab_push), in elements and WITHOUT the peer-pointer deltas:
    off1[k1]: S2 output row k1 of panel (c, i0l) lands in rank k1//n1l's y-slab spectrum [D][N0][n1l][P] at
              c*N0*n1l*P + i0l*n1l*P + off1[k1];
    off0[i0]: S3 output row i0 of component c lands in rank i0//n0l's x-slab spectrum [D][n0l][N1][P] at
              c*n0l*N1*P + off0[i0] + ii."""
    n0l, n1l, r = layout.n0l, layout.n1l, layout.rank
    k1 = np.arange(layout.N[1])
    i0 = np.arange(layout.N[0])
    return (r*n0l*n1l+(k1 % n1l))*P, ((i0 % n0l)*layout.N[1]+r*n1l)*P


def best_exchange(A_local, G, N, group=None, modes=('kblock', 'p2p'), reps=3):
    """Measure, don't guess: build the operator with each exchange mode in `modes`, time `reps` applications on
    the device (max over ranks, so every rank takes the same decision) and return (name, {name: ms}).  Modes
    whose kernels or memory mappings are not available here are skipped; which one wins depends on grid size
    and rank count (2 x B200: 'peer' at 256^3, 'p2p' at 512^3)."""
    import torch
    import torch.distributed as dist
    from . import device as dev
    times = {}
    for m in modes:
        ok = torch.ones(1, device=dev.device())
        op = None
        try:
            op = SlabGA(A_local, G, N, group=group, exchange=m)
        except Exception:
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if ok.item() == 0:
            del op
            continue
        x = dev.zeros(tuple(A_local.shape[1:]))
        x.normal_()
        y = dev.zeros(tuple(A_local.shape[1:]))
        op.apply(x, y)
        torch.cuda.synchronize()
        dist.barrier(group=group)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            op.apply(x, y)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)/reps], dtype=torch.float64, device=dev.device())
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        times[m] = float(t.item())
        del op, x, y
        torch.cuda.empty_cache()
    if not times:
        return None, times
    return min(times, key=lambda k: times[k]), times


class SlabGA(object):
    """y = F^-1 G^ F (A x) on slab-decomposed fields, and the CG loop of general/solver.py:80-139
    over it.  `A_local`: device tensor [D][D][n0l][N1][N2]; `G`: lazy GreenTensor on the GLOBAL grid.
    `exchange`: None = best available ('peer', else 'direct', else 'packed'), or one of those names to
    require it; `nchunk`: x-plane chunks per exchange in direct mode (default: up to 4)."""

    def __init__(self, A_local, G, N, group=None, exchange=None, nchunk=None):
        import torch
        import torch.distributed as dist
        from . import _lib as L
        from . import device as dev
        self.L, self.dev, self.group = L, dev, group
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.layout = lay = SlabLayout(N, world, rank)
        self.D = D = int(A_local.shape[0])
        assert tuple(A_local.shape) == (D, D, lay.n0l, lay.N[1], lay.N[2])
        assert G.lazy and G.fft_form == 'r' and tuple(int(n) for n in G.N) == lay.N
        assert exchange in (None, 'kblock', 'push', 'peer', 'p2p', 'direct', 'packed')
        self.A = A_local.contiguous()
        self.plan = dev.plan(lay.N)
        lib = dev.lib()
        nwork = int(lib.fh_ga_slab_work_doubles(self.plan, D, lay.n0l, lay.n1l))
        # peer mode: the workspace (with the spectrum inside) is symmetric memory, mapped by every rank
        self.symm = None
        if exchange in (None, 'push', 'peer') and world > 1:
            # the allocation is local, the rendezvous collective: agree on the former before entering the latter
            # (a rank that fell back on its own would leave its peers blocked inside the rendezvous)
            work = None
            try:
                import torch.distributed._symmetric_memory as symm_mem
                work = symm_mem.empty(nwork, dtype=torch.float64, device=dev.device())
                work.zero_()
                torch.cuda.synchronize()
            except Exception:
                work = None
            if self._agree(work is not None):
                self.work = work
                self.symm = symm_mem.rendezvous(self.work, group if group is not None else dist.group.WORLD)
                dist.barrier(group=group)      # every rank's zero fill is complete before any peer may store into it
            elif exchange in ('push', 'peer'):
                raise RuntimeError('symmetric memory is not available on every rank (exchange=%r)' % exchange)
        if self.symm is None:
            self.work = dev.zeros((nwork,))
        self.handle = C.c_void_p()
        g = G.descriptor()
        L.check(lib.fh_ga_create_slab(C.byref(self.handle), self.plan, D, dev.ptr(self.A), 0, C.byref(g),
                                      dev.ptr(self.work), lay.n0l, lay.n1l, lay.n1_off))
        spec, specT, pitch = C.c_void_p(), C.c_void_p(), C.c_int()
        L.check(lib.fh_ga_buffers(self.handle, C.byref(spec), C.byref(specT), C.byref(pitch)))
        self.pitch = P = pitch.value
        base = self.work.data_ptr()
        o1 = (spec.value-base)//8
        o2 = (specT.value-base)//8
        n1 = 2*D*lay.n0l*lay.N[1]*P
        n2 = 2*D*lay.N[0]*lay.n1l*P
        self.spec = torch.view_as_complex(self.work[o1:o1+n1].reshape(-1, 2)).reshape(D, lay.n0l, lay.N[1], P)
        self.specT = torch.view_as_complex(self.work[o2:o2+n2].reshape(-1, 2)).reshape(D, lay.N[0], lay.n1l, P)
        self.nloc = lay.n0l*lay.N[1]*lay.N[2]
        self.pN = float(np.prod(lay.N))
        self.exchanged_bytes = 0
        self.sums = dev.zeros((2,))
        self.mode, self.nchunk = 'packed', 1
        if exchange in (None, 'push') and (self.symm is not None or world == 1):
            if self.symm is not None:
                pS = [int(b)+(spec.value-base) for b in self.symm.buffer_ptrs]
                pT = [int(b)+(specT.value-base) for b in self.symm.buffer_ptrs]
            else:
                pS, pT = [spec.value], [specT.value]
            rc = lib.fh_ga_slab_push(self.handle, world, rank, (C.c_void_p*world)(*pS), (C.c_void_p*world)(*pT))
            if self._agree(rc == 0):
                self.mode = 'push'
                # x-plane chunks of S1 | S2: S2 (NVLink stores) of chunk j runs on a side stream under S1 (HBM) of j+1
                import os
                want = int(nchunk) if nchunk else int(os.environ.get('FH_PUSH_CHUNKS', '1'))   # measured on 2 x B200: chunking S1|S2 does not pay (profiles/r02_*)
                need = 8 if (lay.N[2] & (lay.N[2]-1)) == 0 else 24      # rows per S1 launch: a multiple of its rows per CTA
                rows_ok = lambda J: lay.n0l % J == 0 and ((lay.n0l//J)*lay.N[1]) % need == 0   # noqa: E731
                self.nchunk = max([J for J in (8, 4, 2, 1) if J <= max(want, 1) and (J == 1 or rows_ok(J))])
                # high priority: the few persistent S2 CTAs must win the SM slots that S1 frees, or they queue behind all of S1
                self.side = torch.cuda.Stream(priority=-1) if self.nchunk > 1 else None
            elif exchange == 'push':
                L.check(rc if rc else -3)
        if self.mode == 'packed' and exchange == 'kblock':
            self._setup_kblock(nchunk)
        if self.mode == 'packed' and exchange in (None, 'peer') and (self.symm is not None or world == 1):
            if self.symm is not None:
                ptrs = [int(b)+(spec.value-base) for b in self.symm.buffer_ptrs]
                assert ptrs[rank] == spec.value
            else:
                ptrs = [spec.value]
            arr = (C.c_void_p*world)(*ptrs)
            rc = lib.fh_ga_slab_peer(self.handle, world, rank, arr)
            if self._agree(rc == 0):
                self.mode = 'peer'
            elif exchange == 'peer':
                L.check(rc if rc else -3)
        if self.mode == 'packed' and exchange in (None, 'direct', 'p2p'):
            cands = [int(nchunk)] if nchunk else [j for j in (4, 2, 1) if lay.n0l % j == 0]
            nel = D*lay.n0l*lay.N[1]*P
            self.xsymm = None
            if exchange == 'p2p' and world > 1:
                # both exchange buffers in symmetric memory: chunks are PUSHED into the peers' buffers by the
                # copy engines (cudaMemcpyAsync over NVLink on a side stream) while the SMs transform the
                # next chunk; no NCCL kernel competes for SMs and nothing is staged
                import torch.distributed._symmetric_memory as symm_mem
                xb = symm_mem.empty(4*nel, dtype=torch.float64, device=dev.device())
                xb.zero_()
                torch.cuda.synchronize()   # the zero fill must not overtake a peer's first push into this buffer
                self.xsymm = symm_mem.rendezvous(xb, group if group is not None else dist.group.WORLD)
                dist.barrier(group=group)
                cview = lambda t: torch.view_as_complex(t.reshape(-1, 2))
                bufA, bufB = cview(xb[:2*nel]), cview(xb[2*nel:])
                self._xb = xb
            else:
                bufA = torch.zeros(nel, dtype=torch.complex128, device=dev.device())
                # the y-slab workspace doubles as exchange buffer B (on one rank it aliases the x-slab spectrum)
                bufB = self.specT.reshape(-1) if world > 1 else torch.zeros_like(bufA)
            rc = -1
            for J in cands:
                rc = lib.fh_ga_slab_direct(self.handle, world, J, dev.ptr(bufA), dev.ptr(bufB))
                if self._agree(rc == 0):
                    self.mode, self.nchunk = 'direct', J
                    self.bufA, self.bufB = bufA.view(J, -1), bufB.view(J, -1)
                    break
            if self.mode != 'direct':
                del bufA, bufB
                if exchange in ('direct', 'p2p'):
                    L.check(rc)
            elif self.xsymm is not None:
                self.mode = 'p2p'
                J = self.nchunk
                self.blkA, self.blkB = self.bufA.view(J, world, -1), self.bufB.view(J, world, -1)
                self.peerA, self.peerB = [], []
                for g_ in range(world):
                    pb = self.xsymm.get_buffer(g_, (4*nel,), torch.float64, 0)
                    self.peerA.append(cview(pb[:2*nel]).view(J, world, -1))
                    self.peerB.append(cview(pb[2*nel:]).view(J, world, -1))
                import os
                nside = max(1, int(os.environ.get('FH_P2P_STREAMS', '1')))   # measured on 2 x B200: 1 stream 17.2 ms, 2: 17.8, 4: 18.2
                self.cstreams = [torch.cuda.Stream() for _ in range(nside)]
                self.nsplit = max(1, nside//max(1, world-1))   # pieces per remote block when peers are few
        self.direct = self.mode in ('direct', 'p2p')
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier(group=group)          # set-up complete on every rank before the first operator application

    def _setup_kblock(self, nblk):
        """exchange buffers in symmetric memory + block tables (fh_ga_slab_kblock)"""
        import os
        import torch
        import torch.distributed as dist
        L, dev, lay, D, P = self.L, self.dev, self.layout, self.D, self.pitch
        lib = dev.lib()
        world, rank = lay.world, lay.rank
        nel = D*lay.n0l*lay.N[1]*P
        cview = lambda t: torch.view_as_complex(t.reshape(-1, 2))    # noqa: E731
        self.xsymm = None
        if world > 1:
            import torch.distributed._symmetric_memory as symm_mem
            xb = symm_mem.empty(4*nel, dtype=torch.float64, device=dev.device())
            xb.zero_()
            torch.cuda.synchronize()
            self.xsymm = symm_mem.rendezvous(xb, self.group if self.group is not None else dist.group.WORLD)
            dist.barrier(group=self.group)
        else:
            xb = dev.zeros((4*nel,))
        self._xb = xb
        bufA, bufB = cview(xb[:2*nel]), cview(xb[2*nel:])
        J = int(nblk) if nblk else int(os.environ.get('FH_KBLOCKS', '2'))   # 8 x B200, 512^3: 2 blocks 5.03 ms, 3: 5.22, 4: 5.56
        J = max(1, min(J, P//8, 16))
        rc = lib.fh_ga_slab_kblock(self.handle, world, J, dev.ptr(bufA), dev.ptr(bufB))
        if not self._agree(rc == 0):
            L.check(rc if rc else -3)
        self.mode, self.nchunk = 'kblock', J
        self.kbufA, self.kbufB = bufA, bufB
        self.kinfo = []
        for b in range(J):
            base, per = C.c_int64(), C.c_int64()
            L.check(lib.fh_ga_slab_kblock_info(self.handle, b, C.byref(base), C.byref(per), None, None))
            self.kinfo.append((int(base.value), int(per.value)))
        self.kpeerA, self.kpeerB = [], []
        for g_ in range(world):
            if self.xsymm is not None:
                pb = self.xsymm.get_buffer(g_, (4*nel,), torch.float64, 0)
                self.kpeerA.append(cview(pb[:2*nel]))
                self.kpeerB.append(cview(pb[2*nel:]))
            else:
                self.kpeerA.append(bufA)
                self.kpeerB.append(bufB)
        # copy streams per direction (FH_KBLOCK_STREAMS, default 1).  Measured on 8 x B200 at 512^3: one stream per peer
        # (7 copy engines side by side) is SLOWER than one stream for all peers, 5.92 vs 5.03 ms per CG iteration
        # (profiles/r02_slab_8gpu_*): the concurrent copies take HBM and NVLink bandwidth from the running kernels.
        npeer = max(1, min(max(1, world-1), int(os.environ.get('FH_KBLOCK_STREAMS', '1'))))
        self.kstreams = ([torch.cuda.Stream() for _ in range(npeer)], [torch.cuda.Stream() for _ in range(npeer)])

    def _kstage(self, s, blk, p, r, pupdate, y):
        self.L.check(self.dev.lib().fh_ga_slab_kblock_stage(self.handle, s, int(blk), self.dev.ptr(p),
                                                            self.dev.ptr(r) if r is not None else None,
                                                            int(pupdate), self.dev.ptr(y)))

    def _kpush(self, peers, src, blk, streams, ready):
        """block `blk` of this rank's buffer `src`: piece g -> slot `rank` of block `blk` in peer g's buffer, the
        remote pieces spread over the copy streams (`ready`: event the producer kernel recorded); returns with every
        copy joined into streams[0]"""
        import torch
        G, me = self.layout.world, self.layout.rank
        base, per = self.kinfo[blk]
        for st in streams:
            st.wait_event(ready)
        for k in range(1, G):
            g = (me+k) % G
            with torch.cuda.stream(streams[(k-1) % len(streams)]):
                peers[g][base+me*per:base+(me+1)*per].copy_(src[base+g*per:base+(g+1)*per], non_blocking=True)
        with torch.cuda.stream(streams[0]):      # the local piece, then the join
            peers[me][base+me*per:base+(me+1)*per].copy_(src[base+me*per:base+(me+1)*per], non_blocking=True)
        for st in streams[1:]:
            e = torch.cuda.Event()
            e.record(st)
            streams[0].wait_event(e)

    def _apply_kblock(self, x, y, r, pupdate):
        import torch
        J = self.nchunk
        main = torch.cuda.current_stream()
        csF, csB = self.kstreams
        multi = self.xsymm is not None
        self._kstage(1, 0, x, r, pupdate, y)
        landed = []
        for b in range(J):
            self._kstage(2, b, x, r, 0, y)
            ev = torch.cuda.Event()
            ev.record(main)
            self._kpush(self.kpeerB, self.kbufA, b, csF, ev)
            with torch.cuda.stream(csF[0]):
                if multi:
                    self.xsymm.barrier(channel=0)          # block b of every rank has landed everywhere
                e = torch.cuda.Event()
                e.record(csF[0])
                landed.append(e)
        back = []
        for b in range(J):
            main.wait_event(landed[b])
            self._kstage(3, b, x, r, 0, y)
            ev = torch.cuda.Event()
            ev.record(main)
            self._kpush(self.kpeerA, self.kbufB, b, csB, ev)
            with torch.cuda.stream(csB[0]):
                if multi:
                    self.xsymm.barrier(channel=1)
                e = torch.cuda.Event()
                e.record(csB[0])
                back.append(e)
        for b in range(J):
            main.wait_event(back[b])
            self._kstage(4, b, x, None, 0, y)
        self._kstage(5, 0, x, None, 0, y)
        lay = self.layout
        self.exchanged_bytes += 2*self.kbufA.numel()*16*(lay.world-1)//lay.world
        return y

    def _agree(self, ok):
        """collective AND over the ranks: every rank takes the same exchange mode"""
        import torch
        import torch.distributed as dist
        if dist.get_world_size(self.group) == 1:
            return bool(ok)
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.dev.device())
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return bool(t.item())

    def _barrier(self):
        """device-side barrier across the ranks on the current stream (symmetric-memory signal pads)"""
        if self.symm is not None:
            self.symm.barrier()

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.L.load().fh_ga_destroy(self.handle)
        except Exception:
            pass

    def _stage(self, s, chunk, p, r, pupdate, y):
        if self.mode == 'push':
            self._pstage(s, 0, 1, p, r, pupdate, y)
            return
        self.L.check(self.dev.lib().fh_ga_slab_stage(self.handle, s, chunk, self.dev.ptr(p),
                                                     self.dev.ptr(r) if r is not None else None, int(pupdate),
                                                     self.dev.ptr(y)))

    def _pstage(self, s, chunk, nchunk, p, r, pupdate, y):
        self.L.check(self.dev.lib().fh_ga_slab_push_stage(self.handle, s, int(chunk), int(nchunk), self.dev.ptr(p),
                                                          self.dev.ptr(r) if r is not None else None,
                                                          int(pupdate), self.dev.ptr(y)))

    def _apply_push(self, x, y, r, pupdate):
        """S1 | S2 chunked over x-planes on two streams, barrier, S3 (stores into the owners' x-slabs), barrier, S4, S5"""
        import torch
        J = self.nchunk
        lib, L = self.dev.lib(), self.L
        if J == 1:
            self._pstage(1, 0, 1, x, r, pupdate, y)
            self._pstage(2, 0, 1, x, r, 0, y)
        else:
            main, side = torch.cuda.current_stream(), self.side
            side.wait_stream(main)            # the previous application's S4 has read this rank's x-slab spectrum
            for j in range(J):
                self._pstage(1, j, J, x, r, pupdate, y)
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                L.check(lib.fh_set_stream(C.c_void_p(side.cuda_stream)))
                try:
                    self._pstage(2, j, J, x, r, 0, y)
                finally:
                    L.check(lib.fh_set_stream(C.c_void_p(main.cuda_stream)))
            main.wait_stream(side)
        self._barrier()                      # every rank's S2 rows are in place
        self._pstage(3, 0, 1, x, r, 0, y)
        self._barrier()                      # every rank's S3 rows are back
        self._pstage(4, 0, 1, x, r, 0, y)
        self._pstage(5, 0, 1, x, r, 0, y)
        lay = self.layout
        self.exchanged_bytes += 2*self.spec.numel()*16*(lay.world-1)//lay.world
        return y

    def _a2a(self, dst, src):
        """one exchange block; returns a work handle (None on a single rank)"""
        import torch.distributed as dist
        if self.layout.world == 1:
            dst.copy_(src)
            return None
        self.exchanged_bytes += src.numel()*16*(self.layout.world-1)//self.layout.world
        return dist.all_to_all_single(dst, src, group=self.group, async_op=True)

    def apply(self, x, y=None, r=None, pupdate=0):
        """x, y: device tensors [D][n0l][N1][N2] (this rank's slab).  With `pupdate` the CG direction
        update x <- r + beta*x (beta on the device) is folded into S1, as in fh_cg_steps."""
        if y is None:
            y = self.dev.empty(x.shape)
        if self.mode == 'push':
            return self._apply_push(x, y, r, pupdate)
        if self.mode == 'kblock':
            return self._apply_kblock(x, y, r, pupdate)
        if self.mode in ('peer', 'push'):
            self._stage(1, 0, x, r, pupdate, y)
            self._stage(2, 0, x, r, 0, y)        # push: rows stored into the owners' y-slab spectra
            self._barrier()                      # every rank's S2 output is in place
            self._stage(3, 0, x, r, 0, y)        # peer: remote loads -> G^ -> remote stores; push: local loads, remote stores
            self._barrier()                      # every rank's rows are back
            self._stage(4, 0, x, r, 0, y)
            self._stage(5, 0, x, r, 0, y)
            lay = self.layout
            self.exchanged_bytes += 2*self.spec.numel()*16*(lay.world-1)//lay.world
            return y
        if self.mode == 'p2p':
            return self._apply_p2p(x, y, r, pupdate)
        if self.direct:
            J = self.nchunk
            works = []
            for j in range(J):
                self._stage(1, j, x, r, pupdate, y)
                works.append(self._a2a(self.bufB[j], self.bufA[j]))
            for w in works:
                if w is not None:
                    w.wait()
            self._stage(3, 0, x, r, 0, y)
            works = [self._a2a(self.bufA[j], self.bufB[j]) for j in range(J)]
            for j in range(J):
                if works[j] is not None:
                    works[j].wait()
                self._stage(4, j, x, None, 0, y)
            return y
        self._stage(1, 0, x, r, pupdate, y)
        self._stage(2, 0, x, r, 0, y)
        if self.layout.world > 1:
            exchange_fwd(self.spec, self.layout, self.group, out=self.specT)
            self.exchanged_bytes += self.spec.numel()*16*(self.layout.world-1)//self.layout.world
        self._stage(3, 0, x, r, 0, y)
        if self.layout.world > 1:
            exchange_bwd(self.specT, self.layout, self.group, out=self.spec)
            self.exchanged_bytes += self.spec.numel()*16*(self.layout.world-1)//self.layout.world
        self._stage(4, 0, x, r, 0, y)
        self._stage(5, 0, x, r, 0, y)
        return y

    def _push(self, peers, src, j, side):
        """copy-engine pushes of chunk j: blocks src[j, g] -> peers[g][j, me], spread over the side streams
        (`side`: every stream already waits for the producer).  With FH_P2P_STREAMS > 1 different peers (and,
        for few peers, pieces of one block) go to different streams / copy engines; on 2 GPUs that only adds
        HBM contention with the running kernels, so the default is one stream.  The local block goes last."""
        import torch
        G, me = self.layout.world, self.layout.rank
        order = [(me+k) % G for k in range(1, G)]+[me]    # ring order, start with the next rank
        k = 0
        for g in order:
            dst, blk = peers[g][j, me], src[j, g]
            n = blk.numel()
            pieces = self.nsplit if g != me else 1
            step = -(-n//pieces)
            for a in range(0, n, step):
                with torch.cuda.stream(side[k % len(side)]):
                    dst[a:a+step].copy_(blk[a:a+step], non_blocking=True)
                k += 1

    def _apply_p2p(self, x, y, r, pupdate):
        """chunk-pipelined exchange by copy-engine pushes into the peers' symmetric buffers:
        forward  S1+S2(chunk j+1) on the SMs  ||  chunk j -> peers' buffer B over NVLink
        backward chunk j+1 -> peers' buffer A  ||  S4+S5(chunk j) on the SMs"""
        import torch
        J = self.nchunk
        main, side = torch.cuda.current_stream(), self.cstreams
        cs = side[0]
        for j in range(J):
            self._stage(1, j, x, r, pupdate, y)
            ev = torch.cuda.Event()
            ev.record(main)
            for st in side:
                st.wait_event(ev)
            self._push(self.peerB, self.blkA, j, side)
        for st in side:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
        self.xsymm.barrier(channel=0)                     # every rank's pushes have landed
        self._stage(3, 0, x, r, 0, y)
        ev = torch.cuda.Event()
        ev.record(main)
        for st in side:
            st.wait_event(ev)
        evs = []
        for j in range(J):
            self._push(self.peerA, self.blkB, j, side)
            for st in side[1:]:                           # join: chunk j's copies on every side stream
                e = torch.cuda.Event()
                e.record(st)
                cs.wait_event(e)
            with torch.cuda.stream(cs):
                self.xsymm.barrier(channel=1)             # chunk j of every rank is in place
                e = torch.cuda.Event()
                e.record(cs)
                evs.append(e)
        for j in range(J):
            main.wait_event(evs[j])
            self._stage(4, j, x, None, 0, y)
        self.exchanged_bytes += 2*self.bufA.numel()*16*(self.layout.world-1)//self.layout.world
        return y

    def profile_apply(self, x, y=None, reps=3):
        """Where one operator application spends its time on this rank: {phase: ms} from CUDA events on the main
        stream (max over ranks), for the two NVLink schemes.  p2p: 'fwd' = S1+S2 of all chunks up to the moment
        every push has landed, 'S3', 'bwd' = pushes back + S4+S5; 'fwd_compute' / 'bwd_compute' are the same
        stages run back to back without the exchange, so the differences are the exposed (not hidden) transfer
        time.  peer: the five stages with the two device barriers.  A diagnostic, not used by the solve."""
        import torch
        import torch.distributed as dist
        if y is None:
            y = self.dev.empty(x.shape)
        J = self.nchunk
        marks = {}

        def timed(name, fn):
            ts = []
            for _ in range(reps+1):
                torch.cuda.synchronize()
                dist.barrier(group=self.group)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            t = torch.tensor([min(ts[1:])], dtype=torch.float64, device=self.dev.device())
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            marks[name] = float(t.item())

        timed('apply', lambda: self.apply(x, y))
        if self.mode == 'kblock':
            timed('S1', lambda: self._kstage(1, 0, x, None, 0, y))
            timed('S2', lambda: [self._kstage(2, b, x, None, 0, y) for b in range(J)])
            timed('S3', lambda: [self._kstage(3, b, x, None, 0, y) for b in range(J)])
            timed('S4', lambda: [self._kstage(4, b, x, None, 0, y) for b in range(J)])
            timed('S5', lambda: self._kstage(5, 0, x, None, 0, y))
            marks['exchange_exposed'] = marks['apply']-sum(marks['S%d' % k] for k in (1, 2, 3, 4, 5))
        elif self.mode in ('p2p', 'direct'):
            timed('fwd_compute', lambda: [self._stage(1, j, x, None, 0, y) for j in range(J)])
            timed('S3', lambda: self._stage(3, 0, x, None, 0, y))
            timed('bwd_compute', lambda: [self._stage(4, j, x, None, 0, y) for j in range(J)])
            marks['exchange_exposed'] = marks['apply']-marks['fwd_compute']-marks['S3']-marks['bwd_compute']
        else:
            for st in (1, 2, 3, 4, 5):
                timed('S%d' % st, lambda st=st: self._stage(st, 0, x, None, 0, y))
            marks['barriers_and_gaps'] = marks['apply']-sum(marks['S%d' % st] for st in (1, 2, 3, 4, 5))
        return marks

    def last_dot(self):
        """global <x, y> of the most recent apply(x, y): S5 already left this rank's partial sums on the
        device (no extra pass over the fields)"""
        loc = C.c_double()
        self.L.check(self.dev.lib().fh_ga_last_dot(self.handle, C.byref(loc)))
        return allreduce_sum(loc.value, self.dev.device(), self.group)/self.pN

    def dot(self, a, b):
        """global <a,b> = sum over all ranks / prod(N)  (Tensor scalar product, tensors/objects.py:635)"""
        from . import ops
        return allreduce_sum(ops.dot(a, b), a.device, self.group)/self.pN

    def _global_scalar(self, mode, want_norm):
        """local partial sums -> device sum -> all-reduce -> rr / alpha / beta on the device"""
        import torch.distributed as dist
        L, lib, dev = self.L, self.dev.lib(), self.dev
        L.check(lib.fh_cgd_local_sum(self.handle, dev.ptr(self.sums)))
        if self.layout.world > 1:
            dist.all_reduce(self.sums[:1], op=dist.ReduceOp.SUM, group=self.group)
        norm = C.c_double()
        L.check(lib.fh_cgd_scal(self.handle, dev.ptr(self.sums), mode, C.byref(norm) if want_norm else None))
        return norm.value

    def cg_begin(self, B, x0):
        """initial residual of general/solver.py:80-100 (one operator application on x0); returns the
        CG state (x, vecs, r, p, Ap, have_beta, norm_res) that `cg_steps` advances."""
        L, lib, dev = self.L, self.dev.lib(), self.dev
        n = self.D*self.nloc
        shape = tuple(x0.shape)
        # the CG vectors live in buffers owned by the operator: the CUDA graph of the iteration (captured once, see
        # _capture_iteration) holds their addresses and serves every later solve on this operator
        if getattr(self, '_cgbuf', None) is None:
            self._cgbuf = (dev.empty(shape), dev.empty((3*n,)))
        x, vecs = self._cgbuf
        x.copy_(x0)
        r, p, Ap = (vecs[i*n:(i+1)*n].view(shape) for i in range(3))
        self.apply(x, Ap)
        L.check(lib.fh_cgd_init(self.handle, dev.ptr(B), dev.ptr(vecs)))
        norm_res = self._global_scalar(0, True)
        return {'x': x, 'vecs': vecs, 'r': r, 'p': p, 'Ap': Ap, 'have_beta': 0, 'norm_res': norm_res, 'kit': 0,
                'hist': [norm_res]}

    def _capture_iteration(self, st):
        """One steady-state CG iteration (deferred x update pending, p = r + beta p folded into S1) as a CUDA graph:
        operator application with its exchange on the copy streams, both scalar reductions with their all-reduces,
        and the residual update.  Replaying it costs one launch instead of ~100 host calls (8 GPUs, 3 blocks: 48 copies,
        14 kernels, 12 events, 6 barriers) — the host was the bottleneck of the chunked exchanges (profiles/r02_*)."""
        import torch
        import torch.distributed as dist
        L, lib, dev = self.L, self.dev.lib(), self.dev
        main = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(main)
        graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(graph, stream=side, capture_error_mode='thread_local'):
                cur = torch.cuda.current_stream()
                L.check(lib.fh_set_stream(C.c_void_p(cur.cuda_stream)))
                L.check(lib.fh_ga_set_xacc(self.handle, dev.ptr(st['x'])))
                self.apply(st['p'], st['Ap'], r=st['r'], pupdate=1)
                L.check(lib.fh_ga_set_xacc(self.handle, None))
                L.check(lib.fh_cgd_local_sum(self.handle, dev.ptr(self.sums)))
                if self.layout.world > 1:
                    dist.all_reduce(self.sums[:1], op=dist.ReduceOp.SUM, group=self.group)
                L.check(lib.fh_cgd_scal(self.handle, dev.ptr(self.sums), 1, None))
                L.check(lib.fh_cgd_update_r(self.handle, dev.ptr(st['vecs'])))
                L.check(lib.fh_cgd_local_sum(self.handle, dev.ptr(self.sums)))
                if self.layout.world > 1:
                    dist.all_reduce(self.sums[:1], op=dist.ReduceOp.SUM, group=self.group)
        except Exception:
            graph = None
        finally:
            L.check(lib.fh_ga_set_xacc(self.handle, None))
            L.check(lib.fh_set_stream(C.c_void_p(main.cuda_stream)))
        return graph

    def cg_steps(self, st, tol, nsteps):
        """at most `nsteps` further CG iterations (stops early once ||r|| <= tol, the reference's absolute
        test); alpha, beta, rr stay on the device, the host reads only ||r|| (8 bytes) per iteration.
        Returns the number of iterations done."""
        import os
        L, lib, dev = self.L, self.dev.lib(), self.dev
        done = 0
        # deferred x update (as in fh_cg_steps): x += alpha p of iteration k is applied by S1 of iteration k+1,
        # which has p in registers anyway (one field read less per iteration), and flushed before returning
        defer = bool(lib.fh_ga_can_defer_x(self.handle))
        want_graph = defer and self.mode in ('kblock', 'p2p', 'push', 'peer') and os.environ.get('FH_SLAB_GRAPH', '1') != '0'
        pending = st.get('pending', False)
        while st['norm_res'] > tol and done < nsteps:
            done += 1
            if want_graph and pending and st['have_beta']:
                if 'graph' not in st:
                    key = (st['x'].data_ptr(), st['vecs'].data_ptr())
                    if getattr(self, '_graph', (None, None))[0] != key:
                        g = self._capture_iteration(st)       # every rank captures: the capture contains collectives
                        if not self._agree(g is not None):
                            g = None
                        self._graph = (key, g)
                    st['graph'] = self._graph[1]
                if st['graph'] is not None:
                    st['graph'].replay()
                    if self.layout.world > 1:
                        self.exchanged_bytes += 2*self.spec.numel()*16*(self.layout.world-1)//self.layout.world
                    st['norm_res'] = self._scal_only(2)
                    st['hist'].append(st['norm_res'])
                    continue
            if pending:
                L.check(lib.fh_ga_set_xacc(self.handle, dev.ptr(st['x'])))
            try:
                self.apply(st['p'], st['Ap'], r=st['r'], pupdate=st['have_beta'])  # p = r + beta p folded into S1
            finally:
                if pending:
                    L.check(lib.fh_ga_set_xacc(self.handle, None))
            pending = False
            self._global_scalar(1, False)
            if defer:
                L.check(lib.fh_cgd_update_r(self.handle, dev.ptr(st['vecs'])))
                pending = True
            else:
                L.check(lib.fh_cgd_update(self.handle, dev.ptr(st['x']), dev.ptr(st['vecs'])))
            st['norm_res'] = self._global_scalar(2, True)
            st['have_beta'] = 1
            st['hist'].append(st['norm_res'])
        if pending and st.get('flush', True):
            L.check(lib.fh_cgd_xflush(self.handle, dev.ptr(st['x']), dev.ptr(st['vecs'])))
            pending = False
        st['pending'] = pending
        st['kit'] += done
        return done

    def _scal_only(self, mode):
        """the device scalars from the already all-reduced sum + the 8-byte read-back of ||r||"""
        norm = C.c_double()
        self.L.check(self.dev.lib().fh_cgd_scal(self.handle, self.dev.ptr(self.sums), mode, C.byref(norm)))
        return norm.value

    def cg(self, B, x0, tol=1e-6, maxiter=1000):
        """general/solver.py:80-139 on the slab.  Returns x (device, local slab), info."""
        st = self.cg_begin(B, x0)
        self.cg_steps(st, tol, maxiter)
        kit = st['kit']
        return st['x'].clone(), {'kit': kit, 'norm_res': st['norm_res'] if kit > 0 else 0,
                         'norm_res_log': np.array(st['hist'])}
