"""Slab-decomposed Fourier-Galerkin operator and CG across GPUs: one process per GPU,
`torch.distributed` (NCCL over NVLink) for the exchange steps.

Rank g owns the x-planes [g*N0/G, (g+1)*N0/G) of every real-space field (A, x, r, p, Ap).  One
operator application is
    S1, S2  local (A.p, R2C last axis, C2C axis 1)        on [D][n0l][N1][pitch]
    all-to-all: re-partition the half spectrum along axis 1 -> [D][N0][n1l][pitch]
    S3      local (C2C axis 0, G^ with global (k0,k1,k2), inverse axis 0)
    all-to-all back
    S4, S5  local
and CG reduces its two scalars per iteration with all-reduce.  The exchange helpers are plain
tensor plumbing (pack, all_to_all_single, unpack) and run on any backend — tests/test_gloo_slab.py
drives them on CPU with gloo; the transforms themselves are the CUDA kernels of the C ABI
(fh_ga_create_slab / fh_ga_stage).
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


class SlabGA(object):
    """y = F^-1 G^ F (A x) on slab-decomposed fields, and the CG loop of general/solver.py:80-139
    over it.  `A_local`: device tensor [D][D][n0l][N1][N2]; `G`: lazy GreenTensor on the GLOBAL grid."""

    def __init__(self, A_local, G, N, group=None):
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
        self.A = A_local.contiguous()
        self.plan = dev.plan(lay.N)
        lib = dev.lib()
        nwork = int(lib.fh_ga_slab_work_doubles(self.plan, D, lay.n0l, lay.n1l))
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

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.L.load().fh_ga_destroy(self.handle)
        except Exception:
            pass

    def _stage(self, s, x, y):
        self.L.check(self.dev.lib().fh_ga_stage(self.handle, s, self.dev.ptr(x), self.dev.ptr(y)))

    def apply(self, x, y=None):
        """x, y: device tensors [D][n0l][N1][N2] (this rank's slab)."""
        if y is None:
            y = self.dev.empty(x.shape)
        self._stage(1, x, y)
        self._stage(2, x, y)
        if self.layout.world > 1:
            exchange_fwd(self.spec, self.layout, self.group, out=self.specT)
            self.exchanged_bytes += self.spec.numel()*16*(self.layout.world-1)//self.layout.world
        self._stage(3, x, y)
        if self.layout.world > 1:
            exchange_bwd(self.specT, self.layout, self.group, out=self.spec)
            self.exchanged_bytes += self.spec.numel()*16*(self.layout.world-1)//self.layout.world
        self._stage(4, x, y)
        self._stage(5, x, y)
        return y

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

    def cg(self, B, x0, tol=1e-6, maxiter=1000):
        """general/solver.py:80-139; returns x (device, local slab), info."""
        from . import ops
        L, lib, dev = self.L, self.dev.lib(), self.dev
        n = self.D*self.nloc
        x = ops.clone(x0)
        Ap = self.apply(x)
        r = ops.axpby(1., B, -1., Ap)
        p = ops.clone(r)
        rr = self.dot(r, r)
        kit = 0
        norm_res = rr**0.5
        hist = [norm_res]
        while norm_res > tol and kit < maxiter:
            kit += 1
            self.apply(p, Ap)
            alp = rr/self.last_dot()
            loc = C.c_double()
            L.check(lib.fh_cg_xr_update(n, dev.ptr(x), dev.ptr(r), dev.ptr(p), dev.ptr(Ap), float(alp), C.byref(loc)))
            rrnext = allreduce_sum(loc.value, x.device, self.group)/self.pN
            bet = rrnext/rr
            rr = rrnext
            L.check(lib.fh_cg_p_update(n, dev.ptr(p), dev.ptr(r), float(bet)))
            norm_res = rr**0.5
            hist.append(norm_res)
        return x, {'kit': kit, 'norm_res': norm_res if kit > 0 else 0, 'norm_res_log': np.array(hist)}
