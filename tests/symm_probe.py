"""Development probe (run under torchrun on >= 2 GPUs): does torch symmetric memory work on this box,
and what do SM-issued remote stores / loads over NVLink sustain?  Not a pytest."""
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import torch.distributed._symmetric_memory as symm_mem
    n = 1 << 27   # 1 GiB of float64
    t = symm_mem.empty(n, dtype=torch.float64, device=torch.device('cuda', local))
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    if rank == 0:
        print('symmetric memory ok: ptrs', [hex(p) for p in hdl.buffer_ptrs], 'multicast', hdl.has_multicast_support,
              'signal pad', hdl.signal_pad_size, flush=True)
    peer = hdl.get_buffer((rank+1) % world, (n,), torch.float64, 0)
    src = torch.full((n,), float(rank+1), dtype=torch.float64, device='cuda')
    dst = torch.empty_like(src)

    def timeit(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)/reps

    res = {}
    res['copy engine / memcpy store to peer'] = timeit(lambda: peer.copy_(src))
    res['SM kernel remote store (mul out=peer)'] = timeit(lambda: torch.mul(src, 1.0, out=peer))
    res['SM kernel remote load (mul peer -> local)'] = timeit(lambda: torch.mul(peer, 1.0, out=dst))
    res['local copy'] = timeit(lambda: torch.mul(src, 1.0, out=dst))
    hdl.barrier()
    torch.mul(src, 1.0, out=peer)
    hdl.barrier()
    torch.cuda.synchronize()
    ok = bool((t == float((rank-1) % world+1)).all().item())
    t0 = time.perf_counter()
    for _ in range(100):
        hdl.barrier()
    torch.cuda.synchronize()
    tb = (time.perf_counter()-t0)/100
    if rank == 0:
        for k, v in res.items():
            print('%-45s %.3f ms  %.0f GB/s' % (k, v, n*8/v/1e6), flush=True)
        print('contents after barrier ok:', ok, ' device barrier %.1f us' % (tb*1e6), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
