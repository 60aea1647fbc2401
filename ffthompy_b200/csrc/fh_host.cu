// fh_host.cu — device -> host transfer of large results into a FRESH pageable array (what `Tensor.val` hands out,
// ffthompy/tensors/objects.py:119-121: a new NumPy array per field).
//
// Why a dedicated routine: `cudaMemcpy` into fresh pageable memory is dominated by the first-touch page faults of the
// destination, taken one 4 KB page at a time by a single thread (measured: 0.36 s for the 0.8 GB solution of the 256^3
// elasticity solve, 63 % of the whole end-to-end solve; PCIe alone needs 0.016 s).  Here the device data travels through a
// small ring of page-locked staging slots with asynchronous copies, and a handful of worker threads move the slots into
// the destination, so the page faults are taken in parallel (and on 2 MB pages where the kernel grants them) while the
// next slots are in flight on the copy engine.
#include "fh_common.cuh"
#include "../../include/ffthom_b200.h"
#include <string.h>
#include <sys/mman.h>
#include <atomic>
#include <thread>
#include <vector>

namespace {
constexpr int kSlots = 8;
constexpr size_t kSlotBytes = (size_t)8 << 20;
struct Ring {
    unsigned char* host = nullptr;  // kSlots * kSlotBytes, page-locked
    cudaEvent_t ev[kSlots];
    cudaEvent_t ready;
    cudaStream_t stream = nullptr;
    bool ok = false;
};
Ring g_ring;

int ring_init() {
    if (g_ring.ok) return FH_OK;
    FH_CUDA(cudaHostAlloc((void**)&g_ring.host, kSlots * kSlotBytes, cudaHostAllocDefault));
    FH_CUDA(cudaStreamCreateWithFlags(&g_ring.stream, cudaStreamNonBlocking));
    for (int i = 0; i < kSlots; ++i) FH_CUDA(cudaEventCreateWithFlags(&g_ring.ev[i], cudaEventDisableTiming));
    FH_CUDA(cudaEventCreateWithFlags(&g_ring.ready, cudaEventDisableTiming));
    g_ring.ok = true;
    return FH_OK;
}
}  // namespace

// dst: host memory (pageable is fine), src: device memory; ordered after the work already enqueued on the library
// stream, complete on return.
extern "C" int fh_download(void* dst, const void* src, int64_t bytes) {
    FH_REQUIRE(dst && src && bytes >= 0, "fh_download: bad argument");
    if (bytes == 0) return FH_OK;
    int rc;
    if ((rc = ring_init())) return rc;
    // ask for 2 MB pages on the part of the destination that can have them (no-op where THP is off)
    {
        const uintptr_t a = ((uintptr_t)dst + ((size_t)2 << 20) - 1) & ~(((uintptr_t)2 << 20) - 1);
        const uintptr_t e = ((uintptr_t)dst + (size_t)bytes) & ~(((uintptr_t)2 << 20) - 1);
        if (e > a) madvise((void*)a, e - a, MADV_HUGEPAGE);
    }
    FH_CUDA(cudaEventRecord(g_ring.ready, fh_stream()));
    FH_CUDA(cudaStreamWaitEvent(g_ring.stream, g_ring.ready, 0));
    const int64_t nchunk = (bytes + (int64_t)kSlotBytes - 1) / (int64_t)kSlotBytes;
    unsigned hw = std::thread::hardware_concurrency();
    int nthr = hw ? (int)hw : 4;
    if (nthr > kSlots - 2) nthr = kSlots - 2;
    if (nthr > nchunk) nthr = (int)nchunk;
    if (nthr < 1) nthr = 1;
    // chunk i lives in slot i % kSlots; `issued` / `drained` order the producer (this thread) and the workers
    std::atomic<int64_t> issued(0), next(0);
    std::atomic<int64_t> drained[kSlots];
    for (int s = 0; s < kSlots; ++s) drained[s].store(-1);  // index of the last chunk copied out of slot s
    std::atomic<int> err(0);
    auto worker = [&]() {
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= nchunk) return;
            while (issued.load(std::memory_order_acquire) <= i) {
                if (err.load()) return;
                std::this_thread::yield();
            }
            const int s = (int)(i % kSlots);
            if (cudaEventSynchronize(g_ring.ev[s]) != cudaSuccess) {
                err.store(1);
                return;
            }
            const size_t off = (size_t)i * kSlotBytes;
            const size_t len = ((size_t)bytes - off < kSlotBytes) ? (size_t)bytes - off : kSlotBytes;
            memcpy((unsigned char*)dst + off, g_ring.host + (size_t)s * kSlotBytes, len);
            drained[s].store(i, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    try {
        for (int t = 0; t < nthr; ++t) pool.emplace_back(worker);
    } catch (...) {  // no threads to be had (restricted container): plain copy, ordered after the library stream
        err.store(1);
        for (auto& t : pool) t.join();
        FH_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, g_ring.stream));
        FH_CUDA(cudaStreamSynchronize(g_ring.stream));
        return FH_OK;
    }
    cudaError_t ce = cudaSuccess;
    for (int64_t i = 0; i < nchunk && ce == cudaSuccess && !err.load(); ++i) {
        const int s = (int)(i % kSlots);
        // the slot's previous tenant (chunk i - kSlots) must have left
        while (i >= kSlots && drained[s].load(std::memory_order_acquire) < i - kSlots) {
            if (err.load()) break;
            std::this_thread::yield();
        }
        const size_t off = (size_t)i * kSlotBytes;
        const size_t len = ((size_t)bytes - off < kSlotBytes) ? (size_t)bytes - off : kSlotBytes;
        ce = cudaMemcpyAsync(g_ring.host + (size_t)s * kSlotBytes, (const unsigned char*)src + off, len,
                             cudaMemcpyDeviceToHost, g_ring.stream);
        if (ce == cudaSuccess) ce = cudaEventRecord(g_ring.ev[s], g_ring.stream);
        if (ce != cudaSuccess) err.store(1);
        issued.store(i + 1, std::memory_order_release);
    }
    if (ce != cudaSuccess || err.load()) {
        err.store(1);
        issued.store(nchunk, std::memory_order_release);
    }
    for (auto& t : pool) t.join();
    if (ce != cudaSuccess) return fh_set_error(FH_ERR_CUDA, "fh_download: %s", cudaGetErrorString(ce));
    if (err.load()) return fh_set_error(FH_ERR_CUDA, "fh_download: staging copy failed");
    return FH_OK;
}
