// fh_mid2.cu — launchers of the 8-column axis-0 + G^ kernels (fh_mid2.cuh).  Internal C++ interface (fh_mid2.h)
// used by the fused operator in fh_fused.cu; nothing here is part of the C ABI.
#include "fh_mid2.cuh"
#include "fh_mid3.cuh"
#include "fh_mid2.h"
#include <stdlib.h>

static int mid2_env(const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
}
bool fh_mid2_can(int n) { return n == 128 || n == 256; }
bool fh_mid2_len(int n) {
    // 0: k_mid_green_pipe; 1: the 8-column kernel of fh_mid2.cuh (measured slower);
    // 2: the row-group-per-warp kernel of fh_mid3.cuh (N0 = 256)
    static const int on = mid2_env("FH_MID2", 0);
    if (on == 2) return n == 256;
    return on && (n == 128 || n == 256);  // (64 keeps the round-1 kernel: 64 threads per CTA would leave the SM idle)
}

template <int KIND>
static int mid3_launch(cplx* data, const cplx* tw, const GreenDesc& g, const Mid3Map& m, int nh) {
    using Cfg = Mid3Cfg<KIND>;
    if (Cfg::SMEM > (size_t)fh_max_smem_optin())
        return fh_set_error(FH_ERR_UNSUPPORTED, "axis-0 pass: %zu bytes of shared memory", Cfg::SMEM);
    FH_CUDA(cudaFuncSetAttribute(k_mid3<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    int grid = fh_num_sms();
    if (grid > m.ntiles) grid = m.ntiles;
    k_mid3<KIND><<<grid, Cfg::NT, Cfg::SMEM, fh_stream()>>>(data, tw, g, m, nh);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

template <int N, int KIND, int MINB, int PREF>
static int mid2_launch(cplx* data, const cplx* tw, const GreenDesc& g, const Mid2Map& m, int nh) {
    using Cfg = Mid2Cfg<N, KIND>;
    auto kern = k_mid2<N, KIND, MINB, PREF>;
    if (Cfg::SMEM > (size_t)fh_max_smem_optin())
        return fh_set_error(FH_ERR_UNSUPPORTED, "axis-0 pass: %zu bytes of shared memory", Cfg::SMEM);
    FH_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    int per_sm = 1;
    FH_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, Cfg::NT, Cfg::SMEM));
    if (per_sm < 1) return fh_set_error(FH_ERR_UNSUPPORTED, "axis-0 pass: kernel does not fit an SM");
    int grid = per_sm * fh_num_sms();
    if (grid > m.ntiles) grid = m.ntiles;
    kern<<<grid, Cfg::NT, Cfg::SMEM, fh_stream()>>>(data, tw, g, m, nh);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

template <int N, int KIND>
static int mid2_N(cplx* data, const cplx* tw, const GreenDesc& g, const Mid2Map& m, int nh) {
    using Cfg = Mid2Cfg<N, KIND>;
    static const int pref = mid2_env("FH_MID2_PREF", 3);
    // one CTA per SM when the tile fills shared memory (register prefetch hides the load of the next tile),
    // two when two tiles fit (the two CTAs overlap each other's loads)
    if (2 * Cfg::SMEM + 2048 <= (size_t)fh_max_smem_optin() && Cfg::D == 3) return mid2_launch<N, KIND, 2, 0>(data, tw, g, m, nh);
    switch (pref) {
        case 0: return mid2_launch<N, KIND, 1, 0>(data, tw, g, m, nh);
        case 1: return mid2_launch<N, KIND, 1, 1>(data, tw, g, m, nh);
        case 2: return mid2_launch<N, KIND, 1, 2>(data, tw, g, m, nh);
    }
    return mid2_launch<N, KIND, 1, 3>(data, tw, g, m, nh);
}

int fh_mid2_green(int N, int kind, cplx* data, const cplx* tw, const GreenDesc& g, const int64_t* rowoff,
                  int64_t rstride, int64_t cstride, int spitch, int kcol0, int nh, int nrow, int col0, int ncols,
                  cplx* dout, const int64_t* rowoff_out, int64_t cstride_out) {
    if (spitch % 8 || col0 % 8 || ncols % 8 || ncols <= 0)
        return fh_set_error(FH_ERR_UNSUPPORTED, "axis-0 pass: 8-column tiles need 128-byte aligned rows (pitch %d, columns %d+%d)",
                            spitch, col0, ncols);
    static const int which = mid2_env("FH_MID2", 0);
    if (which == 2 && N == 256 && !rowoff && !dout && kcol0 == 0) {
        Mid3Map m3;
        m3.rstride = rstride;
        m3.cstride = cstride;
        m3.spitch = spitch;
        m3.tpr = ncols / 4;
        m3.ntiles = nrow * m3.tpr;
        m3.col0 = col0;
        return (kind == FH_GREEN_ELASTIC) ? mid3_launch<FH_GREEN_ELASTIC>(data, tw, g, m3, nh)
                                          : mid3_launch<FH_GREEN_SCALAR>(data, tw, g, m3, nh);
    }
    Mid2Map m;
    m.rowoff = rowoff;
    m.rstride = rstride;
    m.cstride = cstride;
    m.spitch = spitch;
    m.kcol0 = kcol0;
    m.tpr = ncols / 8;
    m.ntiles = nrow * m.tpr;
    m.col0 = col0;
    m.dout = dout;
    m.rowoff_out = rowoff_out;
    m.cstride_out = cstride_out;
    const bool el = kind == FH_GREEN_ELASTIC;
    switch (N) {
        case 128: return el ? mid2_N<128, FH_GREEN_ELASTIC>(data, tw, g, m, nh) : mid2_N<128, FH_GREEN_SCALAR>(data, tw, g, m, nh);
        case 256: return el ? mid2_N<256, FH_GREEN_ELASTIC>(data, tw, g, m, nh) : mid2_N<256, FH_GREEN_SCALAR>(data, tw, g, m, nh);
    }
    return fh_set_error(FH_ERR_UNSUPPORTED, "no 8-column axis-0 kernel for N0=%d", N);
}
