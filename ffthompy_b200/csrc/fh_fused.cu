// fh_fused.cu — the fused Fourier–Galerkin operator  y = F^-1 G^(xi) F (A x)  and
// the Krylov loops that iterate it on the device.
//
// Reference call path: applications.py:58-81 builds
//   Afun = Operator([[ Operator([[FiN, G^, FN]]), A ]])      (tensors/operators.py:136-144)
// and hands it to linear_solver('CG' | 'richardson')          (general/solver.py:63-139).
// Here the operator is a fixed pipeline of five axis passes over a half spectrum whose rows
// are padded to 128 B:
//   S1  sigma = A p (fused with the CG direction update p = r + beta p) -> R2C along the last axis
//   S2  C2C along axis 1                                   (3-D only)
//   S3  C2C along axis 0, closed-form G^(xi), inverse C2C along axis 0   (fh_green.cuh; G^ is
//       never materialised)
//   S4  inverse C2C along axis 1                           (3-D only)
//   S5  C2R along the last axis, scaled by 1/prod(N), fused with the partial sums of <p, Ap>
// Power-of-two axis lengths (64, 128, 256) use the register-resident kernels of fh_fast.cuh,
// any other length the generic shared-memory passes of fh_fft.cu.
#include "fh_ga.cuh"
#include "fh_fast.cuh"
#include "fh_reg3.h"
#include "fh_mid2.h"
#include "fh_odd.h"
#include "../../include/ffthom_b200.h"
#include <stdlib.h>
#include <string.h>

int fh_fill_green(GreenDesc& g, const fh_green* in);

// rows handled by the next S1 / S5 launch: element offsets into fields (ro) and spectrum rows (so),
// CTA count and first partial-sum slot
struct RowRange {
    int64_t ro, so;
    unsigned nblk, pb;
};
static RowRange row_range(const fh_ga* op, int TRW, bool round_up) {
    const int nlast = op->plan->N[op->plan->dim - 1];
    const int64_t beg = op->row_cnt ? op->row_beg : 0;
    const int64_t cnt = op->row_cnt ? op->row_cnt : op->nrows;
    RowRange rr;
    rr.ro = beg * nlast;
    rr.so = beg * op->pitch;
    rr.nblk = (unsigned)(round_up ? fh_ceil_div(cnt, TRW) : cnt / TRW);
    rr.pb = (unsigned)(beg / TRW);
    return rr;
}

static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
}
static bool use_reg3() {
    static const int v = env_int("FH_REG3", 1);
    return v != 0;
}

// ------------------------------------------------------------------ generic pieces
template <int D>
__global__ void k_apply_A(int64_t n, const double* __restrict__ A, const double* __restrict__ x,
                          double* __restrict__ y) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        double xv[D];
#pragma unroll
        for (int j = 0; j < D; ++j) xv[j] = x[(size_t)j * n + p];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) acc += A[((size_t)i * D + j) * n + p] * xv[j];
            y[(size_t)i * n + p] = acc;
        }
    }
}

// Generic axis-0 pass with G^ (any length): array [D][n0][inner], ping-pong SoA buffers.
template <int KIND, int DIM>
__global__ void __launch_bounds__(GA_NT) k_mid_green(cplx* __restrict__ data, AxisDesc ax, GreenDesc g, int64_t inner,
                                                     int T, int ld, int nh, int pitch) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    double* sm = reinterpret_cast<double*>(fh_smem_raw);
    const int n = ax.n;
    double* b0re = sm;
    double* b0im = sm + (size_t)n * ld;
    double* b1re = sm + (size_t)2 * n * ld;
    double* b1im = sm + (size_t)3 * n * ld;
    const int64_t i0 = (int64_t)blockIdx.x * T;
    const int nl = (int)min((int64_t)T, inner - i0);
    const int nln = n * nl;
    for (int idx0 = threadIdx.x; idx0 < D * nln; idx0 += 4 * blockDim.x) {
        cplx v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * blockDim.x;
            if (idx < D * nln) {
                const int c = idx / nln, rem = idx - c * nln;
                const int row = rem / nl, t = rem - row * nl;
                v[u] = data[((int64_t)c * n + row) * inner + i0 + t];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * blockDim.x;
            if (idx < D * nln) {
                const int c = idx / nln, rem = idx - c * nln;
                const int row = rem / nl, t = rem - row * nl;
                b0re[row * ld + c * nl + t] = v[u].x;
                b0im[row * ld + c * nl + t] = v[u].y;
            }
        }
    }
    __syncthreads();
    int cur = fft_smem<false>(b0re, b0im, b1re, b1im, ax, D * nl, ld);
    double* cre = cur ? b1re : b0re;
    double* cim = cur ? b1im : b0im;
    double* ore = cur ? b0re : b1re;
    double* oim = cur ? b0im : b1im;
    for (int idx = threadIdx.x; idx < nln; idx += blockDim.x) {
        const int row = idx / nl, t = idx - row * nl;
        int k[3];
        k[0] = fh_freq(row, n);
        const int64_t ii = i0 + t;
        bool valid;
        if (DIM == 3) {
            const int i1 = (int)(ii / pitch), i2 = (int)(ii - (int64_t)i1 * pitch);
            k[1] = fh_freq(i1 + g.ioff1, g.N[1]);
            k[2] = fh_freq(i2, g.N[2]);
            valid = i2 < nh;
        } else {
            k[1] = fh_freq((int)ii, g.N[1]);
            k[2] = 0;
            valid = (int)ii < nh;
        }
        cplx e[D];
#pragma unroll
        for (int c = 0; c < D; ++c) e[c] = make_double2(cre[row * ld + c * nl + t], cim[row * ld + c * nl + t]);
        if (valid) {
            green_apply<KIND, DIM>(g, k, e);
        } else {
#pragma unroll
            for (int c = 0; c < D; ++c) e[c] = make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int c = 0; c < D; ++c) {
            cre[row * ld + c * nl + t] = e[c].x;
            cim[row * ld + c * nl + t] = e[c].y;
        }
    }
    __syncthreads();
    cur = fft_smem<true>(cre, cim, ore, oim, ax, D * nl, ld);
    const double* rre = cur ? ore : cre;
    const double* rim = cur ? oim : cim;
    for (int idx = threadIdx.x; idx < D * nln; idx += blockDim.x) {
        const int c = idx / nln, rem = idx - c * nln;
        const int row = rem / nl, t = rem - row * nl;
        data[((int64_t)c * n + row) * inner + i0 + t] = make_double2(rre[row * ld + c * nl + t], rim[row * ld + c * nl + t]);
    }
}

template <int KIND, int DIM>
static int launch_mid_green_generic(fh_ga* op) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    const fh_plan* p = op->plan;
    const AxisDesc& ax = p->ax[0];
    const int64_t inner = (DIM == 3) ? (int64_t)op->n1l * op->pitch : op->pitch;
    int T = 8;
    while (T > 1 && fft_smem_bytes(ax.n, fft_ld(D * T)) > (size_t)113 * 1024) T >>= 1;
    while (T > 1 && fft_smem_bytes(ax.n, fft_ld(D * T)) > (size_t)fh_max_smem_optin()) T >>= 1;
    const int ld = fft_ld(D * T);
    const size_t smem = fft_smem_bytes(ax.n, ld);
    if (smem > (size_t)fh_max_smem_optin())
        return fh_set_error(FH_ERR_UNSUPPORTED, "fused Green pass: N0=%d with D=%d needs %zu B shared memory", ax.n, D, smem);
    if (smem > 48 * 1024)
        FH_CUDA(cudaFuncSetAttribute(k_mid_green<KIND, DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nblk = fh_ceil_div(inner, T);
    k_mid_green<KIND, DIM><<<(unsigned)nblk, GA_NT, smem, fh_stream()>>>(op->specT, ax, op->g, inner, T, ld, p->nh,
                                                                          op->pitch);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// ------------------------------------------------------------------ fast-kernel launchers
template <typename K>
static int smem_attr(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) FH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return FH_OK;
}

// col0/ncols: column chunk of the spectrum rows (L2 blocking); ncols = 0 means whole rows
template <int N, int T>
static int launch_c2c_fast_NT(const cplx* tw, cplx* data, int64_t outer, int64_t inner, bool inv, double scale,
                              int col0, int ncols) {
    const size_t smem = (size_t)N * T * sizeof(cplx);
    const int ntile = ncols ? ncols / T : (int)(inner / T);
    const int tile0 = col0 / T;
    const unsigned nblk = (unsigned)(outer * ntile);
    const int nt = T * FastCfg<N>::TPL;
    int rc;
    if (inv) {
        if ((rc = smem_attr(k_c2c_fast<N, T, true>, smem))) return rc;
        k_c2c_fast<N, T, true><<<nblk, nt, smem, fh_stream()>>>(data, data, tw, inner, ntile, tile0, scale);
    } else {
        if ((rc = smem_attr(k_c2c_fast<N, T, false>, smem))) return rc;
        k_c2c_fast<N, T, false><<<nblk, nt, smem, fh_stream()>>>(data, data, tw, inner, ntile, tile0, scale);
    }
    FH_LAUNCH_CHECK();
    return FH_OK;
}

template <int N, int T>
static int launch_c2c_gen3_NT(const cplx* tw, cplx* data, int64_t outer, int64_t inner, bool inv) {
    const size_t smem = (size_t)(N + N / 16) * T * sizeof(cplx);
    const int ntile = (int)(inner / T);
    const unsigned nblk = (unsigned)(outer * ntile);
    int rc;
    if (inv) {
        if ((rc = smem_attr(k_c2c_gen3<N, T, true>, smem))) return rc;
        k_c2c_gen3<N, T, true><<<nblk, 256, smem, fh_stream()>>>(data, data, tw, inner, ntile, 0, 1.0);
    } else {
        if ((rc = smem_attr(k_c2c_gen3<N, T, false>, smem))) return rc;
        k_c2c_gen3<N, T, false><<<nblk, 256, smem, fh_stream()>>>(data, data, tw, inner, ntile, 0, 1.0);
    }
    FH_LAUNCH_CHECK();
    return FH_OK;
}

template <int N, int T>
static int launch_c2c_reg3_NT(const cplx* tw, cplx* data, int64_t outer, int64_t inner, bool inv) {
    const size_t smem = (size_t)N * T * sizeof(cplx);
    const int ntile = (int)(inner / T);
    const unsigned nblk = (unsigned)(outer * ntile);
    const int nt = T * Reg3Cfg<N>::TPL;
    int rc;
    if (inv) {
        if ((rc = smem_attr(k_c2c_reg3<N, T, true>, smem))) return rc;
        k_c2c_reg3<N, T, true><<<nblk, nt, smem, fh_stream()>>>(data, data, tw, inner, ntile, 0, 1.0);
    } else {
        if ((rc = smem_attr(k_c2c_reg3<N, T, false>, smem))) return rc;
        k_c2c_reg3<N, T, false><<<nblk, nt, smem, fh_stream()>>>(data, data, tw, inner, ntile, 0, 1.0);
    }
    FH_LAUNCH_CHECK();
    return FH_OK;
}

static int launch_c2c_fast(int N, const cplx* tw, cplx* data, int64_t outer, int64_t inner, bool inv, int col0 = 0,
                           int ncols = 0) {
    static const int reg3 = env_int("FH_REG3", 1);  // register-resident three-pass kernels for 512 / 1024 (0: generic smem)
    if (reg3 && N == 512) return launch_c2c_reg3_NT<512, 8>(tw, data, outer, inner, inv);
    if (reg3 && N == 1024) return launch_c2c_reg3_NT<1024, 8>(tw, data, outer, inner, inv);
    switch (N) {
        case 16: return launch_c2c_gen3_NT<16, 8>(tw, data, outer, inner, inv);
        case 32: return launch_c2c_gen3_NT<32, 8>(tw, data, outer, inner, inv);
        case 512: return launch_c2c_gen3_NT<512, 8>(tw, data, outer, inner, inv);
        case 1024: return launch_c2c_gen3_NT<1024, 8>(tw, data, outer, inner, inv);
        case 2048: return launch_c2c_gen3_NT<2048, 4>(tw, data, outer, inner, inv);
        case 64: return launch_c2c_fast_NT<64, 8>(tw, data, outer, inner, inv, 1.0, col0, ncols);
        case 128: return launch_c2c_fast_NT<128, 8>(tw, data, outer, inner, inv, 1.0, col0, ncols);
        case 256: return launch_c2c_fast_NT<256, 8>(tw, data, outer, inner, inv, 1.0, col0, ncols);
    }
    return fh_set_error(FH_ERR_UNSUPPORTED, "no fast strided kernel for N=%d", N);
}

template <int N, int T, int KIND, int DIM>
static int launch_mid_fast_NT(fh_ga* op) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    const fh_plan* p = op->plan;
    const int64_t inner = (DIM == 3) ? (int64_t)op->n1l * op->pitch : op->pitch;
    const size_t smem = (size_t)D * N * T * sizeof(cplx);
    int rc;
    if ((rc = smem_attr(k_mid_green_fast<N, T, KIND, DIM>, smem))) return rc;
    k_mid_green_fast<N, T, KIND, DIM><<<(unsigned)(inner / T), D * T * FastCfg<N>::TPL, smem, fh_stream()>>>(
        op->specT, p->ax[0].tw, op->g, inner, p->nh, op->pitch);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

template <int N, int T, int KIND, int DIM>
static int launch_mid_pipe_NT(fh_ga* op) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    const fh_plan* p = op->plan;
    const int64_t inner = (DIM == 3) ? (int64_t)op->n1l * op->pitch : op->pitch;
    const size_t smem = (size_t)2 * D * (N + N / 16) * T * sizeof(cplx);
    const int tpr = (op->cur_ncols ? op->cur_ncols : op->pitch) / T;
    const int nrow = (DIM == 3) ? op->n1l : 1;
    const int ntiles = nrow * tpr;
    int rc;
    if ((rc = smem_attr(k_mid_green_pipe<N, T, KIND, DIM>, smem))) return rc;
    const int grid = ntiles < fh_num_sms() ? ntiles : fh_num_sms();
    k_mid_green_pipe<N, T, KIND, DIM><<<grid, D * T * FastCfg<N>::TPL, smem, fh_stream()>>>(
        op->specT, p->ax[0].tw, op->g, inner, p->nh, op->pitch, ntiles, tpr, op->cur_col0);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

template <int N, int T, int KIND, int DIM, int CR>
static int launch_mid_2r_NT(fh_ga* op) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    const fh_plan* p = op->plan;
    const int64_t inner = (DIM == 3) ? (int64_t)op->n1l * op->pitch : op->pitch;
    const size_t smem = (size_t)D * (N + N / 16) * T * sizeof(cplx);
    const int tpr = (op->cur_ncols ? op->cur_ncols : op->pitch) / T;
    const int nrow = (DIM == 3) ? op->n1l : 1;
    int rc;
    if ((rc = smem_attr(k_mid_green_2r<N, T, KIND, DIM, CR>, smem))) return rc;
    k_mid_green_2r<N, T, KIND, DIM, CR><<<(unsigned)(nrow * tpr), (D / CR) * T * FastCfg<N>::TPL, smem, fh_stream()>>>(
        op->specT, p->ax[0].tw, op->g, inner, p->nh, op->pitch, tpr, op->cur_col0);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

template <int N, int T, int KIND, int DIM>
static int launch_mid_gen3_NT(fh_ga* op) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    const fh_plan* p = op->plan;
    const int64_t inner = (DIM == 3) ? (int64_t)op->n1l * op->pitch : op->pitch;
    const size_t smem = (size_t)(N + N / 16) * D * T * sizeof(cplx);
    if (smem > (size_t)fh_max_smem_optin())
        return fh_set_error(FH_ERR_UNSUPPORTED, "axis-0 pass: N0=%d D=%d does not fit shared memory", N, D);
    int rc;
    if ((rc = smem_attr(k_mid_green_gen3<N, T, KIND, DIM>, smem))) return rc;
    k_mid_green_gen3<N, T, KIND, DIM><<<(unsigned)(inner / T), 384, smem, fh_stream()>>>(op->specT, p->ax[0].tw, op->g,
                                                                                      inner, p->nh, op->pitch);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

template <int KIND, int DIM>
static int launch_mid_fast(fh_ga* op) {
    const int N = op->plan->N[0];
    const int T = op->mid_T;
    // 8-column tiles, six shared-memory passes, register prefetch (fh_mid2.cuh): 64 / 128 / 256 in 3-D
    if (DIM == 3 && fh_mid2_len(N) && op->mid_pipe == 1) {
        const int64_t inner = (int64_t)op->n1l * op->pitch;
        const int ncols = op->cur_ncols ? op->cur_ncols : op->pitch;
        if (op->pitch % 8 == 0 && ncols % 8 == 0 && op->cur_col0 % 8 == 0)
            return fh_mid2_green(N, KIND, op->specT, op->plan->ax[0].tw, op->g, NULL, inner, (int64_t)N * inner,
                                 op->pitch, 0, op->plan->nh, op->n1l, op->cur_col0, ncols);
    }
    if (use_reg3() && fh_reg3_mid_len(N)) {  // three-pass register kernels (fh_reg3.cu decides which lengths)
        const int64_t inner = (DIM == 3) ? (int64_t)op->n1l * op->pitch : op->pitch;
        if (inner % 8 == 0)
            return fh_reg3_mid_green(N, KIND, DIM, op->specT, op->plan->ax[0].tw, op->g, inner, op->plan->nh, op->pitch);
    }
    if (N == 16) return launch_mid_gen3_NT<16, 4, KIND, DIM>(op);
    if (N == 32) return launch_mid_gen3_NT<32, 4, KIND, DIM>(op);
    if (N == 512) return launch_mid_gen3_NT<512, 2, KIND, DIM>(op);
    if (N == 1024) return launch_mid_gen3_NT<1024, 2, KIND, DIM>(op);
    if (N == 2048) return launch_mid_gen3_NT<2048, 1, KIND, DIM>(op);
    if (op->mid_pipe == 9 && N == 256) {  // debugging: data movement only
        constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
        const fh_plan* p = op->plan;
        const int64_t inner = (DIM == 3) ? (int64_t)op->n1l * op->pitch : op->pitch;
        const size_t smem = (size_t)D * 256 * 4 * sizeof(cplx);
        int rc;
        if ((rc = smem_attr(k_mid_copy_only<256, 4, D>, smem))) return rc;
        k_mid_copy_only<256, 4, D><<<(unsigned)(inner / 4), D * 4 * 16, smem, fh_stream()>>>(op->specT, inner);
        FH_LAUNCH_CHECK();
        return FH_OK;
    }
    if (op->mid_pipe == 3 && N == 256) {  // 8 columns per tile (128-byte segments), one 384-thread CTA per SM
        if constexpr (KIND == FH_GREEN_ELASTIC && DIM == 3) return launch_mid_2r_NT<256, 8, KIND, DIM, 2>(op);
    }
    if (op->mid_pipe == 2 && T == 4) {
        constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
        constexpr int CR = (D % 2 == 0) ? 2 : (D % 3 == 0 ? 3 : 1);
        if (env_int("FH_MID_CR1", 0) && N == 256) return launch_mid_2r_NT<256, 4, KIND, DIM, 1>(op);
        if (N == 64) return launch_mid_2r_NT<64, 4, KIND, DIM, CR>(op);
        if (N == 128) return launch_mid_2r_NT<128, 4, KIND, DIM, CR>(op);
        if (N == 256) return launch_mid_2r_NT<256, 4, KIND, DIM, CR>(op);
    }
    if (op->mid_pipe == 1 && T == 4) {
        constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
        if ((size_t)2 * D * (N + N / 16) * 4 * sizeof(cplx) <= (size_t)fh_max_smem_optin()) {
            if (N == 64) return launch_mid_pipe_NT<64, 4, KIND, DIM>(op);
            if (N == 128) return launch_mid_pipe_NT<128, 4, KIND, DIM>(op);
            if (N == 256) return launch_mid_pipe_NT<256, 4, KIND, DIM>(op);
        }
    }
#define FH_MID_CASE(n, t) \
    if (N == n && T == t) return launch_mid_fast_NT<n, t, KIND, DIM>(op);
    FH_MID_CASE(64, 2)
    FH_MID_CASE(64, 4)
    FH_MID_CASE(128, 2)
    FH_MID_CASE(128, 4)
    FH_MID_CASE(256, 2)
    FH_MID_CASE(256, 4)
#undef FH_MID_CASE
    return fh_set_error(FH_ERR_UNSUPPORTED, "no fast Green kernel for N0=%d T=%d", N, T);
}

template <int N, int D, int TRW, int ALAY>
static int launch_fwd_last_NTA(fh_ga* op, double* p, const double* r, int pupdate) {
    constexpr int NP = D * TRW / 2, NPAD = N + N / 16;
    const size_t smem = (size_t)2 * NP * NPAD * sizeof(double);
    const RowRange rr_ = row_range(op, TRW, false);
    const unsigned nblk = rr_.nblk;
    const int nt = NP * FastCfg<N>::TPL;
    const fh_plan* pl = op->plan;
    int rc;
    if ((rc = smem_attr(k_fwd_last_fast<N, D, TRW, ALAY>, smem))) return rc;
    k_fwd_last_fast<N, D, TRW, ALAY><<<nblk, nt, smem, fh_stream()>>>(
        op->A + rr_.ro, op->phase ? op->phase + rr_.ro : nullptr, op->lut, op->lutc, op->nphase, p + rr_.ro,
        r ? r + rr_.ro : nullptr, op->scal, pupdate, op->spec + rr_.so, pl->ax[pl->dim - 1].tw, op->nrows, pl->nh,
        op->pitch, op->xacc ? op->xacc + rr_.ro : nullptr);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

template <int N, int D, int TRW>
static int launch_fwd_last_NT(fh_ga* op, double* p, const double* r, int pupdate, bool withA) {
    if (!withA) return launch_fwd_last_NTA<N, D, TRW, -1>(op, p, r, pupdate);
    if (op->a_mode == 2 && op->nphase <= 2 && D * D <= 36) return launch_fwd_last_NTA<N, D, TRW, 3>(op, p, r, pupdate);
    if (op->a_mode == 2) return launch_fwd_last_NTA<N, D, TRW, 2>(op, p, r, pupdate);
    if (op->a_mode == 1) return launch_fwd_last_NTA<N, D, TRW, 1>(op, p, r, pupdate);
    return launch_fwd_last_NTA<N, D, TRW, 0>(op, p, r, pupdate);
}

template <int N, int D, int TRW>
static int launch_inv_last_NT(fh_ga* op, double* y, const double* pdot, int* npart) {
    constexpr int NP = D * TRW / 2, NPAD = N + N / 16;
    const size_t smem = (size_t)2 * NP * NPAD * sizeof(double);
    const RowRange rr_ = row_range(op, TRW, false);
    const unsigned nblk = rr_.nblk;
    const int nt = NP * FastCfg<N>::TPL;
    const fh_plan* pl = op->plan;
    int rc;
    if ((rc = smem_attr(k_inv_last_fast<N, D, TRW>, smem))) return rc;
    if (pdot && rr_.pb + nblk > GA_MAXPART) return fh_set_error(FH_ERR_UNSUPPORTED, "too many partial sums (%u)", nblk);
    k_inv_last_fast<N, D, TRW><<<nblk, nt, smem, fh_stream()>>>(op->spec + rr_.so, y + rr_.ro, pdot ? pdot + rr_.ro : nullptr, op->part + rr_.pb, pl->ax[pl->dim - 1].tw,
                                                                op->nrows, pl->nh, op->pitch,
                                                                1.0 / (double)pl->nreal);
    FH_LAUNCH_CHECK();
    if (npart) *npart = (int)(rr_.pb + nblk);
    return FH_OK;
}

// ------------------------------------------------------------------ run-time-length kernels (fh_fast.cuh, RtPlan)
static bool make_rt_plan(const AxisDesc& ax, RtPlan& P) {
    if (ax.nfac < 1 || ax.nfac > 3) return false;
    P.n = ax.n;
    P.ns = ax.nfac;
    int nb = ax.n, ts = 1;
    for (int s = 0; s < 3; ++s) {
        P.R[s] = P.NB[s] = P.TS[s] = 1;
    }
    for (int s = 0; s < ax.nfac; ++s) {
        const int R = ax.fac[s];
        const bool okR = (R == 2 || R == 3 || R == 4 || R == 5 || R == 7 || R == 8 || R == 9 || R == 11 || R == 13 ||
                          R == 15 || R == 16 || R == 17 || R == 19);
        if (!okR) return false;
        P.R[s] = R;
        P.NB[s] = nb;
        P.TS[s] = ts;
        nb /= R;
        ts *= R;
    }
    P.npr = ax.n + ax.n / 16 + 1;
    return true;
}

static int launch_c2c_rt(fh_ga* op, int axis, cplx* data, int64_t outer, int64_t inner, bool inv) {
    const RtPlan& P = op->rt[axis];
    const cplx* tw = op->plan->ax[axis].tw;
    int rc;
    static const int wantT = env_int("FH_RT_T", 16);
    if (wantT == 16 && inner % 16 == 0 && (size_t)P.npr * 16 * sizeof(cplx) <= (size_t)fh_max_smem_optin() / 2) {
        constexpr int T = 16;
        const size_t smem = (size_t)P.npr * T * sizeof(cplx);
        const int ntile = (int)(inner / T);
        const unsigned nblk = (unsigned)(outer * ntile);
        if (inv) {
            if ((rc = smem_attr(k_c2c_rt<T, true>, smem))) return rc;
            k_c2c_rt<T, true><<<nblk, 256, smem, fh_stream()>>>(data, data, tw, P, inner, ntile);
        } else {
            if ((rc = smem_attr(k_c2c_rt<T, false>, smem))) return rc;
            k_c2c_rt<T, false><<<nblk, 256, smem, fh_stream()>>>(data, data, tw, P, inner, ntile);
        }
        FH_LAUNCH_CHECK();
        return FH_OK;
    }
    constexpr int T = 8;
    const size_t smem = (size_t)P.npr * T * sizeof(cplx);
    if (smem > (size_t)fh_max_smem_optin()) return fh_set_error(FH_ERR_UNSUPPORTED, "axis length %d too large", P.n);
    const int ntile = (int)(inner / T);
    const unsigned nblk = (unsigned)(outer * ntile);
    if (inv) {
        if ((rc = smem_attr(k_c2c_rt<T, true>, smem))) return rc;
        k_c2c_rt<T, true><<<nblk, 256, smem, fh_stream()>>>(data, data, tw, P, inner, ntile);
    } else {
        if ((rc = smem_attr(k_c2c_rt<T, false>, smem))) return rc;
        k_c2c_rt<T, false><<<nblk, 256, smem, fh_stream()>>>(data, data, tw, P, inner, ntile);
    }
    FH_LAUNCH_CHECK();
    return FH_OK;
}

template <int T, int KIND, int DIM>
static int launch_mid_rt_T(fh_ga* op) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    const fh_plan* p = op->plan;
    const RtPlan& P = op->rt[0];
    const int64_t inner = (DIM == 3) ? (int64_t)op->n1l * op->pitch : op->pitch;
    const size_t smem = (size_t)P.npr * D * T * sizeof(cplx);
    int rc;
    if ((rc = smem_attr(k_mid_green_rt<T, KIND, DIM>, smem))) return rc;
    k_mid_green_rt<T, KIND, DIM><<<(unsigned)(inner / T), 384, smem, fh_stream()>>>(op->specT, p->ax[0].tw, P, op->g,
                                                                                 inner, p->nh, op->pitch);
    FH_LAUNCH_CHECK();
    return FH_OK;
}
template <int KIND, int DIM>
static int launch_mid_rt(fh_ga* op) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    const size_t per_line = (size_t)op->rt[0].npr * D * sizeof(cplx);
    const size_t cap = (size_t)fh_max_smem_optin();
    if (4 * per_line <= cap / 2 || (4 * per_line <= cap && 2 * per_line > cap / 2)) return launch_mid_rt_T<4, KIND, DIM>(op);
    if (2 * per_line <= cap) return launch_mid_rt_T<2, KIND, DIM>(op);
    if (per_line <= cap) return launch_mid_rt_T<1, KIND, DIM>(op);
    return fh_set_error(FH_ERR_UNSUPPORTED, "axis-0 pass: N0=%d D=%d does not fit shared memory", op->rt[0].n, D);
}

template <int D, int TRW, int ALAY>
static int launch_fwd_last_rtA(fh_ga* op, double* p, const double* r, int pupdate) {
    constexpr int NP = D * TRW / 2;
    const fh_plan* pl = op->plan;
    const int ax = pl->dim - 1;
    const RtPlan& P = op->rt[ax];
    const size_t smem = (size_t)P.npr * NP * sizeof(cplx);
    const RowRange rr_ = row_range(op, TRW, true);
    const unsigned nblk = rr_.nblk;
    int rc;
    if ((rc = smem_attr(k_fwd_last_rt<D, TRW, ALAY>, smem))) return rc;
    k_fwd_last_rt<D, TRW, ALAY><<<nblk, 256, smem, fh_stream()>>>(
        op->A + rr_.ro, op->phase ? op->phase + rr_.ro : nullptr, op->lut, op->lutc, op->nphase, p + rr_.ro,
        r ? r + rr_.ro : nullptr, op->scal, pupdate, op->spec + rr_.so, pl->ax[ax].tw, P, op->nrows, pl->nh,
        op->pitch, op->xacc ? op->xacc + rr_.ro : nullptr);
    FH_LAUNCH_CHECK();
    return FH_OK;
}
template <int D, int TRW>
static int launch_fwd_last_rtD(fh_ga* op, double* p, const double* r, int pupdate) {
    if (op->a_mode == 2 && op->nphase <= 2 && D * D <= 36) return launch_fwd_last_rtA<D, TRW, 3>(op, p, r, pupdate);
    if (op->a_mode == 2) return launch_fwd_last_rtA<D, TRW, 2>(op, p, r, pupdate);
    if (op->a_mode == 1) return launch_fwd_last_rtA<D, TRW, 1>(op, p, r, pupdate);
    return launch_fwd_last_rtA<D, TRW, 0>(op, p, r, pupdate);
}
// rows per CTA: enough lines (D*TRW/2) to keep 256 threads busy with the few wide butterflies of an odd length,
// falling back to the small tiles when the padded lines do not fit shared memory twice
static bool rt_wide(const fh_ga* op, int NPwide) {
    static const int want = env_int("FH_RT_WIDE", 1);
    return want && (size_t)op->rt[op->plan->dim - 1].npr * NPwide * sizeof(cplx) <= (size_t)fh_max_smem_optin() / 2;
}
static int launch_fwd_last_rt(fh_ga* op, double* p, const double* r, int pupdate) {
    switch (op->D) {
        case 6: return rt_wide(op, 18) ? launch_fwd_last_rtD<6, 6>(op, p, r, pupdate) : launch_fwd_last_rtD<6, 2>(op, p, r, pupdate);
        case 3: return rt_wide(op, 12) ? launch_fwd_last_rtD<3, 8>(op, p, r, pupdate) : launch_fwd_last_rtD<3, 4>(op, p, r, pupdate);
        case 2: return rt_wide(op, 12) ? launch_fwd_last_rtD<2, 12>(op, p, r, pupdate) : launch_fwd_last_rtD<2, 4>(op, p, r, pupdate);
    }
    return fh_set_error(FH_ERR_UNSUPPORTED, "fused operator: D=%d", op->D);
}
template <int D, int TRW>
static int launch_inv_last_rtD(fh_ga* op, double* y, const double* pdot, int* npart) {
    constexpr int NP = D * TRW / 2;
    const fh_plan* pl = op->plan;
    const int ax = pl->dim - 1;
    const RtPlan& P = op->rt[ax];
    const size_t smem = (size_t)P.npr * NP * sizeof(cplx);
    const RowRange rr_ = row_range(op, TRW, true);
    const unsigned nblk = rr_.nblk;
    int rc;
    if ((rc = smem_attr(k_inv_last_rt<D, TRW>, smem))) return rc;
    if (pdot && rr_.pb + nblk > GA_MAXPART) return fh_set_error(FH_ERR_UNSUPPORTED, "too many partial sums (%u)", nblk);
    k_inv_last_rt<D, TRW><<<nblk, 256, smem, fh_stream()>>>(op->spec + rr_.so, y + rr_.ro, pdot ? pdot + rr_.ro : nullptr, op->part + rr_.pb, pl->ax[ax].tw, P, op->nrows,
                                                            pl->nh, op->pitch, 1.0 / (double)pl->nreal);
    FH_LAUNCH_CHECK();
    if (npart) *npart = (int)(rr_.pb + nblk);
    return FH_OK;
}
static int launch_inv_last_rt(fh_ga* op, double* y, const double* pdot, int* npart) {
    switch (op->D) {
        case 6: return rt_wide(op, 18) ? launch_inv_last_rtD<6, 6>(op, y, pdot, npart) : launch_inv_last_rtD<6, 2>(op, y, pdot, npart);
        case 3: return rt_wide(op, 12) ? launch_inv_last_rtD<3, 8>(op, y, pdot, npart) : launch_inv_last_rtD<3, 4>(op, y, pdot, npart);
        case 2: return rt_wide(op, 12) ? launch_inv_last_rtD<2, 12>(op, y, pdot, npart) : launch_inv_last_rtD<2, 4>(op, y, pdot, npart);
    }
    return fh_set_error(FH_ERR_UNSUPPORTED, "fused operator: D=%d", op->D);
}

template <int N, int D, int TRW, int ALAY>
static int launch_fwd_last_g3A(fh_ga* op, double* p, const double* r, int pupdate) {
    constexpr int NP = D * TRW / 2;
    const size_t smem = (size_t)(N + N / 16) * NP * sizeof(cplx);
    const RowRange rr_ = row_range(op, TRW, false);
    const unsigned nblk = rr_.nblk;
    const fh_plan* pl = op->plan;
    int rc;
    if ((rc = smem_attr(k_fwd_last_gen3<N, D, TRW, ALAY>, smem))) return rc;
    k_fwd_last_gen3<N, D, TRW, ALAY><<<nblk, 256, smem, fh_stream()>>>(
        op->A + rr_.ro, op->phase ? op->phase + rr_.ro : nullptr, op->lut, op->lutc, op->nphase, p + rr_.ro,
        r ? r + rr_.ro : nullptr, op->scal, pupdate, op->spec + rr_.so, pl->ax[pl->dim - 1].tw, op->nrows, pl->nh,
        op->pitch, op->xacc ? op->xacc + rr_.ro : nullptr);
    FH_LAUNCH_CHECK();
    return FH_OK;
}
template <int N, int D, int TRW>
static int launch_fwd_last_g3(fh_ga* op, double* p, const double* r, int pupdate, bool withA) {
    if (!withA) return launch_fwd_last_g3A<N, D, TRW, -1>(op, p, r, pupdate);
    if (op->a_mode == 2 && op->nphase <= 2 && D * D <= 36) return launch_fwd_last_g3A<N, D, TRW, 3>(op, p, r, pupdate);
    if (op->a_mode == 2) return launch_fwd_last_g3A<N, D, TRW, 2>(op, p, r, pupdate);
    if (op->a_mode == 1) return launch_fwd_last_g3A<N, D, TRW, 1>(op, p, r, pupdate);
    return launch_fwd_last_g3A<N, D, TRW, 0>(op, p, r, pupdate);
}
template <int N, int D, int TRW>
static int launch_inv_last_g3(fh_ga* op, double* y, const double* pdot, int* npart) {
    constexpr int NP = D * TRW / 2;
    const size_t smem = (size_t)(N + N / 16) * NP * sizeof(cplx);
    const RowRange rr_ = row_range(op, TRW, false);
    const unsigned nblk = rr_.nblk;
    const fh_plan* pl = op->plan;
    int rc;
    if ((rc = smem_attr(k_inv_last_gen3<N, D, TRW>, smem))) return rc;
    if (pdot && rr_.pb + nblk > GA_MAXPART) return fh_set_error(FH_ERR_UNSUPPORTED, "too many partial sums (%u)", nblk);
    k_inv_last_gen3<N, D, TRW><<<nblk, 256, smem, fh_stream()>>>(op->spec + rr_.so, y + rr_.ro, pdot ? pdot + rr_.ro : nullptr, op->part + rr_.pb, pl->ax[pl->dim - 1].tw,
                                                                 op->nrows, pl->nh, op->pitch, 1.0 / (double)pl->nreal);
    FH_LAUNCH_CHECK();
    if (npart) *npart = (int)(rr_.pb + nblk);
    return FH_OK;
}

// generic-routine sizes: TRW = 2 rows per CTA (D even) or 4 (D = 3)
#define FH_LAST_DISPATCH_G3(FN, ...)                                             \
    do {                                                                         \
        const int N_ = op->plan->N[op->plan->dim - 1];                           \
        const int D_ = op->D;                                                    \
        if (D_ == 6) {                                                           \
            if (N_ == 16) return FN<16, 6, 2>(__VA_ARGS__);                      \
            if (N_ == 32) return FN<32, 6, 2>(__VA_ARGS__);                      \
            if (N_ == 512) return FN<512, 6, 2>(__VA_ARGS__);                    \
            if (N_ == 1024) return FN<1024, 6, 2>(__VA_ARGS__);                  \
            if (N_ == 2048) return FN<2048, 6, 2>(__VA_ARGS__);                  \
        } else if (D_ == 3) {                                                    \
            if (N_ == 16) return FN<16, 3, 4>(__VA_ARGS__);                      \
            if (N_ == 32) return FN<32, 3, 4>(__VA_ARGS__);                      \
            if (N_ == 512) return FN<512, 3, 4>(__VA_ARGS__);                    \
            if (N_ == 1024) return FN<1024, 3, 4>(__VA_ARGS__);                  \
            if (N_ == 2048) return FN<2048, 3, 4>(__VA_ARGS__);                  \
        } else if (D_ == 2) {                                                    \
            if (N_ == 16) return FN<16, 2, 4>(__VA_ARGS__);                      \
            if (N_ == 32) return FN<32, 2, 4>(__VA_ARGS__);                      \
            if (N_ == 512) return FN<512, 2, 4>(__VA_ARGS__);                    \
            if (N_ == 1024) return FN<1024, 2, 4>(__VA_ARGS__);                  \
            if (N_ == 2048) return FN<2048, 2, 4>(__VA_ARGS__);                  \
        }                                                                        \
        return fh_set_error(FH_ERR_UNSUPPORTED, "no generic-routine last-axis kernel"); \
    } while (0)
static int launch_fwd_last_gen3(fh_ga* op, double* p, const double* r, int pupdate, bool withA) {
    FH_LAST_DISPATCH_G3(launch_fwd_last_g3, op, p, r, pupdate, withA);
}
static int launch_inv_last_gen3(fh_ga* op, double* y, const double* pdot, int* npart) {
    FH_LAST_DISPATCH_G3(launch_inv_last_g3, op, y, pdot, npart);
}

#define FH_LAST_DISPATCH(FN, ...)                                            \
    do {                                                                     \
        const int N_ = op->plan->N[op->plan->dim - 1];                       \
        const int D_ = op->D;                                                \
        if (D_ == 6 && op->trw == 2) {                                       \
            if (N_ == 64) return FN<64, 6, 2>(__VA_ARGS__);                  \
            if (N_ == 128) return FN<128, 6, 2>(__VA_ARGS__);                \
            if (N_ == 256) return FN<256, 6, 2>(__VA_ARGS__);                \
        } else if (D_ == 6) {                                                \
            if (N_ == 64) return FN<64, 6, 4>(__VA_ARGS__);                  \
            if (N_ == 128) return FN<128, 6, 4>(__VA_ARGS__);                \
            if (N_ == 256) return FN<256, 6, 4>(__VA_ARGS__);                \
        } else if (D_ == 3) {                                                \
            if (N_ == 64) return FN<64, 3, 8>(__VA_ARGS__);                  \
            if (N_ == 128) return FN<128, 3, 8>(__VA_ARGS__);                \
            if (N_ == 256) return FN<256, 3, 8>(__VA_ARGS__);                \
        } else if (D_ == 2) {                                                \
            if (N_ == 64) return FN<64, 2, 8>(__VA_ARGS__);                  \
            if (N_ == 128) return FN<128, 2, 8>(__VA_ARGS__);                \
            if (N_ == 256) return FN<256, 2, 8>(__VA_ARGS__);                \
        }                                                                    \
        return fh_set_error(FH_ERR_UNSUPPORTED, "no fast last-axis kernel"); \
    } while (0)

// register-resident three-pass kernels (fh_reg3.cu) for the lengths they cover; FH_REG3=0 falls back to the
// generic shared-memory routine
static int launch_fwd_last_reg3(fh_ga* op, double* p, const double* r, int pupdate, bool withA) {
    const fh_plan* pl = op->plan;
    const RowRange rr_ = row_range(op, op->trw, false);
    Reg3LastArgs a;
    a.A = op->A + rr_.ro;
    a.phase = op->phase ? op->phase + rr_.ro : nullptr;
    a.lut = op->lut;
    a.lutc = &op->lutc;
    a.nphase = op->nphase;
    a.alay = !withA ? -1 : (op->a_mode == 2 ? ((op->nphase <= 2 && op->D * op->D <= 36) ? 3 : 2) : (op->a_mode == 1 ? 1 : 0));
    a.p = p + rr_.ro;
    a.r = r ? r + rr_.ro : nullptr;
    a.scal = op->scal;
    a.pupdate = pupdate;
    a.spec = op->spec + rr_.so;
    a.tw = pl->ax[pl->dim - 1].tw;
    a.nrows = op->nrows;
    a.nh = pl->nh;
    a.pitch = op->pitch;
    a.xacc = op->xacc ? op->xacc + rr_.ro : nullptr;
    a.nblk = rr_.nblk;
    return fh_reg3_fwd_last(pl->N[pl->dim - 1], op->D, op->trw, a);
}
static int launch_inv_last_reg3(fh_ga* op, double* y, const double* pdot, int* npart) {
    const fh_plan* pl = op->plan;
    const RowRange rr_ = row_range(op, op->trw, false);
    if (pdot && rr_.pb + rr_.nblk > GA_MAXPART)
        return fh_set_error(FH_ERR_UNSUPPORTED, "too many partial sums (%u)", rr_.nblk);
    Reg3InvArgs a;
    a.spec = op->spec + rr_.so;
    a.y = y + rr_.ro;
    a.pdot = pdot ? pdot + rr_.ro : nullptr;
    a.part = op->part + rr_.pb;
    a.tw = pl->ax[pl->dim - 1].tw;
    a.nrows = op->nrows;
    a.nh = pl->nh;
    a.pitch = op->pitch;
    a.scale = 1.0 / (double)pl->nreal;
    a.nblk = rr_.nblk;
    int rc;
    if ((rc = fh_reg3_inv_last(pl->N[pl->dim - 1], op->D, op->trw, a))) return rc;
    if (npart) *npart = (int)(rr_.pb + rr_.nblk);
    return FH_OK;
}
static bool reg3_last_ok(const fh_ga* op) {
    const int nl = op->plan->N[op->plan->dim - 1];
    return use_reg3() && fh_reg3_last_len(nl) && (op->D == 6 || op->D == 3 || op->D == 2) &&
           op->trw == ((op->D == 6) ? 2 : 4);
}
static int launch_fwd_last_two_pass(fh_ga* op, double* p, const double* r, int pupdate, bool withA) {
    FH_LAST_DISPATCH(launch_fwd_last_NT, op, p, r, pupdate, withA);
}
static int launch_fwd_last_fast(fh_ga* op, double* p, const double* r, int pupdate, bool withA) {
    if (reg3_last_ok(op)) return launch_fwd_last_reg3(op, p, r, pupdate, withA);
    if (fh_gen3_len(op->plan->N[op->plan->dim - 1])) return launch_fwd_last_gen3(op, p, r, pupdate, withA);
    // S1 and S5 choose their rows-per-CTA independently (the spectrum layout does not depend on it).  Measured at
    // 256^3: the plain S1 (operator application) runs best with 2 rows per CTA (0.419 vs 0.449 ms), its CG form
    // (x and p updates folded in) and S5 with 4 (0.947 vs 0.963 ms)
    const int keep = op->trw;
    if (op->trw_s1 && !pupdate) op->trw = op->trw_s1;
    const int rc = launch_fwd_last_two_pass(op, p, r, pupdate, withA);
    op->trw = keep;
    return rc;
}
static int launch_inv_last_fast(fh_ga* op, double* y, const double* pdot, int* npart) {
    if (reg3_last_ok(op)) return launch_inv_last_reg3(op, y, pdot, npart);
    if (fh_gen3_len(op->plan->N[op->plan->dim - 1])) return launch_inv_last_gen3(op, y, pdot, npart);
    FH_LAST_DISPATCH(launch_inv_last_NT, op, y, pdot, npart);
}
static int trw_for(int D, int nlast) {
    if (fh_gen3_len(nlast)) return D == 6 ? 2 : 4;
    return D == 6 ? 4 : 8;
}

// ------------------------------------------------------------------ coefficient analysis (once per operator)
// exact symmetry check: flag != 0 if any A_ij != A_ji
__global__ void k_sym_check(int64_t n, int D, const double* __restrict__ A, int* __restrict__ flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride)
        for (int i = 0; i < D; ++i)
            for (int j = i + 1; j < D; ++j)
                if (A[((size_t)i * D + j) * n + v] != A[((size_t)j * D + i) * n + v]) *flag = 1;
}
// phase[v] = index of the table matrix that equals A[:, :, v] exactly; voxels matching none
// report the smallest such index through atomicMin
__global__ void k_phase_match(int64_t n, int DD, const double* __restrict__ A, const double* __restrict__ lut,
                              int nph, unsigned char* __restrict__ phase, unsigned long long* __restrict__ first) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        int found = -1;
        for (int m = 0; m < nph && found < 0; ++m) {
            bool eq = true;
            for (int e = 0; e < DD && eq; ++e) eq = (A[(size_t)e * n + v] == lut[m * DD + e]);
            if (eq) found = m;
        }
        if (found >= 0)
            phase[v] = (unsigned char)found;
        else
            atomicMin(first, (unsigned long long)v);
    }
}
__global__ void k_lut_fetch(int64_t n, int DD, const double* __restrict__ A, int64_t v, double* __restrict__ dst) {
    const int e = threadIdx.x;
    if (e < DD) dst[e] = A[(size_t)e * n + v];
}

#define FH_MAX_PHASES 16
static int analyse_coefficients(fh_ga* op) {
    const int D = op->D, DD = D * D;
    const int64_t n = op->nloc;
    const int want = env_int("FH_AMODE", -1);  // -1 auto, 0 full, 1 sym, 2 phase table
    op->a_mode = 0;
    op->phase = NULL;
    op->lut = NULL;
    op->nphase = 0;
    const bool rt_last = op->rt_ok[op->plan->dim - 1];
    if (want == 0 || !(op->fast_last || rt_last)) return FH_OK;
    const unsigned grid = (unsigned)(fh_num_sms() * 8);
    if ((want < 0 || want == 2) && (n % 2 == 0 || !op->fast_last)) {
        unsigned long long* first = NULL;
        FH_CUDA(cudaMalloc((void**)&first, sizeof(unsigned long long)));
        FH_CUDA(cudaMalloc((void**)&op->phase, (size_t)n));
        FH_CUDA(cudaMalloc((void**)&op->lut, sizeof(double) * FH_MAX_PHASES * DD));
        int nph = 0;
        bool ok = false;
        for (int round = 0; round <= FH_MAX_PHASES; ++round) {
            const unsigned long long none = ~0ULL;
            FH_CUDA(cudaMemcpyAsync(first, &none, sizeof(none), cudaMemcpyHostToDevice, fh_stream()));
            k_phase_match<<<grid, 256, 0, fh_stream()>>>(n, DD, op->A, op->lut, nph, op->phase, first);
            FH_LAUNCH_CHECK();
            unsigned long long f = 0;
            FH_CUDA(cudaMemcpyAsync(&f, first, sizeof(f), cudaMemcpyDeviceToHost, fh_stream()));
            FH_CUDA(cudaStreamSynchronize(fh_stream()));
            if (f == none) {
                ok = true;
                break;
            }
            if (nph == FH_MAX_PHASES) break;
            k_lut_fetch<<<1, 64, 0, fh_stream()>>>(n, DD, op->A, (int64_t)f, op->lut + (size_t)nph * DD);
            FH_LAUNCH_CHECK();
            ++nph;
        }
        cudaFree(first);
        if (ok) {
            op->a_mode = 2;
            op->nphase = nph;
            memset(&op->lutc, 0, sizeof(op->lutc));
            if (nph <= 2 && DD <= 36)
                FH_CUDA(cudaMemcpy(&op->lutc.c[0][0], op->lut, sizeof(double) * DD, cudaMemcpyDeviceToHost));
            if (nph == 2 && DD <= 36)
                FH_CUDA(cudaMemcpy(&op->lutc.c[1][0], op->lut + DD, sizeof(double) * DD, cudaMemcpyDeviceToHost));
            return FH_OK;
        }
        cudaFree(op->phase);
        cudaFree(op->lut);
        op->phase = NULL;
        op->lut = NULL;
    }
    if (want < 0 || want == 1) {
        int* flag = NULL;
        FH_CUDA(cudaMalloc((void**)&flag, sizeof(int)));
        FH_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), fh_stream()));
        k_sym_check<<<grid, 256, 0, fh_stream()>>>(n, D, op->A, flag);
        FH_LAUNCH_CHECK();
        int h = 1;
        FH_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, fh_stream()));
        FH_CUDA(cudaStreamSynchronize(fh_stream()));
        cudaFree(flag);
        if (h == 0) op->a_mode = 1;
    }
    return FH_OK;
}

// ------------------------------------------------------------------ operator object
static int pitch_for(const fh_plan* p) { return (p->nh + 7) / 8 * 8; }
static int64_t round16(int64_t v) { return (v + 15) / 16 * 16; }

// work = [sigma: D*nloc (rounded)] [spec: D*n0l*N1*pitch complex] [specT: D*N0*n1l*pitch complex, slabs only]
static int64_t work_doubles(const fh_plan* p, int D, int n0l, int n1l) {
    const int d = p->dim;
    const int64_t rows_other = (d == 3) ? p->N[1] : 1;
    const int64_t nloc = (int64_t)n0l * rows_other * p->N[d - 1];
    int64_t w = round16((int64_t)D * nloc) + 2 * (int64_t)D * n0l * rows_other * pitch_for(p);
    if (d == 3 && (n0l != p->N[0] || n1l != p->N[1])) w += 2 * (int64_t)D * p->N[0] * n1l * pitch_for(p);
    return w;
}

extern "C" int64_t fh_ga_work_doubles(const fh_plan* p, int D) {
    if (!p || D < 1) return 0;
    return work_doubles(p, D, p->N[0], p->dim == 3 ? p->N[1] : 1);
}
extern "C" int64_t fh_ga_slab_work_doubles(const fh_plan* p, int D, int n0_local, int n1_local) {
    if (!p || D < 1 || p->dim != 3) return 0;
    return work_doubles(p, D, n0_local, n1_local);
}

static int ga_create(fh_ga** out, const fh_plan* plan, int D, const double* A, int a_layout, const fh_green* g,
                     double* work, int n0l, int n1l, int n1_off) {
    FH_REQUIRE(out && plan && A && g && work, "fh_ga_create: null argument");
    FH_REQUIRE(plan->dim == 2 || plan->dim == 3, "fh_ga_create: dim must be 2 or 3");
    FH_REQUIRE(a_layout == 0, "fh_ga_create: unsupported coefficient layout %d", a_layout);
    FH_REQUIRE(((uintptr_t)work & 15) == 0 && ((uintptr_t)A & 15) == 0, "fh_ga_create: buffers must be 16-byte aligned");
    const int Dexp = (g->kind == FH_GREEN_SCALAR) ? plan->dim : plan->dim * (plan->dim + 1) / 2;
    FH_REQUIRE(D == Dexp, "fh_ga_create: D=%d does not match Green kind %d in dim %d", D, g->kind, plan->dim);
    for (int a = 0; a < plan->dim; ++a)
        FH_REQUIRE(g->N[a] == plan->N[a], "fh_ga_create: Green descriptor grid differs from the plan grid");
    const int d = plan->dim;
    const bool slab = (d == 3) && (n0l != plan->N[0] || n1l != plan->N[1]);
    FH_REQUIRE(n0l >= 1 && n0l <= plan->N[0] && n1l >= 1 && n1_off >= 0 && (d == 2 || n1_off + n1l <= plan->N[1]),
               "fh_ga_create: bad slab extents");
    fh_ga* op = (fh_ga*)calloc(1, sizeof(fh_ga));
    if (!op) return fh_set_error(FH_ERR_ALLOC, "fh_ga_create: out of host memory");
    int rc = fh_fill_green(op->g, g);
    if (rc) {
        free(op);
        return rc;
    }
    op->g.ioff1 = n1_off;
    op->plan = plan;
    op->D = D;
    op->A = A;
    op->a_layout = a_layout;
    op->pitch = pitch_for(plan);
    op->n0l = n0l;
    op->n1l = (d == 3) ? n1l : 1;
    op->nrows = (d == 3) ? (int64_t)n0l * plan->N[1] : n0l;
    op->nloc = op->nrows * plan->N[d - 1];
    op->nspecp = op->nrows * op->pitch;
    op->work = work;
    op->sigma = work;
    op->spec = (cplx*)(work + round16((int64_t)D * op->nloc));
    op->specT = slab ? op->spec + (size_t)D * op->nspecp : op->spec;
    const int use_fast = env_int("FH_FAST", 1);
    op->trw = trw_for(D, plan->N[d - 1]);
    if (D == 6 && !fh_gen3_len(plan->N[d - 1]) && env_int("FH_TRW", 4) == 2) op->trw = 2;
    auto pow2fast = [](int n) { return fh_fast_len(n) || fh_gen3_len(n); };
    op->fast_last = use_fast && pow2fast(plan->N[d - 1]) && (op->nrows % op->trw == 0);
    op->trw_s1 = (D == 6 && op->trw == 4 && !fh_gen3_len(plan->N[d - 1]) && env_int("FH_TRW_S1", 2) == 2) ? 2 : 0;
    op->fast_mid1 = use_fast && d == 3 && pow2fast(plan->N[1]);
    op->fast_mid0 = use_fast && pow2fast(plan->N[0]) && (((int64_t)op->n1l * op->pitch) % 4 == 0);
    const int use_rt = env_int("FH_RT", 7);
    for (int a = 0; a < 3; ++a) op->rt_ok[a] = false;
    // FH_RT mask: bit 0 = axis 0 (S3), bit 1 = middle axis (S2/S4), bit 2 = last axis (S1/S5)
    for (int a = 0; a < d; ++a) {
        const int bit = (a == 0) ? 1 : (a == d - 1 ? 4 : 2);
        op->rt_ok[a] = use_fast && (use_rt & bit) && make_rt_plan(plan->ax[a], op->rt[a]);
    }
    if (op->rt_ok[d - 1] && (size_t)op->rt[d - 1].npr * 6 * sizeof(cplx) > (size_t)fh_max_smem_optin()) op->rt_ok[d - 1] = false;
    if (d == 3 && op->rt_ok[1] && (size_t)op->rt[1].npr * 8 * sizeof(cplx) > (size_t)fh_max_smem_optin()) op->rt_ok[1] = false;
    if (op->rt_ok[0] && ((size_t)op->rt[0].npr * D * sizeof(cplx) > (size_t)fh_max_smem_optin() ||
                         ((int64_t)op->n1l * op->pitch) % 4 != 0))
        op->rt_ok[0] = false;
    // compile-time odd-length family (3-D, whole grid on this GPU): per axis, ahead of the run-time-length kernels
    for (int a = 0; a < 3; ++a)
        op->odd_ax[a] = use_fast && fh_odd_on() && d == 3 && !slab && (D == 3 || D == 6) && fh_odd_len(plan->N[a]);
    op->mid_T = env_int("FH_MID_T", 4);
    op->mid_pipe = env_int("FH_MID_PIPE", 1);
    if (op->mid_T != 2 && op->mid_T != 4) op->mid_T = 4;
    // L2 blocking (3-D, all-fast, not slab-decomposed): chunk = FH_CHUNK columns (default 8 = one S2 tile)
    op->chunk_cols = 0;
    if (d == 3 && !slab && op->fast_mid0 && op->fast_mid1 && fh_fast_len(plan->N[0]) && fh_fast_len(plan->N[1]) &&
        op->mid_T == 4 && op->mid_pipe != 0 && op->mid_pipe != 9) {
        int cc = env_int("FH_CHUNK", 0);  // off by default: measured slower than whole-row launches
        if (cc > 0 && cc % 8 == 0 && cc < op->pitch) op->chunk_cols = cc;
    }
    if ((size_t)D * plan->N[0] * op->mid_T * sizeof(cplx) > (size_t)fh_max_smem_optin()) op->mid_T = 2;
    cudaError_t e = cudaMalloc((void**)&op->scal, sizeof(double) * (16 + GA_MAXPART));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&op->pinned, sizeof(double) * 16);
    if (e == cudaSuccess) e = cudaMemset(op->scal, 0, sizeof(double) * (16 + GA_MAXPART));
    // padding columns of the spectrum rows are never read as data; zero them once for determinism
    if (e == cudaSuccess) e = cudaMemset(op->spec, 0, sizeof(cplx) * D * op->nspecp);
    if (e == cudaSuccess && slab)
        e = cudaMemset(op->specT, 0, sizeof(cplx) * D * (size_t)plan->N[0] * op->n1l * op->pitch);
    if (e != cudaSuccess) {
        free(op);
        return fh_set_error(FH_ERR_CUDA, "fh_ga_create: %s", cudaGetErrorString(e));
    }
    op->part = op->scal + 16;
    if ((rc = analyse_coefficients(op))) {
        fh_ga_destroy(op);
        return rc;
    }
    *out = op;
    return FH_OK;
}

extern "C" int fh_ga_create(fh_ga** out, const fh_plan* plan, int D, const double* A, int a_layout, const fh_green* g,
                            double* work) {
    FH_REQUIRE(plan, "fh_ga_create: null plan");
    return ga_create(out, plan, D, A, a_layout, g, work, plan->N[0], plan->dim == 3 ? plan->N[1] : 1, 0);
}

// Slab-decomposed operator of one rank (3-D): real-space fields hold n0_local planes of axis 0
// (A, x, y are the LOCAL arrays [.][n0_local][N1][N2]); the axis-0 pass (S3) runs on the transposed
// spectrum [D][N0][n1_local][pitch] holding the global axis-1 indices [n1_offset, n1_offset+n1_local).
// The caller moves the spectrum between the two layouts (all-to-all) between S2/S3 and S3/S4 and
// reduces the CG scalars across ranks; see ffthompy_b200/slab.py.
extern "C" int fh_ga_create_slab(fh_ga** out, const fh_plan* plan, int D, const double* A_local, int a_layout,
                                 const fh_green* g, double* work, int n0_local, int n1_local, int n1_offset) {
    FH_REQUIRE(plan && plan->dim == 3, "fh_ga_create_slab: a 3-D plan is required");
    return ga_create(out, plan, D, A_local, a_layout, g, work, n0_local, n1_local, n1_offset);
}

// device pointers of the two spectrum layouts (x-slab [D][n0l][N1][pitch], y-slab [D][N0][n1l][pitch])
extern "C" int fh_ga_buffers(const fh_ga* op, void** spec, void** specT, int* pitch) {
    FH_REQUIRE(op, "fh_ga_buffers: null argument");
    if (spec) *spec = op->spec;
    if (specT) *specT = op->specT;
    if (pitch) *pitch = op->pitch;
    return FH_OK;
}

extern "C" int fh_ga_destroy(fh_ga* op) {
    if (!op) return FH_OK;
    if (op->phase) cudaFree(op->phase);
    if (op->lut) cudaFree(op->lut);
    if (op->sd_off1) cudaFree(op->sd_off1);
    if (op->sp_off1) cudaFree(op->sp_off1);
    if (op->kb_off) cudaFree(op->kb_off);
    cudaFree(op->scal);
    cudaFreeHost(op->pinned);
    free(op);
    return FH_OK;
}

// which kernels an operator uses: bit0 fast last axis, bit1 fast axis 1, bit2 fast axis 0;
// bits 4-5 coefficient mode (0 full, 1 symmetric, 2 phase table), bits 8.. number of phases,
// bits 16-18 run-time-length kernels (last / middle / first axis), bits 20-22 compile-time odd-length kernels
extern "C" int fh_ga_config(const fh_ga* op, int* flags, int* pitch, int* mid_T) {
    FH_REQUIRE(op, "fh_ga_config: null argument");
    if (flags)
        *flags = (op->fast_last ? 1 : 0) | (op->fast_mid1 ? 2 : 0) | (op->fast_mid0 ? 4 : 0) | (op->a_mode << 4) |
                 (op->nphase << 8) |
                 ((!op->fast_last && !op->odd_ax[op->plan->dim - 1] && op->rt_ok[op->plan->dim - 1]) ? 1 << 16 : 0) |
                 ((op->plan->dim == 3 && !op->fast_mid1 && !op->odd_ax[1] && op->rt_ok[1]) ? 1 << 17 : 0) |
                 ((!op->fast_mid0 && !op->odd_ax[0] && op->rt_ok[0]) ? 1 << 18 : 0) |
                 (op->odd_ax[op->plan->dim - 1] ? 1 << 20 : 0) | ((op->plan->dim == 3 && op->odd_ax[1]) ? 1 << 21 : 0) |
                 (op->odd_ax[0] ? 1 << 22 : 0);
    if (pitch) *pitch = op->pitch;
    if (mid_T) *mid_T = op->mid_T;
    return FH_OK;
}

static unsigned ga_grid(int64_t n) {
    int64_t b = fh_ceil_div(n, GA_NT);
    const int64_t cap = (int64_t)fh_num_sms() * 8;
    if (b > cap) b = cap;
    if (b > GA_MAXPART) b = GA_MAXPART;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// ------------------------------------------------------------------ Krylov kernels
// scal[0]=rr  scal[1]=pAp  scal[2]=alpha  scal[3]=beta  scal[4]=norm_res  (all with the
// 1/prod(N) of Tensor.scalar_product, tensors/objects.py:635)
__global__ void k_dot_part(int64_t n, const double* __restrict__ x, const double* __restrict__ y,
                           double* __restrict__ part) {
    __shared__ double red[32];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += x[i] * y[i];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

// r = b - ax ; p = r ; partial r.r          (solver.py:115-118)
__global__ void k_cg_init(int64_t n, const double* __restrict__ b, const double* __restrict__ ax,
                          double* __restrict__ r, double* __restrict__ p, double* __restrict__ part) {
    __shared__ double red[32];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = b[i] - ax[i];
        r[i] = v;
        if (p) p[i] = v;
        acc += v * v;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

// mode 0: rr = sum/prodN, norm = sqrt(rr)                          (solver.py:118-120)
// mode 1: pAp = sum/prodN, alpha = rr/pAp                          (solver.py:126)
// mode 2: rrn = sum/prodN, beta = rrn/rr, rr = rrn, norm = sqrt(rr) (solver.py:129-133)
__global__ void k_cg_scal(int np, const double* __restrict__ part, double* __restrict__ scal, double inv_prodN,
                          int mode) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) acc += part[i];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) {
        const double v = acc * inv_prodN;
        if (mode == 0) {
            scal[0] = v;
            scal[4] = sqrt(v);
        } else if (mode == 1) {
            scal[1] = v;
            scal[2] = scal[0] / v;
        } else {
            scal[3] = v / scal[0];
            scal[0] = v;
            scal[4] = sqrt(v);
        }
    }
}

// x += alpha p ; r -= alpha Ap ; partial r.r   (solver.py:127-129); 16-byte accesses
__global__ void k_cg_update(int64_t n2, double2* __restrict__ x, double2* __restrict__ r, const double2* __restrict__ p,
                            const double2* __restrict__ Ap, const double* __restrict__ scal,
                            double* __restrict__ part) {
    __shared__ double red[32];
    const double alpha = scal[2];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 xv = x[i], pv = p[i], rv = r[i], av = Ap[i];
        x[i] = make_double2(xv.x + alpha * pv.x, xv.y + alpha * pv.y);
        const double2 v = make_double2(rv.x - alpha * av.x, rv.y - alpha * av.y);
        r[i] = v;
        acc += v.x * v.x;
        acc += v.y * v.y;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
__global__ void k_cg_update1(int64_t n, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                             const double* __restrict__ Ap, const double* __restrict__ scal,
                             double* __restrict__ part) {
    __shared__ double red[32];
    const double alpha = scal[2];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        x[i] = x[i] + alpha * p[i];
        const double v = r[i] - alpha * Ap[i];
        r[i] = v;
        acc += v * v;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

// deferred-x variant of the update: r -= alpha Ap ; partial r.r  (x += alpha p is applied by the next S1,
// which holds p in registers anyway: one full-field read less per iteration)
__global__ void k_cg_update_r(int64_t n2, double2* __restrict__ r, const double2* __restrict__ Ap,
                              const double* __restrict__ scal, double* __restrict__ part) {
    __shared__ double red[32];
    const double alpha = scal[2];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 rv = r[i], av = Ap[i];
        const double2 v = make_double2(rv.x - alpha * av.x, rv.y - alpha * av.y);
        r[i] = v;
        acc += v.x * v.x;
        acc += v.y * v.y;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
// the same two kernels for an odd element count (odd grids: prod(N) odd and D odd), where the fields r | p | Ap of
// `vecs` are only 8-byte aligned: 8-byte accesses, four independent elements per thread and step
__global__ void k_cg_update_r1(int64_t n, double* __restrict__ r, const double* __restrict__ Ap,
                               const double* __restrict__ scal, double* __restrict__ part) {
    __shared__ double red[32];
    const double alpha = scal[2];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        double rv[4], av[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            rv[u] = r[i + u * stride];
            av[u] = Ap[i + u * stride];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const double v = rv[u] - alpha * av[u];
            r[i + u * stride] = v;
            acc += v * v;
        }
    }
    for (; i < n; i += stride) {
        const double v = r[i] - alpha * Ap[i];
        r[i] = v;
        acc += v * v;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
__global__ void k_cg_xflush1(int64_t n, double* __restrict__ x, const double* __restrict__ p,
                             const double* __restrict__ scal) {
    const double alpha = scal[2];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = x[i] + alpha * p[i];
}
// x += alpha p  (flush of the deferred update after the last iteration: scal[2] and p are still that iteration's)
__global__ void k_cg_xflush(int64_t n2, double2* __restrict__ x, const double2* __restrict__ p,
                            const double* __restrict__ scal) {
    const double alpha = scal[2];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        const double2 xv = x[i], pv = p[i];
        x[i] = make_double2(xv.x + alpha * pv.x, xv.y + alpha * pv.y);
    }
}

// p = r + beta p   (solver.py:132)
__global__ void k_cg_pupdate(int64_t n, double* __restrict__ p, const double* __restrict__ r,
                             const double* __restrict__ scal) {
    const double beta = scal[3];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = r[i] + beta * p[i];
}

// x += omega * res  (res = b - Ax), partial res.res     (solver.py:72-74)
__global__ void k_rich_update(int64_t n, double* __restrict__ x, const double* __restrict__ b,
                              const double* __restrict__ ax, double omega, double* __restrict__ part) {
    __shared__ double red[32];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double v = b[i] - ax[i];
        x[i] = x[i] + omega * v;
        acc += v * v;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

// ------------------------------------------------------------------ the pipeline, stage by stage
// stage 1..5 as in the header comment.  `x` is the operand (for the CG loop: p, updated in place
// from r when pupdate != 0).  After stage 5 with dot != 0, op->part[0..*npart) holds the partial
// sums of <x, y>.
static int ga_stage(fh_ga* op, int stage, double* x, const double* r, int pupdate, double* y, int dot, int* npart) {
    const fh_plan* p = op->plan;
    const int D = op->D, d = p->dim;
    const int64_t n = op->nloc;  // local voxels per component
    const int64_t nlines = (int64_t)D * op->nrows;
    int rc;
    switch (stage) {
        case 1:
            if (op->odd_ax[d - 1] && !op->row_cnt) return fh_odd_fwd_last(op, x, r, pupdate);
            if (op->fast_last) return launch_fwd_last_fast(op, x, r, pupdate, true);
            if (op->rt_ok[d - 1]) return launch_fwd_last_rt(op, x, r, pupdate);
            if (pupdate) {
                k_cg_pupdate<<<ga_grid(D * n), GA_NT, 0, fh_stream()>>>(D * n, x, r, op->scal);
                FH_LAUNCH_CHECK();
            }
            switch (D) {
                case 2: k_apply_A<2><<<ga_grid(n), GA_NT, 0, fh_stream()>>>(n, op->A, x, op->sigma); break;
                case 3: k_apply_A<3><<<ga_grid(n), GA_NT, 0, fh_stream()>>>(n, op->A, x, op->sigma); break;
                case 6: k_apply_A<6><<<ga_grid(n), GA_NT, 0, fh_stream()>>>(n, op->A, x, op->sigma); break;
                default: return fh_set_error(FH_ERR_UNSUPPORTED, "fused operator: D=%d", D);
            }
            FH_LAUNCH_CHECK();
            return fh_launch_r2c_last(p, op->sigma, op->spec, nlines, op->pitch);
        case 2:
            if (d != 3) return FH_OK;
            if (op->odd_ax[1]) return fh_odd_c2c(p->N[1], p->ax[1].tw, op->spec, (int64_t)D * op->n0l, op->pitch, false);
            if (op->fast_mid1) return launch_c2c_fast(p->N[1], p->ax[1].tw, op->spec, (int64_t)D * op->n0l, op->pitch, false);
            if (op->rt_ok[1]) return launch_c2c_rt(op, 1, op->spec, (int64_t)D * op->n0l, op->pitch, false);
            return fh_launch_c2c_strided(p->ax[1], op->spec, op->spec, (int64_t)D * op->n0l, op->pitch, false, 1.0);
        case 3:
            if (op->odd_ax[0]) return fh_odd_mid(op);
            if (op->fast_mid0) {
                if (op->g.kind == FH_GREEN_SCALAR)
                    return (d == 3) ? launch_mid_fast<FH_GREEN_SCALAR, 3>(op) : launch_mid_fast<FH_GREEN_SCALAR, 2>(op);
                return (d == 3) ? launch_mid_fast<FH_GREEN_ELASTIC, 3>(op) : launch_mid_fast<FH_GREEN_ELASTIC, 2>(op);
            }
            if (op->rt_ok[0]) {
                if (op->g.kind == FH_GREEN_SCALAR)
                    return (d == 3) ? launch_mid_rt<FH_GREEN_SCALAR, 3>(op) : launch_mid_rt<FH_GREEN_SCALAR, 2>(op);
                return (d == 3) ? launch_mid_rt<FH_GREEN_ELASTIC, 3>(op) : launch_mid_rt<FH_GREEN_ELASTIC, 2>(op);
            }
            if (op->g.kind == FH_GREEN_SCALAR)
                return (d == 3) ? launch_mid_green_generic<FH_GREEN_SCALAR, 3>(op)
                                : launch_mid_green_generic<FH_GREEN_SCALAR, 2>(op);
            return (d == 3) ? launch_mid_green_generic<FH_GREEN_ELASTIC, 3>(op)
                            : launch_mid_green_generic<FH_GREEN_ELASTIC, 2>(op);
        case 4:
            if (d != 3) return FH_OK;
            if (op->odd_ax[1]) return fh_odd_c2c(p->N[1], p->ax[1].tw, op->spec, (int64_t)D * op->n0l, op->pitch, true);
            if (op->fast_mid1) return launch_c2c_fast(p->N[1], p->ax[1].tw, op->spec, (int64_t)D * op->n0l, op->pitch, true);
            if (op->rt_ok[1]) return launch_c2c_rt(op, 1, op->spec, (int64_t)D * op->n0l, op->pitch, true);
            return fh_launch_c2c_strided(p->ax[1], op->spec, op->spec, (int64_t)D * op->n0l, op->pitch, true, 1.0);
        case 5:
            if (op->odd_ax[d - 1] && !op->row_cnt) return fh_odd_inv_last(op, y, dot ? x : NULL, npart);
            if (op->fast_last) return launch_inv_last_fast(op, y, dot ? x : NULL, npart);
            // measured (255^3, 243^3): the generic batched C2R beats the run-time-length one; FH_RT bit 3 opts in
            if (op->rt_ok[d - 1] && (env_int("FH_RT", 7) & 8)) return launch_inv_last_rt(op, y, dot ? x : NULL, npart);
            if ((rc = fh_launch_c2r_last(p, op->spec, y, nlines, op->pitch, 1.0 / (double)p->nreal))) return rc;
            if (dot) {
                const unsigned g = ga_grid(D * n);
                k_dot_part<<<g, GA_NT, 0, fh_stream()>>>(D * n, x, y, op->part);
                FH_LAUNCH_CHECK();
                if (npart) *npart = (int)g;
            }
            return FH_OK;
    }
    return fh_set_error(FH_ERR_ARG, "fused operator: bad stage %d", stage);
}

// S2-S3-S4 over column chunks of the spectrum rows: each chunk (D*N0*N1*chunk_cols*16 B, ~50 MB at
// 256^3 elasticity) is transformed along axis 1, axis 0 (+G^) and back while it is resident in the
// 126 MB L2, so the three stages cost one HBM read and one HBM write of the spectrum instead of three.
static int ga_mid_chunked(fh_ga* op) {
    const fh_plan* p = op->plan;
    const int D = op->D;
    int rc = FH_OK;
    for (int c0 = 0; c0 < op->pitch && !rc; c0 += op->chunk_cols) {
        const int nc = (op->pitch - c0 < op->chunk_cols) ? op->pitch - c0 : op->chunk_cols;
        op->cur_col0 = c0;
        op->cur_ncols = nc;
        rc = launch_c2c_fast(p->N[1], p->ax[1].tw, op->spec, (int64_t)D * op->n0l, op->pitch, false, c0, nc);
        if (!rc)
            rc = (op->g.kind == FH_GREEN_SCALAR) ? launch_mid_fast<FH_GREEN_SCALAR, 3>(op)
                                                 : launch_mid_fast<FH_GREEN_ELASTIC, 3>(op);
        if (!rc) rc = launch_c2c_fast(p->N[1], p->ax[1].tw, op->spec, (int64_t)D * op->n0l, op->pitch, true, c0, nc);
    }
    op->cur_col0 = 0;
    op->cur_ncols = 0;
    return rc;
}

static int ga_matvec(fh_ga* op, double* x, const double* r, int pupdate, double* y, int dot, int* npart) {
    int rc;
    if (op->chunk_cols > 0) {
        if ((rc = ga_stage(op, 1, x, r, pupdate, y, dot, npart))) return rc;
        if ((rc = ga_mid_chunked(op))) return rc;
        return ga_stage(op, 5, x, r, pupdate, y, dot, npart);
    }
    for (int s = 1; s <= 5; ++s)
        if ((rc = ga_stage(op, s, x, r, pupdate, y, dot, npart))) return rc;
    return FH_OK;
}

extern "C" int fh_ga_apply(fh_ga* op, const double* x, double* y) {
    FH_REQUIRE(op && x && y, "fh_ga_apply: null argument");
    return ga_matvec(op, (double*)x, NULL, 0, y, 0, NULL);
}

// Run ONE stage of the pipeline (profiling / roofline accounting in bench.py); the spectrum
// workspace carries the state between stages.
extern "C" int fh_ga_stage(fh_ga* op, int stage, const double* x, double* y) {
    FH_REQUIRE(op && x && y, "fh_ga_stage: null argument");
    if (stage == 6) return op->chunk_cols > 0 ? ga_mid_chunked(op) : fh_set_error(FH_ERR_ARG, "stage 6: L2 blocking is off");
    int np = 0;
    const int rc = ga_stage(op, stage, (double*)x, NULL, 0, y, 1, &np);
    if (stage == 5 && !rc) op->last_npart = np;
    return rc;
}

// sum of the partial <x,y> left by the last stage-5 launch (this rank's part, not normalised)
__global__ void k_sum_part(int np, const double* __restrict__ part, double* __restrict__ out) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < np; i += blockDim.x) acc += part[i];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[0] = acc;
}
extern "C" int fh_ga_last_dot(fh_ga* op, double* result_host) {
    FH_REQUIRE(op && result_host, "fh_ga_last_dot: null argument");
    FH_REQUIRE(op->last_npart > 0, "fh_ga_last_dot: no stage-5 partial sums available");
    k_sum_part<<<1, GA_NT, 0, fh_stream()>>>(op->last_npart, op->part, op->scal + 8);
    FH_LAUNCH_CHECK();
    FH_CUDA(cudaMemcpyAsync(op->pinned, op->scal + 8, sizeof(double), cudaMemcpyDeviceToHost, fh_stream()));
    FH_CUDA(cudaStreamSynchronize(fh_stream()));
    *result_host = op->pinned[0];
    return FH_OK;
}

static int read_norm(fh_ga* op, double* out) {
    FH_CUDA(cudaMemcpyAsync(op->pinned, op->scal + 4, sizeof(double), cudaMemcpyDeviceToHost, fh_stream()));
    FH_CUDA(cudaStreamSynchronize(fh_stream()));
    *out = op->pinned[0];
    return FH_OK;
}

// ------------------------------------------------------------------ zero-copy, chunked slab pipeline
// Exchange buffers are chunk-major: chunk j (x-planes [j*n0c, (j+1)*n0c) of every rank) is one contiguous
// block [G][D][n0c][n1l][pitch] — on the x-slab side G indexes the peer that owns the k1 range, on the
// y-slab side the peer that owns the x-planes — so one all_to_all_single per chunk moves it with no
// pack/unpack pass, and chunk j's exchange overlaps the transforms of chunk j+1 (ffthompy_b200/slab.py).
template <int N, int T>
static int launch_c2c_map_NT(const cplx* tw, const cplx* in, cplx* out, const LineMap& mi, const LineMap& mo,
                             int64_t panels, int pitch, bool inv) {
    const size_t smem = (size_t)(N + N / 16) * T * sizeof(cplx);
    const int ntile = pitch / T;
    const unsigned nblk = (unsigned)(panels * ntile);
    int rc;
    if (inv) {
        if ((rc = smem_attr(k_c2c_map<N, T, true>, smem))) return rc;
        k_c2c_map<N, T, true><<<nblk, 256, smem, fh_stream()>>>(in, out, tw, mi, mo, ntile);
    } else {
        if ((rc = smem_attr(k_c2c_map<N, T, false>, smem))) return rc;
        k_c2c_map<N, T, false><<<nblk, 256, smem, fh_stream()>>>(in, out, tw, mi, mo, ntile);
    }
    FH_LAUNCH_CHECK();
    return FH_OK;
}
static int launch_c2c_map(int N, const cplx* tw, const cplx* in, cplx* out, const LineMap& mi, const LineMap& mo,
                          int64_t panels, int pitch, bool inv) {
    if (use_reg3() && fh_reg3_map_len(N) && pitch % 8 == 0) return fh_reg3_c2c_map(N, tw, in, out, mi, mo, panels, pitch, inv);
    switch (N) {
        case 16: return launch_c2c_map_NT<16, 8>(tw, in, out, mi, mo, panels, pitch, inv);
        case 32: return launch_c2c_map_NT<32, 8>(tw, in, out, mi, mo, panels, pitch, inv);
        case 64: return launch_c2c_map_NT<64, 8>(tw, in, out, mi, mo, panels, pitch, inv);
        case 128: return launch_c2c_map_NT<128, 8>(tw, in, out, mi, mo, panels, pitch, inv);
        case 256: return launch_c2c_map_NT<256, 8>(tw, in, out, mi, mo, panels, pitch, inv);
        case 512: return launch_c2c_map_NT<512, 8>(tw, in, out, mi, mo, panels, pitch, inv);
        case 1024: return launch_c2c_map_NT<1024, 8>(tw, in, out, mi, mo, panels, pitch, inv);
        case 2048: return launch_c2c_map_NT<2048, 4>(tw, in, out, mi, mo, panels, pitch, inv);
    }
    return fh_set_error(FH_ERR_UNSUPPORTED, "no slab-exchange kernel for N1=%d", N);
}
template <int N, int T, int CS, int KIND>
static int launch_mid_mapc(fh_ga* op, size_t smem) {
    const fh_plan* p = op->plan;
    const int64_t inner = (int64_t)op->n1l * op->pitch;
    int rc;
    if ((rc = smem_attr(k_mid_green_mapc<N, T, CS, KIND>, smem))) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(inner / T), 1, 1);
    cfg.blockDim = dim3(384, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = fh_stream();
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    FH_CUDA(cudaLaunchKernelEx(&cfg, k_mid_green_mapc<N, T, CS, KIND>, op->sd_peer ? op->spec : op->sd_bufB,
                               (const cplx*)p->ax[0].tw, op->g, (const int64_t*)op->sd_off0, op->sd_cs0, p->nh,
                               op->pitch));
    fh_count_launch();
    return FH_OK;
}
template <int N, int T, int CS, int KIND>
static int launch_mid_mapp(fh_ga* op, size_t smem) {
    const fh_plan* p = op->plan;
    const int64_t inner = (int64_t)op->n1l * op->pitch;
    const int64_t nseg = inner / (T * CS);
    int rc;
    if ((rc = smem_attr(k_mid_green_mapp<N, T, CS, KIND>, smem))) return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(CS * fh_num_sms()), 1, 1);
    cfg.blockDim = dim3(384, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = fh_stream();
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    static int ncl_cache[9] = {0};  // co-resident clusters of this size (same smem footprint for every N of a T class)
    int ncl = 0;
    FH_CUDA(cudaOccupancyMaxActiveClusters(&ncl, k_mid_green_mapp<N, T, CS, KIND>, &cfg));
    (void)ncl_cache;
    if (ncl < 1) return fh_set_error(FH_ERR_UNSUPPORTED, "axis-0 pass: no cluster of %d CTAs fits", CS);
    if ((int64_t)ncl > nseg) ncl = (int)nseg;
    cfg.gridDim = dim3((unsigned)(CS * ncl), 1, 1);
    FH_CUDA(cudaLaunchKernelEx(&cfg, k_mid_green_mapp<N, T, CS, KIND>, op->sd_peer ? op->spec : op->sd_bufB,
                               (const cplx*)p->ax[0].tw, op->g, (const int64_t*)op->sd_off0, op->sd_cs0, p->nh,
                               op->pitch, nseg));
    fh_count_launch();
    return FH_OK;
}
template <int N, int T, int KIND>
static int launch_mid_map_NT(fh_ga* op) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? 3 : 6;
    const fh_plan* p = op->plan;
    const int64_t inner = (int64_t)op->n1l * op->pitch;
    const size_t smem = (size_t)(N + N / 16) * D * T * sizeof(cplx);
    if (smem > (size_t)fh_max_smem_optin())
        return fh_set_error(FH_ERR_UNSUPPORTED, "axis-0 pass: N0=%d D=%d does not fit shared memory", N, D);
    if (!(op->sd_peer && op->sd_world > 1) && use_reg3() && fh_reg3_map_len(N) && inner % 8 == 0)
        return fh_reg3_mid_green_map(N, KIND, op->sd_peer ? op->spec : op->sd_bufB, p->ax[0].tw, op->g, inner, p->nh,
                                     op->pitch, op->sd_off0, op->sd_cs0);
    // rows in peer memory: move 128-byte (FH_XSEG=256: 256-byte) segments through a CTA cluster
    if (op->sd_peer && op->sd_world > 1) {
        static const int seg = env_int("FH_XSEG", 128);
        const int64_t ntile = inner / T;
        static const int xpipe = env_int("FH_XPIPE", 1);
        if constexpr (D * N * T <= 6144 && (T == 2 || T == 4)) {
            if (xpipe && ntile % (8 / T) == 0) return launch_mid_mapp<N, T, 8 / T, KIND>(op, smem);
        }
        if constexpr (T == 4) {
            if (seg >= 256 && ntile % 4 == 0) return launch_mid_mapc<N, T, 4, KIND>(op, smem);
            if (seg >= 128 && ntile % 2 == 0) return launch_mid_mapc<N, T, 2, KIND>(op, smem);
        } else if constexpr (T == 2) {
            if (seg >= 256 && ntile % 8 == 0) return launch_mid_mapc<N, T, 8, KIND>(op, smem);
            if (seg >= 128 && ntile % 4 == 0) return launch_mid_mapc<N, T, 4, KIND>(op, smem);
        } else {
            if (seg >= 128 && ntile % 8 == 0) return launch_mid_mapc<N, T, 8, KIND>(op, smem);
        }
    }
    int rc;
    if ((rc = smem_attr(k_mid_green_map<N, T, KIND, 3>, smem))) return rc;
    k_mid_green_map<N, T, KIND, 3><<<(unsigned)(inner / T), 384, smem, fh_stream()>>>(
        op->sd_peer ? op->spec : op->sd_bufB, p->ax[0].tw, op->g, op->sd_off0, op->sd_cs0, p->nh, op->pitch);
    FH_LAUNCH_CHECK();
    return FH_OK;
}
template <int KIND>
static int launch_mid_map(fh_ga* op) {
    switch (op->plan->N[0]) {
        case 16: return launch_mid_map_NT<16, 4, KIND>(op);
        case 32: return launch_mid_map_NT<32, 4, KIND>(op);
        case 64: return launch_mid_map_NT<64, 4, KIND>(op);
        case 128: return launch_mid_map_NT<128, 4, KIND>(op);
        case 256: return launch_mid_map_NT<256, 4, KIND>(op);
        case 512: return launch_mid_map_NT<512, 2, KIND>(op);
        case 1024: return launch_mid_map_NT<1024, 2, KIND>(op);
        case 2048: return launch_mid_map_NT<2048, 1, KIND>(op);
    }
    return fh_set_error(FH_ERR_UNSUPPORTED, "no slab-exchange kernel for N0=%d", op->plan->N[0]);
}

int fh_launch_c2c_map(int N, const cplx* tw, const cplx* in, cplx* out, const LineMap& mi, const LineMap& mo, int64_t panels,
                      int pitch, bool inv) {
    return launch_c2c_map(N, tw, in, out, mi, mo, panels, pitch, inv);
}
int fh_ga_stage_local(fh_ga* op, int stage, double* x, const double* r, int pupdate, double* y, int dot, int* npart) {
    const int rc = ga_stage(op, stage, x, r, pupdate, y, dot, npart);
    if (stage == 5 && !rc && npart) op->last_npart = *npart;
    return rc;
}

extern "C" int fh_ga_slab_direct(fh_ga* op, int world, int nchunk, void* bufA, void* bufB) {
    FH_REQUIRE(op && bufA && bufB && bufA != bufB, "fh_ga_slab_direct: null or aliased buffers");
    const fh_plan* p = op->plan;
    FH_REQUIRE(p->dim == 3 && world >= 1 && nchunk >= 1, "fh_ga_slab_direct: a 3-D slab operator is required");
    FH_REQUIRE((int64_t)op->n0l * world == p->N[0] && (int64_t)op->n1l * world == p->N[1],
               "fh_ga_slab_direct: slab extents do not match world=%d", world);
    FH_REQUIRE(op->n0l % nchunk == 0, "fh_ga_slab_direct: %d local planes do not split into %d chunks", op->n0l, nchunk);
    FH_REQUIRE(((uintptr_t)bufA & 15) == 0 && ((uintptr_t)bufB & 15) == 0, "fh_ga_slab_direct: buffers must be 16-byte aligned");
    if (!fh_map_len(p->N[0]) || !fh_map_len(p->N[1]))
        return fh_set_error(FH_ERR_UNSUPPORTED, "fh_ga_slab_direct: N0=%d, N1=%d not in the slab-exchange kernel family",
                            p->N[0], p->N[1]);
    const int n0c = op->n0l / nchunk;
    const int64_t rows_c = (int64_t)n0c * p->N[1];
    if (nchunk > 1) {
        // the run-time-length S5 honours row ranges only in its in-place form (FH_RT bit 3); the generic batched C2R
        // would redo every chunk's rows on each call
        const bool rt_last = !op->fast_last && op->rt_ok[2] && (env_int("FH_RT", 7) & 8);
        if (!(op->fast_last && rows_c % op->trw == 0) && !(rt_last && rows_c % 24 == 0))
            return fh_set_error(FH_ERR_UNSUPPORTED, "fh_ga_slab_direct: the last-axis kernels cannot run %lld-row chunks",
                                (long long)rows_c);
    }
    const int D = op->D, P = op->pitch, n1l = op->n1l, n0l = op->n0l;
    const int64_t inner = (int64_t)n1l * P;
    int64_t* h1 = (int64_t*)malloc(sizeof(int64_t) * p->N[1]);
    int64_t* h0 = (int64_t*)malloc(sizeof(int64_t) * p->N[0]);
    if (!h1 || !h0) {
        free(h1);
        free(h0);
        return fh_set_error(FH_ERR_ALLOC, "fh_ga_slab_direct: out of host memory");
    }
    for (int k1 = 0; k1 < p->N[1]; ++k1) h1[k1] = (int64_t)(k1 / n1l) * D * n0c * inner + (int64_t)(k1 % n1l) * P;
    for (int i0 = 0; i0 < p->N[0]; ++i0) {
        const int g = i0 / n0l, rem = i0 % n0l, j = rem / n0c, i0c = rem % n0c;
        h0[i0] = ((int64_t)(j * world + g) * D) * n0c * inner + (int64_t)i0c * inner;
    }
    if (op->sd_off1) cudaFree(op->sd_off1);
    op->sd_off1 = op->sd_off0 = NULL;
    cudaError_t e = cudaMalloc((void**)&op->sd_off1, sizeof(int64_t) * (p->N[0] + p->N[1]));
    if (e == cudaSuccess) {
        op->sd_off0 = op->sd_off1 + p->N[1];
        e = cudaMemcpy(op->sd_off1, h1, sizeof(int64_t) * p->N[1], cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) e = cudaMemcpy(op->sd_off0, h0, sizeof(int64_t) * p->N[0], cudaMemcpyHostToDevice);
    free(h1);
    free(h0);
    if (e != cudaSuccess) return fh_set_error(FH_ERR_CUDA, "fh_ga_slab_direct: %s", cudaGetErrorString(e));
    op->sd_world = world;
    op->sd_nchunk = nchunk;
    op->sd_n0c = n0c;
    op->sd_cs0 = (int64_t)n0c * inner;
    op->sd_peer = 0;
    op->sd_bufA = (cplx*)bufA;
    op->sd_bufB = (cplx*)bufB;
    // padding columns travel with the rows and must be zero.  The caller hands in zero-filled buffers: with peer-mapped
    // (symmetric-memory) buffers a fill enqueued here could run after a faster rank's first push into this buffer and
    // wipe it, so nothing is written to the exchange buffers in this call.
    return FH_OK;
}

// Fused axis-0 pass + exchange over peer memory: every rank keeps its x-slab half spectrum
// [D][n0l][N1][pitch] in memory its peers can address (NVLink peer mappings, e.g. torch symmetric memory);
// S3 of rank q gathers row i0 of its k1 range straight from the owner of plane i0 (remote loads),
// applies G^ and scatters the result back to the same place (remote stores).  No exchange buffer, no
// all-to-all: the transfer rides inside the kernel's own load/store phases.  The caller puts a device
// barrier across the ranks before and after stage 3.  peer_spec[g] = rank g's spectrum base (fh_ga_buffers)
// as mapped into THIS process.
extern "C" int fh_ga_slab_peer(fh_ga* op, int world, int rank, const void* const* peer_spec) {
    FH_REQUIRE(op && peer_spec && world >= 1 && rank >= 0 && rank < world, "fh_ga_slab_peer: bad argument");
    const fh_plan* p = op->plan;
    FH_REQUIRE(p->dim == 3, "fh_ga_slab_peer: a 3-D slab operator is required");
    FH_REQUIRE((int64_t)op->n0l * world == p->N[0] && (int64_t)op->n1l * world == p->N[1],
               "fh_ga_slab_peer: slab extents do not match world=%d", world);
    FH_REQUIRE(op->g.ioff1 == rank * op->n1l, "fh_ga_slab_peer: rank %d does not own the k1 range of this operator", rank);
    FH_REQUIRE(peer_spec[rank] == (const void*)op->spec, "fh_ga_slab_peer: peer_spec[rank] must be this operator's spectrum");
    if (!fh_map_len(p->N[0]))
        return fh_set_error(FH_ERR_UNSUPPORTED, "fh_ga_slab_peer: N0=%d not in the exchange kernel family", p->N[0]);
    const int P = op->pitch, n0l = op->n0l, N1 = p->N[1];
    int64_t* h0 = (int64_t*)malloc(sizeof(int64_t) * p->N[0]);
    if (!h0) return fh_set_error(FH_ERR_ALLOC, "fh_ga_slab_peer: out of host memory");
    for (int i0 = 0; i0 < p->N[0]; ++i0) {
        const int g = i0 / n0l, i0l = i0 % n0l;
        const intptr_t delta = (intptr_t)peer_spec[g] - (intptr_t)op->spec;
        if (!peer_spec[g] || delta % (intptr_t)sizeof(cplx)) {
            free(h0);
            return fh_set_error(FH_ERR_ARG, "fh_ga_slab_peer: peer %d spectrum pointer is null or misaligned", g);
        }
        h0[i0] = (int64_t)(delta / (intptr_t)sizeof(cplx)) + ((int64_t)i0l * N1 + (int64_t)rank * op->n1l) * P;
    }
    if (op->sd_off1) cudaFree(op->sd_off1);
    op->sd_off1 = op->sd_off0 = NULL;
    cudaError_t e = cudaMalloc((void**)&op->sd_off1, sizeof(int64_t) * p->N[0]);
    if (e == cudaSuccess) e = cudaMemcpy(op->sd_off1, h0, sizeof(int64_t) * p->N[0], cudaMemcpyHostToDevice);
    free(h0);
    if (e != cudaSuccess) return fh_set_error(FH_ERR_CUDA, "fh_ga_slab_peer: %s", cudaGetErrorString(e));
    op->sd_off0 = op->sd_off1;
    op->sd_cs0 = (int64_t)n0l * N1 * P;
    op->sd_world = world;
    op->sd_nchunk = 1;
    op->sd_n0c = n0l;
    op->sd_peer = 1;
    op->sd_bufA = op->sd_bufB = NULL;
    return FH_OK;
}

// One step of the slab pipeline.  Without fh_ga_slab_direct: `stage` = 1..5 is the plain pipeline stage
// (the caller packs/unpacks around its exchange), `chunk` is ignored.  With it:
//   stage 1: S1 (sigma = A p, with p = r + beta p when pupdate) + S2 of chunk -> exchange buffer A
//   stage 3: S3 in place on exchange buffer B (all chunks)
//   stage 4: S4 of chunk from exchange buffer A + S5 (y rows of the chunk, partial sums of <p, y>)
extern "C" int fh_ga_slab_stage(fh_ga* op, int stage, int chunk, double* p, const double* r, int pupdate, double* y) {
    FH_REQUIRE(op && p && y, "fh_ga_slab_stage: null argument");
    int np = 0, rc;
    if (!op->sd_world || op->sd_peer) {
        if (op->sd_peer && stage == 3)
            return (op->g.kind == FH_GREEN_SCALAR) ? launch_mid_map<FH_GREEN_SCALAR>(op) : launch_mid_map<FH_GREEN_ELASTIC>(op);
        rc = ga_stage(op, stage, p, r, pupdate, y, 1, &np);
        if (stage == 5 && !rc) op->last_npart = np;
        return rc;
    }
    const fh_plan* pl = op->plan;
    const int D = op->D, P = op->pitch, N1 = pl->N[1], n0c = op->sd_n0c;
    FH_REQUIRE(chunk >= 0 && chunk < op->sd_nchunk, "fh_ga_slab_stage: chunk %d out of range", chunk);
    const int64_t inner = (int64_t)op->n1l * P;
    const int64_t chunk_elems = (int64_t)op->sd_world * D * n0c * inner;
    const LineMap nat = {NULL, (int64_t)P, (int64_t)op->n0l * N1 * P, (int64_t)N1 * P, n0c};
    const LineMap blk = {op->sd_off1, 0, (int64_t)n0c * inner, inner, n0c};
    cplx* spec_c = op->spec + (size_t)chunk * n0c * N1 * P;
    cplx* bufA_c = op->sd_bufA + (size_t)chunk * chunk_elems;
    if (op->sd_nchunk > 1) {
        op->row_beg = (int64_t)chunk * n0c * N1;
        op->row_cnt = (int64_t)n0c * N1;
    }
    rc = FH_OK;
    if (stage == 1) {
        rc = ga_stage(op, 1, p, r, pupdate, y, 0, NULL);
        if (!rc) rc = launch_c2c_map(N1, pl->ax[1].tw, spec_c, bufA_c, nat, blk, (int64_t)D * n0c, P, false);
    } else if (stage == 3) {
        rc = (op->g.kind == FH_GREEN_SCALAR) ? launch_mid_map<FH_GREEN_SCALAR>(op) : launch_mid_map<FH_GREEN_ELASTIC>(op);
        op->last_npart = 0;
    } else if (stage == 4) {
        rc = launch_c2c_map(N1, pl->ax[1].tw, bufA_c, spec_c, blk, nat, (int64_t)D * n0c, P, true);
        if (!rc) rc = ga_stage(op, 5, p, NULL, 0, y, 1, &np);
        if (!rc && np > op->last_npart) op->last_npart = np;
    } else {
        rc = fh_set_error(FH_ERR_ARG, "fh_ga_slab_stage: stage %d (1, 3 or 4 with direct exchange buffers)", stage);
    }
    op->row_beg = op->row_cnt = 0;
    return rc;
}

// Pieces of the CG iteration for a solve distributed over ranks (general/solver.py:113-136): every scalar
// is a local partial sum -> caller's all-reduce on the device -> fh_cgd_scal.  vecs = [r | p | Ap].
extern "C" int fh_cgd_init(fh_ga* op, const double* B, double* vecs) {
    FH_REQUIRE(op && B && vecs, "fh_cgd_init: null argument");
    const int64_t n = (int64_t)op->D * op->nloc;
    const unsigned g = ga_grid(n);
    k_cg_init<<<g, GA_NT, 0, fh_stream()>>>(n, B, vecs + 2 * n, vecs, vecs + n, op->part);
    FH_LAUNCH_CHECK();
    op->last_npart = (int)g;
    return FH_OK;
}
extern "C" int fh_cgd_update(fh_ga* op, double* x, double* vecs) {
    FH_REQUIRE(op && x && vecs, "fh_cgd_update: null argument");
    const int64_t n = (int64_t)op->D * op->nloc;
    double* r = vecs;
    const double* p = vecs + n;
    const double* Ap = vecs + 2 * n;
    const unsigned g = ga_grid(n / 2 + 1);
    if (n % 2 == 0 && (((uintptr_t)x | (uintptr_t)vecs) & 15) == 0)
        k_cg_update<<<g, GA_NT, 0, fh_stream()>>>(n / 2, (double2*)x, (double2*)r, (const double2*)p, (const double2*)Ap,
                                                  op->scal, op->part);
    else
        k_cg_update1<<<g, GA_NT, 0, fh_stream()>>>(n, x, r, p, Ap, op->scal, op->part);
    FH_LAUNCH_CHECK();
    op->last_npart = (int)g;
    return FH_OK;
}
// Deferred x update for the distributed loop (same scheme as fh_cg_steps): fh_ga_set_xacc(op, x) makes the next S1
// launches with a p update also apply x += alpha p (alpha = the device scalar of the previous iteration);
// fh_cgd_update_r is the update without x; fh_cgd_xflush applies the pending x += alpha p after the last iteration.
extern "C" int fh_ga_set_xacc(fh_ga* op, double* x) {
    FH_REQUIRE(op, "fh_ga_set_xacc: null operator");
    FH_REQUIRE(x == NULL || op->fast_last || op->rt_ok[op->plan->dim - 1],
               "fh_ga_set_xacc: this operator's S1 path does not carry the p update");
    op->xacc = x;
    return FH_OK;
}
extern "C" int fh_ga_can_defer_x(const fh_ga* op) {
    if (!op) return 0;
    const int64_t n = (int64_t)op->D * op->nloc;
    return (n % 2 == 0 && (op->fast_last || op->rt_ok[op->plan->dim - 1]) && env_int("FH_XDEFER", 1)) ? 1 : 0;
}
extern "C" int fh_cgd_update_r(fh_ga* op, double* vecs) {
    FH_REQUIRE(op && vecs, "fh_cgd_update_r: null argument");
    const int64_t n = (int64_t)op->D * op->nloc;
    FH_REQUIRE(n % 2 == 0 && ((uintptr_t)vecs & 15) == 0, "fh_cgd_update_r: even, 16-byte aligned fields required");
    const unsigned g = ga_grid(n / 2 + 1);
    k_cg_update_r<<<g, GA_NT, 0, fh_stream()>>>(n / 2, (double2*)vecs, (const double2*)(vecs + 2 * n), op->scal, op->part);
    FH_LAUNCH_CHECK();
    op->last_npart = (int)g;
    return FH_OK;
}
extern "C" int fh_cgd_xflush(fh_ga* op, double* x, const double* vecs) {
    FH_REQUIRE(op && x && vecs, "fh_cgd_xflush: null argument");
    const int64_t n = (int64_t)op->D * op->nloc;
    FH_REQUIRE(n % 2 == 0 && (((uintptr_t)x | (uintptr_t)vecs) & 15) == 0, "fh_cgd_xflush: even, 16-byte aligned fields required");
    k_cg_xflush<<<ga_grid(n / 2 + 1), GA_NT, 0, fh_stream()>>>(n / 2, (double2*)x, (const double2*)(vecs + n), op->scal);
    FH_LAUNCH_CHECK();
    return FH_OK;
}
// sum_dev[0] = this rank's sum of the partial sums left by the last S5 / init / update launch
extern "C" int fh_cgd_local_sum(fh_ga* op, double* sum_dev) {
    FH_REQUIRE(op && sum_dev, "fh_cgd_local_sum: null argument");
    FH_REQUIRE(op->last_npart > 0, "fh_cgd_local_sum: no partial sums available");
    k_sum_part<<<1, GA_NT, 0, fh_stream()>>>(op->last_npart, op->part, sum_dev);
    FH_LAUNCH_CHECK();
    return FH_OK;
}
// mode 0: rr, norm | mode 1: pAp, alpha | mode 2: beta, rr, norm  from the GLOBAL sum in sum_dev[0];
// norm_host (optional) receives ||r|| (8-byte read-back, synchronises)
extern "C" int fh_cgd_scal(fh_ga* op, const double* sum_dev, int mode, double* norm_host) {
    FH_REQUIRE(op && sum_dev && mode >= 0 && mode <= 2, "fh_cgd_scal: bad argument");
    k_cg_scal<<<1, GA_NT, 0, fh_stream()>>>(1, sum_dev, op->scal, 1.0 / (double)op->plan->nreal, mode);
    FH_LAUNCH_CHECK();
    if (norm_host) return read_norm(op, norm_host);
    return FH_OK;
}


// vecs = [r | p | Ap], each D*prod(N) doubles.
// fh_cg_begin: Ax = Afun(x0); R = B - Ax; P = R; rr = <R,R>      (solver.py:113-120)
extern "C" int fh_cg_begin(fh_ga* op, const double* B, double* x, double* vecs, double* norm_res_host) {
    FH_REQUIRE(op && B && x && vecs && norm_res_host, "fh_cg_begin: null argument");
    const int64_t n = (int64_t)op->D * op->nloc;
    const double inv = 1.0 / (double)op->plan->nreal;
    double* r = vecs;
    double* p = vecs + n;
    double* Ap = vecs + 2 * n;
    const unsigned g = ga_grid(n);
    int rc;
    if ((rc = ga_matvec(op, x, NULL, 0, Ap, 0, NULL))) return rc;
    k_cg_init<<<g, GA_NT, 0, fh_stream()>>>(n, B, Ap, r, p, op->part);
    FH_LAUNCH_CHECK();
    k_cg_scal<<<1, GA_NT, 0, fh_stream()>>>((int)g, op->part, op->scal, inv, 0);
    FH_LAUNCH_CHECK();
    op->kit = 0;
    op->have_beta = 0;
    return read_norm(op, norm_res_host);
}

// up to `nsteps` CG iterations (solver.py:123-136), stopping early when norm_res <= tol.
// Per iteration: 5 pipeline kernels (S1 carries p = r + beta p, S5 carries <p,Ap>), the x/r update
// with <r,r>, two single-CTA scalar kernels, and one 8-byte read-back of the residual norm.
// Deferred x update (default; FH_XDEFER=0 disables): iteration k leaves x += alpha_k p_k pending; S1 of iteration
// k+1 applies it before it overwrites p (same expression, same rounding), and the loop flushes the last one
// before it returns.  Needs an S1 kernel that carries the p update (fast / run-time-length / odd-length paths); an
// odd element count takes the 8-byte forms of the two update kernels.
static int cg_steps_impl(fh_ga* op, double* x, double* vecs, double tol, int64_t nsteps, int64_t* done_host,
                         double* norm_res_host, double* hist_host, int64_t hist_cap) {
    const int64_t n = (int64_t)op->D * op->nloc;
    const double inv = 1.0 / (double)op->plan->nreal;
    double* r = vecs;
    double* p = vecs + n;
    double* Ap = vecs + 2 * n;
    const unsigned g = ga_grid(n / 2 + 1);
    cudaStream_t s = fh_stream();
    int rc;
    double norm_res = *norm_res_host;
    int64_t done = 0;
    static const int want_defer = env_int("FH_XDEFER", 1);
    // (the S1 kernels that apply the pending x += alpha p: two-pass / three-pass, run-time-length and odd-length families)
    const bool defer = want_defer && (op->fast_last || op->rt_ok[op->plan->dim - 1] || op->odd_ax[op->plan->dim - 1]);
    bool pending = false;  // x += alpha p of the last finished iteration not applied yet
    while (norm_res > tol && done < nsteps) {
        int np = 0;
        op->xacc = pending ? x : NULL;
        rc = ga_matvec(op, p, r, op->have_beta, Ap, 1, &np);
        op->xacc = NULL;
        if (rc) return rc;
        pending = false;
        k_cg_scal<<<1, GA_NT, 0, s>>>(np, op->part, op->scal, inv, 1);
        FH_LAUNCH_CHECK();
        if (defer) {
            if (n % 2 == 0)
                k_cg_update_r<<<g, GA_NT, 0, s>>>(n / 2, (double2*)r, (const double2*)Ap, op->scal, op->part);
            else
                k_cg_update_r1<<<g, GA_NT, 0, s>>>(n, r, Ap, op->scal, op->part);
            pending = true;
        } else if (n % 2 == 0) {
            k_cg_update<<<g, GA_NT, 0, s>>>(n / 2, (double2*)x, (double2*)r, (const double2*)p, (const double2*)Ap,
                                            op->scal, op->part);
        } else {
            k_cg_update1<<<g, GA_NT, 0, s>>>(n, x, r, p, Ap, op->scal, op->part);
        }
        FH_LAUNCH_CHECK();
        k_cg_scal<<<1, GA_NT, 0, s>>>((int)g, op->part, op->scal, inv, 2);
        FH_LAUNCH_CHECK();
        op->have_beta = 1;  // P = R + beta P is folded into the next S1
        if ((rc = read_norm(op, &norm_res))) return rc;
        if (hist_host && done < hist_cap) hist_host[done] = norm_res;
        ++done;
        ++op->kit;
    }
    if (pending) {
        if (n % 2 == 0)
            k_cg_xflush<<<g, GA_NT, 0, s>>>(n / 2, (double2*)x, (const double2*)p, op->scal);
        else
            k_cg_xflush1<<<g, GA_NT, 0, s>>>(n, x, p, op->scal);
        FH_LAUNCH_CHECK();
    }
    *done_host = done;
    *norm_res_host = norm_res;
    return FH_OK;
}

extern "C" int fh_cg_steps(fh_ga* op, double* x, double* vecs, double tol, int64_t nsteps, int64_t* done_host,
                           double* norm_res_host, double* hist_host) {
    FH_REQUIRE(op && x && vecs && done_host && norm_res_host, "fh_cg_steps: null argument");
    return cg_steps_impl(op, x, vecs, tol, nsteps, done_host, norm_res_host, hist_host, nsteps);
}

extern "C" int fh_cg(fh_ga* op, const double* B, double* x, double tol, int64_t maxiter, double* vecs,
                     int64_t* kit_host, double* norm_res_host, double* hist_host, int64_t hist_cap) {
    FH_REQUIRE(op && B && x && vecs && kit_host && norm_res_host, "fh_cg: null argument");
    FH_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)vecs & 15) == 0 && ((uintptr_t)B & 15) == 0,
               "fh_cg: buffers must be 16-byte aligned");
    int rc;
    double norm_res = 0.0;
    if ((rc = fh_cg_begin(op, B, x, vecs, &norm_res))) return rc;
    if (hist_host && hist_cap > 0) hist_host[0] = norm_res;
    int64_t kit = 0;
    if ((rc = cg_steps_impl(op, x, vecs, tol, maxiter, &kit, &norm_res, (hist_host && hist_cap > 1) ? hist_host + 1 : NULL,
                            hist_cap - 1)))
        return rc;
    *kit_host = kit;
    *norm_res_host = (kit == 0) ? 0.0 : norm_res;  // solver.py:137-138
    return FH_OK;
}

extern "C" int fh_richardson(fh_ga* op, const double* B, double* x, double alpha, double tol, int64_t maxiter,
                             double* vecs, int64_t* kit_host, double* norm_res_host) {
    FH_REQUIRE(op && B && x && vecs && kit_host && norm_res_host, "fh_richardson: null argument");
    const int64_t n = (int64_t)op->D * op->nloc;
    const double inv = 1.0 / (double)op->plan->nreal;
    const double omega = 1.0 / alpha;
    double* Ax = vecs;
    const unsigned g = ga_grid(n);
    cudaStream_t s = fh_stream();
    int rc;
    double norm_res = 1e15;
    int64_t kit = 0;
    while (norm_res > tol && kit < maxiter) {
        ++kit;
        if ((rc = ga_matvec(op, x, NULL, 0, Ax, 0, NULL))) return rc;
        k_rich_update<<<g, GA_NT, 0, s>>>(n, x, B, Ax, omega, op->part);
        FH_LAUNCH_CHECK();
        k_cg_scal<<<1, GA_NT, 0, s>>>((int)g, op->part, op->scal, inv, 0);
        FH_LAUNCH_CHECK();
        if ((rc = read_norm(op, &norm_res))) return rc;
    }
    *kit_host = kit;
    *norm_res_host = norm_res;
    return FH_OK;
}
