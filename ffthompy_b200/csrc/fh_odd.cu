// fh_odd.cu — the five pipeline stages of the fused operator for ODD axis lengths N = R1 x R2 (odd radices) as
// compile-time, register-resident two-pass kernels: 255 = 15 x 17, the exact-integration grid (Nbar = 2N - 1) of a
// 128^3 problem (BASELINE config 2; reference: ffthompy/tensors/objects.py:144-166,428-467 for the grid doubling,
// ffthompy/tensors/fft.py:39-43 for the transforms, ffthompy/projections.py:54-91,185-240 for G^).
//
// Same scheme as the power-of-two family of fh_fast.cuh (one radix-R butterfly per thread in registers, one exchange
// through shared memory, Stockham order), with what an odd length changes:
//   * rows of N doubles are only 8-byte aligned and the component stride prod(N) is odd: real-space accesses are
//     8-byte, coalesced over a row;
//   * prod(N[:-1]) is odd, so the last CTA of S1 / S5 holds a partial group of rows (masked);
//   * the butterflies are the conjugate-symmetric direct forms of fh_fft.cuh (bfly_direct: (R-1)^2 real FMAs), which
//     need ~130 registers: one CTA-wide barrier scheme, no padding (odd strides are conflict free on their own);
//   * threads per line = max(R1, R2) = 17, so the lines of a warp do not align with warps: every barrier is CTA-wide.
#include "fh_fast.cuh"
#include "fh_odd.h"
#include <stdlib.h>

template <bool INV>
struct Bfly<15, INV> {
    static __device__ __forceinline__ void run(cplx* v) { bfly_direct<15, INV>(v, nullptr, 0); }
};
template <bool INV>
struct Bfly<17, INV> {
    static __device__ __forceinline__ void run(cplx* v) { bfly_direct<17, INV>(v, nullptr, 0); }
};
template <>
struct Fac2<255> {
    static constexpr int R1 = 15, R2 = 17;
};

static int odd_env(const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
}
bool fh_odd_len(int n) { return n == 255; }
bool fh_odd_on() { return odd_env("FH_ODD", 1) != 0; }  // read per operator (fh_ga_create), so tests can compare families
template <typename K>
static int odd_smem_attr(K kernel, size_t bytes) {
    if (bytes > (size_t)fh_max_smem_optin()) return fh_set_error(FH_ERR_UNSUPPORTED, "odd-length kernel: %zu B shared memory", bytes);
    // (the kernels also hold up to 4.6 KB of static shared memory: opt in well below the 48 KB default limit)
    if (bytes > 40 * 1024) FH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return FH_OK;
}

// ------------------------------------------------------------------ S2 / S4: k_c2c_fast<255, 8, .> (fh_fast.cuh)
template <int N, int T>
static int odd_c2c_NT(const cplx* tw, cplx* data, int64_t outer, int64_t inner, bool inv) {
    const size_t smem = (size_t)N * T * sizeof(cplx);
    const int ntile = (int)(inner / T);
    const unsigned nblk = (unsigned)(outer * ntile);
    const int nt = T * FastCfg<N>::TPL;
    int rc;
    if (inv) {
        if ((rc = odd_smem_attr(k_c2c_fast<N, T, true>, smem))) return rc;
        k_c2c_fast<N, T, true><<<nblk, nt, smem, fh_stream()>>>(data, data, tw, inner, ntile, 0, 1.0);
    } else {
        if ((rc = odd_smem_attr(k_c2c_fast<N, T, false>, smem))) return rc;
        k_c2c_fast<N, T, false><<<nblk, nt, smem, fh_stream()>>>(data, data, tw, inner, ntile, 0, 1.0);
    }
    FH_LAUNCH_CHECK();
    return FH_OK;
}
int fh_odd_c2c(int N, const cplx* tw, cplx* data, int64_t outer, int64_t inner, bool inv) {
    if (inner % 8) return fh_set_error(FH_ERR_UNSUPPORTED, "odd-length strided pass: inner %lld", (long long)inner);
    if (N == 255) return odd_c2c_NT<255, 8>(tw, data, outer, inner, inv);
    return fh_set_error(FH_ERR_UNSUPPORTED, "no odd-length strided kernel for N=%d", N);
}

// ------------------------------------------------------------------ S3: axis 0 forward, G^, axis 0 inverse
// data [D][N][inner]; persistent CTAs walk over tiles of T columns of all D components; the next tile is fetched with
// cp.async into the second buffer while the current one is transformed.  Decimation in frequency forward, the mirrored
// network backward, so every stage is in place in shared memory (rows in digit-reversed order in between):
//   F1  rows {j + r Rb}:  y_j[q] = w_N^(jq) sum_r x[j + Rb r] w_Ra^(rq)      thread (c, j < Rb, t)
//   F2  rows {q Rb + j}:  X[q + Ra s] = sum_j y_j[q] w_Rb^(js) -> row q Rb + s  thread (c, q < Ra, t)
//   G^  on every row (k0 = q + Ra s),  I2 = inverse of F2,  I1 = inverse of F1 -> global.
template <int N, int T, int KIND, int MINB>
__global__ void __launch_bounds__(((KIND == FH_GREEN_SCALAR) ? 3 : 6) * T * FastCfg<N>::TPL, MINB)
    k_mid_green_odd(cplx* __restrict__ data, const cplx* __restrict__ tw, const GreenDesc g, const int64_t inner,
                    const int nh, const int pitch, const int ntiles) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? 3 : 6;
    constexpr int Ra = FastCfg<N>::R1, Rb = FastCfg<N>::R2, TPL = FastCfg<N>::TPL;
    constexpr int BUF = D * N * T;  // complex elements per buffer
    constexpr int NT = D * T * TPL;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf0 = reinterpret_cast<cplx*>(fh_smem_raw);
    const int t = threadIdx.x % T;
    const int j = (threadIdx.x / T) % TPL;
    const int c = threadIdx.x / (T * TPL);

    auto prefetch = [&](int tile, cplx* buf) {
        const int64_t i0 = (int64_t)tile * T;
#pragma unroll 4
        for (int e = threadIdx.x; e < D * N * T; e += NT) {
            const int tt = e % T, row = (e / T) % N, cc = e / (T * N);
            cp_async16(buf + (cc * N + row) * T + tt, data + ((int64_t)cc * N + row) * inner + i0 + tt);
        }
    };

    int it = 0;
    if ((int)blockIdx.x < ntiles) prefetch(blockIdx.x, buf0);
    cp_async_commit();
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        cplx* cur = buf0 + (it & 1) * BUF;
        const int next = tile + gridDim.x;
        if (next < ntiles) prefetch(next, buf0 + ((it + 1) & 1) * BUF);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        cplx* sc = cur + c * N * T + t;
        const int64_t i0 = (int64_t)tile * T;
        // F1
        if (j < Rb) {
            cplx v[Ra];
#pragma unroll
            for (int r = 0; r < Ra; ++r) v[r] = sc[(j + r * Rb) * T];
            Bfly<Ra, false>::run(v);
#pragma unroll
            for (int q = 1; q < Ra; ++q) v[q] = cmul(v[q], ldtw(tw, q * j, false));
#pragma unroll
            for (int q = 0; q < Ra; ++q) sc[(j + q * Rb) * T] = v[q];
        }
        __syncthreads();
        // F2
        if (j < Ra) {
            cplx v[Rb];
#pragma unroll
            for (int s = 0; s < Rb; ++s) v[s] = sc[(j * Rb + s) * T];
            Bfly<Rb, false>::run(v);
#pragma unroll
            for (int s = 0; s < Rb; ++s) sc[(j * Rb + s) * T] = v[s];
        }
        __syncthreads();
        // G^: row = q*Rb + s holds frequency index q + Ra*s
        for (int idx = threadIdx.x; idx < N * T; idx += NT) {
            const int row = idx / T, tt = idx - row * T;
            const int q = row / Rb, s = row - q * Rb;
            int k[3];
            k[0] = fh_freq(q + Ra * s, N);
            const int64_t ii = i0 + tt;
            const int i1 = (int)(ii / pitch), i2 = (int)(ii - (int64_t)i1 * pitch);
            k[1] = fh_freq(i1 + g.ioff1, g.N[1]);
            k[2] = fh_freq(i2, g.N[2]);
            cplx* sr = cur + row * T + tt;
            cplx e[D];
#pragma unroll
            for (int cc = 0; cc < D; ++cc) e[cc] = sr[cc * N * T];
            if (i2 < nh) {
                green_apply<KIND, 3>(g, k, e);
            } else {
#pragma unroll
                for (int cc = 0; cc < D; ++cc) e[cc] = make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int cc = 0; cc < D; ++cc) sr[cc * N * T] = e[cc];
        }
        __syncthreads();
        // I2
        if (j < Ra) {
            cplx v[Rb];
#pragma unroll
            for (int s = 0; s < Rb; ++s) v[s] = sc[(j * Rb + s) * T];
            Bfly<Rb, true>::run(v);
#pragma unroll
            for (int s = 0; s < Rb; ++s) sc[(j * Rb + s) * T] = v[s];
        }
        __syncthreads();
        // I1 -> global
        if (j < Rb) {
            cplx v[Ra];
#pragma unroll
            for (int q = 0; q < Ra; ++q) v[q] = sc[(j + q * Rb) * T];
#pragma unroll
            for (int q = 1; q < Ra; ++q) v[q] = cmul(v[q], ldtw(tw, q * j, true));
            Bfly<Ra, true>::run(v);
            cplx* gp = data + (int64_t)c * N * inner + i0 + t;
#pragma unroll
            for (int r = 0; r < Ra; ++r) gp[(int64_t)(j + r * Rb) * inner] = v[r];
        }
        __syncthreads();  // the buffer may be refilled by the prefetch of the next iteration
    }
    cp_async_wait<0>();
}

template <int N, int T, int KIND, int MINB>
static int odd_mid_launch(fh_ga* op) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? 3 : 6;
    const fh_plan* p = op->plan;
    const int64_t inner = (int64_t)op->n1l * op->pitch;
    const size_t smem = (size_t)2 * D * N * T * sizeof(cplx);
    const int ntiles = (int)(inner / T);
    int rc;
    if ((rc = odd_smem_attr(k_mid_green_odd<N, T, KIND, MINB>, smem))) return rc;
    int grid = fh_num_sms() * MINB;
    if (grid > ntiles) grid = ntiles;
    k_mid_green_odd<N, T, KIND, MINB><<<grid, D * T * FastCfg<N>::TPL, smem, fh_stream()>>>(op->specT, p->ax[0].tw, op->g, inner,
                                                                                         p->nh, op->pitch, ntiles);
    FH_LAUNCH_CHECK();
    return FH_OK;
}
int fh_odd_mid(fh_ga* op) {
    const fh_plan* p = op->plan;
    if (p->dim != 3 || p->N[0] != 255 || ((int64_t)op->n1l * op->pitch) % 8)
        return fh_set_error(FH_ERR_UNSUPPORTED, "odd-length axis-0 pass: N0=%d", p->N[0]);
    static const int wantT = odd_env("FH_ODD_T", 4);
    if (op->g.kind == FH_GREEN_SCALAR) {
        // D = 3: two CTAs of 4-column tiles per SM (independent phases), or one of 8-column tiles (128-byte segments)
        if (wantT == 8) return odd_mid_launch<255, 8, FH_GREEN_SCALAR, 1>(op);
        return odd_mid_launch<255, 4, FH_GREEN_SCALAR, 2>(op);
    }
    return odd_mid_launch<255, 4, FH_GREEN_ELASTIC, 1>(op);
}

// ------------------------------------------------------------------ S1: p (CG form), sigma = A p, R2C along the last axis
// Real fields [D][nrows][N]; one CTA transforms TRW consecutive rows of all D components (NL = D*TRW real lines, two
// per complex transform; rows >= nrows of the last CTA are zero lines whose results are not stored).  Shared memory:
// NP = NL/2 complex lines as SoA (re plane, im plane), LP = N + 1 doubles apart.  blockDim = NP * TPL.
// Coefficient layouts as in k_fwd_last_fast: 0 full, 1 upper triangle of the full array, 2 phase table, 3 two phases
// in the constant bank.
template <int N, int D, int TRW, int ALAY, int V>
__global__ void __launch_bounds__((D * TRW / 2) * FastCfg<N>::TPL, 2)
    k_fwd_last_odd(const double* __restrict__ A, const unsigned char* __restrict__ phase, const double* __restrict__ lut,
                   const Lut2C lutc, int nphase, double* __restrict__ p, const double* __restrict__ r,
                   const double* __restrict__ scal, int pupdate, cplx* __restrict__ spec, const cplx* __restrict__ tw,
                   int64_t nrows, int nh, int pitch, double* __restrict__ xacc) {
    constexpr int R1 = FastCfg<N>::R1, R2 = FastCfg<N>::R2, TPL = FastCfg<N>::TPL;
    constexpr int NL = D * TRW, NP = NL / 2, LP = N + 1;
    constexpr int NT = NP * TPL;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    double* zre = reinterpret_cast<double*>(fh_smem_raw);  // [NP][LP]
    double* zim = zre + NP * LP;                           // [NP][LP]
    const int64_t row0 = (int64_t)blockIdx.x * TRW;
    const int64_t n = nrows * N;  // voxels per component
    const double beta = pupdate ? scal[3] : 0.0;
    const double alpha = (pupdate && xacc) ? scal[2] : 0.0;
    __shared__ double slut[(ALAY == 2) ? 16 * D * D : 1];
    if (ALAY == 2) {
        for (int i = threadIdx.x; i < nphase * D * D; i += NT) slut[i] = lut[i];
        __syncthreads();
    }
    // phase 0: V voxels per thread and step, 8-byte accesses coalesced along the rows.  All loads of a step are issued
    // before any is consumed (ncu, round 2: with one or two voxels per step the kernel sat on the load latency, 2.5 TB/s)
    constexpr int NA = (ALAY == 0) ? D * D : (ALAY == 1 ? D * (D + 1) / 2 : 0);  // coefficient loads per voxel
    for (int v0 = threadIdx.x; v0 < TRW * N; v0 += V * NT) {
        double pv[V][D], rv[V][D], xv[V][D], av[V][NA > 0 ? NA : 1];
        int ph[V], rowv[V], i2v[V];
        bool live[V];
        int64_t gvv[V];
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const int v = v0 + u * NT;
            const int row = v / N;
            rowv[u] = row;
            i2v[u] = v - row * N;
            live[u] = (v < TRW * N) && (row0 + row < nrows);
            gvv[u] = (row0 + row) * N + i2v[u];
        }
#pragma unroll
        for (int u = 0; u < V; ++u) {
            ph[u] = 0;
            if ((ALAY == 2 || ALAY == 3) && live[u]) ph[u] = phase[gvv[u]];
#pragma unroll
            for (int jj = 0; jj < D; ++jj) {
                pv[u][jj] = live[u] ? p[(size_t)jj * n + gvv[u]] : 0.0;
                if (pupdate) {
                    rv[u][jj] = live[u] ? r[(size_t)jj * n + gvv[u]] : 0.0;
                    if (xacc) xv[u][jj] = live[u] ? xacc[(size_t)jj * n + gvv[u]] : 0.0;
                }
            }
            if (NA > 0 && live[u]) {
                if (ALAY == 0) {
#pragma unroll
                    for (int e = 0; e < D * D; ++e) av[u][e] = A[(size_t)e * n + gvv[u]];
                } else {
                    int e = 0;
#pragma unroll
                    for (int lo = 0; lo < D; ++lo)
#pragma unroll
                        for (int hi = lo; hi < D; ++hi) av[u][e++] = A[((size_t)lo * D + hi) * n + gvv[u]];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < V; ++u) {
            if (v0 + u * NT >= TRW * N) continue;
            if (pupdate && live[u]) {
#pragma unroll
                for (int jj = 0; jj < D; ++jj) {
                    // deferred x += alpha p of the previous iteration (solver.py:127), then p = r + beta p (solver.py:132)
                    if (xacc) xacc[(size_t)jj * n + gvv[u]] = xv[u][jj] + alpha * pv[u][jj];
                    pv[u][jj] = rv[u][jj] + beta * pv[u][jj];
                    p[(size_t)jj * n + gvv[u]] = pv[u][jj];
                }
            }
#pragma unroll
            for (int i = 0; i < D; ++i) {
                double sg = 0.0;
                if (live[u]) {
#pragma unroll
                    for (int jj = 0; jj < D; ++jj) {
                        double a;
                        if (ALAY == 3) {
                            a = ph[u] ? lutc.c[1][i * D + jj] : lutc.c[0][i * D + jj];
                        } else if (ALAY == 2) {
                            a = slut[ph[u] * D * D + i * D + jj];
                        } else if (ALAY == 1) {  // upper triangle, row-major: (lo, hi) -> lo*D - lo(lo-1)/2 + hi - lo
                            const int lo = i < jj ? i : jj, hi = i < jj ? jj : i;
                            a = av[u][lo * D - lo * (lo - 1) / 2 + hi - lo];
                        } else {
                            a = av[u][i * D + jj];
                        }
                        sg += a * pv[u][jj];
                    }
                }
                const int L = i * TRW + rowv[u];
                ((L & 1) ? zim : zre)[(L >> 1) * LP + i2v[u]] = sg;
            }
        }
    }
    __syncthreads();
    const int j = threadIdx.x % TPL, pr = threadIdx.x / TPL;
    double* lre = zre + pr * LP;
    double* lim = zim + pr * LP;
    // pass 1 (rows j + rr*R2 -> rows j*R1 + q): not in place, so read / barrier / write
    {
        cplx v[R1];
        if (j < R2) {
#pragma unroll
            for (int rr = 0; rr < R1; ++rr) v[rr] = make_double2(lre[j + rr * R2], lim[j + rr * R2]);
            Bfly<R1, false>::run(v);
        }
        __syncthreads();
        if (j < R2) {
#pragma unroll
            for (int q = 0; q < R1; ++q) {
                lre[j * R1 + q] = v[q].x;
                lim[j * R1 + q] = v[q].y;
            }
        }
    }
    __syncthreads();
    // pass 2, in place (thread j reads and writes the rows j + rr*R1 only)
    if (j < R1) {
        cplx v[R2];
#pragma unroll
        for (int rr = 0; rr < R2; ++rr) v[rr] = make_double2(lre[j + rr * R1], lim[j + rr * R1]);
#pragma unroll
        for (int rr = 1; rr < R2; ++rr) v[rr] = cmul(v[rr], ldtw(tw, rr * j, false));
        Bfly<R2, false>::run(v);
#pragma unroll
        for (int q = 0; q < R2; ++q) {
            lre[j + q * R1] = v[q].x;
            lim[j + q * R1] = v[q].y;
        }
    }
    __syncthreads();
    // separate the two real lines of every pair and store the half spectra (padding columns zeroed)
    for (int it = threadIdx.x; it < NL * pitch; it += NT) {
        const int L = it / pitch, k = it - L * pitch;
        const int c = L / TRW, row = L - c * TRW;
        if (row0 + row >= nrows) continue;
        cplx X = make_double2(0.0, 0.0);
        if (k < nh) {
            const double* qre = zre + (L >> 1) * LP;
            const double* qim = zim + (L >> 1) * LP;
            const int km = (k == 0) ? 0 : N - k;
            const double ax_ = qre[k], ay_ = qim[k];
            const double bx_ = qre[km], by_ = qim[km];
            X = (L & 1) ? make_double2(0.5 * (ay_ + by_), -0.5 * (ax_ - bx_))
                        : make_double2(0.5 * (ax_ + bx_), 0.5 * (ay_ - by_));
        }
        spec[((size_t)c * nrows + row0 + row) * pitch + k] = X;
    }
}

template <int N, int D, int TRW, int ALAY, int V>
static int odd_fwd_last_AV(fh_ga* op, double* p, const double* r, int pupdate) {
    constexpr int NP = D * TRW / 2;
    const size_t smem = (size_t)2 * NP * (N + 1) * sizeof(double);
    const unsigned nblk = (unsigned)fh_ceil_div(op->nrows, TRW);
    const fh_plan* pl = op->plan;
    int rc;
    if ((rc = odd_smem_attr(k_fwd_last_odd<N, D, TRW, ALAY, V>, smem))) return rc;
    k_fwd_last_odd<N, D, TRW, ALAY, V><<<nblk, NP * FastCfg<N>::TPL, smem, fh_stream()>>>(
        op->A, op->phase, op->lut, op->lutc, op->nphase, p, r, op->scal, pupdate, op->spec, pl->ax[pl->dim - 1].tw,
        op->nrows, pl->nh, op->pitch, op->xacc);
    FH_LAUNCH_CHECK();
    return FH_OK;
}
// voxels per thread and step of phase 0 (loads in flight).  Measured at 255^3 scalar, symmetric coefficients (S1 plain):
// V = 2 0.412 ms, V = 3 0.484, V = 4 0.491 (the register cap of two CTAs per SM makes the wider steps spill); D = 6: one
template <int N, int D, int TRW, int ALAY>
static int odd_fwd_last_A(fh_ga* op, double* p, const double* r, int pupdate) {
    if (D == 3) {
        static const int v = odd_env("FH_ODD_V", 2);
        if (v == 4) return odd_fwd_last_AV<N, D, TRW, ALAY, 4>(op, p, r, pupdate);
        if (v == 3) return odd_fwd_last_AV<N, D, TRW, ALAY, 3>(op, p, r, pupdate);
        return odd_fwd_last_AV<N, D, TRW, ALAY, 2>(op, p, r, pupdate);
    }
    return odd_fwd_last_AV<N, D, TRW, ALAY, 1>(op, p, r, pupdate);
}
template <int N, int D, int TRW>
static int odd_fwd_last_D(fh_ga* op, double* p, const double* r, int pupdate) {
    if (op->a_mode == 2 && op->nphase <= 2) return odd_fwd_last_A<N, D, TRW, 3>(op, p, r, pupdate);
    if (op->a_mode == 2) return odd_fwd_last_A<N, D, TRW, 2>(op, p, r, pupdate);
    if (op->a_mode == 1) return odd_fwd_last_A<N, D, TRW, 1>(op, p, r, pupdate);
    return odd_fwd_last_A<N, D, TRW, 0>(op, p, r, pupdate);
}
int fh_odd_fwd_last(fh_ga* op, double* p, const double* r, int pupdate) {
    const int nl = op->plan->N[op->plan->dim - 1];
    if (nl != 255 || op->row_cnt) return fh_set_error(FH_ERR_UNSUPPORTED, "odd-length last-axis pass: N=%d", nl);
    static const int trw3 = odd_env("FH_ODD_TRW", 8);
    switch (op->D) {
        case 3: return trw3 == 4 ? odd_fwd_last_D<255, 3, 4>(op, p, r, pupdate) : odd_fwd_last_D<255, 3, 8>(op, p, r, pupdate);
        case 6: return odd_fwd_last_D<255, 6, 4>(op, p, r, pupdate);
    }
    return fh_set_error(FH_ERR_UNSUPPORTED, "odd-length last-axis pass: D=%d", op->D);
}

// ------------------------------------------------------------------ S5: C2R along the last axis, y = scale * result, <p, y>
template <int N, int D, int TRW>
__global__ void __launch_bounds__((D * TRW / 2) * FastCfg<N>::TPL, 2)
    k_inv_last_odd(const cplx* __restrict__ spec, double* __restrict__ y, const double* __restrict__ pdot,
                   double* __restrict__ part, const cplx* __restrict__ tw, int64_t nrows, int nh, int pitch,
                   double scale) {
    constexpr int R1 = FastCfg<N>::R1, R2 = FastCfg<N>::R2, TPL = FastCfg<N>::TPL;
    constexpr int NL = D * TRW, NP = NL / 2, LP = N + 1;
    constexpr int NT = NP * TPL;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    __shared__ double red[32];
    double* zre = reinterpret_cast<double*>(fh_smem_raw);
    double* zim = zre + NP * LP;
    const int64_t row0 = (int64_t)blockIdx.x * TRW;
    // phase 0: Z = X_a + i X_b on the full circle (Hermitian completion), natural order; U independent 16-byte loads
    // per thread are issued before any is consumed
    constexpr int U = 8;
    for (int it0 = threadIdx.x; it0 < NP * nh; it0 += U * NT) {
        cplx a[U], b[U];
        int prs[U], ks[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int it = it0 + u * NT;
            prs[u] = -1;
            if (it < NP * nh) {
                const int pr = it / nh, k = it - pr * nh;
                const int La = 2 * pr, Lb = 2 * pr + 1;
                const int ca = La / TRW, ra = La - ca * TRW, cb = Lb / TRW, rb = Lb - cb * TRW;
                a[u] = make_double2(0.0, 0.0);
                b[u] = make_double2(0.0, 0.0);
                if (row0 + ra < nrows) a[u] = spec[((size_t)ca * nrows + row0 + ra) * pitch + k];
                if (row0 + rb < nrows) b[u] = spec[((size_t)cb * nrows + row0 + rb) * pitch + k];
                prs[u] = pr;
                ks[u] = k;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (prs[u] < 0) continue;
            const int k = ks[u];
            cplx av = a[u], bv = b[u];
            if (k == 0) {  // N odd: no Nyquist bin
                av.y = 0.0;
                bv.y = 0.0;
            }
            double* qre = zre + prs[u] * LP;
            double* qim = zim + prs[u] * LP;
            qre[k] = av.x - bv.y;
            qim[k] = av.y + bv.x;
            if (k > 0) {
                qre[N - k] = av.x + bv.y;
                qim[N - k] = -av.y + bv.x;
            }
        }
    }
    __syncthreads();
    const int j = threadIdx.x % TPL, pr = threadIdx.x / TPL;
    double* lre = zre + pr * LP;
    double* lim = zim + pr * LP;
    // inverse of pass 2, in place
    if (j < R1) {
        cplx v[R2];
#pragma unroll
        for (int q = 0; q < R2; ++q) v[q] = make_double2(lre[j + q * R1], lim[j + q * R1]);
        Bfly<R2, true>::run(v);
#pragma unroll
        for (int rr = 1; rr < R2; ++rr) v[rr] = cmul(v[rr], ldtw(tw, rr * j, true));
#pragma unroll
        for (int rr = 0; rr < R2; ++rr) {
            lre[j + rr * R1] = v[rr].x;
            lim[j + rr * R1] = v[rr].y;
        }
    }
    __syncthreads();
    // inverse of pass 1: registers hold z[j + rr*R2]; re -> line 2*pr, im -> line 2*pr+1
    double acc = 0.0;
    if (j < R2) {
        const int La = 2 * pr, Lb = 2 * pr + 1;
        const int ca = La / TRW, ra = La - ca * TRW, cb = Lb / TRW, rb = Lb - cb * TRW;
        const bool la = row0 + ra < nrows, lb = row0 + rb < nrows;
        const size_t oa = ((size_t)ca * nrows + row0 + ra) * N, ob = ((size_t)cb * nrows + row0 + rb) * N;
        // the operand of <p, y> is fetched before the butterfly: its latency hides behind the arithmetic (ncu, round 2:
        // loaded inside the store loop, every one of the 2*R1 loads was waited for in turn)
        double pa[R1], pb[R1];
        if (pdot) {
#pragma unroll
            for (int rr = 0; rr < R1; ++rr) {
                pa[rr] = la ? pdot[oa + j + rr * R2] : 0.0;
                pb[rr] = lb ? pdot[ob + j + rr * R2] : 0.0;
            }
        }
        cplx v[R1];
#pragma unroll
        for (int q = 0; q < R1; ++q) v[q] = make_double2(lre[j * R1 + q], lim[j * R1 + q]);
        Bfly<R1, true>::run(v);
        double acc2 = 0.0;
#pragma unroll
        for (int rr = 0; rr < R1; ++rr) {
            const int i2 = j + rr * R2;
            const double ya = v[rr].x * scale, yb = v[rr].y * scale;
            if (la) y[oa + i2] = ya;
            if (lb) y[ob + i2] = yb;
            if (pdot) {
                acc += pa[rr] * ya;
                acc2 += pb[rr] * yb;
            }
        }
        acc += acc2;
    }
    if (pdot) {
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) part[blockIdx.x] = acc;
    }
}

template <int N, int D, int TRW>
static int odd_inv_last_D(fh_ga* op, double* y, const double* pdot, int* npart) {
    constexpr int NP = D * TRW / 2;
    const size_t smem = (size_t)2 * NP * (N + 1) * sizeof(double);
    const unsigned nblk = (unsigned)fh_ceil_div(op->nrows, TRW);
    const fh_plan* pl = op->plan;
    int rc;
    if ((rc = odd_smem_attr(k_inv_last_odd<N, D, TRW>, smem))) return rc;
    if (pdot && nblk > GA_MAXPART) return fh_set_error(FH_ERR_UNSUPPORTED, "too many partial sums (%u)", nblk);
    k_inv_last_odd<N, D, TRW><<<nblk, NP * FastCfg<N>::TPL, smem, fh_stream()>>>(
        op->spec, y, pdot, op->part, pl->ax[pl->dim - 1].tw, op->nrows, pl->nh, op->pitch, 1.0 / (double)pl->nreal);
    FH_LAUNCH_CHECK();
    if (npart) *npart = (int)nblk;
    return FH_OK;
}
int fh_odd_inv_last(fh_ga* op, double* y, const double* pdot, int* npart) {
    const int nl = op->plan->N[op->plan->dim - 1];
    if (nl != 255 || op->row_cnt) return fh_set_error(FH_ERR_UNSUPPORTED, "odd-length last-axis pass: N=%d", nl);
    static const int trw3 = odd_env("FH_ODD_TRW", 8);
    switch (op->D) {
        case 3: return trw3 == 4 ? odd_inv_last_D<255, 3, 4>(op, y, pdot, npart) : odd_inv_last_D<255, 3, 8>(op, y, pdot, npart);
        case 6: return odd_inv_last_D<255, 6, 4>(op, y, pdot, npart);
    }
    return fh_set_error(FH_ERR_UNSUPPORTED, "odd-length last-axis pass: D=%d", op->D);
}
