// fh_mid512.cu — stage S3 (C2C along axis 0, closed-form G^(xi), inverse C2C along axis 0, in place) for N0 = 512 as a
// TWO-stage transform, 512 = 32 x 16, replacing the three-stage (8 x 8 x 8) kernel of fh_reg3.cuh on this axis.
//
// Why: at 512^3 the three-stage kernel is 27 % of the CG step at 1.7 TB/s (profiles/r02_stage_times_512_reg3_nt.log): 14
// shared-memory passes per element and 7 CTA barriers per tile.  Here F1 (radix 32) goes global -> registers -> shared, F2
// (radix 16), G^ and I2 work in shared memory, I1 (radix 32) goes shared -> registers -> global: 8 passes, 5 barriers.
//   n = j + 16 r (j < 16, r < 32),  k = q + 32 s (q < 32, s < 16)
//   F1  y_j[q] = w_N^(j q) sum_r x[j + 16 r] w_32^(r q)        task (c, t, j)     -> slot (q, j)
//   F2  X[q + 32 s] = sum_j y_j[q] w_16^(j s)                  task (c, t, q)     slot (q, j) -> slot (q, s), in place
//   G^  on every (q, s), all D components                       2048 frequencies per tile
//   I2, I1 the mirrored inverse.
// Tile: 4 columns of all D components, [D][32][17][4] complex (one padding group per q: the radix-16 gathers of
// 8 different q hit different banks), 209 KB for D = 6; 384 threads (D * 4 columns * 16), one CTA per SM.  (D = 3: 104 KB
// and 192 threads; two CTAs per SM were measured and are SLOWER, 3.08 against 2.80 ms at 512^3 scalar, so one it stays.)
// Addressing as in k_mid_green_reg3: natural [D][N][inner] or exchange buffers (rowoff / cstride), k2-blocks (kcol0),
// optional scattered output rows (push exchange).  Reference semantics: ffthompy/projections.py:54-91,185-240 applied
// between numpy.fft.fftn / ifftn along axis 0 (ffthompy/tensors/fft.py:39-43).
#include "fh_fast.cuh"
#include "fh_mid2.cuh"   // Bfly<32>
#include "fh_mid512.h"
#include <stdlib.h>

template <int KIND>
__global__ void __launch_bounds__(((KIND == FH_GREEN_SCALAR) ? 3 : 6) * 64, 1)
    k_mid_green_512(cplx* __restrict__ data, const cplx* __restrict__ tw, GreenDesc g, int64_t inner, int nh, int pitch,
                    const int64_t* __restrict__ rowoff, int64_t cstride, cplx* __restrict__ dout,
                    const int64_t* __restrict__ rowoff_out, int64_t cstride_out, int kcol0) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? 3 : 6;
    constexpr int N = 512, RA = 32, RB = 16, T = 4, NT = D * 64;
    constexpr int QS = (RB + 1) * T;  // elements between consecutive q: 17 groups of T
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(fh_smem_raw);  // [D][RA][RB + 1][T]
    const int tid = threadIdx.x;
    const int t = tid & 3, j = (tid >> 2) & 15, c = tid >> 6;
    const int64_t i0 = (int64_t)blockIdx.x * T;
    cplx* sc = sm + c * RA * QS;
    // ---- F1: global -> registers -> shared
    {
        const cplx* gp = data + (int64_t)c * cstride + i0 + t;
        cplx v[RA];
#pragma unroll
        for (int r = 0; r < RA; ++r) v[r] = gp[rowoff ? rowoff[j + RB * r] : (int64_t)(j + RB * r) * inner];
        Bfly<RA, false>::run(v);
#pragma unroll
        for (int q0 = 0; q0 < RA; q0 += 8) {
            cplx w[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) w[a] = __ldg(&tw[(q0 + a) * j]);
#pragma unroll
            for (int a = 0; a < 8; ++a) sc[(q0 + a) * QS + j * T + t] = (q0 + a) ? cmul(v[q0 + a], w[a]) : v[0];
        }
    }
    __syncthreads();
    // ---- F2: radix 16 over j for q = j and q = j + 16 (in place: slot (q, j) -> slot (q, s))
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        cplx* sq = sc + (j + 16 * h) * QS + t;
        cplx v[RB];
#pragma unroll
        for (int jj = 0; jj < RB; ++jj) v[jj] = sq[jj * T];
        Bfly<RB, false>::run(v);
#pragma unroll
        for (int s = 0; s < RB; ++s) sq[s * T] = v[s];
    }
    __syncthreads();
    // ---- G^ on every frequency k0 = q + 32 s of the tile
    for (int idx = tid; idx < N * T; idx += NT) {
        const int tt = idx & 3, s = (idx >> 2) & 15, q = idx >> 6;
        int k[3];
        k[0] = fh_freq(q + RA * s, N);
        const int64_t ii = i0 + tt;
        const int i1 = (int)(ii / pitch), i2 = (int)(ii - (int64_t)i1 * pitch) + kcol0;
        k[1] = fh_freq(i1 + g.ioff1, g.N[1]);
        k[2] = fh_freq(i2, g.N[2]);
        cplx* sr = sm + q * QS + s * T + tt;
        cplx e[D];
#pragma unroll
        for (int cc = 0; cc < D; ++cc) e[cc] = sr[cc * RA * QS];
        if (i2 < nh) {
            green_apply<KIND, 3>(g, k, e);
        } else {
#pragma unroll
            for (int cc = 0; cc < D; ++cc) e[cc] = make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int cc = 0; cc < D; ++cc) sr[cc * RA * QS] = e[cc];
    }
    __syncthreads();
    // ---- I2: inverse radix 16 over s
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        cplx* sq = sc + (j + 16 * h) * QS + t;
        cplx v[RB];
#pragma unroll
        for (int s = 0; s < RB; ++s) v[s] = sq[s * T];
        Bfly<RB, true>::run(v);
#pragma unroll
        for (int jj = 0; jj < RB; ++jj) sq[jj * T] = v[jj];
    }
    __syncthreads();
    // ---- I1: shared -> registers -> global
    {
        cplx v[RA];
#pragma unroll
        for (int q0 = 0; q0 < RA; q0 += 8) {
            cplx w[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                w[a] = __ldg(&tw[(q0 + a) * j]);
                v[q0 + a] = sc[(q0 + a) * QS + j * T + t];
            }
#pragma unroll
            for (int a = 0; a < 8; ++a)
                if (q0 + a) v[q0 + a] = cmul(v[q0 + a], make_double2(w[a].x, -w[a].y));
        }
        Bfly<RA, true>::run(v);
        if (dout) {
            cplx* gq = dout + (int64_t)c * cstride_out + i0 + t;
#pragma unroll
            for (int r = 0; r < RA; ++r) gq[rowoff_out[j + RB * r]] = v[r];
        } else {
            cplx* gp = data + (int64_t)c * cstride + i0 + t;
#pragma unroll
            for (int r = 0; r < RA; ++r) gp[rowoff ? rowoff[j + RB * r] : (int64_t)(j + RB * r) * inner] = v[r];
        }
    }
}

bool fh_mid512_on() {
    static const int on = getenv("FH_MID512") ? atoi(getenv("FH_MID512")) : 1;
    return on != 0;
}

template <int KIND>
static int mid512_launch(cplx* data, const cplx* tw, const GreenDesc& g, int64_t inner, int nh, int pitch, const int64_t* rowoff,
                         int64_t cstride, cplx* dout, const int64_t* rowoff_out, int64_t cstride_out, int kcol0) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? 3 : 6;
    const size_t smem = (size_t)D * 32 * 17 * 4 * sizeof(cplx);
    if (smem > (size_t)fh_max_smem_optin()) return fh_set_error(FH_ERR_UNSUPPORTED, "axis-0 pass (512): %zu B shared memory", smem);
    FH_CUDA(cudaFuncSetAttribute(k_mid_green_512<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_mid_green_512<KIND><<<(unsigned)(inner / 4), D * 64, smem, fh_stream()>>>(
        data, tw, g, inner, nh, pitch, rowoff, rowoff ? cstride : (int64_t)512 * inner, dout, rowoff_out, cstride_out, kcol0);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// 3-D only; inner % 4 == 0.  rowoff == NULL: natural layout (cstride = 512 * inner)
int fh_mid512_green(int kind, cplx* data, const cplx* tw, const GreenDesc& g, int64_t inner, int nh, int pitch,
                    const int64_t* rowoff, int64_t cstride, cplx* dout, const int64_t* rowoff_out, int64_t cstride_out,
                    int kcol0) {
    if (inner % 4) return fh_set_error(FH_ERR_UNSUPPORTED, "axis-0 pass (512): inner %lld", (long long)inner);
    if (kind == FH_GREEN_SCALAR)
        return mid512_launch<FH_GREEN_SCALAR>(data, tw, g, inner, nh, pitch, rowoff, cstride, dout, rowoff_out, cstride_out, kcol0);
    return mid512_launch<FH_GREEN_ELASTIC>(data, tw, g, inner, nh, pitch, rowoff, cstride, dout, rowoff_out, cstride_out, kcol0);
}
