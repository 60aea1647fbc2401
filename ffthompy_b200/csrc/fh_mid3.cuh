// fh_mid3.cuh — stage S3 (C2C along axis 0, closed-form G^(xi), inverse C2C along axis 0, in place) for N0 = 256 with the
// middle of the tile handed to INDEPENDENT WARPS.  Reference semantics as k_mid_green_pipe (ffthompy/projections.py:54-91,
// 185-240 between numpy.fft.fftn / ifftn along axis 0, ffthompy/tensors/fft.py:39-43).
//
// 256 = 16 x 16, decimation in frequency:  F1 couples the rows {j + 16 r} of one component, and after it the sixteen row
// groups {16 q .. 16 q + 15} never meet again until I1: F2 (radix 16 inside a group), G^ (per frequency, all components)
// and I2 touch one group only.  k_mid_green_pipe runs them as CTA-wide phases between barriers (ncu, round 1/2: 12 warps in
// lock step, barrier + mio_throttle stalls, the FP64 and shared-memory pipes busy in turn, never together).  Here warp q
// owns row group q for the whole middle part and synchronises with __syncwarp only, so at any time different warps sit
// in different sub-phases (shared-memory loads, butterflies, the Green arithmetic) and the pipes overlap:
//   F1   tasks (c, j, t): rows j + 16 r -> y_j[q] w_N^(jq) -> rows j + 16 q           warps 0..4D-1... (D*4*16 threads)
//   ---- CTA barrier
//   warp q:  F2 lanes (c, t): rows 16 q + s, s < 16   |  G^ lanes (s, t) x 2  |  I2 lanes (c, t)       __syncwarp between
//   ---- CTA barrier
//   I1   tasks (c, j, t) -> global
// Tile: 4 columns of all D components, double buffered with cp.async (2 x 104.5 KB for D = 6); 16 warps, <= 128 registers.
// Shared memory [c][272 rows (one padding row per 16)][4]; the component stride carries 4 extra elements so that the
// D components a warp touches in F2 / I2 alternate between the two 64-byte bank halves.
#pragma once
#include "fh_fast.cuh"

template <int KIND>
struct Mid3Cfg {
    static constexpr int D = (KIND == FH_GREEN_SCALAR) ? 3 : 6;
    static constexpr int N = 256, T = 4, NPR = N + N / 16;
    static constexpr int CS = NPR * T + 4;  // component stride (complex elements)
    static constexpr int BUF = D * CS;
    static constexpr int NT = 512;
    static constexpr size_t SMEM = (size_t)2 * BUF * sizeof(cplx);
};

struct Mid3Map {
    int64_t rstride, cstride;  // element (c, i0, ii) at data[c*cstride + i0*rstride + ii]
    int spitch;                // columns per spectrum row
    int ntiles, tpr, col0;     // tiles walk tpr 4-column tiles per row starting at column col0
};

template <int KIND>
__global__ void __launch_bounds__(512, 1)
    k_mid3(cplx* __restrict__ data, const cplx* __restrict__ tw, const GreenDesc g, const Mid3Map m, const int nh) {
    using Cfg = Mid3Cfg<KIND>;
    constexpr int D = Cfg::D, N = Cfg::N, T = Cfg::T, CS = Cfg::CS, BUF = Cfg::BUF, NT = Cfg::NT;
    constexpr int NF = D * T * 16;  // F1 / I1 tasks
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf0 = reinterpret_cast<cplx*>(fh_smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31, wq = tid >> 5;  // warp wq owns row group wq in the middle part
    // F1 / I1 role
    const int t1 = tid % T, j1 = (tid / T) % 16, c1 = tid / (T * 16);
    // F2 / I2 role inside the warp
    const int t2 = lane % T, c2 = lane / T;

    auto tile_ii = [&](int tile) -> int64_t {
        const int rowi = tile / m.tpr;
        return (int64_t)rowi * m.spitch + m.col0 + (tile - rowi * m.tpr) * T;
    };
    auto prefetch = [&](int tile, cplx* buf) {
        const int64_t ii = tile_ii(tile);
#pragma unroll 4
        for (int e = tid; e < D * N * T; e += NT) {
            const int tt = e % T, row = (e / T) % N, cc = e / (T * N);
            cp_async16(buf + cc * CS + pidx(row) * T + tt, data + (int64_t)cc * m.cstride + (int64_t)row * m.rstride + ii + tt);
        }
    };

    int it = 0;
    if ((int)blockIdx.x < m.ntiles) prefetch(blockIdx.x, buf0);
    cp_async_commit();
    for (int tile = blockIdx.x; tile < m.ntiles; tile += gridDim.x, ++it) {
        cplx* cur = buf0 + (it & 1) * BUF;
        const int next = tile + gridDim.x;
        if (next < m.ntiles) prefetch(next, buf0 + ((it + 1) & 1) * BUF);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const int rowi = tile / m.tpr;
        const int bcol = m.col0 + (tile - rowi * m.tpr) * T;
        // ---- F1
        if (tid < NF) {
            cplx* sc = cur + c1 * CS + t1;
            cplx v[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = sc[pidx(j1 + r * 16) * T];
            Bfly<16, false>::run(v);
#pragma unroll
            for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], ldtw(tw, q * j1, false));
#pragma unroll
            for (int q = 0; q < 16; ++q) sc[pidx(j1 + q * 16) * T] = v[q];
        }
        __syncthreads();
        // ---- middle part: warp wq on rows 16 wq + s (padded position 17 wq + s)
        {
            cplx* grp = cur + (17 * wq) * T;
            if (c2 < D) {
                cplx* sc = grp + c2 * CS + t2;
                cplx v[16];
#pragma unroll
                for (int s = 0; s < 16; ++s) v[s] = sc[s * T];
                Bfly<16, false>::run(v);
#pragma unroll
                for (int s = 0; s < 16; ++s) sc[s * T] = v[s];
            }
            __syncwarp();
            int k[3];
            k[1] = fh_freq(rowi + g.ioff1, g.N[1]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int pnt = lane + 32 * h;  // (s, tt)
                const int s = pnt / T, tt = pnt - s * T;
                k[0] = fh_freq(wq + 16 * s, N);  // row 16 q + s holds frequency q + 16 s
                const int i2 = bcol + tt;
                k[2] = fh_freq(i2, g.N[2]);
                cplx* sr = grp + s * T + tt;
                cplx e[D];
#pragma unroll
                for (int cc = 0; cc < D; ++cc) e[cc] = sr[cc * CS];
                if (i2 < nh) {
                    green_apply<KIND, 3>(g, k, e);
                } else {
#pragma unroll
                    for (int cc = 0; cc < D; ++cc) e[cc] = make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int cc = 0; cc < D; ++cc) sr[cc * CS] = e[cc];
            }
            __syncwarp();
            if (c2 < D) {
                cplx* sc = grp + c2 * CS + t2;
                cplx v[16];
#pragma unroll
                for (int s = 0; s < 16; ++s) v[s] = sc[s * T];
                Bfly<16, true>::run(v);
#pragma unroll
                for (int s = 0; s < 16; ++s) sc[s * T] = v[s];
            }
        }
        __syncthreads();
        // ---- I1 -> global
        if (tid < NF) {
            const cplx* sc = cur + c1 * CS + t1;
            cplx v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = sc[pidx(j1 + q * 16) * T];
#pragma unroll
            for (int q = 1; q < 16; ++q) v[q] = cmul(v[q], ldtw(tw, q * j1, true));
            Bfly<16, true>::run(v);
            cplx* gp = data + (int64_t)c1 * m.cstride + (int64_t)rowi * m.spitch + bcol + t1;
#pragma unroll
            for (int r = 0; r < 16; ++r) gp[(int64_t)(j1 + r * 16) * m.rstride] = v[r];
        }
        __syncthreads();  // the buffer may be refilled by the prefetch of the next iteration
    }
    cp_async_wait<0>();
}
