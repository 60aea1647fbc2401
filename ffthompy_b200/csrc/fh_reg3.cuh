// fh_reg3.cuh — register-resident THREE-pass kernels of the five-stage pipeline for axis lengths
// N = R1*R2*R3 that do not fit the two-pass kernels of fh_fast.cuh (512 = 8*8*8; BASELINE config 4).
//
// Same design as fh_fast.cuh — every thread owns one radix-R butterfly in registers, global memory is
// touched exactly once per element and direction with R independent 16-byte accesses in flight per
// thread — with one more pass.  All three passes are IN PLACE in shared memory (M = N/R1):
//   pass 1, butterfly j < N/R1      : a_q  = DFT_R1(x[j + r*M]) * w_N^(q j)                  -> pos q*M + j
//   pass 2, butterfly (q, j' < R3)  : c_q2 = DFT_R2(pos q*M + j' + r*R3) * w_N^(R1 j' q2)     -> pos q*M + q2*R3 + j'
//   pass 3, butterfly (q, q2)       : X[q + R1*q2 + R1*R2*k] = DFT_R3(pos q*M + q2*R3 + j')[k] -> pos q*M + q2*R3 + k
// so position q*M + q2*R3 + k holds frequency q + R1*q2 + R1*R2*k (reg3_pos_of_freq / reg3_freq_of_pos);
// the inverse runs the mirrored network (conjugate twiddles applied on load), which consumes and
// produces exactly the positions each butterfly already owns.  Rows are padded by one per 8
// (pidx8) so the stride-R3 accesses of pass 3 spread over all banks.
#pragma once
#include "fh_fast.cuh"

__device__ __forceinline__ int pidx8(int row) { return row + (row >> 3); }

template <int N>
__host__ __device__ __forceinline__ int reg3_pos_of_freq(int k) {
    constexpr int R1 = Reg3Cfg<N>::R1, R2 = Reg3Cfg<N>::R2, R3 = Reg3Cfg<N>::R3;
    const int q = k % R1, r1 = k / R1;
    return q * (N / R1) + (r1 % R2) * R3 + r1 / R2;
}
template <int N>
__host__ __device__ __forceinline__ int reg3_freq_of_pos(int pos) {
    constexpr int R1 = Reg3Cfg<N>::R1, R2 = Reg3Cfg<N>::R2, R3 = Reg3Cfg<N>::R3;
    constexpr int M = N / R1;
    const int q = pos / M, rem = pos - q * M;
    const int q2 = rem / R3, k = rem - q2 * R3;
    return q + R1 * (q2 + R2 * k);
}

// Padding of one SoA line (8-byte elements, sixteen 8-byte banks per 128-byte wavefront, conflicts counted per half
// warp).  ncu, round 2 (512^3, profiles/r02b_ncu_512_last_axis.md): with one padding element per 8 (pidx8) S1 and S5 ran at
// 96 % / 92 % of the L1/shared-memory throughput with every second wavefront a bank conflict - the S1 stores of phase 0,
// both sides of pass 1 and above all the digit-reversed gathers / scatters of the R2C / C2R separation (8 wavefronts
// instead of 2).  For 512 = 8 x 8 x 8 an exhaustive search over paddings p + a p/8 + b p/16 + c p/32 + d p/64 + e p/128 and
// the two lane orders of the pass-2 butterflies gives  p + p/16 + 2 (p/32) + 2 (p/64)  with q fastest in pass 2: every
// access of the three passes and of phase 0 is conflict free, the separation costs 4 instead of 8 (24.5 wavefronts per
// element and direction instead of 40; 20 is the floor).
// 1024 = 8 x 8 x 16 (same search, 128 threads per line): p + p/16 + p/64 with q fastest in passes 2 AND 3 gives 22
// wavefronts against 64 for p + p/8 (the separation alone costs 16 + 16 there).
template <int N>
struct LinePad {
    static constexpr int NPAD = N + N / 8;
    static constexpr bool SWAP2 = false, SWAP3 = false;
    static __host__ __device__ __forceinline__ int idx(int p) { return p + (p >> 3); }
};
template <>
struct LinePad<512> {
    static constexpr int NPAD = 588;  // idx(511) = 586
    static constexpr bool SWAP2 = true, SWAP3 = false;
    static __host__ __device__ __forceinline__ int idx(int p) { return p + (p >> 4) + 2 * (p >> 5) + 2 * (p >> 6); }
};
template <>
struct LinePad<1024> {
    static constexpr int NPAD = 1104;  // idx(1023) = 1101
    static constexpr bool SWAP2 = true, SWAP3 = true;
    static __host__ __device__ __forceinline__ int idx(int p) { return p + (p >> 4) + (p >> 6); }
};

// ------------------------------------------------------------------ one line in SoA shared memory (last-axis kernels)
// lre / lim: the line's real / imaginary planes, element `pos` at [LinePad<N>::idx(pos)].  u = butterfly index of this
// thread inside the line (0 .. TPL-1).  All threads of the CTA must call (block-wide barriers inside).
template <int N>
__device__ __forceinline__ void reg3_line_fwd(double* __restrict__ lre, double* __restrict__ lim, int u,
                                              const cplx* __restrict__ tw) {
    constexpr int R1 = Reg3Cfg<N>::R1, R2 = Reg3Cfg<N>::R2, R3 = Reg3Cfg<N>::R3;
    constexpr int M = N / R1;
    if (u < Reg3Cfg<N>::B1) {
        cplx v[R1];
#pragma unroll
        for (int r = 0; r < R1; ++r) v[r] = make_double2(lre[LinePad<N>::idx(u + r * M)], lim[LinePad<N>::idx(u + r * M)]);
        Bfly<R1, false>::run(v);
#pragma unroll
        for (int q = 1; q < R1; ++q) v[q] = cmul(v[q], ldtw(tw, q * u, false));
#pragma unroll
        for (int q = 0; q < R1; ++q) {
            lre[LinePad<N>::idx(q * M + u)] = v[q].x;
            lim[LinePad<N>::idx(q * M + u)] = v[q].y;
        }
    }
    __syncthreads();
    if (u < Reg3Cfg<N>::B2) {
        const int q = LinePad<N>::SWAP2 ? u % R1 : u / R3, jp = LinePad<N>::SWAP2 ? u / R1 : u - (u / R3) * R3;
        const int b = q * M + jp;
        cplx v[R2];
#pragma unroll
        for (int r = 0; r < R2; ++r) v[r] = make_double2(lre[LinePad<N>::idx(b + r * R3)], lim[LinePad<N>::idx(b + r * R3)]);
        Bfly<R2, false>::run(v);
#pragma unroll
        for (int q2 = 1; q2 < R2; ++q2) v[q2] = cmul(v[q2], ldtw(tw, R1 * jp * q2, false));
#pragma unroll
        for (int q2 = 0; q2 < R2; ++q2) {
            lre[LinePad<N>::idx(b + q2 * R3)] = v[q2].x;
            lim[LinePad<N>::idx(b + q2 * R3)] = v[q2].y;
        }
    }
    __syncthreads();
    if (u < Reg3Cfg<N>::B3) {
        const int q2 = LinePad<N>::SWAP3 ? u / R1 : u % R2, q = LinePad<N>::SWAP3 ? u % R1 : u / R2;  // lane order: LinePad
        const int b = q * M + q2 * R3;
        cplx v[R3];
#pragma unroll
        for (int jp = 0; jp < R3; ++jp) v[jp] = make_double2(lre[LinePad<N>::idx(b + jp)], lim[LinePad<N>::idx(b + jp)]);
        Bfly<R3, false>::run(v);
#pragma unroll
        for (int k = 0; k < R3; ++k) {
            lre[LinePad<N>::idx(b + k)] = v[k].x;
            lim[LinePad<N>::idx(b + k)] = v[k].y;
        }
    }
    __syncthreads();
}

// inverse of passes 3 and 2 in place; the inverse of pass 1 leaves x[u + r*M] (unnormalised) in out[r]
template <int N>
__device__ __forceinline__ void reg3_line_inv(double* __restrict__ lre, double* __restrict__ lim, int u,
                                              const cplx* __restrict__ tw, cplx* out) {
    constexpr int R1 = Reg3Cfg<N>::R1, R2 = Reg3Cfg<N>::R2, R3 = Reg3Cfg<N>::R3;
    constexpr int M = N / R1;
    if (u < Reg3Cfg<N>::B3) {
        const int q2 = LinePad<N>::SWAP3 ? u / R1 : u % R2, q = LinePad<N>::SWAP3 ? u % R1 : u / R2;
        const int b = q * M + q2 * R3;
        cplx v[R3];
#pragma unroll
        for (int k = 0; k < R3; ++k) v[k] = make_double2(lre[LinePad<N>::idx(b + k)], lim[LinePad<N>::idx(b + k)]);
        Bfly<R3, true>::run(v);
#pragma unroll
        for (int jp = 0; jp < R3; ++jp) {
            lre[LinePad<N>::idx(b + jp)] = v[jp].x;
            lim[LinePad<N>::idx(b + jp)] = v[jp].y;
        }
    }
    __syncthreads();
    if (u < Reg3Cfg<N>::B2) {
        const int q = LinePad<N>::SWAP2 ? u % R1 : u / R3, jp = LinePad<N>::SWAP2 ? u / R1 : u - (u / R3) * R3;
        const int b = q * M + jp;
        cplx v[R2];
#pragma unroll
        for (int q2 = 0; q2 < R2; ++q2) v[q2] = make_double2(lre[LinePad<N>::idx(b + q2 * R3)], lim[LinePad<N>::idx(b + q2 * R3)]);
#pragma unroll
        for (int q2 = 1; q2 < R2; ++q2) v[q2] = cmul(v[q2], ldtw(tw, R1 * jp * q2, true));
        Bfly<R2, true>::run(v);
#pragma unroll
        for (int r = 0; r < R2; ++r) {
            lre[LinePad<N>::idx(b + r * R3)] = v[r].x;
            lim[LinePad<N>::idx(b + r * R3)] = v[r].y;
        }
    }
    __syncthreads();
    if (u < Reg3Cfg<N>::B1) {
#pragma unroll
        for (int q = 0; q < R1; ++q) out[q] = make_double2(lre[LinePad<N>::idx(q * M + u)], lim[LinePad<N>::idx(q * M + u)]);
#pragma unroll
        for (int q = 1; q < R1; ++q) out[q] = cmul(out[q], ldtw(tw, q * u, true));
        Bfly<R1, true>::run(out);
    }
}

// ------------------------------------------------------------------ S1: sigma = A p (+ CG updates), R2C last axis
// Same contract as k_fwd_last_fast (fh_fast.cuh): real fields [D][rows][N], TRW rows of all D components
// per CTA, two real lines per complex transform; blockDim = NP * TPL.
// (three CTAs per SM: the CG form lives on the loads in flight; without the cap the 512 padding arithmetic takes the
// kernel to 72 registers = two CTAs and S1 in its CG form loses 5 %)
template <int N, int D, int TRW, int ALAY>
__global__ void __launch_bounds__((D * TRW / 2) * Reg3Cfg<N>::TPL, ((D * TRW / 2) * Reg3Cfg<N>::TPL <= 384) ? 3 : 1)
    k_fwd_last_reg3(const double* __restrict__ A, const unsigned char* __restrict__ phase,
                    const double* __restrict__ lut, const Lut2C lutc, int nphase, double* __restrict__ p,
                    const double* __restrict__ r, const double* __restrict__ scal, int pupdate,
                    cplx* __restrict__ spec, const cplx* __restrict__ tw, int64_t nrows, int nh, int pitch,
                    double* __restrict__ xacc) {
    constexpr int TPL = Reg3Cfg<N>::TPL;
    constexpr int NL = D * TRW, NP = NL / 2, NPAD = LinePad<N>::NPAD;
    constexpr int NT = NP * TPL;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    double* smd = reinterpret_cast<double*>(fh_smem_raw);
    double* zre = smd;              // [NP][NPAD]
    double* zim = smd + NP * NPAD;  // [NP][NPAD]
    const int64_t row0 = (int64_t)blockIdx.x * TRW;
    const int64_t n = nrows * N;
    const double beta = pupdate ? scal[3] : 0.0;
    const double alpha = (pupdate && xacc) ? scal[2] : 0.0;
    __shared__ double slut[(ALAY == 2) ? 16 * D * D : 1];
    if (ALAY == 2) {
        for (int i = threadIdx.x; i < nphase * D * D; i += NT) slut[i] = lut[i];
        __syncthreads();
    }
    s1_sigma_phase<N, D, TRW, ALAY, NT>(A, phase, slut, lutc, p, r, beta, pupdate, row0, n,
                                        [&](int L, int i2, double s0, double s1) {
                                            double* dst = ((L & 1) ? zim : zre) + (L >> 1) * NPAD;
                                            dst[LinePad<N>::idx(i2)] = s0;
                                            dst[LinePad<N>::idx(i2 + 1)] = s1;
                                        },
                                        xacc, alpha);
    __syncthreads();
    const int u = threadIdx.x % TPL, pr = threadIdx.x / TPL;
    reg3_line_fwd<N>(zre + pr * NPAD, zim + pr * NPAD, u, tw);
    // separate the two real lines of every pair and store the half spectra (padding columns zeroed)
    for (int it = threadIdx.x; it < NL * pitch; it += NT) {
        const int L = it / pitch, k = it - L * pitch;
        const int c = L / TRW, row = L - c * TRW;
        cplx X = make_double2(0.0, 0.0);
        if (k < nh) {
            const double* qre = zre + (L >> 1) * NPAD;
            const double* qim = zim + (L >> 1) * NPAD;
            const int pk = LinePad<N>::idx(reg3_pos_of_freq<N>(k));
            const int pm = LinePad<N>::idx(reg3_pos_of_freq<N>((k == 0) ? 0 : N - k));
            const double ax_ = qre[pk], ay_ = qim[pk];
            const double bx_ = qre[pm], by_ = qim[pm];
            X = (L & 1) ? make_double2(0.5 * (ay_ + by_), -0.5 * (ax_ - bx_))
                        : make_double2(0.5 * (ax_ + bx_), 0.5 * (ay_ - by_));
        }
        spec[((size_t)c * nrows + row0 + row) * pitch + k] = X;
    }
}

// ------------------------------------------------------------------ S5: C2R last axis, fused with <p, y>
// Same contract as k_inv_last_fast.
template <int N, int D, int TRW>
__global__ void __launch_bounds__((D * TRW / 2) * Reg3Cfg<N>::TPL)
    k_inv_last_reg3(const cplx* __restrict__ spec, double* __restrict__ y, const double* __restrict__ pdot,
                    double* __restrict__ part, const cplx* __restrict__ tw, int64_t nrows, int nh, int pitch,
                    double scale) {
    constexpr int R1 = Reg3Cfg<N>::R1, TPL = Reg3Cfg<N>::TPL, M = N / R1;
    constexpr int NL = D * TRW, NP = NL / 2, NPAD = LinePad<N>::NPAD;
    constexpr int NT = NP * TPL;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    double* smd = reinterpret_cast<double*>(fh_smem_raw);
    __shared__ double red[32];
    double* zre = smd;
    double* zim = smd + NP * NPAD;
    const int64_t row0 = (int64_t)blockIdx.x * TRW;
    // phase 0: Z = X_a + i X_b on the full circle (Hermitian completion), stored at the pass-3 positions
    constexpr int U = 4;
    for (int it0 = threadIdx.x; it0 < NP * nh; it0 += U * NT) {
        cplx a[U], b[U];
        int prs[U], ks[U];
#pragma unroll
        for (int uu = 0; uu < U; ++uu) {
            const int it = it0 + uu * NT;
            prs[uu] = -1;
            if (it < NP * nh) {
                const int pr = it / nh, k = it - pr * nh;
                const int La = 2 * pr, Lb = 2 * pr + 1;
                const int ca = La / TRW, ra = La - ca * TRW, cb = Lb / TRW, rb = Lb - cb * TRW;
                a[uu] = spec[((size_t)ca * nrows + row0 + ra) * pitch + k];
                b[uu] = spec[((size_t)cb * nrows + row0 + rb) * pitch + k];
                prs[uu] = pr;
                ks[uu] = k;
            }
        }
#pragma unroll
        for (int uu = 0; uu < U; ++uu) {
            if (prs[uu] < 0) continue;
            const int k = ks[uu];
            cplx av = a[uu], bv = b[uu];
            if (k == 0 || 2 * k == N) {
                av.y = 0.0;
                bv.y = 0.0;
            }
            double* qre = zre + prs[uu] * NPAD;
            double* qim = zim + prs[uu] * NPAD;
            const int pk = LinePad<N>::idx(reg3_pos_of_freq<N>(k));
            qre[pk] = av.x - bv.y;
            qim[pk] = av.y + bv.x;
            if (k > 0 && 2 * k != N) {
                const int pm = LinePad<N>::idx(reg3_pos_of_freq<N>(N - k));
                qre[pm] = av.x + bv.y;
                qim[pm] = -av.y + bv.x;
            }
        }
    }
    __syncthreads();
    const int u = threadIdx.x % TPL, pr = threadIdx.x / TPL;
    cplx v[R1];
    reg3_line_inv<N>(zre + pr * NPAD, zim + pr * NPAD, u, tw, v);
    double acc = 0.0;
    if (u < Reg3Cfg<N>::B1) {
        const int La = 2 * pr, Lb = 2 * pr + 1;
        const int ca = La / TRW, ra = La - ca * TRW, cb = Lb / TRW, rb = Lb - cb * TRW;
        const size_t oa = ((size_t)ca * nrows + row0 + ra) * N, ob = ((size_t)cb * nrows + row0 + rb) * N;
#pragma unroll
        for (int rr = 0; rr < R1; ++rr) {
            const int i2 = u + rr * M;
            const double ya = v[rr].x * scale, yb = v[rr].y * scale;
            y[oa + i2] = ya;
            y[ob + i2] = yb;
            if (pdot) acc += pdot[oa + i2] * ya + pdot[ob + i2] * yb;
        }
    }
    if (pdot) {
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) part[blockIdx.x] = acc;
    }
}

// ------------------------------------------------------------------ S3: C2C axis 0, G^, inverse C2C axis 0
// data [D][N][inner]; one CTA owns T consecutive inner positions of all D components; the tile sits in
// shared memory as [D][N + N/8][T] complex.  NT threads walk the D*T*(N/R) butterflies of each pass
// (pass 1 straight from global memory, inverse pass 1 straight back).
// Addressing: element (c, row, ii) at  c * cstride + (rowoff ? rowoff[row] : row * inner) + ii  — the natural
// [D][N][inner] array (rowoff = NULL, cstride = N * inner) or a slab-exchange buffer (fh_ga_slab_direct).
template <int N, int T, int KIND, int DIM, int NT>
__global__ void __launch_bounds__(NT, 1)
    k_mid_green_reg3(cplx* __restrict__ data, const cplx* __restrict__ tw, GreenDesc g, int64_t inner, int nh,
                     int pitch, const int64_t* __restrict__ rowoff, int64_t cstride,
                     cplx* __restrict__ dout = nullptr, const int64_t* __restrict__ rowoff_out = nullptr,
                     int64_t cstride_out = 0, int kcol0 = 0) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    constexpr int R1 = Reg3Cfg<N>::R1, R2 = Reg3Cfg<N>::R2, R3 = Reg3Cfg<N>::R3;
    constexpr int B1 = Reg3Cfg<N>::B1, B2 = Reg3Cfg<N>::B2, B3 = Reg3Cfg<N>::B3;
    constexpr int M = N / R1, NPR = N + N / 8;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf = reinterpret_cast<cplx*>(fh_smem_raw);  // [D][NPR][T]
    const int64_t i0 = (int64_t)blockIdx.x * T;
    // F1: global -> registers -> smem
    for (int w = threadIdx.x; w < D * B1 * T; w += NT) {
        const int t = w % T, u = (w / T) % B1, c = w / (T * B1);
        const cplx* gp = data + (int64_t)c * cstride + i0 + t;
        cplx v[R1];
#pragma unroll
        for (int r = 0; r < R1; ++r) v[r] = gp[rowoff ? rowoff[u + r * M] : (int64_t)(u + r * M) * inner];
        Bfly<R1, false>::run(v);
#pragma unroll
        for (int q = 1; q < R1; ++q) v[q] = cmul(v[q], ldtw(tw, q * u, false));
        cplx* sc = buf + (c * NPR) * T + t;
#pragma unroll
        for (int q = 0; q < R1; ++q) sc[pidx8(q * M + u) * T] = v[q];
    }
    __syncthreads();
    // F2
    for (int w = threadIdx.x; w < D * B2 * T; w += NT) {
        const int t = w % T, u = (w / T) % B2, c = w / (T * B2);
        const int q = u / R3, jp = u - q * R3;
        const int b = q * M + jp;
        cplx* sc = buf + (c * NPR) * T + t;
        cplx v[R2];
#pragma unroll
        for (int r = 0; r < R2; ++r) v[r] = sc[pidx8(b + r * R3) * T];
        Bfly<R2, false>::run(v);
#pragma unroll
        for (int q2 = 1; q2 < R2; ++q2) v[q2] = cmul(v[q2], ldtw(tw, R1 * jp * q2, false));
#pragma unroll
        for (int q2 = 0; q2 < R2; ++q2) sc[pidx8(b + q2 * R3) * T] = v[q2];
    }
    __syncthreads();
    // F3
    for (int w = threadIdx.x; w < D * B3 * T; w += NT) {
        const int t = w % T, u = (w / T) % B3, c = w / (T * B3);
        const int q2 = u % R2, q = u / R2;
        const int b = q * M + q2 * R3;
        cplx* sc = buf + (c * NPR) * T + t;
        cplx v[R3];
#pragma unroll
        for (int jp = 0; jp < R3; ++jp) v[jp] = sc[pidx8(b + jp) * T];
        Bfly<R3, false>::run(v);
#pragma unroll
        for (int k = 0; k < R3; ++k) sc[pidx8(b + k) * T] = v[k];
    }
    __syncthreads();
    // G^ on every frequency of the tile: position pos holds k0 = reg3_freq_of_pos(pos)
    for (int idx = threadIdx.x; idx < N * T; idx += NT) {
        const int pos = idx / T, tt = idx - pos * T;
        int k[3];
        k[0] = fh_freq(reg3_freq_of_pos<N>(pos), N);
        const int64_t ii = i0 + tt;
        bool valid = true;
        if (DIM == 3) {
            // (kcol0: first global column of a k2-block exchange buffer whose rows hold `pitch` columns)
            const int i1 = (int)(ii / pitch), i2 = (int)(ii - (int64_t)i1 * pitch) + kcol0;
            k[1] = fh_freq(i1 + g.ioff1, g.N[1]);
            k[2] = fh_freq(i2, g.N[2]);
            valid = i2 < nh;
        } else {
            k[1] = fh_freq((int)ii, g.N[1]);
            k[2] = 0;
            valid = (int)ii < nh;
        }
        cplx* sr = buf + pidx8(pos) * T + tt;
        cplx e[D];
#pragma unroll
        for (int cc = 0; cc < D; ++cc) e[cc] = sr[cc * NPR * T];
        if (valid) {
            green_apply<KIND, DIM>(g, k, e);
        } else {
#pragma unroll
            for (int cc = 0; cc < D; ++cc) e[cc] = make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int cc = 0; cc < D; ++cc) sr[cc * NPR * T] = e[cc];
    }
    __syncthreads();
    // I3
    for (int w = threadIdx.x; w < D * B3 * T; w += NT) {
        const int t = w % T, u = (w / T) % B3, c = w / (T * B3);
        const int q2 = u % R2, q = u / R2;
        const int b = q * M + q2 * R3;
        cplx* sc = buf + (c * NPR) * T + t;
        cplx v[R3];
#pragma unroll
        for (int k = 0; k < R3; ++k) v[k] = sc[pidx8(b + k) * T];
        Bfly<R3, true>::run(v);
#pragma unroll
        for (int jp = 0; jp < R3; ++jp) sc[pidx8(b + jp) * T] = v[jp];
    }
    __syncthreads();
    // I2
    for (int w = threadIdx.x; w < D * B2 * T; w += NT) {
        const int t = w % T, u = (w / T) % B2, c = w / (T * B2);
        const int q = u / R3, jp = u - q * R3;
        const int b = q * M + jp;
        cplx* sc = buf + (c * NPR) * T + t;
        cplx v[R2];
#pragma unroll
        for (int q2 = 0; q2 < R2; ++q2) v[q2] = sc[pidx8(b + q2 * R3) * T];
#pragma unroll
        for (int q2 = 1; q2 < R2; ++q2) v[q2] = cmul(v[q2], ldtw(tw, R1 * jp * q2, true));
        Bfly<R2, true>::run(v);
#pragma unroll
        for (int r = 0; r < R2; ++r) sc[pidx8(b + r * R3) * T] = v[r];
    }
    __syncthreads();
    // I1: smem -> registers -> global
    for (int w = threadIdx.x; w < D * B1 * T; w += NT) {
        const int t = w % T, u = (w / T) % B1, c = w / (T * B1);
        const cplx* sc = buf + (c * NPR) * T + t;
        cplx v[R1];
#pragma unroll
        for (int q = 0; q < R1; ++q) v[q] = sc[pidx8(q * M + u) * T];
#pragma unroll
        for (int q = 1; q < R1; ++q) v[q] = cmul(v[q], ldtw(tw, q * u, true));
        Bfly<R1, true>::run(v);
        if (dout) {  // push mode of the slab pipeline: results go to the x-slab spectra of the ranks that own the planes
            cplx* gq = dout + (int64_t)c * cstride_out + i0 + t;
#pragma unroll
            for (int r = 0; r < R1; ++r) gq[rowoff_out[u + r * M]] = v[r];
        } else {
            cplx* gp = data + (int64_t)c * cstride + i0 + t;
#pragma unroll
            for (int r = 0; r < R1; ++r) gp[rowoff ? rowoff[u + r * M] : (int64_t)(u + r * M) * inner] = v[r];
        }
    }
}

// ------------------------------------------------------------------ S2 / S4 between the natural x-slab spectrum and a
// slab-exchange buffer (out of place, LineMap addressing of fh_fast.cuh): k_c2c_reg3 with mapped rows
template <int N, int T, bool INV>
__global__ void __launch_bounds__(T* Reg3Cfg<N>::TPL) k_c2c_reg3_map(const cplx* __restrict__ in, cplx* __restrict__ out,
                                                                      const cplx* __restrict__ tw, LineMap mi, LineMap mo,
                                                                      int ntile, int64_t nwork) {
    constexpr int R1 = Reg3Cfg<N>::R1, R2 = Reg3Cfg<N>::R2, R3 = Reg3Cfg<N>::R3;
    constexpr int M = N / R1;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* smc = reinterpret_cast<cplx*>(fh_smem_raw);  // [N][T]
    const int t = threadIdx.x % T, u = threadIdx.x / T;
    // grid == nwork: one tile per CTA; a smaller grid walks the tiles (the push exchange launches one CTA per SM so that
    // the NVLink-bound stores of this kernel share the SMs with the HBM-bound S1 of the next x-plane chunk)
    for (int64_t w = blockIdx.x; w < nwork; w += gridDim.x) {
        const int64_t o = w / ntile;
        const int tile = (int)(w - o * ntile);
        const int64_t bi = linemap_base(mi, o) + (int64_t)tile * T + t;
        const int64_t bo = linemap_base(mo, o) + (int64_t)tile * T + t;
        if (u < Reg3Cfg<N>::B1) {
            cplx v[R1];
#pragma unroll
            for (int r = 0; r < R1; ++r) v[r] = in[bi + linemap_row(mi, u + r * M)];
            Bfly<R1, INV>::run(v);
#pragma unroll
            for (int q = 1; q < R1; ++q) v[q] = cmul(v[q], ldtw(tw, q * u, INV));
#pragma unroll
            for (int q = 0; q < R1; ++q) smc[(q * M + u) * T + t] = v[q];
        }
        __syncthreads();
        if (u < Reg3Cfg<N>::B2) {
            const int q = u / R3, jp = u - q * R3;
            cplx* sp = smc + (q * M + jp) * T + t;
            cplx v[R2];
#pragma unroll
            for (int r = 0; r < R2; ++r) v[r] = sp[r * R3 * T];
            Bfly<R2, INV>::run(v);
#pragma unroll
            for (int q2 = 1; q2 < R2; ++q2) v[q2] = cmul(v[q2], ldtw(tw, R1 * jp * q2, INV));
#pragma unroll
            for (int q2 = 0; q2 < R2; ++q2) sp[q2 * R3 * T] = v[q2];
        }
        __syncthreads();
        if (u < Reg3Cfg<N>::B3) {
            const int q = u % R1, q2 = u / R1;
            const cplx* sp = smc + (q * M + q2 * R3) * T + t;
            cplx v[R3];
#pragma unroll
            for (int jp = 0; jp < R3; ++jp) v[jp] = sp[jp * T];
            Bfly<R3, INV>::run(v);
#pragma unroll
            for (int k = 0; k < R3; ++k) out[bo + linemap_row(mo, u + R1 * R2 * k)] = v[k];
        }
        if (w + gridDim.x < nwork) __syncthreads();  // the tile buffer is refilled by the next round
    }
}
