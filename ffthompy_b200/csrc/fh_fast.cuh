// fh_fast.cuh — register-resident two-pass FFT kernels for power-of-two axis lengths
// (N = R1*R2 with radices <= 16: 64, 128, 256), the hot-path versions of the generic
// shared-memory passes in fh_fft.cu.
//
// Design (fp64, HBM-bound): every thread owns one radix-R butterfly (<= 16 complex values in
// registers).  Pass 1 loads its inputs straight from global memory (16 independent 16-byte
// loads in flight per thread — the memory-level parallelism the generic kernels lacked),
// exchanges through shared memory once, and pass 2 stores straight back to global memory in
// natural order (Stockham autosort).  Strided axes move T adjacent lines per CTA so every
// global access is a T*16-byte contiguous segment; the spectrum rows are padded to a multiple
// of 8 complex numbers (128 B) so those segments are sector aligned.
//
//   pass 1 (radix R1, N/R1 = R2 butterflies per line):  a[j*R1 + q] = DFT_R1( x[j + r*R2] )
//   pass 2 (radix R2, N/R2 = R1 butterflies per line):  X[j + q*R1] = DFT_R2( a[j + r*R1] * w_N^(r*j) )
//
// The inverse of the middle (Green) kernel runs the mirrored network (inverse of pass 2, then
// inverse of pass 1), which consumes and produces exactly the rows each thread already owns,
// so the whole forward-G^-inverse sequence is in place in shared memory.
#pragma once
#include "fh_fft.cuh"
#include "fh_green.cuh"
#include "fh_types.cuh"
#include <cooperative_groups.h>

template <int N>
struct Fac2;
template <>
struct Fac2<64> {
    static constexpr int R1 = 8, R2 = 8;
};
template <>
struct Fac2<128> {
    static constexpr int R1 = 8, R2 = 16;
};
template <>
struct Fac2<256> {
    static constexpr int R1 = 16, R2 = 16;
};

static inline bool fh_fast_len(int n) { return n == 64 || n == 128 || n == 256; }

template <int N>
struct FastCfg {
    static constexpr int R1 = Fac2<N>::R1, R2 = Fac2<N>::R2;
    static constexpr int TPL = (R1 > R2) ? R1 : R2;  // threads per line
};

__device__ __forceinline__ cplx ldtw(const cplx* __restrict__ tw, int i, bool inv) {
    cplx w = __ldg(&tw[i]);
    if (inv) w.y = -w.y;
    return w;
}

// ------------------------------------------------------------------ strided complex axis, in place
// array [outer][N][inner]; inner % T == 0.  blockDim = T * TPL.
template <int N, int T, bool INV>
__global__ void __launch_bounds__(T* FastCfg<N>::TPL) k_c2c_fast(const cplx* __restrict__ in, cplx* __restrict__ out,
                                                                  const cplx* __restrict__ tw, int64_t inner,
                                                                  int ntile, int tile0, double scale) {
    constexpr int R1 = FastCfg<N>::R1, R2 = FastCfg<N>::R2;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* smc = reinterpret_cast<cplx*>(fh_smem_raw);  // [N][T]
    const int t = threadIdx.x % T, j = threadIdx.x / T;
    const int64_t o = blockIdx.x / ntile;
    const int tile = blockIdx.x - (int)(o * ntile);
    const int64_t base = o * N * inner + (int64_t)(tile0 + tile) * T + t;  // tile0: first tile of a column chunk
    if (j < R2) {
        cplx v[R1];
#pragma unroll
        for (int r = 0; r < R1; ++r) v[r] = in[base + (int64_t)(j + r * R2) * inner];
        Bfly<R1, INV>::run(v);
#pragma unroll
        for (int r = 0; r < R1; ++r) smc[(j * R1 + r) * T + t] = v[r];
    }
    __syncthreads();
    if (j < R1) {
        cplx v[R2];
#pragma unroll
        for (int r = 0; r < R2; ++r) v[r] = smc[(j + r * R1) * T + t];
#pragma unroll
        for (int r = 1; r < R2; ++r) v[r] = cmul(v[r], ldtw(tw, r * j, INV));
        Bfly<R2, INV>::run(v);
#pragma unroll
        for (int q = 0; q < R2; ++q)
            out[base + (int64_t)(j + q * R1) * inner] = make_double2(v[q].x * scale, v[q].y * scale);
    }
}

// ------------------------------------------------------------------ axis 0: forward, G^, inverse, in place
// data [D][N][inner]; one CTA owns T consecutive inner positions of all D components.
// blockDim = D * T * TPL.  nh = valid entries per spectrum row, pitch = padded row length.
template <int N, int T, int KIND, int DIM>
__global__ void __launch_bounds__(((KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2) * T * FastCfg<N>::TPL,
                                  (T <= 2 ? 3 : 1))
    k_mid_green_fast(cplx* __restrict__ data, const cplx* __restrict__ tw, GreenDesc g, int64_t inner, int nh,
                     int pitch) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    constexpr int R1 = FastCfg<N>::R1, R2 = FastCfg<N>::R2;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* smc = reinterpret_cast<cplx*>(fh_smem_raw);  // [D][N][T]
    const int t = threadIdx.x % T;
    const int c = (threadIdx.x / T) % D;
    const int j = threadIdx.x / (T * D);
    const int64_t i0 = (int64_t)blockIdx.x * T;
    cplx* sc = smc + c * N * T;
    cplx* gp = data + (int64_t)c * N * inner + i0 + t;
    // forward pass 1: global -> registers -> smem
    if (j < R2) {
        cplx v[R1];
#pragma unroll
        for (int r = 0; r < R1; ++r) v[r] = gp[(int64_t)(j + r * R2) * inner];
        Bfly<R1, false>::run(v);
#pragma unroll
        for (int r = 0; r < R1; ++r) sc[(j * R1 + r) * T + t] = v[r];
    }
    __syncthreads();
    // forward pass 2, in place (reads and writes the rows j + r*R1 only)
    if (j < R1) {
        cplx v[R2];
#pragma unroll
        for (int r = 0; r < R2; ++r) v[r] = sc[(j + r * R1) * T + t];
#pragma unroll
        for (int r = 1; r < R2; ++r) v[r] = cmul(v[r], ldtw(tw, r * j, false));
        Bfly<R2, false>::run(v);
#pragma unroll
        for (int q = 0; q < R2; ++q) sc[(j + q * R1) * T + t] = v[q];
    }
    __syncthreads();
    // closed-form Green multiplier on every frequency of the tile
    for (int it = threadIdx.x; it < N * T; it += blockDim.x) {
        const int row = it / T, tt = it - row * T;
        int k[3];
        k[0] = fh_freq(row, N);
        const int64_t ii = i0 + tt;
        bool valid = true;
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
        for (int cc = 0; cc < D; ++cc) e[cc] = smc[(cc * N + row) * T + tt];
        if (valid) {
            green_apply<KIND, DIM>(g, k, e);
        } else {
#pragma unroll
            for (int cc = 0; cc < D; ++cc) e[cc] = make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int cc = 0; cc < D; ++cc) smc[(cc * N + row) * T + tt] = e[cc];
    }
    __syncthreads();
    // inverse of pass 2 (mirrored network), in place
    if (j < R1) {
        cplx v[R2];
#pragma unroll
        for (int q = 0; q < R2; ++q) v[q] = sc[(j + q * R1) * T + t];
        Bfly<R2, true>::run(v);
#pragma unroll
        for (int r = 1; r < R2; ++r) v[r] = cmul(v[r], ldtw(tw, r * j, true));
#pragma unroll
        for (int r = 0; r < R2; ++r) sc[(j + r * R1) * T + t] = v[r];
    }
    __syncthreads();
    // inverse of pass 1: smem -> registers -> global
    if (j < R2) {
        cplx v[R1];
#pragma unroll
        for (int q = 0; q < R1; ++q) v[q] = sc[(j * R1 + q) * T + t];
        Bfly<R1, true>::run(v);
#pragma unroll
        for (int r = 0; r < R1; ++r) gp[(int64_t)(j + r * R2) * inner] = v[r];
    }
}

// ------------------------------------------------------------------ last axis, forward, fused with sigma = A p
// Real fields [D][rows][N]; one CTA transforms TRW consecutive rows of all D components
// (NL = D*TRW real lines, two per complex transform).  Shared memory holds the NP = NL/2 complex
// lines as SoA (re plane, im plane), each line padded by one element per 16 (conflict-free
// radix-16 scatter).  blockDim = NP * TPL.
//   mode bit 0: p = r + beta*p first (the CG direction update, solver.py:132), beta = scal[3]
//   A layout 0: full [D][D][n];  1: symmetric — only the upper triangle of the full array is read;
//   2: piecewise constant — one byte per voxel indexes a table of <= 16 DxD matrices held in shared
//   memory (detected by fh_ga_create, e.g. inclusion-type microstructures);  -1: no multiply
__device__ __forceinline__ int pidx(int row) { return row + (row >> 4); }

// phase 0 of S1: p = r + beta p (optional), sigma = A p on the CTA's TRW x N voxels, two voxels per
// thread and step (16-byte accesses); store(L, i2, s0, s1) receives sigma of real line L = comp*TRW+row
// at positions i2, i2+1.
template <int N, int D, int TRW, int ALAY, int NT, typename Store>
__device__ __forceinline__ void s1_sigma_phase(const double* __restrict__ A, const unsigned char* __restrict__ phase,
                                               const double* slut, const Lut2C& lutc, double* __restrict__ p,
                                               const double* __restrict__ r, double beta, int pupdate, int64_t row0,
                                               int64_t n, Store store, double* __restrict__ xacc = nullptr,
                                               double alpha = 0.0) {
    // phase 0: sigma = A p on TRW x N voxels, two voxels per thread and step (16-byte accesses)
    for (int v = threadIdx.x; v < TRW * (N / 2); v += NT) {
        const int row = v / (N / 2), i2 = 2 * (v - row * (N / 2));
        const int64_t gv = (row0 + row) * N + i2;
        int ph0 = 0, ph1 = 0;
        if (ALAY == 2 || ALAY == 3) {
            const uchar2 ph = *reinterpret_cast<const uchar2*>(phase + gv);
            ph0 = (ALAY == 2) ? ph.x * D * D : ph.x;
            ph1 = (ALAY == 2) ? ph.y * D * D : ph.y;
        }
        double2 pv[D];
#pragma unroll
        for (int jj = 0; jj < D; ++jj) {
            double2 q = *reinterpret_cast<const double2*>(p + (size_t)jj * n + gv);
            if (pupdate) {
                const double2 rr = *reinterpret_cast<const double2*>(r + (size_t)jj * n + gv);
                if (xacc) {  // deferred x += alpha p of the previous iteration (solver.py:127): p is in registers anyway
                    double2* xp = reinterpret_cast<double2*>(xacc + (size_t)jj * n + gv);
                    const double2 xv = *xp;
                    *xp = make_double2(xv.x + alpha * q.x, xv.y + alpha * q.y);
                }
                q = make_double2(rr.x + beta * q.x, rr.y + beta * q.y);
                *reinterpret_cast<double2*>(p + (size_t)jj * n + gv) = q;
            }
            pv[jj] = q;
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double2 s;
            if (ALAY < 0) {
                s = pv[i];
            } else {
                s = make_double2(0.0, 0.0);
#pragma unroll
                for (int jj = 0; jj < D; ++jj) {
                    double2 a;
                    if (ALAY == 3) {
                        const double c0 = lutc.c[0][i * D + jj], c1 = lutc.c[1][i * D + jj];
                        a = make_double2(ph0 ? c1 : c0, ph1 ? c1 : c0);
                    } else if (ALAY == 2) {
                        a = make_double2(slut[ph0 + i * D + jj], slut[ph1 + i * D + jj]);
                    } else if (ALAY == 1) {  // symmetric: (i,j) and (j,i) read the same upper-triangle entry
                        const int lo = i < jj ? i : jj, hi = i < jj ? jj : i;
                        a = *reinterpret_cast<const double2*>(A + ((size_t)lo * D + hi) * n + gv);
                    } else {
                        a = *reinterpret_cast<const double2*>(A + ((size_t)i * D + jj) * n + gv);
                    }
                    s.x += a.x * pv[jj].x;
                    s.y += a.y * pv[jj].y;
                }
            }
            store(i * TRW + row, i2, s.x, s.y);
        }
    }
}

template <int N, int D, int TRW, int ALAY>
__global__ void __launch_bounds__((D * TRW / 2) * FastCfg<N>::TPL)
    k_fwd_last_fast(const double* __restrict__ A, const unsigned char* __restrict__ phase,
                    const double* __restrict__ lut, const Lut2C lutc, int nphase, double* __restrict__ p,
                    const double* __restrict__ r, const double* __restrict__ scal, int pupdate,
                    cplx* __restrict__ spec, const cplx* __restrict__ tw, int64_t nrows, int nh, int pitch,
                    double* __restrict__ xacc) {
    constexpr int R1 = FastCfg<N>::R1, R2 = FastCfg<N>::R2, TPL = FastCfg<N>::TPL;
    constexpr int NL = D * TRW, NP = NL / 2, NPAD = N + N / 16;
    constexpr int NT = NP * TPL;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    double* smd = reinterpret_cast<double*>(fh_smem_raw);
    double* zre = smd;              // [NP][NPAD]
    double* zim = smd + NP * NPAD;  // [NP][NPAD]
    const int64_t row0 = (int64_t)blockIdx.x * TRW;
    const int64_t n = nrows * N;  // voxels per component
    const double beta = pupdate ? scal[3] : 0.0;
    const double alpha = (pupdate && xacc) ? scal[2] : 0.0;
    __shared__ double slut[(ALAY == 2) ? 16 * D * D : 1];
    if (ALAY == 2) {
        for (int i = threadIdx.x; i < nphase * D * D; i += NT) slut[i] = lut[i];
        __syncthreads();
    }
    s1_sigma_phase<N, D, TRW, ALAY, NT>(A, phase, slut, lutc, p, r, beta, pupdate, row0, n,
                                        [&](int L, int i2, double s0, double s1) {
                                            // plain (unpadded) positions, one 16-byte store: the linear phases of this
                                            // kernel (this one, pass-1 loads, the final separation) use the plain
                                            // layout, only the stride-16 scatter / gather between the two radix passes
                                            // the padded one (ncu r01d: 23 % excess shared wavefronts came from linear
                                            // 8-byte accesses crossing the one-per-16 padding)
                                            double* dst = ((L & 1) ? zim : zre) + (L >> 1) * NPAD;
                                            *reinterpret_cast<double2*>(dst + i2) = make_double2(s0, s1);
                                        },
                                        xacc, alpha);
    __syncthreads();
    const int j = threadIdx.x % TPL, pr = threadIdx.x / TPL;
    double* lre = zre + pr * NPAD;
    double* lim = zim + pr * NPAD;
    // pass 1 (rows j + r*R2, plain -> rows j*R1 + q, padded): not in place, so read / barrier / write
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
                lre[pidx(j * R1 + q)] = v[q].x;
                lim[pidx(j * R1 + q)] = v[q].y;
            }
        }
    }
    __syncthreads();
    // pass 2: padded -> plain (the 16 threads of a line sit in one warp: a warp barrier separates the gather from
    // the writes that reuse the line's storage in the other layout)
    {
        cplx v[R2];
        if (j < R1) {
#pragma unroll
            for (int rr = 0; rr < R2; ++rr) v[rr] = make_double2(lre[pidx(j + rr * R1)], lim[pidx(j + rr * R1)]);
#pragma unroll
            for (int rr = 1; rr < R2; ++rr) v[rr] = cmul(v[rr], ldtw(tw, rr * j, false));
            Bfly<R2, false>::run(v);
        }
        __syncwarp();
        if (j < R1) {
#pragma unroll
            for (int q = 0; q < R2; ++q) {
                lre[j + q * R1] = v[q].x;
                lim[j + q * R1] = v[q].y;
            }
        }
    }
    __syncthreads();
    // separate the two real lines of every pair and store the half spectra (padding columns zeroed)
    for (int it = threadIdx.x; it < NL * pitch; it += NT) {
        const int L = it / pitch, k = it - L * pitch;
        const int c = L / TRW, row = L - c * TRW;
        cplx X = make_double2(0.0, 0.0);
        if (k < nh) {
            const double* qre = zre + (L >> 1) * NPAD;
            const double* qim = zim + (L >> 1) * NPAD;
            const int km = (k == 0) ? 0 : N - k;
            const double ax_ = qre[k], ay_ = qim[k];
            const double bx_ = qre[km], by_ = qim[km];
            X = (L & 1) ? make_double2(0.5 * (ay_ + by_), -0.5 * (ax_ - bx_))
                        : make_double2(0.5 * (ax_ + bx_), 0.5 * (ay_ - by_));
        }
        spec[((size_t)c * nrows + row0 + row) * pitch + k] = X;
    }
}

// ------------------------------------------------------------------ last axis, inverse, fused with <p, y>
// y[D][rows][N] = scale * C2R(spec); if pdot != NULL also part[blockIdx.x] = sum over the CTA's
// voxels of pdot*y  (the p.Ap of CG, solver.py:126).
template <int N, int D, int TRW, int MINB = 1>
__global__ void __launch_bounds__((D * TRW / 2) * FastCfg<N>::TPL, MINB)
    k_inv_last_fast(const cplx* __restrict__ spec, double* __restrict__ y, const double* __restrict__ pdot,
                    double* __restrict__ part, const cplx* __restrict__ tw, int64_t nrows, int nh, int pitch,
                    double scale) {
    constexpr int R1 = FastCfg<N>::R1, R2 = FastCfg<N>::R2, TPL = FastCfg<N>::TPL;
    constexpr int NL = D * TRW, NP = NL / 2, NPAD = N + N / 16;
    constexpr int NT = NP * TPL;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    double* smd = reinterpret_cast<double*>(fh_smem_raw);
    __shared__ double red[32];
    double* zre = smd;
    double* zim = smd + NP * NPAD;
    const int64_t row0 = (int64_t)blockIdx.x * TRW;
    // phase 0: Z = X_a + i X_b on the full circle (Hermitian completion), natural order
    // (U independent 16-byte loads per thread are issued before any is consumed)
    constexpr int U = 4;
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
                a[u] = spec[((size_t)ca * nrows + row0 + ra) * pitch + k];
                b[u] = spec[((size_t)cb * nrows + row0 + rb) * pitch + k];
                prs[u] = pr;
                ks[u] = k;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (prs[u] < 0) continue;
            const int k = ks[u];
            cplx av = a[u], bv = b[u];
            if (k == 0 || 2 * k == N) {
                av.y = 0.0;
                bv.y = 0.0;
            }
            double* qre = zre + prs[u] * NPAD;   // plain (unpadded) positions: linear accesses, see k_fwd_last_fast
            double* qim = zim + prs[u] * NPAD;
            qre[k] = av.x - bv.y;
            qim[k] = av.y + bv.x;
            if (k > 0 && 2 * k != N) {
                qre[N - k] = av.x + bv.y;
                qim[N - k] = -av.y + bv.x;
            }
        }
    }
    __syncthreads();
    const int j = threadIdx.x % TPL, pr = threadIdx.x / TPL;
    double* lre = zre + pr * NPAD;
    double* lim = zim + pr * NPAD;
    // inverse of pass 2: plain -> padded (warp barrier between the gather and the writes in the other layout: the
    // TPL threads of a line sit in one warp)
    {
        cplx v[R2];
        if (j < R1) {
#pragma unroll
            for (int q = 0; q < R2; ++q) v[q] = make_double2(lre[j + q * R1], lim[j + q * R1]);
            Bfly<R2, true>::run(v);
#pragma unroll
            for (int rr = 1; rr < R2; ++rr) v[rr] = cmul(v[rr], ldtw(tw, rr * j, true));
        }
        __syncwarp();
        if (j < R1) {
#pragma unroll
            for (int rr = 0; rr < R2; ++rr) {
                lre[pidx(j + rr * R1)] = v[rr].x;
                lim[pidx(j + rr * R1)] = v[rr].y;
            }
        }
    }
    __syncthreads();
    // inverse of pass 1: registers hold z[j + r*R2]; re -> line 2*pr, im -> line 2*pr+1
    double acc = 0.0;
    if (j < R2) {
        cplx v[R1];
#pragma unroll
        for (int q = 0; q < R1; ++q) v[q] = make_double2(lre[pidx(j * R1 + q)], lim[pidx(j * R1 + q)]);
        Bfly<R1, true>::run(v);
        const int La = 2 * pr, Lb = 2 * pr + 1;
        const int ca = La / TRW, ra = La - ca * TRW, cb = Lb / TRW, rb = Lb - cb * TRW;
        const size_t oa = ((size_t)ca * nrows + row0 + ra) * N, ob = ((size_t)cb * nrows + row0 + rb) * N;
#pragma unroll
        for (int rr = 0; rr < R1; ++rr) {
            const int i2 = j + rr * R2;
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

// ------------------------------------------------------------------ axis 0 + G^, software-pipelined
// Persistent variant of k_mid_green_fast: one CTA per SM loops over tiles; the next tile is
// fetched with cp.async (16-byte, L2-only) into the second shared-memory buffer while the current
// one is transformed, so the HBM stream overlaps the FP64 / shared-memory work of the same SM.
// The forward transform is decimation-in-frequency and the inverse its mirror, so every stage is
// in place (no transposing stage, no extra barrier); between them the spectrum sits in shared
// memory in digit-reversed row order, which the Green stage undoes arithmetically:
//   F1: rows {j + r*Rb}: y = DFT_Ra(x) , y[q] *= w_N^(q*j)        (in place)
//   F2: rows {q*Rb + s}: X[q + Ra*s] = DFT_Rb(y_q)[s]             (in place; row q*Rb+s holds k0 = q + Ra*s)
//   G^ on every row, I2 = inverse of F2, I1 = inverse of F1 -> global.
// Rows are padded by one row per 16 so the block-strided F2/I2 accesses are conflict free.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
// named barrier over `nthreads` consecutive threads (a multiple of 32); id 0 is __syncthreads
__device__ __forceinline__ void group_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
template <int NG>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(NG));
}

template <int N, int T, int KIND, int DIM>
__global__ void __launch_bounds__(((KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2) * T * FastCfg<N>::TPL, 1)
    k_mid_green_pipe(cplx* __restrict__ data, const cplx* __restrict__ tw, GreenDesc g, int64_t inner, int nh,
                     int pitch, int ntiles, int tpr, int col0) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    constexpr int Ra = FastCfg<N>::R1, Rb = FastCfg<N>::R2, TPL = FastCfg<N>::TPL;
    constexpr int NPR = N + N / 16;        // padded rows per component
    constexpr int BUF = D * NPR * T;       // complex elements per buffer
    constexpr int NT = D * T * TPL;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf0 = reinterpret_cast<cplx*>(fh_smem_raw);
    const int t = threadIdx.x % T;
    const int j = (threadIdx.x / T) % TPL;
    const int c = threadIdx.x / (T * TPL);

    // tile -> first inner index: tiles walk `tpr` tiles per spectrum row starting at column col0
    // (whole rows: tpr = pitch/T, col0 = 0; a column chunk for L2 blocking: tpr = chunk/T)
    auto tile_i0 = [&](int tile) -> int64_t {
        const int rowi = tile / tpr;
        return (int64_t)rowi * pitch + col0 + (tile - rowi * tpr) * T;
    };
    auto prefetch = [&](int tile, cplx* buf) {
        const int64_t i0 = tile_i0(tile);
#pragma unroll 4
        for (int e = threadIdx.x; e < D * N * T; e += NT) {
            const int tt = e % T, row = (e / T) % N, cc = e / (T * N);
            cp_async16(buf + (cc * NPR + pidx(row)) * T + tt, data + ((int64_t)cc * N + row) * inner + i0 + tt);
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
        cplx* sc = cur + c * NPR * T + t;
        const int64_t i0 = tile_i0(tile);
        // F1
        if (j < Rb) {
            cplx v[Ra];
#pragma unroll
            for (int r = 0; r < Ra; ++r) v[r] = sc[pidx(j + r * Rb) * T];
            Bfly<Ra, false>::run(v);
#pragma unroll
            for (int q = 1; q < Ra; ++q) v[q] = cmul(v[q], ldtw(tw, q * j, false));
#pragma unroll
            for (int q = 0; q < Ra; ++q) sc[pidx(j + q * Rb) * T] = v[q];
        }
        group_sync(c + 1, T * TPL);  // F2 of component c only needs the F1 output of the same T*TPL threads
        // F2
        if (j < Ra) {
            cplx v[Rb];
#pragma unroll
            for (int s = 0; s < Rb; ++s) v[s] = sc[pidx(j * Rb + s) * T];
            Bfly<Rb, false>::run(v);
#pragma unroll
            for (int s = 0; s < Rb; ++s) sc[pidx(j * Rb + s) * T] = v[s];
        }
        __syncthreads();
        // G^: row = q*Rb + s holds frequency index q + Ra*s
        for (int idx = threadIdx.x; idx < N * T; idx += NT) {
            const int row = idx / T, tt = idx - row * T;
            const int q = row / Rb, s = row - q * Rb;
            int k[3];
            k[0] = fh_freq(q + Ra * s, N);
            const int64_t ii = i0 + tt;
            bool valid = true;
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
            cplx* sr = cur + pidx(row) * T + tt;
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
        // I2
        if (j < Ra) {
            cplx v[Rb];
#pragma unroll
            for (int s = 0; s < Rb; ++s) v[s] = sc[pidx(j * Rb + s) * T];
            Bfly<Rb, true>::run(v);
#pragma unroll
            for (int s = 0; s < Rb; ++s) sc[pidx(j * Rb + s) * T] = v[s];
        }
        group_sync(c + 1, T * TPL);
        // I1 -> global
        if (j < Rb) {
            cplx v[Ra];
#pragma unroll
            for (int q = 0; q < Ra; ++q) v[q] = sc[pidx(j + q * Rb) * T];
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

// debugging aid: pure data movement of the axis-0 tiling (global -> smem -> global), to measure what
// the access pattern alone achieves
template <int N, int T, int D>
__global__ void __launch_bounds__(D * T * 16) k_mid_copy_only(cplx* __restrict__ data, int64_t inner) {
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf = reinterpret_cast<cplx*>(fh_smem_raw);
    const int t = threadIdx.x % T;
    const int j = (threadIdx.x / T) % 16;
    const int c = threadIdx.x / (T * 16);
    const int64_t i0 = (int64_t)blockIdx.x * T;
    cplx* gp = data + (int64_t)c * N * inner + i0 + t;
    cplx v[N / 16];
#pragma unroll
    for (int r = 0; r < N / 16; ++r) v[r] = gp[(int64_t)(j + r * 16) * inner];
#pragma unroll
    for (int r = 0; r < N / 16; ++r) buf[(c * N + j + r * 16) * T + t] = v[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < N / 16; ++r) v[r] = buf[(c * N + j * (N / 16) + r) * T + t];
#pragma unroll
    for (int r = 0; r < N / 16; ++r) gp[(int64_t)(j + r * 16) * inner] = v[r];
}

// ------------------------------------------------------------------ axis 0 + G^, two CTAs per SM
// Same in-place DIF / mirrored-inverse scheme as k_mid_green_pipe, but each CTA has only
// (D/CR)*T*TPL threads and walks the D components in CR rounds, loading its inputs straight from
// global memory.  Two such CTAs fit on one SM (registers and shared memory), so the global-load
// phase of one overlaps the FP64 / shared-memory phases of the other.
template <int N, int T, int KIND, int DIM, int CR>
__global__ void __launch_bounds__((((KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2) / CR) * T * FastCfg<N>::TPL,
                                  (T > 4) ? 1 : 2)
    k_mid_green_2r(cplx* __restrict__ data, const cplx* __restrict__ tw, GreenDesc g, int64_t inner, int nh, int pitch,
                   int tpr, int col0) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    constexpr int DC = D / CR;  // components per round
    constexpr int Ra = FastCfg<N>::R1, Rb = FastCfg<N>::R2, TPL = FastCfg<N>::TPL;
    constexpr int NPR = N + N / 16;
    constexpr int NT = DC * T * TPL;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf = reinterpret_cast<cplx*>(fh_smem_raw);  // [D][NPR][T]
    const int t = threadIdx.x % T;
    const int j = (threadIdx.x / T) % TPL;
    const int c0 = threadIdx.x / (T * TPL);
    const int rowi = blockIdx.x / tpr;
    const int64_t i0 = (int64_t)rowi * pitch + col0 + (blockIdx.x - rowi * tpr) * T;
    // F1: global -> registers -> smem (digit-reversed rows come out of F2)
#pragma unroll
    for (int h = 0; h < CR; ++h) {
        const int c = c0 + h * DC;
        if (j < Rb) {
            cplx v[Ra];
            const cplx* gp = data + (int64_t)c * N * inner + i0 + t;
#pragma unroll
            for (int r = 0; r < Ra; ++r) v[r] = gp[(int64_t)(j + r * Rb) * inner];
            Bfly<Ra, false>::run(v);
#pragma unroll
            for (int q = 1; q < Ra; ++q) v[q] = cmul(v[q], ldtw(tw, q * j, false));
            cplx* sc = buf + c * NPR * T + t;
#pragma unroll
            for (int q = 0; q < Ra; ++q) sc[pidx(j + q * Rb) * T] = v[q];
        }
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < CR; ++h) {
        cplx* sc = buf + (c0 + h * DC) * NPR * T + t;
        if (j < Ra) {
            cplx v[Rb];
#pragma unroll
            for (int s = 0; s < Rb; ++s) v[s] = sc[pidx(j * Rb + s) * T];
            Bfly<Rb, false>::run(v);
#pragma unroll
            for (int s = 0; s < Rb; ++s) sc[pidx(j * Rb + s) * T] = v[s];
        }
    }
    __syncthreads();
    {
        // NT is a multiple of T: the tile column (hence k1, k2) is fixed per thread, only k0 varies
        const int tt = threadIdx.x % T;
        const int64_t ii = i0 + tt;
        int k[3];
        bool valid = true;
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
        for (int row = threadIdx.x / T; row < N; row += NT / T) {
            const int q = row / Rb, s = row - q * Rb;
            k[0] = fh_freq(q + Ra * s, N);
            cplx* sr = buf + pidx(row) * T + tt;
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
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < CR; ++h) {
        cplx* sc = buf + (c0 + h * DC) * NPR * T + t;
        if (j < Ra) {
            cplx v[Rb];
#pragma unroll
            for (int s = 0; s < Rb; ++s) v[s] = sc[pidx(j * Rb + s) * T];
            Bfly<Rb, true>::run(v);
#pragma unroll
            for (int s = 0; s < Rb; ++s) sc[pidx(j * Rb + s) * T] = v[s];
        }
    }
    __syncthreads();
#pragma unroll
    for (int h = 0; h < CR; ++h) {
        const int c = c0 + h * DC;
        if (j < Rb) {
            cplx v[Ra];
            const cplx* sc = buf + c * NPR * T + t;
#pragma unroll
            for (int q = 0; q < Ra; ++q) v[q] = sc[pidx(j + q * Rb) * T];
#pragma unroll
            for (int q = 1; q < Ra; ++q) v[q] = cmul(v[q], ldtw(tw, q * j, true));
            Bfly<Ra, true>::run(v);
            cplx* gp = data + (int64_t)c * N * inner + i0 + t;
#pragma unroll
            for (int r = 0; r < Ra; ++r) gp[(int64_t)(j + r * Rb) * inner] = v[r];
        }
    }
}

// ================================================================== generic in-place power-of-two FFT in
// shared memory (2 or 3 decimation-in-frequency stages, radices 4/8/16) — brings every power-of-two
// length that has no register-resident two-pass kernel above (16, 32, 512, 1024, 2048) onto the
// same pipeline.  L lines are interleaved as buf[pidx(row) * L + line] (complex AoS).
//   forward (DIF): stage A rows {j + r*N/Ra}, twiddle w_N^(q j); stage B inside blocks of N/Ra with
//   twiddle w_N^(Ra q j); stage C inside blocks of Rc.  Frequency k = b + Ra*(qb + Rb*s) ends up in
//   row b*(N/Ra) + qb*Rc + s (digit reversed); smem_row_of_freq() gives the map.  The inverse is the
//   mirrored network, so both directions are in place.
template <int N>
struct Fac3;
template <>
struct Fac3<16> {
    static constexpr int Ra = 4, Rb = 4, Rc = 1;
};
template <>
struct Fac3<32> {
    static constexpr int Ra = 4, Rb = 8, Rc = 1;
};
template <>
struct Fac3<512> {
    static constexpr int Ra = 8, Rb = 8, Rc = 8;
};
template <>
struct Fac3<1024> {
    static constexpr int Ra = 8, Rb = 8, Rc = 16;
};
template <>
struct Fac3<2048> {
    static constexpr int Ra = 8, Rb = 16, Rc = 16;
};
// 64/128/256 have dedicated register-resident kernels for the single-GPU pipeline; these three-stage
// splits serve the slab-exchange kernels (k_c2c_map / k_mid_green_map), which exist in this family only
template <>
struct Fac3<64> {
    static constexpr int Ra = 4, Rb = 4, Rc = 4;
};
template <>
struct Fac3<128> {
    static constexpr int Ra = 8, Rb = 4, Rc = 4;
};
template <>
struct Fac3<256> {
    static constexpr int Ra = 8, Rb = 8, Rc = 4;
};
static inline bool fh_gen3_len(int n) { return n == 16 || n == 32 || n == 512 || n == 1024 || n == 2048; }
static inline bool fh_map_len(int n) { return fh_gen3_len(n) || n == 64 || n == 128 || n == 256; }

template <int N>
__host__ __device__ __forceinline__ int smem_freq_of_row(int row) {
    constexpr int Ra = Fac3<N>::Ra, Rb = Fac3<N>::Rb, Rc = Fac3<N>::Rc;
    constexpr int N1 = N / Ra;
    const int b = row / N1, rem = row - b * N1;
    const int qb = rem / Rc, s = rem - qb * Rc;
    return b + Ra * (qb + Rb * s);
}
template <int N>
__host__ __device__ __forceinline__ int smem_row_of_freq(int k) {
    constexpr int Ra = Fac3<N>::Ra, Rb = Fac3<N>::Rb, Rc = Fac3<N>::Rc;
    const int b = k % Ra, r1 = k / Ra;
    const int qb = r1 % Rb, s = r1 / Rb;
    return b * (N / Ra) + qb * Rc + s;
}

// one DIF stage over all lines: blocks of NB rows, radix R, twiddle stride TS (w_N^(TS*q*j)); INV
// runs the mirrored stage (conjugate twiddle before the inverse butterfly)
template <int N, int NB, int R, int TS, bool INV>
__device__ __forceinline__ void smem_stage(cplx* __restrict__ buf, int L, const cplx* __restrict__ tw) {
    constexpr int M = NB / R;               // butterflies per block
    constexpr int NBF = (N / NB) * M;       // butterflies per line
    for (int w = threadIdx.x; w < NBF * L; w += blockDim.x) {
        const int line = w % L, bf = w / L;
        const int blk = bf / M, j = bf - blk * M;
        cplx* base = buf + line;
        const int row0 = blk * NB + j;
        cplx v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = base[pidx(row0 + r * M) * L];
        if (INV) {
            if (M > 1) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmul(v[q], ldtw(tw, TS * q * j, true));
            }
            Bfly<R, true>::run(v);
        } else {
            Bfly<R, false>::run(v);
            if (M > 1) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmul(v[q], ldtw(tw, TS * q * j, false));
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) base[pidx(row0 + r * M) * L] = v[r];
    }
}

template <int N, bool INV>
__device__ __forceinline__ void smem_fft_inplace(cplx* __restrict__ buf, int L, const cplx* __restrict__ tw) {
    constexpr int Ra = Fac3<N>::Ra, Rb = Fac3<N>::Rb, Rc = Fac3<N>::Rc;
    if (!INV) {
        smem_stage<N, N, Ra, 1, false>(buf, L, tw);
        __syncthreads();
        smem_stage<N, N / Ra, Rb, Ra, false>(buf, L, tw);
        __syncthreads();
        if (Rc > 1) {
            smem_stage<N, (Rc > 1 ? Rc : 4), (Rc > 1 ? Rc : 4), 1, false>(buf, L, tw);
            __syncthreads();
        }
    } else {
        if (Rc > 1) {
            smem_stage<N, (Rc > 1 ? Rc : 4), (Rc > 1 ? Rc : 4), 1, true>(buf, L, tw);
            __syncthreads();
        }
        smem_stage<N, N / Ra, Rb, Ra, true>(buf, L, tw);
        __syncthreads();
        smem_stage<N, N, Ra, 1, true>(buf, L, tw);
        __syncthreads();
    }
}

// strided complex axis through the generic routine: [outer][N][inner], T lines per CTA
template <int N, int T, bool INV>
__global__ void __launch_bounds__(256) k_c2c_gen3(const cplx* __restrict__ in, cplx* __restrict__ out,
                                                   const cplx* __restrict__ tw, int64_t inner, int ntile, int tile0,
                                                   double scale) {
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf = reinterpret_cast<cplx*>(fh_smem_raw);  // [N + N/16][T]
    const int64_t o = blockIdx.x / ntile;
    const int tile = blockIdx.x - (int)(o * ntile);
    const int64_t base = o * N * inner + (int64_t)(tile0 + tile) * T;
    // natural-order rows in, digit-reversed after the forward DIF (and the other way round for INV)
    for (int e = threadIdx.x; e < N * T; e += blockDim.x) {
        const int t = e % T, row = e / T;
        const int srow = INV ? smem_row_of_freq<N>(row) : row;
        buf[pidx(srow) * T + t] = in[base + (int64_t)row * inner + t];
    }
    __syncthreads();
    smem_fft_inplace<N, INV>(buf, T, tw);
    for (int e = threadIdx.x; e < N * T; e += blockDim.x) {
        const int t = e % T, row = e / T;
        const int srow = INV ? row : smem_row_of_freq<N>(row);
        const cplx v = buf[pidx(srow) * T + t];
        out[base + (int64_t)row * inner + t] = make_double2(v.x * scale, v.y * scale);
    }
}

// ------------------------------------------------------------------ strided complex axis, three register passes
// N = R1*R2*R3 (512 = 8*8*8, 1024 = 8*8*16, 2048 = 8*16*16): the register-resident scheme of k_c2c_fast with one
// more pass.  Every thread owns one butterfly per pass; pass 1 loads its R1 inputs straight from global memory
// (R1 independent 16-byte loads in flight per thread), two exchanges through shared memory, pass 3 stores straight
// back in natural order.  With M = N/R1:
//   pass 1, thread j < N/R1        : a_q = DFT_R1(x[j + r*M]) * w_N^(q j)                -> smem[q*M + j]
//   pass 2, thread (q, j' < R3)    : c_q2 = DFT_R2(smem[q*M + j' + r*R3]) * w_N^(R1 j' q2) -> smem[q*M + q2*R3 + j']  (in place)
//   pass 3, thread u = q + R1*q2   : X[u + R1*R2*k] = DFT_R3(smem[q*M + q2*R3 + j'])[k]   -> global
// smem is [N][T] complex; T = 8 makes every row one 128-byte line, so all three access patterns are conflict free.
template <int N>
struct Fac3r;
template <>
struct Fac3r<256> {
    static constexpr int R1 = 8, R2 = 8, R3 = 4;
};
template <>
struct Fac3r<512> {
    static constexpr int R1 = 8, R2 = 8, R3 = 8;
};
template <>
struct Fac3r<1024> {
    static constexpr int R1 = 8, R2 = 8, R3 = 16;
};
template <>
struct Fac3r<2048> {
    static constexpr int R1 = 8, R2 = 16, R3 = 16;
};
template <int N>
struct Reg3Cfg {
    static constexpr int R1 = Fac3r<N>::R1, R2 = Fac3r<N>::R2, R3 = Fac3r<N>::R3;
    static constexpr int B1 = N / R1, B2 = N / R2, B3 = N / R3;
    static constexpr int TPL = (B1 > B2 ? (B1 > B3 ? B1 : B3) : (B2 > B3 ? B2 : B3));
};

template <int N, int T, bool INV>
__global__ void __launch_bounds__(T* Reg3Cfg<N>::TPL) k_c2c_reg3(const cplx* __restrict__ in, cplx* __restrict__ out,
                                                                  const cplx* __restrict__ tw, int64_t inner, int ntile,
                                                                  int tile0, double scale) {
    constexpr int R1 = Reg3Cfg<N>::R1, R2 = Reg3Cfg<N>::R2, R3 = Reg3Cfg<N>::R3;
    constexpr int M = N / R1;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* smc = reinterpret_cast<cplx*>(fh_smem_raw);  // [N][T]
    const int t = threadIdx.x % T, u = threadIdx.x / T;
    const int64_t o = blockIdx.x / ntile;
    const int tile = blockIdx.x - (int)(o * ntile);
    const int64_t base = o * N * inner + (int64_t)(tile0 + tile) * T + t;
    if (u < Reg3Cfg<N>::B1) {
        cplx v[R1];
#pragma unroll
        for (int r = 0; r < R1; ++r) v[r] = in[base + (int64_t)(u + r * M) * inner];
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
        for (int k = 0; k < R3; ++k)
            out[base + (int64_t)(u + R1 * R2 * k) * inner] = make_double2(v[k].x * scale, v[k].y * scale);
    }
}

// axis 0 + G^ through the generic routine: data [D][N][inner], tile of T inner positions
template <int N, int T, int KIND, int DIM>
__global__ void __launch_bounds__(384) k_mid_green_gen3(cplx* __restrict__ data, const cplx* __restrict__ tw,
                                                         GreenDesc g, int64_t inner, int nh, int pitch) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    constexpr int L = D * T;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf = reinterpret_cast<cplx*>(fh_smem_raw);  // [N + N/16][D*T]
    const int64_t i0 = (int64_t)blockIdx.x * T;
    for (int e = threadIdx.x; e < D * N * T; e += blockDim.x) {
        const int t = e % T, row = (e / T) % N, c = e / (T * N);
        buf[pidx(row) * L + c * T + t] = data[((int64_t)c * N + row) * inner + i0 + t];
    }
    __syncthreads();
    smem_fft_inplace<N, false>(buf, L, tw);
    for (int idx = threadIdx.x; idx < N * T; idx += blockDim.x) {
        const int row = idx / T, tt = idx - row * T;
        int k[3];
        k[0] = fh_freq(smem_freq_of_row<N>(row), N);
        const int64_t ii = i0 + tt;
        bool valid = true;
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
        cplx* sr = buf + pidx(row) * L + tt;
        cplx e[D];
#pragma unroll
        for (int cc = 0; cc < D; ++cc) e[cc] = sr[cc * T];
        if (valid) {
            green_apply<KIND, DIM>(g, k, e);
        } else {
#pragma unroll
            for (int cc = 0; cc < D; ++cc) e[cc] = make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int cc = 0; cc < D; ++cc) sr[cc * T] = e[cc];
    }
    __syncthreads();
    smem_fft_inplace<N, true>(buf, L, tw);
    for (int e = threadIdx.x; e < D * N * T; e += blockDim.x) {
        const int t = e % T, row = (e / T) % N, c = e / (T * N);
        data[((int64_t)c * N + row) * inner + i0 + t] = buf[pidx(row) * L + c * T + t];
    }
}

// ------------------------------------------------------------------ last axis through the generic routine
// S1: sigma = A p (+ p update), two real lines per complex line, in-place DIF, separation -> spectrum.
// The NP = D*TRW/2 complex lines sit in buf[pidx(row) * NP + pair] (re = even real line, im = odd).
template <int N, int D, int TRW, int ALAY>
__global__ void __launch_bounds__(256)
    k_fwd_last_gen3(const double* __restrict__ A, const unsigned char* __restrict__ phase,
                    const double* __restrict__ lut, const Lut2C lutc, int nphase, double* __restrict__ p,
                    const double* __restrict__ r, const double* __restrict__ scal, int pupdate,
                    cplx* __restrict__ spec, const cplx* __restrict__ tw, int64_t nrows, int nh, int pitch,
                    double* __restrict__ xacc) {
    constexpr int NL = D * TRW, NP = NL / 2, NT = 256;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf = reinterpret_cast<cplx*>(fh_smem_raw);  // [N + N/16][NP]
    const int64_t row0 = (int64_t)blockIdx.x * TRW;
    const int64_t n = nrows * N;
    const double beta = pupdate ? scal[3] : 0.0;
    const double alpha = (pupdate && xacc) ? scal[2] : 0.0;
    __shared__ double slut[(ALAY == 2) ? 16 * D * D : 1];
    if (ALAY == 2) {
        for (int i = threadIdx.x; i < nphase * D * D; i += NT) slut[i] = lut[i];
        __syncthreads();
    }
    double* bd = reinterpret_cast<double*>(buf);
    s1_sigma_phase<N, D, TRW, ALAY, NT>(A, phase, slut, lutc, p, r, beta, pupdate, row0, n,
                                        [&](int L, int i2, double s0, double s1) {
                                            const int pr = L >> 1, part = L & 1;
                                            bd[2 * (pidx(i2) * NP + pr) + part] = s0;
                                            bd[2 * (pidx(i2 + 1) * NP + pr) + part] = s1;
                                        },
                                        xacc, alpha);
    __syncthreads();
    smem_fft_inplace<N, false>(buf, NP, tw);
    for (int it = threadIdx.x; it < NL * pitch; it += NT) {
        const int L = it / pitch, k = it - L * pitch;
        const int c = L / TRW, row = L - c * TRW;
        cplx X = make_double2(0.0, 0.0);
        if (k < nh) {
            const int km = (k == 0) ? 0 : N - k;
            const cplx a = buf[pidx(smem_row_of_freq<N>(k)) * NP + (L >> 1)];
            const cplx b = buf[pidx(smem_row_of_freq<N>(km)) * NP + (L >> 1)];
            X = (L & 1) ? make_double2(0.5 * (a.y + b.y), -0.5 * (a.x - b.x))
                        : make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
        }
        spec[((size_t)c * nrows + row0 + row) * pitch + k] = X;
    }
}

// S5: Hermitian completion of two half spectra into one complex line, mirrored inverse in place,
// re/im -> the two real lines (x scale), optional partial sums of <pdot, y>.
template <int N, int D, int TRW>
__global__ void __launch_bounds__(256)
    k_inv_last_gen3(const cplx* __restrict__ spec, double* __restrict__ y, const double* __restrict__ pdot,
                    double* __restrict__ part, const cplx* __restrict__ tw, int64_t nrows, int nh, int pitch,
                    double scale) {
    constexpr int NL = D * TRW, NP = NL / 2, NT = 256;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    __shared__ double red[32];
    cplx* buf = reinterpret_cast<cplx*>(fh_smem_raw);
    const int64_t row0 = (int64_t)blockIdx.x * TRW;
    for (int it = threadIdx.x; it < NP * nh; it += NT) {
        const int pr = it / nh, k = it - pr * nh;
        const int La = 2 * pr, Lb = 2 * pr + 1;
        const int ca = La / TRW, ra = La - ca * TRW, cb = Lb / TRW, rb = Lb - cb * TRW;
        cplx a = spec[((size_t)ca * nrows + row0 + ra) * pitch + k];
        cplx b = spec[((size_t)cb * nrows + row0 + rb) * pitch + k];
        if (k == 0 || 2 * k == N) {
            a.y = 0.0;
            b.y = 0.0;
        }
        buf[pidx(smem_row_of_freq<N>(k)) * NP + pr] = make_double2(a.x - b.y, a.y + b.x);
        if (k > 0 && 2 * k != N) buf[pidx(smem_row_of_freq<N>(N - k)) * NP + pr] = make_double2(a.x + b.y, -a.y + b.x);
    }
    __syncthreads();
    smem_fft_inplace<N, true>(buf, NP, tw);
    double acc = 0.0;
    for (int it = threadIdx.x; it < NL * N; it += NT) {
        const int L = it / N, i2 = it - L * N;
        const int c = L / TRW, row = L - c * TRW;
        const cplx z = buf[pidx(i2) * NP + (L >> 1)];
        const double v = ((L & 1) ? z.y : z.x) * scale;
        const size_t o = ((size_t)c * nrows + row0 + row) * N + i2;
        y[o] = v;
        if (pdot) acc += pdot[o] * v;
    }
    if (pdot) {
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) part[blockIdx.x] = acc;
    }
}

// ================================================================== run-time-length version of the in-place
// routine: N = R0*R1*R2 with up to three radices from {2,3,4,5,7,8,9,11,13,15,16,17,19}.  This puts the
// odd "doubled" grids of the exact-integration scheme (Nbar = 2N-1: 255 = 15*17, 243 = 9*9*3, 225 = 15*15,
// 125 = 5*5*5, 63 = 9*7, ...) on the same fused five-stage pipeline as the powers of two.
__device__ __forceinline__ int rt_freq_of_row(const RtPlan& P, int row) {
    int k = 0, mul = 1, rem = row;
    for (int s = 0; s < P.ns; ++s) {
        const int sub = P.NB[s] / P.R[s];  // rows per digit step at this stage
        const int dgt = rem / sub;
        rem -= dgt * sub;
        k += dgt * mul;
        mul *= P.R[s];
    }
    return k;
}
__device__ __forceinline__ int rt_row_of_freq(const RtPlan& P, int k) {
    int row = 0, rem = k;
    for (int s = 0; s < P.ns; ++s) {
        const int dgt = rem % P.R[s];
        rem /= P.R[s];
        row += dgt * (P.NB[s] / P.R[s]);
    }
    return row;
}

template <int R, bool INV>
__device__ __forceinline__ void bfly_any(cplx* v) {
    if constexpr (R == 2 || R == 4 || R == 8 || R == 16)
        Bfly<R, INV>::run(v);
    else
        bfly_direct<R, INV>(v, nullptr, 0);
}

template <int R, bool INV>
__device__ __forceinline__ void rt_stage_R(cplx* __restrict__ buf, int L, const cplx* __restrict__ tw, int N, int NB,
                                           int TS) {
    const int M = NB / R;
    const int NBF = (N / NB) * M;
    for (int w = threadIdx.x; w < NBF * L; w += blockDim.x) {
        const int line = w % L, bf = w / L;
        const int blk = bf / M, j = bf - blk * M;
        cplx* base = buf + line;
        const int row0 = blk * NB + j;
        cplx v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = base[pidx(row0 + r * M) * L];
        if (INV) {
            if (M > 1 && j > 0) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmul(v[q], ldtw(tw, TS * q * j, true));
            }
            bfly_any<R, true>(v);
        } else {
            bfly_any<R, false>(v);
            if (M > 1 && j > 0) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[q] = cmul(v[q], ldtw(tw, TS * q * j, false));
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) base[pidx(row0 + r * M) * L] = v[r];
    }
}

template <bool INV>
__device__ __forceinline__ void rt_stage(cplx* buf, int L, const cplx* tw, const RtPlan& P, int s) {
    const int N = P.n, NB = P.NB[s], TS = P.TS[s];
    switch (P.R[s]) {
        case 2: rt_stage_R<2, INV>(buf, L, tw, N, NB, TS); break;
        case 3: rt_stage_R<3, INV>(buf, L, tw, N, NB, TS); break;
        case 4: rt_stage_R<4, INV>(buf, L, tw, N, NB, TS); break;
        case 5: rt_stage_R<5, INV>(buf, L, tw, N, NB, TS); break;
        case 7: rt_stage_R<7, INV>(buf, L, tw, N, NB, TS); break;
        case 8: rt_stage_R<8, INV>(buf, L, tw, N, NB, TS); break;
        case 9: rt_stage_R<9, INV>(buf, L, tw, N, NB, TS); break;
        case 11: rt_stage_R<11, INV>(buf, L, tw, N, NB, TS); break;
        case 13: rt_stage_R<13, INV>(buf, L, tw, N, NB, TS); break;
        case 15: rt_stage_R<15, INV>(buf, L, tw, N, NB, TS); break;
        case 16: rt_stage_R<16, INV>(buf, L, tw, N, NB, TS); break;
        case 17: rt_stage_R<17, INV>(buf, L, tw, N, NB, TS); break;
        case 19: rt_stage_R<19, INV>(buf, L, tw, N, NB, TS); break;
    }
}

template <bool INV>
__device__ __forceinline__ void rt_fft_inplace(cplx* buf, int L, const cplx* tw, const RtPlan& P) {
    if (!INV) {
        for (int s = 0; s < P.ns; ++s) {
            rt_stage<false>(buf, L, tw, P, s);
            __syncthreads();
        }
    } else {
        for (int s = P.ns - 1; s >= 0; --s) {
            rt_stage<true>(buf, L, tw, P, s);
            __syncthreads();
        }
    }
}

// S2 / S4
template <int T, bool INV>
__global__ void __launch_bounds__(256, 2) k_c2c_rt(const cplx* __restrict__ in, cplx* __restrict__ out,
                                                 const cplx* __restrict__ tw, RtPlan P, int64_t inner, int ntile) {
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf = reinterpret_cast<cplx*>(fh_smem_raw);
    const int N = P.n;
    const int64_t o = blockIdx.x / ntile;
    const int tile = blockIdx.x - (int)(o * ntile);
    const int64_t base = o * N * inner + (int64_t)tile * T;
    for (int e0 = threadIdx.x; e0 < N * T; e0 += 4 * blockDim.x) {
        cplx c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * blockDim.x;
            if (e < N * T) c[u] = in[base + (int64_t)(e / T) * inner + (e % T)];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * blockDim.x;
            if (e < N * T) {
                const int t = e % T, row = e / T;
                const int srow = INV ? rt_row_of_freq(P, row) : row;
                buf[pidx(srow) * T + t] = c[u];
            }
        }
    }
    __syncthreads();
    rt_fft_inplace<INV>(buf, T, tw, P);
    for (int e = threadIdx.x; e < N * T; e += blockDim.x) {
        const int t = e % T, row = e / T;
        const int srow = INV ? row : rt_row_of_freq(P, row);
        out[base + (int64_t)row * inner + t] = buf[pidx(srow) * T + t];
    }
}

// S3
template <int T, int KIND, int DIM>
__global__ void __launch_bounds__(384) k_mid_green_rt(cplx* __restrict__ data, const cplx* __restrict__ tw, RtPlan P,
                                                       GreenDesc g, int64_t inner, int nh, int pitch) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    constexpr int L = D * T;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf = reinterpret_cast<cplx*>(fh_smem_raw);
    const int N = P.n;
    const int64_t i0 = (int64_t)blockIdx.x * T;
    const int tot = D * N * T;
    for (int e0 = threadIdx.x; e0 < tot; e0 += 4 * blockDim.x) {
        cplx c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * blockDim.x;
            if (e < tot) {
                const int t = e % T, row = (e / T) % N, cc = e / (T * N);
                c[u] = data[((int64_t)cc * N + row) * inner + i0 + t];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * blockDim.x;
            if (e < tot) {
                const int t = e % T, row = (e / T) % N, cc = e / (T * N);
                buf[pidx(row) * L + cc * T + t] = c[u];
            }
        }
    }
    __syncthreads();
    rt_fft_inplace<false>(buf, L, tw, P);
    for (int idx = threadIdx.x; idx < N * T; idx += blockDim.x) {
        const int row = idx / T, tt = idx - row * T;
        int k[3];
        k[0] = fh_freq(rt_freq_of_row(P, row), N);
        const int64_t ii = i0 + tt;
        bool valid = true;
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
        cplx* sr = buf + pidx(row) * L + tt;
        cplx e[D];
#pragma unroll
        for (int cc = 0; cc < D; ++cc) e[cc] = sr[cc * T];
        if (valid) {
            green_apply<KIND, DIM>(g, k, e);
        } else {
#pragma unroll
            for (int cc = 0; cc < D; ++cc) e[cc] = make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int cc = 0; cc < D; ++cc) sr[cc * T] = e[cc];
    }
    __syncthreads();
    rt_fft_inplace<true>(buf, L, tw, P);
    for (int e = threadIdx.x; e < tot; e += blockDim.x) {
        const int t = e % T, row = (e / T) % N, cc = e / (T * N);
        data[((int64_t)cc * N + row) * inner + i0 + t] = buf[pidx(row) * L + cc * T + t];
    }
}

// S1 (any N, any number of rows: the last CTA masks rows >= nrows)
template <int D, int TRW, int ALAY>
__global__ void __launch_bounds__(256, 2)
    k_fwd_last_rt(const double* __restrict__ A, const unsigned char* __restrict__ phase, const double* __restrict__ lut,
                  const Lut2C lutc, int nphase, double* __restrict__ p, const double* __restrict__ r,
                  const double* __restrict__ scal, int pupdate, cplx* __restrict__ spec, const cplx* __restrict__ tw,
                  RtPlan P, int64_t nrows, int nh, int pitch, double* __restrict__ xacc) {
    constexpr int NL = D * TRW, NP = NL / 2, NT = 256;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf = reinterpret_cast<cplx*>(fh_smem_raw);  // [npr][NP]
    double* bd = reinterpret_cast<double*>(buf);
    const int N = P.n;
    const int64_t row0 = (int64_t)blockIdx.x * TRW;
    const int64_t n = nrows * N;
    const double beta = pupdate ? scal[3] : 0.0;
    const double alpha = (pupdate && xacc) ? scal[2] : 0.0;
    __shared__ double slut[(ALAY == 2) ? 16 * D * D : 1];
    if (ALAY == 2) {
        for (int i = threadIdx.x; i < nphase * D * D; i += NT) slut[i] = lut[i];
        __syncthreads();
    }
    for (int v = threadIdx.x; v < TRW * N; v += NT) {
        const int row = v / N, i2 = v - row * N;
        const bool live = row0 + row < nrows;
        const int64_t gv = (row0 + row) * N + i2;
        int ph = 0;
        if ((ALAY == 2 || ALAY == 3) && live) ph = phase[gv];
        double pv[D];
#pragma unroll
        for (int jj = 0; jj < D; ++jj) {
            double q = 0.0;
            if (live) {
                q = p[(size_t)jj * n + gv];
                if (pupdate) {
                    if (xacc) xacc[(size_t)jj * n + gv] = xacc[(size_t)jj * n + gv] + alpha * q;  // deferred x += alpha p
                    q = r[(size_t)jj * n + gv] + beta * q;
                    p[(size_t)jj * n + gv] = q;
                }
            }
            pv[jj] = q;
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double s = 0.0;
            if (ALAY < 0) {
                s = pv[i];
            } else if (live) {
#pragma unroll
                for (int jj = 0; jj < D; ++jj) {
                    double a;
                    if (ALAY == 3) {
                        a = ph ? lutc.c[1][i * D + jj] : lutc.c[0][i * D + jj];
                    } else if (ALAY == 2) {
                        a = slut[ph * D * D + i * D + jj];
                    } else if (ALAY == 1) {
                        const int lo = i < jj ? i : jj, hi = i < jj ? jj : i;
                        a = A[((size_t)lo * D + hi) * n + gv];
                    } else {
                        a = A[((size_t)i * D + jj) * n + gv];
                    }
                    s += a * pv[jj];
                }
            }
            const int Lr = i * TRW + row;
            bd[2 * (pidx(i2) * NP + (Lr >> 1)) + (Lr & 1)] = s;
        }
    }
    __syncthreads();
    rt_fft_inplace<false>(buf, NP, tw, P);
    for (int it = threadIdx.x; it < NL * pitch; it += NT) {
        const int Lr = it / pitch, k = it - Lr * pitch;
        const int c = Lr / TRW, row = Lr - c * TRW;
        if (row0 + row >= nrows) continue;
        cplx X = make_double2(0.0, 0.0);
        if (k < nh) {
            const int km = (k == 0) ? 0 : N - k;
            const cplx a = buf[pidx(rt_row_of_freq(P, k)) * NP + (Lr >> 1)];
            const cplx b = buf[pidx(rt_row_of_freq(P, km)) * NP + (Lr >> 1)];
            X = (Lr & 1) ? make_double2(0.5 * (a.y + b.y), -0.5 * (a.x - b.x))
                         : make_double2(0.5 * (a.x + b.x), 0.5 * (a.y - b.y));
        }
        spec[((size_t)c * nrows + row0 + row) * pitch + k] = X;
    }
}

// S5
template <int D, int TRW>
__global__ void __launch_bounds__(256, 2)
    k_inv_last_rt(const cplx* __restrict__ spec, double* __restrict__ y, const double* __restrict__ pdot,
                  double* __restrict__ part, const cplx* __restrict__ tw, RtPlan P, int64_t nrows, int nh, int pitch,
                  double scale) {
    constexpr int NL = D * TRW, NP = NL / 2, NT = 256;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    __shared__ double red[32];
    cplx* buf = reinterpret_cast<cplx*>(fh_smem_raw);
    const int N = P.n;
    const int64_t row0 = (int64_t)blockIdx.x * TRW;
    for (int it = threadIdx.x; it < NP * nh; it += NT) {
        const int pr = it / nh, k = it - pr * nh;
        const int La = 2 * pr, Lb = 2 * pr + 1;
        const int ca = La / TRW, ra = La - ca * TRW, cb = Lb / TRW, rb = Lb - cb * TRW;
        cplx a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
        if (row0 + ra < nrows) a = spec[((size_t)ca * nrows + row0 + ra) * pitch + k];
        if (row0 + rb < nrows) b = spec[((size_t)cb * nrows + row0 + rb) * pitch + k];
        if (k == 0 || 2 * k == N) {
            a.y = 0.0;
            b.y = 0.0;
        }
        buf[pidx(rt_row_of_freq(P, k)) * NP + pr] = make_double2(a.x - b.y, a.y + b.x);
        if (k > 0 && 2 * k != N) buf[pidx(rt_row_of_freq(P, N - k)) * NP + pr] = make_double2(a.x + b.y, -a.y + b.x);
    }
    __syncthreads();
    rt_fft_inplace<true>(buf, NP, tw, P);
    double acc = 0.0;
    for (int it = threadIdx.x; it < NL * N; it += NT) {
        const int Lr = it / N, i2 = it - Lr * N;
        const int c = Lr / TRW, row = Lr - c * TRW;
        if (row0 + row >= nrows) continue;
        const cplx z = buf[pidx(i2) * NP + (Lr >> 1)];
        const double v = ((Lr & 1) ? z.y : z.x) * scale;
        const size_t o = ((size_t)c * nrows + row0 + row) * N + i2;
        y[o] = v;
        if (pdot) acc += pdot[o] * v;
    }
    if (pdot) {
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) part[blockIdx.x] = acc;
    }
}


// ------------------------------------------------------------------ slab-exchange addressing
// The multi-GPU pipeline moves the half spectrum between x-slabs and y-slabs with all-to-all.  So that
// no pack/unpack pass is needed, the strided-axis kernels address the exchange buffers directly: a
// line (panel o, row, column t) lives at  base(o) + rowoff(row) + t.
// S2 / S4 between the natural x-slab spectrum and an exchange buffer (out of place)
template <int N, int T, bool INV>
__global__ void __launch_bounds__(256) k_c2c_map(const cplx* __restrict__ in, cplx* __restrict__ out,
                                                  const cplx* __restrict__ tw, LineMap mi, LineMap mo, int ntile) {
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf = reinterpret_cast<cplx*>(fh_smem_raw);  // [N + N/16][T]
    const int64_t o = blockIdx.x / ntile;
    const int tile = blockIdx.x - (int)(o * ntile);
    const int64_t bi = linemap_base(mi, o) + (int64_t)tile * T;
    const int64_t bo = linemap_base(mo, o) + (int64_t)tile * T;
    constexpr int U = 4, NT = 256;
    for (int e0 = threadIdx.x; e0 < N * T; e0 += U * NT) {
        cplx c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * NT;
            if (e < N * T) c[u] = in[bi + linemap_row(mi, e / T) + (e % T)];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * NT;
            if (e < N * T) {
                const int t = e % T, row = e / T;
                const int srow = INV ? smem_row_of_freq<N>(row) : row;
                buf[pidx(srow) * T + t] = c[u];
            }
        }
    }
    __syncthreads();
    smem_fft_inplace<N, INV>(buf, T, tw);
    for (int e = threadIdx.x; e < N * T; e += NT) {
        const int t = e % T, row = e / T;
        const int srow = INV ? row : smem_row_of_freq<N>(row);
        out[bo + linemap_row(mo, row) + t] = buf[pidx(srow) * T + t];
    }
}

// S3 in place on an exchange buffer: element (c, row, ii) at rowoff[row] + c * cstride + ii
template <int N, int T, int KIND, int DIM>
__global__ void __launch_bounds__(384) k_mid_green_map(cplx* __restrict__ data, const cplx* __restrict__ tw, GreenDesc g,
                                                        const int64_t* __restrict__ rowoff, int64_t cstride, int nh,
                                                        int pitch) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    constexpr int L = D * T, NT = 384, U = 4;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf = reinterpret_cast<cplx*>(fh_smem_raw);  // [N + N/16][D*T]
    const int64_t i0 = (int64_t)blockIdx.x * T;
    constexpr int TOT = D * N * T;
    for (int e0 = threadIdx.x; e0 < TOT; e0 += U * NT) {
        cplx c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * NT;
            if (e < TOT) {
                const int t = e % T, row = (e / T) % N, cc = e / (T * N);
                c[u] = data[rowoff[row] + (int64_t)cc * cstride + i0 + t];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * NT;
            if (e < TOT) {
                const int t = e % T, row = (e / T) % N, cc = e / (T * N);
                buf[pidx(row) * L + cc * T + t] = c[u];
            }
        }
    }
    __syncthreads();
    smem_fft_inplace<N, false>(buf, L, tw);
    for (int idx = threadIdx.x; idx < N * T; idx += NT) {
        const int row = idx / T, tt = idx - row * T;
        int k[3];
        k[0] = fh_freq(smem_freq_of_row<N>(row), N);
        const int64_t ii = i0 + tt;
        const int i1 = (int)(ii / pitch), i2 = (int)(ii - (int64_t)i1 * pitch);
        k[1] = fh_freq(i1 + g.ioff1, g.N[1]);
        k[2] = fh_freq(i2, g.N[2]);
        const bool valid = i2 < nh;
        cplx* sr = buf + pidx(row) * L + tt;
        cplx e[D];
#pragma unroll
        for (int cc = 0; cc < D; ++cc) e[cc] = sr[cc * T];
        if (valid) {
            green_apply<KIND, DIM>(g, k, e);
        } else {
#pragma unroll
            for (int cc = 0; cc < D; ++cc) e[cc] = make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int cc = 0; cc < D; ++cc) sr[cc * T] = e[cc];
    }
    __syncthreads();
    smem_fft_inplace<N, true>(buf, L, tw);
    for (int e = threadIdx.x; e < TOT; e += NT) {
        const int t = e % T, row = (e / T) % N, cc = e / (T * N);
        data[rowoff[row] + (int64_t)cc * cstride + i0 + t] = buf[pidx(row) * L + cc * T + t];
    }
}


// S3 on an exchange layout with cluster-cooperative global access.  A CTA can hold only T = 1..4 columns
// of all D*N lines, i.e. 16..64-byte row segments — fine for HBM sectors, poor for NVLink packets when the
// rows live in a peer's memory.  A thread-block cluster of CS CTAs therefore moves CS*T-column segments
// (128 B and more): CTA q of the cluster loads every CS-th (component,row) pair at full segment width and
// deals the columns out to their owners' shared memory over DSMEM; the store phase gathers the same way.
// FFT / G^ / inverse FFT run on each CTA's own tile exactly as in k_mid_green_map.
template <int N, int T, int CS, int KIND>
__global__ void __launch_bounds__(384) k_mid_green_mapc(cplx* __restrict__ data, const cplx* __restrict__ tw,
                                                         GreenDesc g, const int64_t* __restrict__ rowoff,
                                                         int64_t cstride, int nh, int pitch) {
    namespace cg = cooperative_groups;
    constexpr int DIM = 3;
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    constexpr int L = D * T, NT = 384, U = 4, TC = CS * T;
    constexpr int PAIRS = D * N / CS;  // (component,row) pairs per CTA in the load / store phases
    constexpr int TOT = PAIRS * TC;
    static_assert((D * N) % CS == 0, "cluster size must divide D*N");
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf = reinterpret_cast<cplx*>(fh_smem_raw);  // [N + N/16][D*T]
    cg::cluster_group cl = cg::this_cluster();
    const unsigned q = cl.block_rank();
    const int64_t ic = (int64_t)(blockIdx.x / CS) * TC;  // first column of the cluster's segment
    const int64_t i0 = ic + (int64_t)q * T;              // first column of this CTA's tile
    cplx* peer[CS];
#pragma unroll
    for (int u = 0; u < CS; ++u) peer[u] = cl.map_shared_rank(buf, u);
    cl.sync();  // every CTA of the cluster is resident before its shared memory is written remotely
    for (int e0 = threadIdx.x; e0 < TOT; e0 += U * NT) {
        cplx c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * NT;
            if (e < TOT) {
                const int col = e % TC, pi = (e / TC) * CS + (int)q;
                const int cc = pi / N, row = pi - cc * N;
                c[u] = data[rowoff[row] + (int64_t)cc * cstride + ic + col];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * NT;
            if (e < TOT) {
                const int col = e % TC, pi = (e / TC) * CS + (int)q;
                const int cc = pi / N, row = pi - cc * N;
                peer[col / T][pidx(row) * L + cc * T + (col % T)] = c[u];
            }
        }
    }
    cl.sync();
    smem_fft_inplace<N, false>(buf, L, tw);
    for (int idx = threadIdx.x; idx < N * T; idx += NT) {
        const int row = idx / T, tt = idx - row * T;
        int k[3];
        k[0] = fh_freq(smem_freq_of_row<N>(row), N);
        const int64_t ii = i0 + tt;
        const int i1 = (int)(ii / pitch), i2 = (int)(ii - (int64_t)i1 * pitch);
        k[1] = fh_freq(i1 + g.ioff1, g.N[1]);
        k[2] = fh_freq(i2, g.N[2]);
        const bool valid = i2 < nh;
        cplx* sr = buf + pidx(row) * L + tt;
        cplx e[D];
#pragma unroll
        for (int cc = 0; cc < D; ++cc) e[cc] = sr[cc * T];
        if (valid) {
            green_apply<KIND, DIM>(g, k, e);
        } else {
#pragma unroll
            for (int cc = 0; cc < D; ++cc) e[cc] = make_double2(0.0, 0.0);
        }
#pragma unroll
        for (int cc = 0; cc < D; ++cc) sr[cc * T] = e[cc];
    }
    __syncthreads();
    smem_fft_inplace<N, true>(buf, L, tw);
    cl.sync();
    for (int e0 = threadIdx.x; e0 < TOT; e0 += U * NT) {
        cplx c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * NT;
            if (e < TOT) {
                const int col = e % TC, pi = (e / TC) * CS + (int)q;
                const int cc = pi / N, row = pi - cc * N;
                c[u] = peer[col / T][pidx(row) * L + cc * T + (col % T)];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * NT;
            if (e < TOT) {
                const int col = e % TC, pi = (e / TC) * CS + (int)q;
                const int cc = pi / N, row = pi - cc * N;
                data[rowoff[row] + (int64_t)cc * cstride + ic + col] = c[u];
            }
        }
    }
    cl.sync();  // no CTA may exit while a neighbour still reads its tile
}


// Persistent, software-pipelined form of k_mid_green_mapc: one cluster per CS SMs walks over the segments;
// the (remote) loads of the NEXT segment are issued into registers before the transforms of the current
// one and consumed after them, and the stores are fire-and-forget, so NVLink traffic in both directions
// overlaps the FFT / G^ arithmetic instead of alternating with it (CTAs that all wait on the same link
// fall into lock step otherwise).  D*N*T <= 6144 elements per CTA (16 complex numbers per thread in flight).
template <int N, int T, int CS, int KIND>
__global__ void __launch_bounds__(384, 1) k_mid_green_mapp(cplx* __restrict__ data, const cplx* __restrict__ tw,
                                                            GreenDesc g, const int64_t* __restrict__ rowoff,
                                                            int64_t cstride, int nh, int pitch, int64_t nseg) {
    namespace cg = cooperative_groups;
    constexpr int DIM = 3;
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    constexpr int L = D * T, NT = 384, U = 4, TC = CS * T;
    constexpr int PAIRS = D * N / CS;
    constexpr int TOT = PAIRS * TC;
    constexpr int UP = (TOT + NT - 1) / NT;
    static_assert((D * N) % CS == 0 && UP <= 16, "segment does not fit the register pipeline");
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* buf = reinterpret_cast<cplx*>(fh_smem_raw);  // [N + N/16][D*T]
    cg::cluster_group cl = cg::this_cluster();
    const unsigned q = cl.block_rank();
    const int64_t ncl = gridDim.x / CS;
    cplx* peer[CS];
#pragma unroll
    for (int u = 0; u < CS; ++u) peer[u] = cl.map_shared_rank(buf, u);
    cplx pre[UP];
    int64_t seg = blockIdx.x / CS;
    if (seg < nseg) {
        const int64_t ic = seg * TC;
#pragma unroll
        for (int u = 0; u < UP; ++u) {
            const int e = threadIdx.x + u * NT;
            if (e < TOT) {
                const int col = e % TC, pi = (e / TC) * CS + (int)q;
                const int cc = pi / N, row = pi - cc * N;
                pre[u] = data[rowoff[row] + (int64_t)cc * cstride + ic + col];
            }
        }
    }
    cl.sync();
    for (; seg < nseg; seg += ncl) {
        const int64_t ic = seg * TC;
        const int64_t i0 = ic + (int64_t)q * T;
#pragma unroll
        for (int u = 0; u < UP; ++u) {
            const int e = threadIdx.x + u * NT;
            if (e < TOT) {
                const int col = e % TC, pi = (e / TC) * CS + (int)q;
                const int cc = pi / N, row = pi - cc * N;
                peer[col / T][pidx(row) * L + cc * T + (col % T)] = pre[u];
            }
        }
        cl.sync();
        if (seg + ncl < nseg) {  // next segment: in flight during the transforms
            const int64_t icn = (seg + ncl) * TC;
#pragma unroll
            for (int u = 0; u < UP; ++u) {
                const int e = threadIdx.x + u * NT;
                if (e < TOT) {
                    const int col = e % TC, pi = (e / TC) * CS + (int)q;
                    const int cc = pi / N, row = pi - cc * N;
                    pre[u] = data[rowoff[row] + (int64_t)cc * cstride + icn + col];
                }
            }
        }
        smem_fft_inplace<N, false>(buf, L, tw);
        for (int idx = threadIdx.x; idx < N * T; idx += NT) {
            const int row = idx / T, tt = idx - row * T;
            int k[3];
            k[0] = fh_freq(smem_freq_of_row<N>(row), N);
            const int64_t ii = i0 + tt;
            const int i1 = (int)(ii / pitch), i2 = (int)(ii - (int64_t)i1 * pitch);
            k[1] = fh_freq(i1 + g.ioff1, g.N[1]);
            k[2] = fh_freq(i2, g.N[2]);
            const bool valid = i2 < nh;
            cplx* sr = buf + pidx(row) * L + tt;
            cplx e[D];
#pragma unroll
            for (int cc = 0; cc < D; ++cc) e[cc] = sr[cc * T];
            if (valid) {
                green_apply<KIND, DIM>(g, k, e);
            } else {
#pragma unroll
                for (int cc = 0; cc < D; ++cc) e[cc] = make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int cc = 0; cc < D; ++cc) sr[cc * T] = e[cc];
        }
        __syncthreads();
        smem_fft_inplace<N, true>(buf, L, tw);
        cl.sync();
        for (int e0 = threadIdx.x; e0 < TOT; e0 += U * NT) {
            cplx c[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * NT;
                if (e < TOT) {
                    const int col = e % TC, pi = (e / TC) * CS + (int)q;
                    const int cc = pi / N, row = pi - cc * N;
                    c[u] = peer[col / T][pidx(row) * L + cc * T + (col % T)];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = e0 + u * NT;
                if (e < TOT) {
                    const int col = e % TC, pi = (e / TC) * CS + (int)q;
                    const int cc = pi / N, row = pi - cc * N;
                    data[rowoff[row] + (int64_t)cc * cstride + ic + col] = c[u];
                }
            }
        }
        cl.sync();  // tiles are free for the next deal (and nobody exits while a neighbour reads its tile)
    }
}
