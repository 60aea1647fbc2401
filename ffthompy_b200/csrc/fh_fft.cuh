// fh_fft.cuh — shared-memory-staged mixed-radix Stockham FFT building blocks
// (fp64 complex) and the axis-pass kernels built on them.
//
// Replaces numpy.fft.{rfftn,irfftn,fftn,ifftn} as called from
// ffthompy/tensors/fft.py:33-43 (pocketfft, single thread).  Conventions:
// forward = exp(-2*pi*i*k*x/n), un-normalised; inverse = exp(+...), the caller
// chooses the scale factor (1/prod(N) for the 'r' form).
//
// Layout in shared memory: T independent lines of length n are interleaved as
// SoA  re[n][Tp], im[n][Tp]  (Tp = T padded to an odd number, so that both the
// "t fastest" and the "row fastest" thread mappings are bank-conflict free for
// 8-byte words).  One radix pass maps  in[(j + r*n/R)] -> out[j0 + r*Ns]
// (Stockham autosort, decimation in time), so after the last pass the spectrum
// is in natural order and no bit-reversal pass exists.
#pragma once
#include "fh_common.cuh"
#include "fh_roots.cuh"

#define FH_MAX_FAC 12

struct AxisDesc {
    int n;                // transform length
    int nfac;             // number of radix passes
    int fac[FH_MAX_FAC];  // radices, product == n
    const cplx* tw;       // device table tw[m] = exp(-2*pi*i*m/n), m = 0..n-1
};

// ------------------------------------------------------------------ butterflies
template <bool INV>
__device__ __forceinline__ cplx mul_mi(cplx a) {  // forward: multiply by -i ; inverse: by +i
    return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}

template <int R, bool INV>
struct Bfly;

template <bool INV>
struct Bfly<2, INV> {
    static __device__ __forceinline__ void run(cplx* v) {
        cplx a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <bool INV>
struct Bfly<4, INV> {
    static __device__ __forceinline__ void run(cplx* v) {
        cplx t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        cplx t2 = cadd(v[1], v[3]), t3 = mul_mi<INV>(csub(v[1], v[3]));
        v[0] = cadd(t0, t2);
        v[1] = cadd(t1, t3);
        v[2] = csub(t0, t2);
        v[3] = csub(t1, t3);
    }
};

template <bool INV>
struct Bfly<8, INV> {
    static __device__ __forceinline__ void run(cplx* v) {
        const double c = 0.70710678118654752440;
        cplx e[4] = {v[0], v[2], v[4], v[6]};
        cplx o[4] = {v[1], v[3], v[5], v[7]};
        Bfly<4, INV>::run(e);
        Bfly<4, INV>::run(o);
        // o[k] *= W8^k
        cplx o1 = INV ? make_double2(c * (o[1].x - o[1].y), c * (o[1].x + o[1].y))
                      : make_double2(c * (o[1].x + o[1].y), c * (o[1].y - o[1].x));
        cplx o2 = mul_mi<INV>(o[2]);
        cplx o3 = INV ? make_double2(-c * (o[3].x + o[3].y), c * (o[3].x - o[3].y))
                      : make_double2(c * (o[3].y - o[3].x), -c * (o[3].x + o[3].y));
        v[0] = cadd(e[0], o[0]);
        v[4] = csub(e[0], o[0]);
        v[1] = cadd(e[1], o1);
        v[5] = csub(e[1], o1);
        v[2] = cadd(e[2], o2);
        v[6] = csub(e[2], o2);
        v[3] = cadd(e[3], o3);
        v[7] = csub(e[3], o3);
    }
};

template <bool INV>
struct Bfly<16, INV> {
    static __device__ __forceinline__ void run(cplx* v) {
        // cos/sin(k*pi/8), k = 1..7
        const double cs[8] = {1.0,
                              0.92387953251128675613,
                              0.70710678118654752440,
                              0.38268343236508977173,
                              0.0,
                              -0.38268343236508977173,
                              -0.70710678118654752440,
                              -0.92387953251128675613};
        const double sn[8] = {0.0,
                              0.38268343236508977173,
                              0.70710678118654752440,
                              0.92387953251128675613,
                              1.0,
                              0.92387953251128675613,
                              0.70710678118654752440,
                              0.38268343236508977173};
        cplx e[8], o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            e[k] = v[2 * k];
            o[k] = v[2 * k + 1];
        }
        Bfly<8, INV>::run(e);
        Bfly<8, INV>::run(o);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            cplx w = make_double2(cs[k], INV ? sn[k] : -sn[k]);
            cplx t = (k == 0) ? o[0] : cmul(o[k], w);
            v[k] = cadd(e[k], t);
            v[k + 8] = csub(e[k], t);
        }
    }
};

// Odd-radix butterflies (3, 5, 7, 9, 11, 13, 15, 17, 19): direct DFT in its conjugate-symmetric form,
//   a_r = v_r + v_{R-r},  b_r = v_r - v_{R-r}   (r = 1..h, h = (R-1)/2)
//   y_q, y_{R-q} = v_0 + sum_r a_r cos(2 pi q r/R)  -/+  i * sum_r b_r sin(2 pi q r/R)   (forward)
// i.e. (R-1)^2 real FMAs per butterfly; the roots are compile-time constants (fh_roots.cuh).
template <int R, bool INV>
__device__ __forceinline__ void bfly_direct(cplx* v, const cplx* __restrict__ /*tw*/, int /*nR*/) {
    constexpr int H = (R - 1) / 2;
    cplx a[H], b[H];
#pragma unroll
    for (int r = 1; r <= H; ++r) {
        a[r - 1] = cadd(v[r], v[R - r]);
        b[r - 1] = csub(v[r], v[R - r]);
    }
    cplx y0 = v[0];
#pragma unroll
    for (int r = 0; r < H; ++r) {
        y0.x += a[r].x;
        y0.y += a[r].y;
    }
    cplx y[R];
    y[0] = y0;
#pragma unroll
    for (int q = 1; q <= H; ++q) {
        double sx = v[0].x, sy = v[0].y, tx = 0.0, ty = 0.0;
#pragma unroll
        for (int r = 1; r <= H; ++r) {
            const double c = OddRoots<R>::c((q * r) % R);
            const double sn = OddRoots<R>::s((q * r) % R);
            sx += a[r - 1].x * c;
            sy += a[r - 1].y * c;
            tx += b[r - 1].y * sn;
            ty += b[r - 1].x * sn;
        }
        // forward: v e^{-i t}: real a.x c + b.y s, imag a.y c - b.x s ; inverse: signs of s flipped
        if (INV) {
            y[q] = make_double2(sx - tx, sy + ty);
            y[R - q] = make_double2(sx + tx, sy - ty);
        } else {
            y[q] = make_double2(sx + tx, sy - ty);
            y[R - q] = make_double2(sx - tx, sy + ty);
        }
    }
#pragma unroll
    for (int q = 0; q < R; ++q) v[q] = y[q];
}

// ------------------------------------------------------------------ one radix pass
// All threads of the CTA cooperate; nl = number of interleaved lines, ld = Tp.
template <int R, bool INV>
__device__ __forceinline__ void stockham_pass(const double* __restrict__ ire, const double* __restrict__ iim,
                                              double* __restrict__ ore, double* __restrict__ oim, int n, int nl,
                                              int ld, int Ns, const cplx* __restrict__ tw) {
    const int nR = n / R;
    const int nb = nR * nl;
    const int twstep = n / (Ns * R);
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        const int t = b % nl;
        const int j = b / nl;
        const int k = j % Ns;
        const int j0 = (j - k) * R + k;
        cplx v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int a = (j + r * nR) * ld + t;
            v[r] = make_double2(ire[a], iim[a]);
        }
        if (k != 0) {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                cplx w = __ldg(&tw[r * k * twstep]);
                if (INV) w.y = -w.y;
                v[r] = cmul(v[r], w);
            }
        }
        if constexpr (R == 2 || R == 4 || R == 8 || R == 16)
            Bfly<R, INV>::run(v);
        else
            bfly_direct<R, INV>(v, tw, nR);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int a = (j0 + r * Ns) * ld + t;
            ore[a] = v[r].x;
            oim[a] = v[r].y;
        }
    }
}

// Runtime prime radix p (any p): one work item per output element.
template <bool INV>
__device__ __forceinline__ void stockham_pass_prime(const double* __restrict__ ire, const double* __restrict__ iim,
                                                    double* __restrict__ ore, double* __restrict__ oim, int n, int nl,
                                                    int ld, int Ns, int p, const cplx* __restrict__ tw) {
    const int nR = n / p;
    const int nitems = n * nl;  // (j, q, t)
    const int twstep = n / (Ns * p);
    for (int it = threadIdx.x; it < nitems; it += blockDim.x) {
        const int t = it % nl;
        const int jq = it / nl;
        const int j = jq % nR;
        const int q = jq / nR;
        const int k = j % Ns;
        const int j0 = (j - k) * p + k;
        double ax = 0.0, ay = 0.0;
        const int base_e = k * twstep;  // exponent increment per r from the inter-pass twiddle
        int e = 0;                      // (r*k*twstep + ((q*r) % p) * nR) mod n, built incrementally
        int qr = 0;
        for (int r = 0; r < p; ++r) {
            int ee = e + qr * nR;
            if (ee >= n) ee -= n;
            cplx w = __ldg(&tw[ee]);
            if (INV) w.y = -w.y;
            const int a = (j + r * nR) * ld + t;
            const double xr = ire[a], xi = iim[a];
            ax += xr * w.x - xi * w.y;
            ay += xr * w.y + xi * w.x;
            e += base_e;
            if (e >= n) e -= n;
            qr += q;
            if (qr >= p) qr -= p;
        }
        const int a = (j0 + q * Ns) * ld + t;
        ore[a] = ax;
        oim[a] = ay;
    }
}

// Full in-smem FFT of `nl` interleaved lines.  Data starts in (b0re,b0im); the
// function ping-pongs with (b1re,b1im) and returns 0/1 = which buffer holds the
// result.  Every pass ends with __syncthreads(); the caller must have synced
// after filling buffer 0.
template <bool INV>
__device__ __forceinline__ int fft_smem(double* b0re, double* b0im, double* b1re, double* b1im, const AxisDesc& ax,
                                        int nl, int ld) {
    int cur = 0;
    int Ns = 1;
    for (int f = 0; f < ax.nfac; ++f) {
        const int R = ax.fac[f];
        const double* ire = cur ? b1re : b0re;
        const double* iim = cur ? b1im : b0im;
        double* ore = cur ? b0re : b1re;
        double* oim = cur ? b0im : b1im;
        switch (R) {
            case 2: stockham_pass<2, INV>(ire, iim, ore, oim, ax.n, nl, ld, Ns, ax.tw); break;
            case 3: stockham_pass<3, INV>(ire, iim, ore, oim, ax.n, nl, ld, Ns, ax.tw); break;
            case 4: stockham_pass<4, INV>(ire, iim, ore, oim, ax.n, nl, ld, Ns, ax.tw); break;
            case 5: stockham_pass<5, INV>(ire, iim, ore, oim, ax.n, nl, ld, Ns, ax.tw); break;
            case 7: stockham_pass<7, INV>(ire, iim, ore, oim, ax.n, nl, ld, Ns, ax.tw); break;
            case 8: stockham_pass<8, INV>(ire, iim, ore, oim, ax.n, nl, ld, Ns, ax.tw); break;
            case 16: stockham_pass<16, INV>(ire, iim, ore, oim, ax.n, nl, ld, Ns, ax.tw); break;
            case 9: stockham_pass<9, INV>(ire, iim, ore, oim, ax.n, nl, ld, Ns, ax.tw); break;
            case 11: stockham_pass<11, INV>(ire, iim, ore, oim, ax.n, nl, ld, Ns, ax.tw); break;
            case 13: stockham_pass<13, INV>(ire, iim, ore, oim, ax.n, nl, ld, Ns, ax.tw); break;
            case 15: stockham_pass<15, INV>(ire, iim, ore, oim, ax.n, nl, ld, Ns, ax.tw); break;
            case 17: stockham_pass<17, INV>(ire, iim, ore, oim, ax.n, nl, ld, Ns, ax.tw); break;
            case 19: stockham_pass<19, INV>(ire, iim, ore, oim, ax.n, nl, ld, Ns, ax.tw); break;
            default: stockham_pass_prime<INV>(ire, iim, ore, oim, ax.n, nl, ld, Ns, R, ax.tw); break;
        }
        __syncthreads();
        Ns *= R;
        cur ^= 1;
    }
    return cur;
}

// smem bytes needed by fft_smem for nl lines of length n (two SoA buffers)
static inline size_t fft_smem_bytes(int n, int ld) { return (size_t)4 * n * ld * sizeof(double); }
static inline int fft_ld(int nl) { return nl | 1; }
