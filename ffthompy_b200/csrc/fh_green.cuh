// fh_green.cuh — closed-form Green / projection multipliers G^(xi), evaluated per
// frequency in registers (never materialised on the hot path).
//
// Reproduces the arrays assembled by ffthompy/projections.py:9-112 (scalar) and
// :114-267 (elasticity, Mandel notation), including the Nyquist zeroing
// (`NyqNul`, band |k_i| <= (Nred_i-1)/2 of the ORIGINAL grid), the zero-padding
// to the doubled grid and the prod(Nbar)/prod(N) scale that Tensor.enlarge puts
// on an 'r'-form multiplier (tensors/objects.py:144-166,428-467).
//
//   xi = 0                : out = c0 * e                       (G0, the mean)
//   xi != 0, inside band  : out = scale*( cI*e + cS*S e + cH*(v v^T) e
//                                         + cL*Lam e + cW*(W+W^T) e )
//   outside band          : out = 0
//
// scalar kind (vector fields, D = d):  only cI and cH (n (x) n) are used.
// elastic kind (D = d(d+1)/2, Mandel): with eps = unMandel(e), w = eps n:
//   S e        = Mandel(n (x) w + w (x) n)          projections.py:185,211-223
//   (v v^T) e  = (n.w) Mandel(n (x) n)              projections.py:187,216-229
//   Lam e      = tr(eps)/d * Mandel(I)              projections.py:188
//   (W+W^T) e  = tr(eps) Mandel(n (x) n) + (n.w) Mandel(I)    projections.py:222-223,239
// so that G1h=(cH=1), G1s=(cS=1,cH=-2), G2h=(cH,cL,cW)=(1,d,-1)/(d-1),
// G2s = I - G1h - G1s - G2h  (projections.py:237-240).
#pragma once
#include "fh_common.cuh"

#define FH_GREEN_SCALAR 0
#define FH_GREEN_ELASTIC 1

struct GreenDesc {
    int kind;
    int dim;
    int N[3];     // grid the multiplier lives on
    int band[3];  // active iff |k_i| <= band[i] for all i
    double Y[3];
    double invY[3];  // 1/Y (host-computed): xi = k * invY, no fp64 division per frequency
    double c0, cI, cS, cH, cL, cW;
    double scale;
    int ioff1;  // added to the axis-1 storage index (slab-decomposed spectra: this rank owns k1 in [ioff1, ioff1+n1l))
};

// e: D complex components (in place).  k: signed integer frequencies.
template <int DIM>
__device__ __forceinline__ void green_scalar(const GreenDesc& g, const int* k, cplx* e) {
    double xi[DIM];
    double s = 0.0;
    bool inband = true;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
        xi[i] = (double)k[i] * g.invY[i];
        s += xi[i] * xi[i];
        inband = inband && (abs(k[i]) <= g.band[i]);
    }
    bool zero = true;
#pragma unroll
    for (int i = 0; i < DIM; ++i) zero = zero && (k[i] == 0);
    if (zero) {
#pragma unroll
        for (int i = 0; i < DIM; ++i) e[i] = cscale(e[i], g.c0);
        return;
    }
    if (!inband) {
#pragma unroll
        for (int i = 0; i < DIM; ++i) e[i] = make_double2(0.0, 0.0);
        return;
    }
    const double inv = 1.0 / s;
    cplx dot = make_double2(0.0, 0.0);
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
        dot.x += xi[i] * e[i].x;
        dot.y += xi[i] * e[i].y;
    }
    dot.x *= inv * g.cH;
    dot.y *= inv * g.cH;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
        e[i].x = g.scale * (g.cI * e[i].x + xi[i] * dot.x);
        e[i].y = g.scale * (g.cI * e[i].y + xi[i] * dot.y);
    }
}

// Real-valued core of the Mandel-elastic multiplier (applied to re and im).  xi = frequency,
// inv = 1/|xi|^2, v = Mandel(xi (x) xi)/|xi|^2 (shared between the real and imaginary part).
template <int DIM>
__device__ __forceinline__ void green_elastic_real(const GreenDesc& g, const double* n /*xi*/, double inv,
                                                   const double* v, double* e) {
    constexpr int D = DIM * (DIM + 1) / 2;
    const double r2 = 0.70710678118654752440;  // 1/sqrt(2)
    const double s2 = 1.41421356237309504880;
    double w[DIM];
    double tr = 0.0;
    if (DIM == 3) {
        const double e12 = e[5] * r2, e13 = e[4] * r2, e23 = e[3] * r2;
        w[0] = e[0] * n[0] + e12 * n[1] + e13 * n[2];
        w[1] = e12 * n[0] + e[1] * n[1] + e23 * n[2];
        w[2] = e13 * n[0] + e23 * n[1] + e[2] * n[2];
        tr = e[0] + e[1] + e[2];
    } else {
        const double e12 = e[2] * r2;
        w[0] = e[0] * n[0] + e12 * n[1];
        w[1] = e12 * n[0] + e[1] * n[1];
        tr = e[0] + e[1];
    }
    double nw = 0.0;
#pragma unroll
    for (int i = 0; i < DIM; ++i) nw += n[i] * w[i];
    nw *= inv;
#pragma unroll
    for (int i = 0; i < DIM; ++i) w[i] *= inv;
    // Mandel(n (x) w + w (x) n), n = xi/|xi| folded in through `inv`
    double Se[D];
#pragma unroll
    for (int i = 0; i < DIM; ++i) Se[i] = 2.0 * n[i] * w[i];
    if (DIM == 3) {
        Se[3] = s2 * (n[1] * w[2] + w[1] * n[2]);
        Se[4] = s2 * (n[0] * w[2] + w[0] * n[2]);
        Se[5] = s2 * (n[0] * w[1] + w[0] * n[1]);
    } else {
        Se[2] = s2 * (n[0] * w[1] + w[0] * n[1]);
    }
    if (g.cI == 0.0 && g.cL == 0.0 && g.cW == 0.0 && g.cS == 1.0 && g.scale == 1.0) {
        // the strain projector G1h + G1s (and G1s alone) on the solve grid: the dropped terms are exact zeros and
        // multiplications by one, so this shorter form rounds exactly like the general one below
        const double cHn1 = g.cH * nw;
#pragma unroll
        for (int m = 0; m < D; ++m) e[m] = Se[m] + cHn1 * v[m];
        return;
    }
    const double cHn = g.cH * nw + g.cW * tr;          // multiplies v
    const double cdiag = g.cL * tr / DIM + g.cW * nw;  // multiplies Mandel(I)
#pragma unroll
    for (int m = 0; m < D; ++m) {
        double o = g.cI * e[m] + g.cS * Se[m] + cHn * v[m];
        if (m < DIM) o += cdiag;
        e[m] = g.scale * o;
    }
}

template <int DIM>
__device__ __forceinline__ void green_elastic(const GreenDesc& g, const int* k, cplx* e) {
    constexpr int D = DIM * (DIM + 1) / 2;
    const double s2 = 1.41421356237309504880;
    double xi[DIM];
    double s = 0.0;
    bool inband = true, zero = true;
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
        xi[i] = (double)k[i] * g.invY[i];
        s += xi[i] * xi[i];
        inband = inband && (abs(k[i]) <= g.band[i]);
        zero = zero && (k[i] == 0);
    }
    if (zero) {
#pragma unroll
        for (int m = 0; m < D; ++m) e[m] = cscale(e[m], g.c0);
        return;
    }
    if (!inband) {
#pragma unroll
        for (int m = 0; m < D; ++m) e[m] = make_double2(0.0, 0.0);
        return;
    }
    const double inv = 1.0 / s;
    double v[D];
#pragma unroll
    for (int i = 0; i < DIM; ++i) v[i] = xi[i] * xi[i] * inv;
    if (DIM == 3) {
        v[3] = s2 * xi[1] * xi[2] * inv;
        v[4] = s2 * xi[0] * xi[2] * inv;
        v[5] = s2 * xi[0] * xi[1] * inv;
    } else {
        v[2] = s2 * xi[0] * xi[1] * inv;
    }
    double re[D], im[D];
#pragma unroll
    for (int m = 0; m < D; ++m) {
        re[m] = e[m].x;
        im[m] = e[m].y;
    }
    green_elastic_real<DIM>(g, xi, inv, v, re);
    green_elastic_real<DIM>(g, xi, inv, v, im);
#pragma unroll
    for (int m = 0; m < D; ++m) e[m] = make_double2(re[m], im[m]);
}

// Dispatch on (kind, dim); D = number of components held in e.
template <int KIND, int DIM>
__device__ __forceinline__ void green_apply(const GreenDesc& g, const int* k, cplx* e) {
    if (KIND == FH_GREEN_SCALAR)
        green_scalar<DIM>(g, k, e);
    else
        green_elastic<DIM>(g, k, e);
}
