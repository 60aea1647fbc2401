// fh_fused.cu — the fused Fourier–Galerkin operator  y = F^-1 G^(xi) F (A x)  and
// the Krylov loops that iterate it on the device.
//
// Reference call path: applications.py:58-81 builds
//   Afun = Operator([[ Operator([[FiN, G^, FN]]), A ]])      (tensors/operators.py:136-144)
// and hands it to linear_solver('CG' | 'richardson')          (general/solver.py:63-139).
// Here the whole operator is a fixed pipeline of axis passes; G^ is evaluated in
// closed form between the forward and inverse transform of axis 0 (fh_green.cuh)
// and is never materialised.
#include "fh_plan.cuh"
#include "fh_green.cuh"
#include "../../include/ffthom_b200.h"
#include <stdlib.h>

int fh_fill_green(GreenDesc& g, const fh_green* in);

struct fh_ga {
    const fh_plan* plan;
    int D;
    const double* A;
    int a_layout;
    GreenDesc g;
    double* work;   // [D*nreal] sigma  |  [2*D*nspec] spectrum
    double* sigma;  // = work
    cplx* spec;     // = work + D*nreal
    // device scalars for the Krylov loops
    double* scal;  // [8]: rr, pAp, alpha, beta, ...
    double* part;  // partial sums
    double* hist_pinned;
    int npart;
};

#define GA_NT 256
#define GA_MAXPART 2048

// ------------------------------------------------------------------ sigma = A x
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

// ------------------------------------------------------------------ axis 0: forward, G^, inverse
// Array [D][n0][inner] (inner = N1*nh in 3-D, nh in 2-D).  One CTA owns T consecutive
// `inner` positions of all D components, so G^ mixes components in shared memory.
template <int KIND, int DIM>
__global__ void __launch_bounds__(GA_NT) k_mid_green(cplx* __restrict__ data, AxisDesc ax, GreenDesc g, int64_t inner,
                                                     int T, int ld, int nh) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    extern __shared__ double sm[];
    const int n = ax.n;
    double* b0re = sm;
    double* b0im = sm + (size_t)n * ld;
    double* b1re = sm + (size_t)2 * n * ld;
    double* b1im = sm + (size_t)3 * n * ld;
    const int64_t i0 = (int64_t)blockIdx.x * T;
    const int nl = (int)min((int64_t)T, inner - i0);
    const int nln = n * nl;
    for (int idx = threadIdx.x; idx < D * nln; idx += blockDim.x) {
        const int c = idx / nln, rem = idx - c * nln;
        const int row = rem / nl, t = rem - row * nl;
        const cplx v = data[((int64_t)c * n + row) * inner + i0 + t];
        b0re[row * ld + c * nl + t] = v.x;
        b0im[row * ld + c * nl + t] = v.y;
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
        if (DIM == 3) {
            const int i1 = (int)(ii / nh), i2 = (int)(ii - (int64_t)i1 * nh);
            k[1] = fh_freq(i1, g.N[1]);
            k[2] = fh_freq(i2, g.N[2]);
        } else {
            k[1] = fh_freq((int)ii, g.N[1]);
            k[2] = 0;
        }
        cplx e[D];
#pragma unroll
        for (int c = 0; c < D; ++c) e[c] = make_double2(cre[row * ld + c * nl + t], cim[row * ld + c * nl + t]);
        green_apply<KIND, DIM>(g, k, e);
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
static int launch_mid_green(fh_ga* op) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    const fh_plan* p = op->plan;
    const AxisDesc& ax = p->ax[0];
    const int64_t inner = (DIM == 3) ? (int64_t)p->N[1] * p->nh : p->nh;
    // tile width: keep two CTAs per SM if possible
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
    k_mid_green<KIND, DIM><<<(unsigned)nblk, GA_NT, smem, fh_stream()>>>(op->spec, ax, op->g, inner, T, ld, p->nh);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// ------------------------------------------------------------------ operator object
// sigma region rounded up to 128 B so that the spectrum behind it stays 16-byte aligned
static int64_t sigma_doubles(const fh_plan* p, int D) { return ((int64_t)D * p->nreal + 15) / 16 * 16; }
extern "C" int64_t fh_ga_work_doubles(const fh_plan* p, int D) {
    if (!p || D < 1) return 0;
    return sigma_doubles(p, D) + 2 * (int64_t)D * p->nspec;
}

extern "C" int fh_ga_create(fh_ga** out, const fh_plan* plan, int D, const double* A, int a_layout, const fh_green* g,
                            double* work) {
    FH_REQUIRE(out && plan && A && g && work, "fh_ga_create: null argument");
    FH_REQUIRE(plan->dim == 2 || plan->dim == 3, "fh_ga_create: dim must be 2 or 3");
    FH_REQUIRE(a_layout == 0, "fh_ga_create: unsupported coefficient layout %d", a_layout);
    const int Dexp = (g->kind == FH_GREEN_SCALAR) ? plan->dim : plan->dim * (plan->dim + 1) / 2;
    FH_REQUIRE(D == Dexp, "fh_ga_create: D=%d does not match Green kind %d in dim %d", D, g->kind, plan->dim);
    for (int a = 0; a < plan->dim; ++a)
        FH_REQUIRE(g->N[a] == plan->N[a], "fh_ga_create: Green descriptor grid differs from the plan grid");
    fh_ga* op = (fh_ga*)calloc(1, sizeof(fh_ga));
    if (!op) return fh_set_error(FH_ERR_ALLOC, "fh_ga_create: out of host memory");
    int rc = fh_fill_green(op->g, g);
    if (rc) {
        free(op);
        return rc;
    }
    op->plan = plan;
    op->D = D;
    op->A = A;
    op->a_layout = a_layout;
    op->work = work;
    op->sigma = work;
    op->spec = (cplx*)(work + sigma_doubles(plan, D));
    FH_REQUIRE(((uintptr_t)work & 15) == 0, "fh_ga_create: work must be 16-byte aligned");
    cudaError_t e = cudaMalloc((void**)&op->scal, sizeof(double) * (16 + GA_MAXPART));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&op->hist_pinned, sizeof(double) * 16);
    if (e != cudaSuccess) {
        free(op);
        return fh_set_error(FH_ERR_CUDA, "fh_ga_create: %s", cudaGetErrorString(e));
    }
    op->part = op->scal + 16;
    *out = op;
    return FH_OK;
}

extern "C" int fh_ga_destroy(fh_ga* op) {
    if (!op) return FH_OK;
    cudaFree(op->scal);
    cudaFreeHost(op->hist_pinned);
    free(op);
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

extern "C" int fh_ga_apply(fh_ga* op, const double* x, double* y) {
    FH_REQUIRE(op && x && y, "fh_ga_apply: null argument");
    const fh_plan* p = op->plan;
    const int D = op->D;
    const int64_t n = p->nreal;
    const unsigned g = ga_grid(n);
    switch (D) {
        case 2: k_apply_A<2><<<g, GA_NT, 0, fh_stream()>>>(n, op->A, x, op->sigma); break;
        case 3: k_apply_A<3><<<g, GA_NT, 0, fh_stream()>>>(n, op->A, x, op->sigma); break;
        case 6: k_apply_A<6><<<g, GA_NT, 0, fh_stream()>>>(n, op->A, x, op->sigma); break;
        default: return fh_set_error(FH_ERR_UNSUPPORTED, "fh_ga_apply: D=%d", D);
    }
    FH_LAUNCH_CHECK();
    int rc;
    const int64_t nlines = (int64_t)D * (p->nreal / p->N[p->dim - 1]);
    if ((rc = fh_launch_r2c_last(p, op->sigma, op->spec, nlines))) return rc;
    if (p->dim == 3)
        if ((rc = fh_launch_c2c_strided(p->ax[1], op->spec, op->spec, (int64_t)D * p->N[0], p->nh, false, 1.0))) return rc;
    if (op->g.kind == FH_GREEN_SCALAR)
        rc = (p->dim == 3) ? launch_mid_green<FH_GREEN_SCALAR, 3>(op) : launch_mid_green<FH_GREEN_SCALAR, 2>(op);
    else
        rc = (p->dim == 3) ? launch_mid_green<FH_GREEN_ELASTIC, 3>(op) : launch_mid_green<FH_GREEN_ELASTIC, 2>(op);
    if (rc) return rc;
    if (p->dim == 3)
        if ((rc = fh_launch_c2c_strided(p->ax[1], op->spec, op->spec, (int64_t)D * p->N[0], p->nh, true, 1.0))) return rc;
    return fh_launch_c2r_last(p, op->spec, y, nlines, 1.0 / (double)p->nreal);
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

// x += alpha p ; r -= alpha Ap ; partial r.r   (solver.py:127-129)
__global__ void k_cg_update(int64_t n, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
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

static int read_norm(fh_ga* op, double* out) {
    FH_CUDA(cudaMemcpyAsync(op->hist_pinned, op->scal + 4, sizeof(double), cudaMemcpyDeviceToHost, fh_stream()));
    FH_CUDA(cudaStreamSynchronize(fh_stream()));
    *out = op->hist_pinned[0];
    return FH_OK;
}

extern "C" int fh_cg(fh_ga* op, const double* B, double* x, double tol, int64_t maxiter, double* vecs,
                     int64_t* kit_host, double* norm_res_host, double* hist_host, int64_t hist_cap) {
    FH_REQUIRE(op && B && x && vecs && kit_host && norm_res_host, "fh_cg: null argument");
    const int64_t n = (int64_t)op->D * op->plan->nreal;
    const double inv = 1.0 / (double)op->plan->nreal;
    double* r = vecs;
    double* p = vecs + n;
    double* Ap = vecs + 2 * n;
    const unsigned g = ga_grid(n);
    cudaStream_t s = fh_stream();
    int rc;
    // Ax = Afun(x0); R = B - Ax; P = R; rr = scal(R, R)
    if ((rc = fh_ga_apply(op, x, Ap))) return rc;
    k_cg_init<<<g, GA_NT, 0, s>>>(n, B, Ap, r, p, op->part);
    FH_LAUNCH_CHECK();
    k_cg_scal<<<1, GA_NT, 0, s>>>((int)g, op->part, op->scal, inv, 0);
    FH_LAUNCH_CHECK();
    double norm_res;
    if ((rc = read_norm(op, &norm_res))) return rc;
    int64_t kit = 0;
    if (hist_host && hist_cap > 0) hist_host[0] = norm_res;
    while (norm_res > tol && kit < maxiter) {
        ++kit;
        if ((rc = fh_ga_apply(op, p, Ap))) return rc;
        k_dot_part<<<g, GA_NT, 0, s>>>(n, p, Ap, op->part);
        FH_LAUNCH_CHECK();
        k_cg_scal<<<1, GA_NT, 0, s>>>((int)g, op->part, op->scal, inv, 1);
        FH_LAUNCH_CHECK();
        k_cg_update<<<g, GA_NT, 0, s>>>(n, x, r, p, Ap, op->scal, op->part);
        FH_LAUNCH_CHECK();
        k_cg_scal<<<1, GA_NT, 0, s>>>((int)g, op->part, op->scal, inv, 2);
        FH_LAUNCH_CHECK();
        k_cg_pupdate<<<g, GA_NT, 0, s>>>(n, p, r, op->scal);
        FH_LAUNCH_CHECK();
        if ((rc = read_norm(op, &norm_res))) return rc;
        if (hist_host && kit < hist_cap) hist_host[kit] = norm_res;
    }
    *kit_host = kit;
    *norm_res_host = (kit == 0) ? 0.0 : norm_res;  // solver.py:137-138
    return FH_OK;
}

extern "C" int fh_richardson(fh_ga* op, const double* B, double* x, double alpha, double tol, int64_t maxiter,
                             double* vecs, int64_t* kit_host, double* norm_res_host) {
    FH_REQUIRE(op && B && x && vecs && kit_host && norm_res_host, "fh_richardson: null argument");
    const int64_t n = (int64_t)op->D * op->plan->nreal;
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
        if ((rc = fh_ga_apply(op, x, Ax))) return rc;
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
