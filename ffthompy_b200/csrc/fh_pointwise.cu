// fh_pointwise.cu — HBM-streaming kernels of the operator algebra: BLAS-1 on
// fields, fp64 reductions, per-voxel DxD contractions and inverses, spectrum
// re-mapping (fft_form changes, enlarge/decrease), shifts and the Fourier
// differential operators.  Each kernel cites the reference routine it replaces.
#include "fh_plan.cuh"
#include "fh_green.cuh"
#include "../../include/ffthom_b200.h"

#define FH_NT 256

static inline unsigned grid_for(int64_t n, int per_thread = 1) {
    int64_t b = fh_ceil_div(n, (int64_t)FH_NT * per_thread);
    int64_t cap = (int64_t)fh_num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// ------------------------------------------------------------------ scratch for reductions
#define FH_RED_MAX 4096
static double* g_red_dev = NULL;   // FH_RED_MAX*64 partials
static double* g_red_host = NULL;  // pinned, 64 doubles

static int ensure_scratch() {
    if (!g_red_dev) {
        FH_CUDA(cudaMalloc((void**)&g_red_dev, sizeof(double) * (FH_RED_MAX * 64 + 64)));
        FH_CUDA(cudaMallocHost((void**)&g_red_host, sizeof(double) * 64));
    }
    return FH_OK;
}

// final deterministic reduction of `np` partials for each of `nq` quantities
// (partials laid out [q][np]); op 0 = sum, 1 = max.
__global__ void k_finalize(const double* __restrict__ part, int np, int nq, double* __restrict__ out, int op) {
    __shared__ double red[32];
    for (int q = blockIdx.x; q < nq; q += gridDim.x) {
        double acc = 0.0;
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            const double v = part[(size_t)q * np + i];
            acc = op ? fmax(acc, v) : acc + v;
        }
        acc = op ? block_max(acc, red) : block_sum(acc, red);
        if (threadIdx.x == 0) out[q] = acc;
    }
}

static int finish_reduction(int np, int nq, int op, double* result) {
    k_finalize<<<nq, FH_NT, 0, fh_stream()>>>(g_red_dev, np, nq, g_red_dev + (size_t)FH_RED_MAX * 64, op);
    FH_LAUNCH_CHECK();
    FH_CUDA(cudaMemcpyAsync(g_red_host, g_red_dev + (size_t)FH_RED_MAX * 64, sizeof(double) * nq,
                            cudaMemcpyDeviceToHost, fh_stream()));
    FH_CUDA(cudaStreamSynchronize(fh_stream()));
    for (int q = 0; q < nq; ++q) result[q] = g_red_host[q];
    return FH_OK;
}

// ------------------------------------------------------------------ BLAS-1
// out = a*x + b*y   (reference: Tensor.__add__/__sub__/__neg__/__rmul__, tensors/objects.py:194-218)
__global__ void k_axpby(int64_t n, double a, const double* __restrict__ x, double b, const double* __restrict__ y,
                        double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = y ? a * x[i] + b * y[i] : a * x[i];
}
__global__ void k_add_scalar(int64_t n, const double* __restrict__ x, double s, int step, double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = x[i] + ((i % step) == 0 ? s : 0.0);
}

extern "C" int fh_axpby(int64_t n, double a, const double* x, double b, const double* y, double* out) {
    FH_REQUIRE(n >= 0 && x && out, "fh_axpby: bad argument");
    if (n == 0) return FH_OK;
    k_axpby<<<grid_for(n, 4), FH_NT, 0, fh_stream()>>>(n, a, x, b, y, out);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// out = x + s (real arrays) ; is_complex: s is added to the real parts only
extern "C" int fh_add_scalar(int64_t n, const double* x, double s, int is_complex, double* out) {
    FH_REQUIRE(n >= 0 && x && out, "fh_add_scalar: bad argument");
    if (n == 0) return FH_OK;
    k_add_scalar<<<grid_for(n, 4), FH_NT, 0, fh_stream()>>>(n, x, s, is_complex ? 2 : 1, out);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// x[c, :] += vals[c]   (Tensor.add_mean in real space, tensors/objects.py:285-286)
__global__ void k_add_comp(int ncomp, int64_t n, double* __restrict__ x, const double* __restrict__ vals) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tot = (int64_t)ncomp * n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += stride) x[i] += vals[i / n];
}
extern "C" int fh_add_comp(int ncomp, int64_t n, double* x, const double* vals_host) {
    FH_REQUIRE(ncomp >= 0 && ncomp <= 4096 && n >= 0 && x && vals_host, "fh_add_comp: bad argument");
    if (ncomp == 0 || n == 0) return FH_OK;
    int rc;
    if ((rc = ensure_scratch())) return rc;
    // stage the per-component constants through the reduction scratch (stream-ordered)
    FH_CUDA(cudaMemcpyAsync(g_red_dev, vals_host, sizeof(double) * ncomp, cudaMemcpyHostToDevice, fh_stream()));
    FH_CUDA(cudaStreamSynchronize(fh_stream()));  // vals_host may be pageable and short-lived
    k_add_comp<<<grid_for((int64_t)ncomp * n, 4), FH_NT, 0, fh_stream()>>>(ncomp, n, x, g_red_dev);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// ------------------------------------------------------------------ reductions
// mode 0: sum x*y ; 1: sum |x| ; 2: max |x| ; 3: sum |z| (complex) ; 4: max |z| (complex)
__global__ void k_reduce(int64_t n, const double* __restrict__ x, const double* __restrict__ y, int mode,
                         double* __restrict__ part) {
    __shared__ double red[32];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    if (mode == 0) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += x[i] * y[i];
    } else if (mode == 1) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += fabs(x[i]);
    } else if (mode == 2) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc = fmax(acc, fabs(x[i]));
    } else {
        const cplx* z = (const cplx*)x;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const double a = hypot(z[i].x, z[i].y);
            acc = (mode == 3) ? acc + a : fmax(acc, a);
        }
    }
    acc = (mode == 2 || mode == 4) ? block_max(acc, red) : block_sum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}

static int reduce_call(int64_t n, const double* x, const double* y, int mode, double* result) {
    int rc;
    if ((rc = ensure_scratch())) return rc;
    if (n == 0) {
        *result = 0.0;
        return FH_OK;
    }
    unsigned g = grid_for(n, 8);
    if (g > FH_RED_MAX) g = FH_RED_MAX;
    k_reduce<<<g, FH_NT, 0, fh_stream()>>>(n, x, y, mode, g_red_dev);
    FH_LAUNCH_CHECK();
    return finish_reduction((int)g, 1, (mode == 2 || mode == 4) ? 1 : 0, result);
}

// sum_i x_i*y_i   (reference: scalar_product real branch, tensors/objects.py:635, without the 1/prod(N))
extern "C" int fh_dot(int64_t n, const double* x, const double* y, double* result) {
    FH_REQUIRE(n >= 0 && x && y && result, "fh_dot: bad argument");
    return reduce_call(n, x, y, 0, result);
}
extern "C" int fh_asum(int64_t n, const double* x, int is_complex, double* result) {
    FH_REQUIRE(n >= 0 && x && result, "fh_asum: bad argument");
    return reduce_call(n, x, NULL, is_complex ? 3 : 1, result);
}
extern "C" int fh_amax(int64_t n, const double* x, int is_complex, double* result) {
    FH_REQUIRE(n >= 0 && x && result, "fh_amax: bad argument");
    return reduce_call(n, x, NULL, is_complex ? 4 : 2, result);
}

// Weighted half-spectrum product  sum_k w(k_last) Re(y conj x), w = 1 on the
// k_last = 0 plane (and the Nyquist plane for even N_last), 2 elsewhere
// (reference: scalar_product 'r' branch, tensors/objects.py:623-631, without 1/prod(N)^2).
__global__ void k_dot_rspec(int64_t nrows, int nh, int nlast, int elem, const double* __restrict__ x,
                            const double* __restrict__ y, double* __restrict__ part) {
    __shared__ double red[32];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tot = nrows * nh;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += stride) {
        const int k = (int)(i % nh);
        const double w = (k == 0 || 2 * k == nlast) ? 1.0 : 2.0;
        double v = y[i * elem] * x[i * elem];
        if (elem == 2) v += y[i * 2 + 1] * x[i * 2 + 1];
        acc += w * v;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
extern "C" int fh_dot_rspec(const fh_plan* p, int64_t batch, int is_complex, const double* x, const double* y,
                            double* result) {
    FH_REQUIRE(p && x && y && result && batch >= 0, "fh_dot_rspec: bad argument");
    int rc;
    if ((rc = ensure_scratch())) return rc;
    const int64_t nrows = batch * (p->nspec / p->nh);
    if (nrows == 0) {
        *result = 0.0;
        return FH_OK;
    }
    unsigned g = grid_for(nrows * p->nh, 8);
    if (g > FH_RED_MAX) g = FH_RED_MAX;
    k_dot_rspec<<<g, FH_NT, 0, fh_stream()>>>(nrows, p->nh, p->N[p->dim - 1], is_complex ? 2 : 1, x, y, g_red_dev);
    FH_LAUNCH_CHECK();
    return finish_reduction((int)g, 1, 0, result);
}

// real <-> complex copies: out = complex(in, 0) or out = Re(in)
__global__ void k_convert(int64_t n, const double* __restrict__ in, int in_c, double* __restrict__ out, int out_c) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double re = in[i * (in_c ? 2 : 1)];
        const double im = in_c ? in[i * 2 + 1] : 0.0;
        if (out_c) {
            out[i * 2] = re;
            out[i * 2 + 1] = im;
        } else {
            out[i] = re;
        }
    }
}
extern "C" int fh_convert(int64_t n, const double* in, int in_complex, double* out, int out_complex) {
    FH_REQUIRE(n >= 0 && in && out, "fh_convert: bad argument");
    if (n == 0) return FH_OK;
    k_convert<<<grid_for(n, 4), FH_NT, 0, fh_stream()>>>(n, in, in_complex, out, out_complex);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// per-component sums: out[c] = sum_i x[c, i]   (Tensor.mean, tensors/objects.py:273-274)
__global__ void k_sum_comp(int64_t n, const double* __restrict__ x, int np, double* __restrict__ part) {
    __shared__ double red[32];
    const int c = blockIdx.y;
    const double* xc = x + (size_t)c * n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc += xc[i];
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) part[(size_t)c * np + blockIdx.x] = acc;
}
extern "C" int fh_sum_comp(int ncomp, int64_t n, const double* x, double* result) {
    FH_REQUIRE(ncomp >= 0 && ncomp <= 64 && n >= 0 && x && result, "fh_sum_comp: bad argument (ncomp<=64)");
    int rc;
    if ((rc = ensure_scratch())) return rc;
    if (ncomp == 0) return FH_OK;
    if (n == 0) {
        for (int c = 0; c < ncomp; ++c) result[c] = 0.0;
        return FH_OK;
    }
    unsigned g = grid_for(n, 8);
    if (g > FH_RED_MAX) g = FH_RED_MAX;
    k_sum_comp<<<dim3(g, ncomp), FH_NT, 0, fh_stream()>>>(n, x, (int)g, g_red_dev);
    FH_LAUNCH_CHECK();
    return finish_reduction((int)g, ncomp, 0, result);
}

// ------------------------------------------------------------------ small host<->device moves
extern "C" int fh_poke(double* dst, int64_t offset, const double* host_vals, int64_t count) {
    FH_REQUIRE(dst && host_vals && count >= 0, "fh_poke: bad argument");
    FH_CUDA(cudaMemcpyAsync(dst + offset, host_vals, sizeof(double) * count, cudaMemcpyHostToDevice, fh_stream()));
    FH_CUDA(cudaStreamSynchronize(fh_stream()));
    return FH_OK;
}
extern "C" int fh_peek(const double* src, int64_t offset, double* host_vals, int64_t count) {
    FH_REQUIRE(src && host_vals && count >= 0, "fh_peek: bad argument");
    FH_CUDA(cudaMemcpyAsync(host_vals, src + offset, sizeof(double) * count, cudaMemcpyDeviceToHost, fh_stream()));
    FH_CUDA(cudaStreamSynchronize(fh_stream()));
    return FH_OK;
}
extern "C" int fh_memset0(double* dst, int64_t count) {
    FH_REQUIRE(dst && count >= 0, "fh_memset0: bad argument");
    FH_CUDA(cudaMemsetAsync(dst, 0, sizeof(double) * count, fh_stream()));
    return FH_OK;
}
extern "C" int fh_copy(double* dst, const double* src, int64_t count) {
    FH_REQUIRE(dst && src && count >= 0, "fh_copy: bad argument");
    FH_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * count, cudaMemcpyDeviceToDevice, fh_stream()));
    return FH_OK;
}

// out[c, :] = sign[c] * in[perm[c], :]  — component permutations (Tensor.transpose etc.,
// tensors/objects.py:324-343; einsum index shuffles in operators.py:342)
__global__ void k_gather_comps(int64_t n, int ncomp, const int* __restrict__ perm, const double* __restrict__ in,
                               double* __restrict__ out) {
    const int c = blockIdx.y;
    const double* src = in + (size_t)perm[c] * n;
    double* dst = out + (size_t)c * n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}
extern "C" int fh_gather_comps(int64_t n, int ncomp, const int* perm_host, const double* in, double* out) {
    FH_REQUIRE(n >= 0 && ncomp >= 0 && ncomp <= 1024 && perm_host && in && out, "fh_gather_comps: bad argument");
    if (n == 0 || ncomp == 0) return FH_OK;
    int rc;
    if ((rc = ensure_scratch())) return rc;
    int* perm_dev = (int*)g_red_dev;
    FH_CUDA(cudaMemcpyAsync(perm_dev, perm_host, sizeof(int) * ncomp, cudaMemcpyHostToDevice, fh_stream()));
    FH_CUDA(cudaStreamSynchronize(fh_stream()));
    k_gather_comps<<<dim3(grid_for(n, 4), ncomp), FH_NT, 0, fh_stream()>>>(n, ncomp, perm_dev, in, out);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// ------------------------------------------------------------------ per-point contractions
// y[i, k] = sum_j A[i, j] x[j, k] at every grid point / frequency
// (reference: einsum 'ij...,j...->i...', tensors/objects.py:231-232,599-604; with K > 1 it is
// the point-wise matrix product used for P*Q; multype 42 is the same with D = d*d).
template <int D, bool AC, bool XC>
__global__ void k_mul21(int64_t n, int K, const double* __restrict__ A, const double* __restrict__ x,
                        double* __restrict__ y) {
    const int k = blockIdx.y;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        cplx xv[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
            const size_t o = ((size_t)j * K + k) * n + p;
            xv[j] = XC ? ((const cplx*)x)[o] : make_double2(x[o], 0.0);
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
            cplx acc = make_double2(0.0, 0.0);
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const size_t o = ((size_t)i * D + j) * n + p;
                if (AC) {
                    const cplx a = ((const cplx*)A)[o];
                    acc.x += a.x * xv[j].x - a.y * xv[j].y;
                    acc.y += a.x * xv[j].y + a.y * xv[j].x;
                } else {
                    const double a = A[o];
                    acc.x += a * xv[j].x;
                    if (XC) acc.y += a * xv[j].y;
                }
            }
            const size_t o = ((size_t)i * K + k) * n + p;
            if (AC || XC)
                ((cplx*)y)[o] = acc;
            else
                y[o] = acc.x;
        }
    }
}

// generic D (runtime), slower: one thread per (i, point)
template <bool AC, bool XC>
__global__ void k_mul21_gen(int D, int64_t n, int K, const double* __restrict__ A, const double* __restrict__ x,
                            double* __restrict__ y) {
    const int k = blockIdx.y;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tot = (int64_t)D * n;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < tot; t += stride) {
        const int i = (int)(t / n);
        const int64_t p = t - (int64_t)i * n;
        cplx acc = make_double2(0.0, 0.0);
        for (int j = 0; j < D; ++j) {
            const size_t oa = ((size_t)i * D + j) * n + p;
            const size_t ox = ((size_t)j * K + k) * n + p;
            const cplx a = AC ? ((const cplx*)A)[oa] : make_double2(A[oa], 0.0);
            const cplx b = XC ? ((const cplx*)x)[ox] : make_double2(x[ox], 0.0);
            acc.x += a.x * b.x - a.y * b.y;
            acc.y += a.x * b.y + a.y * b.x;
        }
        const size_t o = ((size_t)i * K + k) * n + p;
        if (AC || XC)
            ((cplx*)y)[o] = acc;
        else
            y[o] = acc.x;
    }
}

template <bool AC, bool XC>
static int mul21_dispatch(int D, int64_t n, int K, const double* A, const double* x, double* y) {
    dim3 g(grid_for(n), K);
    switch (D) {
        case 1: k_mul21<1, AC, XC><<<g, FH_NT, 0, fh_stream()>>>(n, K, A, x, y); break;
        case 2: k_mul21<2, AC, XC><<<g, FH_NT, 0, fh_stream()>>>(n, K, A, x, y); break;
        case 3: k_mul21<3, AC, XC><<<g, FH_NT, 0, fh_stream()>>>(n, K, A, x, y); break;
        case 4: k_mul21<4, AC, XC><<<g, FH_NT, 0, fh_stream()>>>(n, K, A, x, y); break;
        case 6: k_mul21<6, AC, XC><<<g, FH_NT, 0, fh_stream()>>>(n, K, A, x, y); break;
        case 9: k_mul21<9, AC, XC><<<g, FH_NT, 0, fh_stream()>>>(n, K, A, x, y); break;
        default:
            g.x = grid_for((int64_t)D * n);
            k_mul21_gen<AC, XC><<<g, FH_NT, 0, fh_stream()>>>(D, n, K, A, x, y);
            break;
    }
    FH_LAUNCH_CHECK();
    return FH_OK;
}

extern "C" int fh_mul21(int D, int64_t n, int K, const double* A, int a_complex, const double* x, int x_complex,
                        double* y) {
    FH_REQUIRE(D >= 1 && D <= 64 && n >= 0 && K >= 1 && K <= 65535 && A && x && y, "fh_mul21: bad argument");
    if (n == 0) return FH_OK;
    if (a_complex && x_complex) return mul21_dispatch<true, true>(D, n, K, A, x, y);
    if (a_complex) return mul21_dispatch<true, false>(D, n, K, A, x, y);
    if (x_complex) return mul21_dispatch<false, true>(D, n, K, A, x, y);
    return mul21_dispatch<false, false>(D, n, K, A, x, y);
}

// Hadamard product with component broadcasting:
// out[c, p] = a[(c / adiv) % ca, p] * b[(c / bdiv) % cb, p]   (einsum '...,...->...' and the 'grad' multype
// 'i...,...->i...', tensors/objects.py:235-238)
template <bool AC, bool BC>
__global__ void k_hadamard(int64_t n, int nc, int adiv, int ca, int bdiv, int cb, const double* __restrict__ a,
                           const double* __restrict__ b, double* __restrict__ out) {
    const int c = blockIdx.y;
    const size_t oa = (size_t)((c / adiv) % ca) * n, ob = (size_t)((c / bdiv) % cb) * n, oo = (size_t)c * n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        const cplx u = AC ? ((const cplx*)a)[oa + p] : make_double2(a[oa + p], 0.0);
        const cplx v = BC ? ((const cplx*)b)[ob + p] : make_double2(b[ob + p], 0.0);
        if (AC || BC)
            ((cplx*)out)[oo + p] = cmul(u, v);
        else
            out[oo + p] = u.x * v.x;
    }
}
extern "C" int fh_hadamard(int64_t n, int nc, int adiv, int ca, int bdiv, int cb, const double* a, int a_complex,
                           const double* b, int b_complex, double* out) {
    FH_REQUIRE(n >= 0 && nc >= 1 && nc <= 65535 && ca >= 1 && cb >= 1 && adiv >= 1 && bdiv >= 1 && a && b && out,
               "fh_hadamard: bad argument");
    if (n == 0) return FH_OK;
    dim3 g(grid_for(n), nc);
    if (a_complex && b_complex)
        k_hadamard<true, true><<<g, FH_NT, 0, fh_stream()>>>(n, nc, adiv, ca, bdiv, cb, a, b, out);
    else if (a_complex)
        k_hadamard<true, false><<<g, FH_NT, 0, fh_stream()>>>(n, nc, adiv, ca, bdiv, cb, a, b, out);
    else if (b_complex)
        k_hadamard<false, true><<<g, FH_NT, 0, fh_stream()>>>(n, nc, adiv, ca, bdiv, cb, a, b, out);
    else
        k_hadamard<false, false><<<g, FH_NT, 0, fh_stream()>>>(n, nc, adiv, ca, bdiv, cb, a, b, out);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// out[k, p] = sum_i a[i, p] * b[i, k, p]   ('div' multype 'i...,i...->...', tensors/objects.py:239-240)
__global__ void k_contract_first(int64_t n, int d, int K, const cplx* __restrict__ a, const cplx* __restrict__ b,
                                 cplx* __restrict__ out) {
    const int k = blockIdx.y;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        cplx acc = make_double2(0.0, 0.0);
        for (int i = 0; i < d; ++i) acc = cadd(acc, cmul(a[(size_t)i * n + p], b[((size_t)i * K + k) * n + p]));
        out[(size_t)k * n + p] = acc;
    }
}
extern "C" int fh_contract_first(int64_t n, int d, int K, const double* a, const double* b, double* out) {
    FH_REQUIRE(n >= 0 && d >= 1 && K >= 1 && K <= 65535 && a && b && out, "fh_contract_first: bad argument");
    if (n == 0) return FH_OK;
    k_contract_first<<<dim3(grid_for(n), K), FH_NT, 0, fh_stream()>>>(n, d, K, (const cplx*)a, (const cplx*)b,
                                                                     (cplx*)out);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// ------------------------------------------------------------------ per-voxel inverse
// Gauss-Jordan without pivoting, same elimination order as trigpol.get_inverse
// (trigpol.py:120-159), one voxel per thread, matrix in registers.
template <int D>
__global__ void k_inv(int64_t n, const double* __restrict__ A, double* __restrict__ Ai) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        double B[D][D], I[D][D];
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) {
                B[i][j] = A[((size_t)i * D + j) * n + p];
                I[i][j] = (i == j) ? 1.0 : 0.0;
            }
#pragma unroll
        for (int m = 0; m < D; ++m) {
            const double diag = B[m][m];
            B[m][m] = 1.0;
#pragma unroll
            for (int c = m + 1; c < D; ++c) B[m][c] = B[m][c] / diag;
#pragma unroll
            for (int c = 0; c < D; ++c) I[m][c] = I[m][c] / diag;
#pragma unroll
            for (int k = m + 1; k < D; ++k) {
                const double f = B[k][m];
#pragma unroll
                for (int l = 0; l < D; ++l) {
                    B[k][l] = B[k][l] - B[m][l] * f;
                    I[k][l] = I[k][l] - I[m][l] * f;
                }
            }
        }
#pragma unroll
        for (int m = D - 1; m >= 0; --m)
#pragma unroll
            for (int k = m - 1; k >= 0; --k) {
                const double f = B[k][m];
#pragma unroll
                for (int l = 0; l < D; ++l) {
                    B[k][l] = B[k][l] - B[m][l] * f;
                    I[k][l] = I[k][l] - I[m][l] * f;
                }
            }
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j) Ai[((size_t)i * D + j) * n + p] = I[i][j];
    }
}
extern "C" int fh_inv_dxd(int D, int64_t n, const double* A, double* Ainv) {
    FH_REQUIRE(n >= 0 && A && Ainv, "fh_inv_dxd: bad argument");
    if (n == 0) return FH_OK;
    const unsigned g = grid_for(n);
    switch (D) {
        case 1: k_inv<1><<<g, FH_NT, 0, fh_stream()>>>(n, A, Ainv); break;
        case 2: k_inv<2><<<g, FH_NT, 0, fh_stream()>>>(n, A, Ainv); break;
        case 3: k_inv<3><<<g, FH_NT, 0, fh_stream()>>>(n, A, Ainv); break;
        case 4: k_inv<4><<<g, FH_NT, 0, fh_stream()>>>(n, A, Ainv); break;
        case 6: k_inv<6><<<g, FH_NT, 0, fh_stream()>>>(n, A, Ainv); break;
        default: return fh_set_error(FH_ERR_UNSUPPORTED, "fh_inv_dxd: D=%d not supported (1,2,3,4,6)", D);
    }
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// ------------------------------------------------------------------ spectrum re-mapping
// One kernel for Tensor.set_fft_form / enlarge / decrease (tensors/objects.py:135-167,
// 428-486; trigpol.py:162-214).  Forms: 0 = FFT order, full; 1 = 'r' (FFT order, last axis
// halved, values un-normalised); 2 = 'c' (centred, full).  For every output bin the signed
// frequency k is mapped to the source grid:
//   axis with M >= N (enlarge): weight 1 if |k| < N/2, 1/2 if N even and |k| = N/2, else 0;
//   axis with M <  N (decrease): take bin k (the stored bin M/2 of an even M means -M/2).
// A source bin in the un-stored half of an 'r' spectrum is read from its Hermitian partner.
struct RemapDesc {
    int dim;
    int N[3], M[3];
    int fin, fout;
    int pure_pad;  // flags. bit 0: trigpol.enlarge semantics (no Nyquist splitting: bin -N/2 kept whole,
                   // +N/2 empty); bit 1: take the Hermitian part (X(k)+conj X(-k))/2 of a full spectrum
    double scale;
};

__device__ __forceinline__ int remap_freq(int idx, int n, int form) { return form == 2 ? idx - n / 2 : fh_freq(idx, n); }

__global__ void k_spec_remap(RemapDesc d, int64_t nout, int64_t nin, int batch, const cplx* __restrict__ in,
                             cplx* __restrict__ out) {
    const int dim = d.dim;
    int so[3], si[3];  // stored extents (out, in)
    for (int a = 0; a < 3; ++a) {
        so[a] = (a < dim) ? d.M[a] : 1;
        si[a] = (a < dim) ? d.N[a] : 1;
    }
    if (d.fout == 1) so[dim - 1] = d.M[dim - 1] / 2 + 1;
    if (d.fin == 1) si[dim - 1] = d.N[dim - 1] / 2 + 1;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nout; o += stride) {
        int io[3];
        int64_t r = o;
        for (int a = dim - 1; a >= 0; --a) {
            io[a] = (int)(r % so[a]);
            r /= so[a];
        }
        double w = d.scale;
        int k[3];
        for (int a = 0; a < dim; ++a) {
            k[a] = remap_freq(io[a], d.M[a], d.fout);
            const int N = d.N[a];
            if (d.M[a] >= N) {
                const int ak = abs(k[a]);
                if (2 * ak > N)
                    w = 0.0;
                else if (2 * ak == N && d.M[a] > N) {
                    if (d.pure_pad & 1)
                        w = (k[a] > 0) ? 0.0 : w;
                    else
                        w *= 0.5;
                }
            }
        }
        if (w == 0.0) {
            for (int b = 0; b < batch; ++b) out[(size_t)b * nout + o] = make_double2(0.0, 0.0);
            continue;
        }
        // Nyquist planes of even source axes: Tensor.enlarge (tensors/objects.py:441-453) splits them axis by axis,
        // new(-N/2) = old(-N/2)/2 and new(+N/2, k') = conj(old(-N/2, -k'))/2, where -k' flips every other axis and keeps
        // the Nyquist index of an even axis that has not been split yet.  For spectra of real fields this is the
        // symmetric halving; for iterates that are NOT Hermitian-consistent (grad of an even-grid Nyquist mode inside the
        // potential-formulation CG) only this exact order reproduces the reference.  Peel the axes in reverse order:
        bool cjn = false;
        if (!(d.pure_pad & 1)) {
            for (int a = dim - 1; a >= 0; --a) {
                const int N = d.N[a];
                if (d.M[a] > N && (N % 2 == 0) && 2 * k[a] == N) {
                    cjn = !cjn;
                    k[a] = -N / 2;
                    for (int c = 0; c < dim; ++c) {
                        if (c == a) continue;
                        if (c > a && (d.N[c] % 2 == 0) && 2 * k[c] == -d.N[c]) continue;
                        k[c] = -k[c];
                    }
                }
            }
        }
        // source storage index
        bool conj = false;
        if (d.fin == 1) {
            const int N = d.N[dim - 1];
            int il = k[dim - 1] % N;
            if (il < 0) il += N;
            if (il > N / 2) conj = true;
        }
        int64_t src = 0, src2 = 0;  // src2: bin of -k (Hermitian-part option, full-spectrum sources only)
        for (int a = 0; a < dim; ++a) {
            const int N = d.N[a];
            int kk = conj ? -k[a] : k[a];
            int ii, i2;
            if (d.fin == 2) {
                ii = kk + N / 2;  // centred storage (k = +N/2 of an even axis folds onto -N/2)
                i2 = -kk + N / 2;
                if (ii >= N) ii -= N;
                if (ii < 0) ii += N;
                if (i2 >= N) i2 -= N;
                if (i2 < 0) i2 += N;
            } else {
                ii = kk % N;
                if (ii < 0) ii += N;
                i2 = (-kk) % N;
                if (i2 < 0) i2 += N;
            }
            src = src * si[a] + ii;
            src2 = src2 * si[a] + i2;
        }
        const bool herm = (d.pure_pad & 2) && d.fin != 1;
        for (int b = 0; b < batch; ++b) {
            cplx v = in[(size_t)b * nin + src];
            if (conj) v.y = -v.y;
            if (herm) {  // (X(k) + conj X(-k)) / 2 : what ifftn(X).real sees
                const cplx u = in[(size_t)b * nin + src2];
                v = make_double2(0.5 * (v.x + u.x), 0.5 * (v.y - u.y));
            }
            if (cjn) v.y = -v.y;
            out[(size_t)b * nout + o] = make_double2(v.x * w, v.y * w);
        }
    }
}

// trigpol.enlarge (trigpol.py:162-189) is positional: the block is copied to
// [ceil((M-N)/2), ceil((M+N)/2)) of a zero array, whatever the parities of N and M.
__global__ void k_pad_centred(RemapDesc d, int64_t nout, int64_t nin, int batch, const cplx* __restrict__ in,
                              cplx* __restrict__ out) {
    const int dim = d.dim;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nout; o += stride) {
        int64_t r = o, src = 0, mul = 1;
        bool inside = true;
        for (int a = dim - 1; a >= 0; --a) {
            const int j = (int)(r % d.M[a]);
            r /= d.M[a];
            const int ibeg = (d.M[a] - d.N[a] + 1) / 2;
            const int i = j - ibeg;
            inside = inside && i >= 0 && i < d.N[a];
            src += (int64_t)i * mul;
            mul *= d.N[a];
        }
        for (int b = 0; b < batch; ++b) {
            cplx v = make_double2(0.0, 0.0);
            if (inside) {
                v = in[(size_t)b * nin + src];
                v.x *= d.scale;
                v.y *= d.scale;
            }
            out[(size_t)b * nout + o] = v;
        }
    }
}

extern "C" int fh_spec_remap(int dim, const int64_t* N, int form_in, const int64_t* M, int form_out, int64_t batch,
                             double scale, int pure_pad, const double* in, double* out) {
    FH_REQUIRE(dim >= 1 && dim <= 3 && N && M && in && out && batch >= 0, "fh_spec_remap: bad argument");
    FH_REQUIRE(form_in >= 0 && form_in <= 2 && form_out >= 0 && form_out <= 2, "fh_spec_remap: bad fft form");
    if (pure_pad & 1) {
        bool grow = true;
        for (int a = 0; a < dim; ++a) grow = grow && M[a] >= N[a];
        if (grow) {
            FH_REQUIRE(form_in == 2 && form_out == 2, "fh_spec_remap: centred padding needs the 'c' form on both sides");
            RemapDesc d;
            d.dim = dim;
            d.scale = scale;
            int64_t nout = 1, nin = 1;
            for (int a = 0; a < 3; ++a) {
                d.N[a] = a < dim ? (int)N[a] : 1;
                d.M[a] = a < dim ? (int)M[a] : 1;
                nout *= d.M[a];
                nin *= d.N[a];
            }
            if (batch == 0 || nout == 0) return FH_OK;
            k_pad_centred<<<grid_for(nout), FH_NT, 0, fh_stream()>>>(d, nout, nin, (int)batch, (const cplx*)in,
                                                                    (cplx*)out);
            FH_LAUNCH_CHECK();
            return FH_OK;
        }
    }
    RemapDesc d;
    d.dim = dim;
    d.fin = form_in;
    d.fout = form_out;
    d.scale = scale;
    d.pure_pad = pure_pad;
    int64_t nout = 1, nin = 1;
    for (int a = 0; a < 3; ++a) {
        d.N[a] = a < dim ? (int)N[a] : 1;
        d.M[a] = a < dim ? (int)M[a] : 1;
    }
    for (int a = 0; a < dim; ++a) {
        const bool last = (a == dim - 1);
        nout *= (last && form_out == 1) ? d.M[a] / 2 + 1 : d.M[a];
        nin *= (last && form_in == 1) ? d.N[a] / 2 + 1 : d.N[a];
    }
    if (batch == 0 || nout == 0) return FH_OK;
    k_spec_remap<<<grid_for(nout), FH_NT, 0, fh_stream()>>>(d, nout, nin, (int)batch, (const cplx*)in, (cplx*)out);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// Circular shift of the grid axes: out[(i + s) mod N] = in[i]; elem = doubles per element.
// (np.fft.fftshift / ifftshift in Tensor.shift, tensors/objects.py:169-186, and the doubly
// centred legacy transforms, matvecs/objects.py:784-800.)
__global__ void k_roll(int dim, int n0, int n1, int n2, int s0, int s1, int s2, int elem, int64_t nvox, int batch,
                       const double* __restrict__ in, double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nvox; o += stride) {
        int i2 = (int)(o % n2);
        int64_t r = o / n2;
        int i1 = (int)(r % n1);
        int i0 = (int)(r / n1);
        int j0 = i0 + s0, j1 = i1 + s1, j2 = i2 + s2;
        if (j0 >= n0) j0 -= n0;
        if (j1 >= n1) j1 -= n1;
        if (j2 >= n2) j2 -= n2;
        const int64_t dst = ((int64_t)j0 * n1 + j1) * n2 + j2;
        for (int b = 0; b < batch; ++b)
            for (int e = 0; e < elem; ++e)
                out[((size_t)b * nvox + dst) * elem + e] = in[((size_t)b * nvox + o) * elem + e];
    }
}
extern "C" int fh_roll(int dim, const int64_t* N, const int64_t* shift, int elem, int64_t batch, const double* in,
                       double* out) {
    FH_REQUIRE(dim >= 1 && dim <= 3 && N && shift && in && out && (elem == 1 || elem == 2), "fh_roll: bad argument");
    int n[3] = {1, 1, 1}, s[3] = {0, 0, 0};
    int64_t nvox = 1;
    for (int a = 0; a < dim; ++a) {
        n[3 - dim + a] = (int)N[a];
        int64_t sh = shift[a] % N[a];
        if (sh < 0) sh += N[a];
        s[3 - dim + a] = (int)sh;
        nvox *= N[a];
    }
    if (nvox == 0 || batch == 0) return FH_OK;
    k_roll<<<grid_for(nvox), FH_NT, 0, fh_stream()>>>(dim, n[0], n[1], n[2], s[0], s[1], s[2], elem, nvox, (int)batch, in,
                                                      out);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// ------------------------------------------------------------------ Fourier differential operators
// grad: out[c, i, k] = 2 pi i xi_i(k) X[c, k]          (operators.py:227-259)
// div : out[c, k]    = sum_i 2 pi i xi_i(k) X[c, i, k]  (operators.py:261-288)
// potential_scalar: u(k) = g_a(k) / (2 pi i xi_a), a = first axis with k_a != 0; u(0) = 0
//                                                        (operators.py:296-309)
struct FreqDesc {
    int dim;
    int N[3];
    int form;  // 0, 1 ('r'), 2 ('c')
    double Y[3];
};
__device__ __forceinline__ void freq_of(const FreqDesc& f, int64_t o, int* k) {
    int s[3];
    for (int a = 0; a < f.dim; ++a) s[a] = f.N[a];
    if (f.form == 1) s[f.dim - 1] = f.N[f.dim - 1] / 2 + 1;
    int64_t r = o;
    for (int a = f.dim - 1; a >= 0; --a) {
        const int i = (int)(r % s[a]);
        r /= s[a];
        k[a] = remap_freq(i, f.N[a], f.form);
    }
}
static int64_t freq_count(const FreqDesc& f) {
    int64_t n = 1;
    for (int a = 0; a < f.dim; ++a) n *= (a == f.dim - 1 && f.form == 1) ? f.N[a] / 2 + 1 : f.N[a];
    return n;
}
static int make_freq(FreqDesc& f, int dim, const int64_t* N, const double* Y, int form) {
    FH_REQUIRE(dim >= 1 && dim <= 3 && N && Y && form >= 0 && form <= 2, "bad grid descriptor");
    f.dim = dim;
    f.form = form;
    for (int a = 0; a < 3; ++a) {
        f.N[a] = a < dim ? (int)N[a] : 1;
        f.Y[a] = a < dim ? Y[a] : 1.0;
    }
    return FH_OK;
}

__global__ void k_grad(FreqDesc f, int64_t nf, int ncomp, const cplx* __restrict__ X, cplx* __restrict__ out) {
    const double tp = 6.283185307179586476925286766559;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nf; o += stride) {
        int k[3];
        freq_of(f, o, k);
        for (int c = 0; c < ncomp; ++c) {
            const cplx v = X[(size_t)c * nf + o];
            for (int i = 0; i < f.dim; ++i) {
                const double m = tp * ((double)k[i] / f.Y[i]);
                out[((size_t)c * f.dim + i) * nf + o] = make_double2(-m * v.y, m * v.x);
            }
        }
    }
}
__global__ void k_div(FreqDesc f, int64_t nf, int ncomp, const cplx* __restrict__ X, cplx* __restrict__ out) {
    const double tp = 6.283185307179586476925286766559;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nf; o += stride) {
        int k[3];
        freq_of(f, o, k);
        for (int c = 0; c < ncomp; ++c) {
            cplx acc = make_double2(0.0, 0.0);
            for (int i = 0; i < f.dim; ++i) {
                const double m = tp * ((double)k[i] / f.Y[i]);
                const cplx v = X[((size_t)c * f.dim + i) * nf + o];
                // un-fused multiply/add: bit-identical to hD(hG(X)) evaluated through the multiplier tensors
                acc.x = __dadd_rn(acc.x, __dmul_rn(-m, v.y));
                acc.y = __dadd_rn(acc.y, __dmul_rn(m, v.x));
            }
            out[(size_t)c * nf + o] = acc;
        }
    }
}
__global__ void k_potential(FreqDesc f, int64_t nf, int ncomp, const cplx* __restrict__ X, cplx* __restrict__ out) {
    const double tp = 6.283185307179586476925286766559;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nf; o += stride) {
        int k[3];
        freq_of(f, o, k);
        int a = -1;
        for (int i = 0; i < f.dim; ++i)
            if (k[i] != 0) {
                a = i;
                break;
            }
        for (int c = 0; c < ncomp; ++c) {
            cplx r = make_double2(0.0, 0.0);
            if (a >= 0) {
                const double m = tp * ((double)k[a] / f.Y[a]);
                const cplx v = X[((size_t)c * f.dim + a) * nf + o];
                // v / (i m) = (v.y - i v.x) / m
                r = make_double2(v.y / m, -v.x / m);
            }
            out[(size_t)c * nf + o] = r;
        }
    }
}
extern "C" int fh_grad(int dim, const int64_t* N, const double* Y, int form, int ncomp, const double* X, double* out) {
    FreqDesc f;
    int rc;
    if ((rc = make_freq(f, dim, N, Y, form))) return rc;
    FH_REQUIRE(X && out && ncomp >= 0, "fh_grad: bad argument");
    const int64_t nf = freq_count(f);
    if (nf == 0 || ncomp == 0) return FH_OK;
    k_grad<<<grid_for(nf), FH_NT, 0, fh_stream()>>>(f, nf, ncomp, (const cplx*)X, (cplx*)out);
    FH_LAUNCH_CHECK();
    return FH_OK;
}
extern "C" int fh_div(int dim, const int64_t* N, const double* Y, int form, int ncomp, const double* X, double* out) {
    FreqDesc f;
    int rc;
    if ((rc = make_freq(f, dim, N, Y, form))) return rc;
    FH_REQUIRE(X && out && ncomp >= 0, "fh_div: bad argument");
    const int64_t nf = freq_count(f);
    if (nf == 0 || ncomp == 0) return FH_OK;
    k_div<<<grid_for(nf), FH_NT, 0, fh_stream()>>>(f, nf, ncomp, (const cplx*)X, (cplx*)out);
    FH_LAUNCH_CHECK();
    return FH_OK;
}
extern "C" int fh_potential(int dim, const int64_t* N, const double* Y, int form, int ncomp, const double* X,
                            double* out) {
    FreqDesc f;
    int rc;
    if ((rc = make_freq(f, dim, N, Y, form))) return rc;
    FH_REQUIRE(X && out && ncomp >= 0, "fh_potential: bad argument");
    const int64_t nf = freq_count(f);
    if (nf == 0 || ncomp == 0) return FH_OK;
    k_potential<<<grid_for(nf), FH_NT, 0, fh_stream()>>>(f, nf, ncomp, (const cplx*)X, (cplx*)out);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// ------------------------------------------------------------------ Green multipliers, unfused
// Apply the closed form to a spectrum of D components stored in `form`
// (Tensor.__call__ of the projection tensors, tensors/objects.py:220-232).
template <int KIND, int DIM>
__global__ void k_green_apply(GreenDesc g, FreqDesc f, int64_t nf, int K, const cplx* __restrict__ X,
                              cplx* __restrict__ out) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    const int kk = blockIdx.y;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nf; o += stride) {
        int k[3];
        freq_of(f, o, k);
        cplx e[D];
#pragma unroll
        for (int c = 0; c < D; ++c) e[c] = X[((size_t)c * K + kk) * nf + o];
        green_apply<KIND, DIM>(g, k, e);
#pragma unroll
        for (int c = 0; c < D; ++c) out[((size_t)c * K + kk) * nf + o] = e[c];
    }
}

// Materialise the D x D multiplier array (real) — what the reference hands out as `.val`.
template <int KIND, int DIM>
__global__ void k_green_materialize(GreenDesc g, FreqDesc f, int64_t nf, double* __restrict__ out) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nf; o += stride) {
        int k[3];
        freq_of(f, o, k);
#pragma unroll
        for (int j = 0; j < D; ++j) {
            cplx e[D];
#pragma unroll
            for (int c = 0; c < D; ++c) e[c] = make_double2(c == j ? 1.0 : 0.0, 0.0);
            green_apply<KIND, DIM>(g, k, e);
#pragma unroll
            for (int i = 0; i < D; ++i) out[((size_t)i * D + j) * nf + o] = e[i].x;
        }
    }
}

static int fill_green(GreenDesc& g, const fh_green* in) {
    FH_REQUIRE(in, "null Green descriptor");
    FH_REQUIRE(in->kind == FH_GREEN_SCALAR || in->kind == FH_GREEN_ELASTIC, "bad Green kind %d", in->kind);
    FH_REQUIRE(in->dim == 2 || in->dim == 3, "Green operators need dim 2 or 3 (got %d)", in->dim);
    g.kind = in->kind;
    g.dim = in->dim;
    for (int a = 0; a < 3; ++a) {
        g.N[a] = a < in->dim ? (int)in->N[a] : 1;
        g.band[a] = a < in->dim ? (int)in->band[a] : 0;
        g.Y[a] = a < in->dim ? in->Y[a] : 1.0;
        g.invY[a] = 1.0 / g.Y[a];
    }
    g.c0 = in->c0;
    g.cI = in->cI;
    g.cS = in->cS;
    g.cH = in->cH;
    g.cL = in->cL;
    g.cW = in->cW;
    g.scale = in->scale;
    g.ioff1 = 0;
    return FH_OK;
}
int fh_fill_green(GreenDesc& g, const fh_green* in) { return fill_green(g, in); }

extern "C" int fh_green_apply(const fh_green* gd, int form, int K, const double* X, double* out) {
    GreenDesc g;
    int rc;
    if ((rc = fill_green(g, gd))) return rc;
    FH_REQUIRE(X && out && K >= 1 && K <= 65535, "fh_green_apply: bad argument");
    FreqDesc f;
    if ((rc = make_freq(f, gd->dim, gd->N, gd->Y, form))) return rc;
    const int64_t nf = freq_count(f);
    if (nf == 0) return FH_OK;
    dim3 grid(grid_for(nf), K);
    const cplx* Xi = (const cplx*)X;
    cplx* Xo = (cplx*)out;
    if (g.kind == FH_GREEN_SCALAR && g.dim == 2)
        k_green_apply<FH_GREEN_SCALAR, 2><<<grid, FH_NT, 0, fh_stream()>>>(g, f, nf, K, Xi, Xo);
    else if (g.kind == FH_GREEN_SCALAR)
        k_green_apply<FH_GREEN_SCALAR, 3><<<grid, FH_NT, 0, fh_stream()>>>(g, f, nf, K, Xi, Xo);
    else if (g.dim == 2)
        k_green_apply<FH_GREEN_ELASTIC, 2><<<grid, FH_NT, 0, fh_stream()>>>(g, f, nf, K, Xi, Xo);
    else
        k_green_apply<FH_GREEN_ELASTIC, 3><<<grid, FH_NT, 0, fh_stream()>>>(g, f, nf, K, Xi, Xo);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

extern "C" int fh_green_materialize(const fh_green* gd, int form, double* out) {
    GreenDesc g;
    int rc;
    if ((rc = fill_green(g, gd))) return rc;
    FH_REQUIRE(out, "fh_green_materialize: bad argument");
    FreqDesc f;
    if ((rc = make_freq(f, gd->dim, gd->N, gd->Y, form))) return rc;
    const int64_t nf = freq_count(f);
    if (nf == 0) return FH_OK;
    const unsigned grid = grid_for(nf);
    if (g.kind == FH_GREEN_SCALAR && g.dim == 2)
        k_green_materialize<FH_GREEN_SCALAR, 2><<<grid, FH_NT, 0, fh_stream()>>>(g, f, nf, out);
    else if (g.kind == FH_GREEN_SCALAR)
        k_green_materialize<FH_GREEN_SCALAR, 3><<<grid, FH_NT, 0, fh_stream()>>>(g, f, nf, out);
    else if (g.dim == 2)
        k_green_materialize<FH_GREEN_ELASTIC, 2><<<grid, FH_NT, 0, fh_stream()>>>(g, f, nf, out);
    else
        k_green_materialize<FH_GREEN_ELASTIC, 3><<<grid, FH_NT, 0, fh_stream()>>>(g, f, nf, out);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// 4th-order Green tensors of tensors/projection.py:33-70 (no Nyquist zeroing):
// kind 0: small strain  G_ijkl = -n_i n_j n_k n_l + (d_ik n_j n_l + d_il n_j n_k + d_jk n_i n_l + d_jl n_i n_k)/2
// kind 1: large deformation  G_ijkl = d_ik n_j n_l
__global__ void k_green4(FreqDesc f, int64_t nf, int kind, double* __restrict__ out) {
    const int d = f.dim;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < nf; o += stride) {
        int k[3];
        freq_of(f, o, k);
        double q[3] = {0.0, 0.0, 0.0};
        double qq = 0.0;
        for (int a = 0; a < d; ++a) {
            q[a] = (double)k[a] / f.Y[a];
            qq += q[a] * q[a];
        }
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j)
                for (int kk = 0; kk < d; ++kk)
                    for (int l = 0; l < d; ++l) {
                        double v = 0.0;
                        if (qq != 0.0) {
                            if (kind == 0)
                                v = -q[i] * q[j] * q[kk] * q[l] / (qq * qq) +
                                    0.5 *
                                        ((i == kk) * q[j] * q[l] + (i == l) * q[j] * q[kk] + (j == kk) * q[i] * q[l] +
                                         (j == l) * q[i] * q[kk]) /
                                        qq;
                            else
                                v = (i == kk) * q[j] * q[l] / qq;
                        }
                        out[((((size_t)i * d + j) * d + kk) * d + l) * nf + o] = v;
                    }
    }
}
extern "C" int fh_green4_materialize(int kind, int dim, const int64_t* N, const double* Y, int form, double* out) {
    FreqDesc f;
    int rc;
    if ((rc = make_freq(f, dim, N, Y, form))) return rc;
    FH_REQUIRE(out && (kind == 0 || kind == 1), "fh_green4_materialize: bad argument");
    const int64_t nf = freq_count(f);
    if (nf == 0) return FH_OK;
    k_green4<<<grid_for(nf), FH_NT, 0, fh_stream()>>>(f, nf, kind, out);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// ------------------------------------------------------------------ CG vector updates as stand-alone calls
// (used by the slab-decomposed loop, where the scalars are reduced across ranks by the caller)
// x += alpha p ; r -= alpha Ap ; *rr_local = sum r.r over this rank's entries   (solver.py:127-129)
__global__ void k_xr_update(int64_t n, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ p,
                            const double* __restrict__ Ap, double alpha, double* __restrict__ part) {
    __shared__ double red[32];
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
extern "C" int fh_cg_xr_update(int64_t n, double* x, double* r, const double* p, const double* Ap, double alpha,
                               double* rr_local_host) {
    FH_REQUIRE(n >= 0 && x && r && p && Ap && rr_local_host, "fh_cg_xr_update: bad argument");
    int rc;
    if ((rc = ensure_scratch())) return rc;
    if (n == 0) {
        *rr_local_host = 0.0;
        return FH_OK;
    }
    unsigned g = grid_for(n, 8);
    if (g > FH_RED_MAX) g = FH_RED_MAX;
    k_xr_update<<<g, FH_NT, 0, fh_stream()>>>(n, x, r, p, Ap, alpha, g_red_dev);
    FH_LAUNCH_CHECK();
    return finish_reduction((int)g, 1, 0, rr_local_host);
}
// ------------------------------------------------------------------ homogenised matrix in one pass (K10)
// AH[s][s2] = sum_v (A e_s)(v) . e_s2(v)   for the NS minimisers e_s (D components each): the coefficient array is read
// ONCE per voxel for all NS*NS entries (ffthompy/postprocess.py:53-70 evaluates Afun(sol[ii]) * sol[jj] for every
// pair, i.e. NS*NS matrix-vector passes over A).  Per-CTA partial sums [NS*NS][gridDim], fixed-order final reduction.
struct SolPtrs {
    const double* p[6];
};
template <int D, int NS>
__global__ void __launch_bounds__(FH_NT) k_assemble_AH(int64_t n, const double* __restrict__ A, SolPtrs sol,
                                                       double* __restrict__ part) {
    __shared__ double red[32];
    double acc[NS][NS];
#pragma unroll
    for (int a = 0; a < NS; ++a)
#pragma unroll
        for (int b = 0; b < NS; ++b) acc[a][b] = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        double e[NS][D];
#pragma unroll
        for (int a = 0; a < NS; ++a)
#pragma unroll
            for (int j = 0; j < D; ++j) e[a][j] = sol.p[a][(size_t)j * n + v];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            double row[D];
#pragma unroll
            for (int j = 0; j < D; ++j) row[j] = A[((size_t)i * D + j) * n + v];
#pragma unroll
            for (int a = 0; a < NS; ++a) {
                double t = 0.0;
#pragma unroll
                for (int j = 0; j < D; ++j) t += row[j] * e[a][j];
#pragma unroll
                for (int b = 0; b < NS; ++b) acc[a][b] += t * e[b][i];
            }
        }
    }
#pragma unroll
    for (int a = 0; a < NS; ++a)
#pragma unroll
        for (int b = 0; b < NS; ++b) {
            const double r = block_sum(acc[a][b], red);
            if (threadIdx.x == 0) part[(size_t)(a * NS + b) * gridDim.x + blockIdx.x] = r;
        }
}
// A: [D][D][n] real coefficients; sols_host: nsol DEVICE pointers to real fields [D][n]; AH_host: nsol*nsol sums
// (not normalised: the caller divides by prod(N), tensors/objects.py:635).  nsol == D in {2, 3, 6}.
extern "C" int fh_assemble_AH(int D, int nsol, int64_t n, const double* A, const double* const* sols_host,
                              double* AH_host) {
    FH_REQUIRE(A && sols_host && AH_host && n > 0, "fh_assemble_AH: bad argument");
    FH_REQUIRE(nsol == D && (D == 2 || D == 3 || D == 6), "fh_assemble_AH: D = nsol in {2, 3, 6} (got %d, %d)", D, nsol);
    int rc;
    if ((rc = ensure_scratch())) return rc;
    SolPtrs sp;
    for (int a = 0; a < 6; ++a) sp.p[a] = (a < nsol) ? sols_host[a] : NULL;
    for (int a = 0; a < nsol; ++a) FH_REQUIRE(sp.p[a], "fh_assemble_AH: null solution pointer %d", a);
    unsigned g = grid_for(n, 1);
    if (g > 1024) g = 1024;
    switch (D) {
        case 2: k_assemble_AH<2, 2><<<g, FH_NT, 0, fh_stream()>>>(n, A, sp, g_red_dev); break;
        case 3: k_assemble_AH<3, 3><<<g, FH_NT, 0, fh_stream()>>>(n, A, sp, g_red_dev); break;
        case 6: k_assemble_AH<6, 6><<<g, FH_NT, 0, fh_stream()>>>(n, A, sp, g_red_dev); break;
    }
    FH_LAUNCH_CHECK();
    return finish_reduction((int)g, nsol * nsol, 0, AH_host);
}

// p = r + beta p   (solver.py:132)
__global__ void k_p_update(int64_t n, double* __restrict__ p, const double* __restrict__ r, double beta) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = r[i] + beta * p[i];
}
extern "C" int fh_cg_p_update(int64_t n, double* p, const double* r, double beta) {
    FH_REQUIRE(n >= 0 && p && r, "fh_cg_p_update: bad argument");
    if (n == 0) return FH_OK;
    k_p_update<<<grid_for(n, 4), FH_NT, 0, fh_stream()>>>(n, p, r, beta);
    FH_LAUNCH_CHECK();
    return FH_OK;
}
