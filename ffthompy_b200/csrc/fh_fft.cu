// fh_fft.cu — plan management, generic axis-pass kernels and the n-D real
// transforms of the 'r' form (numpy rfftn / irfftn semantics,
// reference: ffthompy/tensors/fft.py:39-43, tensors/operators.py:49-58).
#include "fh_plan.cuh"
#include "../../include/ffthom_b200.h"
#include <stdlib.h>
#include <string.h>
#include <vector>

// ------------------------------------------------------------------ global state
static thread_local char g_err[1024] = "";
static cudaStream_t g_stream = 0;
static int g_num_sms = 148;
static int g_smem_optin = 232448;
static int g_device = -1;

int fh_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
cudaStream_t fh_stream() { return g_stream; }
int fh_num_sms() { return g_num_sms; }
int fh_max_smem_optin() { return g_smem_optin; }
static int64_t g_launches = 0;
void fh_count_launch() { ++g_launches; }
extern "C" int64_t fh_launch_count(void) { return g_launches; }

extern "C" const char* fh_last_error(void) { return g_err; }
extern "C" int fh_version(void) { return 100; }

extern "C" int fh_init(int device) {
    FH_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    FH_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fh_set_error(FH_ERR_UNSUPPORTED, "libffthom_b200 is built for sm_100a only; device %d is sm_%d%d", device,
                            prop.major, prop.minor);
    g_num_sms = prop.multiProcessorCount;
    g_smem_optin = (int)prop.sharedMemPerBlockOptin;
    g_device = device;
    // L2 fill granularity (FH_L2FETCH = 32 | 64 | 128 bytes; unset: driver default).  The axis-0 pass reads 64-byte
    // row segments whose other half belongs to the neighbouring tile; with 128-byte fills the second reader hits L2.
    const char* fg = getenv("FH_L2FETCH");
    if (fg && atoi(fg) > 0) {
        FH_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(fg)));
    }
    return FH_OK;
}

extern "C" int fh_set_stream(void* stream) {
    g_stream = (cudaStream_t)stream;
    return FH_OK;
}

extern "C" int fh_sync(void) {
    FH_CUDA(cudaStreamSynchronize(g_stream));
    return FH_OK;
}

extern "C" int fh_device_info(int* num_sms, int* smem_optin, int* l2_bytes) {
    int dev = 0;
    FH_CUDA(cudaGetDevice(&dev));
    int l2 = 0;
    FH_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
    if (num_sms) *num_sms = g_num_sms;
    if (smem_optin) *smem_optin = g_smem_optin;
    if (l2_bytes) *l2_bytes = l2;
    return FH_OK;
}

// ------------------------------------------------------------------ plan
static void factorize(int n, AxisDesc& ax) {
    ax.n = n;
    ax.nfac = 0;
    int e2 = 0;
    int m = n;
    while (m % 2 == 0) {
        m /= 2;
        ++e2;
    }
    // powers of two: as few passes as possible with radix <= 16, evenly split
    if (e2 > 0) {
        int np = (e2 + 3) / 4;
        int base = e2 / np, extra = e2 % np;
        for (int i = 0; i < np; ++i) ax.fac[ax.nfac++] = 1 << (base + (i < extra ? 1 : 0));
    }
    // odd part: prime factors, with pairs of small ones merged into the composite radices 9 and 15
    // (one register pass instead of two)
    int odd[FH_MAX_FAC + 8], nodd = 0;
    for (int p = 3; (int64_t)p * p <= m; p += 2)
        while (m % p == 0 && nodd < FH_MAX_FAC + 8) {
            odd[nodd++] = p;
            m /= p;
        }
    if (m > 1) odd[nodd++] = m;
    bool used[FH_MAX_FAC + 8] = {false};
    for (int i = 0; i < nodd; ++i) {
        if (used[i]) continue;
        int r = odd[i];
        if (r == 3) {
            for (int j = i + 1; j < nodd; ++j)
                if (!used[j] && (odd[j] == 5 || odd[j] == 3)) {
                    // prefer 3*5 = 15, else 3*3 = 9
                    int pick = -1;
                    for (int k = i + 1; k < nodd; ++k)
                        if (!used[k] && odd[k] == 5) {
                            pick = k;
                            break;
                        }
                    if (pick < 0) pick = j;
                    r *= odd[pick];
                    used[pick] = true;
                    break;
                }
        }
        used[i] = true;
        if (ax.nfac < FH_MAX_FAC) ax.fac[ax.nfac] = r;
        ax.nfac++;
    }
}

extern "C" int fh_plan_create(fh_plan** out, int dim, const int64_t* N) {
    FH_REQUIRE(out != NULL && N != NULL, "fh_plan_create: null argument");
    FH_REQUIRE(dim >= 1 && dim <= 3, "fh_plan_create: dim must be 1..3 (got %d)", dim);
    fh_plan* p = (fh_plan*)calloc(1, sizeof(fh_plan));
    if (!p) return fh_set_error(FH_ERR_ALLOC, "fh_plan_create: out of host memory");
    p->dim = dim;
    p->nreal = 1;
    for (int a = 0; a < dim; ++a) {
        if (N[a] < 1 || N[a] > (1 << 20)) {
            free(p);
            return fh_set_error(FH_ERR_ARG, "fh_plan_create: bad grid size N[%d]=%lld", a, (long long)N[a]);
        }
        p->N[a] = (int)N[a];
        p->nreal *= N[a];
    }
    for (int a = dim; a < 3; ++a) p->N[a] = 1;
    p->nh = p->N[dim - 1] / 2 + 1;
    p->nspec = p->nreal / p->N[dim - 1] * p->nh;
    for (int a = 0; a < dim; ++a) {
        const int n = p->N[a];
        factorize(n, p->ax[a]);
        if (p->ax[a].nfac > FH_MAX_FAC) {
            free(p);
            return fh_set_error(FH_ERR_UNSUPPORTED, "fh_plan_create: too many radix factors for n=%d", n);
        }
        std::vector<cplx> tw(n);
        for (int m = 0; m < n; ++m) {
            // exact octant reduction is unnecessary with 80-bit long double: |err| < 1e-19
            const long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / (long double)n;
            tw[m] = make_double2((double)cosl(ang), (double)(-sinl(ang)));
        }
        cudaError_t e = cudaMalloc((void**)&p->tw_dev[a], sizeof(cplx) * n);
        if (e == cudaSuccess) e = cudaMemcpy(p->tw_dev[a], tw.data(), sizeof(cplx) * n, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            for (int b = 0; b < a; ++b) cudaFree(p->tw_dev[b]);
            free(p);
            return fh_set_error(FH_ERR_CUDA, "fh_plan_create: twiddle upload failed: %s", cudaGetErrorString(e));
        }
        p->ax[a].tw = p->tw_dev[a];
    }
    *out = p;
    return FH_OK;
}

extern "C" int fh_plan_destroy(fh_plan* p) {
    if (!p) return FH_OK;
    for (int a = 0; a < p->dim; ++a)
        if (p->tw_dev[a]) cudaFree(p->tw_dev[a]);
    free(p);
    return FH_OK;
}

extern "C" int fh_plan_factors(const fh_plan* p, int axis, int* nfac, int* fac) {
    FH_REQUIRE(p && axis >= 0 && axis < p->dim, "fh_plan_factors: bad argument");
    *nfac = p->ax[axis].nfac;
    for (int i = 0; i < p->ax[axis].nfac; ++i) fac[i] = p->ax[axis].fac[i];
    return FH_OK;
}

// ------------------------------------------------------------------ kernels
// Strided complex transform of the middle axis of a [outer][n][inner] array.
// One CTA handles T consecutive `inner` indices (T*16 contiguous bytes per row).
template <bool INV>
__global__ void __launch_bounds__(256) k_c2c_strided(const cplx* __restrict__ in, cplx* __restrict__ out, AxisDesc ax,
                                                     int64_t inner, int T, int ld, int ntiles, double scale) {
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    double* sm = reinterpret_cast<double*>(fh_smem_raw);
    const int n = ax.n;
    double* b0re = sm;
    double* b0im = sm + (size_t)n * ld;
    double* b1re = sm + (size_t)2 * n * ld;
    double* b1im = sm + (size_t)3 * n * ld;
    const int64_t o = blockIdx.x / ntiles;
    const int tile = blockIdx.x % ntiles;
    const int64_t i0 = (int64_t)tile * T;
    const int nl = (int)min((int64_t)T, inner - i0);
    const int64_t base = o * n * inner + i0;
    // 4 independent 16-byte loads in flight per thread
    for (int idx0 = threadIdx.x; idx0 < n * nl; idx0 += 4 * blockDim.x) {
        cplx c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * blockDim.x;
            if (idx < n * nl) {
                const int row = idx / nl, t = idx - row * nl;
                c[u] = in[base + (int64_t)row * inner + t];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * blockDim.x;
            if (idx < n * nl) {
                const int row = idx / nl, t = idx - row * nl;
                b0re[row * ld + t] = c[u].x;
                b0im[row * ld + t] = c[u].y;
            }
        }
    }
    __syncthreads();
    const int cur = fft_smem<INV>(b0re, b0im, b1re, b1im, ax, nl, ld);
    const double* rre = cur ? b1re : b0re;
    const double* rim = cur ? b1im : b0im;
    for (int idx = threadIdx.x; idx < n * nl; idx += blockDim.x) {
        const int row = idx / nl, t = idx - row * nl;
        out[base + (int64_t)row * inner + t] = make_double2(rre[row * ld + t] * scale, rim[row * ld + t] * scale);
    }
}

// Real -> half-spectrum along the contiguous last axis, two real lines per
// complex transform (valid for even and odd n alike).
__global__ void __launch_bounds__(256) k_r2c_last(const double* __restrict__ x, cplx* __restrict__ X, AxisDesc ax,
                                                  int64_t nlines, int nh, int pitch, int LP, int ld) {
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    double* sm = reinterpret_cast<double*>(fh_smem_raw);
    const int n = ax.n;
    double* b0re = sm;
    double* b0im = sm + (size_t)n * ld;
    double* b1re = sm + (size_t)2 * n * ld;
    double* b1im = sm + (size_t)3 * n * ld;
    const int64_t line0 = (int64_t)blockIdx.x * 2 * LP;
    const int nll = (int)min((int64_t)2 * LP, nlines - line0);
    const int npairs = (nll + 1) >> 1;
    for (int idx0 = threadIdx.x; idx0 < 2 * npairs * n; idx0 += 4 * blockDim.x) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * blockDim.x;
            v[u] = 0.0;
            if (idx < 2 * npairs * n) {
                const int l = idx / n, i = idx - l * n;
                if (l < nll) v[u] = x[(line0 + l) * n + i];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = idx0 + u * blockDim.x;
            if (idx < 2 * npairs * n) {
                const int l = idx / n, i = idx - l * n;
                if (l & 1)
                    b0im[i * ld + (l >> 1)] = v[u];
                else
                    b0re[i * ld + (l >> 1)] = v[u];
            }
        }
    }
    __syncthreads();
    const int cur = fft_smem<false>(b0re, b0im, b1re, b1im, ax, npairs, ld);
    const double* zre = cur ? b1re : b0re;
    const double* zim = cur ? b1im : b0im;
    for (int idx = threadIdx.x; idx < nll * nh; idx += blockDim.x) {
        const int l = idx / nh, k = idx - l * nh;
        const int pr = l >> 1;
        const int km = (k == 0) ? 0 : n - k;
        const double ax_ = zre[k * ld + pr], ay_ = zim[k * ld + pr];
        const double bx_ = zre[km * ld + pr], by_ = zim[km * ld + pr];
        cplx r;
        if (l & 1)
            r = make_double2(0.5 * (ay_ + by_), -0.5 * (ax_ - bx_));
        else
            r = make_double2(0.5 * (ax_ + bx_), 0.5 * (ay_ - by_));
        X[(line0 + l) * pitch + k] = r;
    }
}

// Half-spectrum -> real along the contiguous last axis (numpy irfft semantics:
// the imaginary parts of the DC and, for even n, Nyquist bins are ignored).
__global__ void __launch_bounds__(256) k_c2r_last(const cplx* __restrict__ X, double* __restrict__ x, AxisDesc ax,
                                                  int64_t nlines, int nh, int pitch, int LP, int ld, double scale) {
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    double* sm = reinterpret_cast<double*>(fh_smem_raw);
    const int n = ax.n;
    double* b0re = sm;
    double* b0im = sm + (size_t)n * ld;
    double* b1re = sm + (size_t)2 * n * ld;
    double* b1im = sm + (size_t)3 * n * ld;
    const int64_t line0 = (int64_t)blockIdx.x * 2 * LP;
    const int nll = (int)min((int64_t)2 * LP, nlines - line0);
    const int npairs = (nll + 1) >> 1;
    for (int idx0 = threadIdx.x; idx0 < npairs * nh; idx0 += 2 * blockDim.x)
      for (int u = 0; u < 2; ++u) {
        const int idx = idx0 + u * blockDim.x;
        if (idx >= npairs * nh) break;
        const int pr = idx / nh, k = idx - pr * nh;
        const int64_t la = line0 + 2 * pr;
        cplx a = X[la * pitch + k];
        cplx b = (2 * pr + 1 < nll) ? X[(la + 1) * pitch + k] : make_double2(0.0, 0.0);
        if (k == 0 || 2 * k == n) {
            a.y = 0.0;
            b.y = 0.0;
        }
        b0re[k * ld + pr] = a.x - b.y;
        b0im[k * ld + pr] = a.y + b.x;
        if (k > 0 && 2 * k != n) {
            b0re[(n - k) * ld + pr] = a.x + b.y;
            b0im[(n - k) * ld + pr] = -a.y + b.x;
        }
    }
    __syncthreads();
    const int cur = fft_smem<true>(b0re, b0im, b1re, b1im, ax, npairs, ld);
    const double* zre = cur ? b1re : b0re;
    const double* zim = cur ? b1im : b0im;
    for (int idx = threadIdx.x; idx < nll * n; idx += blockDim.x) {
        const int l = idx / n, i = idx - l * n;
        const double v = (l & 1) ? zim[i * ld + (l >> 1)] : zre[i * ld + (l >> 1)];
        x[(line0 + l) * n + i] = v * scale;
    }
}

// ------------------------------------------------------------------ launch configuration
static int pick_lines(int n, int want) {
    // largest T in {want, want/2, ...} whose two SoA ping-pong buffers leave >= 2 CTAs/SM,
    // falling back to whatever fits at all.
    const size_t pref = 113 * 1024;
    int T = want;
    while (T > 4 && fft_smem_bytes(n, fft_ld(T)) > pref) T >>= 1;
    while (T > 1 && fft_smem_bytes(n, fft_ld(T)) > (size_t)fh_max_smem_optin()) T >>= 1;
    return T;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > (size_t)fh_max_smem_optin())
        return fh_set_error(FH_ERR_UNSUPPORTED, "transform length needs %zu B of shared memory (> %d)", bytes,
                            fh_max_smem_optin());
    if (bytes > 48 * 1024) FH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return FH_OK;
}

int fh_launch_c2c_strided(const AxisDesc& ax, const cplx* in, cplx* out, int64_t outer, int64_t inner, bool inverse,
                          double scale) {
    if (outer <= 0 || inner <= 0) return FH_OK;
    int T = pick_lines(ax.n, 8);
    int64_t ntiles = fh_ceil_div(inner, T);
    T = (int)fh_ceil_div(inner, ntiles);
    const int ld = fft_ld(T);
    const size_t smem = fft_smem_bytes(ax.n, ld);
    const int64_t nblk = outer * ntiles;
    FH_REQUIRE(nblk < 2147483647LL, "c2c_strided: grid too large");
    int rc;
    if (inverse) {
        if ((rc = set_smem(k_c2c_strided<true>, smem))) return rc;
        k_c2c_strided<true><<<(unsigned)nblk, 256, smem, fh_stream()>>>(in, out, ax, inner, T, ld, (int)ntiles, scale);
    } else {
        if ((rc = set_smem(k_c2c_strided<false>, smem))) return rc;
        k_c2c_strided<false><<<(unsigned)nblk, 256, smem, fh_stream()>>>(in, out, ax, inner, T, ld, (int)ntiles, scale);
    }
    FH_LAUNCH_CHECK();
    return FH_OK;
}

int fh_launch_r2c_last(const fh_plan* p, const double* x, cplx* X, int64_t nlines, int pitch) {
    if (nlines <= 0) return FH_OK;
    const AxisDesc& ax = p->ax[p->dim - 1];
    const int LP = pick_lines(ax.n, 8);
    const int ld = fft_ld(LP);
    const size_t smem = fft_smem_bytes(ax.n, ld);
    const int64_t nblk = fh_ceil_div(nlines, 2 * LP);
    FH_REQUIRE(nblk < 2147483647LL, "r2c_last: grid too large");
    int rc;
    if ((rc = set_smem(k_r2c_last, smem))) return rc;
    k_r2c_last<<<(unsigned)nblk, 256, smem, fh_stream()>>>(x, X, ax, nlines, p->nh, pitch, LP, ld);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

int fh_launch_c2r_last(const fh_plan* p, const cplx* X, double* x, int64_t nlines, int pitch, double scale) {
    if (nlines <= 0) return FH_OK;
    const AxisDesc& ax = p->ax[p->dim - 1];
    const int LP = pick_lines(ax.n, 8);
    const int ld = fft_ld(LP);
    const size_t smem = fft_smem_bytes(ax.n, ld);
    const int64_t nblk = fh_ceil_div(nlines, 2 * LP);
    FH_REQUIRE(nblk < 2147483647LL, "c2r_last: grid too large");
    int rc;
    if ((rc = set_smem(k_c2r_last, smem))) return rc;
    k_c2r_last<<<(unsigned)nblk, 256, smem, fh_stream()>>>(X, x, ax, nlines, p->nh, pitch, LP, ld, scale);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// ------------------------------------------------------------------ n-D real transforms
// X[b, k0, k1, k2<=N2/2] = sum_x x[b, x] exp(-2 pi i k.x/N)   (no normalisation)
extern "C" int fh_rfftn(const fh_plan* p, const double* x, double* Xout, int64_t batch) {
    FH_REQUIRE(p && x && Xout && batch >= 0, "fh_rfftn: bad argument");
    cplx* X = (cplx*)Xout;
    const int d = p->dim;
    const int64_t nlines = batch * (p->nreal / p->N[d - 1]);
    int rc;
    if ((rc = fh_launch_r2c_last(p, x, X, nlines, p->nh))) return rc;
    if (d == 3) {
        if ((rc = fh_launch_c2c_strided(p->ax[1], X, X, batch * p->N[0], p->nh, false, 1.0))) return rc;
        if ((rc = fh_launch_c2c_strided(p->ax[0], X, X, batch, (int64_t)p->N[1] * p->nh, false, 1.0))) return rc;
    } else if (d == 2) {
        if ((rc = fh_launch_c2c_strided(p->ax[0], X, X, batch, p->nh, false, 1.0))) return rc;
    }
    return FH_OK;
}

// x = scale * sum_k X[k] exp(+2 pi i k.x/N) with Hermitian completion along the
// last axis.  `work` (same size as X) keeps X intact; work == NULL transforms X
// in place (destroying it).  numpy.fft.irfftn corresponds to scale = 1/prod(N).
extern "C" int fh_irfftn(const fh_plan* p, const double* Xin, double* x, int64_t batch, double scale, double* work) {
    FH_REQUIRE(p && Xin && x && batch >= 0, "fh_irfftn: bad argument");
    const cplx* X = (const cplx*)Xin;
    cplx* W = work ? (cplx*)work : (cplx*)Xin;
    const int d = p->dim;
    const int64_t nlines = batch * (p->nreal / p->N[d - 1]);
    int rc;
    const cplx* src = X;
    if (d == 3) {
        if ((rc = fh_launch_c2c_strided(p->ax[0], src, W, batch, (int64_t)p->N[1] * p->nh, true, 1.0))) return rc;
        if ((rc = fh_launch_c2c_strided(p->ax[1], W, W, batch * p->N[0], p->nh, true, 1.0))) return rc;
        src = W;
    } else if (d == 2) {
        if ((rc = fh_launch_c2c_strided(p->ax[0], src, W, batch, p->nh, true, 1.0))) return rc;
        src = W;
    }
    return fh_launch_c2r_last(p, src, x, nlines, p->nh, scale);
}
