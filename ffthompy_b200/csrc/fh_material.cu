// fh_material.cu — device kernels behind ffthompy_b200/materials.py (`Material`: ffthompy/materials.py:54-425):
// characteristic functions of the inclusions at the nodal points, phase combination, separable spectral weights and
// the periodic extension / truncation of centred spectra.  Plain HBM-bound elementwise kernels (set-up, not the
// per-iteration path).
#include "fh_common.cuh"
#include "../../include/ffthom_b200.h"

#define FM_NT 256
static inline unsigned fm_grid(int64_t n) {
    int64_t b = (n + FM_NT - 1) / FM_NT;
    const int64_t cap = (int64_t)fh_num_sms() * 16;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

struct TopoDesc {
    int dim, ninc;
    int N[3];
    int kind[16];        // 0 cube, 1 ball, 2 pyramid, 3 otherwise, 4 all
    double Y[3];
    double pos[16][3], par[16][3];
};

// topologies at the nodes x_d[i_d] (ffthompy/materials.py:216-308): for every periodic image Ym = Y * {-1,0,1}^d
//   cube:    prod_d [ (x - pos + Ym) > -par/2 ] [ (x - pos + Ym) <= par/2 ]
//   ball:    sqrt(sum_d (x - pos - Ym)^2) < par/2
//   pyramid: prod_d max(1 - |x - pos + 2 Ym| / (par/2), 0)
// summed over the images; 'otherwise' = 1 - (all listed before it), 'all' = 1.  Same expressions, same order of
// operations as the reference (no FMA contraction), so the values are bit-identical.
__global__ void __launch_bounds__(FM_NT) k_topologies(TopoDesc d, int64_t n, const double* __restrict__ coords,
                                                       double* __restrict__ out, int* __restrict__ negative) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int nimg = (d.dim == 3) ? 27 : (d.dim == 2 ? 9 : 3);
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        double x[3] = {0.0, 0.0, 0.0};
        {
            int64_t r = v;
            int off = 0;
            int idx[3] = {0, 0, 0};
            for (int a = d.dim - 1; a >= 0; --a) {
                idx[a] = (int)(r % d.N[a]);
                r /= d.N[a];
            }
            for (int a = 0; a < d.dim; ++a) {
                x[a] = coords[off + idx[a]];
                off += d.N[a];
            }
        }
        double rest = 1.0;
        for (int ii = 0; ii < d.ninc; ++ii) {
            double topo = 0.0;
            const int kind = d.kind[ii];
            if (kind == 4) {
                topo = 1.0;
            } else if (kind == 3) {
                topo = rest;
                if (topo < 0.0) *negative = 1;
            } else {
                for (int img = 0; img < nimg; ++img) {
                    // image order of the reference's Yiter: axis 0 slowest, last axis fastest (materials.py:243-247)
                    int c[3] = {0, 0, 0};
                    {
                        int q = img;
                        for (int a = d.dim - 1; a >= 0; --a) {
                            c[a] = q % 3 - 1;
                            q /= 3;
                        }
                    }
                    double loc = 1.0, norm2 = 0.0;
                    for (int a = 0; a < d.dim; ++a) {
                        const double Ym = __dmul_rn(d.Y[a], (double)c[a]);
                        if (kind == 0) {
                            const double t = __dadd_rn(__dsub_rn(x[a], d.pos[ii][a]), Ym);
                            const double h = d.par[ii][a] / 2;
                            loc = __dmul_rn(loc, (t > -h) ? 1.0 : 0.0);
                            loc = __dmul_rn(loc, (t <= h) ? 1.0 : 0.0);
                        } else if (kind == 1) {
                            const double t = __dsub_rn(__dsub_rn(x[a], d.pos[ii][a]), Ym);
                            norm2 = __dadd_rn(norm2, __dmul_rn(t, t));
                        } else {
                            const double t = __dadd_rn(__dsub_rn(x[a], d.pos[ii][a]), __dmul_rn(2.0, Ym));
                            const double h = d.par[ii][a] / 2.;
                            const double tri = fmax(__dsub_rn(1.0, fabs(t) / h), 0.0);
                            loc = __dmul_rn(loc, tri);
                        }
                    }
                    if (kind == 1) loc = (sqrt(norm2) < d.par[ii][0] / 2) ? 1.0 : 0.0;
                    topo = __dadd_rn(topo, loc);
                }
            }
            if (kind != 3 && kind != 4) rest = __dsub_rn(rest, topo);
            else if (kind == 4) rest = __dsub_rn(rest, topo);
            out[(size_t)ii * n + v] = topo;
        }
    }
}

extern "C" int fh_topologies(int dim, const int64_t* N, const double* coords, const double* Y_host, int ninc,
                             const int* kinds_host, const double* pos_host, const double* par_host, double* out,
                             int* overlap_host) {
    FH_REQUIRE(dim >= 1 && dim <= 3 && N && coords && Y_host && kinds_host && pos_host && par_host && out && overlap_host,
               "fh_topologies: bad argument");
    FH_REQUIRE(ninc >= 1 && ninc <= 16, "fh_topologies: 1..16 inclusions (got %d)", ninc);
    TopoDesc d;
    d.dim = dim;
    d.ninc = ninc;
    int64_t n = 1;
    for (int a = 0; a < 3; ++a) {
        d.N[a] = a < dim ? (int)N[a] : 1;
        d.Y[a] = Y_host[a];
        n *= d.N[a];
    }
    for (int i = 0; i < ninc; ++i) {
        d.kind[i] = kinds_host[i];
        for (int a = 0; a < 3; ++a) {
            d.pos[i][a] = pos_host[i * 3 + a];
            d.par[i][a] = par_host[i * 3 + a];
        }
    }
    int* flag = NULL;
    FH_CUDA(cudaMalloc((void**)&flag, sizeof(int)));
    FH_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), fh_stream()));
    k_topologies<<<fm_grid(n), FM_NT, 0, fh_stream()>>>(d, n, coords, out, flag);
    fh_count_launch();
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(overlap_host, flag, sizeof(int), cudaMemcpyDeviceToHost, fh_stream());
    if (e == cudaSuccess) e = cudaStreamSynchronize(fh_stream());
    cudaFree(flag);
    if (e != cudaSuccess) return fh_set_error(FH_ERR_CUDA, "fh_topologies: %s", cudaGetErrorString(e));
    return FH_OK;
}

// out[c][v] = sum_p coef[c][p] * chars[p][v], accumulated from zero in the order p = 0, 1, ... with separate
// multiply and add (the reference's `A_val += einsum(vals[ii], topos[ii])`, materials.py:203-204)
struct CombDesc {
    double coef[36 * 16];
};
__global__ void __launch_bounds__(FM_NT) k_combine(int ncomp, int nph, int64_t n, CombDesc cd, const double* __restrict__ chars,
                                                    double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        double t[16];
        for (int p = 0; p < nph; ++p) t[p] = chars[(size_t)p * n + v];
        for (int c = 0; c < ncomp; ++c) {
            double acc = 0.0;
            for (int p = 0; p < nph; ++p) acc = __dadd_rn(acc, __dmul_rn(cd.coef[c * nph + p], t[p]));
            out[(size_t)c * n + v] = acc;
        }
    }
}
extern "C" int fh_combine_phases(int ncomp, int nphase, int64_t n, const double* coef_host, const double* chars,
                                 double* out) {
    FH_REQUIRE(ncomp >= 1 && ncomp <= 36 && nphase >= 1 && nphase <= 16 && n > 0 && coef_host && chars && out,
               "fh_combine_phases: bad argument (ncomp %d, nphase %d)", ncomp, nphase);
    CombDesc cd;
    for (int i = 0; i < ncomp * nphase; ++i) cd.coef[i] = coef_host[i];
    k_combine<<<fm_grid(n), FM_NT, 0, fh_stream()>>>(ncomp, nphase, n, cd, chars, out);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// out[b][k] = (in ? in[b][k] : 1) * f0[k0] * f1[k1] * f2[k2]   (complex; `factors` = the per-axis vectors concatenated)
__global__ void __launch_bounds__(FM_NT) k_sep_product(int dim, int n0, int n1, int n2, int64_t n, int batch,
                                                        const cplx* __restrict__ f, const cplx* __restrict__ in,
                                                        cplx* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += stride) {
        const int i2 = (int)(v % n2);
        const int64_t r = v / n2;
        const int i1 = (int)(r % n1), i0 = (int)(r / n1);
        cplx w = make_double2(1.0, 0.0);
        // vectors are stored for the `dim` real axes; leading (padded) axes have length 1
        int off = 0;
        if (dim == 3) {
            w = f[i0];
            off = n0;
        }
        if (dim >= 2) {
            w = (dim == 3) ? cmul(w, f[off + i1]) : f[off + i1];
            off += n1;
        }
        w = (dim >= 2) ? cmul(w, f[off + i2]) : f[off + i2];
        for (int b = 0; b < batch; ++b) out[(size_t)b * n + v] = in ? cmul(in[(size_t)b * n + v], w) : w;
    }
}
extern "C" int fh_sep_product(int dim, const int64_t* N, const double* factors, int batch, const double* in, double* out) {
    FH_REQUIRE(dim >= 1 && dim <= 3 && N && factors && out && batch >= 1, "fh_sep_product: bad argument");
    int n[3] = {1, 1, 1};
    int64_t nv = 1;
    for (int a = 0; a < dim; ++a) {
        n[3 - dim + a] = (int)N[a];
        nv *= N[a];
    }
    k_sep_product<<<fm_grid(nv), FM_NT, 0, fh_stream()>>>(dim, n[0], n[1], n[2], nv, batch, (const cplx*)factors,
                                                          (const cplx*)in, (cplx*)out);
    FH_LAUNCH_CHECK();
    return FH_OK;
}

// out[b][j] = in[b][(start + j) mod P] per axis (complex): the centre block of a periodically tiled centred spectrum
// (np.tile + trigpol.decrease, ffthompy/materials.py:95-102) or a plain centred truncation (P >= M)
__global__ void __launch_bounds__(FM_NT) k_gather_periodic(int p0, int p1, int p2, int m0, int m1, int m2, int s0, int s1,
                                                            int s2, int64_t nin, int64_t nout, int batch,
                                                            const cplx* __restrict__ in, cplx* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nout; v += stride) {
        const int j2 = (int)(v % m2);
        const int64_t r = v / m2;
        const int j1 = (int)(r % m1), j0 = (int)(r / m1);
        const int i0 = (s0 + j0) % p0, i1 = (s1 + j1) % p1, i2 = (s2 + j2) % p2;
        const int64_t src = ((int64_t)i0 * p1 + i1) * p2 + i2;
        for (int b = 0; b < batch; ++b) out[(size_t)b * nout + v] = in[(size_t)b * nin + src];
    }
}
extern "C" int fh_gather_periodic(int dim, const int64_t* P, const int64_t* M, const int64_t* start, int batch,
                                  const double* in, double* out) {
    FH_REQUIRE(dim >= 1 && dim <= 3 && P && M && start && in && out && batch >= 1, "fh_gather_periodic: bad argument");
    int p[3] = {1, 1, 1}, m[3] = {1, 1, 1}, s[3] = {0, 0, 0};
    int64_t nin = 1, nout = 1;
    for (int a = 0; a < dim; ++a) {
        p[3 - dim + a] = (int)P[a];
        m[3 - dim + a] = (int)M[a];
        s[3 - dim + a] = (int)(start[a] % P[a]);
        nin *= P[a];
        nout *= M[a];
    }
    k_gather_periodic<<<fm_grid(nout), FM_NT, 0, fh_stream()>>>(p[0], p[1], p[2], m[0], m[1], m[2], s[0], s[1], s[2], nin, nout,
                                                               batch, (const cplx*)in, (cplx*)out);
    FH_LAUNCH_CHECK();
    return FH_OK;
}
