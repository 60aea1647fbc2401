// fh_ga.cuh — the fused-operator object behind the opaque `fh_ga*` of the C ABI (include/ffthom_b200.h), shared by
// the translation units that implement its stages (fh_fused.cu: single-GPU pipeline and x-plane-chunked slab
// pipeline; fh_slab2.cu: k2-block slab pipeline).
#pragma once
#include "fh_plan.cuh"
#include "fh_green.cuh"
#include "fh_types.cuh"

#define GA_NT 256
#define GA_MAXPART 524288

struct fh_ga {
    const fh_plan* plan;
    int D;
    const double* A;
    int a_layout;
    int a_mode;            // how S1 reads the coefficients: 0 full, 1 symmetric (upper triangle), 2 phase table
    unsigned char* phase;  // [prod(N)] phase index per voxel (a_mode 2), owned by the operator
    double* lut;           // [nphase][D][D]
    Lut2C lutc;            // host copy of the table when nphase <= 2 (passed by value to S1)
    int nphase;
    GreenDesc g;
    int pitch;       // padded spectrum row length (complex elements)
    int64_t nrows;   // rows of the local real fields: prod(N[:-1]), or n0_local*N1 for a slab
    int64_t nloc;    // local voxels per component = nrows * N_last
    int n0l, n1l;    // slab decomposition (3-D): local planes of axis 0 (real space) / axis 1 (axis-0 pass)
    cplx* specT;     // [D][N0][n1l][pitch] y-slab spectrum (== spec when not decomposed)
    int64_t nspecp;  // nrows * pitch
    double* work;
    double* sigma;  // [D*nreal] (generic last-axis path only)
    cplx* spec;     // [D][nrows][pitch]
    // configuration
    bool fast_last, fast_mid1, fast_mid0;
    bool rt_ok[3];   // run-time-length in-place kernels usable on axis a (any N = up to 3 supported radices)
    bool odd_ax[3];  // compile-time odd-length kernels (fh_odd.cu: 255 = 15 x 17) serve axis a; they take precedence over rt
    RtPlan rt[3];
    int mid_T, trw, mid_pipe;
    int trw_s1;          // rows per CTA of S1 when it differs from trw (0: same)
    int chunk_cols;      // L2 blocking of S2-S3-S4: columns of the spectrum rows per chunk (0 = off)
    int cur_col0, cur_ncols;  // chunk the next S3 launch works on (0,0 = whole rows)
    // device scalars / partial sums of the Krylov loops
    double* scal;  // [16]: rr, pAp, alpha, beta, norm_res
    double* part;  // [GA_MAXPART]
    double* pinned;
    // CG state (fh_cg_begin / fh_cg_steps)
    int64_t kit;
    int have_beta;
    double* xacc;    // non-null inside fh_cg_steps: S1 applies the deferred x += alpha p while it has p in registers
    int last_npart;  // partial sums left in `part` by the last stage-5 launch
    // row range of the next S1 / S5 launch (chunked slab pipeline); row_cnt = 0 means all rows
    int64_t row_beg, row_cnt;
    // zero-copy slab exchange (fh_ga_slab_direct): chunk-major blocks [J][G][D][n0c][n1l][pitch]
    int sd_world, sd_nchunk, sd_n0c;
    cplx* sd_bufA;      // S2 output / S4 input (send buffer forward, receive buffer backward)
    cplx* sd_bufB;      // S3 in place (receive buffer forward, send buffer backward)
    int64_t* sd_off1;   // [N1] row offsets of the axis-1 passes inside one chunk of bufA
    int64_t* sd_off0;   // [N0] row offsets of the axis-0 pass inside bufB
    int64_t sd_cs0;     // component stride of the axis-0 pass
    int sd_peer;        // 1: the axis-0 pass reads/writes the peers' x-slab spectra directly (fh_ga_slab_peer)
    // push-mode slab exchange (fh_slab2.cu): S2 stores into the peers' y-slab spectra, S3 into their x-slab spectra
    int sp_world;       // 0: off
    int64_t* sp_off1;   // device [N1]: S2 output row k1 -> element offset from this rank's specT (peer-mapped memory)
    int64_t* sp_off0;   // device [N0]: S3 output row i0 -> element offset from this rank's spec
    // k2-block exchange pipeline (fh_slab2.cu): exchange buffers [block][G][D][n0l][n1l][w_block]
    int kb_world, kb_nblk;
    int kb_col0[16], kb_w[16];   // first spectrum column and width (multiples of 8) of every block
    int64_t kb_base[16];         // element offset of block b inside bufA / bufB
    int64_t* kb_off;             // device: per block [N1] axis-1 row offsets, then per block [N0] axis-0 row offsets
    cplx* kb_bufA;               // S2 output / S4 input of this rank (x-slab side)
    cplx* kb_bufB;               // S3 in place (y-slab side)
};

// internals of fh_fused.cu used by fh_slab2.cu
int fh_ga_stage_local(fh_ga* op, int stage, double* x, const double* r, int pupdate, double* y, int dot, int* npart);
int fh_launch_c2c_map(int N, const cplx* tw, const cplx* in, cplx* out, const LineMap& mi, const LineMap& mo, int64_t panels,
                      int pitch, bool inv);

