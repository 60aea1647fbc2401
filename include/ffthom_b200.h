/* ffthom_b200.h — C ABI of libffthom_b200.so: the B200 (sm_100a) implementation of
 * FFTHomPy's Fourier–Galerkin solve loop.
 *
 * The reference (vondrejc/FFTHomPy) is pure Python/NumPy and has no FFI; the entry
 * points below are what a ctypes binding for its hot path binds.
 * Each declaration cites the reference routine (path:line under the FFTHomPy tree)
 * it replaces.
 *
 * Conventions
 *  - every function returns 0 on success or a negative FH_ERR_* code; the message is
 *    available from fh_last_error() (thread-local);
 *  - all array arguments are caller-owned DEVICE pointers unless the name ends in
 *    `_host`; the library never frees them and keeps none past the call except the
 *    `A`/`work` pointers registered in an fh_ga operator;
 *  - fields are C-contiguous, component-major: shape + N  (fp64) in real space and
 *    shape + N_fft (complex128 = interleaved re,im doubles) in Fourier space, exactly
 *    the layout of `Tensor.val` (ffthompy/tensors/objects.py:92-133);
 *  - all work is enqueued on the stream given to fh_set_stream (default stream 0);
 *    only functions that return scalars to the host synchronise.
 */
#ifndef FFTHOM_B200_H
#define FFTHOM_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define FH_OK 0
#define FH_ERR_CUDA -1
#define FH_ERR_ARG -2
#define FH_ERR_UNSUPPORTED -3
#define FH_ERR_ALLOC -4

/* fft_form codes (ffthompy/tensors/objects.py:47-66): 0 -> 0, 'r' -> 1, 'c' -> 2 */
#define FH_FORM_0 0
#define FH_FORM_R 1
#define FH_FORM_C 2

typedef struct fh_plan fh_plan; /* FFT plan of one grid N (twiddles, radix schedule) */
typedef struct fh_ga fh_ga;     /* fused operator  x -> F^-1 G^ F (A x)               */

/* Closed-form Green/projection multiplier (ffthompy/projections.py:9-112,114-267):
 *   xi = 0               -> c0 * I
 *   xi != 0, |k_i|<=band -> scale * (cI*I + cS*S + cH*v v^T + cL*Lambda + cW*(W+W^T))
 *   otherwise            -> 0
 * kind 0 (scalar problems, D = dim): only cI, cH (n (x) n) are used.
 * kind 1 (elasticity, Mandel, D = dim(dim+1)/2).                                      */
typedef struct fh_green {
    int32_t kind;
    int32_t dim;
    int64_t N[3];
    int64_t band[3];
    double Y[3];
    double c0, cI, cS, cH, cL, cW;
    double scale;
} fh_green;

/* ---- lifecycle ------------------------------------------------------------------ */
int fh_init(int device);
int fh_set_stream(void* cuda_stream);
int fh_sync(void);
const char* fh_last_error(void);
int fh_version(void);
int fh_device_info(int* num_sms, int* smem_optin, int* l2_bytes);
/* device -> host copy of a result into fresh pageable memory (the new NumPy array `Tensor.val` hands out,
 * ffthompy/tensors/objects.py:119-121): pinned staging ring + parallel first-touch of the destination; ordered
 * after the work enqueued on the library stream, complete on return                                          */
int fh_download(void* dst_host, const void* src_device, int64_t bytes);

/* ---- FFT plan and transforms (ffthompy/tensors/fft.py:39-43; operators.py:14-58) -- */
int fh_plan_create(fh_plan** plan, int dim, const int64_t* N);
int fh_plan_destroy(fh_plan* plan);
int fh_plan_factors(const fh_plan* plan, int axis, int* nfac, int* fac);
/* numpy.fft.rfftn over the last `dim` axes, `batch` leading components, un-normalised */
int fh_rfftn(const fh_plan* plan, const double* x, double* X, int64_t batch);
/* x = scale * irfftn-sum (numpy.fft.irfftn == scale 1/prod(N)); work: spectrum-sized
 * scratch keeping X intact, or NULL to transform X in place (destroying it)           */
int fh_irfftn(const fh_plan* plan, const double* X, double* x, int64_t batch, double scale, double* work);

/* ---- field algebra (ffthompy/tensors/objects.py:194-300,606-636) ------------------ */
int fh_axpby(int64_t n, double a, const double* x, double b, const double* y /*or NULL*/, double* out);
int fh_add_scalar(int64_t n, const double* x, double s, int is_complex, double* out);
int fh_add_comp(int ncomp, int64_t n, double* x, const double* vals_host);
int fh_dot(int64_t n, const double* x, const double* y, double* result_host);
int fh_dot_rspec(const fh_plan* plan, int64_t batch, int is_complex, const double* x, const double* y,
                 double* result_host);
int fh_convert(int64_t n, const double* in, int in_complex, double* out, int out_complex);
int fh_asum(int64_t n, const double* x, int is_complex, double* result_host);
int fh_amax(int64_t n, const double* x, int is_complex, double* result_host);
int fh_sum_comp(int ncomp, int64_t n, const double* x, double* result_host);
int fh_poke(double* dst, int64_t offset, const double* vals_host, int64_t count);
int fh_peek(const double* src, int64_t offset, double* vals_host, int64_t count);
int fh_memset0(double* dst, int64_t count);
int fh_copy(double* dst, const double* src, int64_t count);
int fh_gather_comps(int64_t n, int ncomp, const int* perm_host, const double* in, double* out);

/* ---- per-point contractions (tensors/objects.py:220-245,599-604; matvecs/objects.py:471-499) */
int fh_mul21(int D, int64_t n, int K, const double* A, int a_complex, const double* x, int x_complex, double* y);
/* out[c] = a[(c/adiv) % ca] * b[(c/bdiv) % cb], c < nc  (einsum '...,...->...' and 'i...,...->i...') */
int fh_hadamard(int64_t n, int nc, int adiv, int ca, int bdiv, int cb, const double* a, int a_complex, const double* b,
                int b_complex, double* out);
int fh_contract_first(int64_t n, int d, int K, const double* a, const double* b, double* out);
/* Gauss-Jordan inverse per voxel (ffthompy/trigpol.py:120-159) */
int fh_inv_dxd(int D, int64_t n, const double* A, double* Ainv);
/* homogenised matrix in one pass over the coefficients (ffthompy/postprocess.py:53-70: AH[i][j] = Afun(sol[i]) * sol[j]
 * for every pair): A [D][D][n] real, sols_host = nsol DEVICE pointers to the minimisers [D][n] (nsol == D in {2,3,6});
 * AH_host[i*nsol+j] = sum over voxels of (A e_i).e_j, NOT divided by prod(N) */
int fh_assemble_AH(int D, int nsol, int64_t n, const double* A, const double* const* sols_host, double* AH_host);

/* ---- material coefficients (ffthompy/materials.py:54-425; set-up of the solve loop's largest input) -------------
 * fh_topologies: characteristic functions of `ninc` inclusions (kind 0 cube, 1 ball, 2 pyramid, 3 otherwise, 4 all;
 *   pos / par: ninc x 3 host arrays) at the nodes x_a[i_a] (`coords` = the per-axis coordinate vectors concatenated, device),
 *   out [ninc][prod N]; *overlap_host != 0 if an 'otherwise' phase became negative (materials.py:296-297).
 * fh_combine_phases: out[c] = sum_p coef[c][p] * chars[p]  (materials.py:203-204, :70-75).
 * fh_sep_product: out[b][k] = (in ? in[b][k] : 1) * prod_a f_a[k_a], complex (weights of materials.py:318-390).
 * fh_gather_periodic: out[b][j] = in[b][(start + j) mod P] per axis, complex (tile + decrease, materials.py:95-102). */
int fh_topologies(int dim, const int64_t* N, const double* coords, const double* Y_host, int ninc, const int* kinds_host,
                  const double* pos_host, const double* par_host, double* out, int* overlap_host);
int fh_combine_phases(int ncomp, int nphase, int64_t n, const double* coef_host, const double* chars, double* out);
int fh_sep_product(int dim, const int64_t* N, const double* factors, int batch, const double* in, double* out);
int fh_gather_periodic(int dim, const int64_t* P, const int64_t* M, const int64_t* start, int batch, const double* in,
                       double* out);

/* ---- spectra: form changes, enlarge/decrease, shifts (tensors/objects.py:135-186,428-486;
 *      trigpol.py:162-214) ----------------------------------------------------------- */
/* flags bit 0: trigpol.enlarge semantics (centred zero padding, no Nyquist splitting);
 * flags bit 1: Hermitian part (X(k)+conj X(-k))/2 of a full-spectrum source (what ifftn(X).real sees) */
int fh_spec_remap(int dim, const int64_t* N, int form_in, const int64_t* M, int form_out, int64_t batch, double scale,
                  int flags, const double* in, double* out);
int fh_roll(int dim, const int64_t* N, const int64_t* shift, int elem_doubles, int64_t batch, const double* in,
            double* out);

/* ---- Fourier differential operators (ffthompy/tensors/operators.py:227-312) ------- */
int fh_grad(int dim, const int64_t* N, const double* Y, int form, int ncomp, const double* X, double* out);
int fh_div(int dim, const int64_t* N, const double* Y, int form, int ncomp, const double* X, double* out);
int fh_potential(int dim, const int64_t* N, const double* Y, int form, int ncomp, const double* X, double* out);

/* ---- Green operators, unfused (projections.py; tensors/projection.py:33-70) ------- */
int fh_green_apply(const fh_green* g, int form, int K, const double* X, double* out);
int fh_green_materialize(const fh_green* g, int form, double* out);
int fh_green4_materialize(int kind, int dim, const int64_t* N, const double* Y, int form, double* out);

/* ---- fused operator  y = F^-1 G^ F (A x)  and Krylov loops -------------------------
 * (Operator([[Operator([[FiN, G^, FN]]), A]]), tensors/operators.py:136-144;
 *  CG / richardson, ffthompy/general/solver.py:63-139)
 * a_layout: 0 = full D*D*n array (Tensor.val of the coefficient tensor)               */
int64_t fh_ga_work_doubles(const fh_plan* plan, int D);
int fh_ga_create(fh_ga** op, const fh_plan* plan, int D, const double* A, int a_layout, const fh_green* g,
                 double* work);
/* slab-decomposed operator of one rank (3-D): local real fields [.][n0_local][N1][N2];
 * S3 runs on the transposed spectrum [D][N0][n1_local][pitch] with global axis-1 offset n1_offset;
 * the caller exchanges the spectrum between the two layouts (fh_ga_buffers) around S3 */
int64_t fh_ga_slab_work_doubles(const fh_plan* plan, int D, int n0_local, int n1_local);
int fh_ga_create_slab(fh_ga** op, const fh_plan* plan, int D, const double* A_local, int a_layout, const fh_green* g,
                      double* work, int n0_local, int n1_local, int n1_offset);
int fh_ga_buffers(const fh_ga* op, void** spec, void** specT, int* pitch);
/* (bufA / bufB below must be handed in ZERO-FILLED: the padding columns travel with the rows, and the library writes
 * nothing into peer-mapped buffers outside the stages) */
/* Zero-copy, chunked slab pipeline (SURVEY §8e "overlap the all-to-all with the local FFT passes"):
 * bufA / bufB are caller-owned exchange buffers of D*n0_local*N1*pitch complex128 each, organised as
 * nchunk contiguous blocks [world][D][n0_local/nchunk][n1_local][pitch]; the axis-1 and axis-0 kernels
 * address them directly, so each exchange is one all_to_all_single(bufB_j, bufA_j) (forward) or
 * (bufA_j, bufB_j) (backward) per chunk with no pack/unpack pass.  N0, N1 in {16,...,2048} powers of two. */
int fh_ga_slab_direct(fh_ga* op, int world, int nchunk, void* bufA, void* bufB);
/* Fused axis-0 pass + exchange over NVLink peer memory (no all-to-all, no exchange buffer): stage 3 of
 * rank `rank` loads row i0 of its k1 range from the x-slab spectrum of the rank that owns plane i0
 * (peer_spec[g] = rank g's fh_ga_buffers spectrum as mapped into this process, e.g. torch symmetric
 * memory), applies G^ and stores the result back in place.  The caller issues a device barrier across
 * the ranks before and after stage 3 of fh_ga_slab_stage.  N0 in {16,...,2048} powers of two. */
int fh_ga_slab_peer(fh_ga* op, int world, int rank, const void* const* peer_spec);
/* "Push" exchange (csrc/fh_slab2.cu): S2 stores every output row k1 into the y-slab spectrum of the rank that owns
 * k1, S3 stores every output row i0 into the x-slab spectrum of the rank that owns plane i0 (NVLink peer stores, no
 * remote loads, no exchange buffer, no extra pass).  peer_spec[g] / peer_specT[g] = rank g's fh_ga_buffers pointers as
 * mapped into this process (symmetric memory).  fh_ga_slab_push_stage runs stage 1..5 with the CG fusions; the caller
 * places a device barrier across the ranks between stages 2|3 and 3|4.  Replaces one Afun(P) of general/solver.py:125
 * on a decomposed field.  Stages 1 and 2 work on x-plane chunk `chunk` of `nchunk` (S2 of one chunk is NVLink-bound,
 * S1 of the next HBM-bound: run on two streams they overlap).  N0 in {128, 256, 512}, N1 a power of two 16..2048. */
int fh_ga_slab_push(fh_ga* op, int world, int rank, const void* const* peer_spec, const void* const* peer_specT);
int fh_ga_slab_push_stage(fh_ga* op, int stage, int chunk, int nchunk, double* p, const double* r, int pupdate,
                          double* y);
/* k2-block exchange pipeline (csrc/fh_slab2.cu): the half-spectrum columns are split into nblk blocks of whole 8-column
 * tiles; S2, exchange, S3, exchange back and S4 are independent across blocks, so block b travels on the copy engines
 * while its neighbours are transformed.  bufA / bufB: zero-filled exchange buffers of D*n0_local*N1*pitch complex128,
 * block b = [world][D][n0_local][n1_local][width_b] at element offset `base` (fh_ga_slab_kblock_info): the piece for
 * peer g is the g-th of `world` contiguous runs of `per_peer` elements.  Stages: 1 = S1 (whole slab), 2 = S2 of block
 * blk -> bufA, 3 = S3 in place on block blk of bufB, 4 = S4 of block blk from bufA, 5 = S5 (whole slab).
 * N0 in {128, 256, 512}, N1 a power of two 16..2048. */
int fh_ga_slab_kblock(fh_ga* op, int world, int nblk, void* bufA, void* bufB);
int fh_ga_slab_kblock_info(const fh_ga* op, int blk, int64_t* base, int64_t* per_peer, int* col0, int* width);
int fh_ga_slab_kblock_stage(fh_ga* op, int stage, int blk, double* p, const double* r, int pupdate, double* y);
/* one pipeline step with the CG fusions of fh_cg_steps (p = r + beta p in S1 when pupdate, <p,y> in S5):
 * plain stages 1..5 without fh_ga_slab_direct; with it stage 1 = S1+S2 of `chunk` -> bufA,
 * 3 = S3 on bufB, 4 = S4+S5 of `chunk` from bufA.  Replaces one Afun(P) of general/solver.py:125 */
int fh_ga_slab_stage(fh_ga* op, int stage, int chunk, double* p, const double* r, int pupdate, double* y);
/* CG of general/solver.py:113-136 split at its two global reductions: each piece leaves per-CTA partial
 * sums, fh_cgd_local_sum folds them into sum_dev[0] (device), the caller all-reduces sum_dev over the
 * ranks, fh_cgd_scal turns the global sum into rr/norm (mode 0), alpha (1) or beta/rr/norm (2). */
int fh_cgd_init(fh_ga* op, const double* B, double* vecs);
int fh_cgd_update(fh_ga* op, double* x, double* vecs);
/* deferred x update (what fh_cg_steps does internally): with fh_ga_set_xacc(op, x) the next S1 launches that
 * carry the p update also apply x += alpha p of the previous iteration (solver.py:127) while p is in registers;
 * fh_cgd_update_r is then the update without x (r -= alpha Ap, partial <r,r>), fh_cgd_xflush applies the
 * pending x += alpha p after the last iteration.  fh_ga_can_defer_x: 1 if this operator's S1 path supports it. */
int fh_ga_set_xacc(fh_ga* op, double* x);
int fh_ga_can_defer_x(const fh_ga* op);
int fh_cgd_update_r(fh_ga* op, double* vecs);
int fh_cgd_xflush(fh_ga* op, double* x, const double* vecs);
int fh_cgd_local_sum(fh_ga* op, double* sum_dev);
int fh_cgd_scal(fh_ga* op, const double* sum_dev, int mode, double* norm_host);
/* this rank's sum of x*y left on the device by the last fh_ga_stage(op, 5, x, y) */
int fh_ga_last_dot(fh_ga* op, double* result_host);
int fh_ga_destroy(fh_ga* op);
int fh_ga_apply(fh_ga* op, const double* x, double* y);
/* flags: bit0/1/2 = register-resident power-of-two kernels on the last axis / axis 1 / axis 0 */
int fh_ga_config(const fh_ga* op, int* flags, int* pitch, int* mid_T);
/* run one stage (1..5) of the pipeline: S1 A.x + R2C, S2 C2C axis 1, S3 axis 0 + G^, S4, S5 C2R (+<x,y>) */
int fh_ga_stage(fh_ga* op, int stage, const double* x, double* y);
/* vecs: 3*D*prod(N) doubles of scratch (r, p, Ap); x holds x0 on entry, the solution on exit;
 * hist_host (optional, hist_cap entries) receives the residual norm after every iteration,
 * entry 0 = initial residual                                                          */
int fh_cg(fh_ga* op, const double* B, double* x, double tol, int64_t maxiter, double* vecs, int64_t* kit_host,
          double* norm_res_host, double* hist_host, int64_t hist_cap);
/* the same loop in two calls: begin = initial residual (one operator application), steps = up to
 * nsteps iterations (stops early at norm_res <= tol); *norm_res_host carries the state in between */
int fh_cg_begin(fh_ga* op, const double* B, double* x, double* vecs, double* norm_res_host);
int fh_cg_steps(fh_ga* op, double* x, double* vecs, double tol, int64_t nsteps, int64_t* done_host,
                double* norm_res_host, double* hist_host);
/* vecs: 2*D*prod(N) doubles */
int fh_richardson(fh_ga* op, const double* B, double* x, double alpha, double tol, int64_t maxiter, double* vecs,
                  int64_t* kit_host, double* norm_res_host);
/* stand-alone CG vector updates (slab loop: scalars are reduced across ranks by the caller) */
int fh_cg_xr_update(int64_t n, double* x, double* r, const double* p, const double* Ap, double alpha,
                    double* rr_local_host);
int fh_cg_p_update(int64_t n, double* p, const double* r, double beta);
/* kernels launched by this library since fh_init (for bench.py's gpu_launches) */
int64_t fh_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
