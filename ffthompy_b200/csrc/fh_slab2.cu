// fh_slab2.cu — "push" exchange of the slab-decomposed operator: the two FFT transposes ride inside the
// store phases of the kernels that produce the data, over NVLink peer mappings.
//
//   S1   local  (A p, R2C last axis)                                  -> this rank's x-slab spectrum   [D][n0l][N1][P]
//   S2   C2C along axis 1, each output row k1 STORED into the y-slab spectrum [D][N0][n1l][P] of the rank that owns k1
//   ---- device barrier across the ranks (every rank's rows have landed)
//   S3   C2C axis 0 + G^ + inverse on this rank's y-slab spectrum, each output row i0 STORED into the x-slab spectrum
//        of the rank that owns plane i0
//   ---- device barrier
//   S4, S5 local
// No exchange buffer, no copy-engine or NCCL traffic, no extra pass over the data, and no remote LOADS (round 1's
// "peer" scheme pulled its rows and paid the NVLink round-trip latency inside S3): stores are fire-and-forget, so the
// transfer overlaps the transforms of the same kernel.  Six kernels + two barriers per operator application.
// The caller (ffthompy_b200/slab.py, exchange 'push') owns the symmetric-memory workspace and issues the barriers.
#include "fh_ga.cuh"
#include "fh_reg3.h"
#include "fh_mid2.h"
#include "../../include/ffthom_b200.h"
#include <stdlib.h>

// axis-1 lengths served by the mapped (out-of-place, LineMap) C2C kernels (fh_fast.cuh: k_c2c_map, fh_reg3.cuh)
static inline bool fh_map_len_host(int n) { return n >= 16 && n <= 2048 && (n & (n - 1)) == 0; }

// peer_spec[g] / peer_specT[g]: rank g's x-slab / y-slab spectrum (fh_ga_buffers) as mapped into THIS process
extern "C" int fh_ga_slab_push(fh_ga* op, int world, int rank, const void* const* peer_spec,
                               const void* const* peer_specT) {
    FH_REQUIRE(op && peer_spec && peer_specT && world >= 1 && rank >= 0 && rank < world, "fh_ga_slab_push: bad argument");
    const fh_plan* p = op->plan;
    FH_REQUIRE(p->dim == 3, "fh_ga_slab_push: a 3-D slab operator is required");
    FH_REQUIRE((int64_t)op->n0l * world == p->N[0] && (int64_t)op->n1l * world == p->N[1],
               "fh_ga_slab_push: slab extents do not match world=%d", world);
    FH_REQUIRE(op->g.ioff1 == rank * op->n1l, "fh_ga_slab_push: rank %d does not own the k1 range of this operator", rank);
    FH_REQUIRE(peer_spec[rank] == (const void*)op->spec && peer_specT[rank] == (const void*)op->specT,
               "fh_ga_slab_push: the entries of this rank must be this operator's spectra");
    const int N0 = p->N[0], N1 = p->N[1], P = op->pitch, n0l = op->n0l, n1l = op->n1l;
    if (!(N0 == 512 || fh_mid2_can(N0)))
        return fh_set_error(FH_ERR_UNSUPPORTED, "fh_ga_slab_push: N0=%d has no push variant of the axis-0 kernel", N0);
    if (!fh_map_len_host(N1))
        return fh_set_error(FH_ERR_UNSUPPORTED, "fh_ga_slab_push: N1=%d not in the mapped axis-1 kernel family", N1);
    if (((int64_t)n1l * P) % 8)
        return fh_set_error(FH_ERR_UNSUPPORTED, "fh_ga_slab_push: %d local rows of pitch %d", n1l, P);
    int64_t* h = (int64_t*)malloc(sizeof(int64_t) * (N0 + N1));
    if (!h) return fh_set_error(FH_ERR_ALLOC, "fh_ga_slab_push: out of host memory");
    for (int g = 0; g < world; ++g) {
        const intptr_t dS = (intptr_t)peer_spec[g] - (intptr_t)op->spec, dT = (intptr_t)peer_specT[g] - (intptr_t)op->specT;
        if (!peer_spec[g] || !peer_specT[g] || dS % (intptr_t)sizeof(cplx) || dT % (intptr_t)sizeof(cplx)) {
            free(h);
            return fh_set_error(FH_ERR_ARG, "fh_ga_slab_push: spectrum pointers of peer %d are null or misaligned", g);
        }
        // S2: row k1 of panel (c, i0l) -> peer g = k1 / n1l, element ((c*N0 + rank*n0l + i0l)*n1l + k1 % n1l)*P
        for (int k1l = 0; k1l < n1l; ++k1l)
            h[g * n1l + k1l] = (int64_t)(dT / (intptr_t)sizeof(cplx)) + ((int64_t)rank * n0l * n1l + k1l) * P;
        // S3: row i0 of component c -> peer g = i0 / n0l, element ((c*n0l + i0l)*N1 + rank*n1l)*P + ii
        for (int i0l = 0; i0l < n0l; ++i0l)
            h[N1 + g * n0l + i0l] = (int64_t)(dS / (intptr_t)sizeof(cplx)) + ((int64_t)i0l * N1 + (int64_t)rank * n1l) * P;
    }
    if (op->sp_off1) cudaFree(op->sp_off1);
    op->sp_off1 = op->sp_off0 = NULL;
    cudaError_t e = cudaMalloc((void**)&op->sp_off1, sizeof(int64_t) * (N0 + N1));
    if (e == cudaSuccess) e = cudaMemcpy(op->sp_off1, h, sizeof(int64_t) * (N0 + N1), cudaMemcpyHostToDevice);
    free(h);
    if (e != cudaSuccess) return fh_set_error(FH_ERR_CUDA, "fh_ga_slab_push: %s", cudaGetErrorString(e));
    op->sp_off0 = op->sp_off1 + N1;
    op->sp_world = world;
    return FH_OK;
}

// one stage (1..5) of the push pipeline with the CG fusions of fh_cg_steps (p = r + beta p in S1 when pupdate,
// partial sums of <p, y> in S5); the caller puts a device barrier between stages 2 | 3 and 3 | 4.
// Stages 1 and 2 take an x-plane chunk (chunk of nchunk; nchunk = 1: the whole slab): S2 of chunk j is bound by the
// NVLink stores, S1 of chunk j+1 by HBM, so the caller runs them on two streams (fh_set_stream) and they overlap.
extern "C" int fh_ga_slab_push_stage(fh_ga* op, int stage, int chunk, int nchunk, double* p, const double* r, int pupdate,
                                     double* y) {
    FH_REQUIRE(op && p && y && op->sp_world >= 1, "fh_ga_slab_push_stage: fh_ga_slab_push has not been set up");
    const fh_plan* pl = op->plan;
    const int D = op->D, P = op->pitch, N0 = pl->N[0], N1 = pl->N[1];
    const int64_t inner = (int64_t)op->n1l * P;
    FH_REQUIRE(nchunk >= 1 && chunk >= 0 && chunk < nchunk && op->n0l % nchunk == 0,
               "fh_ga_slab_push_stage: chunk %d of %d over %d planes", chunk, nchunk, op->n0l);
    const int n0c = op->n0l / nchunk;
    int np = 0, rc;
    switch (stage) {
        case 1:
            if (nchunk > 1) {
                op->row_beg = (int64_t)chunk * n0c * N1;
                op->row_cnt = (int64_t)n0c * N1;
            }
            rc = fh_ga_stage_local(op, 1, p, r, pupdate, y, 0, NULL);
            op->row_beg = op->row_cnt = 0;
            return rc;
        case 2: {
            const LineMap nat = {NULL, (int64_t)P, (int64_t)op->n0l * N1 * P, (int64_t)N1 * P, n0c};
            const LineMap rem = {op->sp_off1, 0, (int64_t)N0 * inner, inner, n0c};
            const cplx* src = op->spec + (size_t)chunk * n0c * N1 * P;
            cplx* dst = op->specT + (size_t)chunk * n0c * inner;
            if (nchunk > 1 && fh_reg3_map_len(N1) && P % 8 == 0) {
                // persistent: FH_PUSH_S2_CTAS CTAs per SM (default 1) leave room for the S1 CTAs of the next chunk
                static const int per_sm = getenv("FH_PUSH_S2_CTAS") ? atoi(getenv("FH_PUSH_S2_CTAS")) : 1;
                return fh_reg3_c2c_map(N1, pl->ax[1].tw, src, dst, nat, rem, (int64_t)D * n0c, P, false,
                                       per_sm > 0 ? per_sm * fh_num_sms() : 0);
            }
            return fh_launch_c2c_map(N1, pl->ax[1].tw, src, dst, nat, rem, (int64_t)D * n0c, P, false);
        }
        case 3:
            if (N0 == 512)
                return fh_reg3_mid_green_push(N0, op->g.kind, op->specT, op->spec, pl->ax[0].tw, op->g, inner, pl->nh, P,
                                              op->sp_off0, (int64_t)op->n0l * N1 * P);
            return fh_mid2_green(N0, op->g.kind, op->specT, pl->ax[0].tw, op->g, NULL, inner, (int64_t)N0 * inner, P, 0,
                                 pl->nh, op->n1l, 0, P, op->spec, op->sp_off0, (int64_t)op->n0l * N1 * P);
        case 4: return fh_ga_stage_local(op, 4, p, NULL, 0, y, 0, NULL);
        case 5: return fh_ga_stage_local(op, 5, p, NULL, 0, y, 1, &np);
    }
    return fh_set_error(FH_ERR_ARG, "fh_ga_slab_push_stage: stage %d", stage);
}

// ------------------------------------------------------------------ k2-block exchange pipeline
// The half-spectrum columns are split into nblk blocks of whole 8-column tiles.  S2, the exchange, S3, the exchange back
// and S4 are independent across blocks (only S1 and S5 need whole rows), so with the blocks in flight on the copy
// engines the NVLink transfer of block b overlaps S2 / S3 / S4 of its neighbours and only one block's transfer per
// direction is exposed (task 3):
//   S1 (all rows)  |  for b: S2_b -> bufA_b, push bufA_b[g] -> peer g's bufB_b[me]
//                  |  for b: (pushes of block b landed everywhere)  S3_b in place on bufB_b, push back into bufA_b
//                  |  for b: (block b is back)  S4_b: bufA_b -> spectrum columns of block b   |  S5 (all rows)
// Exchange buffers: block b = [G][D][n0l][n1l][w_b] complex at element offset kb_base[b]; on the x-slab side G indexes the
// peer that owns the k1 range, on the y-slab side the peer that owns the x-planes, so every (block, peer) piece is one
// contiguous copy.  bufA / bufB: D*n0l*N1*pitch complex each, zero-filled by the caller (symmetric memory).
extern "C" int fh_ga_slab_kblock(fh_ga* op, int world, int nblk, void* bufA, void* bufB) {
    FH_REQUIRE(op && bufA && bufB && bufA != bufB && world >= 1, "fh_ga_slab_kblock: bad argument");
    const fh_plan* p = op->plan;
    FH_REQUIRE(p->dim == 3, "fh_ga_slab_kblock: a 3-D slab operator is required");
    FH_REQUIRE((int64_t)op->n0l * world == p->N[0] && (int64_t)op->n1l * world == p->N[1],
               "fh_ga_slab_kblock: slab extents do not match world=%d", world);
    const int N0 = p->N[0], N1 = p->N[1], P = op->pitch, n0l = op->n0l, n1l = op->n1l, D = op->D;
    const int ntile = P / 8;
    FH_REQUIRE(P % 8 == 0 && nblk >= 1 && nblk <= 16 && nblk <= ntile, "fh_ga_slab_kblock: %d blocks over %d tiles", nblk, ntile);
    if (!(N0 == 512 || fh_mid2_can(N0)))
        return fh_set_error(FH_ERR_UNSUPPORTED, "fh_ga_slab_kblock: N0=%d has no column-block variant of the axis-0 kernel", N0);
    if (!fh_map_len_host(N1))
        return fh_set_error(FH_ERR_UNSUPPORTED, "fh_ga_slab_kblock: N1=%d not in the mapped axis-1 kernel family", N1);
    int64_t* h = (int64_t*)malloc(sizeof(int64_t) * (size_t)nblk * (N0 + N1));
    if (!h) return fh_set_error(FH_ERR_ALLOC, "fh_ga_slab_kblock: out of host memory");
    int c0 = 0;
    for (int b = 0; b < nblk; ++b) {
        const int tiles = ntile / nblk + (b < ntile % nblk ? 1 : 0);
        const int w = tiles * 8;
        op->kb_col0[b] = c0;
        op->kb_w[b] = w;
        op->kb_base[b] = (int64_t)D * n0l * N1 * c0;
        const int64_t inner = (int64_t)n1l * w;
        for (int k1 = 0; k1 < N1; ++k1) h[(size_t)b * N1 + k1] = (int64_t)(k1 / n1l) * D * n0l * inner + (int64_t)(k1 % n1l) * w;
        for (int i0 = 0; i0 < N0; ++i0)
            h[(size_t)nblk * N1 + (size_t)b * N0 + i0] = ((int64_t)(i0 / n0l) * D * n0l + (i0 % n0l)) * inner;
        c0 += w;
    }
    if (op->kb_off) cudaFree(op->kb_off);
    op->kb_off = NULL;
    cudaError_t e = cudaMalloc((void**)&op->kb_off, sizeof(int64_t) * (size_t)nblk * (N0 + N1));
    if (e == cudaSuccess) e = cudaMemcpy(op->kb_off, h, sizeof(int64_t) * (size_t)nblk * (N0 + N1), cudaMemcpyHostToDevice);
    free(h);
    if (e != cudaSuccess) return fh_set_error(FH_ERR_CUDA, "fh_ga_slab_kblock: %s", cudaGetErrorString(e));
    op->kb_world = world;
    op->kb_nblk = nblk;
    op->kb_bufA = (cplx*)bufA;
    op->kb_bufB = (cplx*)bufB;
    return FH_OK;
}

// element offset of block `blk` inside the exchange buffers and the elements one peer receives from this rank
extern "C" int fh_ga_slab_kblock_info(const fh_ga* op, int blk, int64_t* base, int64_t* per_peer, int* col0, int* width) {
    FH_REQUIRE(op && op->kb_world >= 1 && blk >= 0 && blk < op->kb_nblk, "fh_ga_slab_kblock_info: bad block");
    if (base) *base = op->kb_base[blk];
    if (per_peer) *per_peer = (int64_t)op->D * op->n0l * op->n1l * op->kb_w[blk];
    if (col0) *col0 = op->kb_col0[blk];
    if (width) *width = op->kb_w[blk];
    return FH_OK;
}

// stage 1: S1 (whole slab);  2: S2 of column block blk -> bufA;  3: S3 in place on block blk of bufB;
// 4: S4 of block blk from bufA into the x-slab spectrum;  5: S5 (whole slab, partial sums of <p, y>)
extern "C" int fh_ga_slab_kblock_stage(fh_ga* op, int stage, int blk, double* p, const double* r, int pupdate, double* y) {
    FH_REQUIRE(op && p && y && op->kb_world >= 1, "fh_ga_slab_kblock_stage: fh_ga_slab_kblock has not been set up");
    FH_REQUIRE(blk >= 0 && blk < op->kb_nblk, "fh_ga_slab_kblock_stage: block %d out of range", blk);
    const fh_plan* pl = op->plan;
    const int D = op->D, P = op->pitch, N0 = pl->N[0], N1 = pl->N[1], n0l = op->n0l, n1l = op->n1l;
    const int w = op->kb_w[blk], c0 = op->kb_col0[blk];
    const int64_t inner = (int64_t)n1l * w;
    const int64_t* off1 = op->kb_off + (size_t)blk * N1;
    const int64_t* off0 = op->kb_off + (size_t)op->kb_nblk * N1 + (size_t)blk * N0;
    const LineMap nat = {NULL, (int64_t)P, (int64_t)n0l * N1 * P, (int64_t)N1 * P, n0l};
    const LineMap blkmap = {off1, 0, (int64_t)n0l * inner, inner, n0l};
    int np = 0;
    switch (stage) {
        case 1: return fh_ga_stage_local(op, 1, p, r, pupdate, y, 0, NULL);
        case 2:
            return fh_launch_c2c_map(N1, pl->ax[1].tw, op->spec + c0, op->kb_bufA + op->kb_base[blk], nat, blkmap,
                                     (int64_t)D * n0l, w, false);
        case 3:
            if (N0 == 512)
                return fh_reg3_mid_green_map(N0, op->g.kind, op->kb_bufB + op->kb_base[blk], pl->ax[0].tw, op->g, inner,
                                             pl->nh, w, off0, (int64_t)n0l * inner, c0);
            return fh_mid2_green(N0, op->g.kind, op->kb_bufB + op->kb_base[blk], pl->ax[0].tw, op->g, off0, 0,
                                 (int64_t)n0l * inner, w, c0, pl->nh, n1l, 0, w);
        case 4:
            return fh_launch_c2c_map(N1, pl->ax[1].tw, op->kb_bufA + op->kb_base[blk], op->spec + c0, blkmap, nat,
                                     (int64_t)D * n0l, w, true);
        case 5: return fh_ga_stage_local(op, 5, p, NULL, 0, y, 1, &np);
    }
    return fh_set_error(FH_ERR_ARG, "fh_ga_slab_kblock_stage: stage %d", stage);
}
