// fh_odd.h — internal: compile-time two-pass kernels of the fused operator for ODD axis lengths N = R1 x R2 with odd
// radices (fh_odd.cu).  255 = 15 x 17 is the exact-integration ("Ga", Nbar = 2N - 1) grid of a 128^3 problem — BASELINE
// config 2 — which until round 2 ran through the run-time-length kernels (k_*_rt) at a quarter of its roofline.
#pragma once
#include "fh_ga.cuh"

bool fh_odd_len(int n);  // lengths served (255)
bool fh_odd_on();        // FH_ODD=0 keeps the run-time-length kernels
// S2 / S4: C2C along a strided axis of [outer][N][inner], in place; inner % 8 == 0
int fh_odd_c2c(int N, const cplx* tw, cplx* data, int64_t outer, int64_t inner, bool inv);
// S3 (3-D, not slab-decomposed): forward axis 0, G^, inverse axis 0 on op->specT
int fh_odd_mid(fh_ga* op);
// S1 / S5 on all rows of the local fields (op->row_cnt must be 0)
int fh_odd_fwd_last(fh_ga* op, double* p, const double* r, int pupdate);
int fh_odd_inv_last(fh_ga* op, double* y, const double* pdot, int* npart);
