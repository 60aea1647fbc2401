// fh_mid2.h — internal interface between fh_fused.cu and the 8-column axis-0 + G^ kernels of fh_mid2.cu (not C ABI).
#pragma once
#include "fh_green.cuh"

// lengths served (3-D, scalar D = 3 or elastic D = 6)
bool fh_mid2_len(int n);
// S3 in place.  Natural layout: rowoff = NULL, rstride = inner (= n1l * pitch), cstride = N * inner.
// Exchange-buffer layout: rowoff[i0] + c*cstride + ii.  The buffer rows hold `spitch` columns = the global columns
// kcol0.. (whole rows: spitch = pitch, kcol0 = 0); tiles cover buffer columns [col0, col0 + ncols) of `nrow` rows.
// dout != NULL (push mode of the slab pipeline): results of row i0 go to dout[c*cstride_out + rowoff_out[i0] + ii]
int fh_mid2_green(int N, int kind, cplx* data, const cplx* tw, const GreenDesc& g, const int64_t* rowoff,
                  int64_t rstride, int64_t cstride, int spitch, int kcol0, int nh, int nrow, int col0, int ncols,
                  cplx* dout = nullptr, const int64_t* rowoff_out = nullptr, int64_t cstride_out = 0);
bool fh_mid2_can(int n);  // lengths the kernel family covers (fh_mid2_len: those it is the default for)
