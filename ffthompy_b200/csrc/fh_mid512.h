// fh_mid512.h — internal: two-stage (32 x 16) axis-0 + G^ kernel for N0 = 512 (fh_mid512.cu), called from fh_reg3.cu
#pragma once
#include "fh_green.cuh"
bool fh_mid512_on();   // FH_MID512=0 keeps the three-stage kernel of fh_reg3.cuh
int fh_mid512_green(int kind, cplx* data, const cplx* tw, const GreenDesc& g, int64_t inner, int nh, int pitch,
                    const int64_t* rowoff, int64_t cstride, cplx* dout, const int64_t* rowoff_out, int64_t cstride_out,
                    int kcol0);
