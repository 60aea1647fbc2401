// fh_reg3.h — internal interface between fh_fused.cu and the three-pass kernels of fh_reg3.cu (not C ABI).
#pragma once
#include "fh_green.cuh"
struct Lut2C;

struct Reg3LastArgs {  // S1 (k_fwd_last_reg3): pointers already offset to the first row of the launch
    const double* A;
    const unsigned char* phase;
    const double* lut;
    const Lut2C* lutc;
    int nphase, alay;
    double* p;
    const double* r;
    const double* scal;
    int pupdate;
    cplx* spec;
    const cplx* tw;
    int64_t nrows;
    int nh, pitch;
    double* xacc;
    unsigned nblk;
};
struct Reg3InvArgs {  // S5 (k_inv_last_reg3)
    const cplx* spec;
    double* y;
    const double* pdot;
    double* part;
    const cplx* tw;
    int64_t nrows;
    int nh, pitch;
    double scale;
    unsigned nblk;
};
bool fh_reg3_last_len(int n);
bool fh_reg3_mid_len(int n);
int fh_reg3_fwd_last(int N, int D, int trw, const Reg3LastArgs& a);
int fh_reg3_inv_last(int N, int D, int trw, const Reg3InvArgs& a);
int fh_reg3_mid_green(int N, int kind, int dim, cplx* data, const cplx* tw, const GreenDesc& g, int64_t inner, int nh,
                      int pitch);

// slab-exchange layouts (LineMap of fh_fast.cuh; rowoff / cstride of fh_ga_slab_direct)
struct LineMap;
bool fh_reg3_map_len(int n);
// kcol0: global column of buffer column 0 (k2-block exchange buffers whose rows hold `pitch` = block-width columns)
int fh_reg3_mid_green_map(int N, int kind, cplx* data, const cplx* tw, const GreenDesc& g, int64_t inner, int nh,
                          int pitch, const int64_t* rowoff, int64_t cstride, int kcol0 = 0);
// max_ctas > 0: persistent launch with at most that many CTAs walking the tiles (0: one CTA per tile)
int fh_reg3_c2c_map(int N, const cplx* tw, const cplx* in, cplx* out, const LineMap& mi, const LineMap& mo,
                    int64_t panels, int pitch, bool inv, int max_ctas = 0);
// push mode of the slab pipeline: S3 reads the natural y-slab spectrum `data` and writes row i0 of component c to
// dout[c*cstride_out + rowoff_out[i0] + ii] (the x-slab spectrum of the rank that owns plane i0, peer-mapped)
int fh_reg3_mid_green_push(int N, int kind, cplx* data, cplx* dout, const cplx* tw, const GreenDesc& g, int64_t inner, int nh,
                           int pitch, const int64_t* rowoff_out, int64_t cstride_out);
