// fh_plan.cuh — host-side FFT plan (per grid N): radix factorisation and fp64
// twiddle tables per axis.  Opaque to the C ABI (struct fh_plan).
#pragma once
#include "fh_fft.cuh"

struct fh_plan {
    int dim;
    int N[3];     // grid
    int nh;       // N[dim-1]/2 + 1   (reference: tensors/objects.py:79-83)
    AxisDesc ax[3];
    cplx* tw_dev[3];
    int64_t nreal;  // prod(N)
    int64_t nspec;  // prod(N[:-1]) * nh
};

// kernels launchers shared between translation units
int fh_launch_r2c_last(const fh_plan* p, const double* x, cplx* X, int64_t nlines, int pitch);
int fh_launch_c2r_last(const fh_plan* p, const cplx* X, double* x, int64_t nlines, int pitch, double scale);
int fh_launch_c2c_strided(const AxisDesc& ax, const cplx* in, cplx* out, int64_t outer, int64_t inner, bool inverse,
                          double scale);
