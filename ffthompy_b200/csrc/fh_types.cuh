// fh_types.cuh — small plain structs shared between the kernel headers (fh_fast.cuh, fh_reg3.cuh) and the host-side
// operator object (fh_ga.cuh), so that translation units can see `struct fh_ga` without the kernel bodies.
#pragma once
#include "fh_common.cuh"

// two-phase coefficient table passed BY VALUE (kernel parameter = constant bank): both matrices
// are read with uniform addresses and selected per voxel, so the per-voxel gather costs no shared
// memory bandwidth (A layout 3)
struct Lut2C {
    double c[2][36];
};

// run-time-length in-place FFT plan of one axis (kernels: fh_fast.cuh, k_*_rt)
struct RtPlan {
    int n;      // length
    int ns;     // number of stages (1..3)
    int R[3];   // radices, DIF order
    int NB[3];  // block size of stage s:  n, n/R0, n/(R0*R1)
    int TS[3];  // twiddle stride of stage s: 1, R0, R0*R1
    int npr;    // padded rows: n + n/16 + 1
};

// slab-exchange addressing (kernels: k_c2c_map, k_c2c_reg3_map): a line (panel o, row, column t) lives at
// base(o) + rowoff(row) + t
struct LineMap {
    const int64_t* off;        // per-row offsets (nullptr: row * rstride)
    int64_t rstride;
    int64_t cstride, istride;  // panel o = c * nper + i  ->  c * cstride + i * istride
    int nper;
};
__device__ __forceinline__ int64_t linemap_base(const LineMap& m, int64_t o) {
    const int64_t c = o / m.nper, i = o - c * m.nper;
    return c * m.cstride + i * m.istride;
}
__device__ __forceinline__ int64_t linemap_row(const LineMap& m, int row) {
    return m.off ? m.off[row] : (int64_t)row * m.rstride;
}

