// fh_reg3.cu — launchers of the register-resident three-pass kernels (fh_reg3.cuh) for axis length 512.
// Internal C++ interface (fh_reg3.h) used by the fused operator in fh_fused.cu; nothing here is part of the C ABI.
#include "fh_reg3.cuh"
#include "fh_reg3.h"
#include "fh_mid512.h"
#include <stdlib.h>

template <typename K>
static int reg3_smem_attr(K kernel, size_t bytes) {
    if (bytes > (size_t)fh_max_smem_optin())
        return fh_set_error(FH_ERR_UNSUPPORTED, "three-pass kernel needs %zu bytes of shared memory", bytes);
    FH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return FH_OK;
}

static int reg3_env(const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
}
// 512 always; 1024 = 8 x 8 x 16 unless FH_REG3_1024=0 (then the generic shared-memory routine of fh_fast.cuh serves it)
bool fh_reg3_last_len(int n) {
    static const int l1024 = reg3_env("FH_REG3_1024", 1);
    return n == 512 || (n == 1024 && l1024);
}
// axis-0 kernel: 512 always; 256 (8 columns per tile, 128-byte segments) when FH_MID256_REG3=1
bool fh_reg3_mid_len(int n) {
    static const int m256 = reg3_env("FH_MID256_REG3", 0);
    return n == 512 || (n == 256 && m256);
}

template <int N, int D, int TRW, int ALAY>
static int fwd_last_A(const Reg3LastArgs& a) {
    constexpr int NP = D * TRW / 2, NPAD = LinePad<N>::NPAD;
    const size_t smem = (size_t)2 * NP * NPAD * sizeof(double);
    int rc;
    if ((rc = reg3_smem_attr(k_fwd_last_reg3<N, D, TRW, ALAY>, smem))) return rc;
    k_fwd_last_reg3<N, D, TRW, ALAY><<<a.nblk, NP * Reg3Cfg<N>::TPL, smem, fh_stream()>>>(
        a.A, a.phase, a.lut, *a.lutc, a.nphase, a.p, a.r, a.scal, a.pupdate, a.spec, a.tw, a.nrows, a.nh, a.pitch,
        a.xacc);
    FH_LAUNCH_CHECK();
    return FH_OK;
}
template <int N, int D, int TRW>
static int fwd_last_D(const Reg3LastArgs& a) {
    switch (a.alay) {
        case -1: return fwd_last_A<N, D, TRW, -1>(a);
        case 0: return fwd_last_A<N, D, TRW, 0>(a);
        case 1: return fwd_last_A<N, D, TRW, 1>(a);
        case 2: return fwd_last_A<N, D, TRW, 2>(a);
        case 3: return fwd_last_A<N, D, TRW, 3>(a);
    }
    return fh_set_error(FH_ERR_ARG, "three-pass S1: bad coefficient layout %d", a.alay);
}
int fh_reg3_fwd_last(int N, int D, int trw, const Reg3LastArgs& a) {
    if (N == 512 && D == 6 && trw == 2) return fwd_last_D<512, 6, 2>(a);
    if (N == 512 && D == 3 && trw == 4) return fwd_last_D<512, 3, 4>(a);
    if (N == 512 && D == 2 && trw == 4) return fwd_last_D<512, 2, 4>(a);
    if (N == 1024 && D == 6 && trw == 2) return fwd_last_D<1024, 6, 2>(a);
    if (N == 1024 && D == 3 && trw == 4) return fwd_last_D<1024, 3, 4>(a);
    if (N == 1024 && D == 2 && trw == 4) return fwd_last_D<1024, 2, 4>(a);
    return fh_set_error(FH_ERR_UNSUPPORTED, "no three-pass last-axis kernel for N=%d D=%d TRW=%d", N, D, trw);
}

template <int N, int D, int TRW>
static int inv_last_D(const Reg3InvArgs& a) {
    constexpr int NP = D * TRW / 2, NPAD = LinePad<N>::NPAD;
    const size_t smem = (size_t)2 * NP * NPAD * sizeof(double);
    int rc;
    if ((rc = reg3_smem_attr(k_inv_last_reg3<N, D, TRW>, smem))) return rc;
    k_inv_last_reg3<N, D, TRW><<<a.nblk, NP * Reg3Cfg<N>::TPL, smem, fh_stream()>>>(a.spec, a.y, a.pdot, a.part, a.tw,
                                                                                   a.nrows, a.nh, a.pitch, a.scale);
    FH_LAUNCH_CHECK();
    return FH_OK;
}
int fh_reg3_inv_last(int N, int D, int trw, const Reg3InvArgs& a) {
    if (N == 512 && D == 6 && trw == 2) return inv_last_D<512, 6, 2>(a);
    if (N == 512 && D == 3 && trw == 4) return inv_last_D<512, 3, 4>(a);
    if (N == 512 && D == 2 && trw == 4) return inv_last_D<512, 2, 4>(a);
    if (N == 1024 && D == 6 && trw == 2) return inv_last_D<1024, 6, 2>(a);
    if (N == 1024 && D == 3 && trw == 4) return inv_last_D<1024, 3, 4>(a);
    if (N == 1024 && D == 2 && trw == 4) return inv_last_D<1024, 2, 4>(a);
    return fh_set_error(FH_ERR_UNSUPPORTED, "no three-pass last-axis kernel for N=%d D=%d TRW=%d", N, D, trw);
}

template <int N, int T, int KIND, int DIM, int NT>
static int mid_KD_NT(cplx* data, const cplx* tw, const GreenDesc& g, int64_t inner, int nh, int pitch, const int64_t* rowoff,
                     int64_t cstride, cplx* dout, const int64_t* rowoff_out, int64_t cstride_out, int kcol0) {
    constexpr int D = (KIND == FH_GREEN_SCALAR) ? DIM : DIM * (DIM + 1) / 2;
    const size_t smem = (size_t)D * (N + N / 8) * T * sizeof(cplx);
    int rc;
    if ((rc = reg3_smem_attr(k_mid_green_reg3<N, T, KIND, DIM, NT>, smem))) return rc;
    k_mid_green_reg3<N, T, KIND, DIM, NT><<<(unsigned)(inner / T), NT, smem, fh_stream()>>>(
        data, tw, g, inner, nh, pitch, rowoff, rowoff ? cstride : (int64_t)N * inner, dout, rowoff_out, cstride_out, kcol0);
    FH_LAUNCH_CHECK();
    return FH_OK;
}
// threads per CTA of the axis-0 kernel: 768 (85 registers, the Green stage spills ~200 B) or 512 (128 registers, no
// spill, three rounds per pass instead of two); FH_REG3_NT picks, the default is what measured faster on B200
template <int N, int T, int KIND, int DIM>
static int mid_KD(cplx* data, const cplx* tw, const GreenDesc& g, int64_t inner, int nh, int pitch,
                  const int64_t* rowoff = nullptr, int64_t cstride = 0, cplx* dout = nullptr,
                  const int64_t* rowoff_out = nullptr, int64_t cstride_out = 0, int kcol0 = 0) {
    // N0 = 512 in 3-D: the two-stage 32 x 16 kernel (fh_mid512.cu) unless FH_MID512=0
    if (N == 512 && DIM == 3 && T == 4 && fh_mid512_on())
        return fh_mid512_green(KIND, data, tw, g, inner, nh, pitch, rowoff, cstride, dout, rowoff_out, cstride_out, kcol0);
    static const int nt = reg3_env("FH_REG3_NT", 768);
    if (nt == 512)
        return mid_KD_NT<N, T, KIND, DIM, 512>(data, tw, g, inner, nh, pitch, rowoff, cstride, dout, rowoff_out, cstride_out, kcol0);
    if (nt == 384)
        return mid_KD_NT<N, T, KIND, DIM, 384>(data, tw, g, inner, nh, pitch, rowoff, cstride, dout, rowoff_out, cstride_out, kcol0);
    return mid_KD_NT<N, T, KIND, DIM, 768>(data, tw, g, inner, nh, pitch, rowoff, cstride, dout, rowoff_out, cstride_out, kcol0);
}
// push mode (fh_slab2.cu): natural y-slab input `data`, output rows scattered through rowoff_out into `dout`
int fh_reg3_mid_green_push(int N, int kind, cplx* data, cplx* dout, const cplx* tw, const GreenDesc& g, int64_t inner, int nh,
                           int pitch, const int64_t* rowoff_out, int64_t cstride_out) {
    if (N != 512 || inner % 4 != 0)
        return fh_set_error(FH_ERR_UNSUPPORTED, "no three-pass axis-0 push kernel for N0=%d inner=%lld", N, (long long)inner);
    if (kind == FH_GREEN_SCALAR)
        return mid_KD<512, 4, FH_GREEN_SCALAR, 3>(data, tw, g, inner, nh, pitch, nullptr, 0, dout, rowoff_out, cstride_out);
    return mid_KD<512, 4, FH_GREEN_ELASTIC, 3>(data, tw, g, inner, nh, pitch, nullptr, 0, dout, rowoff_out, cstride_out);
}
int fh_reg3_mid_green(int N, int kind, int dim, cplx* data, const cplx* tw, const GreenDesc& g, int64_t inner, int nh,
                      int pitch) {
    if (N == 256 && inner % 8 == 0 && kind == FH_GREEN_ELASTIC && dim == 3)
        return mid_KD<256, 8, FH_GREEN_ELASTIC, 3>(data, tw, g, inner, nh, pitch);
    if (N != 512 || inner % 4 != 0)
        return fh_set_error(FH_ERR_UNSUPPORTED, "no three-pass axis-0 kernel for N0=%d inner=%lld", N, (long long)inner);
    if (kind == FH_GREEN_SCALAR)
        return (dim == 3) ? mid_KD<512, 4, FH_GREEN_SCALAR, 3>(data, tw, g, inner, nh, pitch)
                          : mid_KD<512, 4, FH_GREEN_SCALAR, 2>(data, tw, g, inner, nh, pitch);
    return (dim == 3) ? mid_KD<512, 4, FH_GREEN_ELASTIC, 3>(data, tw, g, inner, nh, pitch)
                      : mid_KD<512, 4, FH_GREEN_ELASTIC, 2>(data, tw, g, inner, nh, pitch);
}

// slab-exchange layout (fh_ga_slab_direct): rows through rowoff[], component stride cstride; 3-D only
bool fh_reg3_map_len(int n) {
    static const int m256 = reg3_env("FH_MAP256_REG3", 0);
    return n == 512 || (n == 256 && m256);
}
int fh_reg3_mid_green_map(int N, int kind, cplx* data, const cplx* tw, const GreenDesc& g, int64_t inner, int nh,
                          int pitch, const int64_t* rowoff, int64_t cstride, int kcol0) {
    if (N == 512 && inner % 4 == 0) {
        if (kind == FH_GREEN_SCALAR)
            return mid_KD<512, 4, FH_GREEN_SCALAR, 3>(data, tw, g, inner, nh, pitch, rowoff, cstride, nullptr, nullptr, 0, kcol0);
        return mid_KD<512, 4, FH_GREEN_ELASTIC, 3>(data, tw, g, inner, nh, pitch, rowoff, cstride, nullptr, nullptr, 0, kcol0);
    }
    if (kcol0) return fh_set_error(FH_ERR_UNSUPPORTED, "three-pass axis-0 exchange kernel: column blocks need N0 = 512");
    if (N == 256 && inner % 8 == 0 && kind == FH_GREEN_ELASTIC)
        return mid_KD<256, 8, FH_GREEN_ELASTIC, 3>(data, tw, g, inner, nh, pitch, rowoff, cstride);
    return fh_set_error(FH_ERR_UNSUPPORTED, "no three-pass axis-0 exchange kernel for N0=%d", N);
}
template <int N, int T>
static int c2c_map_NT(const cplx* tw, const cplx* in, cplx* out, const LineMap& mi, const LineMap& mo, int64_t panels,
                      int pitch, bool inv, int max_ctas) {
    const size_t smem = (size_t)N * T * sizeof(cplx);
    const int ntile = pitch / T;
    const int64_t nwork = panels * ntile;
    const unsigned nblk = (unsigned)((max_ctas > 0 && max_ctas < nwork) ? max_ctas : nwork);
    const int nt = T * Reg3Cfg<N>::TPL;
    int rc;
    if (inv) {
        if ((rc = reg3_smem_attr(k_c2c_reg3_map<N, T, true>, smem))) return rc;
        k_c2c_reg3_map<N, T, true><<<nblk, nt, smem, fh_stream()>>>(in, out, tw, mi, mo, ntile, nwork);
    } else {
        if ((rc = reg3_smem_attr(k_c2c_reg3_map<N, T, false>, smem))) return rc;
        k_c2c_reg3_map<N, T, false><<<nblk, nt, smem, fh_stream()>>>(in, out, tw, mi, mo, ntile, nwork);
    }
    FH_LAUNCH_CHECK();
    return FH_OK;
}
int fh_reg3_c2c_map(int N, const cplx* tw, const cplx* in, cplx* out, const LineMap& mi, const LineMap& mo,
                    int64_t panels, int pitch, bool inv, int max_ctas) {
    if (pitch % 8) return fh_set_error(FH_ERR_UNSUPPORTED, "three-pass exchange kernel: pitch %d", pitch);
    if (N == 512) return c2c_map_NT<512, 8>(tw, in, out, mi, mo, panels, pitch, inv, max_ctas);
    if (N == 256) return c2c_map_NT<256, 8>(tw, in, out, mi, mo, panels, pitch, inv, max_ctas);
    return fh_set_error(FH_ERR_UNSUPPORTED, "no three-pass exchange kernel for N1=%d", N);
}
