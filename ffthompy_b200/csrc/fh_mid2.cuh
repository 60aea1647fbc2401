// fh_mid2.cuh — stage S3 of the fused operator (C2C along axis 0, closed-form G^(xi), inverse C2C along axis 0,
// in place) for power-of-two N0 = 64, 128, 256 in 3-D: the round-2 replacement of k_mid_green_pipe.
//
// Reference semantics (unchanged): the middle factor of Operator([[FiN, G^, FN]]) — ffthompy/projections.py:54-91
// (scalar), :185-240 (elasticity) applied per frequency between numpy.fft.fftn / ifftn along axis 0
// (ffthompy/tensors/fft.py:39-43).
//
// What round 1's kernel lost, and what this one does about it (profiles/r01d_ncu_full_summary.md):
//  * 64-byte row segments (4-column tiles)  ->  8-column tiles: every global access is a full 128-byte line, the
//    pattern S2/S4 already run at copy speed;
//  * ten shared-memory passes per element (cp.async staging, F1, F2, G^, I2, I1 each read + write)  ->  six: F1
//    goes global -> registers -> shared, the last forward stage, G^ and the first inverse stage happen in the
//    registers of ONE thread that holds eight frequencies of ALL D components, and I1 goes shared -> registers ->
//    global;
//  * five CTA-wide barriers per tile  ->  three;
//  * no room for a second buffer at 8 columns (D*N0*8*16 B = 192 KB for D = 6, N0 = 256)  ->  the next tile's loads
//    are issued into registers (48 per thread) while I1 of the current tile computes and stores;
//  * 168 registers per thread  ->  256 threads per CTA (255 registers): the 96 data registers of the G^ stage fit.
//
//   N0 = RA x 16:  n = j + 16 r (j < 16, r < RA),  k = q + RA s (q < RA, s = 2a + p < 16)
//   F1   y_j[q] = w_N^(j q) sum_r x[j + 16 r] w_RA^(r q)                    task (c, t, j)     -> slot j
//   F2   z_p[j'] = (y_j' + (-1)^p y_(j'+8)) w_16^(j' p),  X[q + RA(2a+p)] = DFT8(z_p)[a]   task (t, q, p): all D
//   G^   on the eight frequencies k0 = q + RA (2a + p)                                      components, registers
//   I2   z'_p = IDFT8 over a                                                               -> slot j' + 8 p
//   I1   y'_j = z'_0[j mod 8] + w_16^(-j) z'_1[j mod 8],  x'[j + 16 r] = sum_q w_RA^(-r q) w_N^(-j q) y'_j[q]
// Shared memory [c][q][slot 0..15][t 0..7] (t = column, 128 B): every warp access is whole 128-byte segments,
// conflict free without padding.  256 threads: 3 F1/I1 tasks each for D = 6, one G^ half-task (p) each.
#pragma once
#include "fh_fft.cuh"
#include "fh_green.cuh"

// radix-32 butterfly from two radix-16 halves (decimation in time), natural-order output
template <bool INV>
struct Bfly<32, INV> {
    static __device__ __forceinline__ void run(cplx* v) {
        // cos / sin (k*pi/16), k = 0..15
        const double cs[16] = {1.0,
                               0.98078528040323044913,
                               0.92387953251128675613,
                               0.83146961230254523708,
                               0.70710678118654752440,
                               0.55557023301960222474,
                               0.38268343236508977173,
                               0.19509032201612826785,
                               0.0,
                               -0.19509032201612826785,
                               -0.38268343236508977173,
                               -0.55557023301960222474,
                               -0.70710678118654752440,
                               -0.83146961230254523708,
                               -0.92387953251128675613,
                               -0.98078528040323044913};
        const double sn[16] = {0.0,
                               0.19509032201612826785,
                               0.38268343236508977173,
                               0.55557023301960222474,
                               0.70710678118654752440,
                               0.83146961230254523708,
                               0.92387953251128675613,
                               0.98078528040323044913,
                               1.0,
                               0.98078528040323044913,
                               0.92387953251128675613,
                               0.83146961230254523708,
                               0.70710678118654752440,
                               0.55557023301960222474,
                               0.38268343236508977173,
                               0.19509032201612826785};
        cplx e[16], o[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            e[k] = v[2 * k];
            o[k] = v[2 * k + 1];
        }
        Bfly<16, INV>::run(e);
        Bfly<16, INV>::run(o);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            cplx t;
            if (k == 0)
                t = o[0];
            else if (k == 8)
                t = mul_mi<INV>(o[8]);
            else
                t = cmul(o[k], make_double2(cs[k], INV ? sn[k] : -sn[k]));
            v[k] = cadd(e[k], t);
            v[k + 16] = csub(e[k], t);
        }
    }
};

// where a tile lives: element (c, i0, ii) of the y-slab spectrum is data[c*cstride + row(i0) + ii], row(i0) =
// rowoff[i0] (exchange-buffer blocks, fh_ga_slab_*) or i0*rstride (natural layout); a spectrum row of the buffer
// holds `spitch` columns which are the global columns kcol0 .. kcol0+spitch-1 (k2-blocks of the slab pipeline)
struct Mid2Map {
    const int64_t* rowoff;
    int64_t rstride, cstride;
    int spitch, kcol0;
    int ntiles, tpr, col0;  // tiles walk tpr 8-column tiles per buffer row starting at buffer column col0
    // push mode of the slab pipeline (fh_slab2.cu): results of row i0 go to dout[c*cstride_out + rowoff_out[i0] + ii]
    cplx* dout;
    const int64_t* rowoff_out;
    int64_t cstride_out;
};

template <int N, int KIND>
struct Mid2Cfg {
    static constexpr int D = (KIND == FH_GREEN_SCALAR) ? 3 : 6;
    static constexpr int RA = N / 16;
    static constexpr int NT = RA * 16;                   // G^ half-tasks (q, p, t): 256 threads for N = 256
    static constexpr int NT1 = D * 128;                  // F1 / I1 tasks (c, j, t)
    static constexpr int ROUNDS = (NT1 + NT - 1) / NT;
    static constexpr size_t SMEM = (size_t)D * N * 8 * sizeof(cplx);
};

// PREF = number of F1 rounds whose loads are prefetched across tiles (issued during I1 of the previous tile);
// the remaining rounds load at the top of F1, one round ahead of the butterfly that consumes them
template <int N, int KIND, int MINB, int PREF>
__global__ void __launch_bounds__((Mid2Cfg<N, KIND>::NT), MINB)
    k_mid2(cplx* __restrict__ data, const cplx* __restrict__ tw, const GreenDesc g, const Mid2Map m, const int nh) {
    using Cfg = Mid2Cfg<N, KIND>;
    constexpr int D = Cfg::D, RA = Cfg::RA, NT = Cfg::NT, ROUNDS = Cfg::ROUNDS;
    extern __shared__ __align__(16) unsigned char fh_smem_raw[];
    cplx* sm = reinterpret_cast<cplx*>(fh_smem_raw);  // [D][RA][16][8]
    const int tid = threadIdx.x;
    const int t = tid & 7;
    // F1 / I1 role: task id = tid + rd*NT -> (c = id >> 7, j = (id >> 3) & 15, t)
    const int j = (tid >> 3) & 15;
    // G^ role: half-task (q, p, t)
    // (p is uniform over a warp, so the two variants of the load below do not diverge)
    const int p = (tid >> 5) & 1, q = ((tid >> 6) << 2) | ((tid >> 3) & 3);
    // cos / sin (k*pi/8), k = 0..7: w_16^k = cs - i sn
    constexpr double cs16[8] = {1.0, 0.92387953251128675613, 0.70710678118654752440, 0.38268343236508977173,
                            0.0, -0.38268343236508977173, -0.70710678118654752440, -0.92387953251128675613};
    constexpr double sn16[8] = {0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128675613,
                            1.0, 0.92387953251128675613, 0.70710678118654752440, 0.38268343236508977173};

    auto tile_ii = [&](int tile) -> int64_t {
        const int rowi = tile / m.tpr;
        return (int64_t)rowi * m.spitch + m.col0 + (tile - rowi * m.tpr) * 8;
    };
    auto row_off = [&](int i0) -> int64_t { return m.rowoff ? m.rowoff[i0] : (int64_t)i0 * m.rstride; };
    auto comp_of = [&](int rd) -> int { return (tid + rd * NT) >> 7; };
    auto load_round = [&](cplx (&v)[ROUNDS][RA], int rd, int64_t ii) {
        const int c = comp_of(rd);
        if (c < D) {
            const cplx* gp = data + (int64_t)c * m.cstride + ii + t;
#pragma unroll
            for (int r = 0; r < RA; ++r) v[rd][r] = gp[row_off(j + 16 * r)];
        }
    };
    // w_16^(-j) of this thread's I1 tasks = conj(tw[j * N/16])
    const cplx wj0 = __ldg(&tw[j * RA]);
    const cplx wj = make_double2(wj0.x, -wj0.y);

    cplx v[ROUNDS][RA];
    int tile = blockIdx.x;
    constexpr int PR = PREF < ROUNDS ? PREF : ROUNDS;
    if (tile < m.ntiles) {
#pragma unroll
        for (int rd = 0; rd < PR; ++rd) load_round(v, rd, tile_ii(tile));
    }
    for (; tile < m.ntiles; tile += gridDim.x) {
        const int rowi = tile / m.tpr;
        const int bcol = m.col0 + (tile - rowi * m.tpr) * 8;
        const int64_t ii = (int64_t)rowi * m.spitch + bcol;
        // ---- F1: radix RA over r, twiddle w_N^(j q) -> slot j
        if (PR < ROUNDS) load_round(v, PR, ii);
#pragma unroll
        for (int rd = 0; rd < ROUNDS; ++rd) {
            const int c = comp_of(rd);
            if (rd + 1 >= PR + 1 && rd + 1 < ROUNDS) load_round(v, rd + 1, ii);
            if (c < D) {
                cplx* s1 = sm + (c * RA * 16 + j) * 8 + t;  // + q*128
                Bfly<RA, false>::run(v[rd]);
#pragma unroll
                for (int qq = 1; qq < RA; ++qq) v[rd][qq] = cmul(v[rd][qq], __ldg(&tw[qq * j]));
#pragma unroll
                for (int qq = 0; qq < RA; ++qq) s1[qq * 128] = v[rd][qq];
            }
        }
        __syncthreads();
        // ---- last forward stage (radix 2 x 8 over j), G^ on k0 = q + RA (2a + p), first inverse stage
        {
            cplx w[D][8];
            cplx* s2 = sm + q * 128 + t;
            if (p == 0) {
#pragma unroll
                for (int cc = 0; cc < D; ++cc) {
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj)
                        w[cc][jj] = cadd(s2[cc * RA * 128 + jj * 8], s2[cc * RA * 128 + (jj + 8) * 8]);
                    asm volatile("" ::: "memory");  // 16 shared-memory loads in flight per thread, not 96
                }
            } else {
#pragma unroll
                for (int cc = 0; cc < D; ++cc) {
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const cplx dlt = csub(s2[cc * RA * 128 + jj * 8], s2[cc * RA * 128 + (jj + 8) * 8]);
                        w[cc][jj] = (jj == 0) ? dlt
                                              : (jj == 4) ? mul_mi<false>(dlt) : cmul(dlt, make_double2(cs16[jj], -sn16[jj]));
                    }
                    asm volatile("" ::: "memory");
                }
            }
#pragma unroll
            for (int cc = 0; cc < D; ++cc) Bfly<8, false>::run(w[cc]);
            int k[3];
            k[1] = fh_freq(rowi + g.ioff1, g.N[1]);
            const int i2 = bcol + t + m.kcol0;
            k[2] = fh_freq(i2, g.N[2]);
            const bool valid = i2 < nh;
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                k[0] = fh_freq(q + RA * (2 * a + p), N);
                cplx e[D];
#pragma unroll
                for (int cc = 0; cc < D; ++cc) e[cc] = w[cc][a];
                if (valid) {
                    green_apply<KIND, 3>(g, k, e);
                } else {
#pragma unroll
                    for (int cc = 0; cc < D; ++cc) e[cc] = make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int cc = 0; cc < D; ++cc) w[cc][a] = e[cc];
            }
#pragma unroll
            for (int cc = 0; cc < D; ++cc) Bfly<8, true>::run(w[cc]);
            __syncthreads();  // both half-tasks of (q, t) have read slots 0..15 before either overwrites them
#pragma unroll
            for (int cc = 0; cc < D; ++cc)
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) s2[cc * RA * 128 + (jj + 8 * p) * 8] = w[cc][jj];
        }
        __syncthreads();
        // ---- I1: y'_j = z'_0 + w_16^(-j) z'_1, conj twiddle, inverse radix RA, store.  The next tile's loads are
        // issued first and stay in flight in registers (there is no room for a second tile in shared memory)
        const int next = tile + gridDim.x;
        const int64_t iin = (next < m.ntiles) ? tile_ii(next) : 0;
#pragma unroll
        for (int rd = 0; rd < ROUNDS; ++rd) {
            const int c = comp_of(rd);
            if (c < D) {
                const cplx* s1 = sm + (c * RA * 16 + (j & 7)) * 8 + t;
                cplx u[RA];
#pragma unroll
                for (int qq = 0; qq < RA; ++qq) {
                    const cplx z0 = s1[qq * 128], z1 = s1[qq * 128 + 64];
                    u[qq] = cadd(z0, cmul(z1, wj));
                }
                if (rd < PR && next < m.ntiles) {
                    const cplx* gn = data + (int64_t)c * m.cstride + iin + t;
#pragma unroll
                    for (int r = 0; r < RA; ++r) v[rd][r] = gn[row_off(j + 16 * r)];
                }
#pragma unroll
                for (int qq = 1; qq < RA; ++qq) {
                    const cplx wq = __ldg(&tw[qq * j]);
                    u[qq] = cmul(u[qq], make_double2(wq.x, -wq.y));
                }
                Bfly<RA, true>::run(u);
                if (m.dout) {
                    cplx* gq = m.dout + (int64_t)c * m.cstride_out + ii + t;
#pragma unroll
                    for (int r = 0; r < RA; ++r) gq[m.rowoff_out[j + 16 * r]] = u[r];
                } else {
                    cplx* gp = data + (int64_t)c * m.cstride + ii + t;
#pragma unroll
                    for (int r = 0; r < RA; ++r) gp[row_off(j + 16 * r)] = u[r];
                }
            }
        }
        __syncthreads();  // the next F1 writes slots that the partner task (j +- 8) of the same (c, t) reads in I1
    }
}
