// K4 - fused back substitution and result recovery per (subset, load case): U_m = D L^-T D_piv^-1 y,
// scatter to DOF order, compliance, fixity reactions R_f = K_fm U_m - P_f, element end forces
// eR = -R12 (K_G,e U_e), max nodal translation/rotation and the success flag.
//
// Reference: numpy_stiffness.py:353-359 (solve, unscale, compliance), :366-372 (reactions),
// :386-391 (element reactions, sign flipped: quirk Q4), :438-441 / :457-464 (criterion:
// max_v ||u_v[0:3]||_2 < trans_tol, strict, NaN -> False).  One CTA per subset.
#include "cm_frag.cuh"

namespace {

constexpr int K4_THREADS = 256;
constexpr int K4_WARPS = K4_THREADS / 32;

__device__ __forceinline__ bool elem_on(const uint32_t *mask, int e) { return (mask[e >> 5] >> (e & 31)) & 1u; }

// NB right-hand sides at once.  cs[j][ks] (valid on lanes 0..3, column 4*ks + lane) =
// sum_r T[r][4*ks + lane] * w_j[r] for a tile in fragment-major layout; w4[j] holds
// w_j[i*8 + lane/4], i = 0..3.  One 32-byte load per lane and k-step, shared by the NB vectors.
template <int NB>
__device__ __forceinline__ void tile_tmatvec(const double *__restrict__ T, const double (&w4)[NB][4], double (&cs)[NB][8], int lane) {
    const double *p = T + lane * 4;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        double v[4];
        ldg4(v, p + q * 128);
#pragma unroll
        for (int j = 0; j < NB; ++j) cs[j][q] = v[0] * w4[j][0] + v[1] * w4[j][1] + v[2] * w4[j][2] + v[3] * w4[j][3];
    }
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            cs[j][q] += __shfl_xor_sync(0xffffffffu, cs[j][q], 4);
            cs[j][q] += __shfl_xor_sync(0xffffffffu, cs[j][q], 8);
            cs[j][q] += __shfl_xor_sync(0xffffffffu, cs[j][q], 16);
        }
}

// x_j = M^T w_j for a lower-triangular M_J tile in the cm_mtile_off layout: cs[j][2cj + e] (valid on
// lanes 0..3) = column 8cj + 2*lane + e.  Only the stored blocks (row block >= column block) are read.
template <int NB>
__device__ __forceinline__ void mtile_tmatvec(const double *__restrict__ M, const double (&w4)[NB][4], double (&cs)[NB][8], int lane) {
    const double *p = M + lane * 2;
#pragma unroll
    for (int cj = 0; cj < 4; ++cj) {
        double sx[NB], sy[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) { sx[j] = 0.0; sy[j] = 0.0; }
#pragma unroll
        for (int i = cj; i < 4; ++i) {
            const double2 b = ldg2(p + i * 256 + cj * 64);
#pragma unroll
            for (int j = 0; j < NB; ++j) { sx[j] += b.x * w4[j][i]; sy[j] += b.y * w4[j][i]; }
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) { cs[j][2 * cj] = sx[j]; cs[j][2 * cj + 1] = sy[j]; }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            cs[j][q] += __shfl_xor_sync(0xffffffffu, cs[j][q], 4);
            cs[j][q] += __shfl_xor_sync(0xffffffffu, cs[j][q], 8);
            cs[j][q] += __shfl_xor_sync(0xffffffffu, cs[j][q], 16);
        }
}
// where mtile_tmatvec leaves column q of its result: block 8*(q/2), lane pair 2*lane, parity q%2
__device__ __forceinline__ int mcol(int q, int lane) { return ((q >> 1) << 3) + 2 * lane + (q & 1); }

// NaN-propagating max (np.max semantics)
__device__ __forceinline__ double nanmax(double a, double b) {
    if (a != a) return a;
    if (b != b) return b;
    return a > b ? a : b;
}

__device__ double block_sum(double v, double *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < K4_WARPS; ++w) tot += s_red[w];
    return tot;
}

__device__ double block_nanmax(double v, double *s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = s_red[0];
    for (int w = 1; w < K4_WARPS; ++w) tot = nanmax(tot, s_red[w]);
    return tot;
}

// XS: the solution vectors of the load cases in progress live in shared memory (n_pad doubles each).
// The back substitution is a chain of one step per tile row, and every step reads and writes a few
// entries of x: from shared memory the chain step is one DRAM access (the tile) instead of five
// dependent global round trips.  NB load cases share one pass over L (every tile is read once for
// the NB right-hand sides); the recovery below then runs per load case.
template <bool XS, int NB>
__global__ void __launch_bounds__(K4_THREADS, NB == 1 ? 4 : (NB == 2 ? 2 : 1))      // 64 / 128 / 255 registers
k4_solve_recover(cm_model_dev m, cm_ws_layout L, char *ws, const uint32_t *masks, const int *geom, long long first,
                 cm_out_dev out) {
    static_assert(XS || NB == 1, "several right-hand sides per pass need the shared-memory vectors");
    __shared__ double s_red[K4_WARPS];
    extern __shared__ __align__(16) double s_x[];            // XS: [NB][n_pad] solutions, then [n_pad] pivots; then [mask_words] ints

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long long gs = first + blockIdx.x;
    char *slab = ws + (size_t)blockIdx.x * L.per_subset;
    const int *info = reinterpret_cast<const int *>(slab + L.off_info);
    const int status = info[CM_INFO_STATUS];
    const int NT = info[CM_INFO_NT];
    const double *tiles = reinterpret_cast<const double *>(slab + L.off_tiles);
    const int *kfirst = reinterpret_cast<const int *>(slab + L.off_kfirst);
    const int *sidx = reinterpret_cast<const int *>(slab + L.off_sidx);
    const double *dinv = reinterpret_cast<const double *>(slab + L.off_dinv);
    const double *dpiv = reinterpret_cast<const double *>(slab + L.off_dpiv);
    const double *minv = reinterpret_cast<const double *>(slab + L.off_minv);
    double *X = reinterpret_cast<double *>(slab + L.off_x);
    const double *P = reinterpret_cast<const double *>(slab + L.off_p);
    const uint32_t *mask = masks + (size_t)gs * m.mask_words;
    const size_t gv = geom ? (size_t)min(max(geom[gs], 0), m.n_geom - 1) : 0;      // geometry variant of this subset
    const double *Kg = m.Kg + gv * m.nE * 144;
    const double *Kt = m.Kt + gv * m.nE * 144;        // the entries the assembly summed (|k| <= 1e-14 dropped)
    const double *R3g = m.R3 + gv * m.nE * 9;
    const int bt = m.band_tiles, nV = m.nV, nE = m.nE, n_lc = m.n_lc;
    const int zero_bits = info[CM_INFO_ZERO_LOAD];
    const int n_pad = L.n_pad;
    // prefix-sharing group: tile rows < trows of L, M, the pivots and the forward solutions are the trunk subset's
    const int trows = info[CM_INFO_TROWS];
    const char *tslab = ws + (size_t)info[CM_INFO_TRUNK] * L.per_subset;
    const double *ttiles = reinterpret_cast<const double *>(tslab + L.off_tiles);
    const double *tminv = reinterpret_cast<const double *>(tslab + L.off_minv);
    const double *tdpiv = reinterpret_cast<const double *>(tslab + L.off_dpiv);
    const double *tX = reinterpret_cast<const double *>(tslab + L.off_x);

    if (out.min_pivot && t == 0)
        out.min_pivot[gs] = (status == CM_ST_OK || status == CM_ST_SINGULAR) ? dpiv[n_pad] : nan("");
    // compact element rows: rank of an existing element among the subset's elements = popcount of the mask bits below it
    int *s_wpre = reinterpret_cast<int *>(s_x + (XS ? (size_t)(NB + 1) * n_pad : 0));      // [mask_words] exclusive prefix of the word popcounts
    long long crow0 = 0, ccnt = 0;
    if (out.eR_rows) {
        crow0 = out.eR_row_offsets[gs] - out.row_bias;
        ccnt = out.eR_row_offsets[gs + 1] - out.eR_row_offsets[gs];
        if (t == 0) {
            int run = 0;
            for (int w = 0; w < m.mask_words; ++w) { s_wpre[w] = run; run += __popc(mask[w]); }
        }
        __syncthreads();
    }

    for (int lc0 = 0; lc0 < n_lc; lc0 += NB) {
      // ---- back substitution of the load cases lc0 .. lc0+NB-1 in one pass over L (status OK only;
      // a load case with a zero load vector rides along as x = 0)
      const double *dpl = XS ? s_x + NB * n_pad : dpiv;       // pivots next to x (copied below)
      if (status == CM_ST_OK) {
          bool any = false;
#pragma unroll
          for (int j = 0; j < NB; ++j) {
              const int lc = lc0 + j;
              const bool live = lc < n_lc && !((zero_bits >> lc) & 1);
              any |= live;
              double *Xg = X + (size_t)(lc < n_lc ? lc : 0) * n_pad;
              const double *Xt = tX + (size_t)(lc < n_lc ? lc : 0) * n_pad;
              double *Xl = XS ? s_x + j * n_pad : Xg;
              for (int r = t; r < NT * CM_TILE; r += K4_THREADS) {
                  const bool shared_row = XS && r < trows * CM_TILE;       // (sharing needs the shared-memory vectors: host checks)
                  const double d = shared_row ? tdpiv[r] : dpiv[r];
                  if (XS && j == 0) s_x[NB * n_pad + r] = d;
                  Xl[r] = live ? (shared_row ? Xt[r] : Xg[r]) / d : 0.0;
              }
          }
          __syncthreads();
          if (any) {
            // L^T x = D^-1 y, row-oriented: once x_I is known every tile (I,J) of block row I
            // (contiguous in memory) subtracts L(I,J)^T x_I from the pending sum of block J; the warp
            // that owns tile (I,I-1) then finishes x_{I-1} = M^T (d .* s) with M = D^-1 L_JJ^-1 left by
            // the factorisation.  One barrier per row.
            double *Xb = XS ? s_x : X + (size_t)lc0 * n_pad;   // vector j at Xb + j*n_pad (NB == 1 without XS)
            const int g = lane >> 2;
            auto finish_block = [&](int J) {                   // x_J = M_J^T (d .* s_J), one warp
                double w4[NB][4], cs[NB][8];
#pragma unroll
                for (int j = 0; j < NB; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) w4[j][i] = dpl[J * CM_TILE + i * 8 + g] * Xb[j * n_pad + J * CM_TILE + i * 8 + g];
                mtile_tmatvec<NB>((J < trows ? tminv : minv) + (size_t)J * CM_TILE_ELEMS, w4, cs, lane);
                __syncwarp();
                if (lane < 4)
#pragma unroll
                    for (int j = 0; j < NB; ++j)
#pragma unroll
                        for (int q = 0; q < 8; ++q) Xb[j * n_pad + J * CM_TILE + mcol(q, lane)] = cs[j][q];
            };
            if (warp == 0) finish_block(NT - 1);
            for (int I = NT - 1; I > 0; --I) {
                // M_{I-1} is needed at the end of this row's chain: start its DRAM read now
                if (warp == 0) {
                    const double *Mn = (I - 1 < trows ? tminv : minv) + (size_t)(I - 1) * CM_TILE_ELEMS;
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(Mn + lane * 16));
                    asm volatile("prefetch.global.L2 [%0];" :: "l"(Mn + 512 + lane * 16));
                }
                // ... and so are the tiles of the next row: request this warp's share one row ahead
                if (I >= 2) {
                    const int kfn = kfirst[I - 1];
                    for (int J = I - 2 - warp; J >= kfn; J -= K4_WARPS) {
                        const double *Tn = cm_tile_ptr(I - 1 < trows ? ttiles : tiles, bt, I - 1, J) + lane * 16;
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(Tn));
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(Tn + 512));
                    }
                }
                __syncthreads();
                const int kf = kfirst[I];
                double x4[NB][4];
#pragma unroll
                for (int j = 0; j < NB; ++j)
#pragma unroll
                    for (int i = 0; i < 4; ++i) x4[j][i] = Xb[j * n_pad + I * CM_TILE + i * 8 + g];
                for (int J = I - 1 - warp; J >= kf; J -= K4_WARPS) {
                    double cs[NB][8];
                    tile_tmatvec<NB>(cm_tile_ptr(I < trows ? ttiles : tiles, bt, I, J), x4, cs, lane);
                    if (lane < 4)
#pragma unroll
                        for (int j = 0; j < NB; ++j)
#pragma unroll
                            for (int q = 0; q < 8; ++q) Xb[j * n_pad + J * CM_TILE + q * 4 + lane] -= cs[j][q];
                    if (J == I - 1) {
                        __syncwarp();
                        finish_block(J);
                    }
                }
                // row I has no tile left of the diagonal: block I-1 is complete without it
                if (kf > I - 1 && warp == 0) finish_block(I - 1);
            }
            __syncthreads();
          }
      }
      for (int lc = lc0; lc < lc0 + NB && lc < n_lc; ++lc) {
        const size_t ol = (size_t)gs * n_lc + lc;
        double *Uo = out.U ? out.U + ol * 6 * nV : nullptr;
        double *Ro = out.Rf ? out.Rf + ol * 6 * nV : nullptr;
        double *Eo = out.eR ? out.eR + ol * (size_t)nE * 12 : nullptr;
        double *Rs = out.Rf_sup ? out.Rf_sup + ol * (size_t)m.nS * 6 : nullptr;
        double *Ec = out.eR_rows ? out.eR_rows + ((size_t)crow0 * n_lc + (size_t)lc * ccnt) * 12 : nullptr;
        const bool zero_load = (zero_bits >> lc) & 1;
        // a zero load vector skips the factorisation in the reference (numpy_stiffness.py:267), so a
        // singular matrix does not fail that load case
        if (status != CM_ST_OK && !(status == CM_ST_SINGULAR && zero_load)) {
            // no stored result in the reference (has_stored_result() == False)
            if (Uo) for (int i = t; i < 6 * nV; i += K4_THREADS) Uo[i] = 0.0;
            if (Ro) for (int i = t; i < 6 * nV; i += K4_THREADS) Ro[i] = 0.0;
            if (Eo) for (int i = t; i < 12 * nE; i += K4_THREADS) Eo[i] = 0.0;
            if (Rs) for (int i = t; i < 6 * m.nS; i += K4_THREADS) Rs[i] = 0.0;
            if (Ec) for (long long i = t; i < 12 * ccnt; i += K4_THREADS) Ec[i] = 0.0;
            if (out.sigma_max) for (int i = t; i < nE; i += K4_THREADS) out.sigma_max[ol * (size_t)nE + i] = 0.0;
            if (t == 0) {
                if (out.max_abs_sigma) out.max_abs_sigma[ol] = nan("");
                if (out.flag) out.flag[ol] = 0;
                if (out.status) out.status[ol] = (unsigned char)status;
                if (out.compliance) out.compliance[ol] = 0.0;
                if (out.max_trans) out.max_trans[ol] = nan("");
                if (out.max_rot) out.max_rot[ol] = nan("");
            }
            continue;
        }
        double *Xl = XS ? s_x + (lc - lc0) * n_pad : X + (size_t)lc * n_pad;
        if (status != CM_ST_OK) {            // singular matrix, zero load: the solve is skipped, U = 0
            for (int r = t; r < NT * CM_TILE; r += K4_THREADS) Xl[r] = 0.0;
            __syncthreads();
        }

        // ---- displacements in DOF order (U = D x on free DOFs, 0 elsewhere), compliance
        auto u_at = [&](int d) -> double {
            int r = sidx[d];
            return r >= 0 ? dinv[r] * Xl[r] : 0.0;
        };
        const double *Pl = P + (size_t)lc * 6 * nV;
        double comp = 0.0;
        for (int d = t; d < 6 * nV; d += K4_THREADS) {
            double u = u_at(d);
            if (Uo) Uo[d] = u;
            if (sidx[d] >= 0) comp += u * Pl[d];
        }
        comp = 0.5 * block_sum(comp, s_red);

        // ---- max nodal translation / rotation over existing nodes (row = 6 consecutive DOFs)
        double mt = -1.0, mr = -1.0;
        for (int v = t; v < nV; v += K4_THREADS) {
            bool ex = false;
            for (int p = m.ne_ptr[v]; p < m.ne_ptr[v + 1]; ++p) ex |= elem_on(mask, m.ne_elem[p]);
            if (!ex) continue;
            double u0 = u_at(6 * v), u1 = u_at(6 * v + 1), u2 = u_at(6 * v + 2);
            double r0 = u_at(6 * v + 3), r1 = u_at(6 * v + 4), r2 = u_at(6 * v + 5);
            // np.linalg.norm(axis=1): sqrt(x0^2 + x1^2 + x2^2), no fused multiply-add
            double nt = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(u0, u0), __dmul_rn(u1, u1)), __dmul_rn(u2, u2)));
            double nr = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(r0, r0), __dmul_rn(r1, r1)), __dmul_rn(r2, r2)));
            mt = nanmax(mt, nt);
            mr = nanmax(mr, nr);
        }
        mt = block_nanmax(mt, s_red);
        mr = block_nanmax(mr, s_red);

        // ---- fixity reactions at the fixed DOFs of existing support nodes
        if (Ro || Rs) {
            if (Ro) {
                for (int d = t; d < 6 * nV; d += K4_THREADS) Ro[d] = 0.0;
                __syncthreads();
            }
            for (int idx = t; idx < 6 * m.nS; idx += K4_THREADS) {
                int s = idx / 6, i = idx - 6 * s;
                int v = m.sup_node[s];
                if (!m.sup_cond[6 * s + i]) { if (Rs) Rs[idx] = 0.0; continue; }
                bool ex = false;
                double acc = 0.0;
                for (int p = m.ne_ptr[v]; p < m.ne_ptr[v + 1]; ++p) {
                    int e = m.ne_elem[p];
                    if (!elem_on(mask, e)) continue;
                    ex = true;
                    const double *krow = Kt + (size_t)e * 144 + (6 * m.ne_end[p] + i) * 12;
                    int n0 = m.conn[2 * e], n1 = m.conn[2 * e + 1];
#pragma unroll
                    for (int j = 0; j < 6; ++j) acc += krow[j] * u_at(6 * n0 + j);
#pragma unroll
                    for (int j = 0; j < 6; ++j) acc += krow[6 + j] * u_at(6 * n1 + j);
                }
                const double rv = ex ? acc - (zero_load ? 0.0 : Pl[6 * v + i]) : 0.0;   // P_f stays 0 when the solve is skipped (:260-270)
                if (ex && Ro) Ro[6 * v + i] = rv;
                if (Rs) Rs[idx] = rv;
            }
        }

        // ---- element end forces in local axes, four threads per element (one per 3-row block:
        // F end 0, M end 0, F end 1, M end 1), and from them the reference's stress measure
        // (stiffness_checker.py:330-365): sigma = N/A, sign(sigma) * (|sigma| + My/W + Mz/W) with
        // r = sqrt(A/pi), W = Iy/r for BOTH axes, My/Mz = larger end moment in magnitude
        double *So = out.sigma_max ? out.sigma_max + ol * (size_t)nE : nullptr;
        double smax = -1.0;
        if (Eo || Ec || So || out.max_abs_sigma) {
            for (int i0 = 0; i0 < 4 * nE; i0 += K4_THREADS) {       // uniform trip count: shuffles below
                const int idx = i0 + t;
                const bool valid = idx < 4 * nE;
                const int e = valid ? idx >> 2 : 0, pblk = idx & 3;
                const bool on = valid && elem_on(mask, e);
                double d3[3] = {0.0, 0.0, 0.0};
                if (on) {
                    int n0 = m.conn[2 * e], n1 = m.conn[2 * e + 1];
                    double ue[12];
#pragma unroll
                    for (int j = 0; j < 6; ++j) { ue[j] = u_at(6 * n0 + j); ue[6 + j] = u_at(6 * n1 + j); }
                    double tv[3];
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const double *krow = Kg + (size_t)e * 144 + (3 * pblk + a) * 12;
                        double s = 0.0;
#pragma unroll
                        for (int j = 0; j < 12; ++j) s += krow[j] * ue[j];
                        tv[a] = s;
                    }
                    const double *R3 = R3g + (size_t)e * 9;
#pragma unroll
                    for (int a = 0; a < 3; ++a)
                        d3[a] = -(R3[3 * a] * tv[0] + R3[3 * a + 1] * tv[1] + R3[3 * a + 2] * tv[2]);
                }
                if (Eo && valid) {
                    double *dst = Eo + (size_t)e * 12 + 3 * pblk;
                    dst[0] = d3[0]; dst[1] = d3[1]; dst[2] = d3[2];
                }
                if (Ec && on) {
                    const int rank = s_wpre[e >> 5] + __popc(mask[e >> 5] & ((1u << (e & 31)) - 1u));
                    double *dst = Ec + (size_t)rank * 12 + 3 * pblk;
                    dst[0] = d3[0]; dst[1] = d3[1]; dst[2] = d3[2];
                }
                if (So || out.max_abs_sigma) {
                    const int l0 = lane & ~3;
                    const double N = __shfl_sync(0xffffffffu, d3[0], l0);
                    const double My0 = __shfl_sync(0xffffffffu, d3[1], l0 + 1), My1 = __shfl_sync(0xffffffffu, d3[1], l0 + 3);
                    const double Mz0 = __shfl_sync(0xffffffffu, d3[2], l0 + 1), Mz1 = __shfl_sync(0xffffffffu, d3[2], l0 + 3);
                    if (valid && pblk == 0) {
                        double sg = 0.0;
                        if (on) {
                            const double A = m.props[(size_t)e * 7], Iy = m.props[(size_t)e * 7 + 2];
                            const double My = fmax(fabs(My0), fabs(My1)), Mz = fmax(fabs(Mz0), fabs(Mz1));
                            const double rr = sqrt(__ddiv_rn(A, 3.141592653589793));
                            const double W = __ddiv_rn(Iy, rr);
                            const double sigma = __ddiv_rn(N, A);
                            const double mag = __dadd_rn(__dadd_rn(fabs(sigma), __ddiv_rn(My, W)), __ddiv_rn(Mz, W));
                            sg = sigma > 0.0 ? mag : -mag;
                            smax = nanmax(smax, fabs(sg));
                        }
                        if (So) So[e] = sg;
                    }
                }
            }
        }
        if (out.max_abs_sigma) smax = block_nanmax(smax, s_red);

        if (t == 0) {
            if (out.flag) out.flag[ol] = (mt < m.trans_tol) ? 1 : 0;
            if (out.status) out.status[ol] = zero_load ? CM_ST_ZERO_LOAD : CM_ST_OK;
            if (out.compliance) out.compliance[ol] = comp;
            if (out.max_trans) out.max_trans[ol] = mt;
            if (out.max_rot) out.max_rot[ol] = mr;
            if (out.max_abs_sigma) out.max_abs_sigma[ol] = smax < 0.0 ? 0.0 : smax;
        }
        __syncthreads();
      }
    }
}

}  // namespace

template <int NB>
static cudaError_t launch_k4_xs(const cm_model_dev &m, const cm_ws_layout &L, char *ws, const uint32_t *masks, const int *geom,
                                long long first, int count, const cm_out_dev &out, cudaStream_t st) {
    const size_t xs_bytes = (size_t)(NB + 1) * L.n_pad * sizeof(double) + (size_t)m.mask_words * sizeof(int);      // NB solutions and the pivots; word prefix of the mask
    static cm_smem_cache cache;
    cudaError_t e = cm_ensure_smem(k4_solve_recover<true, NB>, xs_bytes, cache, m.device);
    if (e != cudaSuccess) return e;
    k4_solve_recover<true, NB><<<count, K4_THREADS, xs_bytes, st>>>(m, L, ws, masks, geom, first, out);
    return cudaGetLastError();
}

// Load cases per pass over L: as many as keep the shared-memory vectors within 64 KB per CTA (the
// kernel is HBM-bound through many resident CTAs), at most 4.  CM_K4_NB overrides (timing).
cudaError_t cm_launch_solve_recover(const cm_model_dev &m, const cm_ws_layout &L, char *ws,
                                    const uint32_t *masks, const int *geom, long long first, int count,
                                    const cm_out_dev &out, cudaStream_t st) {
    const size_t vec = (size_t)L.n_pad * sizeof(double);
    static const int nb_env = getenv("CM_K4_NB") ? atoi(getenv("CM_K4_NB")) : 0;
    int nb = 0;
    if (2 * vec <= 100 * 1024) {          // (two CTAs per SM at the upper end; beyond that the vectors stay in global memory)
        nb = 1;
        if (m.n_lc >= 2 && 3 * vec <= 64 * 1024) nb = 2;
        if (m.n_lc >= 3 && 5 * vec <= 64 * 1024) nb = 4;
        if (nb_env == 1 || (nb_env == 2 && 3 * vec <= 200 * 1024) || (nb_env == 4 && 5 * vec <= 200 * 1024)) nb = nb_env;
    }
    if (nb == 4) return launch_k4_xs<4>(m, L, ws, masks, geom, first, count, out, st);
    if (nb == 2) return launch_k4_xs<2>(m, L, ws, masks, geom, first, count, out, st);
    if (nb == 1) return launch_k4_xs<1>(m, L, ws, masks, geom, first, count, out, st);
    k4_solve_recover<false, 1><<<count, K4_THREADS, (size_t)m.mask_words * sizeof(int), st>>>(m, L, ws, masks, geom, first, out);
    return cudaGetLastError();
}
