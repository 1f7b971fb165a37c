// K3 (default) - batched fp64 LDL^T of the Jacobi-scaled K_mm in block-band tile storage on the
// FP64 tensor cores, with pivot-based mechanism detection and the forward substitution fused in.
//
// Replaces the reference's SuperLU factorisation + definiteness test + first triangular solve
// (numpy_stiffness.py:299-336, :353).  The reference factors with diagonal pivots only
// (diag_pivot_thresh=0, SymmetricMode), raises on an exactly zero pivot (:301-319) and rejects any
// pivot <= -1e-14 (:326-336); the same rules are applied here to the pivots d_j of the LDL^T.
//
// One CTA (NW warps) per subset, left-looking by block column J, three barrier-separated phases:
//   L(J)  "late"   every tile (I,J) of the column receives its newest term  - L(I,J-1) D L(J,J-1)^T
//   D(J)  "diag"   one warp (rotating with J over the SM sub-partitions) factors the diagonal tile
//                  (LDL^T in registers, pivot rules) and forms M_J = D_J^-1 L_JJ^-1; meanwhile the
//                  other warps run the look-ahead - all older terms (K <= J-1) of the tiles of column
//                  J+1, one linear operand stream per tile - and the forward substitution of block
//                  row J-1
//   T(J)  "trsm"   L(I,J) = acc(I,J) M_J^T for the tiles below the diagonal - the triangular solve as
//                  one more tile product on the tensor cores (M_J is read from shared memory)
// Tiles are stored fragment-major (cm_tile_off), so accumulators and operands move with 32-byte
// vector loads/stores straight between global memory (L1/L2 resident window) and MMA registers.
// The serial diagonal factorisation is the latency the design hides: several small CTAs share an SM
// so that while one waits on its diagonal warp the others keep the FP64 tensor pipe busy.
#include "cm_frag.cuh"
#include <cstdlib>

namespace {

constexpr int MAX_ITEMS = 64;               // band_tiles covered by the look-ahead slot table

__device__ unsigned long long g_col_prof[16];   // instrumented build: cycles per phase

// 1/d to full double precision (not correctly rounded): hardware seed + two Newton steps.
// d is a pivot of a Jacobi-scaled matrix (or 1 after a rejected pivot), far from the fp64 range ends.
__device__ __forceinline__ double fast_rcp(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);      // seed: relative error <= 2^-23
    x = fma(x, e, x);                // 2^-46
    e = fma(-d, x, 1.0);
    return fma(x, e, x);             // below one ulp
}

// LDL^T of the lower triangle of the 32x32 tile T (global, tile layout) by one warp, lane r owning
// row r, followed by M = D^-1 L^-1 with lane j owning column j of L^-1.  Outputs: s_M (tile layout,
// shared), s_d (pivots, shared).  *s_fail is set when a pivot violates the reference's rules (exact
// zero raises in SuperLU: numpy_stiffness.py:301-319; <= -1e-14 is rejected: :326-336; NaN fails
// both); such a pivot is replaced by 1 so the arithmetic stays finite.
__device__ __noinline__ void diag_factor(const double *T, double *s_M, double *s_L, double *s_col,
                                         double *s_d, double *s_rinv, int lane, int *s_fail, double *s_minpiv) {
    double a[32];
    int fail = 0;
    double minpiv = *s_minpiv;
#pragma unroll
    for (int k = 0; k < 32; ++k) a[k] = T[cm_tile_off(lane, k)];
    double dmine = 1.0, rmine = 1.0;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        s_col[lane] = a[c];
        __syncwarp();
        double d = s_col[c];
        if (!(d > -CM_EPS) || d == 0.0) { fail = 1; d = 1.0; }
        minpiv = fmin(minpiv, d);
        const double rinv = fast_rcp(d);
        if (lane == c) { dmine = d; rmine = rinv; }
        const double l = a[c] * rinv;           // L[lane][c] for lane > c (unused otherwise)
        s_L[c * 32 + lane] = l;
        if (c < 31) {
            // column values two at a time (16-byte shared loads)
            if ((c + 1) & 1) a[c + 1] -= l * s_col[c + 1];
#pragma unroll
            for (int k = (c + 2) & ~1; k < 32; k += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(s_col + k);
                a[k] -= l * v.x;
                a[k + 1] -= l * v.y;
            }
        }
        __syncwarp();
    }
    s_d[lane] = dmine;
    s_rinv[lane] = rmine;
    __syncwarp();
    if (lane == 0) { *s_minpiv = minpiv; if (fail) *s_fail = 1; }
    // column `lane` of L^-1 by forward substitution on e_lane; M[n][lane] = Linv[n][lane] / d_n
    double x[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) x[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const double f = x[c];
        s_M[cm_tile_off(c, lane)] = f * s_rinv[c];
        if (c < 31) {
            if ((c + 1) & 1) x[c + 1] -= s_L[c * 32 + c + 1] * f;
#pragma unroll
            for (int k = (c + 2) & ~1; k < 32; k += 2) {
                const double2 v = *reinterpret_cast<const double2 *>(s_L + c * 32 + k);
                x[k] -= v.x * f;
                x[k + 1] -= v.y * f;
            }
        }
    }
    __syncwarp();
}

// Whole-tile tasks as separate functions: each gets its own register allocation (64 accumulator +
// 32 operand registers), so the MMA loops never spill whatever the caller keeps live.
template <bool SYRK>
__device__ __noinline__ void tile_update(double *T, const double *A, const double *B, const double *d,
                                         int nsteps, int lane) {
    TileAcc acc;
    acc_load(acc, T, lane);
    acc_update_stream<SYRK>(acc, A, B, d, nsteps, lane);
    acc_store(acc, T, lane);
}

__device__ __noinline__ void tile_trsm(double *T, const double *s_M, int lane) {
    TileAcc acc;
    acc_zero(acc);
    acc_mul_smem(acc, T, s_M, lane);
    acc_store(acc, T, lane);
}

// forward substitution of block row J for every load case (one warp):
//   y_J = D_J M_J (b_J - sum_K L(J,K) y_K),  M_J = D_J^-1 L_JJ^-1 read back from global memory
__device__ __noinline__ void forward_row(const double *tiles, const double *minv, const double *dpiv, double *X,
                                         double *s_r, int n_pad, int n_lc, int zero_bits, int bt, int J, int kfJ,
                                         int lane) {
    const int g = lane >> 2, tt = lane & 3;
    const double *Mj = minv + (size_t)J * CM_TILE_ELEMS + lane * 4;
    for (int lc = 0; lc < n_lc; ++lc) {
        if ((zero_bits >> lc) & 1) continue;
        double *Xl = X + (size_t)lc * n_pad;
        double part[4] = {0.0, 0.0, 0.0, 0.0};
        const double *Ljk = cm_tile_ptr(tiles, bt, J, kfJ) + lane * 4;
        const double *yk = Xl + kfJ * CM_TILE + tt;
        const int nst = (J - kfJ) * 8;
#pragma unroll 4
        for (int q = 0; q < nst; ++q) {
            double v[4];
            ldg4(v, Ljk + q * 128);
            const double y = yk[q * 4];
#pragma unroll
            for (int i = 0; i < 4; ++i) part[i] += v[i] * y;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            part[i] += __shfl_xor_sync(0xffffffffu, part[i], 1);
            part[i] += __shfl_xor_sync(0xffffffffu, part[i], 2);
        }
        if (tt == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) s_r[i * 8 + g] = Xl[J * CM_TILE + i * 8 + g] - part[i];
        }
        __syncwarp();
        double zp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            double v[4];
            ldg4(v, Mj + q * 128);
            const double rr = s_r[q * 4 + tt];
#pragma unroll
            for (int i = 0; i < 4; ++i) zp[i] += v[i] * rr;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            zp[i] += __shfl_xor_sync(0xffffffffu, zp[i], 1);
            zp[i] += __shfl_xor_sync(0xffffffffu, zp[i], 2);
        }
        if (tt == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) Xl[J * CM_TILE + i * 8 + g] = dpiv[J * CM_TILE + i * 8 + g] * zp[i];
        }
        __syncwarp();
    }
}

template <int NW, bool PROF, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB)
k3_factor_col(cm_model_dev m, cm_ws_layout L, char *ws, int dbg) {
    __shared__ __align__(32) double s_M[CM_TILE_ELEMS];     // M_J, tile layout
    __shared__ __align__(16) double s_L[CM_TILE_ELEMS];     // columns of L_JJ
    __shared__ __align__(16) double s_col[32], s_d[32], s_rinv[32], s_r[32];
    __shared__ int s_slot[MAX_ITEMS];                       // look-ahead item -> worker slot
    __shared__ int s_fail, s_fwd_slot;
    __shared__ double s_minpiv;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    char *slab = ws + (size_t)blockIdx.x * L.per_subset;
    int *info = reinterpret_cast<int *>(slab + L.off_info);
    if (info[CM_INFO_STATUS] != CM_ST_OK) return;
    const int NT = info[CM_INFO_NT];
    double *tiles = reinterpret_cast<double *>(slab + L.off_tiles);
    const int *kfirst = reinterpret_cast<const int *>(slab + L.off_kfirst);
    double *dpiv = reinterpret_cast<double *>(slab + L.off_dpiv);
    double *minv = reinterpret_cast<double *>(slab + L.off_minv);
    double *X = reinterpret_cast<double *>(slab + L.off_x);
    const int bt = m.band_tiles;

    if (t == 0) {
        s_fail = 0;
        s_minpiv = __longlong_as_double(0x7ff0000000000000LL);
        // static slot table of the look-ahead items of a full band: item i = tile (J+1+i, J+1) has
        // bt-2-i terms; longest first onto the least loaded of the NW-1 worker slots.  The forward
        // substitution (about one term's worth of work) goes to the least loaded slot at the end.
        int load[NW];
        const int nslots = (dbg & 4) ? 3 : NW - 1;
        for (int s = 0; s < NW - 1; ++s) load[s] = 0;
        for (int i = 0; i < bt && i < MAX_ITEMS; ++i) {
            int best = 0;
            for (int s = 1; s < nslots; ++s) if (load[s] < load[best]) best = s;
            load[best] += bt - 2 - i > 0 ? bt - 2 - i : 0;
            s_slot[i] = best;
        }
        int best = 0;
        for (int s = 1; s < nslots; ++s) if (load[s] < load[best]) best = s;
        s_fwd_slot = best;
    }
    __syncthreads();

    const int n_lc = m.n_lc;
    const int zero_bits = info[CM_INFO_ZERO_LOAD];
    long long tp = PROF ? clock64() : 0;
    long long pc[3] = {0, 0, 0};

    for (int J = 0; J < NT; ++J) {
        const int kfJ = kfirst[J];
        // ---- phase L(J): newest term (K = J-1) of every tile of column J
        if (J > 0) {
            const int cnt = min(bt, NT - J);
            for (int i = warp; i < cnt; i += NW) {
                const int I = J + i;
                if (max(kfirst[I], kfJ) > J - 1) continue;
                double *T = cm_tile_ptr(tiles, bt, I, J);
                if (i == 0)
                    tile_update<true>(T, cm_tile_ptr(tiles, bt, J, J - 1), nullptr, dpiv + (J - 1) * CM_TILE, 8, lane);
                else
                    tile_update<false>(T, cm_tile_ptr(tiles, bt, I, J - 1), cm_tile_ptr(tiles, bt, J, J - 1),
                                       dpiv + (J - 1) * CM_TILE, 8, lane);
            }
            __syncthreads();
        }
        if (PROF && t == 0) { long long n = clock64(); pc[0] += n - tp; tp = n; }

        // ---- phase D(J): diagonal factor (warp J % NW) | look-ahead + forward substitution (others)
        const int dwarp = (dbg & 1) ? 0 : J % NW;
        if (warp == dwarp) {
            long long t0 = PROF ? clock64() : 0;
            diag_factor(cm_tile_ptr(tiles, bt, J, J), s_M, s_L, s_col, s_d, s_rinv, lane, &s_fail, &s_minpiv);
            dpiv[J * CM_TILE + lane] = s_d[lane];
            {   // keep M_J for the forward and back substitutions
                double *dst = minv + (size_t)J * CM_TILE_ELEMS;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    double v[4];
                    lds4(v, s_M + q * 128 + lane * 4);
                    stg4(dst + q * 128 + lane * 4, v[0], v[1], v[2], v[3]);
                }
            }
            if (PROF && lane == 0) atomicAdd(&g_col_prof[3], (unsigned long long)(clock64() - t0));
        } else {
            const int slot = (warp - dwarp + NW) % NW - 1;      // 0 .. NW-2
            if (J + 1 < NT) {
                const int J1 = J + 1;
                const int kfJ1 = kfirst[J1];
                const int cnt = min(bt, NT - J1);
                for (int i = 0; i < cnt; ++i) {
                    if ((i < MAX_ITEMS ? s_slot[i] : i % (NW - 1)) != slot) continue;
                    const int I = J1 + i;
                    const int K0 = max(kfirst[I], kfJ1);
                    if (K0 > J - 1) continue;
                    double *T = cm_tile_ptr(tiles, bt, I, J1);
                    if (i == 0)
                        tile_update<true>(T, cm_tile_ptr(tiles, bt, J1, K0), nullptr, dpiv + K0 * CM_TILE, (J - K0) * 8, lane);
                    else
                        tile_update<false>(T, cm_tile_ptr(tiles, bt, I, K0), cm_tile_ptr(tiles, bt, J1, K0),
                                           dpiv + K0 * CM_TILE, (J - K0) * 8, lane);
                }
            }
            if (J > 0 && slot == ((dbg & 2) ? 0 : s_fwd_slot))
                forward_row(tiles, minv, dpiv, X, s_r, L.n_pad, n_lc, zero_bits, bt, J - 1, kfirst[J - 1], lane);
        }
        __syncthreads();
        if (PROF && t == 0) { long long n = clock64(); pc[1] += n - tp; tp = n; }
        if (s_fail) break;

        // ---- phase T(J): L(I,J) = acc(I,J) M_J^T for the tiles below the diagonal
        {
            const int cnt = min(bt - 1, NT - 1 - J);
            for (int i = warp; i < cnt; i += NW) {
                const int I = J + 1 + i;
                if (kfirst[I] > J) continue;
                tile_trsm(cm_tile_ptr(tiles, bt, I, J), s_M, lane);
            }
        }
        __syncthreads();
        if (PROF && t == 0) { long long n = clock64(); pc[2] += n - tp; tp = n; }
    }
    if (!s_fail && warp == 0)
        forward_row(tiles, minv, dpiv, X, s_r, L.n_pad, n_lc, zero_bits, bt, NT - 1, kfirst[NT - 1], lane);

    if (t == 0) {
        if (s_fail) info[CM_INFO_STATUS] = CM_ST_SINGULAR;
        dpiv[L.n_pad] = s_minpiv;      // slot after the pivots
        if (PROF) for (int i = 0; i < 3; ++i) atomicAdd(&g_col_prof[i], (unsigned long long)pc[i]);
    }
}

}  // namespace

cudaError_t cm_launch_factor_col(const cm_model_dev &m, const cm_ws_layout &L, char *ws, int count,
                                 int nwarps, int prof, cudaStream_t st) {
    static int dbg = -1;
    if (dbg < 0) { const char *e = getenv("CM_DBG"); dbg = e ? atoi(e) : 0; }
    if (nwarps == 8) {
        if (dbg & 8) k3_factor_col<8, false, 1><<<count, 256, 0, st>>>(m, L, ws, dbg);
        else if (prof) k3_factor_col<8, true, 2><<<count, 256, 0, st>>>(m, L, ws, dbg);
        else k3_factor_col<8, false, 2><<<count, 256, 0, st>>>(m, L, ws, dbg);
    } else {
        if (prof) k3_factor_col<4, true, 4><<<count, 128, 0, st>>>(m, L, ws, dbg);
        else k3_factor_col<4, false, 4><<<count, 128, 0, st>>>(m, L, ws, dbg);
    }
    return cudaGetLastError();
}

// read (and clear) the phase counters of the instrumented build: 0 late, 1 diag|look-ahead,
// 2 trsm (thread 0 of each CTA), 3 diagonal factor alone (its warp); cycles, summed over CTAs
cudaError_t cm_factor_col_read_prof(unsigned long long *out16) {
    cudaError_t e = cudaMemcpyFromSymbol(out16, g_col_prof, 16 * sizeof(unsigned long long));
    if (e != cudaSuccess) return e;
    unsigned long long z[16] = {0};
    return cudaMemcpyToSymbol(g_col_prof, z, sizeof(z));
}
