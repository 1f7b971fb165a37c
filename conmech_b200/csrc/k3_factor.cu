// K3 - batched fp64 LDL^T of the Jacobi-scaled K_mm in block-band (32x32 tile) storage, with
// pivot-based mechanism detection and the forward substitution of all load cases fused in.
//
// Replaces the reference's SuperLU factorisation + definiteness test + first triangular solve
// (numpy_stiffness.py:299-336, :353).  The reference factors with diagonal pivots only
// (diag_pivot_thresh=0, SymmetricMode), i.e. an LDL^T in another ordering, raises on an exactly
// zero pivot and rejects any pivot <= -1e-14; the same three rules are applied here to the
// pivots d_j.  Any pivot order gives the same inertia, so flags agree on well-posed subsets.
//
// Algorithm: left-looking over block columns J.  Tile (I,J) accumulates
//     A(I,J) - sum_K L(I,K) D_K L(J,K)^T,   K in [max(kfirst[I],kfirst[J]), J)
// in registers; the products are FP64 tensor-core MMAs (mma.sync m8n8k4, the only FP64 MMA
// shape family on sm_100a - tcgen05 has no f64 kind) or, in the validation variant, scalar
// FMAs.  The diagonal tile is factored by one warp, the others are solved against it
// (Y L_JJ^T = acc, L = Y D^-1) with one matrix row per lane.  One CTA per subset.
#include "cm_common.cuh"

namespace {

constexpr int K3_WARPS = 8;
constexpr int K3_THREADS = K3_WARPS * 32;
constexpr int LDS = 33;     // padded leading dimension of 32x32 shared tiles

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ const double *tile_ptr(const double *tiles, int bt, int I, int K) {
    return tiles + ((size_t)I * bt + (K - I + bt - 1)) * CM_TILE_ELEMS;
}
__device__ __forceinline__ double *tile_ptr(double *tiles, int bt, int I, int K) {
    return tiles + ((size_t)I * bt + (K - I + bt - 1)) * CM_TILE_ELEMS;
}

// acc[c] (row `lane` of the tile) -= sum_k Lik[lane][k] * d[k] * Ljk[c][k] -- scalar variant
__device__ __forceinline__ void update_scalar(double (&acc)[32], const double *Lik, const double *Ljk,
                                              const double *dK, int lane) {
    double a[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) a[k] = Lik[cm_tile_off(lane, k)] * dK[k];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 32; ++k) s += a[k] * Ljk[cm_tile_off(c, k)];
        acc[c] -= s;
    }
}

template <bool MMA>
__global__ void __launch_bounds__(K3_THREADS)
k3_factor(cm_model_dev m, cm_ws_layout L, char *ws) {
    extern __shared__ __align__(16) double smem[];
    double *s_Ljj = smem;                       // [32*LDS] unit lower factor of the diagonal tile
    double *s_dj = s_Ljj + 32 * LDS;            // [32] pivots of this block column
    double *s_col = s_dj + 32;                  // [32] column broadcast buffer of the diagonal factor
    double *s_scr = s_col + 32;                 // [K3_WARPS][32*LDS] fragment -> row-per-lane scratch
    __shared__ int s_fail;
    __shared__ double s_minpiv;

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    char *slab = ws + (size_t)blockIdx.x * L.per_subset;
    int *info = reinterpret_cast<int *>(slab + L.off_info);
    if (info[CM_INFO_STATUS] != CM_ST_OK) return;
    const int NT = info[CM_INFO_NT];
    double *tiles = reinterpret_cast<double *>(slab + L.off_tiles);
    const int *kfirst = reinterpret_cast<const int *>(slab + L.off_kfirst);
    double *dpiv = reinterpret_cast<double *>(slab + L.off_dpiv);
    double *X = reinterpret_cast<double *>(slab + L.off_x);
    const int bt = m.band_tiles;
    const int n_lc = m.n_lc;
    const int zero_bits = info[CM_INFO_ZERO_LOAD];
    double *scr = s_scr + warp * 32 * LDS;

    if (t == 0) { s_fail = 0; s_minpiv = __longlong_as_double(0x7ff0000000000000LL); }
    __syncthreads();

    for (int J = 0; J < NT; ++J) {
        const int kfJ = kfirst[J];
        const int Iend = min(J + bt - 1, NT - 1);
        bool first_round = true;
        for (int I0 = J; I0 <= Iend; I0 += K3_WARPS) {
            const int I = I0 + warp;
            const bool active = (I <= Iend) && (kfirst[I] <= J);
            double acc[32];             // row `lane` of tile (I,J)
            if (active) {
                const int K0 = max(kfirst[I], kfJ);
                if (MMA) {
                    // accumulators in MMA fragment layout: c[ri][cj] = rows ri*8+lane/4,
                    // cols cj*8+(lane%4)*2+{0,1}
                    double c[4][4][2];
                    const double *A = tile_ptr(tiles, bt, I, J);
                    const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
                    for (int ri = 0; ri < 4; ++ri)
#pragma unroll
                        for (int cj = 0; cj < 4; ++cj) {
                            const double2 v = *reinterpret_cast<const double2 *>(
                                A + cm_tile_off(ri * 8 + fr, cj * 8 + fk * 2));
                            c[ri][cj][0] = v.x; c[ri][cj][1] = v.y;
                        }
                    for (int K = K0; K < J; ++K) {
                        const double *Lik = tile_ptr(tiles, bt, I, K);
                        const double *Ljk = tile_ptr(tiles, bt, J, K);
                        const double *dK = dpiv + K * CM_TILE;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {
                            const double dk = -dK[ks * 4 + fk];
                            double a[4], b[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                a[i] = Lik[cm_tile_off(i * 8 + fr, ks * 4 + fk)];
                                b[i] = Ljk[cm_tile_off(i * 8 + fr, ks * 4 + fk)] * dk;
                            }
#pragma unroll
                            for (int ri = 0; ri < 4; ++ri)
#pragma unroll
                                for (int cj = 0; cj < 4; ++cj)
                                    dmma_m8n8k4(c[ri][cj][0], c[ri][cj][1], a[ri], b[cj]);
                        }
                    }
                    // fragment layout -> one row per lane through shared memory
#pragma unroll
                    for (int ri = 0; ri < 4; ++ri)
#pragma unroll
                        for (int cj = 0; cj < 4; ++cj) {
                            scr[(ri * 8 + fr) * LDS + cj * 8 + fk * 2] = c[ri][cj][0];
                            scr[(ri * 8 + fr) * LDS + cj * 8 + fk * 2 + 1] = c[ri][cj][1];
                        }
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < 32; ++k) acc[k] = scr[lane * LDS + k];
                    __syncwarp();
                } else {
                    const double *A = tile_ptr(tiles, bt, I, J);
#pragma unroll
                    for (int k = 0; k < 32; ++k) acc[k] = A[cm_tile_off(lane, k)];
                    for (int K = K0; K < J; ++K)
                        update_scalar(acc, tile_ptr(tiles, bt, I, K), tile_ptr(tiles, bt, J, K),
                                      dpiv + K * CM_TILE, lane);
                }
            }

            if (first_round) {
                // ---- warp 0 holds the diagonal tile: LDL^T, one row per lane
                if (warp == 0) {
                    double dval = 1.0;
                    double mp = s_minpiv;
                    int fail = 0;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        s_col[lane] = acc[c];
                        __syncwarp();
                        double d = s_col[c];
                        // reference: exactly-zero pivot raises (numpy_stiffness.py:301-319);
                        // any pivot <= -1e-14 is rejected (:326-336); NaN fails both tests
                        if (!(d > -CM_EPS) || d == 0.0) { fail = 1; d = 1.0; }
                        mp = fmin(mp, d);
                        if (lane == c) dval = d;
                        if (lane > c) {
                            const double l = acc[c] / d;
                            acc[c] = l;
#pragma unroll
                            for (int k = c + 1; k < 32; ++k) acc[k] -= l * s_col[k];
                        }
                        __syncwarp();
                    }
                    double *Ld = tile_ptr(tiles, bt, J, J);
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        const double v = (k < lane) ? acc[k] : (k == lane ? 1.0 : 0.0);
                        s_Ljj[lane * LDS + k] = v;
                        Ld[cm_tile_off(lane, k)] = v;
                    }
                    s_dj[lane] = dval;
                    dpiv[J * CM_TILE + lane] = dval;
                    if (lane == 0) { s_minpiv = mp; if (fail) s_fail = 1; }
                }
                __syncthreads();
                if (s_fail) break;
            }

            // ---- off-diagonal tiles: Y L_JJ^T = acc, L(I,J) = Y D_J^-1
            if (active && I != J) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const double y = acc[c];
#pragma unroll
                    for (int k = c + 1; k < 32; ++k) acc[k] -= y * s_Ljj[k * LDS + c];
                    acc[c] = y / s_dj[c];
                }
                double *Lo = tile_ptr(tiles, bt, I, J);
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    double2 *dst = reinterpret_cast<double2 *>(Lo + cm_tile_off(lane, g * 4));
                    dst[0] = make_double2(acc[g * 4], acc[g * 4 + 1]);
                    dst[1] = make_double2(acc[g * 4 + 2], acc[g * 4 + 3]);
                }
            }
            first_round = false;
        }
        if (s_fail) break;

        // ---- fused forward substitution of the load cases for block row J (last warp):
        // y(J) = L_JJ^-1 ( b(J) - sum_K L(J,K) y(K) )
        if (warp == K3_WARPS - 1) {
            for (int lc = 0; lc < n_lc; ++lc) {
                if ((zero_bits >> lc) & 1) continue;
                double *Xl = X + (size_t)lc * L.n_pad;
                double s = Xl[J * CM_TILE + lane];
                for (int K = kfJ; K < J; ++K) {
                    const double *Ljk = tile_ptr(tiles, bt, J, K);
                    const double *yK = Xl + K * CM_TILE;
#pragma unroll 8
                    for (int k = 0; k < 32; ++k) s -= Ljk[cm_tile_off(lane, k)] * yK[k];
                }
#pragma unroll
                for (int c = 0; c < 31; ++c) {
                    const double yc = __shfl_sync(0xffffffffu, s, c);
                    if (lane > c) s -= s_Ljj[lane * LDS + c] * yc;
                }
                Xl[J * CM_TILE + lane] = s;
            }
        }
        __syncthreads();    // block column J complete and visible to the whole CTA
    }

    if (t == 0) {
        if (s_fail) info[CM_INFO_STATUS] = CM_ST_SINGULAR;
        reinterpret_cast<double *>(slab + L.off_dpiv)[L.n_pad] = s_minpiv;    // slot after the pivots
    }
}

}  // namespace

cudaError_t cm_launch_factor(const cm_model_dev &m, const cm_ws_layout &L, char *ws, int count,
                             int variant, cudaStream_t st) {
    size_t smem = (size_t)(32 * LDS + 64 + K3_WARPS * 32 * LDS) * sizeof(double);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k3_factor<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k3_factor<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (variant == 0)
        k3_factor<false><<<count, K3_THREADS, smem, st>>>(m, L, ws);
    else
        k3_factor<true><<<count, K3_THREADS, smem, st>>>(m, L, ws);
    return cudaGetLastError();
}
