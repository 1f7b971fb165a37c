// K3, scalar validation variant - the same block-band LDL^T as k3_factor_flow.cu (pivot rules,
// forward substitution, M_J = D_J^-1 L_JJ^-1 output) with plain FP64 FMAs, true triangular solves
// and element-wise tile accesses.  Slow on purpose; the parity tests use it to cross-check the
// tensor-core kernel (cm_set_variant(h, 0)).
//
// Reference: numpy_stiffness.py:299-336 (factor + definiteness test), :353 (first solve).
// Algorithm: left-looking over block columns J; tile (I,J) accumulates
//     A(I,J) - sum_K L(I,K) D_K L(J,K)^T,   K in [max(kfirst[I],kfirst[J]), J)
// one tile row per lane; the diagonal tile is factored by warp 0, the others are solved against
// it (Y L_JJ^T = acc, L = Y D^-1).  One CTA per subset.
#include "cm_frag.cuh"

namespace {

constexpr int K3_WARPS = 8;
constexpr int K3_THREADS = K3_WARPS * 32;
constexpr int LDS = 33;     // padded leading dimension of 32x32 shared tiles

__global__ void __launch_bounds__(K3_THREADS)
k3_factor_scalar(cm_model_dev m, cm_ws_layout L, char *ws) {
    __shared__ double s_Ljj[32 * LDS];         // unit lower factor of the diagonal tile
    __shared__ double s_dj[32];                // pivots of this block column
    __shared__ double s_col[32];
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
    double *minv = reinterpret_cast<double *>(slab + L.off_minv);
    const unsigned char *aflag = reinterpret_cast<const unsigned char *>(slab + L.off_aflag);
    double *X = reinterpret_cast<double *>(slab + L.off_x);
    const int bt = m.band_tiles;
    const int n_lc = m.n_lc;
    const int zero_bits = info[CM_INFO_ZERO_LOAD];

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
                const double *A = cm_tile_ptr(tiles, bt, I, J);
                const bool has_a = aflag[I * bt + (J - I + bt - 1)] != 0;     // else the storage is not initialised
#pragma unroll
                for (int k = 0; k < 32; ++k) acc[k] = has_a ? A[cm_tile_off(lane, k)] : 0.0;
                for (int K = max(kfirst[I], kfJ); K < J; ++K) {
                    const double *Lik = cm_tile_ptr(tiles, bt, I, K);
                    const double *Ljk = cm_tile_ptr(tiles, bt, J, K);
                    const double *dK = dpiv + K * CM_TILE;
                    for (int k = 0; k < 32; ++k) {
                        const double a = Lik[cm_tile_off(lane, k)] * dK[k];
#pragma unroll
                        for (int c = 0; c < 32; ++c) acc[c] -= a * Ljk[cm_tile_off(c, k)];
                    }
                }
            }
            if (first_round) {
                if (warp == 0) {
                    double dval = 1.0;
                    double mp = s_minpiv;
                    int fail = 0;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        s_col[lane] = acc[c];
                        __syncwarp();
                        double d = s_col[c];
                        // exactly-zero pivot raises in the reference (numpy_stiffness.py:301-319); any
                        // pivot <= -1e-14 is rejected (:326-336); NaN fails both tests
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
#pragma unroll
                    for (int k = 0; k < 32; ++k) s_Ljj[lane * LDS + k] = (k < lane) ? acc[k] : (k == lane ? 1.0 : 0.0);
                    s_dj[lane] = dval;
                    dpiv[J * CM_TILE + lane] = dval;
                    __syncwarp();
                    // column `lane` of L_JJ^-1 by forward substitution on e_lane; M = D^-1 L^-1
                    double x[32];
                    double *Mj = minv + (size_t)J * CM_TILE_ELEMS;
                    for (int r = 0; r < 32; ++r) {
                        double s = (r == lane) ? 1.0 : 0.0;
                        for (int k = 0; k < r; ++k) s -= s_Ljj[r * LDS + k] * x[k];
                        x[r] = s;
                        Mj[cm_mtile_off(r, lane)] = s / s_dj[r];
                    }
                    if (lane == 0) { s_minpiv = mp; if (fail) s_fail = 1; }
                }
                __syncthreads();
                if (s_fail) break;
            }
            // off-diagonal tiles: Y L_JJ^T = acc, L(I,J) = Y D_J^-1
            if (active && I != J) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const double y = acc[c];
#pragma unroll
                    for (int k = c + 1; k < 32; ++k) acc[k] -= y * s_Ljj[k * LDS + c];
                    acc[c] = y / s_dj[c];
                }
                double *Lo = cm_tile_ptr(tiles, bt, I, J);
#pragma unroll
                for (int k = 0; k < 32; ++k) Lo[cm_tile_off(lane, k)] = acc[k];
            }
            first_round = false;
        }
        if (s_fail) break;
        __syncthreads();
        // forward substitution of the load cases for block row J (last warp):
        // y(J) = L_JJ^-1 ( b(J) - sum_K L(J,K) y(K) )
        if (warp == K3_WARPS - 1) {
            for (int lc = 0; lc < n_lc; ++lc) {
                if ((zero_bits >> lc) & 1) continue;
                double *Xl = X + (size_t)lc * L.n_pad;
                double s = Xl[J * CM_TILE + lane];
                for (int K = kfJ; K < J; ++K) {
                    const double *Ljk = cm_tile_ptr(tiles, bt, J, K);
                    const double *yK = Xl + K * CM_TILE;
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
        dpiv[L.n_pad] = s_minpiv;    // slot after the pivots
    }
}

}  // namespace

cudaError_t cm_launch_factor(const cm_model_dev &m, const cm_ws_layout &L, char *ws, int count,
                             int variant, cudaStream_t st) {
    (void)variant;
    k3_factor_scalar<<<count, K3_THREADS, 0, st>>>(m, L, ws);
    return cudaGetLastError();
}
