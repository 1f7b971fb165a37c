// K3 (pipelined variant, the default) - batched fp64 LDL^T in block-band tile storage with no
// block-wide barrier on the hot loop.
//
// One CTA per subset, NW "row" warps + 1 "rhs" warp.  Row warp w owns tile rows I = w, w+NW, ...
// and walks each row left to right: tile (I,J) accumulates A(I,J) - sum_K L(I,K) D_K L(J,K)^T on the
// FP64 tensor cores (the A operand is the warp's own row, the B operand is row J), then is solved
// against the diagonal factor of column J (J < I) or factored itself (J == I).  The only
// synchronisation is a per-row progress counter in shared memory (prog[J] = number of finished
// tile columns of row J): term K of tile (I,J) may run once prog[J] > K, the triangular solve
// once row J is complete.  Rows advance as a staggered wavefront; the serial 32-step diagonal
// factorisations overlap with other rows' MMAs.  To keep the row-to-row critical path short the
// diagonal tile's update with all but the newest column is done ahead of time (before the warp
// waits on the neighbouring row) and parked in the tile's own storage.
// The rhs warp trails the wavefront and performs the forward substitution of all load cases
// (y(J) = L_JJ^-1 (b(J) - sum_K L(J,K) y(K))), so the factor is read once while hot in L2.
//
// Pivot rules and their reference lines: see cm_tile_ops.cuh / k3_factor.cu.
#include "cm_tile_ops.cuh"

namespace {

constexpr int NW = 7;                       // row warps (+1 rhs warp = 8 warps, 2 per SM sub-partition)
constexpr int PIPE_THREADS = (NW + 1) * 32;
constexpr int RING = 10;                    // diagonal factors kept in shared memory (>= band_tiles)
constexpr int SCR = 32 * 33;                // per-warp transposition scratch (doubles)

__device__ __forceinline__ void wait_prog_gt(volatile int *prog, int J, int K) {
    if (prog[J] > K) return;
    while (prog[J] <= K) __nanosleep(20);
}

// phase cycle counters of the instrumented build (PROF = true), summed over all row warps:
// 0 wait-for-operand  1 MMA update (incl. fragment loads)  2 transposition  3 wait-for-diagonal
// 4 triangular solve  5 store+publish  6 look-ahead  7 diagonal MMA  8 LDL^T  9 rhs warp busy
// 10 rhs warp waiting  11 row-warp total
__device__ unsigned long long g_k3_prof[16];

#define PROF_T(var) long long var = PROF ? clock64() : 0
#define PROF_ADD(slot, t0) do { if (PROF) { long long _n = clock64(); pc[slot] += _n - (t0); (t0) = _n; } } while (0)

template <int MINB, bool PROF>
__global__ void __launch_bounds__(PIPE_THREADS, MINB)
k3_factor_pipe(cm_model_dev m, cm_ws_layout L, char *ws) {
    extern __shared__ __align__(16) double smem[];
    double *s_Lp = smem;                                // [RING][496] packed unit lower diagonal factors
    double *s_dj = s_Lp + RING * CM_TRI_ELEMS;          // [RING][32]  their pivots
    double *s_scr = s_dj + RING * 32;                   // [NW][32*33] per-row-warp scratch
    double *s_col = s_scr + NW * SCR;                   // [NW][32]
    int *s_prog = reinterpret_cast<int *>(s_col + NW * 32);   // [n_tile_rows_max]
    __shared__ int s_fail;
    __shared__ double s_minpiv[NW];

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
    volatile int *prog = s_prog;

    for (int I = t; I < NT; I += PIPE_THREADS) s_prog[I] = kfirst[I];
    if (t == 0) s_fail = 0;
    __syncthreads();

    if (warp < NW) {
        double *scr = s_scr + warp * SCR;
        double *col = s_col + warp * 32;
        double minpiv = __longlong_as_double(0x7ff0000000000000LL);
        int fail = 0;
        long long pc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        PROF_T(tstart);
        PROF_T(tp);
        for (int I = warp; I < NT; I += NW) {
            const int kfI = kfirst[I];
            double *Tdiag = cm_tile_ptr(tiles, bt, I, I);
            for (int J = kfI; J < I; ++J) {
                if (J == I - 1 && I - 2 >= kfI) {
                    // look-ahead: diagonal tile minus every term that does not need row I-1
                    TileAcc dacc;
                    acc_load(dacc, Tdiag, lane);
                    for (int K = kfI; K <= I - 2; ++K) {
                        const double *Lk = cm_tile_ptr(tiles, bt, I, K);
                        acc_update(dacc, Lk, Lk, dpiv + K * CM_TILE, lane);
                    }
                    acc_store(dacc, Tdiag, lane);
                    PROF_ADD(6, tp);
                }
                const int kfJ = kfirst[J];
                double *T = cm_tile_ptr(tiles, bt, I, J);
                double row[32];
                {
                    TileAcc acc;
                    acc_load(acc, T, lane);
                    for (int K = max(kfI, kfJ); K < J; ++K) {
                        wait_prog_gt(prog, J, K);                     // tile (J,K) finished
                        __threadfence_block();
                        PROF_ADD(0, tp);
                        acc_update(acc, cm_tile_ptr(tiles, bt, I, K), cm_tile_ptr(tiles, bt, J, K),
                                   dpiv + K * CM_TILE, lane);
                        PROF_ADD(1, tp);
                    }
                    acc_to_rows(acc, row, scr, lane);
                    PROF_ADD(2, tp);
                }
                wait_prog_gt(prog, J, J);                             // row J complete: L_JJ, d_J published
                __threadfence_block();
                PROF_ADD(3, tp);
                tile_trsm_packed(row, s_Lp + (J % RING) * CM_TRI_ELEMS, s_dj + (J % RING) * 32);
                PROF_ADD(4, tp);
                row_store(T, row, lane);
                __threadfence_block();
                __syncwarp();
                if (lane == 0) prog[I] = J + 1;
                PROF_ADD(5, tp);
            }
            // ---- diagonal tile: newest term (or all of them when there was no look-ahead), LDL^T
            {
                double row[32];
                {
                    TileAcc acc;
                    __syncwarp();
                    acc_load(acc, Tdiag, lane);
                    const int K0 = (I - 2 >= kfI) ? I - 1 : kfI;
                    for (int K = K0; K < I; ++K) {
                        const double *Lk = cm_tile_ptr(tiles, bt, I, K);
                        acc_update(acc, Lk, Lk, dpiv + K * CM_TILE, lane);
                    }
                    acc_to_rows(acc, row, scr, lane);
                    PROF_ADD(7, tp);
                }
                const double dval = tile_ldlt(row, col, lane, fail, minpiv);
                PROF_ADD(8, tp);
                double *Lp = s_Lp + (I % RING) * CM_TRI_ELEMS;
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    if (k < lane) Lp[cm_tri_off(lane, k)] = row[k];
                    row[k] = (k < lane) ? row[k] : (k == lane ? 1.0 : 0.0);
                }
                row_store(Tdiag, row, lane);
                s_dj[(I % RING) * 32 + lane] = dval;
                dpiv[I * CM_TILE + lane] = dval;
                __threadfence_block();
                __syncwarp();
                if (lane == 0) prog[I] = I + 1;
                PROF_ADD(5, tp);
            }
        }
        if (PROF && lane == 0) {
            pc[11] = clock64() - tstart;
            for (int i = 0; i < 12; ++i) if (i != 9 && i != 10) atomicAdd(&g_k3_prof[i], (unsigned long long)pc[i]);
        }
        if (lane == 0) { s_minpiv[warp] = minpiv; if (fail) s_fail = 1; }
    } else {
        // ---- rhs warp: forward substitution trailing the wavefront
        const int n_lc = m.n_lc;
        const int zero_bits = info[CM_INFO_ZERO_LOAD];
        long long pc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        PROF_T(tp);
        for (int J = 0; J < NT; ++J) {
            wait_prog_gt(prog, J, J);
            __threadfence_block();
            PROF_ADD(10, tp);
            const int kfJ = kfirst[J];
            const double *Ljj = cm_tile_ptr(tiles, bt, J, J);
            for (int lc = 0; lc < n_lc; ++lc) {
                if ((zero_bits >> lc) & 1) continue;
                double *Xl = X + (size_t)lc * L.n_pad;
                double s = Xl[J * CM_TILE + lane];
                for (int K = kfJ; K < J; ++K) {
                    const double *Ljk = cm_tile_ptr(tiles, bt, J, K);
                    const double *yK = Xl + K * CM_TILE;
#pragma unroll 8
                    for (int k = 0; k < 32; ++k) s -= Ljk[cm_tile_off(lane, k)] * yK[k];
                }
#pragma unroll
                for (int c = 0; c < 31; ++c) {
                    const double yc = __shfl_sync(0xffffffffu, s, c);
                    if (lane > c) s -= Ljj[cm_tile_off(lane, c)] * yc;
                }
                Xl[J * CM_TILE + lane] = s;
            }
            __syncwarp();
            PROF_ADD(9, tp);
        }
        if (PROF && lane == 0) { atomicAdd(&g_k3_prof[9], (unsigned long long)pc[9]); atomicAdd(&g_k3_prof[10], (unsigned long long)pc[10]); }
    }
    __syncthreads();
    if (t == 0) {
        double mp = s_minpiv[0];
        for (int w = 1; w < NW; ++w) mp = fmin(mp, s_minpiv[w]);
        if (s_fail) info[CM_INFO_STATUS] = CM_ST_SINGULAR;
        dpiv[L.n_pad] = mp;
    }
}

}  // namespace

size_t cm_factor_pipe_smem(const cm_model_dev &m) {
    return (size_t)(RING * CM_TRI_ELEMS + RING * 32 + NW * SCR + NW * 32) * sizeof(double) +
           (size_t)m.n_tile_rows_max * sizeof(int);
}

bool cm_factor_pipe_supported(const cm_model_dev &m) {
    return m.band_tiles <= RING && cm_factor_pipe_smem(m) <= 112 * 1024;
}

cudaError_t cm_launch_factor_pipe(const cm_model_dev &m, const cm_ws_layout &L, char *ws, int count,
                                  int minb, cudaStream_t st) {
    const size_t smem = cm_factor_pipe_smem(m);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(k3_factor_pipe<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k3_factor_pipe<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k3_factor_pipe<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    if (minb == 1)
        k3_factor_pipe<1, false><<<count, PIPE_THREADS, smem, st>>>(m, L, ws);
    else if (minb == 2)
        k3_factor_pipe<2, false><<<count, PIPE_THREADS, smem, st>>>(m, L, ws);
    else
        k3_factor_pipe<2, true><<<count, PIPE_THREADS, smem, st>>>(m, L, ws);      // instrumented
    return cudaGetLastError();
}

// read (and clear) the phase counters of the instrumented build
cudaError_t cm_factor_pipe_read_prof(unsigned long long *out16) {
    cudaError_t e = cudaMemcpyFromSymbol(out16, g_k3_prof, 16 * sizeof(unsigned long long));
    if (e != cudaSuccess) return e;
    unsigned long long z[16] = {0};
    return cudaMemcpyToSymbol(g_k3_prof, z, sizeof(z));
}
