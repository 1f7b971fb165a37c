// Shared declarations of the conmech_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <atomic>
#include <vector>
#include <string>
#include "../../include/conmech_b200.h"

#define CM_TILE 32                 // tile edge of the block-band storage
#define CM_TILE_ELEMS 1024         // doubles per tile (8 KB)
#define CM_EPS 1e-14               // reference DOUBLE_EPS (numpy_stiffness_assembly.py:8)
#define CM_BIG 1e14                // reference DOUBLE_INF (numpy_stiffness_assembly.py:9)
#define CM_MAX_LC 8                // load cases solved together per subset
enum { CM_K2T_ORDER = 0, CM_K2T_NE_PTR, CM_K2T_NE_ELEM, CM_K2T_LP_PTR, CM_K2T_LP_NODE, CM_K2T_LP_EPTR, CM_K2T_LP_ELEM,
       CM_K2T_NODE_SUP, CM_K2T_CONN, CM_K2T_NE_END, CM_K2T_LP_END, CM_K2T_SUP_COND };
#define CM_INFO_INTS 16            // per-subset info record (ints)

// info record slots
#define CM_INFO_NS 0               // free DOFs of the subset (solver matrix order)
#define CM_INFO_NT 1               // tile rows
#define CM_INFO_STATUS 2           // CM_ST_OK / NO_SUPPORT / NO_FIXED_DOF / SINGULAR
#define CM_INFO_NFIXED 3
#define CM_INFO_ZERO_LOAD 4        // bit lc set: ||P_lc|| <= eps
#define CM_INFO_JACOBI 5           // 1: Jacobi scaling applied, 0: identity (a |k_ii| <= eps)
#define CM_INFO_COST 6             // factorisation cost of the subset (float bits of its envelope flops; 0: nothing to factor)
#define CM_INFO_SCHED 7            // scheduling: record r holds the slab index of the subset with the r-th largest cost
// Prefix-sharing groups (cm_solve_many_host_seq): a "tail" subset shares the first TROWS tile rows of its factor with
// the "trunk" subset of its group (slab index TRUNK within the wave) instead of computing them again.
#define CM_INFO_TROWS 8            // shared tile rows (0: the subset stands alone / is a trunk)
#define CM_INFO_TRUNK 9            // slab index of the trunk
#define CM_INFO_FAILROW 10         // 0, or 1 + the first tile row whose diagonal factorisation hit a bad pivot

// Element (r,k) of a 32x32 tile, "fragment-major" order: [k/4][(r%8)*4 + k%4][r/8].
// For the FP64 MMA m8n8k4 (lane = 4*g + t) the four A (or B) fragments of k-step ks that a lane
// needs - rows g, 8+g, 16+g, 24+g at k = 4*ks + t - are the 4 consecutive doubles at
// ks*128 + lane*4, so a warp fetches one k-step of an operand tile as one contiguous 1 KB run of
// 32-byte vector loads; accumulator fragments (rows ri*8+g, columns cj*8 + 2t, 2t+1, ri = 0..3)
// are 8 consecutive doubles per (lane, cj).  The same storage therefore serves as MMA operand
// and as accumulator without any shared-memory transposition.
__host__ __device__ __forceinline__ int cm_tile_off(int r, int k) {
    return ((k >> 2) << 7) + ((((r & 7) << 2) + (k & 3)) << 2) + (r >> 3);
}

// Element (r,k) of a lower-triangular M_J = D_J^-1 L_JJ^-1 tile, "Mq" order: [r/8][k/8][lane][k%2]
// with lane = 4*(r%8) + (k%8)/2.  As the B operand of the solve product L(I,J) = acc * M_J^T taken
// straight from accumulator registers (k-slot t of k-step (cj,e) <-> column 8cj + 2t + e, which is
// where the accumulator fragment of that lane already sits), the two fragments a lane needs for
// output block r/8 and k-block cj are one 16-byte vector; blocks above the diagonal are never
// stored or read.  tools/frag_emulator.py checks these maps lane by lane.
__host__ __device__ __forceinline__ int cm_mtile_off(int r, int k) {
    return ((r >> 3) << 8) + ((k >> 3) << 6) + ((((r & 7) << 2) + ((k & 7) >> 1)) << 1) + (k & 1);
}

// Model tables on the device (all sizes fixed at cm_create).
struct cm_model_dev {
    int device, n_sm;      // CUDA device the tables live on and its SM count (fixed at cm_create)
    int nV, nE, nS, n_lc;
    int n_geom;            // geometry variants of the frame (same topology, different node coordinates): the tables marked
                           // [n_geom x ...] below are stacked, a subset picks its variant by index
    int mask_words;
    int n_tile_rows_max;   // tile rows of the full structure
    int band_tiles;        // tiles stored per tile row: columns I-band_tiles+1 .. I
    const double *xyz;     // [n_geom x nV*3]
    const int *conn;       // [nE*2]
    const double *props;   // [nE*7] A,Jx,Iy,Iz,E,mu,rho
    const double *cr;      // [nE*6]
    double *Kg;            // [n_geom x nE*144] row major global element stiffness
    double *Kt;            // [n_geom x nE*144] the same with entries |k| <= 1e-14 set to zero: what the assembly sums (the
                           //   reference drops those entries per element, numpy_stiffness_assembly.py:384); written by K1 so
                           //   that the gather of K2 and the reactions of K4 do not test every entry of every subset again
    double *R3;            // [n_geom x nE*9]
    double *elen;          // [n_geom x nE]
    const int *ne_ptr;     // [nV+1] node -> incident elements (ascending ids)
    const int *ne_elem;    // [2nE]
    const unsigned char *ne_end;   // [2nE] which end of that element the node is
    const int *sup_node;   // [nS]
    const unsigned char *sup_cond; // [nS*6]
    const int *node_sup;   // [nV] support index or -1
    // "lower pairs": for node a the distinct neighbours c that precede it in the solver order, and
    // per pair the elements joining them (ascending ids) with the end at which a sits
    const int *lp_ptr;     // [nV+1]
    const int *lp_node;    // [n_pairs]
    const int *lp_eptr;    // [n_pairs+1]
    const int *lp_elem;    // [<= nE]
    const unsigned char *lp_end;   // [<= nE]
    const int *order;      // [nV] solver position -> node
    const int *pos;        // [nV] node -> solver position
    // The index tables K2 walks (order, ne_*, lp_*, node_sup, conn, sup_cond), once more as ONE
    // contiguous blob: K2 copies it to shared memory when it fits, so that the dependent address
    // chains of its gather run at shared-memory latency.  Offsets in bytes (CM_K2T_* below).
    const char *k2_blob;
    int k2_blob_bytes;
    int k2_off[12];
    const double *q;       // [n_geom x n_lc*nE*12] lumped element loads
    const double *pnodal;  // [n_lc*6nV]
    double trans_tol;
};

// Opt-in dynamic shared memory of a kernel.  The attribute is PER DEVICE (and per kernel), handles on
// several devices may live in one process and be driven from several host threads: the cache is keyed
// by device and atomic; setting the attribute twice is harmless.  Always requested (not only above
// 48 KB): the no-opt-in limit counts static + dynamic shared memory.
#define CM_MAX_DEVICES 64
struct cm_smem_cache { std::atomic<size_t> bytes[CM_MAX_DEVICES]; };
template <typename K>
static inline cudaError_t cm_ensure_smem(K kernel, size_t bytes, cm_smem_cache &c, int device) {
    // (the cache holds bytes + 1: 0 = never set on that device)
    if (device >= 0 && device < CM_MAX_DEVICES && c.bytes[device].load(std::memory_order_acquire) > bytes)
        return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    if (device >= 0 && device < CM_MAX_DEVICES) c.bytes[device].store(bytes + 1, std::memory_order_release);
    return cudaSuccess;
}

// Per-subset workspace layout: byte offsets inside one subset's slab.
struct cm_ws_layout {
    size_t per_subset;
    size_t off_tiles;      // double [n_tile_rows_max*band_tiles*1024]
    size_t off_sidx;       // int    [6nV]  DOF -> solver row (-1: not free)
    size_t off_dinv;       // double [n_pad] Jacobi scale per solver row
    size_t off_dpiv;       // double [n_pad] LDL^T pivots
    size_t off_x;          // double [n_lc*n_pad] rhs -> forward solution -> solution
    size_t off_p;          // double [n_lc*6nV] full load vector, DOF order
    size_t off_kfirst;     // int    [n_tile_rows_max] first tile column in the envelope
    size_t off_minv;       // double [n_tile_rows_max*1024] M_J = D_J^-1 L_JJ^-1 per diagonal tile (cm_mtile_off layout)
    size_t off_aflag;      // uchar  [n_tile_rows_max*band_tiles] 1: the tile holds entries of K_mm (assembled by K2);
                           //        0: structurally empty before fill-in, its storage is NOT initialised
    size_t off_kcol;       // int    [n_tile_rows_max] first nonzero column of the tile row, in k-steps of 4
                           //        (rounded down to even): columns left of it are zero in A and in L
    size_t off_info;       // int    [CM_INFO_INTS]
    int n_pad;             // n_tile_rows_max*32
};

// Output pointers as typed device pointers (0 = not wanted).
struct cm_out_dev {
    unsigned char *flag, *status;
    int *n_free, *n_fixed, *dof_map;
    unsigned char *node_exists;
    double *U, *Rf, *eR, *compliance, *max_trans, *max_rot, *min_pivot, *profile_flops, *sigma_max, *max_abs_sigma;
    double *profile_entries, *envelope_tiles;
    // compact results (include/conmech_b200.h): reactions of the support nodes, end forces of the existing elements
    double *Rf_sup, *eR_rows;
    const long long *eR_row_offsets;      // [B+1], indexed like the masks
    long long row_bias;                   // rows of subset gs start at eR_rows + (off[gs] - row_bias) * n_lc * 12
};

// kernel launchers (one per stage); all enqueue on `st` and return the launch error
cudaError_t cm_launch_element_tables(const cm_model_dev &m, cudaStream_t st);
cudaError_t cm_launch_element_loads(const cm_model_dev &m, int n_lc, const double *w_udl,
                                    const unsigned char *has_udl, const unsigned char *grav_on,
                                    const double *grav_dir, double *q, cudaStream_t st);
cudaError_t cm_launch_element_generic(const double *xyz_u, const double *xyz_v, const double *props,
                                      const double *cr, long long n, double *Kg, double *R3,
                                      cudaStream_t st);
// ktilde_rowmajor: K2 writes the assembled K~ tiles row-major (r*32 + k) instead of fragment-major.
// K~ is only ever read as the INITIAL ACCUMULATOR of its tile (never as an MMA operand), the warp
// that reads it owns the tile at that moment and overwrites it with L in the fragment-major layout -
// so the layout K2 writes is free, and row-major makes the six entries of a block row contiguous
// (2 sectors instead of 6 partial-sector writes).  The tensor-core factorisation reads it that way
// (acc_load_rm); the scalar validation kernel wants fragment-major.
// geom: per-subset geometry index (device, indexed like masks) or nullptr = geometry 0
// trows / group: prefix-sharing groups - subset gs shares trows[gs] tile rows with subset (gs / group) * group, which is
// in the same wave (first and count are multiples of group); nullptr / 0 = no sharing
// part: 0 = all `count` subsets; 1 = only the first subset of every group (the trunks: count / group CTAs);
// 2 = all but those
cudaError_t cm_launch_dofmap_assemble(const cm_model_dev &m, const cm_ws_layout &L, char *ws,
                                      const uint32_t *masks, const int *geom, const int *trows, int group, int part,
                                      long long first, int count, const cm_out_dev &out, int ktilde_rowmajor, cudaStream_t st);
// ranks the subsets of a wave by factorisation cost, largest first; subsets without shared rows (trunks) before the tails
cudaError_t cm_launch_schedule(const cm_ws_layout &L, char *ws, int count, cudaStream_t st);
cudaError_t cm_launch_factor(const cm_model_dev &m, const cm_ws_layout &L, char *ws, int count,
                             int variant, cudaStream_t st);
// factors the subsets of ranks rank0 .. rank0 + count - 1 of the wave's schedule; stride > 0: no schedule, CTA b
// factors the subset in slab b * stride (the trunks of prefix-sharing groups, before the wave has been ranked)
cudaError_t cm_launch_factor_flow(const cm_model_dev &m, const cm_ws_layout &L, char *ws, int rank0, int count, int stride,
                                  int config, int prof, cudaStream_t st);
cudaError_t cm_factor_flow_read_prof(unsigned long long *out16);
cudaError_t cm_k2_read_prof(unsigned long long *out8);
cudaError_t cm_launch_solve_recover(const cm_model_dev &m, const cm_ws_layout &L, char *ws,
                                    const uint32_t *masks, const int *geom, long long first, int count,
                                    const cm_out_dev &out, cudaStream_t st);
