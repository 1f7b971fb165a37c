// K2 - per subset: DOF status / permutation (bit-exact integers), compacted solver numbering in
// the static node ordering, and the deterministic assembly of the Jacobi-scaled K_mm and the
// right-hand sides into block-band tile storage.
//
// Reference: numpy_stiffness.py:147-219 (compute_dof_permutation), numpy_stiffness_assembly.py:
// 347-389 (assemble_global_stiffness_matrix, entries |k| <= 1e-14 dropped),
// numpy_stiffness.py:244-257 (K_mm slice, load vector), :283-291 (Jacobi scaling).
//
// Assembly is a GATHER: every matrix entry is owned by one thread, which sums the contributing
// element entries in ascending element id.  No atomics touch floating-point data, so results
// are run-to-run deterministic.  One CTA per subset.
#include "cm_common.cuh"

namespace {

constexpr int K2_THREADS = 256;

// instrumented build (-DK2_PROF=1, tools/build_variant.py): cycles of thread 0 between the block barriers that separate the
// phases, summed over CTAs (slots: 0 table / mask staging, 1 node existence, 2 numbering + envelope, 3 diagonal gather,
// 4 Jacobi scale, 5 fill + scatter, 6 load vectors, 7 total)
__device__ unsigned long long g_k2_prof[16];     // 8..: inside fill + scatter, per warp (lane 0): 8 zero fill incl. wait, 9 row lookup, 10 diagonal step, 11 pair steps, 12 idle at the closing barrier
#if defined(K2_PROF) && K2_PROF
#define K2P_T0() long long _k2t = clock64(); const long long _k2start = _k2t
#define K2P(slot) do { if (threadIdx.x == 0) { const long long _n = clock64(); atomicAdd(&g_k2_prof[slot], (unsigned long long)(_n - _k2t)); _k2t = _n; } } while (0)
#define K2P_END() do { if (threadIdx.x == 0) atomicAdd(&g_k2_prof[7], (unsigned long long)(clock64() - _k2start)); } while (0)
#define K2W_T0() long long _k2w = clock64()
#define K2W(slot) do { const long long _n = clock64(); if ((threadIdx.x & 31) == 0) atomicAdd(&g_k2_prof[slot], (unsigned long long)(_n - _k2w)); _k2w = _n; } while (0)
#else
#define K2W_T0()
#define K2W(slot)
#define K2P_T0()
#define K2P(slot)
#define K2P_END()
#endif


// exclusive scan of n ints in shared memory, in place; returns the total (all threads).
// scratch: K2_THREADS+1 ints.
__device__ int block_excl_scan(int *data, int n, int *scratch) {
    const int t = threadIdx.x, T = blockDim.x;
    const int chunk = (n + T - 1) / T;
    const int lo = min(t * chunk, n), hi = min(lo + chunk, n);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += data[i];
    scratch[t] = s;
    __syncthreads();
    if (t == 0) {
        int run = 0;
        for (int i = 0; i < T; ++i) { int v = scratch[i]; scratch[i] = run; run += v; }
        scratch[T] = run;
    }
    __syncthreads();
    int run = scratch[t];
    for (int i = lo; i < hi; ++i) { int v = data[i]; data[i] = run; run += v; }
    int total = scratch[T];
    __syncthreads();
    return total;
}

__device__ __forceinline__ bool elem_on(const uint32_t *mask, int e) { return (mask[e >> 5] >> (e & 31)) & 1u; }

// solver row of local DOF i of node v (-1 if fixed); base = first solver row of the node
__device__ __forceinline__ int node_row(const int *node_sup, const unsigned char *sup_cond, int v, int i, int base) {
    const int s = node_sup[v];
    if (s < 0) return base + i;            // (the common case: not a support node)
    const unsigned char *c = sup_cond + 6 * s;
    if (c[i]) return -1;
    int r = base;
    for (int j = 0; j < i; ++j) r += c[j] ? 0 : 1;
    return r;
}

// next tile row of the CTA for the calling warp
__device__ __forceinline__ int next_row(int *counter, int lane) {
    int r = 0;
    if (lane == 0) r = atomicAdd(counter, 1);
    return __shfl_sync(0xffffffffu, r, 0);
}

#ifndef K2_BULK_FILL
#define K2_BULK_FILL 1
#endif
#ifndef K2_MIN_CTAS
#define K2_MIN_CTAS 4          // 64 registers, no spills: 4 CTAs 2.32 ms, 5 CTAs (48 registers, spills) 2.59, 6 CTAs 2.73 per 4096 subsets
                               // (round 1, with the scatter's value array still in local memory, 5 CTAs won: 3.0 against 3.35)
#endif
template <bool TS, bool DS>      // TS: the index tables are staged in shared memory; DS: so is the Jacobi scale (n_pad doubles)
__global__ void __launch_bounds__(K2_THREADS, K2_MIN_CTAS)
k2_dofmap_assemble(cm_model_dev m, cm_ws_layout L, char *ws, const uint32_t *masks, const int *geom, const int *trows_of,
                   int group, int part, long long first, cm_out_dev out, int ktilde_rm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // shared layout
    int *s_base = reinterpret_cast<int *>(smem_raw);             // [nV] first solver row of node
    int *s_tmp = s_base + m.nV;                                    // [nV] scan input / scratch
    int *s_kfirst = s_tmp + m.nV;                                  // [n_tile_rows_max]
    int *s_kcol = s_kfirst + m.n_tile_rows_max;                    // [n_tile_rows_max] first column / 4
    int *s_scan = s_kcol + m.n_tile_rows_max;                      // [K2_THREADS+1]
    uint32_t *s_mask = reinterpret_cast<uint32_t *>(s_scan + K2_THREADS + 1);   // [mask_words]
    unsigned char *s_free = reinterpret_cast<unsigned char *>(s_mask + m.mask_words);   // [nV] free DOFs (0 if absent)
    unsigned char *s_exist = s_free + m.nV;                        // [nV]
    unsigned char *s_aflag = s_exist + m.nV;                       // [n_tile_rows_max*band_tiles] tile holds K_mm entries
    char *s_blob = reinterpret_cast<char *>((reinterpret_cast<size_t>(s_aflag + m.n_tile_rows_max * m.band_tiles) + 15) & ~(size_t)15);   // [k2_blob_bytes] if TS
    // the Jacobi scale of every solver row: read back per matrix entry by the scatter (row and six columns) and by the
    // right-hand sides - from global memory that was an L2 round trip on the latency chain of every block row
    double *s_dinv = reinterpret_cast<double *>(s_blob + (TS ? ((m.k2_blob_bytes + 15) & ~15) : 0));              // [n_pad] if DS
    __shared__ int s_red[8];
#if K2_BULK_FILL
    __shared__ __align__(128) double s_zero[512];      // source of the bulk zero fill
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_zero[i] = 0.0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // (ordered before the first bulk copy by the barriers below)
#endif
    __shared__ unsigned long long s_work, s_ent;
    __shared__ int s_ntile, s_next;
    __shared__ double s_dred[K2_THREADS / 32];

    const int t = threadIdx.x, T = blockDim.x;
    K2P_T0();
    // part 1: CTA b assembles the first subset of group b (the trunks); part 2: every subset but those
    const int sb = part == 1 ? blockIdx.x * group : blockIdx.x;
    if (part == 2 && sb % group == 0) return;
    const long long gs = first + sb;
    char *slab = ws + (size_t)sb * L.per_subset;
    double *tiles = reinterpret_cast<double *>(slab + L.off_tiles);
    int *sidx = reinterpret_cast<int *>(slab + L.off_sidx);
    double *dinv = reinterpret_cast<double *>(slab + L.off_dinv);
    double *X = reinterpret_cast<double *>(slab + L.off_x);
    double *P = reinterpret_cast<double *>(slab + L.off_p);
    int *kfirst_g = reinterpret_cast<int *>(slab + L.off_kfirst);
    int *kcol_g = reinterpret_cast<int *>(slab + L.off_kcol);
    unsigned char *aflag_g = reinterpret_cast<unsigned char *>(slab + L.off_aflag);
    int *info = reinterpret_cast<int *>(slab + L.off_info);
    const uint32_t *gmask = masks + (size_t)gs * m.mask_words;
    const int nV = m.nV;
    // geometry variant of this subset: its slice of the stacked stiffness / load tables
    const size_t gv = geom ? (size_t)min(max(geom[gs], 0), m.n_geom - 1) : 0;
    const double *Kg = m.Kt + gv * m.nE * 144;        // entries |k| <= 1e-14 already dropped (K1)
    const double *qtab = m.q + gv * m.n_lc * m.nE * 12;

    // index tables: shared-memory copies when they fit (one coalesced pass), the global blob otherwise
    const char *blob = TS ? s_blob : m.k2_blob;
    if (TS) {
        const int4 *src = reinterpret_cast<const int4 *>(m.k2_blob);
        int4 *dst = reinterpret_cast<int4 *>(s_blob);
        for (int i = t; i < (m.k2_blob_bytes >> 4); i += T) dst[i] = src[i];
    }
    const int *order = reinterpret_cast<const int *>(blob + m.k2_off[CM_K2T_ORDER]);
    const int *ne_ptr = reinterpret_cast<const int *>(blob + m.k2_off[CM_K2T_NE_PTR]);
    const int *ne_elem = reinterpret_cast<const int *>(blob + m.k2_off[CM_K2T_NE_ELEM]);
    const int *lp_ptr = reinterpret_cast<const int *>(blob + m.k2_off[CM_K2T_LP_PTR]);
    const int *lp_node = reinterpret_cast<const int *>(blob + m.k2_off[CM_K2T_LP_NODE]);
    const int *lp_eptr = reinterpret_cast<const int *>(blob + m.k2_off[CM_K2T_LP_EPTR]);
    const int *lp_elem = reinterpret_cast<const int *>(blob + m.k2_off[CM_K2T_LP_ELEM]);
    const int *node_sup = reinterpret_cast<const int *>(blob + m.k2_off[CM_K2T_NODE_SUP]);
    const int *conn = reinterpret_cast<const int *>(blob + m.k2_off[CM_K2T_CONN]);
    const unsigned char *ne_end = reinterpret_cast<const unsigned char *>(blob + m.k2_off[CM_K2T_NE_END]);
    const unsigned char *lp_end = reinterpret_cast<const unsigned char *>(blob + m.k2_off[CM_K2T_LP_END]);
    const unsigned char *sup_cond = reinterpret_cast<const unsigned char *>(blob + m.k2_off[CM_K2T_SUP_COND]);

    for (int w = t; w < m.mask_words; w += T) s_mask[w] = gmask[w];
    if (t < 8) s_red[t] = 0;
    if (t == 0) { s_work = 0ull; s_ent = 0ull; s_ntile = 0; }
    __syncthreads();
    K2P(0);

    // ---- node existence, support status (numpy_stiffness.py:156-184)
    int my_fixed = 0, my_exist = 0, my_sup = 0;
    for (int v = t; v < nV; v += T) {
        bool ex = false;
        for (int p = ne_ptr[v]; p < ne_ptr[v + 1]; ++p) ex |= elem_on(s_mask, ne_elem[p]);
        int nfix = 0;
        int s = node_sup[v];
        if (ex && s >= 0) {
            for (int i = 0; i < 6; ++i) nfix += sup_cond[6 * s + i] ? 1 : 0;
            my_sup = 1;
        }
        s_exist[v] = ex ? 1 : 0;
        s_free[v] = ex ? (unsigned char)(6 - nfix) : 0;
        my_fixed += nfix;
        my_exist += ex ? 1 : 0;
    }
    atomicAdd(&s_red[0], my_fixed);
    atomicAdd(&s_red[1], my_exist);
    atomicOr(&s_red[2], my_sup);
    __syncthreads();
    K2P(1);
    const int n_fixed = s_red[0], n_exist = s_red[1];
    const int status = (s_red[2] == 0) ? CM_ST_NO_SUPPORT : (n_fixed == 0 ? CM_ST_NO_FIXED_DOF : CM_ST_OK);
    const int n_free = 6 * n_exist - n_fixed;

    if (out.node_exists)
        for (int v = t; v < nV; v += T) out.node_exists[(size_t)gs * nV + v] = s_exist[v];
    if (t == 0) {
        if (out.n_free) out.n_free[gs] = status == CM_ST_OK ? n_free : 0;
        if (out.n_fixed) out.n_fixed[gs] = status == CM_ST_OK ? n_fixed : 0;
        if (out.profile_flops && status != CM_ST_OK) out.profile_flops[gs] = 0.0;
        if (out.profile_entries && status != CM_ST_OK) out.profile_entries[gs] = 0.0;
        if (out.envelope_tiles && status != CM_ST_OK) out.envelope_tiles[gs] = 0.0;
        info[CM_INFO_STATUS] = status;
        info[CM_INFO_NS] = status == CM_ST_OK ? n_free : 0;
        info[CM_INFO_NT] = status == CM_ST_OK ? (n_free + CM_TILE - 1) / CM_TILE : 0;
        info[CM_INFO_NFIXED] = n_fixed;
        info[CM_INFO_ZERO_LOAD] = 0;
        info[CM_INFO_JACOBI] = 1;
        info[CM_INFO_COST] = 0;
        info[CM_INFO_FAILROW] = 0;
        // prefix-sharing group: the first trows tile rows of this subset's factor are those of the group's first subset
        // (recorded for K3 / K4; the assembly below is complete all the same, so that the factorisation can fall back to
        // factoring the subset on its own when the trunk turns out not to be usable)
        info[CM_INFO_TROWS] = (trows_of && status == CM_ST_OK) ? min(trows_of[gs], n_free / CM_TILE) : 0;
        info[CM_INFO_TRUNK] = group > 0 ? (int)((gs / group) * group - first) : 0;
        s_next = 0;
    }
    if (status != CM_ST_OK) return;      // uniform for the block

    // ---- the reference's id_map_RO: stable partition free | fixed | absent in DOF order
    // (numpy_stiffness.py:193-208); three node-level exclusive scans
    if (out.dof_map) {
        int *dm = out.dof_map + (size_t)gs * 6 * nV;
        // free
        for (int v = t; v < nV; v += T) s_tmp[v] = s_free[v];
        __syncthreads();
        block_excl_scan(s_tmp, nV, s_scan);
        for (int v = t; v < nV; v += T) {
            if (!s_exist[v]) continue;
            int r = s_tmp[v];
            int s = node_sup[v];
            for (int i = 0; i < 6; ++i)
                if (s < 0 || !sup_cond[6 * s + i]) dm[r++] = 6 * v + i;
        }
        __syncthreads();
        // fixed
        for (int v = t; v < nV; v += T) s_tmp[v] = s_exist[v] ? 6 - s_free[v] : 0;
        __syncthreads();
        block_excl_scan(s_tmp, nV, s_scan);
        for (int v = t; v < nV; v += T) {
            if (!s_exist[v]) continue;
            int s = node_sup[v];
            if (s < 0) continue;
            int r = n_free + s_tmp[v];
            for (int i = 0; i < 6; ++i)
                if (sup_cond[6 * s + i]) dm[r++] = 6 * v + i;
        }
        __syncthreads();
        // absent
        for (int v = t; v < nV; v += T) s_tmp[v] = s_exist[v] ? 0 : 6;
        __syncthreads();
        block_excl_scan(s_tmp, nV, s_scan);
        for (int v = t; v < nV; v += T) {
            if (s_exist[v]) continue;
            int r = n_free + n_fixed + s_tmp[v];
            for (int i = 0; i < 6; ++i) dm[r + i] = 6 * v + i;
        }
        __syncthreads();
    }

    // ---- compacted solver numbering: free DOFs of existing nodes in the static node ordering
    for (int p = t; p < nV; p += T) s_tmp[p] = s_free[order[p]];
    __syncthreads();
    block_excl_scan(s_tmp, nV, s_scan);
    for (int p = t; p < nV; p += T) s_base[order[p]] = s_tmp[p];
    const int ns = n_free;
    const int NT = (ns + CM_TILE - 1) / CM_TILE;
    for (int I = t; I < NT; I += T) { s_kfirst[I] = I; s_kcol[I] = 8 * I; }
    for (int i = t; i < NT * m.band_tiles; i += T) s_aflag[i] = 0;
    __syncthreads();
    for (int v = t; v < nV; v += T) {
        int base = s_base[v];
#pragma unroll
        for (int i = 0; i < 6; ++i) sidx[6 * v + i] = s_exist[v] ? node_row(node_sup, sup_cond, v, i, base) : -1;
        // envelope: first column touched by the rows of this node
        int nf = s_free[v];
        if (nf == 0) continue;
        int cmin = base;
        const int I0 = base / CM_TILE, I1 = (base + nf - 1) / CM_TILE;
        const int btf = m.band_tiles;
        // tiles that receive entries of K_mm: the node's own diagonal block and the blocks towards
        // neighbours that come earlier in the solver order
        s_aflag[I0 * btf + btf - 1] = 1;
        if (I1 != I0) { s_aflag[I1 * btf + btf - 1] = 1; s_aflag[I1 * btf + btf - 2] = 1; }
        for (int p = ne_ptr[v]; p < ne_ptr[v + 1]; ++p) {
            int e = ne_elem[p];
            if (!elem_on(s_mask, e)) continue;
            int c = conn[2 * e + (ne_end[p] ? 0 : 1)];
            if (!s_free[c]) continue;
            const int bc = s_base[c];
            cmin = min(cmin, bc);
            if (bc < base) {
                const int K0 = bc / CM_TILE, K1 = (bc + s_free[c] - 1) / CM_TILE;
                s_aflag[I0 * btf + (K0 - I0 + btf - 1)] = 1;
                if (K1 <= I0) s_aflag[I0 * btf + (K1 - I0 + btf - 1)] = 1;
                if (I1 != I0) {
                    s_aflag[I1 * btf + (K0 - I1 + btf - 1)] = 1;
                    s_aflag[I1 * btf + (K1 - I1 + btf - 1)] = 1;
                }
            }
        }
        atomicMin(&s_kfirst[I0], cmin / CM_TILE);
        atomicMin(&s_kcol[I0], cmin / 4);
        if (I1 != I0) { atomicMin(&s_kfirst[I1], cmin / CM_TILE); atomicMin(&s_kcol[I1], cmin / 4); }
        unsigned long long w = 0;      // algorithmic work: sum over rows of (row - first column)^2
        for (int i = 0; i < nf; ++i) { unsigned long long hgt = (unsigned long long)(base + i - cmin); w += hgt * hgt; }
        atomicAdd(&s_work, w);
        if (out.profile_entries)       // sum over the node's rows of (row - first column + 1)
            atomicAdd(&s_ent, (unsigned long long)nf * (unsigned long long)(base - cmin + 1) + (unsigned long long)(nf * (nf - 1) / 2));
    }
    __syncthreads();
    K2P(2);
    if (t == 0 && out.profile_flops) out.profile_flops[gs] = (double)s_work;
    // cost key of the schedule: > 0, monotone in the work.  (For a tail the shared rows still count: an upper bound.)
    if (t == 0) info[CM_INFO_COST] = __float_as_int((float)s_work + 1.0f);
    for (int I = t; I < NT; I += T) {
        kfirst_g[I] = s_kfirst[I]; kcol_g[I] = s_kcol[I] & ~1;
        if (out.envelope_tiles) atomicAdd(&s_ntile, I - s_kfirst[I] + 1);
    }
    if (out.profile_entries || out.envelope_tiles) {
        __syncthreads();
        if (t == 0 && out.profile_entries) out.profile_entries[gs] = (double)s_ent;
        if (t == 0 && out.envelope_tiles) out.envelope_tiles[gs] = (double)s_ntile;
    }
    for (int i = t; i < NT * m.band_tiles; i += T) aflag_g[i] = s_aflag[i];

    // ---- diagonal of K_mm and the Jacobi scale (numpy_stiffness.py:283-290)
    int ok = 1;
    for (int idx = t; idx < 6 * nV; idx += T) {
        int v = idx / 6, i = idx - 6 * v;
        if (!s_exist[v]) continue;
        int r = node_row(node_sup, sup_cond, v, i, s_base[v]);
        if (r < 0) continue;
        double kii = 0.0;
        // (table reads are unconditional and the mask only selects: the loads of consecutive elements
        // are then independent of the branch and overlap; adding 0.0 for an absent element is exact)
#pragma unroll 4
        for (int p = ne_ptr[v]; p < ne_ptr[v + 1]; ++p) {
            const int e = ne_elem[p];
            const int l = 6 * ne_end[p] + i;
            const double kv = Kg[(size_t)e * 144 + l * 12 + l];
            kii += elem_on(s_mask, e) ? kv : 0.0;
        }
        (DS ? s_dinv : dinv)[r] = kii;
        if (!(fabs(kii) > CM_EPS)) ok = 0;
    }
    if (!ok) atomicOr(&s_red[3], 1);
    __syncthreads();
    K2P(3);
    const int jacobi = s_red[3] ? 0 : 1;
    for (int r = t; r < NT * CM_TILE; r += T) {
        double d = 1.0;
        if (r < ns && jacobi) d = 1.0 / sqrt((DS ? s_dinv : dinv)[r]);
        dinv[r] = d;
        if (DS) s_dinv[r] = d;
    }
    if (t == 0) info[CM_INFO_JACOBI] = jacobi;
    __syncthreads();
    K2P(4);

    // ---- gather assembly.  Per node: the 6x6 diagonal block (lower triangle) and one 6x6 block per
    // lower pair (a,c) - 36 entries each, all entries of the node spread over the lanes of a warp.
    // An entry sums its contributing element entries in ascending element id.
    //
    // One warp per TILE ROW, no block barrier: the warp zero-fills the row's tiles that receive
    // entries (16-byte stores; structurally empty tiles of the envelope are left untouched - the
    // factorisation starts their accumulators from zero) and then scatters the entries of the
    // nodes whose rows fall into this tile row - a node that straddles two tile rows is visited by
    // both owners, each taking its rows.  The 8-byte entry stores meet their freshly zeroed sectors
    // in L1/L2; with a whole-matrix fill first (2.4 MB per CTA, some 900 CTAs resident) the zeros
    // were evicted to DRAM before the scatter and every entry cost a sector read-modify-write there.
    // s_tmp[p] = first solver row of the node at position p of the static order (exclusive scan above).
    {
        const int bt = m.band_tiles;
        const int lane = t & 31;
        K2W_T0();
        // tile rows are handed out dynamically (shared counter): rows differ in work (nodes, neighbours, tiles to fill),
        // with a static round robin a quarter of the warps sat at the barrier below for 5 % of the kernel
#if defined(K2_STATIC_ROWS) && K2_STATIC_ROWS
        for (int IR = t >> 5; IR < NT; IR += T >> 5) {
#else
        for (int IR = next_row(&s_next, lane); IR < NT; IR = next_row(&s_next, lane)) {
#endif
#if K2_BULK_FILL
            // zero fill by the copy engine: lane 0 streams a zeroed 4 KB shared-memory block over every
            // tile of the row that receives entries (cp.async.bulk, two per tile) and waits for the
            // writes to complete before the warp's entry stores follow.  (2.83 ms per 4096 subsets
            // against 3.1 with 16-byte stores from all lanes; issuing the fill one row ahead is slower,
            // 3.29: the zeros must still be in L2 when the entries arrive.)
            if (lane == 0) {
                for (int K = s_kfirst[IR]; K <= IR; ++K) {
                    if (!s_aflag[IR * bt + (K - IR + bt - 1)]) continue;
                    double *dst = tiles + ((size_t)IR * bt + (K - IR + bt - 1)) * CM_TILE_ELEMS;
                    const unsigned src = (unsigned)__cvta_generic_to_shared(s_zero);
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src), "r"(4096) : "memory");
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst + 512), "r"(src), "r"(4096) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            }
#else
            {
                const double2 z2 = make_double2(0.0, 0.0);
                for (int K = s_kfirst[IR]; K <= IR; ++K) {
                    if (!s_aflag[IR * bt + (K - IR + bt - 1)]) continue;
                    double2 *dst = reinterpret_cast<double2 *>(tiles + ((size_t)IR * bt + (K - IR + bt - 1)) * CM_TILE_ELEMS);
#pragma unroll 4
                    for (int i = lane; i < CM_TILE_ELEMS / 2; i += 32) dst[i] = z2;
                }
            }
#endif
            __syncwarp();
            K2W(8);
            // first position whose node ends beyond the first row of the tile row (binary search on the scan)
            const int row0 = IR * CM_TILE, row1 = min(row0 + CM_TILE, ns);
            int lo = 0, hi = nV;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_tmp[mid] + s_free[order[mid]] > row0) hi = mid; else lo = mid + 1;
            }
#if defined(K2_EXP) && K2_EXP == 2          // timing experiment (wrong results): no gather/scatter
            lo = nV;
#endif
#if !defined(K2_NODE_CENTRIC) || !K2_NODE_CENTRIC
            // One lane per matrix ROW of the tile row.  Step s = -1 is the row's part of its node's 6x6 diagonal block (lower
            // triangle), steps s = 0, 1, .. its 6x6 blocks towards the node's lower pairs: in every step all lanes run the
            // same path, whereas one lane per (node, block, row) - one node at a time - left 6 of 32 lanes on the long
            // diagonal path (14.7 active lanes per instruction over the kernel, profiles/r02_experiments.txt).  The 6
            // entries of row i of a block are 48 contiguous bytes of the element's K_G (three 16-byte loads per element);
            // an entry sums its elements in ascending element id as before.
            const int r = row0 + lane;
            int a = -1, i = 0, base_a = 0;
            for (int pa = lo; pa < nV && s_tmp[pa] < row1; ++pa) {            // (uniform: at most 7 nodes touch a tile row)
                const int v = order[pa], nf = s_free[v], b = s_tmp[pa];
                if (r >= b && r < b + nf) {
                    a = v; base_a = b; i = r - b;
                    const int sv = node_sup[v];
                    if (sv >= 0) {                                             // support node with free DOFs: i-th free one
                        const unsigned char *cnd = sup_cond + 6 * sv;
                        int k = i; i = 0;
                        while (cnd[i] || k > 0) { k -= cnd[i] ? 0 : 1; ++i; }
                    }
                }
            }
            __syncwarp();
            K2W(9);
            if (a >= 0) {
                const int p0 = ne_ptr[a], p1 = ne_ptr[a + 1];
                const int q0 = lp_ptr[a], npairs = lp_ptr[a + 1] - q0;
                const double *dv = DS ? s_dinv : dinv;
                const double dr = dv[r];
                const int I = IR;
                auto tile_of = [&](int K) { return tiles + ((size_t)I * bt + (K - I + bt - 1)) * CM_TILE_ELEMS; };
                for (int s = -1; s < npairs; ++s) {
                    double s6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                    int c = a, base_c = base_a;
                    if (s < 0) {
#pragma unroll 2
                        for (int p = p0; p < p1; ++p) {
                            const int e = ne_elem[p];
                            const int o = 6 * ne_end[p];
                            const double2 *kp = reinterpret_cast<const double2 *>(Kg + (size_t)e * 144 + (o + i) * 12 + o);
                            const double2 v0 = kp[0], v1 = kp[1], v2 = kp[2];
                            if (elem_on(s_mask, e)) {
                                s6[0] += v0.x; s6[1] += v0.y; s6[2] += v1.x;
                                s6[3] += v1.y; s6[4] += v2.x; s6[5] += v2.y;
                            }
                        }
                    } else {
                        const int q = q0 + s;
                        c = lp_node[q];
                        if (s_free[c] == 0) continue;
                        base_c = s_base[c];
                        bool any = false;
                        const int pe1 = lp_eptr[q + 1];
                        for (int pe = lp_eptr[q]; pe < pe1; ++pe) {
                            const int e = lp_elem[pe];
                            const int ea = lp_end[pe];
                            const double2 *kp = reinterpret_cast<const double2 *>(Kg + (size_t)e * 144 + (6 * ea + i) * 12 + 6 * (1 - ea));
                            const double2 v0 = kp[0], v1 = kp[1], v2 = kp[2];
                            if (elem_on(s_mask, e)) {
                                s6[0] += v0.x; s6[1] += v0.y; s6[2] += v1.x;
                                s6[3] += v1.y; s6[4] += v2.x; s6[5] += v2.y;
                                any = true;
                            }
                        }
                        if (!any) continue;          // no existing element joins the pair: the tile may be unassembled
                    }
                    // scale and store; columns = free DOFs of node c in order (diagonal block: lower triangle only).
                    // Every index into the value array is a compile-time constant (a run-time index puts the array
                    // into local memory: that was 1.3 GB of L1-missing local loads per 4096 subsets).
                    const int sc = node_sup[c];
                    const int jmax = s < 0 ? i : 5;
                    if (sc < 0) {
                        // node c is not a support: its six DOFs are the columns base_c .. base_c + 5
                        double v[6];
#pragma unroll
                        for (int j = 0; j < 6; ++j) v[j] = (dr * s6[j]) * dv[base_c + j];
                        const int K0 = base_c >> 5;
                        if (ktilde_rm && K0 == ((base_c + jmax) >> 5)) {
                            // row-major K~ tile: the row's entries are contiguous - 16-byte stores where aligned
                            double *dst = tile_of(K0) + (r & 31) * 32 + (base_c & 31);
                            if (base_c & 1) {
                                dst[0] = v[0];
                                if (jmax >= 2) *reinterpret_cast<double2 *>(dst + 1) = make_double2(v[1], v[2]); else if (jmax == 1) dst[1] = v[1];
                                if (jmax >= 4) *reinterpret_cast<double2 *>(dst + 3) = make_double2(v[3], v[4]); else if (jmax == 3) dst[3] = v[3];
                                if (jmax == 5) dst[5] = v[5];
                            } else {
                                if (jmax >= 1) *reinterpret_cast<double2 *>(dst) = make_double2(v[0], v[1]); else dst[0] = v[0];
                                if (jmax >= 3) *reinterpret_cast<double2 *>(dst + 2) = make_double2(v[2], v[3]); else if (jmax == 2) dst[2] = v[2];
                                if (jmax == 5) *reinterpret_cast<double2 *>(dst + 4) = make_double2(v[4], v[5]); else if (jmax == 4) dst[4] = v[4];
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 6; ++j)
                                if (j <= jmax) {
                                    const int cc = base_c + j;
                                    tile_of(cc >> 5)[ktilde_rm ? (r & 31) * 32 + (cc & 31) : cm_tile_off(r & 31, cc & 31)] = v[j];
                                }
                        }
                    } else {
                        // support node with some free DOFs: compact the columns (rare; entry by entry)
                        int cc = base_c;
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            if (sup_cond[6 * sc + j]) continue;        // fixed DOF: no solver column
                            if (j <= jmax) tile_of(cc >> 5)[ktilde_rm ? (r & 31) * 32 + (cc & 31) : cm_tile_off(r & 31, cc & 31)] = (dr * s6[j]) * dv[cc];
                            ++cc;
                        }
                    }
                    if (s < 0) K2W(10);
                }
            }
            __syncwarp();
            K2W(11);
#else
            for (int pa = lo; pa < nV && s_tmp[pa] < row1; ++pa) {
            const int a = order[pa];
            if (s_free[a] == 0) continue;
            const int base_a = s_base[a];
            const int p0 = ne_ptr[a], p1 = ne_ptr[a + 1];
            const int q0 = lp_ptr[a];
            // One lane per (block, row): the 6 entries of row i of a 6x6 block are 48 contiguous bytes of the
            // element's K_G - three 16-byte loads in flight per lane and element instead of one 8-byte
            // load, and the index arithmetic is paid once per six entries.
            const int n_items = 6 * (1 + lp_ptr[a + 1] - q0);
            for (int it = lane; it < n_items; it += 32) {
                const int blk = it / 6, i = it - 6 * blk;
                const int r = node_row(node_sup, sup_cond, a, i, base_a);
                if (r < 0 || (r >> 5) != IR) continue;
                double s6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                int c = a, base_c = base_a;
                if (blk == 0) {
#pragma unroll 2
                    for (int p = p0; p < p1; ++p) {
                        const int e = ne_elem[p];
                        const int o = 6 * ne_end[p];
                        const double2 *kp = reinterpret_cast<const double2 *>(Kg + (size_t)e * 144 + (o + i) * 12 + o);
#if defined(K2_EXP) && K2_EXP == 4          // timing experiment (wrong results): no stiffness-table reads
                        const double2 v0 = make_double2(1.0 + e, 2.0), v1 = v0, v2 = v0; (void)kp;
#else
                        const double2 v0 = kp[0], v1 = kp[1], v2 = kp[2];
#endif
                        if (elem_on(s_mask, e)) {
                            s6[0] += v0.x; s6[1] += v0.y; s6[2] += v1.x;
                            s6[3] += v1.y; s6[4] += v2.x; s6[5] += v2.y;
                        }
                    }
                } else {
                    const int q = q0 + blk - 1;
                    c = lp_node[q];
                    if (s_free[c] == 0) continue;
                    base_c = s_base[c];
                    bool any = false;
                    const int pe1 = lp_eptr[q + 1];
                    for (int pe = lp_eptr[q]; pe < pe1; ++pe) {
                        const int e = lp_elem[pe];
                        const int ea = lp_end[pe];
                        const double2 *kp = reinterpret_cast<const double2 *>(Kg + (size_t)e * 144 + (6 * ea + i) * 12 + 6 * (1 - ea));
#if defined(K2_EXP) && K2_EXP == 4
                        const double2 v0 = make_double2(1.0 + e, 2.0), v1 = v0, v2 = v0; (void)kp;
#else
                        const double2 v0 = kp[0], v1 = kp[1], v2 = kp[2];
#endif
                        if (elem_on(s_mask, e)) {
                            s6[0] += v0.x; s6[1] += v0.y; s6[2] += v1.x;
                            s6[3] += v1.y; s6[4] += v2.x; s6[5] += v2.y;
                            any = true;
                        }
                    }
                    if (!any) continue;          // no existing element joins the pair: the tile may be unassembled
                }
                // scale and store; columns = free DOFs of node c in order (diagonal block: lower triangle only).
                // Every index into the value array is a compile-time constant (a run-time index puts the array
                // into local memory: that was 1.3 GB of L1-missing local loads per 4096 subsets).
                const double *dv = DS ? s_dinv : dinv;
                const double dr = dv[r];
                const int sc = node_sup[c];
                const int jmax = blk == 0 ? i : 5;
                const int I = r >> 5;
                auto tile_of = [&](int K) { return tiles + ((size_t)I * bt + (K - I + bt - 1)) * CM_TILE_ELEMS; };
                if (sc < 0) {
                    // node c is not a support: its six DOFs are the columns base_c .. base_c + 5
                    double v[6];
#pragma unroll
                    for (int j = 0; j < 6; ++j) v[j] = (dr * s6[j]) * dv[base_c + j];
                    const int K0 = base_c >> 5;
                    if (ktilde_rm && K0 == ((base_c + jmax) >> 5)) {
                        // row-major K~ tile: the row's entries are contiguous - 16-byte stores where aligned
#if defined(K2_EXP) && K2_EXP == 5          // timing experiment (wrong results): all entry stores into one 8 KB tile of the slab
                        double *dst = tiles + (r & 31) * 32 + (base_c & 31);       // (same alignment as the real destination)
#else
                        double *dst = tile_of(K0) + (r & 31) * 32 + (base_c & 31);
#endif
                        if (base_c & 1) {
                            dst[0] = v[0];
                            if (jmax >= 2) *reinterpret_cast<double2 *>(dst + 1) = make_double2(v[1], v[2]); else if (jmax == 1) dst[1] = v[1];
                            if (jmax >= 4) *reinterpret_cast<double2 *>(dst + 3) = make_double2(v[3], v[4]); else if (jmax == 3) dst[3] = v[3];
                            if (jmax == 5) dst[5] = v[5];
                        } else {
                            if (jmax >= 1) *reinterpret_cast<double2 *>(dst) = make_double2(v[0], v[1]); else dst[0] = v[0];
                            if (jmax >= 3) *reinterpret_cast<double2 *>(dst + 2) = make_double2(v[2], v[3]); else if (jmax == 2) dst[2] = v[2];
                            if (jmax == 5) *reinterpret_cast<double2 *>(dst + 4) = make_double2(v[4], v[5]); else if (jmax == 4) dst[4] = v[4];
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 6; ++j)
                            if (j <= jmax) {
                                const int cc = base_c + j;
                                tile_of(cc >> 5)[ktilde_rm ? (r & 31) * 32 + (cc & 31) : cm_tile_off(r & 31, cc & 31)] = v[j];
                            }
                    }
                } else {
                    // support node with some free DOFs: compact the columns (rare; entry by entry)
                    int cc = base_c;
#pragma unroll
                    for (int j = 0; j < 6; ++j) {
                        if (sup_cond[6 * sc + j]) continue;        // fixed DOF: no solver column
                        if (j <= jmax) tile_of(cc >> 5)[ktilde_rm ? (r & 31) * 32 + (cc & 31) : cm_tile_off(r & 31, cc & 31)] = (dr * s6[j]) * dv[cc];
                        ++cc;
                    }
                }
            }
            }
#endif
        }
        K2W(12);
        __syncthreads();
        K2P(5);
        // unit diagonal on the padding rows of the last tile
        for (int r = ns + t; r < NT * CM_TILE; r += T) {
            int I = r >> 5;
            tiles[((size_t)I * bt + (bt - 1)) * CM_TILE_ELEMS + (ktilde_rm ? (r & 31) * 33 : cm_tile_off(r & 31, r & 31))] = 1.0;
        }
    }

    // ---- load vectors (numpy_stiffness.py:250-270): P = point + sum_e q_e; rhs = D * P_m
    int zero_bits = 0;
    for (int lc = 0; lc < m.n_lc; ++lc) {
        double *Pl = P + (size_t)lc * 6 * nV;
        double *Xl = X + (size_t)lc * L.n_pad;
        for (int r = ns + t; r < NT * CM_TILE; r += T) Xl[r] = 0.0;
        double sq = 0.0;
        for (int idx = t; idx < 6 * nV; idx += T) {
            int v = idx / 6, i = idx - 6 * v;
            double pv = m.pnodal[(size_t)lc * 6 * nV + idx];
            double qs = 0.0;
#pragma unroll 4
            for (int p = ne_ptr[v]; p < ne_ptr[v + 1]; ++p) {
                const int e = ne_elem[p];
                const double qv = qtab[((size_t)lc * m.nE + e) * 12 + 6 * ne_end[p] + i];
                qs += elem_on(s_mask, e) ? qv : 0.0;
            }
            pv += qs;
            Pl[idx] = pv;
            sq += pv * pv;
            if (s_exist[v]) {
                int r = node_row(node_sup, sup_cond, v, i, s_base[v]);
                if (r >= 0) Xl[r] = (DS ? s_dinv : dinv)[r] * pv;
            }
        }
        // deterministic block reduction of the squared norm
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if ((t & 31) == 0) s_dred[t >> 5] = sq;
        __syncthreads();
        if (t == 0) {
            double tot = 0.0;
            for (int w = 0; w < T / 32; ++w) tot += s_dred[w];
            if (!(sqrt(tot) > CM_EPS)) zero_bits |= 1 << lc;
        }
        __syncthreads();
    }
    if (t == 0) info[CM_INFO_ZERO_LOAD] = zero_bits;
    K2P(6);
    K2P_END();
}

}  // namespace

// Scheduling order of a wave for the factorisation: largest subsets first.  The subsets of a wave
// differ in cost by more than an order of magnitude (20-90 % of the elements, cost ~ n b^2), the
// hardware hands out CTAs in blockIdx order, and the launch ends with whatever was started last -
// so the big ones go first and the tail of the launch is made of short CTAs.  Rank by counting
// (stable: ties keep slab order), deterministic, no atomics: thread i compares its key with every
// other key of the wave (staged through shared memory) and writes its slab index into record
// `rank`.  count^2 integer comparisons - microseconds for the wave sizes in use.
namespace {
__global__ void __launch_bounds__(256) k2_schedule(cm_ws_layout L, char *ws, int count) {
    __shared__ long long s_key[256];
    const int i = blockIdx.x * 256 + threadIdx.x;
    auto info_of = [&](int s) { return reinterpret_cast<int *>(ws + (size_t)s * L.per_subset + L.off_info); };
    // key: cost, with the first subset of every prefix-sharing group (the trunk) ahead of all others - the trunks are
    // factored by an earlier launch, the others may read their rows.  (Without groups only slab 0 is "a trunk": harmless.)
    auto key_of = [&](int s) -> long long {
        const int *inf = info_of(s);
        return (long long)inf[CM_INFO_COST] + (inf[CM_INFO_TRUNK] == s ? (1ll << 32) : 0ll);      // first subset of its group
    };
    const long long mine = i < count ? key_of(i) : -1;
    int rank = 0;
    for (int j0 = 0; j0 < count; j0 += 256) {
        const int j = j0 + threadIdx.x;
        __syncthreads();
        s_key[threadIdx.x] = j < count ? key_of(j) : -1;
        __syncthreads();
        const int n = min(256, count - j0);
        for (int k = 0; k < n; ++k) {
            const long long key = s_key[k];
            rank += (key > mine || (key == mine && j0 + k < i)) ? 1 : 0;
        }
    }
    if (i < count) info_of(rank)[CM_INFO_SCHED] = i;
}
}  // namespace

// read (and clear) the phase counters of the instrumented build (zeros otherwise)
cudaError_t cm_k2_read_prof(unsigned long long *out8) {
    cudaError_t e = cudaMemcpyFromSymbol(out8, g_k2_prof, 16 * sizeof(unsigned long long));
    if (e != cudaSuccess) return e;
    unsigned long long z[16] = {0};
    return cudaMemcpyToSymbol(g_k2_prof, z, sizeof(z));
}

cudaError_t cm_launch_schedule(const cm_ws_layout &L, char *ws, int count, cudaStream_t st) {
    k2_schedule<<<(count + 255) / 256, 256, 0, st>>>(L, ws, count);
    return cudaGetLastError();
}

cudaError_t cm_launch_dofmap_assemble(const cm_model_dev &m, const cm_ws_layout &L, char *ws,
                                      const uint32_t *masks, const int *geom, const int *trows, int group, int part, long long first, int count,
                                      const cm_out_dev &out, int ktilde_rowmajor, cudaStream_t st) {
    const size_t base = (size_t)(2 * m.nV + 2 * m.n_tile_rows_max + K2_THREADS + 1 + m.mask_words) * 4 + 2 * (size_t)m.nV +
                        (size_t)m.n_tile_rows_max * m.band_tiles + 16;
    // index tables and the Jacobi scale in shared memory while K2_MIN_CTAS CTAs still fit an SM (the 1k-element frame needs
    // 35 KB + 16.6 KB per CTA)
    const size_t budget = (size_t)(227 * 1024) / K2_MIN_CTAS - 1024 - 4500;      // per CTA: 1 KB reserved, 4.4 KB static
    const size_t blob = ((size_t)m.k2_blob_bytes + 15) & ~(size_t)15;
    const size_t dsz = (size_t)L.n_pad * sizeof(double);
    // the index tables first (2.32 against 2.57 ms per 4096 subsets), the Jacobi scale if it still fits (2.52 alone)
    bool ts = base + 16 + blob <= budget;
    bool ds = base + 16 + (ts ? blob : 0) + dsz <= budget;
    static const int force = getenv("CM_K2_SMEM") ? atoi(getenv("CM_K2_SMEM")) : -1;      // timing experiments: bit 0 tables, bit 1 scale
    if (force >= 0) { ts = force & 1; ds = (force & 2) != 0; }
    const size_t smem = base + 16 + (ts ? blob : 0) + (ds ? dsz : 0);
    static cm_smem_cache cache[4];
    auto go = [&](auto kernel, cm_smem_cache &c) -> cudaError_t {
        cudaError_t e = cm_ensure_smem(kernel, smem, c, m.device);
        if (e != cudaSuccess) return e;
        const int ctas = part == 1 ? (count + group - 1) / group : count;
        kernel<<<ctas, K2_THREADS, smem, st>>>(m, L, ws, masks, geom, trows, group, part, first, out, ktilde_rowmajor);
        return cudaGetLastError();
    };
    if (ts && ds) return go(k2_dofmap_assemble<true, true>, cache[3]);
    if (ts) return go(k2_dofmap_assemble<true, false>, cache[2]);
    if (ds) return go(k2_dofmap_assemble<false, true>, cache[1]);
    return go(k2_dofmap_assemble<false, false>, cache[0]);
}
