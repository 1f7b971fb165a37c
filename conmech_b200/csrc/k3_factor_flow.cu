// K3 (default) - batched fp64 LDL^T of the Jacobi-scaled K_mm in block-band tile storage on the
// FP64 tensor cores, with pivot-based mechanism detection and the forward substitution fused in.
//
// Replaces the reference's SuperLU factorisation + definiteness test + first triangular solve
// (numpy_stiffness.py:299-336, :353).  The reference factors with diagonal pivots only
// (diag_pivot_thresh=0, SymmetricMode), raises on an exactly zero pivot (:301-319) and rejects any
// pivot <= -1e-14 (:326-336); the same rules are applied here to the pivots d_j of the LDL^T.
// Off-diagonal tiles are solved as a product with M_J = D_J^-1 L_JJ^-1 (one more tensor-core tile
// product instead of a scalar triangular solve).  Tiles are stored fragment-major (cm_tile_off), so
// accumulators and operands move with vector loads/stores straight between global memory (L1/L2
// resident window) and MMA registers.  No block-wide barrier on the hot loop:
//
// One CTA per subset, NW "row" warps.  Warp w owns tile rows I = w, w+NW, ... and walks each row
// left to right: tile (I,J) accumulates A(I,J) - sum_K L(I,K) D_K L(J,K)^T as one linear operand
// stream, is solved against M_J, and the row ends with its diagonal tile (SYRK stream, LDL^T,
// inverse) and the forward substitution of the row.  The only synchronisation is one progress
// counter per tile row in shared memory: prog[J] = number of finished tile columns of row J
// (J+1 once its diagonal factor and M_J are published).  Term K of tile (I,J) may run once
// prog[J] > K; the solve once prog[J] == J+1.  Rows therefore advance as a staggered wavefront, the
// serial diagonal factorisations of consecutive rows overlap with the other rows' tensor-core work,
// and every warp always has a long stream of independent MMAs queued.  The critical path per row
// is  M_{I-1} -> solve of tile (I,I-1) -> newest SYRK term -> LDL^T + inverse of tile (I,I);  all
// older terms of those two tiles are accumulated before the warp starts waiting (look-ahead).
#include "cm_factor_common.cuh"

namespace {

__device__ __forceinline__ bool t_is_zero() { return threadIdx.x == 0; }

// instrumented build: cycles per activity, summed over all row warps (slots: 0 wait for an update
// operand, 1 update streams, 2 wait for M_J, 3 solve product, 4 diagonal-tile updates, 5 wait for the
// previous diagonal factor, 6 diagonal factor, 7 wait for the forward chain, 8 forward substitution,
// 9 total, 10 solve product of the chain tile (I,I-1), 11 newest diagonal term, 12 gap between the
// publication of M_{I-1} and the owner of row I noticing it)
__device__ unsigned long long g_flow_prof[16];
// (PROF is a template parameter of every function that uses these: the production build carries no trace of them)
#define FP_T0(pc) long long _t0 = (PROF && (pc)) ? clock64() : 0
#define FP_ADD(pc, slot) do { if (PROF && (pc)) { long long _n = clock64(); if ((threadIdx.x & 31) == 0) (pc)[slot] += _n - _t0; _t0 = _n; } } while (0)

// tile (I,J), J < I: all update terms as they become available, then - still in registers - the
// solve against M_J, the tile's share of the forward substitution and, for the chain tile
// (J == I-1), the newest term of the diagonal tile, handed to the diagonal factorisation through
// the shared scratch.  Returns true when the diagonal tile was placed in the scratch.
//
// Look-ahead builds (LA): a visit may stop with the accumulator PARKED in tensor memory (slot J & 1 of the warp's lanes).
//   Kfrom < 0: fresh visit; Kfrom >= 0: resume - the accumulator comes back from the slot, terms < Kfrom are in it.
//   Kstop:     accumulate the terms K < min(Kstop, J).
//   LA_AHEAD:  look-ahead call - park after those terms and return LA_K0 + (terms done); LA_NONE if there was nothing to do.
//   LA_TRY:    after ALL terms: if M_J is not published yet, park and return LA_PARKED instead of waiting for it.
// Return otherwise: 1 = chain tile (the diagonal tile was placed in the scratch), 0 = done.
enum { LA_AHEAD = 1, LA_TRY = 2, LA_PARKED = 2, LA_NONE = 3, LA_K0 = 16 };
#ifndef CM_TILE_SMEM
#define CM_TILE_SMEM 1
#endif
// a tile visit's own arguments, per warp: written at entry and re-read behind the update streams, so that they are not
// live (= spilled to local memory, which misses L1 under the streams) while accumulator and operand buffers fill the
// register file
__shared__ int s_tilestate[16 * 4];
#ifndef CM_LA_FAR
#define CM_LA_FAR 0
#endif
#ifndef CM_LA_CHUNK
#define CM_LA_CHUNK 0
#endif
template <bool PROF, int OPS, bool LA>
__device__ __noinline__ int offdiag_tile(int I, int J, int kcI, int lane, int Kfrom, int Kstop, int la_flags) {
#if CM_TILE_SMEM
    volatile int *ts = s_tilestate + (threadIdx.x >> 5) * 4;
    ts[0] = I; ts[1] = J; ts[2] = kcI; ts[3] = la_flags;
#endif
    const int bt = CTX(bt);
    long long *pc = PROF ? CTX(pc) + (threadIdx.x >> 5) * 16 : nullptr;
    double *T = cm_tile_ptr(CTX(tiles), bt, I, J);
    const int ab = J - I + bt - 1, aw = (bt + 63) >> 6;
    const unsigned long long *arow = CTX(arow);
    const bool has_a = (arow[I * aw + (ab >> 6)] >> (ab & 63)) & 1ull;         // K2 assembled entries into this tile
    const unsigned slot = LA ? CTX(tmem) + ((unsigned)(threadIdx.x & ~31u) << 16) + 64u * (unsigned)(J & 1) : 0u;
    const bool resume = LA && Kfrom >= 0, ahead = LA && (la_flags & LA_AHEAD);
    const int Kto = LA ? min(Kstop, J) : J;
    // the next tile of the row that K2 assembled (the diagonal tile always is) comes from DRAM: start that read now
    if (!resume && !ahead && ((arow[I * aw + ((ab + 1) >> 6)] >> ((ab + 1) & 63)) & 1ull)) {
        const double *nx = T + CM_TILE_ELEMS + lane * 16;
        asm volatile("prefetch.global.L2 [%0];" :: "l"(nx));
        asm volatile("prefetch.global.L2 [%0];" :: "l"(nx + 512));
    }
    // operand columns left of k-step qs are zero in row I or in row J (envelope): the stream starts there
    int qs = resume ? Kfrom * 8 : max(kcI, CTX(kcol)[J]);
    int K = qs >> 3;
    FP_T0(pc);
    const int trows = CTX(trows);                   // rows < trows live in the trunk's slab
    TileAcc acc;
    if (resume) {
        tm_unpark(slot, acc);
    } else {
        if (K >= J && !has_a && J != I - 1) {         // no entries and no fill-in: the tile of L is zero (the chain tile still hands the diagonal over)
            if (ahead) return LA_NONE;                // (left to the regular visit)
#pragma unroll 1
            for (int i = 0; i < 8; ++i) stg4(T + i * 128 + lane * 4, 0.0, 0.0, 0.0, 0.0);
            return 0;
        }
        if (ahead && Kto <= K) return LA_NONE;        // no term of this tile can run yet
        if (has_a) acc_load_rm(acc, T, lane); else acc_zero(acc);      // assembled K~ tiles are row-major (K2)
    }
    while (K < Kto) {
        int Kend = min(wait_prog_gt(CTX(prog), J, K), Kto);
        if (K < trows && Kend > trows) Kend = trows;           // the pivots of the shared columns come from the trunk
        if (CM_LA_CHUNK && ahead) Kend = K + 1;                // one operand tile at a time, then look at M_{J-1} again
        FP_ADD(pc, 0);
        const double *tiles = CTX(tiles);
        const double *A = tiles + ((size_t)I * bt + (bt - 1 - I)) * CM_TILE_ELEMS + (size_t)qs * 128;   // tile (I,0) + qs k-steps
        const double *Bm = (J < trows ? CTX(ttiles) : tiles) + ((size_t)J * bt + (bt - 1 - J)) * CM_TILE_ELEMS + (size_t)qs * 128;
        const double *dq = (K < trows ? CTX(tdpiv) : CTX(dpiv)) + qs * 4;
        if (OPS) {
            const int w = threadIdx.x >> 5;
            unsigned *itp = CTX(ring_it) + w;
            OpRing R{CTX(ring0) + w * (OPS * CM_RING_STAGE_BYTES), CTX(ringbar0) + w * (OPS * 8), *itp};
            acc_update_stream_ring<OPS ? OPS : 1>(acc, A, Bm, dq, Kend * 8 - qs, lane, R);
            if (lane == 0) *itp = R.it;
            __syncwarp();
        } else
            acc_update_stream<false>(acc, A, Bm, dq, Kend * 8 - qs, lane);
        K = Kend;
        qs = K * 8;
        FP_ADD(pc, 1);
#if CM_LA_CHUNK
        // look-ahead call: the tile that was parked for this one (I, J-1) can go on as soon as M_{J-1} is there
        if (ahead && K < Kto && flag_acquire(CTX(prog) + J - 1) > J - 1) break;
#endif
    }
#if CM_TILE_SMEM
    I = ts[0]; J = ts[1]; kcI = ts[2]; la_flags = ts[3]; lane = threadIdx.x & 31;
#endif
    if (LA) {
        if (ahead) { tm_park(slot, acc); return LA_K0 + K; }
        // M_J not there yet: let the caller start the next tile of the row meanwhile - if row J is far enough from publishing
        // it (CM_LA_FAR tile columns short of its diagonal factor; 0 = any wait) that a round trip through tensor memory pays
        if ((la_flags & LA_TRY) && flag_acquire(CTX(prog) + J) <= J - CM_LA_FAR) { tm_park(slot, acc); return LA_PARKED; }
    }
    wait_prog_gt(CTX(prog), J, J);                // M_J published
    FP_ADD(pc, 2);
    if (PROF && pc && J == I - 1 && (threadIdx.x & 31) == 0) pc[12] += clock64() - *CTX(tpub);     // reaction gap on the chain
    acc_solve_inplace(acc, (J < trows ? CTX(tminv) : CTX(minv)) + (size_t)J * CM_TILE_ELEMS, max(kcI - 8 * J, 0) >> 1, lane);
    acc_store(acc, cm_tile_ptr(CTX(tiles), bt, I, J), lane);
    if (PROF && pc && J == I - 1) FP_ADD(pc, 10); else FP_ADD(pc, 3);
    // share of this tile in the forward substitution of row I:  s_I += L(I,J) y_J
    {
        volatile int *yprog = CTX(yprog);
        int y = flag_acquire(yprog);
        while (y <= J) { __nanosleep(CM_SLEEP_NS); y = flag_acquire(yprog); }
    }
    FP_ADD(pc, 7);
    double2 dk[4];                                // pivots of block J for the diagonal term below: in flight with the y loads
    {
        const double *dJ = (J < trows ? CTX(tdpiv) : CTX(dpiv)) + J * CM_TILE + 2 * (lane & 3);
#pragma unroll
        for (int cs = 0; cs < 4; ++cs) dk[cs] = ldg2(dJ + cs * 8);
    }
    {
        const int n_lc = CTX(n_lc), zero_bits = CTX(zero_bits);
        double *s_part = CTX(s_part) + (threadIdx.x >> 5) * n_lc * 32;
        for (int lc = 0; lc < n_lc; ++lc) {
            if ((zero_bits >> lc) & 1) continue;
            double part[4];
            acc_matvec(acc, (J < trows ? CTX(tX) : CTX(X)) + (size_t)lc * CTX(n_pad) + J * CM_TILE, part, lane);
            if ((lane & 3) == 0) {
#pragma unroll
                for (int ri = 0; ri < 4; ++ri) s_part[lc * 32 + ri * 8 + (lane >> 2)] += part[ri];
            }
        }
    }
    FP_ADD(pc, 8);
    // this tile's term of the diagonal tile of the row, from the same registers (pending sum in shared
    // memory); the chain tile completes the diagonal tile and hands it to the factorisation
    const bool chain = J == I - 1;
    diag_term(acc, chain, dk, CTX(s_scr),
#if defined(CM_EXP_DACC1)
              CTX(s_dacc),
#else
              CTX(s_dacc) + (threadIdx.x >> 5) * 320,
#endif
              max(kcI - 8 * J, 0) >> 1, lane);
    if (chain) FP_ADD(pc, 11); else FP_ADD(pc, 4);
    return chain;
}

// OPS: stages of the shared-memory operand ring per warp (0 = operands through registers, one k-step of look-ahead)
template <int NW, bool PROF, int OPS, bool LA = false>
#ifndef CM_K3_CTAS4
#define CM_K3_CTAS4 4          // resident 4-warp CTAs per SM the register allocation aims at (timing experiments: 5)
#endif
#ifndef CM_K3_RING_CTAS
#define CM_K3_RING_CTAS 3       // the same for the operand-ring builds
#endif
__global__ void __launch_bounds__(NW * 32, NW == 1 ? 12 : (NW == 4 ? (OPS ? CM_K3_RING_CTAS : CM_K3_CTAS4) : (NW <= 16 ? 16 / NW : 1)))
k3_factor_flow(cm_model_dev m, cm_ws_layout L, char *ws, int rank0, int stride) {
    extern __shared__ __align__(16) double smem[];
    double *s_scr = smem;                                            // [DSCR] diagonal scratch (shared by the CTA)
    double *s_rhs = s_scr + DSCR;                                    // [NW][32] forward substitution scratch
    double *s_part = s_rhs + NW * 32;                                // [NW][n_lc][32] pending forward sums of the rows in progress
    double *s_minpiv = s_part + NW * m.n_lc * 32;                    // [NW]
    double2 *s_dacc = reinterpret_cast<double2 *>(s_minpiv + NW + (NW & 1));   // [NW][10][32] pending diagonal-tile sums (16-byte aligned)
#if defined(CM_EXP_DACC1)       // timing experiment (wrong results): one pending-sum region for all warps
    int *s_prog = reinterpret_cast<int *>(s_dacc + 320);
#else
    int *s_prog = reinterpret_cast<int *>(s_dacc + NW * 320);        // [n_tile_rows_max] factor progress
#endif
    int *s_yprog = s_prog + m.n_tile_rows_max;                       // rows forward-substituted
    int *s_fail = s_yprog + 1;
    int *s_kcol = s_yprog + 4;                                       // [n_tile_rows_max] envelope: first nonzero k-step of each row
    unsigned long long *s_arow = reinterpret_cast<unsigned long long *>(s_kcol + m.n_tile_rows_max);   // 8-byte aligned: 2n+4 ints after s_prog
                                                                     // [n_tile_rows_max][aw] bit b: tile (I, I-bt+1+b) was assembled by K2
    const int aw = (m.band_tiles + 63) >> 6;
    // operand ring: [NW][OPS] stages of CM_RING_STAGE_BYTES (128-byte aligned), then [NW][OPS] mbarriers, then [NW] counters
    unsigned char *s_ring = reinterpret_cast<unsigned char *>((reinterpret_cast<size_t>(s_arow + m.n_tile_rows_max * aw) + 127) & ~(size_t)127);
    unsigned long long *s_ringbar = reinterpret_cast<unsigned long long *>(s_ring + (size_t)NW * OPS * CM_RING_STAGE_BYTES);
    unsigned *s_ring_it = reinterpret_cast<unsigned *>(s_ringbar + NW * OPS);
    __shared__ long long s_pc[PROF ? NW : 1][16];
    __shared__ long long s_tpub;
    __shared__ unsigned s_tmem;
    __shared__ int s_rowstate[NW * 4];       // the tile loop's state across the tile calls (below)

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // CTA r factors the subset with the r-th largest cost of the wave (k2_schedule): big ones first
    const int sb = stride > 0 ? (int)blockIdx.x * stride
                              : reinterpret_cast<const int *>(ws + (size_t)(rank0 + blockIdx.x) * L.per_subset + L.off_info)[CM_INFO_SCHED];
    char *slab = ws + (size_t)sb * L.per_subset;
    int *info = reinterpret_cast<int *>(slab + L.off_info);
    if (info[CM_INFO_STATUS] != CM_ST_OK) return;
    const int NT = info[CM_INFO_NT];
    // prefix-sharing group: the first trows tile rows are the trunk's (factored by the previous launch) - if the trunk
    // is usable: solved, or failed behind the shared rows; same Jacobi-scaling decision (a zero diagonal anywhere switches
    // the scaling off for the whole matrix, numpy_stiffness.py:283-290); at least trows complete tile rows.  Otherwise the
    // subset is factored on its own, exactly as without sharing (K2 has assembled all of it).
    int trows = info[CM_INFO_TROWS];
    const char *tslab = ws + (size_t)info[CM_INFO_TRUNK] * L.per_subset;
    if (trows > 0) {
        const int *tinfo = reinterpret_cast<const int *>(tslab + L.off_info);
        const int tst = tinfo[CM_INFO_STATUS], tfail = tinfo[CM_INFO_FAILROW];
        const bool usable = (tst == CM_ST_OK || (tst == CM_ST_SINGULAR && tfail - 1 >= trows)) &&
                            tinfo[CM_INFO_JACOBI] == info[CM_INFO_JACOBI] && tinfo[CM_INFO_NS] / CM_TILE >= trows;
        if (!usable) {
            trows = 0;
            if (t_is_zero()) info[CM_INFO_TROWS] = 0;          // (K4 reads it)
        }
    }
    double *tiles = reinterpret_cast<double *>(slab + L.off_tiles);
    const int *kfirst = reinterpret_cast<const int *>(slab + L.off_kfirst);
    const int *kcol = reinterpret_cast<const int *>(slab + L.off_kcol);
    const unsigned char *aflag = reinterpret_cast<const unsigned char *>(slab + L.off_aflag);
    double *dpiv = reinterpret_cast<double *>(slab + L.off_dpiv);
    double *minv = reinterpret_cast<double *>(slab + L.off_minv);
    double *X = reinterpret_cast<double *>(slab + L.off_x);
    const int bt = m.band_tiles;
    const int n_lc = m.n_lc;
    const int zero_bits = info[CM_INFO_ZERO_LOAD];
    volatile int *prog = s_prog;
    volatile int *yprog = s_yprog;

    if (LA) {
        // 128 columns of tensor memory per CTA (four CTAs per SM = all 512): two parked accumulator tiles per warp
        static_assert(!LA || NW == 4, "one lane quarter of the allocation per warp");
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(&s_tmem)), "n"(128) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (t == 0) {
        s_ctx.tmem = LA ? s_tmem : 0u;
        s_ctx.tiles = tiles; s_ctx.minv = minv; s_ctx.dpiv = dpiv; s_ctx.X = X;
        s_ctx.kcol = s_kcol; s_ctx.arow = s_arow; s_ctx.prog = s_prog; s_ctx.yprog = s_yprog;
        s_ctx.s_part = s_part; s_ctx.s_rhs = s_rhs; s_ctx.s_scr = s_scr; s_ctx.s_minpiv = s_minpiv; s_ctx.s_fail = s_fail; s_ctx.s_dacc = s_dacc;
        s_ctx.pc = PROF ? &s_pc[0][0] : nullptr; s_ctx.tpub = &s_tpub;
        s_ctx.bt = bt; s_ctx.n_pad = L.n_pad; s_ctx.n_lc = n_lc; s_ctx.zero_bits = zero_bits;
        s_ctx.trows = trows;
        s_ctx.ttiles = reinterpret_cast<const double *>(tslab + L.off_tiles);
        s_ctx.tminv = reinterpret_cast<const double *>(tslab + L.off_minv);
        s_ctx.tdpiv = reinterpret_cast<const double *>(tslab + L.off_dpiv);
        s_ctx.tX = reinterpret_cast<const double *>(tslab + L.off_x);
        if (OPS) {
            s_ctx.ring0 = (unsigned)__cvta_generic_to_shared(s_ring);
            s_ctx.ringbar0 = (unsigned)__cvta_generic_to_shared(s_ringbar);
            s_ctx.ring_it = s_ring_it;
            for (int i = 0; i < NW * OPS; ++i) mbar_init((unsigned)__cvta_generic_to_shared(s_ringbar + i), 1);
            for (int i = 0; i < NW; ++i) s_ring_it[i] = 0;
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    for (int I = t; I < NT; I += NW * 32) {
        s_prog[I] = I < trows ? I + 1 : kfirst[I];        // shared rows are complete
        s_kcol[I] = kcol[I];
        for (int w = 0; w < aw; ++w) {
            unsigned long long bits = 0;
#pragma unroll 1
            for (int b = 64 * w; b < bt && b < 64 * w + 64; ++b) bits |= (unsigned long long)(aflag[I * bt + b] != 0) << (b & 63);
            s_arow[I * aw + w] = bits;
        }
    }
    if (t == 0) { *s_yprog = trows; *s_fail = 0; }
    if (t < NW) s_minpiv[t] = __longlong_as_double(0x7ff0000000000000LL);
    __syncthreads();

    long long *pc = PROF ? s_pc[warp] : nullptr;
    if (PROF && lane == 0) for (int i = 0; i < 16; ++i) pc[i] = 0;
    const long long tstart = PROF ? clock64() : 0;
    double *s_part_w = s_part + warp * n_lc * 32;
#if defined(CM_EXP_DACC1)
    double2 *s_dacc_w = s_dacc;
#else
    double2 *s_dacc_w = s_dacc + warp * 320;
#endif
    for (int I = warp + (trows + NW - 1 - warp) / NW * NW; I < NT; I += NW) {      // this warp's first row >= trows
        const int kcI = s_kcol[I];               // first nonzero column of the row in k-steps (even)
        const int kfI = kcI >> 3;                // = kfirst[I], the first tile column of the row
        for (int lc = 0; lc < n_lc; ++lc) s_part_w[lc * 32 + lane] = 0.0;
        if (I + NW < NT) {                       // the diagonal tile of this warp's next row comes from DRAM: request it now
            const double *nx = cm_tile_ptr(tiles, bt, I + NW, I + NW) + lane * 16;
            asm volatile("prefetch.global.L2 [%0];" :: "l"(nx));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(nx + 512));
        }
        if (kfI < I) {                           // pending sum of the diagonal tile (diag_term) starts from the assembled tile
            const double *pd = cm_tile_ptr(tiles, bt, I, I) + (lane >> 2) * 32 + 2 * (lane & 3);      // row-major K~ (K2)
#pragma unroll
            for (int cj = 0; cj < 4; ++cj)
#pragma unroll
                for (int ri = cj; ri < 4; ++ri) {
                    double2 v;
                    asm volatile("ld.global.L1::evict_first.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(pd + ri * 256 + cj * 8) : "memory");
                    s_dacc_w[(ri * (ri + 1) / 2 + cj) * 32 + lane] = v;
                }
        }
        __syncwarp();
        bool handed = false;                     // diagonal tile already in the scratch (chain tile hand-over)
        if (LA) {
            // Look-ahead over the waits for M_J.  At a band of NW..2 NW tiles nearly every tile of a row depends on a row that
            // is still being factored, and a warp that waits for M_J holds 64 accumulator registers idle (10 % of the kernel).
            // Instead the finished accumulator is parked in tensor memory and the NEXT tile of the row runs the terms that are
            // available (all but the one with the parked tile), is parked in the other slot, and the first is taken up again.
            int pend = -1;                       // >= 0: tile J was started ahead, terms < pend are in its parked accumulator
            for (int J = kfI; J < I; ++J) {
                int r = offdiag_tile<PROF, OPS, LA>(I, J, kcI, lane, pend, J, J + 1 < I ? LA_TRY : 0);
                pend = -1;
                if (r == LA_PARKED) {
                    const int avail = min(J, flag_acquire(prog + J + 1));        // terms K < J of tile (I, J+1) that row J+1 has reached
                    const int r2 = offdiag_tile<PROF, OPS, LA>(I, J + 1, kcI, lane, -1, avail, LA_AHEAD);
                    if (r2 >= LA_K0) pend = r2 - LA_K0;
                    r = offdiag_tile<PROF, OPS, LA>(I, J, kcI, lane, J, J, 0);  // back to tile J: wait for M_J, finish
                }
                handed = r == 1;
                release_fence<OPS>();
                __syncwarp();
                if (lane == 0) prog[I] = J + 1;
            }
        } else {
            // The tile loop keeps its state (row, envelope start, tile column) in SHARED memory and re-reads it after every
            // call: the tile routine uses all 128 registers, so whatever the loop holds across the call is saved to local
            // memory and comes back through L1 - which the operand streams keep flushing (31.6 M L1-missing local-load
            // sectors per 4096 subsets, a dependent L2 round trip each).  136 -> 32 bytes of spill stores, 9.89 -> 9.72 ms
            // per 4096 subsets.  (The whole row loop rewritten that way measured 9.84: only this loop.)
            volatile int *rw = s_rowstate + (threadIdx.x >> 5) * 4;
            rw[0] = I; rw[1] = kcI;
            for (rw[2] = kfI; rw[2] < rw[0]; rw[2] = rw[2] + 1) {
                handed = offdiag_tile<PROF, OPS, false>(rw[0], rw[2], rw[1], threadIdx.x & 31, -1, rw[2], 0) == 1;
                release_fence<OPS>();
                __syncwarp();
                if ((threadIdx.x & 31) == 0) CTX(prog)[rw[0]] = rw[2] + 1;
            }
        }
        FP_T0(pc);
        if (!handed && I > 0) wait_prog_gt(prog, I - 1, I - 1);      // one diagonal factorisation at a time (shared scratch)
        FP_ADD(pc, 5);
        diag_factor_blk(I, handed, lane);
        release_fence<OPS>();
        __syncwarp();
        if (lane == 0) { if (PROF) s_tpub = clock64(); prog[I] = I + 1; }
        FP_ADD(pc, 6);
        // forward substitution of row I once rows < I are done (they finish one chain step earlier)
        {
            int y = flag_acquire(yprog);
            while (y < I) { __nanosleep(CM_SLEEP_NS); y = flag_acquire(yprog); }
        }
        FP_ADD(pc, 7);
        forward_row_w(I, lane);
        release_fence<OPS>();
        __syncwarp();
        if (lane == 0) *yprog = I + 1;
        FP_ADD(pc, 8);
    }
    if (PROF && lane == 0) {
        pc[9] = clock64() - tstart;
        for (int i = 0; i < 16; ++i) atomicAdd(&g_flow_prof[i], (unsigned long long)pc[i]);
    }
    __syncthreads();
    if (t == 0) {
        double mp = s_minpiv[0];
        for (int w = 1; w < NW; ++w) mp = fmin(mp, s_minpiv[w]);
        if (*s_fail) { info[CM_INFO_STATUS] = CM_ST_SINGULAR; info[CM_INFO_FAILROW] = *s_fail; }
        if (trows > 0) mp = fmin(mp, reinterpret_cast<const double *>(tslab + L.off_dpiv)[L.n_pad]);
        dpiv[L.n_pad] = mp;            // slot after the pivots
    }
    if (LA && warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem), "n"(128) : "memory");
}

size_t flow_smem(const cm_model_dev &m, int nw, int ops) {
#if defined(CM_EXP_DACC1)
    const int ndacc = 1;
#else
    const int ndacc = nw;
#endif
    return (size_t)(DSCR + nw * 32 + nw * std::max(1, m.n_lc) * 32 + nw + (nw & 1) + ndacc * 640) * sizeof(double) +
           (size_t)(2 * m.n_tile_rows_max + 8) * sizeof(int) + (size_t)m.n_tile_rows_max * ((m.band_tiles + 63) >> 6) * sizeof(unsigned long long) +
           (ops ? 128 + (size_t)nw * ops * (CM_RING_STAGE_BYTES + 8) + (size_t)nw * 4 : 0);
}

}  // namespace

template <int NW, bool PROF, int OPS, bool LA = false>
static cudaError_t launch_flow(const cm_model_dev &m, const cm_ws_layout &L, char *ws, int rank0, int count, int stride, cudaStream_t st) {
    static cm_smem_cache cache;
    const size_t smem = flow_smem(m, NW, OPS);
    cudaError_t e = cm_ensure_smem(k3_factor_flow<NW, PROF, OPS, LA>, smem, cache, m.device);
    if (e != cudaSuccess) return e;
    k3_factor_flow<NW, PROF, OPS, LA><<<count, NW * 32, smem, st>>>(m, L, ws, rank0, stride);
    return cudaGetLastError();
}

// config: 0 = automatic, 1 = 8 warps x 2 CTAs/SM, 2 = 4 warps x 4 CTAs/SM, 3 = 16 warps x 1 CTA/SM.
// Automatic: four small CTAs per SM hide the serial chain of a subset behind other subsets and are
// best whenever the launch fills the GPU; a launch with few subsets of a wide band (the 10k-element
// frames) instead needs its parallelism INSIDE the subset - more tile rows in flight per CTA.
cudaError_t cm_launch_factor_flow(const cm_model_dev &m, const cm_ws_layout &L, char *ws, int rank0, int count, int stride,
                                  int config, int prof, cudaStream_t st) {
    if (config == 0) {
        const int n_sm = m.n_sm > 0 ? m.n_sm : 148;
        config = 2;
        if (count <= 2 * n_sm && m.band_tiles > 8) config = 1;
        if (m.band_tiles > 16) config = 3;       // config 5 (band of 35 tiles), 512 subsets, largest first: 16 warps 142 ms, 8 warps 147 ms, 4 warps 209 ms
    }
    // operand ring (cm_set_variant 10 / 11: 4-warp CTAs with one / two stages per warp; 12: 16-warp CTAs, two stages): CM_K3_RING=1 makes the 16-warp
    // configuration use it by default (one CTA per SM: the shared memory is there)
    if (config == 10) return launch_flow<4, false, 1>(m, L, ws, rank0, count, stride, st);     // 59 KB per CTA: three CTAs per SM
    if (config == 11) return launch_flow<4, false, 2>(m, L, ws, rank0, count, stride, st);     // 76 KB: three CTAs per SM just fit
    if (config == 12) return launch_flow<8, false, 2>(m, L, ws, rank0, count, stride, st);      // 8-warp CTAs (wide bands)
    if (config == 13) return launch_flow<4, false, 0, true>(m, L, ws, rank0, count, stride, st);   // look-ahead over the M_J waits, accumulators parked in tensor memory
    static const int ring16 = getenv("CM_K3_RING") ? atoi(getenv("CM_K3_RING")) : 0;
    if (config == 1) return prof ? launch_flow<8, true, 0>(m, L, ws, rank0, count, stride, st) : launch_flow<8, false, 0>(m, L, ws, rank0, count, stride, st);
    if (config == 3) return ring16 ? launch_flow<8, false, 2>(m, L, ws, rank0, count, stride, st) : launch_flow<16, false, 0>(m, L, ws, rank0, count, stride, st);
    if (config == 4) return launch_flow<2, false, 0>(m, L, ws, rank0, count, stride, st);
    if (config == 5) return launch_flow<3, false, 0>(m, L, ws, rank0, count, stride, st);
    if (config == 6) return launch_flow<1, false, 0>(m, L, ws, rank0, count, stride, st);
    return prof ? launch_flow<4, true, 0>(m, L, ws, rank0, count, stride, st) : launch_flow<4, false, 0>(m, L, ws, rank0, count, stride, st);
}

// read (and clear) the activity counters of the instrumented build
cudaError_t cm_factor_flow_read_prof(unsigned long long *out16) {
    cudaError_t e = cudaMemcpyFromSymbol(out16, g_flow_prof, 16 * sizeof(unsigned long long));
    if (e != cudaSuccess) return e;
    unsigned long long z[16] = {0};
    return cudaMemcpyToSymbol(g_flow_prof, z, sizeof(z));
}
