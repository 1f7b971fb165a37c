// Pieces shared by the factorisation kernels (k3_factor_flow.cu, k3_factor_pair.cu): the blocked
// diagonal-tile factorisation, the forward substitution of one block row, progress-flag waits.
#pragma once
#include "cm_frag.cuh"

#ifndef CM_SLEEP_NS
#define CM_SLEEP_NS 40          // back-off of the progress-flag polls (5 and 150 ns measured the same)
#endif

namespace {

constexpr int DS = 34;                          // row stride of the diagonal scratch matrices (doubles)
constexpr int DSCR = 2 * 32 * DS + 2 * 32;      // CTA scratch of the diagonal factorisation: two 32x32
                                                // matrices, a column, the reciprocal pivots.  One copy
                                                // per CTA: the diagonal factorisations of a subset form
                                                // a dependency chain, so only one is ever in progress.

// Per-CTA context of the factorisation kernel in SHARED memory.  The device functions below are real
// calls (not inlined: the register allocator does best with the stream loops in functions of their
// own); with two dozen pointer/size arguments most of them travelled through the stack - local
// memory, which misses the L1 under the operand streams - and were reloaded from L2 at every use.
// Instead every function takes two or three integers in registers and reads what it needs from this
// block at the point of use (volatile: never cached in registers across a stream, never spilled).
struct K3Ctx {
    double *tiles; double *minv; double *dpiv; double *X;
    const int *kcol; const unsigned long long *arow; int *prog; int *yprog;
    double *s_part; double *s_rhs; double *s_scr; double *s_minpiv; int *s_fail; double2 *s_dacc;
    long long *pc; long long *tpub;
    // prefix-sharing group: tile rows < trows of the factor (tiles, M_J, pivots, forward solutions) are read from the
    // trunk subset's slab, which an earlier launch has completed; trows == 0: none
    const double *ttiles; const double *tminv; const double *tdpiv; const double *tX;
    int trows;
    unsigned ring0, ringbar0;      // operand ring (OPS builds): shared-window address of warp 0's stages / barriers
    unsigned *ring_it;             // [NW] pairs each warp has consumed so far (stage index and barrier phase follow from it)
    unsigned tmem;                 // look-ahead builds: base address of the CTA's tensor-memory allocation (128 columns)
    int bt, n_pad, n_lc, zero_bits;
};
__shared__ K3Ctx s_ctx;
#define CTX(f) (((volatile K3Ctx *)&s_ctx)->f)

__device__ __forceinline__ double fast_rcp(double d) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    double e = fma(-d, x, 1.0);      // seed: relative error <= 2^-23
    x = fma(x, e, x);                // 2^-46
    e = fma(-d, x, 1.0);
    return fma(x, e, x);             // below one ulp
}

// Progress flags live in shared memory, the data they guard in global memory (written by other warps of the
// SAME CTA, i.e. through the same L1).  Consumer: an acquire load of the flag - on sm_100a a plain LDS; the loads
// that follow cannot issue before it has returned (in-order issue, control dependence) - and NO fence: a
// __threadfence_block() there is a MEMBAR.SC.CTA that also drains the polling warp's OWN outstanding stores (the 8 KB
// tile it has just written).  Producer: release fence (fence.acq_rel.cta = MEMBAR.ALL.CTA) after its stores, then the
// flag store.  (9.95 -> 9.91 ms per 4096 subsets against __threadfence_block() on both sides.)
__device__ __forceinline__ int flag_acquire(const volatile int *p) {
    int v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(const_cast<const int *>(p))) : "memory");
    return v;
}
// (OPS builds read the tiles through the bulk-copy engine, i.e. from L2: their producers fence at GPU scope)
template <int OPS = 0>
__device__ __forceinline__ void release_fence() {
    if (OPS) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    else asm volatile("fence.acq_rel.cta;" ::: "memory");
}
__device__ __forceinline__ int wait_prog_gt(volatile int *prog, int J, int K) {
    int p = flag_acquire(prog + J);
    while (p <= K) { __nanosleep(CM_SLEEP_NS); p = flag_acquire(prog + J); }
    return p;
}

__device__ __forceinline__ double dneg(double x) {       // sign flip on the integer pipe
    return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x));
}

// Blocked LDL^T of the lower triangle of tile T (global, tile layout) by one warp, then
// M = D^-1 L^-1, written to Mg (global, tile layout); pivots to dg.  8-column panels are eliminated
// with one matrix row per lane (the pivot column travels through shared memory); the trailing
// update and the block-wise inverse (X = L^-1: X_ii by substitution, X_ij = -X_ii sum_k L_ik X_kj)
// run as 8x8x4 FP64 MMAs on fragments read from the shared scratch matrices.  The serial chain is
// 32 pivots of ~100 cycles plus a dozen short MMA stages, and the warp issues ~4x fewer FP64
// instructions than an unblocked elimination - this routine is the critical path of every tile row.
// A pivot that violates the reference's rules (exact zero raises in SuperLU: numpy_stiffness.py:
// 301-319; <= -1e-14 is rejected: :326-336; NaN fails both) sets *s_fail and is replaced by 1.
// in_smem: the lower blocks of the tile were already placed in scr (row stride DS) by
// diag_term(); otherwise the tile is read from T.
__device__ __noinline__ void diag_factor_blk(int I, bool in_smem, int lane) {
    double *scr = CTX(s_scr);
    double *sA = scr, *sB = scr + 32 * DS, *s_col = sB + 32 * DS, *s_rinv = s_col + 32;
    double *s_minpiv_w = CTX(s_minpiv) + (threadIdx.x >> 5);
    const int g = lane >> 2, t = lane & 3;
    if (!in_smem) {                      // a row with nothing left of the diagonal: the assembled K~ tile, ROW-MAJOR (K2)
        const double *T = cm_tile_ptr(CTX(tiles), CTX(bt), I, I);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            double v[4];
            ldg4(v, T + q * 128 + lane * 4);                 // elements 128q + 4*lane .. +3 = row 4q + lane/8, columns 4*(lane%8) ..
#pragma unroll
            for (int i = 0; i < 4; ++i) sA[(4 * q + (lane >> 3)) * DS + 4 * (lane & 7) + i] = v[i];
        }
    }
    __syncwarp();
    int fail = 0;
    double minpiv = *s_minpiv_w;
    double dmine = 1.0, rmine = 1.0;
#pragma unroll 1                     // code size: the kernel's working set must stay inside the instruction cache
    for (int b = 0; b < 4; ++b) {
        const int c0 = 8 * b;
        double p[8], pk[8];
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(sA + lane * DS + c0 + k);
            p[k] = v.x; p[k + 1] = v.y;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            s_col[lane] = p[c];
            __syncwarp();
            double d = s_col[c0 + c];
            // The pivot -> next pivot recurrence is the critical path of the whole kernel, and every
            // dependent FP64 instruction on it queues behind the other warps' DMMAs.  Keep it to the
            // reciprocal (hardware seed, relative error <= 2^-23, and ONE cubic Newton step:
            // x (1 + e + e^2), error e^3) and one FMA per updated entry: the products with the pivot
            // row do not need the reciprocal and are formed while it is computed.
            double u[8];
#pragma unroll
            for (int k = c + 1; k < 8; ++k) u[k] = p[c] * s_col[c0 + k];
            double x;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
            const double e = fma(-d, x, 1.0);
            const double t2 = fma(e, e, e);
            double rinv = fma(x, t2, x);
            if (!(d > -CM_EPS) || d == 0.0) { fail = 1; d = 1.0; rinv = 1.0; }
            minpiv = fmin(minpiv, d);
            if (lane == c0 + c) { dmine = d; rmine = rinv; }
            pk[c] = p[c];                        // (L D)[lane][c0+c]
#pragma unroll
            for (int k = c + 1; k < 8; ++k) p[k] = fma(-u[k], rinv, p[k]);
            p[c] = p[c] * rinv;                  // L[lane][c0+c] for lane > c0+c (unused otherwise)
            __syncwarp();
        }
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
            *reinterpret_cast<double2 *>(sA + lane * DS + c0 + k) = make_double2(p[k], p[k + 1]);
            *reinterpret_cast<double2 *>(sB + lane * DS + c0 + k) = make_double2(pk[k], pk[k + 1]);
        }
        __syncwarp();
        if (b < 3) {
            // trailing blocks (ri,cj), b < cj <= ri:  A -= (L D)(ri,b) L(cj,b)^T
            for (int ri = b + 1; ri < 4; ++ri)
                for (int cj = b + 1; cj <= ri; ++cj) {
                    double2 cc = *reinterpret_cast<double2 *>(sA + (8 * ri + g) * DS + 8 * cj + 2 * t);
#pragma unroll
                    for (int s2 = 0; s2 < 2; ++s2)
                        dmma_m8n8k4(cc.x, cc.y, dneg(sB[(8 * ri + g) * DS + c0 + 4 * s2 + t]),
                                    sA[(8 * cj + g) * DS + c0 + 4 * s2 + t]);
                    *reinterpret_cast<double2 *>(sA + (8 * ri + g) * DS + 8 * cj + 2 * t) = cc;
                }
            __syncwarp();
        }
    }
    CTX(dpiv)[I * CM_TILE + lane] = dmine;
    s_rinv[lane] = rmine;
    if (lane == 0) {
        *s_minpiv_w = minpiv;
        if (fail && *CTX(s_fail) == 0) *CTX(s_fail) = I + 1;      // (diagonal tiles are factored one at a time, in row order)
    }
    // ---- X = L^-1.  Diagonal blocks: lane 8*bb + j owns column j of X_bb (substitution on e_j)
    {
        const int bb = lane >> 3, j = lane & 7;
        const double *Lb = sA + (8 * bb) * DS + 8 * bb;
        double x[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) x[r] = (r == j) ? 1.0 : 0.0;
#pragma unroll
        for (int c = 0; c < 7; ++c)
#pragma unroll
            for (int r = c + 1; r < 8; ++r) x[r] -= Lb[r * DS + c] * x[c];
#pragma unroll
        for (int r = 0; r < 8; ++r) sB[(8 * bb + r) * DS + 8 * bb + j] = x[r];
    }
    __syncwarp();
    // off-diagonal blocks by distance d from the diagonal: T_ij = sum_{k=j}^{i-1} L_ik X_kj is parked in
    // the unused upper block (j,i) of sA, then X_ij = -X_ii T_ij
#pragma unroll
    for (int d = 1; d < 4; ++d) {
#pragma unroll
        for (int i = d; i < 4; ++i) {
            const int j = i - d;
            double c0v = 0.0, c1v = 0.0;
#pragma unroll
            for (int k = j; k < i; ++k)
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2)
                    dmma_m8n8k4(c0v, c1v, sA[(8 * i + g) * DS + 8 * k + 4 * s2 + t],
                                sB[(8 * k + 4 * s2 + t) * DS + 8 * j + g]);
            *reinterpret_cast<double2 *>(sA + (8 * j + g) * DS + 8 * i + 2 * t) = make_double2(c0v, c1v);
        }
        __syncwarp();
#pragma unroll
        for (int i = d; i < 4; ++i) {
            const int j = i - d;
            double c0v = 0.0, c1v = 0.0;
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2)
                dmma_m8n8k4(c0v, c1v, dneg(sB[(8 * i + g) * DS + 8 * i + 4 * s2 + t]),
                            sA[(8 * j + 4 * s2 + t) * DS + 8 * i + g]);
            *reinterpret_cast<double2 *>(sB + (8 * i + g) * DS + 8 * j + 2 * t) = make_double2(c0v, c1v);
        }
        __syncwarp();
    }
    // ---- M = D^-1 X to global memory, cm_mtile_off layout (blocks above the diagonal are not stored)
    double *Mg = CTX(minv) + (size_t)I * CM_TILE_ELEMS;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double ri = s_rinv[8 * i + g];
#pragma unroll
        for (int cj = 0; cj <= i; ++cj) {
            const double2 x = *reinterpret_cast<const double2 *>(sB + (8 * i + g) * DS + 8 * cj + 2 * t);
            asm volatile("st.global.v2.f64 [%0], {%1,%2};" :: "l"(Mg + i * 256 + cj * 64 + lane * 2), "d"(x.x * ri), "d"(x.y * ri) : "memory");
        }
    }
    __syncwarp();
}

// Diagonal-tile term of a finished off-diagonal tile R = L(I,J), both MMA operands straight from the
// registers of R (k-slot t of k-step (cs,e) is column 8cs + 2t + e: the A fragment of row block ri is
// r.c[cs][e][ri], the B fragment of output column block cj is d_k * r.c[cs][e][cj] of the same lane):
//     dacc -= R diag(d_J) R^T      (lower atoms only)
// dacc is the warp's pending sum for the diagonal tile of its row in SHARED memory ([atom][lane]
// double2, atom = ri(ri+1)/2 + cj; 16-byte accesses, conflict free) - the row of L is never read
// back from global memory for the SYRK part of the diagonal tile.  The sum starts from the assembled
// diagonal tile (loaded at the start of the row, off the critical path).  For the chain tile
// (J = I-1) the result is written to the diagonal scratch sA (row stride DS) where
// diag_factor_blk() continues: the chain tile -> diagonal tile hand-over never touches global memory.  Column blocks left of cs0
// are zero in R (envelope).
__device__ __forceinline__ void diag_term(const TileAcc &r, bool chain, const double2 (&dk)[4], double *sA,
                                          double2 *dacc, int cs0, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int cj = 0; cj < 4; ++cj) {
        double c0[4], c1[4];
#pragma unroll
        for (int ri = cj; ri < 4; ++ri) {
            const double2 v = dacc[(ri * (ri + 1) / 2 + cj) * 32 + lane];
            c0[ri] = v.x; c1[ri] = v.y;
        }
#pragma unroll
        for (int cs = 0; cs < 4; ++cs) {
            if (cs < cs0) continue;
            const double nb0 = -(dk[cs].x * r.c[cs][0][cj]);
            const double nb1 = -(dk[cs].y * r.c[cs][1][cj]);
#pragma unroll
            for (int ri = cj; ri < 4; ++ri) dmma_m8n8k4(c0[ri], c1[ri], r.c[cs][0][ri], nb0);
#pragma unroll
            for (int ri = cj; ri < 4; ++ri) dmma_m8n8k4(c0[ri], c1[ri], r.c[cs][1][ri], nb1);
        }
#pragma unroll
        for (int ri = cj; ri < 4; ++ri) {
            if (chain) *reinterpret_cast<double2 *>(sA + (8 * ri + g) * DS + 8 * cj + 2 * t) = make_double2(c0[ri], c1[ri]);
            else dacc[(ri * (ri + 1) / 2 + cj) * 32 + lane] = make_double2(c0[ri], c1[ri]);
        }
    }
}

// forward substitution of block row J for every load case:  y_J = D_J M_J (b_J - s_J), where
// s_J = sum_K L(J,K) y_K was accumulated tile by tile (acc_matvec) into s_part while the tiles of
// the row were still in registers - the row of L is not read again.
__device__ __noinline__ void forward_row_w(int J, int lane) {
    const int g = lane >> 2, tt = lane & 3;
    const int n_lc = CTX(n_lc), zero_bits = CTX(zero_bits), warp = threadIdx.x >> 5;
    const double *pm = CTX(minv) + (size_t)J * CM_TILE_ELEMS + lane * 2;
    const double *dpiv = CTX(dpiv);
    const double *s_part = CTX(s_part) + warp * n_lc * 32;
    double *s_r = CTX(s_rhs) + warp * 32;
    for (int lc = 0; lc < n_lc; ++lc) {
        if ((zero_bits >> lc) & 1) continue;
        double *Xl = CTX(X) + (size_t)lc * CTX(n_pad);
        s_r[lane] = Xl[J * CM_TILE + lane] - s_part[lc * 32 + lane];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double zp = 0.0;
#pragma unroll
            for (int cj = 0; cj <= i; ++cj) {
                const double2 b = ldg2(pm + i * 256 + cj * 64);
                const double2 rr = *reinterpret_cast<const double2 *>(s_r + cj * 8 + 2 * tt);
                zp += b.x * rr.x + b.y * rr.y;
            }
            zp += __shfl_xor_sync(0xffffffffu, zp, 1);
            zp += __shfl_xor_sync(0xffffffffu, zp, 2);
            if (tt == 0) Xl[J * CM_TILE + i * 8 + g] = dpiv[J * CM_TILE + i * 8 + g] * zp;
        }
        __syncwarp();
    }
}

}  // namespace
