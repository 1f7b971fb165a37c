// Warp-level building blocks on 32x32 fp64 tiles in the fragment-major layout of cm_tile_off()
// (cm_common.cuh), shared by the factorisation and substitution kernels.
#pragma once
#include "cm_common.cuh"

// FP64 tensor-core MMA, D(8x8) += A(8x4) * B(4x8).  Fragment ownership (lane = 4*g + t):
//   a = A[g][t],  b = B[t][g],  c0,c1 = C[g][2t], C[g][2t+1].
// (tcgen05 has no f64 kind; mma.sync m8n8k4 is the FP64 tensor path on sm_100a.)
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ const double *cm_tile_ptr(const double *tiles, int bt, int I, int K) {
    return tiles + ((size_t)I * bt + (K - I + bt - 1)) * CM_TILE_ELEMS;
}
__device__ __forceinline__ double *cm_tile_ptr(double *tiles, int bt, int I, int K) {
    return tiles + ((size_t)I * bt + (K - I + bt - 1)) * CM_TILE_ELEMS;
}

// 32-byte accesses as two 16-byte vector instructions; p must be 32-byte aligned.  (The single
// 256-bit forms ld/st.global.v4.f64 exist on sm_100a, but ptxas 12.9 was seen to drop three of the
// four components of such stores in one instantiation of the factorisation kernel, so they are
// not used.)
__device__ __forceinline__ void ldg4(double (&v)[4], const double *p) {
    asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p) : "memory");
    asm volatile("ld.global.v2.f64 {%0,%1}, [%2+16];" : "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
}
// streaming variant for the operand streams and the accumulator reads: the two 16-byte halves of a
// lane's 32 bytes still meet in L1 (same sector), but the line is the first to be evicted, so the
// small data that IS re-read through L1 (spilled registers, M_J, pivots, envelope) survives the
// streams.
__device__ __forceinline__ void ldg4s(double (&v)[4], const double *p) {
#if defined(CM_LD256) && CM_LD256       // experiment: one 256-bit load per lane instead of two 128-bit halves
    asm volatile("ld.global.L1::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
#elif defined(CM_L2_EL) && CM_L2_EL      // experiment: operand tiles are re-read by the following rows - last to be evicted from L2
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L1::evict_first.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v[0]), "=d"(v[1]) : "l"(p), "l"(pol) : "memory");
    asm volatile("ld.global.L1::evict_first.L2::cache_hint.v2.f64 {%0,%1}, [%2+16], %3;" : "=d"(v[2]), "=d"(v[3]) : "l"(p), "l"(pol) : "memory");
#else
    asm volatile("ld.global.L1::evict_first.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p) : "memory");
    asm volatile("ld.global.L1::evict_first.v2.f64 {%0,%1}, [%2+16];" : "=d"(v[2]), "=d"(v[3]) : "l"(p) : "memory");
#endif
}
__device__ __forceinline__ void stg4(double *p, double a, double b, double c, double d) {
    asm volatile("st.global.v2.f64 [%0], {%1,%2};" :: "l"(p), "d"(a), "d"(b) : "memory");
    asm volatile("st.global.v2.f64 [%0+16], {%1,%2};" :: "l"(p), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ double2 ldg2(const double *p) {
    double2 v;
    asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lds4(double (&v)[4], const double *p) {
    const double2 x = reinterpret_cast<const double2 *>(p)[0];
    const double2 y = reinterpret_cast<const double2 *>(p)[1];
    v[0] = x.x; v[1] = x.y; v[2] = y.x; v[3] = y.y;
}

// Accumulator fragments of one warp: a 32x32 tile as 4x4 MMA atoms; c[cj][e][ri] is the element
// at row ri*8 + lane/4, column cj*8 + (lane%4)*2 + e.  In tile storage the 8 values of one cj are
// the 8 consecutive doubles at cbase(lane) + cj*256.
struct TileAcc { double c[4][2][4]; };

__device__ __forceinline__ int acc_cbase(int lane) {
    const int g = lane >> 2, t = lane & 3;
    return ((t >> 1) << 7) + (((g << 2) + ((t & 1) << 1)) << 2);
}

__device__ __forceinline__ void acc_zero(TileAcc &a) {
#pragma unroll
    for (int cj = 0; cj < 4; ++cj)
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int ri = 0; ri < 4; ++ri) a.c[cj][e][ri] = 0.0;
}

__device__ __forceinline__ void acc_load(TileAcc &a, const double *T, int lane) {
    const double *p = T + acc_cbase(lane);
#pragma unroll
    for (int cj = 0; cj < 4; ++cj) {
        ldg4s(a.c[cj][0], p + cj * 256);
        ldg4s(a.c[cj][1], p + cj * 256 + 4);
    }
}

// the same from an assembled K~ tile, which K2 leaves ROW-MAJOR (r*32 + k): the two values of one
// (cj, ri) are 16 contiguous bytes, a warp load covers eight 64-byte row segments
__device__ __forceinline__ void acc_load_rm(TileAcc &a, const double *T, int lane) {
    const double *p = T + (lane >> 2) * 32 + 2 * (lane & 3);
#if defined(CM_L2_EF) && CM_L2_EF      // experiment: the assembled tile is read exactly once - first to be evicted from L2 as well
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#endif
#pragma unroll
    for (int cj = 0; cj < 4; ++cj)
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) {
            double2 v;
#if defined(CM_L2_EF) && CM_L2_EF
            asm volatile("ld.global.L1::evict_first.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p + ri * 256 + cj * 8), "l"(pol) : "memory");
#else
            asm volatile("ld.global.L1::evict_first.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p + ri * 256 + cj * 8) : "memory");
#endif
            a.c[cj][0][ri] = v.x; a.c[cj][1][ri] = v.y;
        }
}

__device__ __forceinline__ void acc_store(const TileAcc &a, double *T, int lane) {
    double *p = T + acc_cbase(lane);
#pragma unroll
    for (int cj = 0; cj < 4; ++cj) {
        stg4(p + cj * 256, a.c[cj][0][0], a.c[cj][0][1], a.c[cj][0][2], a.c[cj][0][3]);
        stg4(p + cj * 256 + 4, a.c[cj][1][0], a.c[cj][1][1], a.c[cj][1][2], a.c[cj][1][3]);
    }
}

// 16 (or, LOWER: the 10 on and below the block diagonal) MMAs of one k-step
template <bool LOWER>
__device__ __forceinline__ void frag_mma(TileAcc &acc, const double (&a)[4], const double (&b)[4]) {
#pragma unroll
    for (int ri = 0; ri < 4; ++ri)
#pragma unroll
        for (int cj = 0; cj < 4; ++cj)
            if (!LOWER || cj <= ri) dmma_m8n8k4(acc.c[cj][0][ri], acc.c[cj][1][ri], a[ri], b[cj]);
}

// acc -= sum over nsteps k-steps (4 columns each) of  A[:, k] * d[k] * B[:, k]^T, where the operand
// tiles of consecutive block columns are contiguous in memory (one tile row of the block band), so
// the whole sum is one linear stream: A and B advance 128 doubles per k-step, d advances 4.
// nsteps is a multiple of 8.  Software pipelined by hand: the fragments of step q+1 are in flight
// while the MMAs of step q issue.  SYRK: A == B (diagonal tile; only the lower atoms are computed).
#ifndef CM_PF_DIST
#define CM_PF_DIST 2
#endif
template <bool SYRK>
__device__ __forceinline__ void acc_update_stream(TileAcc &acc, const double *__restrict__ A,
                                                  const double *__restrict__ B, const double *__restrict__ d,
                                                  int nsteps, int lane) {
    const double *pa = A + lane * 4;
#if defined(CM_EXP) && CM_EXP >= 1          // timing experiment (wrong results): B stream = A stream, meets it in L1
    const double *pb = A + lane * 4;
#else
    const double *pb = B + lane * 4;
#endif
    const double *pd = d + (lane & 3);
    double a0[4], b0[4], a1[4], b1[4], d0, d1;
    ldg4s(a0, pa);
    if (!SYRK) ldg4s(b0, pb);
    d0 = pd[0];
    for (int q = 0; q < nsteps; q += 2) {
#if CM_PF_DIST > 0      // pull the lane's sectors of k-steps q+D, q+D+1 into L1 with real (dummy) 4-byte loads: the register look-ahead
                        // is one k-step, this makes the following pair an L1 hit (9.93 -> 9.87 ms per 4096 subsets at D = 2; the
                        // addresses stay inside the tile row's storage; with all operands resident in a larger L1 - timing only -
                        // the gain is 1.8 %: operand latency is not what bounds the kernel, profiles/r02_experiments.txt)
        {
            unsigned z0, z1;
            asm volatile("ld.global.L1::evict_last.b32 %0, [%1];" : "=r"(z0) : "l"(pa + 128 * CM_PF_DIST) : "memory");
            asm volatile("ld.global.L1::evict_last.b32 %0, [%1];" : "=r"(z1) : "l"(pa + 128 * (CM_PF_DIST + 1)) : "memory");
            if (!SYRK) {
                asm volatile("ld.global.L1::evict_last.b32 %0, [%1];" : "=r"(z0) : "l"(pb + 128 * CM_PF_DIST) : "memory");
                asm volatile("ld.global.L1::evict_last.b32 %0, [%1];" : "=r"(z1) : "l"(pb + 128 * (CM_PF_DIST + 1)) : "memory");
            }
        }
#endif
        ldg4s(a1, pa + 128);
        if (!SYRK) ldg4s(b1, pb + 128);
        d1 = pd[4];
        {
            const double nd = -d0;
#pragma unroll
            for (int i = 0; i < 4; ++i) b0[i] = (SYRK ? a0[i] : b0[i]) * nd;
            frag_mma<SYRK>(acc, a0, b0);
        }
#if defined(CM_EXP) && CM_EXP >= 2          // timing experiment: the streams do not advance (operands from L1)
        pd += 8;
#else
        pa += 256; pb += 256; pd += 8;
#endif
        if (q + 2 < nsteps) {
            ldg4s(a0, pa);
            if (!SYRK) ldg4s(b0, pb);
            d0 = pd[0];
        }
        {
            const double nd = -d1;
#pragma unroll
            for (int i = 0; i < 4; ++i) b1[i] = (SYRK ? a1[i] : b1[i]) * nd;
            frag_mma<SYRK>(acc, a1, b1);
        }
    }
}

// ---- operand ring in shared memory, filled by the bulk-copy engine (TMA, cp.async.bulk + mbarrier)
// The operand streams are linear in memory (one k-step of a tile row = 1 KB of A or B, consecutive k-steps and tiles
// contiguous), so a stage of the ring - one PAIR of k-steps: 2 KB of A, 2 KB of B, 64 bytes of pivots - is three 1-D
// bulk copies issued by one lane, completing on the stage's mbarrier.  NST stages per warp run ahead of the MMAs: the
// L2 / DRAM latency of the operands no longer sits in registers (one k-step of look-ahead), but in shared memory
// (2 NST k-steps).  Producer and consumer are the same warp, so a stage is refilled in program order after the warp
// has read it (no "empty" barriers).  Costs 4.1 KB of shared memory per stage and warp.
#define CM_RING_STAGE_BYTES 4224u       // 2048 A + 2048 B + 64 pivots, padded to a multiple of 128
struct OpRing { unsigned stage0, bar0, it; };     // shared-window addresses of the warp's stages and barriers; pairs consumed so far

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void lds4a(double (&v)[4], unsigned addr) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(addr) : "memory");
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2+16];" : "=d"(v[2]), "=d"(v[3]) : "r"(addr) : "memory");
}

// the update stream of acc_update_stream<false> with its operands taken from the ring
template <int NST>
__device__ __forceinline__ void acc_update_stream_ring(TileAcc &acc, const double *__restrict__ A, const double *__restrict__ B,
                                                       const double *__restrict__ d, int nsteps, int lane, OpRing &R) {
    const int npairs = nsteps >> 1;
    auto issue = [&](int p, unsigned it) {          // lane 0: fill the stage of pair p
        const unsigned s = it & (NST - 1), stage = R.stage0 + s * CM_RING_STAGE_BYTES, bar = R.bar0 + 8 * s;
        mbar_expect(bar, 4160);
        bulk_g2s(stage, A + (size_t)p * 256, 2048, bar);
        bulk_g2s(stage + 2048, B + (size_t)p * 256, 2048, bar);
        bulk_g2s(stage + 4096, d + (size_t)p * 8, 64, bar);
    };
    if (lane == 0) {
        asm volatile("fence.proxy.async;" ::: "memory");     // the tiles were written through the generic proxy (by other warps)
        for (int p = 0; p < NST && p < npairs; ++p) issue(p, R.it + p);
    }
#pragma unroll 1
    for (int p = 0; p < npairs; ++p) {
        const unsigned it = R.it + p, s = it & (NST - 1);
        mbar_wait(R.bar0 + 8 * s, (it / NST) & 1);
        const unsigned st = R.stage0 + s * CM_RING_STAGE_BYTES;
        double a0[4], b0[4], a1[4], b1[4], d0, d1;
        lds4a(a0, st + lane * 32);
        lds4a(b0, st + 2048 + lane * 32);
        lds4a(a1, st + 1024 + lane * 32);
        lds4a(b1, st + 3072 + lane * 32);
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(d0) : "r"(st + 4096 + (lane & 3) * 8) : "memory");
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(d1) : "r"(st + 4128 + (lane & 3) * 8) : "memory");
#pragma unroll
        for (int i = 0; i < 4; ++i) { b0[i] *= -d0; b1[i] *= -d1; }
        __syncwarp();                                        // every lane holds its fragments: the stage may be refilled
        if (lane == 0 && p + NST < npairs) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the refill (async proxy) follows generic-proxy reads of the stage
            issue(p + NST, it + NST);
        }
        frag_mma<false>(acc, a0, b0);
        frag_mma<false>(acc, a1, b1);
    }
    R.it += npairs;
}

// ---- tensor memory as a parking place for an accumulator tile (the look-ahead variant of the factorisation kernel):
// tcgen05.st / tcgen05.ld move the 64 registers of a lane to / from 64 columns of the lane's row of the CTA's allocation
// (SASS STTM.x64 / LDTM.x64); warp w of a 4-warp CTA owns lanes 32w .. 32w+31.  No f64 MMA reads tensor memory - it is
// used as 256 KB per SM of register-speed storage that the kernel otherwise leaves idle.
// (Four x16 transfers, one per block column of the tile; a single x64 transfer - 64 consecutive registers - measured the same.)
__device__ __forceinline__ void tm_park(unsigned taddr, const TileAcc &a) {
#pragma unroll
    for (int cj = 0; cj < 4; ++cj) {
        unsigned r[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double v = a.c[cj][i >> 2][i & 3];
            r[2 * i] = (unsigned)__double2loint(v); r[2 * i + 1] = (unsigned)__double2hiint(v);
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                     :: "r"(taddr + 16u * cj),
                     "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15])
                     : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_unpark(unsigned taddr, TileAcc &a) {
    unsigned r[4][16];
#pragma unroll
    for (int cj = 0; cj < 4; ++cj)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[cj][0]),"=r"(r[cj][1]),"=r"(r[cj][2]),"=r"(r[cj][3]),"=r"(r[cj][4]),"=r"(r[cj][5]),"=r"(r[cj][6]),"=r"(r[cj][7]),
                       "=r"(r[cj][8]),"=r"(r[cj][9]),"=r"(r[cj][10]),"=r"(r[cj][11]),"=r"(r[cj][12]),"=r"(r[cj][13]),"=r"(r[cj][14]),"=r"(r[cj][15])
                     : "r"(taddr + 16u * cj) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int cj = 0; cj < 4; ++cj)
#pragma unroll
        for (int i = 0; i < 8; ++i) a.c[cj][i >> 2][i & 3] = __hiloint2double((int)r[cj][2 * i + 1], (int)r[cj][2 * i]);
}

// acc <- acc * M^T in place, M = D_J^-1 L_JJ^-1 lower triangular in the cm_mtile_off layout (the
// triangular solve of an off-diagonal tile as a tile product).  The accumulator fragments ARE the A
// operand: k-slot t of k-step (cj,e) stands for column 8cj + 2t + e, which is exactly the element
// c[cj][e][ri] this lane already holds, so the tile never leaves the registers between its update
// stream and its solve.  Output block column co only needs the inputs cj <= co: going from the right,
// every finished block overwrites an input block that is dead.  Columns left of block cj0 are zero.
__device__ __forceinline__ void acc_solve_inplace(TileAcc &acc, const double *__restrict__ Mq, int cj0, int lane) {
    const double *pm = Mq + lane * 2;
#pragma unroll
    for (int co = 3; co >= 0; --co) {
        double r0[4] = {0.0, 0.0, 0.0, 0.0}, r1[4] = {0.0, 0.0, 0.0, 0.0};
        double2 b[4];
#pragma unroll
        for (int cj = 0; cj <= co; ++cj) b[cj] = ldg2(pm + co * 256 + cj * 64);
#pragma unroll
        for (int cj = 0; cj <= co; ++cj) {
            if (cj < cj0) continue;
#pragma unroll
            for (int ri = 0; ri < 4; ++ri) dmma_m8n8k4(r0[ri], r1[ri], acc.c[cj][0][ri], b[cj].x);
#pragma unroll
            for (int ri = 0; ri < 4; ++ri) dmma_m8n8k4(r0[ri], r1[ri], acc.c[cj][1][ri], b[cj].y);
        }
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) { acc.c[co][0][ri] = r0[ri]; acc.c[co][1][ri] = r1[ri]; }
    }
}

// partial forward substitution of one finished tile R = L(I,J):  out[ri] = (R y)[ri*8 + lane/4]
// (valid on every lane), y = the 32 forward-solved values of block J.
__device__ __forceinline__ void acc_matvec(const TileAcc &r, const double *__restrict__ y, double (&out)[4], int lane) {
    const double *py = y + 2 * (lane & 3);
#pragma unroll
    for (int ri = 0; ri < 4; ++ri) out[ri] = 0.0;
#pragma unroll
    for (int cj = 0; cj < 4; ++cj) {
        const double2 yy = ldg2(py + cj * 8);
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) out[ri] += r.c[cj][0][ri] * yy.x + r.c[cj][1][ri] * yy.y;
    }
#pragma unroll
    for (int ri = 0; ri < 4; ++ri) {
        out[ri] += __shfl_xor_sync(0xffffffffu, out[ri], 1);
        out[ri] += __shfl_xor_sync(0xffffffffu, out[ri], 2);
    }
}
