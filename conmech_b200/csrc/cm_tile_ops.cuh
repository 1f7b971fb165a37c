// Warp-level building blocks on 32x32 fp64 tiles shared by the factorisation kernels.
#pragma once
#include "cm_common.cuh"

// FP64 tensor-core MMA, D(8x8) += A(8x4) * B(4x8).  Fragment ownership (lane = 4*g + t):
//   a = A[g][t],  b = B[t][g],  c0,c1 = C[g][2t], C[g][2t+1].
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

// Accumulator fragments of one warp: a 32x32 tile as 4x4 MMA tiles, c[ri][cj][0..1] holding rows
// ri*8 + lane/4, columns cj*8 + (lane%4)*2 + {0,1}.
struct TileAcc { double c[4][4][2]; };

__device__ __forceinline__ void acc_load(TileAcc &t, const double *A, int lane) {
    const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
    for (int ri = 0; ri < 4; ++ri)
#pragma unroll
        for (int cj = 0; cj < 4; ++cj) {
            const double2 v = *reinterpret_cast<const double2 *>(A + cm_tile_off(ri * 8 + fr, cj * 8 + fk * 2));
            t.c[ri][cj][0] = v.x;
            t.c[ri][cj][1] = v.y;
        }
}

// acc -= Lik * diag(dK) * Ljk^T   (both operand tiles row-major [row][k] in global / L2).
// Software pipelined by hand: the fragments of k-step s+1 are requested before the 16 MMAs of
// step s issue; the compiler barrier keeps at most two steps of fragments live (register budget).
__device__ __forceinline__ void frag_load(double (&a)[4], double (&b)[4], const double *__restrict__ Lik,
                                          const double *__restrict__ Ljk, const double *__restrict__ dK,
                                          int ks, int fr, int fk) {
    const double dk = -dK[ks * 4 + fk];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a[i] = Lik[cm_tile_off(i * 8 + fr, ks * 4 + fk)];
        b[i] = Ljk[cm_tile_off(i * 8 + fr, ks * 4 + fk)] * dk;
    }
}

__device__ __forceinline__ void frag_mma(TileAcc &t, const double (&a)[4], const double (&b)[4]) {
#pragma unroll
    for (int ri = 0; ri < 4; ++ri)
#pragma unroll
        for (int cj = 0; cj < 4; ++cj) dmma_m8n8k4(t.c[ri][cj][0], t.c[ri][cj][1], a[ri], b[cj]);
}

__device__ __forceinline__ void acc_update(TileAcc &t, const double *__restrict__ Lik,
                                           const double *__restrict__ Ljk, const double *__restrict__ dK,
                                           int lane) {
    const int fr = lane >> 2, fk = lane & 3;
    double a0[4], b0[4], a1[4], b1[4];
    frag_load(a0, b0, Lik, Ljk, dK, 0, fr, fk);
#pragma unroll
    for (int ks = 0; ks < 8; ks += 2) {
        asm volatile("" ::: "memory");
        frag_load(a1, b1, Lik, Ljk, dK, ks + 1, fr, fk);
        frag_mma(t, a0, b0);
        asm volatile("" ::: "memory");
        if (ks + 2 < 8) frag_load(a0, b0, Lik, Ljk, dK, ks + 2, fr, fk);
        frag_mma(t, a1, b1);
    }
}

// Fragment layout -> one tile row per lane through a warp-private 32 x 33 shared scratch (all
// fragments are written before any row is read, so accumulators and row never coexist in registers).
__device__ __forceinline__ void acc_to_rows(const TileAcc &t, double (&row)[32], double *scr, int lane) {
    const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
    for (int ri = 0; ri < 4; ++ri)
#pragma unroll
        for (int cj = 0; cj < 4; ++cj) {
            scr[(ri * 8 + fr) * 33 + cj * 8 + fk * 2] = t.c[ri][cj][0];
            scr[(ri * 8 + fr) * 33 + cj * 8 + fk * 2 + 1] = t.c[ri][cj][1];
        }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 32; ++k) row[k] = scr[lane * 33 + k];
    __syncwarp();
}

__device__ __forceinline__ void acc_store(const TileAcc &t, double *A, int lane) {
    const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
    for (int ri = 0; ri < 4; ++ri)
#pragma unroll
        for (int cj = 0; cj < 4; ++cj)
            *reinterpret_cast<double2 *>(A + cm_tile_off(ri * 8 + fr, cj * 8 + fk * 2)) =
                make_double2(t.c[ri][cj][0], t.c[ri][cj][1]);
}

// LDL^T of a 32x32 tile held one row per lane (lower triangle meaningful).  On return lane r holds
// L[r][0..r-1] in row[], the pivot d_r is returned, `fail` is set when a pivot violates the
// reference's rules (exact zero raises in SuperLU: numpy_stiffness.py:301-319; <= -1e-14 is
// rejected: :326-336; NaN fails both) and `minpiv` tracks the smallest pivot.  col: 32 doubles of
// warp-private shared scratch.
__device__ __forceinline__ double tile_ldlt(double (&row)[32], double *col, int lane, int &fail, double &minpiv) {
    double dval = 1.0;
#pragma unroll 32
    for (int c = 0; c < 32; ++c) {
        col[lane] = row[c];
        __syncwarp();
        double d = col[c];
        if (!(d > -CM_EPS) || d == 0.0) { fail = 1; d = 1.0; }
        minpiv = fmin(minpiv, d);
        if (lane == c) dval = d;
        if (lane > c) {
            const double l = row[c] / d;
            row[c] = l;
#pragma unroll
            for (int k = 0; k < 32; ++k)
                if (k > c) row[k] -= l * col[k];
        }
        __syncwarp();
    }
    return dval;
}

// Y Ljj^T = row (unit lower Ljj, swizzled tile layout in shared memory), then L = Y D^-1
__device__ __forceinline__ void tile_trsm(double (&row)[32], const double *Ljj, const double *dj) {
#pragma unroll 32
    for (int c = 0; c < 32; ++c) {
        const double y = row[c];
#pragma unroll
        for (int k = 0; k < 32; ++k)
            if (k > c) row[k] -= y * Ljj[cm_tile_off(k, c)];
        row[c] = y / dj[c];
    }
}

// strictly-lower triangle of a unit lower 32x32 factor, packed by rows: (k,c), c<k -> k(k-1)/2 + c
__host__ __device__ constexpr int cm_tri_off(int k, int c) { return k * (k - 1) / 2 + c; }
#define CM_TRI_ELEMS 496

// same as tile_trsm with the diagonal factor in packed storage
__device__ __forceinline__ void tile_trsm_packed(double (&row)[32], const double *Lp, const double *dj) {
#pragma unroll 32
    for (int c = 0; c < 32; ++c) {
        const double y = row[c];
#pragma unroll
        for (int k = 0; k < 32; ++k)
            if (k > c) row[k] -= y * Lp[cm_tri_off(k, c)];
        row[c] = y / dj[c];
    }
}

__device__ __forceinline__ void row_store(double *T, const double (&row)[32], int lane) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        double2 *dst = reinterpret_cast<double2 *>(T + cm_tile_off(lane, g * 4));
        dst[0] = make_double2(row[g * 4], row[g * 4 + 1]);
        dst[1] = make_double2(row[g * 4 + 2], row[g * 4 + 3]);
    }
}
