// K1 - fused per-element kernel: 12x12 Euler-Bernoulli local stiffness with end rotational
// springs, global->local rotation, K_G = R12^T K_loc R12, and the fixed-end lumping of uniform
// line loads (self weight + UDL) per load case.
//
// Follows, operation for operation, reference numpy_stiffness_assembly.py:11-71 (axial/torsion),
// :101-206 (bending with springs), :208-281 (12x12 scatter, quirk Q1: Iz in both planes),
// :283-327 (rotation, quirk Q3: mirrored frame for vertical members), :393-420 (load lumping)
// and numpy_stiffness.py:579-614 / :98-123.  Compiled with -fmad=false so that the scalar
// expressions round like the reference's Python floats.
#include "cm_common.cuh"

namespace {

struct BendCoef { double a, aa, a1_2, a1_3, a2_2, a2_3; };

__device__ __forceinline__ double axial_d(double L, double A, double E, double c1, double c2) {
    // numpy_stiffness_assembly.py:63-69
    if (fabs(c1) < CM_EPS || fabs(c2) < CM_EPS) return 0.0;
    double i1 = (c1 < CM_BIG) ? 1.0 / c1 : 0.0;
    double i2 = (c2 < CM_BIG) ? 1.0 / c2 : 0.0;
    return 1.0 / (L / (E * A) + i1 + i2);
}

__device__ __forceinline__ BendCoef bend_coef(double a1, double a2) {
    // numpy_stiffness_assembly.py:155-189
    BendCoef c;
    if (fabs(a1) < CM_BIG && fabs(a2) < CM_BIG) {
        double den = a1 * a2 + 4 * a1 + 4 * a2 + 12;
        c.a = a1 * a2 / den;
        c.aa = (a1 * a2 + a1 + a2) / den;
        c.a2_2 = (a2 + 2) * a1 / den;
        c.a2_3 = (a2 + 3) * a1 / den;
        c.a1_2 = (a1 + 2) * a2 / den;
        c.a1_3 = (a1 + 3) * a2 / den;
    } else if (fabs(a1) < CM_EPS && fabs(a2) > CM_BIG) {
        c.a = 0.0; c.aa = 0.25; c.a2_2 = 0.0; c.a2_3 = 0.0; c.a1_2 = 0.5; c.a1_3 = 0.75;
    } else if (fabs(a2) < CM_EPS && fabs(a1) > CM_BIG) {
        c.a = 0.0; c.aa = 0.25; c.a1_2 = 0.0; c.a1_3 = 0.0; c.a2_2 = 0.5; c.a2_3 = 0.75;
    } else {
        double i1 = (fabs(a1) > CM_BIG) ? 0.0 : 1.0 / a1;
        double i2 = (fabs(a2) > CM_BIG) ? 0.0 : 1.0 / a2;
        double den = 1 + 4 * i1 + 4 * i2 + 12 * i1 * i2;
        c.a = 1 / den;
        c.aa = (1 + i1 + i2) / den;
        c.a2_2 = (1 + 2 * i1) / den;
        c.a2_3 = (1 + 3 * i1) / den;
        c.a1_2 = (1 + 2 * i2) / den;
        c.a1_3 = (1 + 3 * i2) / den;
    }
    return c;
}

// 4x4 bending block on (v1, th1, v2, th2); s = +1 about local z, -1 about local y
__device__ __forceinline__ void bend_block(double L, double E, double Iz, double s, double c1,
                                           double c2, double K[4][4]) {
    double a1 = c1 * L / E * Iz;     // quirk Q2: evaluated left to right
    double a2 = c2 * L / E * Iz;
    BendCoef c = bend_coef(a1, a2);
    double L2 = L * L;
    K[0][0] = 12. / L2 * c.aa;   K[0][1] = s * 6 / L * c.a2_2;    K[0][2] = -12. / L2 * c.aa;  K[0][3] = s * 6 / L * c.a1_2;
    K[1][0] = s * 6 / L * c.a2_2; K[1][1] = 4 * c.a2_3;           K[1][2] = s * (-6 / L) * c.a2_2; K[1][3] = 2 * c.a;
    K[3][0] = s * 6 / L * c.a1_2; K[3][1] = 2 * c.a;              K[3][2] = s * (-6 / L) * c.a1_2; K[3][3] = 4 * c.a1_3;
    double f = E * Iz / L;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        K[2][j] = -K[0][j];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) K[i][j] *= f;
}

// which uncoupled mechanism a local DOF belongs to, and its index inside that block
// (numpy_stiffness_assembly.py:277-280): axial [0,6], torsion [3,9], bend-z [1,5,7,11],
// bend-y [2,4,8,10]
__host__ __device__ constexpr int dof_kind(int l) {
    return (l == 0 || l == 6) ? 0 : (l == 3 || l == 9) ? 1 : (l == 1 || l == 5 || l == 7 || l == 11) ? 2 : 3;
}
__host__ __device__ constexpr int dof_idx(int l) {
    return l == 0 ? 0 : l == 6 ? 1 : l == 3 ? 0 : l == 9 ? 1 : l == 1 ? 0 : l == 5 ? 1 : l == 7 ? 2
         : l == 11 ? 3 : l == 2 ? 0 : l == 4 ? 1 : l == 8 ? 2 : 3;
}

struct LocalK {
    double ax, to;          // d of the 2x2 blocks
    double bz[4][4], by[4][4];
};

__device__ __forceinline__ double local_entry(const LocalK &k, int l, int m) {
    const int kl = dof_kind(l), km = dof_kind(m);
    if (kl != km) return 0.0;
    const int il = dof_idx(l), im = dof_idx(m);
    if (kl == 0) return (il == im) ? k.ax : -k.ax;
    if (kl == 1) return (il == im) ? k.to : -k.to;
    if (kl == 2) return k.bz[il][im];
    return k.by[il][im];
}

__device__ __forceinline__ bool rotation3(const double *pu, const double *pv, double R[3][3], double &Lout) {
    // numpy_stiffness_assembly.py:283-327
    double dx = pu[0] - pv[0], dy = pu[1] - pv[1], dz = pu[2] - pv[2];
    double L = sqrt(dx * dx + dy * dy + dz * dz);
    Lout = L;
    if (L < CM_EPS) return false;
    double cx = (pv[0] - pu[0]) / L, cy = (pv[1] - pu[1]) / L, cz = (pv[2] - pu[2]) / L;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) R[i][j] = 0.0;
    if (fabs(fabs(cz) - 1.0) < CM_EPS) {        // quirk Q3
        R[0][2] = -cz; R[1][1] = 1.0; R[2][0] = cz;
    } else {
        double yx = -(cy * 1.0 - cz * 0.0);
        double yy = -(cz * 0.0 - cx * 1.0);
        double yz = -(cx * 0.0 - cy * 0.0);
        double ny = sqrt(yx * yx + yy * yy + yz * yz);
        yx /= ny; yy /= ny; yz /= ny;
        R[0][0] = cx; R[0][1] = cy; R[0][2] = cz;
        R[1][0] = yx; R[1][1] = yy; R[1][2] = yz;
        R[2][0] = cy * yz - cz * yy;
        R[2][1] = cz * yx - cx * yz;
        R[2][2] = cx * yy - cy * yx;
    }
    return true;
}

// K_G of one element per thread.  The 12x12 result leaves the thread three rows (36 entries) at a
// time through a shared staging slab [36][K1_THREADS+1]: the block then writes the slice of all its
// elements with coalesced 288-byte runs instead of one 8-byte store per thread and entry (which
// touched a separate 32-byte sector each: 4x write amplification, 21 % of the HBM copy rate on a
// 1M-element batch before the change).
constexpr int K1_THREADS = 128;
constexpr int K1_SSTR = K1_THREADS + 1;

__device__ void element_tables(const double *pu, const double *pv, const double *pr, const double *cr,
                               bool valid, double *Kg_block, double *Kt_block, int n_in_block, double *R3out, double *len_out,
                               double *s_stage) {
    const int tid = threadIdx.x;
    double R[3][3], L = 0.0;
    bool ok = false;
    LocalK k;
    if (valid) {
        ok = rotation3(pu, pv, R, L);
        if (len_out) *len_out = L;
        if (ok) {
            const double A = pr[0], Jx = pr[1], Iz = pr[3], E = pr[4], mu = pr[5];
            const double G = E / (2 * (1 + mu));                      // io_base.py:303
            const double inf = __longlong_as_double(0x7ff0000000000000LL);
            k.ax = axial_d(L, A, E, inf, inf);
            k.to = axial_d(L, Jx, G, cr[0], cr[3]);
            bend_block(L, E, Iz, 1.0, cr[2], cr[5], k.bz);            // about z: DOFs 1,5,7,11
            bend_block(L, E, Iz, -1.0, cr[1], cr[4], k.by);           // about y: DOFs 2,4,8,10 (Q1: Iz)
#pragma unroll
            for (int i = 0; i < 9; ++i) R3out[i] = R[i / 3][i % 3];
        } else {                    // host validated lengths; keep outputs defined
            for (int i = 0; i < 9; ++i) R3out[i] = 0.0;
        }
    }
    // K_G block (p,q) = R^T * Kloc(p,q) * R, evaluated as (R^T Kloc) R like numpy_stiffness.py:612
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        if (valid) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                double T[3][3];
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        double sacc = 0.0;
                        if (ok) {
#pragma unroll
                            for (int a = 0; a < 3; ++a) sacc += R[a][i] * local_entry(k, 3 * p + a, 3 * q + j);
                        }
                        T[i][j] = sacc;
                    }
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        double sacc = 0.0;
                        if (ok) {
#pragma unroll
                            for (int b = 0; b < 3; ++b) sacc += T[i][b] * R[b][j];
                        }
                        s_stage[(i * 12 + 3 * q + j) * K1_SSTR + tid] = sacc;
                    }
            }
        }
        __syncthreads();
        for (int idx = tid; idx < n_in_block * 36; idx += K1_THREADS) {
            const int el = idx / 36, jj = idx - 36 * el;
            const double kv = s_stage[jj * K1_SSTR + el];
            Kg_block[(size_t)el * 144 + p * 36 + jj] = kv;
            if (Kt_block) Kt_block[(size_t)el * 144 + p * 36 + jj] = (fabs(kv) > CM_EPS) ? kv : 0.0;
        }
        __syncthreads();
    }
}

// numpy_stiffness_assembly.py:393-420
__device__ __forceinline__ void lump_udl(const double w[3], const double *R3, double L, double q[12]) {
    double wl[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        q[i] = w[i] * L / 2;
        q[6 + i] = w[i] * L / 2;
        wl[i] = R3[3 * i + 0] * w[0] + R3[3 * i + 1] * w[1] + R3[3 * i + 2] * w[2];
    }
    double L2 = L * L;
    double M0[3] = {0.0, -wl[2] * L2 / 12., wl[1] * L2 / 12.};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double s0 = R3[0 * 3 + i] * M0[0] + R3[1 * 3 + i] * M0[1] + R3[2 * 3 + i] * M0[2];
        double s1 = R3[0 * 3 + i] * (-M0[0]) + R3[1 * 3 + i] * (-M0[1]) + R3[2 * 3 + i] * (-M0[2]);
        q[3 + i] = s0;
        q[9 + i] = s1;
    }
}

// blockIdx.y = geometry variant: node coordinates of variant g, tables of variant g (stacked)
__global__ void __launch_bounds__(K1_THREADS) k1_element_tables(cm_model_dev m) {
    __shared__ double s_stage[36 * K1_SSTR];
    const int e0 = blockIdx.x * K1_THREADS;
    const int e = e0 + threadIdx.x;
    const size_t g = blockIdx.y;
    const bool valid = e < m.nE;
    const int ec = valid ? e : 0;
    const int u = m.conn[2 * ec], v = m.conn[2 * ec + 1];
    const double *xyz = m.xyz + g * 3 * m.nV;
    element_tables(xyz + 3 * u, xyz + 3 * v, m.props + 7 * ec, m.cr + 6 * ec, valid, m.Kg + 144 * (g * m.nE + e0),
                   m.Kt + 144 * (g * m.nE + e0), min(K1_THREADS, m.nE - e0), m.R3 + 9 * (g * m.nE + ec), m.elen + g * m.nE + ec, s_stage);
}

__global__ void __launch_bounds__(K1_THREADS) k1_element_generic(const double *xyz_u, const double *xyz_v,
                                                                 const double *props, const double *cr, long long n,
                                                                 double *Kg, double *R3) {
    __shared__ double s_stage[36 * K1_SSTR];
    const long long e0 = (long long)blockIdx.x * K1_THREADS;
    const long long e = e0 + threadIdx.x;
    const bool valid = e < n;
    const long long ec = valid ? e : 0;
    element_tables(xyz_u + 3 * ec, xyz_v + 3 * ec, props + 7 * ec, cr + 6 * ec, valid, Kg + 144 * e0, nullptr,
                   (int)((n - e0) < K1_THREADS ? (n - e0) : K1_THREADS), R3 + 9 * ec, nullptr, s_stage);
}

__global__ void __launch_bounds__(128) k1_element_loads(cm_model_dev m, int n_lc, const double *w_udl,
                                                        const unsigned char *has_udl,
                                                        const unsigned char *grav_on,
                                                        const double *grav_dir, double *q) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    int lc = blockIdx.y;
    const size_t g = blockIdx.z;                                // geometry variant
    if (e >= m.nE) return;
    const double *R3 = m.R3 + 9 * (g * m.nE + e);
    double L = m.elen[g * m.nE + e];
    q += g * n_lc * m.nE * 12;
    double out[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) out[i] = 0.0;
    if (has_udl && has_udl[(size_t)lc * m.nE + e]) {           // numpy_stiffness.py:113-123
        double w[3] = {w_udl[((size_t)lc * m.nE + e) * 3 + 0], w_udl[((size_t)lc * m.nE + e) * 3 + 1],
                       w_udl[((size_t)lc * m.nE + e) * 3 + 2]};
        double t[12];
        lump_udl(w, R3, L, t);
#pragma unroll
        for (int i = 0; i < 12; ++i) out[i] += t[i];
    }
    if (grav_on[lc]) {                                          // numpy_stiffness.py:98-111
        double qsw = m.props[7 * e + 6] * m.props[7 * e + 0];  // density * A
        double w[3] = {grav_dir[3 * lc + 0] * qsw, grav_dir[3 * lc + 1] * qsw, grav_dir[3 * lc + 2] * qsw};
        double t[12];
        lump_udl(w, R3, L, t);
#pragma unroll
        for (int i = 0; i < 12; ++i) out[i] += t[i];
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) q[((size_t)lc * m.nE + e) * 12 + i] = out[i];
}

}  // namespace

cudaError_t cm_launch_element_tables(const cm_model_dev &m, cudaStream_t st) {
    dim3 grid((m.nE + 127) / 128, m.n_geom);
    k1_element_tables<<<grid, 128, 0, st>>>(m);
    return cudaGetLastError();
}

cudaError_t cm_launch_element_generic(const double *xyz_u, const double *xyz_v, const double *props,
                                      const double *cr, long long n, double *Kg, double *R3, cudaStream_t st) {
    k1_element_generic<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(xyz_u, xyz_v, props, cr, n, Kg, R3);
    return cudaGetLastError();
}

cudaError_t cm_launch_element_loads(const cm_model_dev &m, int n_lc, const double *w_udl,
                                    const unsigned char *has_udl, const unsigned char *grav_on,
                                    const double *grav_dir, double *q, cudaStream_t st) {
    dim3 grid((m.nE + 127) / 128, n_lc, m.n_geom);
    k1_element_loads<<<grid, 128, 0, st>>>(m, n_lc, w_udl, has_udl, grav_on, grav_dir, q);
    return cudaGetLastError();
}
