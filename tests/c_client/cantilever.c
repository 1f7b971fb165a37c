/* A plain C client of include/conmech_b200.h - no Python, no torch: the drop-in boundary used the way a host
 * written in C (or cgo / JNI / N-API stubs) would use it.  A cantilever along x of N elements, clamped at node 0,
 * tip load Fz and tip load Fy: Euler-Bernoulli gives the tip deflection P L^3 / (3 E I) and the clamp reactions in
 * closed form (the same textbook checks as the reference's tests, tests/python/test_stiffness_checker.py:573-632).
 * Also: a subset that keeps only the first half of the beam (shorter cantilever, its loaded node absent: no motion),
 * a subset without the clamped element (no support: status 1), compact results against dense ones.
 *   gcc -O2 -I include tests/c_client/cantilever.c -L conmech_b200/lib -lconmech_b200 -lm -o cantilever */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "conmech_b200.h"

#define N 8
#define CHECK(call) do { int rc_ = (call); if (rc_ != CM_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, cm_last_error()); return 2; } } while (0)

#define EXPECT(cond) do { if (!(cond)) { fprintf(stderr, "FAILED CHECK: %s\n", #cond); ok = 0; } } while (0)

static int close_to(double a, double b, double rel, const char *what) {
    if (fabs(a - b) <= rel * fmax(fabs(b), 1e-300)) return 1;
    fprintf(stderr, "MISMATCH %s: got %.15g expected %.15g\n", what, a, b);
    return 0;
}

int main(void) {
    const double L = 4.0, E = 2.1e8, mu = 0.3, r = 0.05, PI = 3.141592653589793;
    const double A = PI * r * r, Jx = PI * pow(r, 4) / 2, I = PI * pow(r, 4) / 4;
    double xyz[(N + 1) * 3] = {0}, props[N * 7], cr[N * 6];
    int32_t conn[N * 2];
    for (int v = 0; v <= N; ++v) xyz[3 * v] = L * v / N;
    for (int e = 0; e < N; ++e) {
        conn[2 * e] = e; conn[2 * e + 1] = e + 1;
        const double p[7] = {A, Jx, I, I, E, mu, 0.0};
        memcpy(props + 7 * e, p, sizeof(p));
        for (int k = 0; k < 6; ++k) cr[6 * e + k] = INFINITY;     /* rigid joints */
    }
    const int32_t sup_node[1] = {0};
    const uint8_t sup_cond[6] = {1, 1, 1, 1, 1, 1};
    cm_model_desc d = {N + 1, N, 1, 0, xyz, conn, props, cr, sup_node, sup_cond, NULL};
    cm_handle *h = NULL;
    CHECK(cm_create(&d, &h));
    cm_model_info info;
    CHECK(cm_get_info(h, &info));
    if (info.n_free_full != 6 * N || info.mask_words != 1) { fprintf(stderr, "unexpected model info\n"); return 3; }

    /* two load cases: tip force -z, tip force +y */
    const double P = 3.5;
    static double nodal[2 * 6 * (N + 1)], wudl[2 * N * 3], gdir[2 * 3];
    static uint8_t gon[2];
    nodal[6 * N + 2] = -P;
    nodal[6 * (N + 1) + 6 * N + 1] = P;
    CHECK(cm_set_loadcases(h, 2, nodal, wudl, NULL, gon, gdir));
    CHECK(cm_set_tol(h, 1e-2, 0.0));

    /* three subsets: everything; the clamped half (no loaded node -> zero load); the free half (no support) */
    enum { B = 3, LC = 2 };
    uint32_t masks[B] = {(1u << N) - 1u, (1u << (N / 2)) - 1u, ((1u << N) - 1u) & ~((1u << (N / 2)) - 1u)};
    static uint8_t flag[B * LC], status[B * LC];
    static double U[B * LC * 6 * (N + 1)], Rf[B * LC * 6 * (N + 1)], eR[B * LC * N * 12], comp[B * LC], mt[B * LC];
    static double Rs[B * LC * 6], rows[(N + N / 2 + N / 2) * LC * 12];
    int64_t off[B + 1];
    CHECK(cm_element_row_offsets(masks, B, 1, off));
    if (off[B] != N + N / 2 + N / 2) { fprintf(stderr, "row offsets\n"); return 3; }
    cm_solve_out_host o;
    memset(&o, 0, sizeof(o));
    o.flag = flag; o.status = status; o.U = U; o.Rf = Rf; o.eR = eR; o.compliance = comp; o.max_trans = mt;
    o.Rf_sup = Rs; o.eR_rows = rows; o.eR_row_offsets = off;
    CHECK(cm_solve_many_host(h, masks, B, &o));

    int ok = 1;
    const double tip = P * L * L * L / (3 * E * I), rot = P * L * L / (2 * E * I);
    /* subset 0, load case 0: tip -z */
    EXPECT(status[0] == CM_ST_OK && status[1] == CM_ST_OK);
    ok &= close_to(U[6 * N + 2], -tip, 1e-9, "tip deflection z");
    ok &= close_to(U[6 * N + 4], rot, 1e-9, "tip rotation about y");
    ok &= close_to(Rf[2], P, 1e-9, "clamp force z");
    ok &= close_to(Rf[4], -P * L, 1e-9, "clamp moment y");
    ok &= close_to(comp[0], 0.5 * P * tip, 1e-9, "compliance");
    ok &= close_to(mt[0], tip, 1e-9, "max translation");
    EXPECT(flag[0] == (tip < 1e-2));
    /* load case 1: tip +y */
    const double *U1 = U + 6 * (N + 1), *R1 = Rf + 6 * (N + 1);
    ok &= close_to(U1[6 * N + 1], tip, 1e-9, "tip deflection y");
    ok &= close_to(R1[1], -P, 1e-9, "clamp force y");
    ok &= close_to(R1[5], -P * L, 1e-9, "clamp moment z");
    /* subset 1: the loaded node does not exist - the existing half is unloaded and does not move (the reference keeps
     * the load in the norm test of numpy_stiffness.py:267, so this is a regular solve with a zero right-hand side);
     * subset 2: no support */
    fprintf(stderr, "status per (subset, load case): %d %d | %d %d | %d %d\n", status[0], status[1], status[2], status[3], status[4], status[5]);
    EXPECT(status[2] == CM_ST_OK && status[3] == CM_ST_OK && flag[2] == 1 && flag[3] == 1);
    EXPECT(mt[2] == 0.0 && comp[2] == 0.0);
    EXPECT(status[4] == CM_ST_NO_SUPPORT && flag[4] == 0 && status[5] == CM_ST_NO_SUPPORT);
    /* compact results == dense results without the padding */
    for (int lc = 0; lc < LC; ++lc) {
        for (int k = 0; k < 6; ++k) EXPECT(Rs[lc * 6 + k] == Rf[lc * 6 * (N + 1) + k]);
        for (int e = 0; e < N; ++e)
            for (int k = 0; k < 12; ++k) EXPECT(rows[(lc * N + e) * 12 + k] == eR[(lc * N + e) * 12 + k]);
    }
    /* the element end forces of the clamped element: shear P, moment P L at the clamp (local axes, Q4 sign) */
    ok &= close_to(fabs(eR[2]), P, 1e-9, "element 0 end shear") & close_to(fabs(eR[4]), P * L, 1e-9, "element 0 end moment");
    CHECK(cm_destroy(h));
    printf("%s: cantilever of %d elements, tip deflection %.6e (analytic %.6e), %s\n", ok ? "OK" : "FAILED", N, -U[6 * N + 2], tip, cm_version());
    return ok ? 0 : 1;
}
