"""Lane-level emulation of the FP64 MMA m8n8k4 fragment maps used by the factorisation kernel.

No GPU here: the register-level index gymnastics of k3_factor_flow.cu (in-place triangular solve
product from accumulator registers, newest SYRK term from registers, the "Mq" layout of
M_J = D_J^-1 L_JJ^-1, K4's transposed product with M) are checked against plain numpy before a
kernel is sent to the GPU box.      python tools/frag_emulator.py
"""
import numpy as np

LANES = np.arange(32)
G = LANES >> 2
T = LANES & 3


def dmma(c0, c1, a, b):
    """D(8x8) += A(8x4) B(4x8); lane 4g+t holds a=A[g][t], b=B[t][g], c0/c1 = C[g][2t], C[g][2t+1]."""
    A = np.zeros((8, 4)); B = np.zeros((4, 8)); C = np.zeros((8, 8))
    A[G, T] = a
    B[T, G] = b
    C[G, 2 * T] = c0
    C[G, 2 * T + 1] = c1
    C = C + A @ B
    return C[G, 2 * T].copy(), C[G, 2 * T + 1].copy()


def acc_from_matrix(Tm):
    """TileAcc c[cj][e][ri] (each a 32-lane vector) of a 32x32 matrix."""
    c = np.zeros((4, 2, 4, 32))
    for cj in range(4):
        for e in range(2):
            for ri in range(4):
                c[cj, e, ri] = Tm[ri * 8 + G, cj * 8 + 2 * T + e]
    return c


def matrix_from_acc(c):
    Tm = np.zeros((32, 32))
    for cj in range(4):
        for e in range(2):
            for ri in range(4):
                Tm[ri * 8 + G, cj * 8 + 2 * T + e] = c[cj, e, ri]
    return Tm


def mq_off(r, k):
    """cm_mtile_off: Mq[co][cj][lane = 4*(r%8) + (k%8)/2][e = k%2], co = r/8, cj = k/8."""
    return (r >> 3) * 256 + (k >> 3) * 64 + (((r & 7) << 2) + ((k & 7) >> 1)) * 2 + (k & 1)


def mq_from_matrix(M):
    buf = np.zeros(1024)
    for r in range(32):
        for k in range(32):
            buf[mq_off(r, k)] = M[r, k]
    return buf


def solve_inplace(c, mq, cj0=0):
    """acc <- acc * M^T, M lower triangular, block column by block column from the right."""
    for co in range(3, -1, -1):
        r0 = np.zeros((4, 32)); r1 = np.zeros((4, 32))
        for cj in range(co + 1):
            if cj < cj0:
                continue
            bx = mq[co * 256 + cj * 64 + LANES * 2]
            by = mq[co * 256 + cj * 64 + LANES * 2 + 1]
            for ri in range(4):
                r0[ri], r1[ri] = dmma(r0[ri], r1[ri], c[cj, 0, ri], bx)
                r0[ri], r1[ri] = dmma(r0[ri], r1[ri], c[cj, 1, ri], by)
        for ri in range(4):
            c[co, 0, ri] = r0[ri]
            c[co, 1, ri] = r1[ri]
    return c


def syrk_from_regs(diag_c, r, d):
    """lower atoms of diag -= R diag(d) R^T with both operands taken from the registers of R."""
    for cj in range(4):
        for cs in range(4):
            for e in range(2):
                dk = d[cs * 8 + 2 * T + e]
                nb = -(dk * r[cs, e, cj])
                for ri in range(cj, 4):
                    diag_c[cj, 0, ri], diag_c[cj, 1, ri] = dmma(diag_c[cj, 0, ri], diag_c[cj, 1, ri], r[cs, e, ri], nb)
    return diag_c


def fwd_partial(r, y):
    """(R y)[ri*8+g], valid on every lane after the xor-1/xor-2 reduction."""
    out = np.zeros(32)
    for ri in range(4):
        part = np.zeros(32)
        for cj in range(4):
            part += r[cj, 0, ri] * y[cj * 8 + 2 * T] + r[cj, 1, ri] * y[cj * 8 + 2 * T + 1]
        part = part + part[LANES ^ 1]
        part = part + part[LANES ^ 2]
        out[ri * 8 + G[T == 0]] = part[T == 0]
    return out


def m_matvec(mq, rhs):
    """z = M rhs (forward substitution): lane (g,t), row block i."""
    z = np.zeros(32)
    for i in range(4):
        zp = np.zeros(32)
        for cj in range(i + 1):
            zp += mq[i * 256 + cj * 64 + LANES * 2] * rhs[cj * 8 + 2 * T] + mq[i * 256 + cj * 64 + LANES * 2 + 1] * rhs[cj * 8 + 2 * T + 1]
        zp = zp + zp[LANES ^ 1]
        zp = zp + zp[LANES ^ 2]
        z[i * 8 + G[T == 0]] = zp[T == 0]
    return z


def m_tmatvec(mq, w):
    """x = M^T w (back substitution): column 8cj+2t+e, valid on lanes g == 0 after the xor-4/8/16 reduction."""
    x = np.zeros(32)
    for cj in range(4):
        for e in range(2):
            cs = np.zeros(32)
            for i in range(cj, 4):
                cs += mq[i * 256 + cj * 64 + LANES * 2 + e] * w[i * 8 + G]
            for o in (4, 8, 16):
                cs = cs + cs[LANES ^ o]
            x[cj * 8 + 2 * T[:4] + e] = cs[:4]
    return x


def main():
    rng = np.random.default_rng(0)
    Tm = rng.standard_normal((32, 32))
    M = np.tril(rng.standard_normal((32, 32)))
    mq = mq_from_matrix(M)
    c = solve_inplace(acc_from_matrix(Tm), mq)
    assert np.allclose(matrix_from_acc(c), Tm @ M.T, atol=1e-12), 'solve_inplace'
    T2 = Tm.copy(); T2[:, :16] = 0.0
    c = solve_inplace(acc_from_matrix(T2), mq, cj0=2)
    assert np.allclose(matrix_from_acc(c), T2 @ M.T, atol=1e-12), 'solve_inplace cj0'
    R = rng.standard_normal((32, 32)); d = rng.standard_normal(32)
    Dg = rng.standard_normal((32, 32))
    out = matrix_from_acc(syrk_from_regs(acc_from_matrix(Dg), acc_from_matrix(R), d))
    ref = Dg - R @ np.diag(d) @ R.T
    low = np.zeros((32, 32), bool)
    for ri in range(4):
        for cj in range(ri + 1):
            low[ri * 8:ri * 8 + 8, cj * 8:cj * 8 + 8] = True
    assert np.allclose(out[low], ref[low], atol=1e-12), 'syrk'
    y = rng.standard_normal(32)
    assert np.allclose(fwd_partial(acc_from_matrix(R), y), R @ y, atol=1e-12), 'fwd_partial'
    assert np.allclose(m_matvec(mq, y), M @ y, atol=1e-12), 'm_matvec'
    assert np.allclose(m_tmatvec(mq, y), M.T @ y, atol=1e-12), 'm_tmatvec'
    print('fragment maps ok')


if __name__ == '__main__':
    main()
