"""Host-side count of the FP64 MMAs the block-band factorisation (k3_factor_flow) executes on config-3
subsets, by category, against the algorithmic count (sum h_j^2 / 512 per m8n8k4 MMA), and what finer
envelope granularities would save.  Mirrors K2's envelope (kfirst / kcol) and K3's stream ranges.
Usage: python tools/mma_count.py [n_subsets]"""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conmech_b200 import synthetic as syn   # noqa: E402
from tools.envelope_study import rcm         # noqa: E402


def subset_rows(model, order, ids):
    """first column per solver row, n_s"""
    nV = len(model.nodes)
    conn = np.array([e.end_node_inds for e in model.elements])
    fixed = np.zeros(nV, bool)
    for v in model.supports:
        fixed[v] = True
    c = conn[np.asarray(ids)]
    exist = np.zeros(nV, bool)
    exist[c.ravel()] = True
    nfree = np.where(exist & ~fixed, 6, 0)
    order = np.array(order)
    nf_ord = nfree[order]
    base = np.empty(nV, int)
    base[order] = np.concatenate([[0], np.cumsum(nf_ord)])[:-1]
    ns = int(nf_ord.sum())
    first = base.copy()
    u, v = c[:, 0], c[:, 1]
    ok = (nfree[u] > 0) & (nfree[v] > 0)
    np.minimum.at(first, u[ok], base[v[ok]])
    np.minimum.at(first, v[ok], base[u[ok]])
    rf = np.zeros(ns, int)
    for n in np.flatnonzero(nfree):
        rf[base[n]:base[n] + 6] = first[n]
    return rf, ns


def count(rf, ns, T=32):
    NT = (ns + T - 1) // T
    pad = NT * T
    rfp = np.concatenate([rf, np.arange(ns, pad)])          # padding rows: unit diagonal
    alg = float(((np.arange(ns) - rf) ** 2).sum()) / 256.0 / 2 * 2 / 2   # MACs h^2/2 -> MMAs of 256 MACs
    alg = float(((np.arange(ns) - rf) ** 2).sum()) / 512.0
    kcol = np.array([(rfp[I * T:(I + 1) * T].min() // 4) & ~1 for I in range(NT)])     # k-steps of 4, even
    kfirst = kcol // 8
    # row-block (8 rows) first k-step
    rb = np.array([[(rfp[I * T + 8 * r:I * T + 8 * r + 8].min() // 4) for r in range(4)] for I in range(NT)])
    stream = stream_rb = stream_rb2 = 0
    solve = dterm = 0
    ntiles = 0
    for I in range(NT):
        for J in range(kfirst[I], I):
            ntiles += 1
            qs = max(kcol[I], kcol[J])
            n = max(0, 8 * J - qs)
            stream += 16 * n
            # per (ri, cj) atom: starts at max(rb[I][ri], rb[J][cj]) rounded down to even
            for ri in range(4):
                for cj in range(4):
                    s = max(rb[I][ri], rb[J][cj])
                    stream_rb += max(0, 8 * J - s)
                    s2 = max(rb[I][ri] & ~1, rb[J][cj] & ~1)
                    stream_rb2 += max(0, 8 * J - s2)
            cj0 = max(kcol[I] - 8 * J, 0) >> 1
            # solve: for co in 0..3: for cj in cj0..co: 8 MMAs
            solve += sum(8 for co in range(4) for cj in range(co + 1) if cj >= cj0)
            dterm += sum(2 * (4 - cj) for cj in range(4) for cs in range(4) if cs >= cj0)
    diag = NT * (3 * 2 * 6 // 2 + 40)       # rough: trailing updates + inverse products
    return dict(alg=alg, stream=stream, stream_rb=stream_rb, stream_rb2=stream_rb2, solve=solve, dterm=dterm, diag=diag,
                ntiles=ntiles, NT=NT)


if __name__ == '__main__':
    nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    model = syn.grid_frame(7, 7, 8)
    nV = len(model.nodes)
    adj = [set() for _ in range(nV)]
    for e in model.elements:
        u, v = e.end_node_inds
        adj[u].add(v); adj[v].add(u)
    adj = [sorted(a) for a in adj]
    order = rcm(nV, adj)
    tot = {}
    for s in range(nsub):
        rf, ns = subset_rows(model, order, syn.connected_subset(model, s))
        c = count(rf, ns)
        for k, v in c.items():
            tot[k] = tot.get(k, 0) + v
    for k in tot:
        tot[k] /= nsub
    ex = tot['stream'] + tot['solve'] + tot['dterm'] + tot['diag']
    print({k: round(v, 1) for k, v in tot.items()})
    print('executed/alg = {:.3f}; with row-block masks (k-step granularity) {:.3f}; (even k-step) {:.3f}'.format(
        ex / tot['alg'], (ex - tot['stream'] + tot['stream_rb']) / tot['alg'], (ex - tot['stream'] + tot['stream_rb2']) / tot['alg']))
