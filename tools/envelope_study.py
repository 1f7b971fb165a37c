"""Host-side study of the factorisation work of config-3 subsets: algorithmic profile flops
(sum h_j^2) against the flops a tiled block-band LDL^T executes at several tile sizes and node
orderings.  Pure numpy; no GPU.  Usage: python tools/envelope_study.py [n_subsets]"""
import sys
import os
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conmech_b200 import synthetic as syn   # noqa: E402


def rcm(nV, adj):
    """Same rule as cm_api.cu rcm_order (pseudo-peripheral root, degree-sorted BFS, reversed)."""
    from collections import deque
    order, seen = [], [False] * nV
    for s in range(nV):
        if seen[s]:
            continue
        root, ecc = s, -1
        for _ in range(8):
            level = {root: 0}
            q = deque([root])
            comp = []
            while q:
                v = q.popleft()
                comp.append(v)
                for w in adj[v]:
                    if w not in level:
                        level[w] = level[v] + 1
                        q.append(w)
            last = level[comp[-1]]
            best = min((v for v in comp if level[v] == last), key=lambda v: (len(adj[v]),))
            if last <= ecc:
                break
            ecc, root = last, best
        cm = []
        q = deque([root])
        seen[root] = True
        while q:
            v = q.popleft()
            cm.append(v)
            nb = [w for w in adj[v] if not seen[w]]
            for w in nb:
                seen[w] = True
            nb.sort(key=lambda a: (len(adj[a]), a))
            q.extend(nb)
        order.extend(reversed(cm))
    return order


def study(model, order, subsets, tiles=(32, 16, 8)):
    nV = len(model.nodes)
    conn = np.array([e.end_node_inds for e in model.elements])
    fixed = np.zeros(nV, bool)
    for v in model.supports:
        fixed[v] = True
    pos = np.empty(nV, int)
    pos[np.array(order)] = np.arange(nV)
    out = {T: [] for T in tiles}
    alg = []
    ns_l = []
    for ids in subsets:
        ids = np.asarray(ids)
        c = conn[ids]
        exist = np.zeros(nV, bool)
        exist[c.ravel()] = True
        nfree = np.where(exist & ~fixed, 6, 0)
        nf_ord = nfree[np.array(order)]
        base_ord = np.concatenate([[0], np.cumsum(nf_ord)])[:-1]
        base = np.empty(nV, int)
        base[np.array(order)] = base_ord
        ns = int(nf_ord.sum())
        first = base.copy()
        u, v = c[:, 0], c[:, 1]
        ok = (nfree[u] > 0) & (nfree[v] > 0)
        np.minimum.at(first, u[ok], base[v[ok]])
        np.minimum.at(first, v[ok], base[u[ok]])
        # per row: first column
        rows_first = np.full(ns, 0)
        a = 0.0
        for n in np.flatnonzero(nfree):
            b = base[n]
            rows_first[b:b + 6] = first[n]
            h = np.arange(b, b + 6) - first[n]
            a += float((h * h).sum())
        alg.append(a)
        ns_l.append(ns)
        for T in tiles:
            NT = (ns + T - 1) // T
            kf = np.arange(NT)
            rf = rows_first // T
            for I in range(NT):
                kf[I] = min(I, rf[I * T:(I + 1) * T].min())
            ops = 0
            for I in range(NT):
                for J in range(kf[I], I + 1):
                    k0 = max(kf[I], kf[J])
                    ops += (J - k0) + (0.5 if J < I else 1.0 / 3)    # gemm terms + trsm / ldlt
            out[T].append(ops * 2.0 * T ** 3)
    return np.array(alg), {T: np.array(v) for T, v in out.items()}, np.array(ns_l)


if __name__ == '__main__':
    nsub = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    model = syn.grid_frame(7, 7, 8)
    nV = len(model.nodes)
    adj = [set() for _ in range(nV)]
    for e in model.elements:
        u, v = e.end_node_inds
        adj[u].add(v)
        adj[v].add(u)
    adj = [sorted(a) for a in adj]
    subsets = [syn.connected_subset(model, s) for s in range(nsub)]
    full = [list(range(len(model.elements)))]
    for name, order in (('natural', list(range(nV))), ('rcm', rcm(nV, adj))):
        for label, ss in (('full', full), ('subsets', subsets)):
            alg, ex, ns = study(model, order, ss)
            print('{:8s} {:8s} n_s {:7.1f}  algorithmic {:.3e}'.format(name, label, ns.mean(), alg.mean()),
                  '  '.join('T{}: {:.3e} (x{:.2f})'.format(T, v.mean(), v.mean() / alg.mean()) for T, v in ex.items()))
