"""Compare the factorisation workspace of two kernel variants for one subset (debug aid)."""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from conmech_b200 import StiffnessChecker, LoadCase, GravityLoad, synthetic as syn
va, vb, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
model = syn.grid_frame(7, 7, 8)
sc = StiffnessChecker(model, checker_engine='cuda')
sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
eng = sc._sc_ins
info = eng.info
masks = syn.pack_masks([syn.connected_subset(model, seed)], len(model.elements))
dm = torch.from_numpy(masks.view(np.int32)).cuda()
al = lambda x: (x + 255) & ~255
ntr, bt, nV = info['n_tile_rows_max'], info['band_tiles'], info['n_nodes']
npad = ntr * 32
o = 0
off = {}
for name, size in (('tiles', ntr * bt * 1024 * 8), ('sidx', 6 * nV * 4), ('dinv', npad * 8), ('dpiv', (npad + 8) * 8),
                   ('x', npad * 8), ('p', 6 * nV * 8), ('kfirst', ntr * 4), ('minv', ntr * 1024 * 8), ('info', 64)):
    off[name] = o
    o = al(o + size)
snaps = {}
for v in (va, vb):
    eng.set_factor_variant(v)
    r = eng.solve_many_device(dm, want=('flag', 'status', 'min_pivot'))
    torch.cuda.synchronize()
    ws = list(eng._ws_cache.values())[0]
    snaps[v] = ws.cpu().numpy().copy()
    print('variant', v, 'status', r['status'].cpu().numpy().ravel(), 'minpiv', r['min_pivot'].cpu().numpy())
a, b = snaps[va], snaps[vb]
inf = a[off['info']:off['info'] + 64].view(np.int32)
NT = int(inf[1])
print('NS, NT', inf[0], NT, 'kfirst', a[off['kfirst']:off['kfirst'] + 4 * NT].view(np.int32))
da = a[off['dpiv']:off['dpiv'] + npad * 8].view(np.float64)
db = b[off['dpiv']:off['dpiv'] + npad * 8].view(np.float64)
for J in range(NT):
    e = np.abs(da[J * 32:J * 32 + 32] - db[J * 32:J * 32 + 32]).max()
    if not (e < 1e-9):
        print('first pivot difference in column', J, 'err', e)
        print(da[J * 32:J * 32 + 32]); print(db[J * 32:J * 32 + 32])
        break
ta = a[:ntr * bt * 8192].view(np.float64).reshape(ntr, bt, 1024)
tb = b[:ntr * bt * 8192].view(np.float64).reshape(ntr, bt, 1024)
kf = a[off['kfirst']:off['kfirst'] + 4 * NT].view(np.int32)
bad = []
for I in range(NT):
    for K in range(kf[I], I):
        e = np.abs(ta[I, K - I + bt - 1] - tb[I, K - I + bt - 1]).max()
        if not (e < 1e-9):
            bad.append((I, K, float(e)))
print('differing tiles (I,K,err) first 20:', bad[:20])
ma = a[off['minv']:off['minv'] + ntr * 8192].view(np.float64).reshape(ntr, 1024)
mb = b[off['minv']:off['minv'] + ntr * 8192].view(np.float64).reshape(ntr, 1024)
print('minv[0] max diff', np.abs(ma[0] - mb[0]).max(), 'minv[1] max diff', np.abs(ma[1] - mb[1]).max())
def untile(t):
    out = np.zeros((32, 32))
    for r in range(32):
        for k in range(32):
            out[r, k] = t[((k >> 2) << 7) + ((((r & 7) << 2) + (k & 3)) << 2) + (r >> 3)]
    return out
d10 = untile(ta[1, 0 - 1 + bt - 1]) - untile(tb[1, 0 - 1 + bt - 1])
np.set_printoptions(linewidth=250, precision=1, threshold=100000)
print('tile(1,0) diff pattern (rows x cols, 1 = differs):')
print((np.abs(d10) > 1e-9).astype(int))
xa = a[off['x']:off['x'] + npad * 8].view(np.float64); xb = b[off['x']:off['x'] + npad * 8].view(np.float64)
print('x diff first block', np.abs(xa[:32] - xb[:32]).max())

d20 = untile(ta[2, 0 - 2 + bt - 1]) - untile(tb[2, 0 - 2 + bt - 1])
print('tile(2,0) diff pattern:')
print((np.abs(d20) > 1e-9).astype(int))
print('xa[:40]', xa[:40]); print('xb[:40]', xb[:40])
