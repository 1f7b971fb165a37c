import os, sys, time
import numpy as np
sys.path.insert(0, '/root/repo')
from conmech_b200 import StiffnessChecker, LoadCase, GravityLoad, synthetic as syn, masks as _m
B = 12500
model = syn.grid_frame(7, 7, 8)
sc = StiffnessChecker(model, checker_engine='cuda')
sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
masks = syn.pack_masks([syn.connected_subset(model, s) for s in range(B)], len(model.elements))
lists = [_m.mask_to_ids(masks[b], len(model.elements)).tolist() for b in range(B)]
for rep in range(3):
    sc.solve_many(lists)
t0 = time.perf_counter()
for rep in range(6):
    flags = sc.solve_many(lists)
dt = (time.perf_counter() - t0) / 6
print('facade solve_many(id lists): {:.1f} ms -> {:.1f} k solves/s'.format(dt * 1e3, B / dt / 1e3))
t0 = time.perf_counter(); _m.pack_masks(lists, len(model.elements)); print('packing alone {:.1f} ms'.format((time.perf_counter() - t0) * 1e3))
r = sc._sc_ins.solve_many(masks, want=('flag',))
assert [bool(f) for f in r['flag'][:, 0]] == list(flags)
