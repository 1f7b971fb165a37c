"""Where does the host-buffer path spend its time? (debug aid)"""
import os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from conmech_b200 import StiffnessChecker, LoadCase, GravityLoad, synthetic as syn
B = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
model = syn.grid_frame(7, 7, 8)
sc = StiffnessChecker(model, checker_engine='cuda')
sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
eng = sc._sc_ins
masks = syn.pack_masks([syn.connected_subset(model, s) for s in range(B)], len(model.elements))
pm = torch.from_numpy(masks.view(np.int32)).pin_memory()
hm = pm.numpy().view(np.uint32)
dm = pm.cuda()
want = ('flag', 'status', 'max_trans', 'compliance')
pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
out_host = {'flag': pin((B, 1), torch.uint8), 'status': pin((B, 1), torch.uint8),
            'max_trans': pin((B, 1), torch.float64), 'compliance': pin((B, 1), torch.float64)}
for timing in (False, True):
    eng.enable_timing(timing)
    for rep in range(3):
        t0 = time.perf_counter(); eng.solve_many(hm, want=want, out=out_host); t1 = time.perf_counter()
    print('host path   timing={} wall {:.1f} ms stage {}'.format(timing, (t1 - t0) * 1e3, eng.last_stage_ms() if timing else ''))
    od = None
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); od = eng.solve_many_device(dm, want=want, out=od); torch.cuda.synchronize(); t1 = time.perf_counter()
    print('device path timing={} wall {:.1f} ms stage {}'.format(timing, (t1 - t0) * 1e3, eng.last_stage_ms() if timing else ''))
# the facade call a user makes: Python id lists in, flags out (packing by the native packer included)
from conmech_b200 import masks as _m
lists = [_m.mask_to_ids(masks[b], len(model.elements)).tolist() for b in range(B)]
for rep in range(3):
    t0 = time.perf_counter(); flags = sc.solve_many(lists); t1 = time.perf_counter()
print('facade solve_many(list of id lists): wall {:.1f} ms -> {:.0f} solves/s ({} True)'.format((t1 - t0) * 1e3, B / (t1 - t0), int(np.sum(flags))))
t0 = time.perf_counter(); _m.pack_masks(lists, len(model.elements)); t1 = time.perf_counter()
print('  of which packing {:.1f} ms'.format((t1 - t0) * 1e3))
