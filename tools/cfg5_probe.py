"""Config 5 (10k-element frame, few subsets) throughput probe."""
import os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from conmech_b200 import StiffnessChecker, LoadCase, GravityLoad, synthetic as syn
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
model = syn.grid_frame(15, 15, 16)
sc = StiffnessChecker(model, checker_engine='cuda')
sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
eng = sc._sc_ins
print('info', eng.info)
t0 = time.time()
masks = syn.pack_masks([syn.connected_subset(model, s) for s in range(B)], len(model.elements))
print('sampler {:.1f} s'.format(time.time() - t0))
dm = torch.from_numpy(masks.view(np.int32)).cuda()
eng.enable_timing(True)
if len(sys.argv) > 2:
    eng.set_factor_variant(int(sys.argv[2]))
for rep in range(2):
    r = eng.solve_many_device(dm, want=('flag', 'status', 'max_trans', 'profile_flops'), max_in_flight=B)
    torch.cuda.synchronize()
    ms = eng.last_stage_ms()
fl = float(r['profile_flops'].sum().item())
print('B={} stage ms {} -> {:.1f} solves/s; K3 {:.2f} TFLOP/s algorithmic; status {}'.format(
    B, ms, B / ms.sum() * 1e3, fl / ms[1] / 1e9, np.bincount(r['status'].cpu().numpy().ravel())))
