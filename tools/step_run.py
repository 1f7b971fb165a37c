"""One bench step, nothing else (for ncu launch lists): config 3, B subsets, FULL results on the device, `steps` timed
calls after `warmup`: python tools/step_run.py [B] [steps] [warmup]"""
import os, sys
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from conmech_b200 import StiffnessChecker, LoadCase, GravityLoad, synthetic as syn
B = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
warmup = int(sys.argv[3]) if len(sys.argv) > 3 else 1
model = syn.grid_frame(7, 7, 8)
sc = StiffnessChecker(model, checker_engine='cuda')
sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
eng = sc._sc_ins
masks = syn.pack_masks([syn.connected_subset(model, s) for s in range(B)], len(model.elements))
dm = torch.from_numpy(masks.view(np.int32)).cuda()
FULL = ('flag', 'status', 'max_trans', 'compliance', 'U', 'Rf', 'eR')
out = None
for k in range(warmup):
    out = eng.solve_many_device(dm, want=FULL, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(steps):
    out = eng.solve_many_device(dm, want=FULL, out=out)
e1.record()
torch.cuda.synchronize()
print('{} steps of {} subsets: {:.2f} ms per step, {} kernel launches in all'.format(steps, B, e0.elapsed_time(e1) / steps, eng.kernel_launches()))
