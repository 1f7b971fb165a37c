"""Config 4 (joint releases, four load cases per subset, height-field subsets) throughput probe:
one factorisation per subset serves all load cases; a solve = one (subset, load case)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from conmech_b200 import StiffnessChecker, synthetic as syn      # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
model = syn.grid_frame(7, 7, 8, joints=True)
cases = syn.config4_loadcases(model)
sc = StiffnessChecker(model, checker_engine='cuda')
sc.set_load_cases(cases)
eng = sc._sc_ins
masks = syn.pack_masks([syn.heightfield_subset(model, 7, 7, 8, seed=s) for s in range(B)], len(model.elements))
dm = torch.from_numpy(masks.view(np.int32)).cuda()
eng.enable_timing(True)
for rep in range(2):
    r = eng.solve_many_device(dm, want=('flag', 'status', 'max_trans', 'profile_flops'))
    torch.cuda.synchronize()
    ms = eng.last_stage_ms()
n_lc = r['flag'].shape[1]
print('config4 B={} load cases={} stage ms {} -> {:.0f} solves/s (subset x load case); status {}'.format(
    B, n_lc, ms, B * n_lc / ms.sum() * 1e3, np.bincount(r['status'].cpu().numpy().ravel())))
