"""Small driver for ncu captures: G1k frame, B subsets, one warm-up call and one measured call of
the device path with the chosen factorisation variant."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from conmech_b200 import StiffnessChecker, LoadCase, GravityLoad, synthetic as syn      # noqa: E402

variant = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
grid = tuple(int(x) for x in sys.argv[3].split('x')) if len(sys.argv) > 3 else (7, 7, 8)
model = syn.grid_frame(*grid)
sc = StiffnessChecker(model, checker_engine='cuda')
sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
eng = sc._sc_ins
eng.set_factor_variant(variant)
masks = syn.pack_masks([syn.connected_subset(model, s) for s in range(B)], len(model.elements))
dm = torch.from_numpy(masks.view(np.int32)).cuda()
eng.enable_timing(True)
for rep in range(2):
    r = eng.solve_many_device(dm, want=('flag', 'status', 'max_trans', 'compliance', 'profile_flops'))
    torch.cuda.synchronize()
    ms = eng.last_stage_ms()
if os.environ.get('CM_PROF_REPS'):           # a steadier number: mean over more repetitions
    acc = np.zeros(3)
    n = int(os.environ['CM_PROF_REPS'])
    for rep in range(n):
        r = eng.solve_many_device(dm, want=('flag', 'status', 'max_trans', 'compliance', 'profile_flops'))
        torch.cuda.synchronize()
        acc += eng.last_stage_ms()
    ms = acc / n
import hashlib
digest = hashlib.sha1(r['max_trans'].cpu().numpy().tobytes() + r['compliance'].cpu().numpy().tobytes()).hexdigest()[:12]
print('result digest (max_trans, compliance bits):', digest)
fl = float(r['profile_flops'].sum().item())
print('variant {} B={} stage ms {} | K3 {:.2f} TFLOP/s algorithmic | status {} flags {}'.format(variant, B, ms, fl / ms[1] / 1e9, np.bincount(r['status'].cpu().numpy().ravel()), int(r['flag'].sum().item())))
if variant in (4, 5):
    import ctypes as C
    out = np.zeros(16, dtype=np.uint64)
    eng._lib.cm_debug_k3_phase_cycles(eng._handle, out.ctypes.data_as(C.c_void_p))
    names = ['wait_operand', 'update_streams', 'wait_M', 'solve_product', 'diag_updates', 'wait_prev_diag',
             'diag_factor', 'wait_fwd_chain', 'forward_subst', 'total', 'chain_solve', 'chain_last_syrk', 'chain_gap']
    tot = float(out[9])
    for n, v in zip(names, out[:len(names)]):
        print('  {:18s} {:14d} cycles  {:6.2f}%'.format(n, int(v), 100.0 * float(v) / tot))
if os.environ.get('CM_K2_PROF'):
    import ctypes as C
    out = np.zeros(16, dtype=np.uint64)
    eng._lib.cm_debug_k2_phase_cycles(eng._handle, out.ctypes.data_as(C.c_void_p))
    names = ['stage tables/mask', 'node existence', 'numbering+envelope', 'diagonal gather', 'jacobi scale', 'fill+scatter', 'load vectors', 'total',
             'w: zero fill + wait', 'w: row lookup', 'w: diagonal step', 'w: pair steps', 'w: closing barrier']
    tot = float(out[7]) or 1.0
    for n, v in zip(names, out):
        if n.startswith('w:'):
            tot = float(out[8:13].sum()) or 1.0
        print('  K2 {:20s} {:14d} cycles  {:6.2f}%'.format(n, int(v), 100.0 * float(v) / tot))
