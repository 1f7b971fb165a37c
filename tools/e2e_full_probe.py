"""Host-buffer path with FULL results (U, Rf, eR + scalars to pinned host arrays) against the device-resident
path, config 3: tools/e2e_full_probe.py [B] [reps].  Wave plan knobs are environment variables of the library
(CM_TAPER, CM_TAPER_MIN, CM_HOST_WAVE_MB), read once per process: one process per setting."""
import os, sys, time
import numpy as np, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from conmech_b200 import StiffnessChecker, LoadCase, GravityLoad, synthetic as syn
B = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
model = syn.grid_frame(7, 7, 8)
nE, nV = len(model.elements), len(model.nodes)
sc = StiffnessChecker(model, checker_engine='cuda')
sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
eng = sc._sc_ins
nb = 3
hms = []
for k in range(nb):
    masks = syn.pack_masks([syn.connected_subset(model, k * B + s) for s in range(B)], nE)
    hms.append(torch.from_numpy(masks.view(np.int32)).pin_memory())
dms = [m.cuda() for m in hms]
hnp = [m.numpy().view(np.uint32) for m in hms]
FULL = ('flag', 'status', 'max_trans', 'compliance', 'U', 'Rf', 'eR')
pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
out_host = {'flag': pin((B, 1), torch.uint8), 'status': pin((B, 1), torch.uint8),
            'max_trans': pin((B, 1), torch.float64), 'compliance': pin((B, 1), torch.float64),
            'U': pin((B, 1, 6 * nV), torch.float64), 'Rf': pin((B, 1, 6 * nV), torch.float64),
            'eR': pin((B, 1, nE, 12), torch.float64)}
if os.environ.get('PROBE_COMPACT'):
    # compact results: reactions of the support nodes, end forces of the existing elements only
    COMPACT = ('flag', 'status', 'max_trans', 'compliance', 'U', 'Rf_sup', 'eR_rows')
    nS = len(model.supports)
    outs_c = []
    for k in range(nb):
        rows = int(eng.element_row_offsets(hnp[k])[-1])
        oc = {n: out_host[n] for n in ('flag', 'status', 'max_trans', 'compliance', 'U')}
        oc['Rf_sup'] = pin((B, 1, nS, 6), torch.float64)
        oc['eR_rows'] = pin((rows, 12), torch.float64)
        outs_c.append(oc)
    for k in range(3):
        eng.solve_many(hnp[k % nb], want=COMPACT, out=outs_c[k % nb])
    t0 = time.perf_counter()
    for k in range(reps):
        eng.solve_many(hnp[k % nb], want=COMPACT, out=outs_c[k % nb])
    c_ms = (time.perf_counter() - t0) / reps * 1e3
    nbytes = sum(outs_c[0][n].nbytes for n in COMPACT)
    print('compact results: host {:.2f} ms = {:.1f} k/s, {:.1f} KB per solve ({} waves)'.format(c_ms, B / c_ms, nbytes / B / 1e3, eng._lib.cm_last_wave_count(eng._handle)))
for k in range(3):
    eng.solve_many(hnp[k % nb], want=FULL, out=out_host)
t0 = time.perf_counter()
for k in range(reps):
    eng.solve_many(hnp[k % nb], want=FULL, out=out_host)
host_ms = (time.perf_counter() - t0) / reps * 1e3
waves = eng._lib.cm_last_wave_count(eng._handle)
fly = int(os.environ['PROBE_FLY']) if os.environ.get('PROBE_FLY') else None
od = None
for k in range(3):
    od = eng.solve_many_device(dms[k % nb], want=FULL, out=od, max_in_flight=fly)
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(reps):
    od = eng.solve_many_device(dms[k % nb], want=FULL, out=od, max_in_flight=fly)
torch.cuda.synchronize()
dev_ms = (time.perf_counter() - t0) / reps * 1e3
eng.solve_many(hnp[0], want=FULL, out=out_host)
od = eng.solve_many_device(dms[0], want=FULL, out=od, max_in_flight=fly)
torch.cuda.synchronize()
same = all(np.array_equal(out_host[n], od[n].cpu().numpy(), equal_nan=True) for n in FULL)
print('FLY={} TAPER={} MIN={} MB={} | host full {:.2f} ms = {:.1f} k/s ({} waves) | device full {:.2f} ms = {:.1f} k/s | ratio {:.3f} | same bits {}'.format(
    fly, os.environ.get('CM_TAPER', '-'), os.environ.get('CM_TAPER_MIN', '-'), os.environ.get('CM_HOST_WAVE_MB', '-'),
    host_ms, B / host_ms, waves, dev_ms, B / dev_ms, dev_ms / host_ms, same))
if os.environ.get('PROBE_COMPACT') and os.environ.get('PROBE_PYTIME'):
    # where the Python side of one call spends its time
    t0 = time.perf_counter(); off = eng.element_row_offsets(hnp[0]); t1 = time.perf_counter()
    print('element_row_offsets {:.3f} ms'.format((t1 - t0) * 1e3))
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable(); eng.solve_many(hnp[0], want=COMPACT, out=outs_c[0]); pr.disable()
    pstats.Stats(pr).sort_stats('cumulative').print_stats(8)
