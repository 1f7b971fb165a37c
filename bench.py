#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's headline metric: partial-structure solves/sec on
the ~1k-element space frame (config 3: grid 7x7x8, 931 elements, 2058 free DOFs, random connected
element subsets of 20-90 %, gravity load, trans_tol 1e-3), subsets range-partitioned over GPUs.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference itself on the host CPU

One "step" = one pass of the hot path (assembly -> factorisation -> recovery) over one batch of B
subsets per GPU (default 12 500, i.e. 100 000 across 8 GPUs), with the FULL result of every solve -
displacements U, fixity reactions Rf, element end forces eR and the scalars flag / status /
max_trans / compliance - materialised, as the reference does on every solve
(numpy_stiffness.py:366-391).

  value             device-resident throughput: masks already in HBM, full results written to HBM,
                    CUDA events on the launching stream, max over ranks
  value_flags_only  the same with only the four scalars per solve requested (the stiffness-CHECK
                    variant: no U / Rf / eR stores, no reaction recovery)
  e2e               through the host-buffer C-ABI call (cm_solve_many_host): pinned host masks in,
                    full results out to pinned host arrays, host<->device copies inside the timed
                    region; e2e.flags_only_value = the four scalars out only; e2e.id_lists_value =
                    StiffnessChecker.solve_many(list of Python id lists) -> list of bools (the facade
                    call, packing included)
  roofline          the dominant kernel (K3, FP64 tensor pipe) + roofline.kernels = K2 / K3 / K4
  other_configs     (N = 1) BASELINE configs 1, 2, 4, 5 on the GPU next to the CPU rate of each
  parity_in_run     (N = 1) the CUDA results of this run against the CPU arm's results for the same first subsets
                    (flags equal, U / Rf / eR / max_trans / compliance within 1e-9)
  single_process    (N > 1, rank 0 after the SPMD legs) ONE process driving all N GPUs through
                    cm_solve_many_host_multi, host gather inside the timed region
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

GRID = (7, 7, 8)
TRANS_TOL = 1e-3
METRIC = 'partial-structure solves/sec (1k-elem frame)'
UNIT = 'solves/s'
FP64_FALLBACK_TFLOPS = 35.5      # cuBLAS DGEMM 8192^3 measured on this pool's B200 (round 1)
HBM_FALLBACK_GBS = 6549.8        # MEASURED_PEAKS.json of this pool (round 1)
# DRAM bytes (read + write) per subset of ONE launch over 4096 config-3 subsets, from the ncu --set full
# capture summarised in profiles/ (file named in `traffic_source`)
NCU_DRAM_BYTES_PER_SUBSET = {'k2': (0.407226e9 + 7.938545e9) / 4096, 'k3': (23.888439e9 + 9.466585e9) / 4096,
                             'k4': (11.314762e9 + 0.017875e9) / 4096}      # (K4 captured with the scalar outputs only)
NCU_SOURCE = 'profiles/r02d_ncu_k2_k3_k4.txt'
FULL = ('flag', 'status', 'max_trans', 'compliance', 'U', 'Rf', 'eR')
SCALARS = ('flag', 'status', 'max_trans', 'compliance')


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--batch', type=int, default=12500, help='subsets per GPU per step')
    ap.add_argument('--cpu-sample', type=int, default=0, help='subsets in the CPU baseline sample (0 = auto)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--parity-sample', type=int, default=128, help='subsets of the in-run parity check against the CPU arm')
    ap.add_argument('--no-other-configs', action='store_true')
    ap.add_argument('--distinct-batches', type=int, default=2)
    return ap.parse_args()


def workload_config(args, extra=None):
    cfg = {'workload': 'config3: synthetic space-frame grid 7x7x8 (392 nodes / 931 elements / 2058 free DOFs), '
                       'gravity [0,0,-1], random connected element subsets U(0.2,0.9)*nE (seeded), trans_tol 1e-3',
           'subsets_per_gpu_per_step': args.batch,
           'global_subsets_per_step': args.batch * args.gpus,
           'load_cases': 1,
           'results': 'full: U, Rf, eR + flag/status/max_trans/compliance per solve (127 KB per solve)',
           'partition': 'contiguous seed ranges per GPU, no collective on the solve path',
           'l2': 'per-step streamed working set (block-band factors, ~4.8 MB/subset) is >> 126 MB L2; '
                 'consecutive steps use different subset batches; element tables (1.1 MB) stay L2-resident by design'}
    if extra:
        cfg.update(extra)
    return cfg


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
        'clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.1:
                continue
            f = [x.strip() for x in line.split(',')]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nme)
        if not sm:        # region shorter than a sample period: take the nearest sample
            for ts, line in self.rows[-3:]:
                f = [x.strip() for x in line.split(',')]
                try:
                    sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
                except (ValueError, IndexError):
                    pass
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def run_reference(args):
    """The reference's own CPU implementation of the path on the box's host cores: the unmodified
    reference from baseline/_ref (kind "reference"), else the oracle port (kind "port").  Rank 0 only;
    every step solves a bounded sample of the same workload on all cores."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    from oracle import cpu_bench
    n_procs = os.cpu_count() or 1
    kind = 'reference' if cpu_bench.reference_available() else 'port'
    per_step = args.cpu_sample or (30 if kind == 'reference' else 100) * n_procs          # ~3 s per step
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    seeds_all = list(range((args.warmup + args.steps) * per_step))
    spec = dict(cpu_bench.CONFIG3, model=('grid',) + GRID + (False,), trans_tol=TRANS_TOL)
    with ctx.Pool(n_procs, initializer=cpu_bench._init_worker, initargs=(spec, kind)) as pool:
        def step(k):
            seeds = seeds_all[k * per_step:(k + 1) * per_step]
            chunks = [c for c in (seeds[i::n_procs] for i in range(n_procs)) if c]
            t0 = time.perf_counter()
            pool.map(cpu_bench._solve_seeds, chunks)
            return time.perf_counter() - t0
        for k in range(args.warmup):
            step(k)
        dts = [step(args.warmup + k) for k in range(args.steps)]
    total = sum(dts)
    value = per_step * args.steps / total
    what = ('unmodified reference: pyconmech StiffnessChecker(checker_engine="numpy").solve(ids) from baseline/_ref'
            if kind == 'reference' else 'oracle port oracle/frame_oracle.py (baseline/_ref absent)')
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(args),        # identical to the CUDA arm's `config`
            'reference_sample_per_step': per_step,
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': n_procs, 'kind': kind, 'what': what,
                             'sample': '{} subsets per step (seeds 0..), {} processes, {}'.format(
                                 per_step, n_procs, cpu_bench.cpu_model_name())},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def bind_to_gpu_numa_node(index):
    """One process per GPU: run this rank's host threads - and so place its pinned result buffers (first touch) - on
    the CPUs of the NUMA node the GPU hangs off, so that eight ranks' device->host result copies do not all land in
    one socket's memory.  Best effort; returns a note for the bench line."""
    try:
        out = subprocess.run(['nvidia-smi', '-i', str(index), '--query-gpu=pci.bus_id', '--format=csv,noheader'],
                             capture_output=True, text=True).stdout.strip().lower()
        bus = out[-12:] if len(out) >= 12 else out            # 00000000:1B:00.0 -> 0000:1b:00.0
        with open('/sys/bus/pci/devices/{}/local_cpulist'.format(bus)) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(','):
            if '-' in part:
                a, b = part.split('-')
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return 'rank bound to {} CPUs local to GPU {} ({})'.format(len(cpus), index, spec)
    except Exception as ex:
        return 'numa binding skipped: {!r}'.format(ex)
    return 'numa binding skipped: no local cpu list'


def measured_peaks():
    try:
        with open(os.path.join(REPO, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (burst copy rate; driver-written)'
    except (OSError, KeyError, ValueError):
        return HBM_FALLBACK_GBS, 'fallback: the copy rate MEASURED_PEAKS.json held in round 1 (file absent)'


def parity_in_run(ref, got, masks, n_elems, tol=1e-9):
    """Results of THIS run's CUDA path against the CPU arm's results for the same subsets (BASELINE.md section 3,
    step 6).  ``ref``: the arrays oracle.cpu_bench.parity_sample stored (flag, has, max_trans, compliance, U, R, eR
    with NaN rows for absent elements); ``got``: flag / status / max_trans / compliance / U / Rf / eR of the first
    ``len(ref['flag'])`` subsets as numpy arrays shaped [n, 1, ...]; ``masks``: their uint32 bit masks.  Flags must be
    equal; values are compared with the survey's metric - max |a - b| / max |b| per array and solve, translations /
    forces and rotations / moments separately - against ``tol``."""
    import numpy as np
    from conmech_b200 import masks as cm_masks

    def rel(a, b):
        a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
        return 0.0 if b.size == 0 else float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

    def split(a, b):
        a, b = np.asarray(a, dtype=float).reshape(-1, 6), np.asarray(b, dtype=float).reshape(-1, 6)
        return max(rel(a[:, :3], b[:, :3]), rel(a[:, 3:], b[:, 3:]))

    n = int(len(ref['flag']))
    gflag = np.asarray(got['flag']).reshape(len(got['flag']), -1)[:n, 0].astype(bool)
    gstat = np.asarray(got['status']).reshape(len(got['status']), -1)[:n, 0]
    has = np.asarray(ref['has']).astype(bool)
    worst = {'U': 0.0, 'Rf': 0.0, 'eR': 0.0, 'max_trans': 0.0, 'compliance': 0.0}
    for b in np.flatnonzero(has):
        ids = cm_masks.mask_to_ids(masks[b], n_elems)
        worst['U'] = max(worst['U'], split(got['U'][b, 0], ref['U'][b]))
        worst['Rf'] = max(worst['Rf'], split(got['Rf'][b, 0], ref['R'][b]))
        worst['eR'] = max(worst['eR'], split(np.asarray(got['eR'][b, 0]).reshape(n_elems, 12)[ids], ref['eR'][b][ids]))
        worst['max_trans'] = max(worst['max_trans'], rel(got['max_trans'][b, 0], ref['max_trans'][b]))
        worst['compliance'] = max(worst['compliance'], rel(got['compliance'][b, 0], ref['compliance'][b]))
    flags_equal = bool(np.array_equal(gflag, np.asarray(ref['flag']).astype(bool)))
    solved_equal = bool(np.array_equal((gstat == 0) | (gstat == 4), has))       # 0 ok, 4 zero load: a result is stored
    return {'subsets': n, 'against': str(ref['kind']) if 'kind' in ref else 'unknown', 'flags_equal': flags_equal,
            'solved_equal': solved_equal, 'flags_true': int(gflag.sum()), 'with_result': int(has.sum()),
            'max_rel_err': worst, 'tolerance': tol,
            'pass': bool(flags_equal and solved_equal and max(worst.values()) <= tol)}


def other_configs_gpu(torch, np, dev, cpu_rates):
    """BASELINE configs 1, 2, 4 and 5 on one GPU (full results on the device, CUDA events), each
    next to the CPU rate of the same workload."""
    from conmech_b200 import StiffnessChecker, LoadCase, GravityLoad, Model
    from conmech_b200 import synthetic as syn
    out = {}

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    def with_cpu(name, rec):
        if cpu_rates and name in cpu_rates:
            rec['cpu'] = cpu_rates[name]
        out[name] = rec

    # config 1 / 2: the bundled tower, single solve() latency through the facade and the prefix batch
    path = os.path.join(REPO, 'tests', 'golden', 'models', 'tower_3D.json')
    model = Model.from_json(path)
    sc = StiffnessChecker(model, checker_engine='cuda', device=dev.index)
    sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
    sc.set_nodal_displacement_tol(trans_tol=1e-3)
    for _ in range(5):
        sc.solve([])
    t0 = time.perf_counter()
    n = 200
    for _ in range(n):
        ok = sc.solve([])
    lat = (time.perf_counter() - t0) / n
    with_cpu('config1_tower_full_solve', {'value': 1.0 / lat, 'unit': UNIT, 'latency_ms': 1e3 * lat, 'flag': bool(ok),
                                           'what': 'StiffnessChecker.solve([]) on tower_3D (24 elements), host wall clock per call, '
                                                   'nD / fR / eR stored like the reference'})
    order = syn.bfs_sequence(model)
    t0 = time.perf_counter()
    for _ in range(50):
        flags = sc.solve_sequence(order)
    dt = (time.perf_counter() - t0) / 50
    with_cpu('config2_tower_bfs_prefixes', {'value': len(order) / dt, 'unit': UNIT, 'ms_per_sequence': 1e3 * dt,
                                             'flags': ''.join('1' if f else '0' for f in flags),
                                             'what': 'solve_sequence(order): all 24 prefixes of the BFS construction sequence in one call'})
    del sc

    # config 2 on the largest bundled model: every prefix of the BFS sequence of cantilever_beam_truss (2197 elements)
    from conmech_b200 import LoadCase as LC
    mdir = os.path.join(REPO, 'tests', 'golden', 'models')
    model = Model.from_json(os.path.join(mdir, 'cantilever_beam_truss.json'))
    sc = StiffnessChecker(model, checker_engine='cuda', device=dev.index)
    sc.set_loads(LC.from_json(os.path.join(mdir, 'cantilever_beam_truss_loadcase.json')))
    sc.set_nodal_displacement_tol(trans_tol=1e-3)
    order = syn.bfs_sequence(model)
    for _ in range(2):
        flags = sc.solve_sequence(order)
    t0 = time.perf_counter()
    for _ in range(5):
        flags = sc.solve_sequence(order)
    dt = (time.perf_counter() - t0) / 5
    for _ in range(2):
        flags_ind = sc.solve_sequence(order, reuse=False)
    t0 = time.perf_counter()
    for _ in range(5):
        flags_ind = sc.solve_sequence(order, reuse=False)
    dt_ind = (time.perf_counter() - t0) / 5
    eng = sc._sc_ins
    o = np.asarray(order)
    words = np.zeros((len(o), eng.info['mask_words']), dtype=np.uint32)
    words[np.arange(len(o)), o >> 5] = np.uint32(1) << (o & 31).astype(np.uint32)
    dmask = torch.from_numpy(np.bitwise_or.accumulate(words, axis=0).view(np.int32)).to(dev)
    eng.enable_timing(True)
    for _ in range(2):
        rr = eng.solve_many_device(dmask, want=SCALARS + ('profile_flops',))
        torch.cuda.synchronize(dev)
    st = eng.last_stage_ms()
    eng.enable_timing(False)
    with_cpu('config2_cantilever_bfs_prefixes', {
        'value': len(order) / dt, 'unit': UNIT, 'ms_per_sequence': 1e3 * dt, 'prefixes': len(order), 'flags_true': int(sum(flags)),
        'ms_per_sequence_independent_prefixes': 1e3 * dt_ind, 'flags_equal': flags == flags_ind,
        'stage_ms_independent_prefixes': [float(x) for x in st], 'k3_tflops_algorithmic': float(rr['profile_flops'].sum().item()) / st[1] / 1e9,
        'what': 'StiffnessChecker.solve_sequence(order): all 2197 prefixes of the BFS construction sequence of cantilever_beam_truss '
                '(817 nodes, 4284 free DOFs, band of 14 tiles, 68 point loads) in one call, host wall clock, flags to the host; '
                'the factorisation is reused along the sequence (groups of 32 prefixes share the factor rows in front of the rows '
                'their added elements touch); *_independent_prefixes = the same prefixes as 2197 independent subsets'})
    del sc, eng, rr, dmask

    # config 4: joints on horizontals, four load cases per subset, height-field subsets
    model = syn.grid_frame(7, 7, 8, joints=True)
    sc = StiffnessChecker(model, checker_engine='cuda', device=dev.index)
    sc.set_load_cases(syn.config4_loadcases(model))
    eng = sc._sc_ins
    B4 = 4096
    masks = syn.pack_masks([syn.heightfield_subset(model, 7, 7, 8, seed=s) for s in range(B4)], len(model.elements))
    dm = torch.from_numpy(masks.view(np.int32)).to(dev)
    r = {}

    def run4():
        r['o'] = eng.solve_many_device(dm, want=FULL, out=r.get('o'))
    ms = timed(run4, 3)
    n_lc = r['o']['flag'].shape[1]
    with_cpu('config4_joints_4_loadcases', {'value': B4 * n_lc / ms * 1e3, 'unit': UNIT, 'subsets': B4, 'load_cases': n_lc,
                                             'ms_per_batch': ms, 'status_ok': int((r['o']['status'] == 0).sum().item()),
                                             'what': 'grid 7x7x8 + joint releases, 4 load cases per subset (one factorisation serves '
                                                     'all four), solve = (subset, load case), full results on device'})
    del sc, eng, r, dm
    torch.cuda.empty_cache()

    # config 5: 10k-element frame, 512 subsets, factorisation-bound
    model = syn.grid_frame(15, 15, 16)
    sc = StiffnessChecker(model, checker_engine='cuda', device=dev.index)
    sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
    eng = sc._sc_ins
    B5 = 512
    masks = syn.pack_masks([syn.connected_subset(model, s) for s in range(B5)], len(model.elements))
    dm = torch.from_numpy(masks.view(np.int32)).to(dev)
    r = {}

    def run5():
        r['o'] = eng.solve_many_device(dm, want=FULL + ('profile_flops',), out=r.get('o'), max_in_flight=B5)
    eng.enable_timing(True)
    run5()
    torch.cuda.synchronize(dev)
    run5()
    torch.cuda.synchronize(dev)
    st = eng.last_stage_ms()
    eng.enable_timing(False)
    ms = timed(run5, 2)
    fl = float(r['o']['profile_flops'].sum().item())
    with_cpu('config5_10k_frame', {'value': B5 / ms * 1e3, 'unit': UNIT, 'subsets': B5, 'ms_per_batch': ms,
                                    'k3_tflops_algorithmic': fl / st[1] / 1e9, 'stage_ms': [float(x) for x in st],
                                    'status_ok': int((r['o']['status'] == 0).sum().item()),
                                    'what': 'grid 15x15x16 (9675 elements / 20 250 free DOFs, band of 35 tiles), random connected subsets, '
                                            'full results on device'})
    del sc, eng, r, dm
    torch.cuda.empty_cache()

    # SURVEY 8f-4, many models of one topology: G shape variants of the config-3 frame, the element kernel over
    # all G x nE elements in one launch (HBM-bound), then one full-structure solve per variant
    model = syn.grid_frame(7, 7, 8)
    nE, nV = len(model.elements), len(model.nodes)
    base = np.array([n.point for n in model.nodes], dtype=float)
    G = 4096
    shapes = np.stack([syn.shape_variant(base, s) for s in range(G)])
    sc = StiffnessChecker(model, checker_engine='cuda', device=dev.index)
    eng = sc._sc_ins
    t0 = time.perf_counter()
    eng.set_geometries(shapes)                                  # first call: allocates the stacked tables (8.8 GB)
    t_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    eng.set_geometries(shapes)                                  # steady state of an optimisation loop: tables reused
    t_set = time.perf_counter() - t0
    k1_ms = eng.last_element_kernel_ms()
    sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
    masks = torch.from_numpy(syn.pack_masks([[]] * G, nE).view(np.int32)).to(dev)
    gidx = torch.arange(G, dtype=torch.int32, device=dev)
    r = {}

    def run_shapes():
        r['o'] = eng.solve_many_device(masks, want=FULL, out=r.get('o'), geometry=gidx)
    ms = timed(run_shapes, 3)
    # 152 B in, (144 + 9) doubles out; here + the member length and the second stiffness table
    # (entries |k| <= 1e-14 dropped, what the assembly reads)
    k1_bytes = G * nE * (152 + (2 * 144 + 9 + 1) * 8)
    hbm_peak, _ = measured_peaks()
    with_cpu('many_shapes_one_topology', {
        'value': G / (ms * 1e-3 + t_set), 'unit': 'shapes/s', 'shapes': G, 'solve_ms': ms, 'set_geometries_ms': 1e3 * t_set, 'first_set_geometries_ms': 1e3 * t_first,
        'k1_ms': k1_ms, 'k1_elements': G * nE, 'k1_gbs': k1_bytes / (k1_ms * 1e-3) / 1e9, 'k1_frac_of_hbm': k1_bytes / (k1_ms * 1e-3) / 1e9 / hbm_peak,
        'status_ok': int((r['o']['status'] == 0).sum().item()),
        'what': '4096 geometry variants (node coordinates jittered) of the 7x7x8 frame: cm_set_geometries (host upload + ONE element-kernel '
                'launch over 3.8 M elements) + one full-structure solve per variant, full results on device; value counts both'})
    del sc, eng, r
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))

    # ---- CPU baseline (rank 0, N=1 only) in a child process of its own, before the GPU work starts
    cpu_baseline = None
    cpu_others = None
    parity_file = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cmd = [sys.executable, '-m', 'oracle.cpu_bench', str(GRID[0]), str(GRID[1]), str(GRID[2]),
               str(TRANS_TOL), str(args.batch), str(args.cpu_sample)]
        if not args.no_other_configs:
            cmd.append('others')
        import tempfile
        parity_path = os.path.join(tempfile.gettempdir(), 'conmech_b200_parity_{}.npz'.format(os.getpid()))
        cmd.append('parity={},{}'.format(parity_path, min(args.batch, args.parity_sample)))
        r = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True)
        if r.returncode == 0 and r.stdout.strip():
            cpu_baseline = json.loads(r.stdout.strip().splitlines()[-1])
            cpu_others = cpu_baseline.pop('other_configs', None)
            parity_file = cpu_baseline.pop('parity_file', None)
        else:
            sys.stderr.write('cpu baseline failed: ' + r.stderr[-2000:] + '\n')

    import numpy as np
    import torch
    import torch.distributed as dist
    from conmech_b200 import StiffnessChecker, LoadCase, GravityLoad
    from conmech_b200 import synthetic as syn
    from conmech_b200 import masks as cm_masks

    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    numa_note = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        tns = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tns, op=dist.ReduceOp.MAX)
        return float(tns.item())

    model = syn.grid_frame(*GRID)
    nE, nV = len(model.elements), len(model.nodes)
    sc = StiffnessChecker(model, checker_engine='cuda', device=local)
    sc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
    sc.set_nodal_displacement_tol(trans_tol=TRANS_TOL)
    eng = sc._sc_ins
    info = eng.info

    # ---- synthetic input: this rank's contiguous range of subset seeds, a few distinct batches
    B = args.batch
    n_batches = max(1, min(args.distinct_batches, args.steps + args.warmup))
    host_masks = []
    for k in range(n_batches):
        base = (k * world + rank) * B
        m = syn.pack_masks([syn.connected_subset(model, base + s) for s in range(B)], nE)
        host_masks.append(torch.from_numpy(m.view(np.int32)).pin_memory())
    dev_masks = [m.to(dev) for m in host_masks]
    outs = {}

    def step_device(k, want):
        outs[want] = eng.solve_many_device(dev_masks[k % n_batches], want=want, out=outs.get(want))

    def timed_device(want):
        for k in range(args.warmup):
            step_device(k, want)
        barrier()
        eng.reset_counters()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for k in range(args.steps):
            step_device(args.warmup + k, want)
        ev1.record()
        torch.cuda.synchronize(dev)
        n_launch = eng.kernel_launches()
        ms = ev0.elapsed_time(ev1)
        barrier()
        return max_over_ranks(ms), n_launch

    # ---- FP64 peak (cuBLAS DGEMM) measured live: the roofline denominator for the factorisation
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    fp64_peak = 2 * 8192 ** 3 / best / 1e9
    del a, b
    torch.cuda.empty_cache()
    hbm_peak, hbm_source = measured_peaks()

    # ---- device-resident timed regions: full results (the headline), then the scalars-only variant
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    t0 = time.perf_counter()
    ms_max, launches = timed_device(FULL)
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    value = world * B * args.steps / (ms_max * 1e-3)
    flags_true = int(outs[FULL]['flag'].sum().item())
    status_ok = int((outs[FULL]['status'] == 0).sum().item())
    ms_flags, _ = timed_device(SCALARS)
    value_flags = world * B * args.steps / (ms_flags * 1e-3)

    # ---- end to end through the host-buffer C-ABI call (pinned host masks in, pinned host results out)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    out_host = {'flag': pin((B, 1), torch.uint8), 'status': pin((B, 1), torch.uint8),
                'max_trans': pin((B, 1), torch.float64), 'compliance': pin((B, 1), torch.float64),
                'U': pin((B, 1, 6 * nV), torch.float64), 'Rf': pin((B, 1, 6 * nV), torch.float64),
                'eR': pin((B, 1, nE, 12), torch.float64)}
    host_np = [m.numpy().view(np.uint32) for m in host_masks]

    def timed_host(fn):
        for k in range(max(1, args.warmup)):
            fn(k)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            fn(args.warmup + k)
        dt = time.perf_counter() - t0
        barrier()
        return world * B * args.steps / max_over_ranks(dt)

    # raw device->host rate of this box for pinned memory (one 1 GiB copy): what bounds the full-result leg
    probe_d = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    probe_h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
    probe_h.copy_(probe_d, non_blocking=True)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); probe_h.copy_(probe_d, non_blocking=True); e1.record(); torch.cuda.synchronize(dev)
    pcie_d2h_gbs = (1 << 30) / (e0.elapsed_time(e1) * 1e-3) / 1e9
    del probe_d, probe_h
    torch.cuda.empty_cache()
    e2e_dense = timed_host(lambda k: eng.solve_many(host_np[k % n_batches], want=FULL, out=out_host))
    d2h_dense = int(sum(out_host[n].nbytes for n in FULL))
    # the same results without zero padding - what the reference's get_solved_results() returns: displacements, the
    # reactions of the support nodes, the end forces of the EXISTING elements (rows in ascending element id)
    COMPACT = ('flag', 'status', 'max_trans', 'compliance', 'U', 'Rf_sup', 'eR_rows')
    outs_c = []
    for hm in host_np:
        oc = {n: out_host[n] for n in ('flag', 'status', 'max_trans', 'compliance', 'U')}
        oc['Rf_sup'] = pin((B, 1, len(model.supports), 6), torch.float64)
        oc['eR_rows'] = pin((int(eng.element_row_offsets(hm)[-1]), 12), torch.float64)
        outs_c.append(oc)
    e2e_full = timed_host(lambda k: eng.solve_many(host_np[k % n_batches], want=COMPACT, out=outs_c[k % n_batches]))
    d2h_full = int(sum(sum(oc[n].nbytes for n in COMPACT) for oc in outs_c) // len(outs_c))
    # compact rows == dense rows of the existing elements (same bits), support reactions == dense reactions there
    rc = eng.solve_many(host_np[0], want=COMPACT, out=outs_c[0])
    eng.solve_many(host_np[0], want=FULL, out=out_host)
    off = rc['eR_row_offsets']
    sup_ids = np.array(list(model.supports.keys()))
    assert np.array_equal(rc['Rf_sup'], out_host['Rf'].reshape(B, 1, -1, 6)[:, :, sup_ids])
    for b in range(0, B, 97):
        ids = cm_masks.mask_to_ids(host_np[0][b], nE)
        assert np.array_equal(rc['eR_rows'][off[b]:off[b + 1]], out_host['eR'][b, 0][ids]), b
    # the host path and the device path deliver the same bits
    eng.solve_many(host_np[0], want=FULL, out=out_host)
    step_device(0, FULL)
    torch.cuda.synchronize(dev)
    for n in FULL:
        assert np.array_equal(out_host[n], outs[FULL][n].cpu().numpy(), equal_nan=(n in ('max_trans',))), n
    # the same subsets the CPU arm solved (seeds 0.. of rank 0's first batch): flags equal, values within 1e-9
    parity = None
    if parity_file is not None:
        try:
            if 'error' in parity_file:
                raise RuntimeError(parity_file['error'])
            with np.load(parity_file['path']) as z:
                ref = {k: z[k] for k in z.files}
            n_par = len(ref['flag'])
            got = {n: outs[FULL][n][:n_par].cpu().numpy() for n in FULL}
            parity = parity_in_run(ref, got, host_np[0][:n_par], nE)
            os.remove(parity_file['path'])
        except Exception as ex:          # reported, never fatal for the line
            parity = {'error': repr(ex)}
    e2e_flags = timed_host(lambda k: eng.solve_many(host_np[k % n_batches], want=SCALARS, out=out_host))
    d2h_flags = int(sum(out_host[n].nbytes for n in SCALARS))
    # the facade's own calling convention: Python id lists in, list of bools out
    id_lists = [[cm_masks.mask_to_ids(row, nE).tolist() for row in hm] for hm in host_np[:1]]
    e2e_lists = timed_host(lambda k: sc.solve_many(id_lists[0]))
    h2d = int(host_np[0].nbytes)

    # ---- per-stage device times (CUDA events on the launching stream, inside the library) and the
    # algorithmic work of the same batches -> roofline per kernel
    eng.enable_timing(True)
    stage_ms = np.zeros(3)
    n_waves = 0
    for k in range(args.steps):
        eng.solve_many_device(dev_masks[(args.warmup + k) % n_batches], want=FULL, out=outs[FULL])
        torch.cuda.synchronize(dev)
        stage_ms += eng.last_stage_ms()
        n_waves += eng.last_wave_count()
    eng.enable_timing(False)
    flops = entries = tiles = nfree = elems = nfixed = 0.0
    for k in range(args.steps):              # algorithmic work of exactly the batches that were timed
        pf = eng.solve_many_device(dev_masks[(args.warmup + k) % n_batches],
                                   want=('profile_flops', 'profile_entries', 'envelope_tiles', 'n_free', 'n_fixed'))
        torch.cuda.synchronize(dev)
        flops += float(pf['profile_flops'].sum().item())
        entries += float(pf['profile_entries'].sum().item())
        tiles += float(pf['envelope_tiles'].sum().item())
        nfree += float(pf['n_free'].double().sum().item())
        hm = host_np[(args.warmup + k) % n_batches]
        elems += float(np.unpackbits(hm.view(np.uint8), axis=1).sum())
        nfixed += float(pf['n_fixed'].double().sum().item())
    n_sub = B * args.steps
    spl = n_sub / n_waves                                   # subsets per launch
    mask_bytes = info['mask_words'] * 4
    # algorithmic bytes per subset: S_K = 8 bytes x scalar-profile entries of K_mm (the smallest
    # of the listed formats; the tile format actually stored is given beside it)
    sk_profile = 8.0 * entries / n_sub
    sk_tiles = 8192.0 * tiles / n_sub
    k2_bytes = mask_bytes + sk_profile + 8.0 * nfree / n_sub
    k4_bytes = sk_profile + 8.0 * (6 * nV + nfixed / n_sub + 12.0 * elems / n_sub) + 18
    k2_ms, k3_ms, k4_ms = (stage_ms / n_waves)

    def hbm(name, bytes_per_subset, ms, extra):
        ach = bytes_per_subset * spl / (ms * 1e-3) / 1e9
        d = {'bound': 'hbm', 'achieved': ach, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach / hbm_peak,
             'traffic': NCU_DRAM_BYTES_PER_SUBSET[name] * spl, 'algorithmic_bytes_per_subset': bytes_per_subset,
             'avg_launch_ms': ms, 'subsets_per_launch': spl}
        d.update(extra)
        return d
    k3_ach = flops / n_waves / (k3_ms * 1e-3) / 1e12
    k3 = {'kernel': 'k3_factor_flow (block-band LDL^T, FP64 DMMA m8n8k4, dataflow rows)', 'bound': 'tensor',
          'achieved': k3_ach, 'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': k3_ach / fp64_peak,
          'traffic': NCU_DRAM_BYTES_PER_SUBSET['k3'] * spl,
          'traffic_source': 'ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch over 4096 subsets, scaled to this '
                            'launch size; ' + NCU_SOURCE,
          'peak_source': 'cuBLAS DGEMM 8192^3 fp64 measured live in this run (MEASURED_PEAKS.json has no '
                         'fp64 entry; earlier measurement on this pool {} TFLOP/s)'.format(FP64_FALLBACK_TFLOPS),
          'algorithmic_flops_per_launch': flops / n_waves,
          'flops_definition': 'sum_j h_j^2 over the active DOFs of each subset (envelope LDL^T, multiply-add = 2)',
          'algorithmic_bytes_per_subset': 2 * sk_profile,
          'avg_launch_ms': k3_ms, 'subsets_per_launch': spl}
    roofline = dict(k3)
    roofline['stage_ms_per_step'] = {'k2_dofmap_assemble': stage_ms[0] / args.steps, 'k3_factor': stage_ms[1] / args.steps,
                                     'k4_solve_recover': stage_ms[2] / args.steps}
    roofline['kernels'] = {
        'k2_dofmap_assemble': hbm('k2', k2_bytes, k2_ms, {
            'bytes_definition': 'mask + S_K + 8 n_s per subset, S_K = 8 x scalar-profile entries of K_mm (SURVEY 8d)',
            'achieved_tile_format': (mask_bytes + sk_tiles + 8.0 * nfree / n_sub) * spl / (k2_ms * 1e-3) / 1e9,
            'includes': 'k2_schedule (largest-first rank of the wave, ~60 us)'}),
        'k3_factor_flow': k3,
        'k4_solve_recover': hbm('k4', k4_bytes, k4_ms, {
            'bytes_definition': 'S_K (forward sweep fused into K3) + 8 (6nV + n_fixed + 12 #S) + 18 per solve, full results stored',
            'achieved_tile_format': (sk_tiles + k4_bytes - sk_profile) * spl / (k4_ms * 1e-3) / 1e9}),
    }
    roofline['peak_hbm_source'] = hbm_source
    roofline['s_k_bytes_per_subset'] = {'profile': sk_profile, 'tile_format': sk_tiles}

    # ---- the other BASELINE configurations (one GPU)
    others = None
    outs.clear()
    out_host = None
    if rank == 0 and world == 1 and not args.no_other_configs:
        eng.close()                      # frees the handle's workspace and staging: config 5 needs the memory
        torch.cuda.empty_cache()
        try:
            others = other_configs_gpu(torch, np, dev, cpu_others)
        except Exception as ex:          # never lose the headline line to a secondary leg
            others = {'error': repr(ex)}

    # ---- one process, all GPUs (N > 1): rank 0 drives every device through cm_solve_many_host_multi.  The other
    # ranks leave first (a rank waiting in an NCCL barrier keeps a spinning kernel on its GPU, which is what this leg
    # would then measure): the process group is torn down on every rank, ranks > 0 exit, rank 0 goes on alone.
    single = None
    if world > 1:
        eng.close()
        torch.cuda.empty_cache()
        barrier()
        dist.destroy_process_group()
        if rank != 0:
            return
        time.sleep(3.0)                     # the other ranks' contexts are going away
        try:
            scm = StiffnessChecker(model, checker_engine='cuda', devices=list(range(world)))
            scm.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
            scm.set_nodal_displacement_tol(trans_tol=TRANS_TOL)
            em = scm._sc_ins
            BG = B * world
            mm = syn.pack_masks([syn.connected_subset(model, s) for s in range(BG)], nE)
            mm = torch.from_numpy(mm.view(np.int32)).pin_memory().numpy().view(np.uint32)
            oh = {'flag': pin((BG, 1), torch.uint8), 'status': pin((BG, 1), torch.uint8),
                  'max_trans': pin((BG, 1), torch.float64), 'compliance': pin((BG, 1), torch.float64)}
            for _ in range(2):
                em.solve_many(mm, want=SCALARS, out=oh)
            t0 = time.perf_counter()
            reps = max(2, args.steps // 2)
            for _ in range(reps):
                em.solve_many(mm, want=SCALARS, out=oh)
            rate = BG * reps / (time.perf_counter() - t0)
            single = {'n_gpus': world, 'unit': UNIT, 'subsets_per_call': BG, 'value': rate,
                      'outputs': 'flag,status,max_trans,compliance',
                      'what': 'ONE process, one handle and one host thread per GPU (cm_solve_many_host_multi): pinned host '
                              'masks in, results written by every device straight into the caller\'s arrays (the host '
                              'gather) inside the timed region; host wall clock; run after the SPMD ranks have exited'}
            em.close()
        except Exception as ex:
            single = {'error': repr(ex)}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_max / args.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': workload_config(args),                         # identical to the reference arm's `config`
                'workload_stats': {'mean_free_dofs_per_subset': nfree / n_sub,
                                   'half_bandwidth_full': info['half_bandwidth'],
                                   'flags_true_last_step_rank0': flags_true,
                                   'status_ok_last_step_rank0': status_ok},
                'value_flags_only': value_flags, 'host_placement': numa_note,
                'clocks': clocks,
                'e2e': {'value': e2e_full, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h_full,
                        'outputs': 'per subset, to pinned host arrays: flag, status, max_trans, compliance, U (6 nV), the reactions of the '
                                   'support nodes (nS x 6) and the end forces of the EXISTING elements (12 per element) - the content of '
                                   'the reference\'s get_solved_results(), no zero padding; checked against the dense arrays in this run',
                        'd2h_gbs': d2h_full * e2e_full / (world * B) / 1e9,
                        'pcie_d2h_gbs_measured': pcie_d2h_gbs,
                        'fraction_of_device_resident': e2e_full / value,
                        'dense_value': e2e_dense, 'dense_d2h_bytes_per_step': d2h_dense,
                        'dense_what': 'the same with zero-padded arrays: Rf over all 6 nV DOFs, eR rows for all nE elements (127 KB per solve)',
                        'note': 'every wave\'s results are copied to the host by the copy stream of its lane while later waves compute; the '
                                'waves shrink towards the end of the batch so that little is left to copy when the last kernel ends',
                        'flags_only_value': e2e_flags, 'flags_only_d2h_bytes_per_step': d2h_flags,
                        'id_lists_value': e2e_lists,
                        'id_lists_what': 'StiffnessChecker.solve_many(list of Python id lists) -> list of bools, packing included'},
                'gpu_launches': launches, 'roofline': roofline}
        if cpu_baseline is not None:
            line['cpu_baseline'] = cpu_baseline
        if parity is not None:
            line['parity_in_run'] = parity
        if others is not None:
            line['other_configs'] = others
        if single is not None:
            line['single_process'] = single
        print(json.dumps(line))


if __name__ == '__main__':
    main()
