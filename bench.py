#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's headline metric: partial-structure solves/sec on
the ~1k-element space frame (config 3: grid 7x7x8, 931 elements, 2058 free DOFs, random connected
element subsets of 20-90 %, gravity load, trans_tol 1e-3), subsets range-partitioned over GPUs.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host CPU

One "step" = one pass of the hot path (assembly -> factorisation -> recovery) over one batch of
B subsets per GPU (default 12 500, i.e. 100 000 across 8 GPUs).  `value` is device-resident
throughput (masks already in HBM, CUDA events on the launching stream, max over ranks);
`e2e` goes through the host-buffer C-ABI call with pinned host masks in and flags / status /
max_trans / compliance out, host<->device copies inside the timed region.
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
# DRAM bytes (read + write) of ONE k3_factor_flow launch over 4096 config-3 subsets, from the ncu --set full
# capture summarised in profiles/r01c_ncu_k2_k3_k4.txt (24.30 GB read + 9.52 GB written)
K3_DRAM_BYTES_PER_SUBSET = (24.302537e9 + 9.517715e9) / 4096


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--batch', type=int, default=12500, help='subsets per GPU per step')
    ap.add_argument('--cpu-sample', type=int, default=0, help='subsets in the CPU baseline sample (0 = auto)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--distinct-batches', type=int, default=2)
    return ap.parse_args()


def workload_config(args, extra=None):
    cfg = {'workload': 'config3: synthetic space-frame grid 7x7x8 (392 nodes / 931 elements / 2058 free DOFs), '
                       'gravity [0,0,-1], random connected element subsets U(0.2,0.9)*nE (seeded), trans_tol 1e-3',
           'subsets_per_gpu_per_step': args.batch,
           'global_subsets_per_step': args.batch * args.gpus,
           'load_cases': 1,
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
    """The reference algorithm on the host CPU (oracle port; the Python reference cannot travel to
    the GPU box).  Rank 0 only; every step solves a bounded sample of the same workload."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    from oracle import cpu_bench
    n_procs = os.cpu_count() or 1
    per_step = args.cpu_sample or 100 * n_procs                          # ~3 s per step
    import multiprocessing as mp
    ctx = mp.get_context('spawn')
    seeds_all = list(range((args.warmup + args.steps) * per_step))
    with ctx.Pool(n_procs, initializer=cpu_bench._init_worker, initargs=(GRID, TRANS_TOL)) as pool:
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
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': workload_config(args, {'reference_sample_per_step': per_step}),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': n_procs, 'kind': 'port',
                             'sample': '{} subsets per step (seeds 0..), {} processes, {}'.format(
                                 per_step, n_procs, cpu_bench.cpu_model_name())},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == 'reference':
        return run_reference(args)

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))

    # ---- CPU baseline (rank 0, N=1 only) in a child process of its own, before the GPU work starts
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = subprocess.run([sys.executable, '-m', 'oracle.cpu_bench', str(GRID[0]), str(GRID[1]), str(GRID[2]),
                            str(TRANS_TOL), str(args.batch), str(args.cpu_sample)],
                           cwd=REPO, capture_output=True, text=True)
        if r.returncode == 0 and r.stdout.strip():
            cpu_baseline = json.loads(r.stdout.strip().splitlines()[-1])
        else:
            sys.stderr.write('cpu baseline failed: ' + r.stderr[-2000:] + '\n')

    import numpy as np
    import torch
    import torch.distributed as dist
    from conmech_b200 import StiffnessChecker, LoadCase, GravityLoad
    from conmech_b200 import synthetic as syn

    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    model = syn.grid_frame(*GRID)
    nE = len(model.elements)
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
    want = ('flag', 'status', 'max_trans', 'compliance')
    out_dev = None

    def step_device(k):
        nonlocal out_dev
        out_dev = eng.solve_many_device(dev_masks[k % n_batches], want=want, out=out_dev)

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

    # ---- device-resident timed region
    for k in range(args.warmup):
        step_device(k)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    eng.reset_counters()
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for k in range(args.steps):
        step_device(args.warmup + k)
    ev1.record()
    torch.cuda.synchronize(dev)
    launches = eng.kernel_launches()
    ms = ev0.elapsed_time(ev1)
    barrier()
    t1 = time.perf_counter()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world * B * args.steps / (ms_max * 1e-3)
    flags_true = int(out_dev['flag'].sum().item())
    status_ok = int((out_dev['status'] == 0).sum().item())

    # ---- end to end through the host-buffer C-ABI call (pinned host masks in, host results out)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    out_host = {'flag': pin((B, 1), torch.uint8), 'status': pin((B, 1), torch.uint8),
                'max_trans': pin((B, 1), torch.float64), 'compliance': pin((B, 1), torch.float64)}
    host_np = [m.numpy().view(np.uint32) for m in host_masks]
    for k in range(max(1, args.warmup)):
        eng.solve_many(host_np[k % n_batches], want=want, out=out_host)
    barrier()
    t0 = time.perf_counter()
    step_times = []
    for k in range(args.steps):
        ts = time.perf_counter()
        eng.solve_many(host_np[(args.warmup + k) % n_batches], want=want, out=out_host)
        step_times.append(time.perf_counter() - ts)
    e2e_s = time.perf_counter() - t0
    if os.environ.get('CM_BENCH_DEBUG'):
        sys.stderr.write('rank {} e2e step ms {}\n'.format(rank, ['{:.1f}'.format(1e3 * x) for x in step_times]))
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(te.item())
    h2d = int(host_np[0].nbytes)
    d2h = int(sum(v.nbytes for v in out_host.values()))
    assert np.array_equal(out_host['flag'], out_dev['flag'].cpu().numpy()) or n_batches > 1

    # ---- per-stage device times (CUDA events on the launching stream, inside the library) and
    # the algorithmic work of the same batch -> roofline of the dominant kernel (factorisation)
    eng.enable_timing(True)
    stage_ms = np.zeros(3)
    n_waves = 0
    for k in range(args.steps):
        eng.solve_many_device(dev_masks[(args.warmup + k) % n_batches], want=want, out=out_dev)
        torch.cuda.synchronize(dev)
        stage_ms += eng.last_stage_ms()
        n_waves += eng.last_wave_count()
    eng.enable_timing(False)
    flops_batches = 0.0
    nfree_mean = 0.0
    for k in range(args.steps):              # algorithmic work of exactly the batches that were timed
        pf = eng.solve_many_device(dev_masks[(args.warmup + k) % n_batches], want=('profile_flops', 'n_free'))
        torch.cuda.synchronize(dev)
        flops_batches += float(pf['profile_flops'].sum().item())
        nfree_mean += float(pf['n_free'].double().mean().item()) / args.steps
    # achieved = algorithmic flops of the timed batches / device time of their K3 launches
    k3_ms_per_launch = stage_ms[1] / n_waves
    flops_per_launch = flops_batches / n_waves
    subsets_per_launch = B * args.steps / n_waves
    achieved = flops_per_launch / (k3_ms_per_launch * 1e-3) / 1e12
    roofline = {'kernel': 'k3_factor_flow (block-band LDL^T, FP64 DMMA m8n8k4, dataflow rows)', 'bound': 'tensor',
                'achieved': achieved, 'peak': fp64_peak, 'unit': 'TFLOP/s', 'frac': achieved / fp64_peak,
                'traffic': K3_DRAM_BYTES_PER_SUBSET * subsets_per_launch,
                'traffic_source': 'ncu dram__bytes_read.sum + dram__bytes_write.sum of one k3_factor_flow launch '
                                  '(4096 subsets), scaled to this launch size; profiles/r01c_ncu_k2_k3_k4.txt',
                'peak_source': 'cuBLAS DGEMM 8192^3 fp64 measured live in this run (MEASURED_PEAKS.json has no '
                               'fp64 entry; earlier measurement on this pool {} TFLOP/s)'.format(FP64_FALLBACK_TFLOPS),
                'algorithmic_flops_per_launch': flops_per_launch,
                'flops_definition': 'sum_j h_j^2 over the active DOFs of each subset (envelope LDL^T, multiply-add = 2)',
                'avg_launch_ms': k3_ms_per_launch, 'subsets_per_launch': subsets_per_launch,
                'stage_ms_per_step': {'k2_dofmap_assemble': stage_ms[0] / args.steps,
                                      'k3_factor': stage_ms[1] / args.steps,
                                      'k4_solve_recover': stage_ms[2] / args.steps}}

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_max / args.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
                'config': workload_config(args, {'mean_free_dofs_per_subset': nfree_mean,
                                                 'half_bandwidth_full': info['half_bandwidth'],
                                                 'flags_true_last_step_rank0': flags_true,
                                                 'status_ok_last_step_rank0': status_ok}),
                'clocks': clocks,
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                        'outputs': 'flag,status,max_trans,compliance per subset'},
                'gpu_launches': launches, 'roofline': roofline}
        if cpu_baseline is not None:
            line['cpu_baseline'] = cpu_baseline
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
