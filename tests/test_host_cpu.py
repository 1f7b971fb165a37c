"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol the header
declares, host logic (masks, model IO, synthetic workloads, subset sharding)."""
import json
import os
import re

import numpy as np
import pytest
from numpy.testing import assert_array_equal

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_capi_loads_and_exports_header_symbols():
    import __graft_entry__ as ge
    ge.build()
    from conmech_b200 import _capi
    lib = _capi.load()
    with open(os.path.join(REPO, 'include', 'conmech_b200.h')) as f:
        header = f.read()
    declared = set(re.findall(r'\b(cm_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found'
    assert declared == set(_capi.SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert b'sm_100a' in lib.cm_version()
    # struct layouts the header promises
    import ctypes as C
    assert C.sizeof(_capi.SolveOut) == 21 * 8 and C.sizeof(_capi.SolveOutHost) == 21 * 8
    assert C.sizeof(_capi.ModelDesc) == 16 + 7 * 8


def test_engine_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from conmech_b200 import StiffnessChecker, Model
    from conmech_b200._capi import CmError
    m = Model.from_json(os.path.join(REPO, 'tests', 'golden', 'models', 'tower_3D.json'))
    with pytest.raises(CmError):
        StiffnessChecker(m, checker_engine='cuda')


def test_no_oracle_import_in_product():
    pkg = os.path.join(REPO, 'conmech_b200')
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(root, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f


def test_masks_roundtrip():
    from conmech_b200 import pack_masks, ids_to_mask, mask_to_ids
    rng = np.random.RandomState(0)
    for nE in (1, 31, 32, 33, 931):
        lists = [sorted(rng.choice(nE, size=rng.randint(1, nE + 1), replace=False).tolist()) for _ in range(7)]
        m = pack_masks(lists, nE)
        assert m.shape == (7, (nE + 31) // 32) and m.dtype == np.uint32
        for b, ids in enumerate(lists):
            assert_array_equal(mask_to_ids(m[b], nE), ids)
            assert_array_equal(ids_to_mask(ids, nE), m[b])
    full = pack_masks([[]], 40)
    assert_array_equal(mask_to_ids(full[0], 40), np.arange(40))
    with pytest.raises(IndexError):
        pack_masks([[40]], 40)


def test_native_mask_packer_matches_numpy():
    """The CPython packer (csrc_host/cmpack.c) and the numpy path produce the same masks for lists,
    tuples, sets, ranges and integer arrays; empty = full structure; out-of-range ids raise."""
    import random
    import numpy as np
    from conmech_b200 import masks
    assert masks._native_packer() is not None, 'run python -m conmech_b200.build'
    nE = 931
    rng = random.Random(1)
    subs = [rng.sample(range(nE), rng.randint(1, 900)) for _ in range(200)]
    subs[3] = []
    subs[4] = tuple(subs[4])
    subs[5] = set(subs[5])
    subs[6] = range(10, 500, 3)
    subs[7] = np.array(subs[7], dtype=np.int32)
    subs[8] = np.array(subs[8], dtype=np.int64)
    subs[9] = np.array(subs[9], dtype=np.uint16)
    subs[10] = [nE - 1, 0, 31, 32]
    native = masks.pack_masks(subs, nE)
    masks.USE_NATIVE = False
    try:
        ref = masks.pack_masks([list(x) for x in subs], nE)
    finally:
        masks.USE_NATIVE = True
    assert native.dtype == np.uint32 and native.shape == (200, 30)
    assert (native == ref).all()
    assert set(masks.mask_to_ids(native[10], nE)) == {0, 31, 32, nE - 1}
    assert len(masks.mask_to_ids(native[3], nE)) == nE
    for bad in ([[0, nE]], [[-1]], [np.array([5, 5000])]):
        with pytest.raises(IndexError):
            masks.pack_masks(bad, nE)
    # a repeated id is rejected by every packer (the reference would assemble the element twice,
    # numpy_stiffness_assembly.py:371-388; a bit mask cannot express that)
    for dup in ([[3, 7, 3]], [np.array([1, 1], dtype=np.int64)], [(5, 6, 5)]):
        with pytest.raises(ValueError):
            masks.pack_masks(dup, nE)
        masks.USE_NATIVE = False
        try:
            with pytest.raises(ValueError):
                masks.pack_masks([list(dup[0])], nE)
        finally:
            masks.USE_NATIVE = True
    with pytest.raises(ValueError):
        masks.pack_masks_csr(np.array([4, 9, 4]), np.array([0, 3]), nE)
    # every kind of id object takes the same decisions on the fast path (exact one-digit ints read from the object)
    # and on the number-protocol path (numpy scalars, bools, big ints)
    mixed = [[1, 2, 3], [np.int64(5), np.int32(7), True], (nE - 1, 0), range(3), {4, 5}, np.array([1, 2], dtype=np.int16)]
    got = [masks.mask_to_ids(r, nE).tolist() for r in masks.pack_masks(mixed, nE)]
    assert got == [[1, 2, 3], [1, 5, 7], [0, nE - 1], [0, 1, 2], [4, 5], [1, 2]]
    for bad, exc in (([[-1]], IndexError), ([[2 ** 40]], IndexError), ([[2 ** 80]], IndexError), ([[1.5]], TypeError),
                     ([['a']], TypeError), ([[None]], TypeError)):
        with pytest.raises(exc):
            masks.pack_masks(bad, nE)
    # release-store hand-over flag and the mask-matrix detection of solve_many
    fl = np.zeros(4, dtype=np.int32)
    masks.publish(fl, 2, 1)
    masks.publish(fl, 3, -1)
    assert fl.tolist() == [0, 0, 1, -1]
    mm = masks.pack_masks([[1, 2], [3]], nE)
    assert masks.as_mask_matrix(mm, 30) is not None and masks.as_mask_matrix(mm.view(np.int32), 30).dtype == np.uint32
    assert masks.as_mask_matrix([[1, 2], [3]], 30) is None
    assert masks.as_mask_matrix(np.zeros((2, 5), dtype=np.int64), 30) is None       # id rows
    with pytest.raises(ValueError):
        masks.as_mask_matrix(np.zeros((2, 5), dtype=np.int32), 30)                  # masks of another model?
    # the same subsets as two arrays (flat ids + offsets)
    lists = [list(x) for x in subs]
    flat = np.concatenate([np.asarray(x, dtype=np.int64) for x in lists])
    off = np.concatenate([[0], np.cumsum([len(x) for x in lists])])
    assert (masks.pack_masks_csr(flat, off, nE) == ref).all()
    with pytest.raises(IndexError):
        masks.pack_masks_csr(np.array([0, nE]), np.array([0, 2]), nE)
    with pytest.raises(ValueError):
        masks.pack_masks_csr(np.array([1, 2]), np.array([0, 3]), nE)


def test_pack_elements_matches_the_reference_rules():
    """cm_pack_elements (host-only C function of the library): per-element property / spring arrays from
    per-tag tables, against the per-element rules of numpy_stiffness.py:579-614 written out in Python."""
    import ctypes as C
    from conmech_b200 import _capi, synthetic as syn
    lib = _capi.load()
    model = syn.grid_frame(4, 4, 4, joints=True)
    model.elements[5].bending_stiff = False
    nE = len(model.elements)
    tags = sorted(set(e.elem_tag for e in model.elements))
    jt = {j_tag: k for k, j in enumerate(model.joints) for j_tag in j.elem_tags}
    sections = np.array([[cs.A, cs.Jx, cs.Iy, cs.Iz] for cs in model.crosssecs])
    materials = np.array([[m.E, m.mu, m.density] for m in model.materials])
    joints = np.array([[np.inf if x is None else x for x in list(j.c_conditions[3:6]) + list(j.c_conditions[9:12])]
                       for j in model.joints])
    sec_of = np.zeros(nE, np.int32)
    mat_of = np.zeros(nE, np.int32)
    joint_of = np.array([jt.get(e.elem_tag, -1) for e in model.elements], dtype=np.int32)
    stiff = np.array([1 if e.bending_stiff else 0 for e in model.elements], dtype=np.uint8)
    props, cr = np.zeros((nE, 7)), np.zeros((nE, 6))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.cm_pack_elements(nE, p(sec_of), p(mat_of), p(joint_of), p(stiff), len(sections), p(sections), len(materials),
                              p(materials), len(joints), p(joints), p(props), p(cr))
    assert rc == 0, lib.cm_last_error()
    assert len(tags) == 4
    for i, e in enumerate(model.elements):
        assert_array_equal(props[i], list(sections[0]) + list(materials[0]))
        want = np.full(6, np.inf)
        if joint_of[i] >= 0:
            want = joints[joint_of[i]]
        if not e.bending_stiff:
            want = np.zeros(6)
        assert_array_equal(cr[i], want)
    sec_of[3] = 7                                   # out-of-range table row
    assert lib.cm_pack_elements(nE, p(sec_of), p(mat_of), p(joint_of), p(stiff), len(sections), p(sections), len(materials),
                                p(materials), len(joints), p(joints), p(props), p(cr)) == -1
    assert b'does not exist' in lib.cm_last_error()


def test_model_io_roundtrip(models_dir):
    from conmech_b200 import Model, LoadCase
    for f in sorted(os.listdir(models_dir)):
        if f.endswith('_loads.json') or f.endswith('_loadcase.json'):
            lc = LoadCase.from_json(os.path.join(models_dir, f))
            lc2 = LoadCase.from_data(json.loads(json.dumps(lc.to_data())))
            assert len(lc2.point_loads) == len(lc.point_loads)
            continue
        m = Model.from_json(os.path.join(models_dir, f))
        m2 = Model.from_data(json.loads(json.dumps(m.to_data())))
        assert [n.point for n in m.nodes] == [n.point for n in m2.nodes]
        assert [e.end_node_inds for e in m.elements] == [e.end_node_inds for e in m2.elements]
        assert list(m.supports.keys()) == list(m2.supports.keys())
        assert m.crosssec_id_from_etag == m2.crosssec_id_from_etag
    mm = Model.from_json(os.path.join(models_dir, 'tower_3D.json'))
    data = mm.to_data()
    data['unit'] = 'millimeter'
    scaled = Model.from_data(data)
    assert np.allclose(np.array([n.point for n in scaled.nodes]) * 1e3, [n.point for n in mm.nodes])


def test_synthetic_workloads_match_survey():
    from conmech_b200 import synthetic as syn, Model
    m = syn.grid_frame(7, 7, 8)
    assert (len(m.nodes), len(m.elements), len(m.supports)) == (392, 931, 49)
    assert len(syn.grid_frame(15, 15, 16).elements) == 9675
    s0 = syn.connected_subset(m, 0)
    assert s0 == syn.connected_subset(m, 0) and s0 == sorted(set(s0))
    assert 0.19 * 931 <= len(s0) <= 0.91 * 931
    fixed = set(m.supports)
    seen, todo = set(), [e for e in s0 if set(m.elements[e].end_node_inds) & fixed]
    inc = syn.node_incidence(m)
    chosen = set(s0)
    while todo:                               # connected to the ground by construction
        e = todo.pop()
        if e in seen:
            continue
        seen.add(e)
        for v in m.elements[e].end_node_inds:
            todo.extend(f for f in inc[v] if f in chosen and f not in seen)
    assert seen == chosen
    tower = Model.from_json(os.path.join(REPO, 'tests', 'golden', 'models', 'tower_3D.json'))
    assert syn.bfs_sequence(tower) == [0, 1, 2, 3, 8, 9, 10, 11, 12, 13, 14, 15, 4, 7, 17, 22, 5, 19, 20, 6, 18, 21, 16, 23]
    mj = syn.grid_frame(5, 5, 5, joints=True)
    tags = {e.elem_tag for e in mj.elements}
    assert tags == {'', 'hinge_a', 'spring', 'both'}
    assert len(syn.config4_loadcases(mj)) == 4
    hf = syn.heightfield_subset(mj, 5, 5, 5, seed=3)
    assert hf == sorted(hf) and len(hf) > 0


# ---------------------------------------------------------------- multi-rank sharding (gloo, CPU)
def test_shard_range_partitions():
    from conmech_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 100000, 12345):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


class _FakeEngine(object):
    """Stands in for CudaStiffness on a CPU box: deterministic per-subset outputs from the masks."""
    info = {'mask_words': 30}

    def solve_many(self, masks, want):
        pop = np.unpackbits(np.ascontiguousarray(masks).view(np.uint8), axis=1).sum(axis=1)
        res = {'flag': (pop % 2 == 0).astype(np.uint8).reshape(-1, 1), 'status': np.zeros((len(masks), 1), np.uint8),
               'max_trans': pop.astype(np.float64).reshape(-1, 1) * 1e-4,
               'compliance': pop.astype(np.float64).reshape(-1, 1) * 0.5}
        if 'eR_rows' in want:           # compact element rows: one row per existing element, value = 1000 * subset hash + element id
            m = np.ascontiguousarray(masks)
            bits = np.unpackbits(m.view(np.uint8), axis=1, bitorder='little')
            rows = [np.outer(1000.0 * float(m[b].sum() % 997) + np.flatnonzero(bits[b]), np.ones(12)) for b in range(len(m))]
            res['eR_rows'] = np.concatenate(rows, axis=0) if rows else np.zeros((0, 12))
        return {k: res[k] for k in want}


def _sharded_worker(rank, world, port, B, ret):
    import torch.distributed as dist
    from conmech_b200.sharding import solve_many_sharded
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:{}'.format(port), rank=rank, world_size=world)
    rng = np.random.RandomState(0)
    masks = rng.randint(0, 2 ** 32, size=(B, 30), dtype=np.uint64).astype(np.uint32)
    out = solve_many_sharded(_FakeEngine(), masks.view(np.int32) if rank else masks)      # int32 views are masks too
    out.update(solve_many_sharded(_FakeEngine(), masks, want=('eR_rows',)))
    ret[rank] = {k: v.copy() for k, v in out.items()}
    dist.destroy_process_group()


def test_solve_many_sharded_gloo_world2():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    B, world = 37, 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_sharded_worker, args=(world, port, B, ret), nprocs=world, join=True)
    rng = np.random.RandomState(0)
    masks = rng.randint(0, 2 ** 32, size=(B, 30), dtype=np.uint64).astype(np.uint32)
    from conmech_b200.cuda_stiffness import CudaStiffness
    ref = _FakeEngine().solve_many(masks, ('flag', 'status', 'max_trans', 'compliance', 'eR_rows'))
    ref['eR_row_offsets'] = CudaStiffness.element_row_offsets(masks)
    assert ref['eR_rows'].shape[0] == ref['eR_row_offsets'][-1]
    for r in range(world):
        for k, v in ref.items():
            assert np.array_equal(ret[r][k], v), (r, k)


def test_fragment_maps_of_the_factorisation_kernel():
    """tools/frag_emulator.py: lane-level emulation of the FP64 MMA fragment maps that
    k3_factor_flow.cu relies on (in-register solve product, SYRK from registers, M_J layout)."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools', 'frag_emulator.py')
    spec = importlib.util.spec_from_file_location('frag_emulator', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main()


def test_element_row_offsets_host_helper():
    """cm_element_row_offsets (host-only C function): exclusive prefix sum of the masks' popcounts, odd and even
    word counts, empty batch."""
    from conmech_b200.cuda_stiffness import CudaStiffness
    rng = np.random.default_rng(3)
    for words in (1, 2, 7, 30):
        m = rng.integers(0, 2 ** 32, size=(257, words), dtype=np.uint64).astype(np.uint32)
        off = CudaStiffness.element_row_offsets(m)
        ref = np.concatenate([[0], np.cumsum(np.bitwise_count(m).sum(axis=1, dtype=np.int64))])
        assert off.dtype == np.int64 and np.array_equal(off, ref)
    assert np.array_equal(CudaStiffness.element_row_offsets(np.zeros((0, 4), dtype=np.uint32)), [0])


def _build_c_client(tmp_path):
    import subprocess
    exe = str(tmp_path / 'cantilever')
    subprocess.check_call(['gcc', '-O2', '-Wall', '-Werror', '-I', os.path.join(REPO, 'include'),
                           os.path.join(REPO, 'tests', 'c_client', 'cantilever.c'),
                           '-L', os.path.join(REPO, 'conmech_b200', 'lib'), '-lconmech_b200', '-lm', '-o', exe])
    return exe


def test_c_client_compiles_against_header_and_fails_loudly_without_gpu(tmp_path):
    """The drop-in boundary is a C ABI: a plain C program (tests/c_client/cantilever.c) compiles against
    include/conmech_b200.h with -Wall -Werror and links the library.  Without a GPU cm_create reports the CUDA
    error and the program exits 2 - nothing falls back to the CPU."""
    import subprocess
    import torch
    from conmech_b200 import build as b
    b.build()
    exe = _build_c_client(tmp_path)
    if torch.cuda.is_available():
        pytest.skip('GPU present: the GPU suite runs the client')
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(REPO, 'conmech_b200', 'lib'))
    r = subprocess.run([exe], capture_output=True, text=True, env=env)
    assert r.returncode == 2 and 'cm_create' in r.stderr


def test_cm_create_validates_the_description_before_any_cuda_call():
    """Programmer errors in a model description are reported as CM_ERR_INVALID with the reference's messages
    (numpy_stiffness_assembly.py:288-289 bar length, :146 E / Iz, stiffness_base.py:25-26 supports) - checked on the
    host before the library touches CUDA, so the same answers come back with and without a GPU."""
    import ctypes as C
    from conmech_b200 import _capi
    lib = _capi.load()

    def create(xyz, conn, props=None, cr=None, sup_node=(0,), sup_cond=None, node_order=None, null=None):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        nE = len(conn)
        props = np.ascontiguousarray(np.tile([1e-3, 1e-6, 1e-6, 1e-6, 2.1e8, 0.3, 78.5], (nE, 1)) if props is None else props)
        cr = np.full((nE, 6), np.inf) if cr is None else np.ascontiguousarray(cr, dtype=np.float64)
        sup_node = np.ascontiguousarray(sup_node, dtype=np.int32)
        sup_cond = np.ones((len(sup_node), 6), dtype=np.uint8) if sup_cond is None else sup_cond
        d = _capi.ModelDesc(len(xyz), nE, len(sup_node), 0,
                            xyz.ctypes.data_as(_capi.c_double_p), conn.ctypes.data_as(_capi.c_int32_p),
                            props.ctypes.data_as(_capi.c_double_p), cr.ctypes.data_as(_capi.c_double_p),
                            sup_node.ctypes.data_as(_capi.c_int32_p), sup_cond.ctypes.data_as(_capi.c_uint8_p), None)
        if node_order is not None:
            no = np.ascontiguousarray(node_order, dtype=np.int32)
            d.node_order = no.ctypes.data_as(_capi.c_int32_p)
        if null:
            setattr(d, null, None)
        h = C.c_void_p()
        rc = lib.cm_create(C.byref(d), C.byref(h))
        msg = lib.cm_last_error().decode()
        if rc == 0:
            lib.cm_destroy(h)
        return rc, msg

    xyz = [[0, 0, 0], [1, 0, 0], [1, 1, 0]]
    conn = [[0, 1], [1, 2]]
    INVALID = -1
    assert lib.cm_create(None, None) == INVALID
    for field in ('xyz', 'conn', 'props', 'cr', 'sup_node', 'sup_cond'):
        rc, msg = create(xyz, conn, null=field)
        assert rc == INVALID and 'null table' in msg, (field, rc, msg)
    rc, msg = create(xyz, conn, sup_node=())
    assert rc == INVALID and 'at least one support' in msg
    rc, msg = create(xyz, conn, sup_node=(3,))
    assert rc == INVALID and 'support node out of range' in msg
    rc, msg = create(xyz, conn, sup_node=(0, 0))
    assert rc == INVALID and 'duplicate support node' in msg
    rc, msg = create(xyz, [[0, 1], [1, 3]])
    assert rc == INVALID and 'element 1 has invalid end nodes' in msg
    rc, msg = create(xyz, [[0, 1], [2, 2]])
    assert rc == INVALID and 'invalid end nodes' in msg
    rc, msg = create([[0, 0, 0], [1, 0, 0], [1, 0, 1e-15]], conn)
    assert rc == INVALID and msg == 'E#1 - Bar length too small!'
    p = np.tile([1e-3, 1e-6, 1e-6, 1e-6, 2.1e8, 0.3, 78.5], (2, 1))
    p[1, 4] = 0.0
    rc, msg = create(xyz, conn, props=p)
    assert rc == INVALID and msg == 'E#1 - E and Iz must be non-zero'
    rc, msg = create(xyz, conn, node_order=[0, 1, 1])
    assert rc == INVALID and 'not a permutation' in msg
    # a valid description gets as far as the device: OK with a GPU, the CUDA error (never a CPU fallback) without
    rc, msg = create(xyz, conn)
    import torch
    if torch.cuda.is_available():
        assert rc == 0, msg
    else:
        assert rc not in (0, INVALID) and 'cuda' in msg.lower()


def test_bench_parity_in_run_check(tmp_path):
    """bench.py compares the CUDA results of a run with the CPU arm's results for the same subsets (BASELINE.md
    section 3, step 6).  Here the comparison itself is exercised without a GPU: the CPU arm's parity set (the
    unmodified reference when baseline/_ref is installed, else the port) against the oracle port's results laid out
    like the device outputs - which also pins the port to the reference on config-3 subsets once more."""
    import importlib.util
    from oracle import cpu_bench
    from conmech_b200 import synthetic as syn
    spec = importlib.util.spec_from_file_location('bench_module', os.path.join(REPO, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    n = 6
    wl = dict(cpu_bench.CONFIG3)
    path = str(tmp_path / 'parity.npz')
    kind = cpu_bench.parity_sample(wl, n, path)
    assert kind in ('reference', 'port')
    with np.load(path) as z:
        ref = {k: z[k] for k in z.files}
    assert ref['has'].all() and ref['U'].shape == (n, 6 * 392) and ref['eR'].shape == (n, 931, 12)

    # "device" results: the port, dense, zero rows for absent elements, [n, load case, ...]
    model = syn.grid_frame(7, 7, 8)
    from oracle.frame_oracle import FrameOracle
    from conmech_b200 import LoadCase, GravityLoad
    orc = FrameOracle(model)
    orc.set_loads(LoadCase(gravity_load=GravityLoad([0, 0, -1])))
    orc.set_nodal_displacement_tol(1e-3)
    subsets = [syn.connected_subset(model, s) for s in range(n)]
    got = {'flag': np.zeros((n, 1), np.uint8), 'status': np.zeros((n, 1), np.uint8), 'max_trans': np.zeros((n, 1)),
           'compliance': np.zeros((n, 1)), 'U': np.zeros((n, 1, 6 * 392)), 'Rf': np.zeros((n, 1, 6 * 392)),
           'eR': np.zeros((n, 1, 931, 12))}
    for b, ids in enumerate(subsets):
        o = orc.solve(ids)
        got['flag'][b, 0], got['max_trans'][b, 0], got['compliance'][b, 0] = o.success, o.max_trans, o.compliance
        got['U'][b, 0], got['Rf'][b, 0] = o.U, o.R
        got['eR'][b, 0][ids] = np.asarray(o.eR)[ids]
    masks = syn.pack_masks(subsets, 931)
    rec = bench.parity_in_run(ref, got, masks, 931)
    assert rec['pass'] and rec['flags_equal'] and rec['solved_equal'] and rec['subsets'] == n and rec['against'] == kind
    assert max(rec['max_rel_err'].values()) <= 1e-9
    # a wrong displacement, a wrong flag and a missing result are all caught
    bad = {k: v.copy() for k, v in got.items()}
    bad['U'][2, 0, 6 * 200 + 2] += 1e-6
    r2 = bench.parity_in_run(ref, bad, masks, 931)
    assert not r2['pass'] and r2['max_rel_err']['U'] > 1e-9 and r2['flags_equal']
    bad = {k: v.copy() for k, v in got.items()}
    bad['flag'][1, 0] ^= 1
    assert not bench.parity_in_run(ref, bad, masks, 931)['flags_equal']
    bad = {k: v.copy() for k, v in got.items()}
    bad['status'][3, 0] = 3
    r3 = bench.parity_in_run(ref, bad, masks, 931)
    assert not r3['solved_equal'] and not r3['pass']
