"""``CudaStiffness``: the B200 engine behind ``StiffnessChecker(checker_engine="cuda")``.

Host-side mirror of the reference's ``NumpyStiffness`` (``src/pyconmech/frame_analysis/
numpy_stiffness.py:34-634``): same constructor, load setters, ``solve`` / ``get_solved_results``
and procedural getters, same return conventions (numerical failure is a ``False`` return,
programmer errors raise).  All arithmetic of the hot path runs in hand-written sm_100a
kernels reached through the C ABI of ``include/conmech_b200.h``; this module only flattens the
``Model`` into arrays once, packs element-id lists into bit masks and reshapes results.
There is no CPU fallback: without the CUDA library or a CUDA device the constructor raises.

Beyond the reference interface it adds the batched entry points ``solve_many`` (host
buffers) and ``solve_many_device`` (torch CUDA tensors, caller's stream).  With
``devices=[0, 1, ...]`` one process drives several GPUs: the model tables are replicated, a batch
is range-partitioned over the devices (one host thread each, inside the library) and the results
land directly in the caller's arrays - no collective.
"""
import ctypes as C
import os

import numpy as np

from . import _capi
from ._capi import OUT_FIELDS, ST_OK, ST_NO_SUPPORT, ST_NO_FIXED_DOF, ST_SINGULAR, ST_ZERO_LOAD
from .masks import pack_masks, as_mask_matrix, publish
from .stiffness_base import StiffnessBase

DOUBLE_EPS = 1e-14
DEFAULT_GRAVITY_DIRECTION = np.array([0., 0., -1.])

_OUT_DTYPE = {'flag': np.uint8, 'status': np.uint8, 'n_free': np.int32, 'n_fixed': np.int32,
              'dof_map': np.int32, 'node_exists': np.uint8, 'U': np.float64, 'Rf': np.float64,
              'eR': np.float64, 'compliance': np.float64, 'max_trans': np.float64,
              'Rf_sup': np.float64, 'eR_rows': np.float64, 'eR_row_offsets': np.int64,
              'max_rot': np.float64, 'min_pivot': np.float64, 'profile_flops': np.float64,
              'sigma_max': np.float64, 'max_abs_sigma': np.float64, 'profile_entries': np.float64,
              'envelope_tiles': np.float64}


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class CudaStiffness(StiffnessBase):
    def __init__(self, model, verbose=False, output_json=False, device=None, devices=None, node_order=None):
        super(CudaStiffness, self).__init__(model, verbose, output_json)
        self._lib = _capi.load()
        self._model = model
        self._node_order = None if node_order is None else np.ascontiguousarray(node_order, dtype=np.int32)
        self._static_pos = None                # solve_sequence: node -> position in the handle's static ordering
        if devices is not None:
            devices = [int(d) for d in devices]
            if len(devices) == 0 or len(set(devices)) != len(devices):
                raise ValueError('devices must be a non-empty list of distinct CUDA device ordinals')
            device = devices[0]
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', 0)) if 'LOCAL_RANK' in os.environ else 0
        self._device = int(device)
        self._devices = devices or [self._device]
        self._handle = C.c_void_p()
        self._handles = []                     # one per device; _handle is the first

        self._trans_tol = 1e-3
        self._rot_tol = np.pi / 180 * 3
        self._include_self_weight_load = False
        self._gravity_direction = np.array(DEFAULT_GRAVITY_DIRECTION)
        self._node_neighbors = self.get_node_neighbors()
        self._e_ind_from_tag = self.get_elem_ind_from_elem_tag()

        nV, nE = self.nV, self.nE
        self._nodal_load = np.zeros(nV * 6)
        self._udl = None                       # None or {elem index: q (3,)}
        self._loadcase_list = None             # set_load_cases(): explicit multi load case table
        self._loads_dirty = True
        self._n_lc = 1
        self._n_geom = 1

        self._has_stored_deformation = False
        self._stored_existing_ids = []
        self._stored_nD = None
        self._stored_eR = None
        self._stored_fR = None
        self._stored_compliance = None
        self._last_status = None

        self._element_k_list = None
        self._rot_list = None
        self._ws_cache = {}

        self._v_id_map = np.arange(6 * nV, dtype=int).reshape(nV, 6)         # numpy_stiffness.py:563-565
        self._flatten_and_create(model)
        for h in self._handles:
            self._lib.cm_set_tol(h, self._trans_tol, self._rot_tol)

    def close(self):
        """Destroy the device handles (model tables, workspace, staging) now instead of at collection."""
        self.__del__()
        self._ws_cache = {}

    def __del__(self):
        for h in getattr(self, '_handles', []):
            if h is not None and h.value:
                try:
                    self._lib.cm_destroy(h)
                except Exception:
                    pass
        self._handles = []
        self._handle = C.c_void_p()

    # ------------------------------------------------------------------ model flattening
    @property
    def nV(self):
        return len(self._nodes)

    @property
    def nS(self):
        return len(self._sup_node)

    @property
    def nE(self):
        return len(self._elements)

    def _flatten_and_create(self, model):
        """Model -> SoA arrays -> cm_create.  Per-element lookups follow
        numpy_stiffness.py:579-614; the support table follows :172-184 (Q11: the rotational
        pin looks at neighbours in the FULL model)."""
        nV, nE = self.nV, self.nE
        xyz = np.array([n.point for n in self._nodes], dtype=np.float64).reshape(nV, 3)
        conn = np.array([e.end_node_inds for e in self._elements], dtype=np.int32).reshape(nE, 2)
        lengths = np.linalg.norm(xyz[conn[:, 0]] - xyz[conn[:, 1]], axis=1)
        if (lengths < DOUBLE_EPS).any():
            i = int(np.argmax(lengths < DOUBLE_EPS))
            raise ValueError('E#{} - Bar length too small!: Node {} - Node {}'.format(i, xyz[conn[i, 0]], xyz[conn[i, 1]]))
        # tag -> table row once per TAG (stiffness_base.py:131-156), then the per-element arrays in C
        # (cm_pack_elements) instead of the reference's per-element Python loop
        sec_rows, mat_rows, joint_rows = [], [], []
        sec_of, mat_of, joint_of = np.empty(nE, np.int32), np.empty(nE, np.int32), np.full(nE, -1, np.int32)
        by_tag = {}
        for i, element in enumerate(self._elements):
            tag = element.elem_tag
            if tag not in by_tag:
                cs, mat, joint = self.get_element_crosssec(i), self.get_element_material(i), self.get_element_joint(i)
                assert abs(mat.E) > DOUBLE_EPS and abs(cs.Iz) > DOUBLE_EPS
                sec_rows.append([cs.A, cs.Jx, cs.Iy, cs.Iz])
                mat_rows.append([mat.E, mat.mu, mat.density])
                j = -1
                if joint is not None:
                    c = joint.c_conditions
                    assert all([x is None or x == np.inf for x in c[0:3]]) and \
                           all([x is None or x == np.inf for x in c[6:9]]), "We do not support translational joint now!"
                    joint_rows.append([np.inf if x is None else x for x in list(c[3:6]) + list(c[9:12])])
                    j = len(joint_rows) - 1
                by_tag[tag] = (len(sec_rows) - 1, len(mat_rows) - 1, j)
            sec_of[i], mat_of[i], joint_of[i] = by_tag[tag]
        stiff = np.array([1 if e.bending_stiff else 0 for e in self._elements], dtype=np.uint8)
        sections = np.array(sec_rows, dtype=np.float64).reshape(-1, 4)
        materials = np.array(mat_rows, dtype=np.float64).reshape(-1, 3)
        joints = np.array(joint_rows, dtype=np.float64).reshape(-1, 6)
        props = np.zeros((nE, 7))
        cr = np.zeros((nE, 6))
        _capi.check(self._lib.cm_pack_elements(nE, _ptr(sec_of), _ptr(mat_of), _ptr(joint_of), _ptr(stiff), len(sections),
                                               _ptr(sections), len(materials), _ptr(materials), len(joints),
                                               _ptr(joints) if len(joints) else None, _ptr(props), _ptr(cr)))
        sup_node = np.array(list(self._supports.keys()), dtype=np.int32)
        sup_cond = np.zeros((len(sup_node), 6), dtype=np.uint8)
        for k, v in enumerate(sup_node):
            cond = np.array(self._supports[int(v)].condition, dtype=int)
            assert all(s >= 0 for s in cond)
            if all([not self._elements[e].bending_stiff for e in self._node_neighbors[int(v)]]):
                cond[3:] = 1
            sup_cond[k] = cond
        self._xyz, self._conn, self._props, self._cr = xyz, conn, props, cr
        self._sup_node, self._sup_cond = sup_node, sup_cond
        self._id_map = np.hstack([self._v_id_map[conn[:, 0]], self._v_id_map[conn[:, 1]]])   # :567-570

        for dev in self._devices:
            desc = _capi.ModelDesc(nV, nE, len(sup_node), dev,
                                   xyz.ctypes.data_as(_capi.c_double_p), conn.ctypes.data_as(_capi.c_int32_p),
                                   props.ctypes.data_as(_capi.c_double_p), cr.ctypes.data_as(_capi.c_double_p),
                                   sup_node.ctypes.data_as(_capi.c_int32_p), sup_cond.ctypes.data_as(_capi.c_uint8_p),
                                   self._node_order.ctypes.data_as(_capi.c_int32_p) if self._node_order is not None else None)
            h = C.c_void_p()
            _capi.check(self._lib.cm_create(C.byref(desc), C.byref(h)))
            self._handles.append(h)
        self._handle = self._handles[0]
        self._info = _capi.ModelInfo()
        _capi.check(self._lib.cm_get_info(self._handle, C.byref(self._info)))

    @property
    def info(self):
        """Static solver facts (free DOFs, half bandwidth, tile counts, profile flops)."""
        i = self._info
        return {f: getattr(i, f) for f, _ in i._fields_}

    # ------------------------------------------------------------------ loads
    def set_self_weight_load(self, include_self_weight=True, gravity_direction=DEFAULT_GRAVITY_DIRECTION):
        """numpy_stiffness.py:98-111."""
        self._include_self_weight_load = include_self_weight
        if include_self_weight:
            assert len(gravity_direction) == 3
            self._gravity_direction = np.array(gravity_direction, dtype=float)
        self._loadcase_list = None
        self._loads_dirty = True

    def set_uniformly_distributed_loads(self, unif_dist_loads):
        """numpy_stiffness.py:113-123: {elem_tag: UniformlyDistLoad} or None to reset."""
        if unif_dist_loads is not None:
            d = {}
            for e_tag, uload in unif_dist_loads.items():
                for e_ind in self._e_ind_from_tag[e_tag]:
                    d[e_ind] = np.array(uload.q, dtype=float)
            self._udl = d
        else:
            self._udl = None
        self._loadcase_list = None
        self._loads_dirty = True

    def set_load(self, nodal_forces):
        """numpy_stiffness.py:125-138: {node id: PointLoad}; None resets.  Like the reference,
        a dict only overwrites the listed nodes (Q6)."""
        if nodal_forces is not None:
            for v_id, pf in nodal_forces.items():
                self._nodal_load[self._v_id_map[v_id, :]] = np.array(list(pf.force) + list(pf.moment))
        else:
            self._nodal_load = np.zeros(self.nV * 6)
        self._loadcase_list = None
        self._loads_dirty = True

    def set_load_cases(self, loadcases):
        """Batched extension: solve several ``LoadCase`` objects per subset in one pass (the
        factorisation is shared).  Each case is taken on its own (no carry-over between cases)."""
        tabs = []
        for lc in loadcases:
            nodal = np.zeros(6 * self.nV)
            for pl in lc.point_loads:
                nodal[6 * pl.node_ind:6 * pl.node_ind + 6] = list(pl.force) + list(pl.moment)
            udl = None
            if lc.uniform_element_loads:
                udl = {}
                for ul in lc.uniform_element_loads:
                    tags = ul.elem_tags if len(ul.elem_tags) > 0 else [""]
                    for tg in tags:
                        for e in self._e_ind_from_tag[tg]:
                            udl[e] = np.array(ul.q, dtype=float)
            g = lc.gravity_load
            tabs.append((nodal, udl, g is not None, np.array(g.force, dtype=float) if g is not None else DEFAULT_GRAVITY_DIRECTION))
        assert 1 <= len(tabs) <= 8
        self._loadcase_list = tabs
        self._loads_dirty = True

    def _push_loads(self):
        if not self._loads_dirty:
            return
        tabs = self._loadcase_list or [(self._nodal_load, self._udl, self._include_self_weight_load,
                                        self._gravity_direction)]
        n_lc, nV, nE = len(tabs), self.nV, self.nE
        nodal = np.zeros((n_lc, 6 * nV))
        w = np.zeros((n_lc, nE, 3))
        has = np.zeros((n_lc, nE), dtype=np.uint8)
        on = np.zeros(n_lc, dtype=np.uint8)
        gd = np.zeros((n_lc, 3))
        for k, (nl, udl, sw, gdir) in enumerate(tabs):
            nodal[k] = nl
            if udl:
                for e, q in udl.items():
                    w[k, e] = q
                    has[k, e] = 1
            on[k] = 1 if sw else 0
            gd[k] = gdir
        any_udl = bool(has.any())
        for h in self._handles:
            _capi.check(self._lib.cm_set_loadcases(h, n_lc, _ptr(nodal), _ptr(w) if any_udl else None,
                                                   _ptr(has) if any_udl else None, _ptr(on), _ptr(gd)))
        self._n_lc = n_lc
        self._loads_dirty = False

    # ------------------------------------------------------------------ geometry variants
    def set_geometries(self, node_points_list):
        """Many models of ONE topology (e.g. the iterates of a shape optimisation): ``node_points_list`` is
        a (G, nV, 3) array (or list of (nV, 3) arrays) of node coordinates in metres.  One launch of the
        element kernel builds the tables of all G variants on the device; ``solve_many(...,
        geometry=ids)`` then picks a variant per subset.  Replaces constructing one engine per shape
        (``Model.from_data`` io_base.py:107-144 + numpy_stiffness.py:579-614)."""
        xyz = np.ascontiguousarray(np.asarray(node_points_list, dtype=np.float64))
        if xyz.ndim != 3 or xyz.shape[1:] != (self.nV, 3):
            raise ValueError('node_points_list must have shape (G, {}, 3)'.format(self.nV))
        for h in self._handles:
            try:
                _capi.check(self._lib.cm_set_geometries(h, xyz.shape[0], _ptr(xyz)))
            except _capi.CmError as ex:
                if 'Bar length too small' in str(ex):
                    raise ValueError(str(ex))
                raise
        self._n_geom = xyz.shape[0]
        self._element_k_list = self._rot_list = None
        self._loads_dirty = True               # lumped element loads depend on the geometry

    def n_geometries(self):
        return int(self._lib.cm_geometry_count(self._handle))

    def last_element_kernel_ms(self):
        ms = C.c_double()
        _capi.check(self._lib.cm_last_element_kernel_ms(self._handle, C.byref(ms)))
        return ms.value

    # ------------------------------------------------------------------ tolerances
    def set_nodal_displacement_tol(self, transl_tol, rot_tol):
        self._trans_tol = transl_tol
        self._rot_tol = rot_tol
        for h in self._handles:
            self._lib.cm_set_tol(h, float(transl_tol), float(rot_tol))

    def get_nodal_deformation_tol(self):
        return (self._trans_tol, self._rot_tol)

    def set_factor_variant(self, variant):
        """1 = FP64 tensor-core factorisation, CTA size chosen per launch (default); 2 / 3 / 6 = pinned to
        4 / 8 / 16-warp CTAs; 4 / 5 = 4 / 8-warp CTAs with activity cycle counters; 7 / 8 / 9 = 2 / 3 / 1-warp
        CTAs; 10 / 11 / 12 = operands through a bulk-copy (TMA) shared-memory ring; 13 = look-ahead over the M_J waits with
        accumulators parked in tensor memory; 0 = scalar validation
        variant."""
        for h in self._handles:
            _capi.check(self._lib.cm_set_variant(h, int(variant)))

    # ------------------------------------------------------------------ batched solves
    def _as_masks(self, subsets):
        m = as_mask_matrix(subsets, self._info.mask_words)
        return m if m is not None else pack_masks(subsets, self.nE)

    def _out_shapes(self, B):
        nV, nE, L = self.nV, self.nE, self._n_lc
        return {'flag': (B, L), 'status': (B, L), 'n_free': (B,), 'n_fixed': (B,), 'dof_map': (B, 6 * nV),
                'node_exists': (B, nV), 'U': (B, L, 6 * nV), 'Rf': (B, L, 6 * nV), 'eR': (B, L, nE, 12),
                'compliance': (B, L), 'max_trans': (B, L), 'max_rot': (B, L), 'min_pivot': (B,),
                'profile_flops': (B,), 'sigma_max': (B, L, nE), 'max_abs_sigma': (B, L),
                'profile_entries': (B,), 'envelope_tiles': (B,), 'Rf_sup': (B, L, self.nS, 6)}

    @staticmethod
    def element_row_offsets(masks):
        """Exclusive prefix sum of the subsets' element counts, int64 [B+1]: where each subset's rows start in the
        compact ``eR_rows`` output (``cm_solve_out.eR_row_offsets``)."""
        masks = np.ascontiguousarray(masks, dtype=np.uint32)
        off = np.zeros(len(masks) + 1, dtype=np.int64)
        if len(masks):
            _capi.check(_capi.load().cm_element_row_offsets(_ptr(masks), len(masks), masks.shape[1], _ptr(off)))
        return off

    def solve_many(self, subsets, want=('flag', 'status', 'max_trans', 'compliance'), out=None, geometry=None):
        """Solve B element subsets (list of id lists, [] = full structure, or a (B, words)
        uint32 mask matrix) with HOST buffers: masks are copied to the GPU, the kernels run, and
        the requested outputs are copied back.  Returns {name: ndarray}; see ``cm_solve_out``
        in include/conmech_b200.h for names and shapes.  ``out`` may hold preallocated
        (e.g. pinned) arrays to write into.  ``geometry``: one geometry-variant index per subset
        (``set_geometries``), default variant 0.

        Compact results (what the reference's ``get_solved_results`` returns, without zero padding): ``'Rf_sup'`` =
        reactions of the support nodes ``(B, L, nS, 6)`` in the order of the model's supports; ``'eR_rows'`` = end
        forces of the EXISTING elements only, ``(rows * L, 12)``: subset b owns ``rows[off[b] * L:off[b + 1] * L]``
        viewed as ``(L, off[b + 1] - off[b], 12)``, elements in ascending id; ``off`` is returned as
        ``'eR_row_offsets'``."""
        self._push_loads()
        masks0 = as_mask_matrix(subsets, self._info.mask_words)
        prepacked = masks0 is not None
        if not prepacked and not isinstance(subsets, (list, tuple)):
            subsets = list(subsets)            # generators, sets of tuples, id-row arrays
        B = len(masks0) if prepacked else len(subsets)
        shapes = self._out_shapes(B)
        res = {}
        if 'eR_rows' in want:                  # the row offsets come from the masks: pack them all first
            if not prepacked:
                masks0, prepacked = pack_masks(subsets, self.nE), True
            off = self.element_row_offsets(masks0)
            shapes['eR_rows'] = (int(off[-1]) * self._n_lc, 12)
            res['eR_row_offsets'] = off
        for name in OUT_FIELDS:
            if name in want and name != 'eR_row_offsets':
                a = out[name] if out is not None and name in out else np.empty(shapes[name], dtype=_OUT_DTYPE[name])
                assert a.flags['C_CONTIGUOUS'] and a.dtype == _OUT_DTYPE[name] and a.size == int(np.prod(shapes[name]))
                res[name] = a.reshape(shapes[name])

        def out_struct(b0=0):
            oh = _capi.SolveOutHost()
            for name in OUT_FIELDS:
                if name in ('eR_rows', 'eR_row_offsets'):
                    assert b0 == 0 or name not in res
                    setattr(oh, name, res[name].ctypes.data if name in res and res[name].size else None)
                else:
                    setattr(oh, name, res[name][b0:].ctypes.data if name in res and b0 < B else None)
            if 'eR_rows' in res and not res['eR_rows'].size:
                oh.eR_row_offsets = None
            return oh

        if B == 0:
            return res
        geom = None
        if geometry is not None:
            geom = np.ascontiguousarray(np.asarray(geometry, dtype=np.int32).reshape(-1))
            if geom.size != B:
                raise ValueError('geometry must hold one variant index per subset')
        if geom is not None or prepacked or B < self.PIPELINE_MIN:
            masks = masks0 if prepacked else pack_masks(subsets, self.nE)
            oh = out_struct()
            if geom is not None:
                if len(self._handles) > 1:
                    raise NotImplementedError('geometry variants on several devices: shard the batch and use one engine per device')
                _capi.check(self._lib.cm_solve_many_host_g(self._handle, _ptr(masks), _ptr(geom), B, C.byref(oh)))
            elif len(self._handles) > 1:
                hs = (C.c_void_p * len(self._handles))(*[h.value for h in self._handles])
                _capi.check(self._lib.cm_solve_many_host_multi(hs, len(self._handles), _ptr(masks), B, C.byref(oh)))
            else:
                _capi.check(self._lib.cm_solve_many_host(self._handle, _ptr(masks), B, C.byref(oh)))
            return res
        # Python id lists, large batch: packing (this thread, GIL held, ~5 ns per id) is pipelined with
        # the GPU work.  ONE worker thread per device makes ONE call into the library for its share of the
        # batch (cm_solve_many_host_chunks); the library enqueues a wave of subsets as soon as the chunks it
        # covers are there - it waits, without the GIL, for the 'ready' flags this thread release-stores after
        # writing each chunk's masks.
        import threading
        from .sharding import shard_range
        G = len(self._handles)
        masks = np.empty((B, self._info.mask_words), dtype=np.uint32)
        jobs = []                              # (handle, lo, hi, step, ready flags)
        for g, h in enumerate(self._handles):
            lo, hi = shard_range(B, G, g)
            if hi == lo:
                continue
            n_chunks = max(1, min(64, (hi - lo) // self.PIPELINE_CHUNK))
            step = -(-(hi - lo) // n_chunks)
            n_chunks = -(-(hi - lo) // step)
            jobs.append((h, lo, hi, step, np.zeros(n_chunks, dtype=np.int32)))
        results = [None] * len(jobs)

        def worker(j):
            h, lo, hi, step, ready = jobs[j]
            oh = out_struct(lo)
            rc = self._lib.cm_solve_many_host_chunks(h, masks[lo:].ctypes.data, hi - lo, C.byref(oh), step,
                                                     ready.ctypes.data)
            # the library's error message is thread-local: read it on THIS thread
            results[j] = (rc, self._lib.cm_last_error().decode() if rc != _capi.CM_OK else '')

        threads = [threading.Thread(target=worker, args=(j,), name='conmech-b200-solve-{}'.format(j), daemon=True)
                   for j in range(len(jobs))]
        for th in threads:
            th.start()
        try:
            # chunks are packed round-robin over the devices so that every device starts early
            order = sorted((k, j) for j, job in enumerate(jobs) for k in range(len(job[4])))
            for k, j in order:
                h, lo, hi, step, ready = jobs[j]
                b0, b1 = lo + k * step, min(hi, lo + (k + 1) * step)
                pack_masks(subsets[b0:b1], self.nE, out=masks[b0:b1])
                publish(ready, k, 1)
        except BaseException:
            for job in jobs:                   # cancel the chunks that were not handed over
                for k in np.flatnonzero(job[4] == 0):
                    publish(job[4], int(k), -1)
            for th in threads:
                th.join()
            raise
        for th in threads:
            th.join()
        for rc, msg in results:
            if rc != _capi.CM_OK:
                raise _capi.CmError('conmech_b200 error {}: {}'.format(rc, msg))
        return res

    # ------------------------------------------------------------------ construction sequences
    SEQUENCE_GROUP = 32            # consecutive prefixes that share the factor rows of the group's first prefix
                                   # (cantilever_beam_truss, 2197 prefixes: 8 -> 20.7 ms, 16 -> 16.8, 32 -> 15.5, 64 -> 16.4; independent 21.9)

    def solve_sequence(self, order, want=('flag', 'status', 'max_trans', 'compliance'), group=None):
        """Every prefix ``order[:k+1]`` of a construction sequence, with the factorisation REUSED along the sequence:
        result row k is what ``solve_many`` returns for the subset ``order[:k+1]`` (same arithmetic in the same order
        for every row of the factor, whether computed or shared).

        The solver numbers the existing nodes in the handle's static ordering; a prefix differs from an earlier one by
        the nodes and elements added since.  Every solver row in front of the first row that an added element touches -
        or that a new node is inserted before - is identical in both.  Prefixes are therefore solved in groups of
        ``group``: the first of a group completely (the trunk), the others share the trunk's factor rows (tiles of L,
        inverted diagonal factors, pivots, forward solutions) in front of that row and factor only the rows behind it
        (``cm_solve_many_host_seq``).  Replaces the loop ``for k: sc.solve(order[:k])`` of the sequencing planners
        around the reference (README.rst:39-42, examples/notebook_demo/demo.ipynb), which re-assembles and re-factors
        everything N times."""
        self._push_loads()
        order = [int(e) for e in order]
        N, nE, nV = len(order), self.nE, self.nV
        if N == 0:
            return {name: np.zeros((0,) + self._out_shapes(0)[name][1:], dtype=_OUT_DTYPE[name]) for name in want}
        oa = np.asarray(order, dtype=np.int64)
        if oa.min() < 0 or oa.max() >= nE:
            raise IndexError('element id not within range!')
        if len(set(order)) != N:
            raise ValueError('duplicate element id in the sequence')
        group = int(group or self.SEQUENCE_GROUP)
        t_rows = self._sequence_shared_rows(oa, group)
        # prefix masks by a cumulative OR
        W = self._info.mask_words
        words = np.zeros((N, W), dtype=np.uint32)
        words[np.arange(N), oa >> 5] = np.uint32(1) << (oa & 31).astype(np.uint32)
        masks = np.bitwise_or.accumulate(words, axis=0)
        shapes = self._out_shapes(N)
        res = {name: np.empty(shapes[name], dtype=_OUT_DTYPE[name]) for name in OUT_FIELDS if name in want}
        oh = _capi.SolveOutHost()
        for name in OUT_FIELDS:
            setattr(oh, name, res[name].ctypes.data if name in res else None)
        _capi.check(self._lib.cm_solve_many_host_seq(self._handle, _ptr(masks), N, group, _ptr(t_rows), C.byref(oh)))
        return res

    def _sequence_shared_rows(self, oa, group):
        """t_rows[k]: 32-row tile rows of prefix k's factor that equal those of the first prefix of its group.
        In the trunk's compacted numbering a node at static position p - present or not - sits in front of solver row
        cum[p] = (free DOFs of the trunk's nodes at positions < p); an added element touches the rows of its end nodes,
        a new node is inserted in front of row cum[p]: everything in front of the smallest such row is untouched."""
        N, nV = len(oa), self.nV
        if getattr(self, '_static_pos', None) is None:
            no = np.zeros(nV, dtype=np.int32)
            _capi.check(self._lib.cm_get_node_order(self._handle, _ptr(no)))
            pos = np.empty(nV, dtype=np.int64)
            pos[no] = np.arange(nV)
            self._static_pos, self._static_order = pos, no.astype(np.int64)
        pos, no = self._static_pos, self._static_order
        nfree = np.full(nV, 6, dtype=np.int64)
        nfree[self._sup_node] = 6 - self._sup_cond.sum(axis=1)
        ends = self._conn[oa].astype(np.int64)                                   # (N, 2)
        appeared = np.full(nV, N, dtype=np.int64)                                # step at which a node first exists
        np.minimum.at(appeared, ends.ravel(), np.repeat(np.arange(N), 2))
        free_end = nfree[ends] > 0
        t_rows = np.zeros(N, dtype=np.int32)
        big = np.iinfo(np.int64).max
        for g0 in range(0, N, group):
            g1 = min(N, g0 + group)
            if g1 - g0 < 2:
                continue
            trunk_rows = np.where(appeared[no] <= g0, nfree[no], 0)              # by static position
            cum = np.concatenate([[0], np.cumsum(trunk_rows)])[:-1]
            r = np.where(free_end[g0 + 1:g1], cum[pos[ends[g0 + 1:g1]]], big).min(axis=1)
            t_rows[g0 + 1:g1] = np.minimum(np.minimum.accumulate(r) // 32, 2 ** 30).astype(np.int32)
        return t_rows

    def solve_many_csr(self, flat_ids, offsets, want=('flag', 'status', 'max_trans', 'compliance'), out=None):
        """``solve_many`` for subsets kept as two arrays: subset b = ``flat_ids[offsets[b]:offsets[b+1]]``
        (an empty range = the full structure).  No Python object per id: packing costs ~2 ns per id."""
        from .masks import pack_masks_csr
        return self.solve_many(pack_masks_csr(flat_ids, offsets, self.nE), want=want, out=out)

    PIPELINE_CHUNK = 512           # subsets per pipelined chunk of solve_many (id lists only): the GPU starts after the first ones
    PIPELINE_MIN = 4096            # smaller batches are packed in one go

    def solve_many_device(self, masks, want=('flag', 'status', 'max_trans', 'compliance'), out=None,
                          max_in_flight=None, geometry=None):
        """Same on DEVICE buffers: ``masks`` is a (B, words) int32/uint32 torch CUDA tensor; outputs
        are torch CUDA tensors; kernels are enqueued on torch's current stream and the call
        returns without synchronising."""
        import torch
        self._push_loads()
        assert masks.is_cuda and masks.dim() == 2 and masks.shape[1] == self._info.mask_words
        assert masks.element_size() == 4 and masks.is_contiguous()
        dev = masks.device
        B = masks.shape[0]
        shapes = self._out_shapes(B)
        tdt = {np.uint8: torch.uint8, np.int32: torch.int32, np.float64: torch.float64, np.int64: torch.int64}
        res = {}
        od = _capi.SolveOut()
        if 'eR_rows' in want:
            # compact element rows: offsets = exclusive prefix sum of the masks' popcounts (a caller that already has
            # them passes out['eR_row_offsets'], int64 [B+1] on the device, first entry 0)
            off = out.get('eR_row_offsets') if out is not None else None
            if off is None:
                x = masks.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
                x = x - ((x >> 1) & 0x55555555)
                x = (x & 0x33333333) + ((x >> 2) & 0x33333333)
                x = (((x + (x >> 4)) & 0x0F0F0F0F) * 0x01010101 >> 24) & 0xFF
                off = torch.zeros(B + 1, dtype=torch.int64, device=dev)
                torch.cumsum(x.sum(dim=1), 0, out=off[1:])
            assert off.is_cuda and off.dtype == torch.int64 and off.numel() == B + 1 and off.is_contiguous()
            res['eR_row_offsets'] = off
            n_rows = int(out['eR_rows'].numel() // (12 * self._n_lc)) if out is not None and 'eR_rows' in out else int(off[-1].item())
            shapes['eR_rows'] = (n_rows * self._n_lc, 12)
        for name in OUT_FIELDS:
            if name == 'eR_row_offsets':
                setattr(od, name, res[name].data_ptr() if name in res else 0)
            elif name in want:
                tns = out[name] if out is not None and name in out else \
                    torch.empty(shapes[name], dtype=tdt[_OUT_DTYPE[name]], device=dev)
                assert tns.is_contiguous()
                res[name] = tns
                setattr(od, name, tns.data_ptr())
            else:
                setattr(od, name, 0)
        if max_in_flight is None:
            max_in_flight = int(os.environ.get('CM_WS_SUBSETS', 8192))
        n_fly = int(min(B, max_in_flight))
        nbytes = int(self._lib.cm_workspace_bytes(self._handle, n_fly))
        key = (dev.index, self._n_lc)
        ws = self._ws_cache.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._ws_cache = {key: ws}
        stream = torch.cuda.current_stream(dev).cuda_stream
        gptr = 0
        if geometry is not None:              # (B,) int32 CUDA tensor of geometry-variant indices
            assert geometry.is_cuda and geometry.dtype == torch.int32 and geometry.numel() == B and geometry.is_contiguous()
            gptr = geometry.data_ptr()
        _capi.check(self._lib.cm_solve_many_g(self._handle, masks.data_ptr(), gptr, B, C.byref(od), ws.data_ptr(),
                                              ws.numel(), stream))
        return res

    def kernel_launches(self):
        return int(self._lib.cm_kernel_launches(self._handle))

    def reset_counters(self):
        self._lib.cm_reset_counters(self._handle)

    def enable_timing(self, on=True):
        self._lib.cm_enable_timing(self._handle, 1 if on else 0)

    def last_stage_ms(self):
        ms = np.zeros(3)
        self._lib.cm_last_stage_ms(self._handle, _ptr(ms))
        return ms

    def last_wave_count(self):
        return int(self._lib.cm_last_wave_count(self._handle))

    # ------------------------------------------------------------------ the reference interface
    def solve(self, exist_element_ids=[], if_cond_num=False):
        """numpy_stiffness.py:228-397 for one subset; stores nD / fR / eR like the reference
        (row orders per SURVEY Q8)."""
        ids = list(exist_element_ids)
        if len(ids) == 0:
            ids = list(range(self.nE))
        r = self.solve_many([ids], want=('flag', 'status', 'U', 'Rf', 'eR', 'compliance', 'n_free', 'n_fixed'))
        status = int(r['status'][0, 0])
        self._last_status = status
        if status in (ST_NO_SUPPORT, ST_NO_FIXED_DOF, ST_SINGULAR):
            if self._verbose:
                print({ST_NO_SUPPORT: 'Structure floating: no fixed node exists in the partial structure.',
                       ST_NO_FIXED_DOF: 'Not stable: At least one node needs to be fixed in the considered substructure!',
                       ST_SINGULAR: 'ERROR: Stiffness Solver fail! Stiffness matrix not Positive Definite, '
                                    'the sub-structure might contain mechanism.'}[status])
            self._has_stored_deformation = False
            return False
        U, R, E = r['U'][0, 0], r['Rf'][0, 0], r['eR'][0, 0]
        exist_node_set = set()
        for e in ids:
            exist_node_set.update(self._elements[e].end_node_inds)
        self._stored_nD = np.zeros((len(exist_node_set), 7))
        for i, v in enumerate(list(exist_node_set)):
            self._stored_nD[i, 0] = v
            self._stored_nD[i, 1:] = U[6 * v:6 * v + 6]
        fix_rows = [v for v in self._supports if v in exist_node_set]
        self._stored_fR = np.zeros((len(fix_rows), 7))
        for i, v in enumerate(fix_rows):
            self._stored_fR[i, 0] = v
            self._stored_fR[i, 1:] = R[6 * v:6 * v + 6]
        self._stored_eR = np.zeros((len(ids), 13))
        for i, e in enumerate(ids):
            self._stored_eR[i, 0] = e
            self._stored_eR[i, 1:] = E[e]
        self._stored_compliance = float(r['compliance'][0, 0])
        self._stored_existing_ids = np.copy(ids)
        self._has_stored_deformation = True
        return bool(r['flag'][0, 0])

    def check_stiffness_criteria(self, node_displ, fixities_reaction, element_reaction, ord=None):
        nD_trans_max, _, _, _ = self.get_max_nodal_deformation(ord=ord)
        return bool(nD_trans_max < self._trans_tol)

    def has_stored_result(self):
        return self._has_stored_deformation

    def get_solved_results(self):
        assert self.has_stored_result()
        success = self.check_stiffness_criteria(self._stored_nD, self._stored_fR, self._stored_eR)
        return (success, self._stored_nD, self._stored_fR, self._stored_eR)

    def get_compliance(self):
        assert self.has_stored_result()
        return self._stored_compliance

    def get_max_nodal_deformation(self, ord=None):
        """numpy_stiffness.py:457-464 (Q5: the indices are ROW indices of nD)."""
        assert self.has_stored_result()
        t = np.linalg.norm(self._stored_nD[:, 1:4], ord=ord, axis=1)
        r = np.linalg.norm(self._stored_nD[:, 4:7], ord=ord, axis=1)
        return (np.max(t), np.max(r), np.argmax(t), np.argmax(r))

    def get_frame_stat(self):
        return (self.nV, self.nE)

    # ---- procedural data
    def _element_loads(self, lc=0):
        self._push_loads()
        q = np.zeros((self.nE, 12))
        _capi.check(self._lib.cm_get_element_loads(self._handle, lc, _ptr(q)))
        return q

    def _scatter_element_loads(self, q, ids):
        P = np.zeros(6 * self.nV)
        ids = np.asarray(list(ids), dtype=int)
        np.add.at(P, self._id_map[ids].ravel(), q[ids].ravel())
        return P

    def get_gravity_nodal_loads(self, existing_ids=[]):
        assert self._include_self_weight_load, 'Call `set_self_weight_load` first to init the load!'
        if len(existing_ids) == 0:
            existing_ids = range(self.nE)
        # gravity-only table: push a temporary load case, fetch, restore lazily
        saved = (self._loadcase_list, self._loads_dirty)
        self._loadcase_list = [(np.zeros(6 * self.nV), None, True, self._gravity_direction)]
        self._loads_dirty = True
        self._push_loads()
        q = self._element_loads(0)
        self._loadcase_list = saved[0]
        self._loads_dirty = True
        return self._scatter_element_loads(q, existing_ids)

    def get_lumped_nodal_loads(self, existing_ids=[]):
        if len(existing_ids) == 0:
            existing_ids = range(self.nE)
        self._push_loads()
        P = np.copy(self._nodal_load) if self._loadcase_list is None else np.copy(self._loadcase_list[0][0])
        return P + self._scatter_element_loads(self._element_loads(0), existing_ids)

    def _fetch_tables(self):
        if self._element_k_list is None:
            K = np.zeros((self.nE, 12, 12))
            R3 = np.zeros((self.nE, 3, 3))
            _capi.check(self._lib.cm_get_element_tables(self._handle, _ptr(K), _ptr(R3), None))
            R12 = np.zeros((self.nE, 12, 12))
            for i in range(4):
                R12[:, 3 * i:3 * i + 3, 3 * i:3 * i + 3] = R3
            self._element_k_list = [K[e] for e in range(self.nE)]
            self._rot_list = [R12[e] for e in range(self.nE)]

    def get_element_stiffness_matrices(self):
        self._fetch_tables()
        return self._element_k_list

    def get_element_local2global_rot_matrices(self):
        self._fetch_tables()
        return self._rot_list

    def get_element2dof_id_map(self):
        return self._id_map

    def get_node2dof_id_map(self):
        return self._v_id_map

    def compute_full_stiffness_matrix(self, exist_element_ids=[]):
        """numpy_stiffness.py:221-226: the (6nV)^2 sparse K of a subset, from the GPU-built
        element tables (a procedural getter, not on the batched path)."""
        from scipy.sparse import csc_matrix
        ids = np.asarray(list(exist_element_ids) or list(range(self.nE)), dtype=int)
        K = np.asarray(self.get_element_stiffness_matrices())
        rows = np.repeat(self._id_map[ids][:, :, None], 12, axis=2).ravel()
        cols = np.repeat(self._id_map[ids][:, None, :], 12, axis=1).ravel()
        data = K[ids].ravel()
        keep = np.abs(data) > DOUBLE_EPS
        n = 6 * self.nV
        return csc_matrix((data[keep], (rows[keep], cols[keep])), shape=(n, n), dtype=float)

    def compute_dof_permutation(self, exist_element_ids=[]):
        """numpy_stiffness.py:147-219 from the GPU DOF map: (Perm, (node set, n_free, n_fixed,
        n existing fixed nodes)); ValueError when no support exists, bare False when the
        supports fix nothing (Q10)."""
        from scipy.sparse import csc_matrix
        ids = list(exist_element_ids) or list(range(self.nE))
        r = self.solve_many([ids], want=('status', 'n_free', 'n_fixed', 'dof_map', 'node_exists'))
        status = int(r['status'][0, 0])
        if status == ST_NO_SUPPORT:
            raise ValueError('Structure floating: no fixed node exists in the partial structure.')
        if status == ST_NO_FIXED_DOF:
            return False
        n = 6 * self.nV
        Perm = csc_matrix((np.ones(n, dtype=int), (np.arange(n), r['dof_map'][0])), shape=(n, n))
        node_set = set(int(v) for v in np.flatnonzero(r['node_exists'][0]))
        n_fix_nodes = len([v for v in self._supports if v in node_set])
        return Perm, (node_set, int(r['n_free'][0]), int(r['n_fixed'][0]), n_fix_nodes)

    def get_node_neighbors(self):
        nb = {}
        for e in self._elements:
            for n in e.end_node_inds:
                nb.setdefault(n, set()).add(e.elem_ind)
        for v in range(len(self._nodes)):
            nb.setdefault(v, set())
        return nb

    def get_elem_ind_from_elem_tag(self):
        d = {}
        for e in self._elements:
            d.setdefault(e.elem_tag, []).append(e.elem_ind)
        return d

    def get_original_shape(self):
        raise NotImplementedError()

    def get_deformed_shape(self, exagg_ratio=1.0, disc=10):
        raise NotImplementedError()

    def set_output_json_path(self, file_path, file_name):
        raise NotImplementedError()

    def set_output_json(self, output_json=False):
        raise NotImplementedError()
