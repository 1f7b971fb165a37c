"""User facade: ``StiffnessChecker`` with ``checker_engine="cuda"``.

Keeps the public API of the reference's facade (``src/pyconmech/frame_analysis/
stiffness_checker.py:29-769``) - ``from_json``, ``set_loads`` / ``LoadCase``,
``set_nodal_displacement_tol``, ``solve(existing_ids)``, ``get_solved_results`` and the query
helpers - and adds the batched ``solve_many(list_of_existing_ids)``.  Only the engine differs:
every solve runs on the GPU through :class:`conmech_b200.CudaStiffness`.  Other engine names
raise ``NotImplementedError`` exactly like unknown engines do in the reference
(``stiffness_checker.py:65-66``); there is no CPU engine in this package.
"""
import json
import os
import warnings
from collections import OrderedDict, defaultdict

import numpy as np

from .io_base import Model
from .cuda_stiffness import CudaStiffness

DEFAULT_ENGINE = 'cuda'


class StiffnessChecker(object):
    def __init__(self, model, verbose=False, checker_engine=DEFAULT_ENGINE, device=None, devices=None):
        """``device``: CUDA ordinal (default: LOCAL_RANK or 0).  ``devices=[0, 1, ...]``: one process
        drives several GPUs - ``solve_many`` range-partitions every batch over them."""
        if checker_engine == 'cuda':
            engine = CudaStiffness(model, verbose=verbose, device=device, devices=devices)
        else:
            raise NotImplementedError('Engine not exist: {}'.format(checker_engine))
        self._sc_ins = engine
        self.set_nodal_displacement_tol()
        self.model = model
        if not len(model.supports) > 0:
            raise RuntimeError('there needs to be at least one support (fixed) vertex in the model!')

    @classmethod
    def from_json(cls, json_file_path=None, verbose=False, checker_engine=DEFAULT_ENGINE, device=None, devices=None):
        return cls(Model.from_json(json_file_path, verbose=verbose), verbose=verbose,
                   checker_engine=checker_engine, device=device, devices=devices)

    # ---------------------------------------------------------------- properties
    @property
    def model_name(self):
        return self.model.model_name

    @property
    def node_points(self):
        return np.array([n.point for n in self.model.nodes])

    @property
    def elements(self):
        return [e.end_node_inds for e in self.model.elements]

    @property
    def fix_node_ids(self):
        return list(self.model.supports.keys())

    @property
    def fix_element_ids(self):
        fixed = set(self.fix_node_ids)
        return [i for i, e in enumerate(self.elements) if any(v in fixed for v in e)]

    @property
    def supports(self):
        return self.model.supports

    # ---------------------------------------------------------------- solve
    def solve(self, exist_element_ids=[], if_cond_num=False, eid_sanity_check=False):
        """True iff the partial structure passes the stiffness criterion
        (reference ``stiffness_checker.py:153-182``)."""
        if eid_sanity_check:
            nE = len(self.model.elements)
            for e_id in exist_element_ids:
                assert 0 <= e_id < nE and int(e_id) == e_id, 'element id not within range!'
        return self._sc_ins.solve(exist_element_ids, if_cond_num=if_cond_num)

    def solve_many(self, list_of_existing_ids, return_details=False, with_stress=False):
        """Batched ``solve``: one flag per subset, all solved in one GPU pass.  With
        ``return_details`` also returns {'status','max_trans','compliance'} arrays;
        ``with_stress`` adds 'max_abs_sigma' (largest |sigma_max| over the subset's elements, the
        batched counterpart of ``get_sigma_max_per_element``) to them."""
        names = ('status', 'max_trans', 'compliance') + (('max_abs_sigma',) if with_stress else ())
        r = self._sc_ins.solve_many(list_of_existing_ids, want=('flag',) + names)
        flags = [bool(f) for f in r['flag'][:, 0]]
        if return_details or with_stress:
            return flags, {k: r[k][:, 0].copy() for k in names}
        return flags

    def solve_many_csr(self, flat_ids, offsets, return_details=False):
        """``solve_many`` for subsets kept as two integer arrays (flat element ids + offsets, an empty
        range = the full structure): the cheapest host format next to pre-packed bit masks."""
        from .masks import pack_masks_csr
        return self.solve_many(pack_masks_csr(flat_ids, offsets, len(self.model.elements)), return_details=return_details)

    def get_sigma_max_per_element_many(self, list_of_existing_ids):
        """``get_sigma_max_per_element`` for many subsets at once, computed on the device: a
        (B, nE) array (first load case), 0 for elements that are not in the subset."""
        r = self._sc_ins.solve_many(list_of_existing_ids, want=('flag', 'sigma_max'))
        return r['sigma_max'][:, 0].copy()

    def solve_shapes(self, node_points_list, existing_ids=None, return_details=False):
        """Stiffness check of G geometry variants of this frame - same elements, sections, joints and
        supports, node coordinates ``node_points_list[g]`` ((nV, 3), metres) - in one batched pass; the
        loop a shape optimiser around the reference runs as "new Model, new StiffnessChecker, solve" per
        iterate (``Model.from_data`` io_base.py:107-144).  ``existing_ids`` may hold
        one element subset per variant (default: the full structure).  Loads are those set on this
        checker; self weight and lumped line loads follow each variant's member lengths and axes.  The
        variants stay installed (``solve_many(..., )`` afterwards refers to variant 0)."""
        eng = self._sc_ins
        eng.set_geometries(node_points_list)
        G = eng.n_geometries()
        subsets = list(existing_ids) if existing_ids is not None else [[] for _ in range(G)]
        if len(subsets) != G:
            raise ValueError('existing_ids must hold one id list per geometry variant')
        r = eng.solve_many(subsets, want=('flag', 'status', 'max_trans', 'compliance'), geometry=np.arange(G, dtype=np.int32))
        flags = [bool(f) for f in r['flag'][:, 0]]
        if return_details:
            return flags, {k: r[k][:, 0].copy() for k in ('status', 'max_trans', 'compliance')}
        return flags

    SEQUENCE_REUSE_MIN = 256       # shorter sequences are one small batch: the independent path has less launch overhead

    def solve_sequence(self, order, return_details=False, reuse=True):
        """Stiffness check of EVERY prefix of a construction sequence in one call: flag k is
        ``solve(order[:k+1])``.  This is the loop the sequencing planners around the reference run one solve at a
        time (README.rst:39-42, examples/notebook_demo/demo.ipynb).  ``reuse=True`` (default) reuses the
        factorisation along the sequence (``CudaStiffness.solve_sequence``: prefixes share the factor rows in front of
        the rows their added elements touch); ``reuse=False`` solves the N prefix masks as N independent subsets."""
        order = np.asarray(list(order), dtype=np.int64)
        nE = len(self.model.elements)
        if order.size and (order.min() < 0 or order.max() >= nE):
            raise IndexError('element id not within range!')
        if reuse and order.size >= self.SEQUENCE_REUSE_MIN:
            r = self._sc_ins.solve_sequence(order.tolist())
            flags = [bool(f) for f in r['flag'][:, 0]]
            if return_details:
                return flags, {k: r[k][:, 0].copy() for k in ('status', 'max_trans', 'compliance')}
            return flags
        W = (nE + 31) // 32
        words = np.zeros((len(order), W), dtype=np.uint32)
        words[np.arange(len(order)), order >> 5] = np.uint32(1) << (order & 31).astype(np.uint32)
        masks = np.bitwise_or.accumulate(words, axis=0)
        return self.solve_many(masks, return_details=return_details)

    # ---------------------------------------------------------------- loads
    def set_loads(self, loadcase):
        """reference ``stiffness_checker.py:188-218``."""
        self.set_self_weight_load(loadcase.gravity_load)
        pls = loadcase.point_loads
        self._sc_ins.set_load({pl.node_ind: pl for pl in pls} if pls is not None and len(pls) > 0 else None)
        uls = loadcase.uniform_element_loads
        if uls is not None and len(uls) > 0:
            by_tag = {}
            for eload in uls:
                if len(eload.elem_tags) == 0:
                    eload.elem_tags = [""]
                for e_tag in eload.elem_tags:
                    by_tag[e_tag] = eload
        else:
            by_tag = None
        self._sc_ins.set_uniformly_distributed_loads(by_tag)

    def set_self_weight_load(self, gravity_load):
        if gravity_load is not None:
            self._sc_ins.set_self_weight_load(include_self_weight=True, gravity_direction=gravity_load.force)
        else:
            self._sc_ins.set_self_weight_load(include_self_weight=False)

    def set_load_cases(self, loadcases):
        """Batched extension: several load cases solved per subset (see CudaStiffness)."""
        self._sc_ins.set_load_cases(loadcases)

    # ---------------------------------------------------------------- criteria
    def set_nodal_displacement_tol(self, trans_tol=1e-3, rot_tol=np.inf):
        self._sc_ins.set_nodal_displacement_tol(trans_tol, rot_tol)

    def get_nodal_deformation_tol(self):
        return self._sc_ins.get_nodal_deformation_tol()

    # ---------------------------------------------------------------- results
    def has_stored_result(self):
        return self._sc_ins.has_stored_result()

    def get_solved_results(self, existing_ids=[]):
        """(success, {node: u[6]}, {fixed node: r[6]}, {elem: {0: f[6], 1: f[6]}})
        (reference ``stiffness_checker.py:281-328``; the reference raises NameError for a
        non-empty argument, Q13 - here the argument filters the dicts)."""
        assert self.has_stored_result(), "No result stored! Previous analysis might have failed."
        e_keep = set(existing_ids) if existing_ids else set(range(len(self.elements)))
        n_keep = set(self.get_element_connected_node_ids(existing_ids=existing_ids))
        success, nD_np, fR_np, eR_np = self._sc_ins.get_solved_results()
        fixed = set(self.fix_node_ids)
        nD = {int(r[0]): np.array(r[1:7]) for r in nD_np if int(r[0]) in n_keep}
        fR = {int(r[0]): np.array(r[1:7]) for r in fR_np if int(r[0]) in n_keep and int(r[0]) in fixed}
        eR = {int(r[0]): {0: np.array(r[1:7]), 1: np.array(r[7:13])} for r in eR_np if int(r[0]) in e_keep}
        return success, nD, fR, eR

    def get_sigma_max_per_element(self, existing_ids=[]):
        """Max axial normal stress per element, solid round sections only
        (reference ``stiffness_checker.py:330-365``; both section moduli use Iy there)."""
        warnings.warn('Maximal stress computation only support solid, round cross sec for now!')
        _, _, _, eR = self.get_solved_results(existing_ids)
        sigma_max = {}
        for e_id, er in eR.items():
            cs = self.get_element_crosssec(e_id)
            N = er[0][0]
            My = max(abs(er[0][4]), abs(er[1][4]))
            Mz = max(abs(er[0][5]), abs(er[1][5]))
            r = np.sqrt(cs.A / np.pi)
            Wy = cs.Iy / r
            Wz = cs.Iy / r
            sigma = N / cs.A
            sign = 1.0 if sigma > 0 else -1.0
            sigma_max[e_id] = sign * (abs(sigma) + My / Wy + Mz / Wz)
        return sigma_max

    def get_max_nodal_deformation(self):
        return self._sc_ins.get_max_nodal_deformation()

    def get_compliance(self):
        return self._sc_ins.get_compliance()

    def _loads_as(self, flat, existing_ids, dof_flattened):
        if dof_flattened:
            return flat
        n2d = self.get_node2dof_id_map()
        return {nid: flat[n2d[nid]] for nid in self.get_element_connected_node_ids(existing_ids)}

    def get_self_weight_loads(self, existing_ids=[], dof_flattened=False):
        return self._loads_as(self._sc_ins.get_gravity_nodal_loads(existing_ids), existing_ids, dof_flattened)

    def get_nodal_loads(self, existing_ids=[], dof_flattened=False):
        return self._loads_as(self._sc_ins.get_lumped_nodal_loads(existing_ids), existing_ids, dof_flattened)

    def get_element_stiffness_matrices(self, in_local_coordinate=False):
        Ks = self._sc_ins.get_element_stiffness_matrices()
        if not in_local_coordinate:
            return {e: K for e, K in enumerate(Ks)}
        Rs = self._sc_ins.get_element_local2global_rot_matrices()
        return {e: Rs[e].dot(K).dot(Rs[e].T) for e, K in enumerate(Ks)}

    def get_global_stiffness_matrix(self, exist_element_ids=[]):
        return self._sc_ins.compute_full_stiffness_matrix(exist_element_ids)

    def get_dof_permutation_matrix(self, exist_element_ids=[]):
        Perm, (_, n_free_dof, n_fixed_dof, _) = self._sc_ins.compute_dof_permutation(exist_element_ids)
        return Perm, (n_free_dof, n_fixed_dof)

    def get_element_local2global_rot_matrices(self):
        return {e: R for e, R in enumerate(self._sc_ins.get_element_local2global_rot_matrices())}

    def get_element2dof_id_map(self):
        return {int(e): {0: m[0:6], 1: m[6:12]} for e, m in enumerate(self._sc_ins.get_element2dof_id_map())}

    def get_node2dof_id_map(self):
        return {int(n): m for n, m in enumerate(self._sc_ins.get_node2dof_id_map())}

    def get_element_crosssec(self, elem_id):
        return self._sc_ins.get_element_crosssec(elem_id)

    def get_element_material(self, elem_id):
        return self._sc_ins.get_element_material(elem_id)

    # ---------------------------------------------------------------- result export
    @property
    def result_data(self):
        return self.get_result_data(False)

    def get_result_data(self, output_frame_transf=False):
        """Result dictionary in the reference's JSON layout (``stiffness_checker.py:609-665``)."""
        if not self.has_stored_result():
            print('no result to output!')
            return None
        success, nD, fR, eR = self.get_solved_results()
        trans_tol, rot_tol = self.get_nodal_deformation_tol()
        max_trans, max_rot, max_t_id, max_r_id = self.get_max_nodal_deformation()
        data = OrderedDict()
        data['model_name'] = self.model_name
        data['solve_success'] = bool(success)
        data['trans_tol'] = trans_tol
        data['rot_tol'] = rot_tol
        data['length_unit'] = 'meter'
        data['angle_unit'] = 'rad'
        data['force_unit'] = 'kN'
        data['moment_unit'] = 'kN-m'
        data['compliance'] = self.get_compliance()
        data['max_trans'] = {'node_id': int(max_t_id), 'value': max_trans}
        data['max_rot'] = {'node_id': int(max_r_id), 'value': max_rot}
        pts = self.node_points
        data['node_displacement'] = {
            n: OrderedDict([('node_id', n), ('node_pose', list(pts[n])), ('displacement', d.tolist())])
            for n, d in nD.items()}
        rots = self.get_element_local2global_rot_matrices() if output_frame_transf else None
        er_data = {}
        for e, er in eR.items():
            d = OrderedDict([('element_id', e), ('node_ids', self.elements[e]),
                             ('reaction', OrderedDict([(0, er[0].tolist()), (1, er[1].tolist())]))])
            if output_frame_transf:
                d['local_to_global_transformation'] = rots[e][:3, :3].tolist()
            er_data[e] = d
        data['element_reaction'] = er_data
        data['fixity_reaction'] = {
            v: OrderedDict([('node_id', v), ('node_pose', list(pts[v])), ('reaction', r.tolist())])
            for v, r in fR.items()}
        return data

    def write_result_to_json(self, file_path, formatted_json=False):
        with open(file_path, 'w+') as f:
            json.dump(self.result_data, f, indent=4 if formatted_json else None)

    # ---------------------------------------------------------------- graph queries
    def set_output_json(self, do_output=True, output_dir=None, file_name=None):
        """DEPRECATED in the reference (``stiffness_checker.py:589-603``); use ``write_result_to_json``."""
        raise NotImplementedError('deprecated in the reference; use write_result_to_json(file_path)')

    def get_element_local_node_id(self, element_id, point):
        """0 / 1 if ``point`` is end 0 / end 1 of the element, else None.  (The reference's
        version, ``stiffness_checker.py:678-692``, calls an undefined ``norm``; this one compares
        coordinates within 1e-9.)"""
        u, v = self.elements[element_id]
        p = np.asarray(point, dtype=float)
        if np.linalg.norm(p - np.asarray(self.node_points[u])) < 1e-9:
            return 0
        if np.linalg.norm(p - np.asarray(self.node_points[v])) < 1e-9:
            return 1
        return None

    def get_node_neighbors(self, return_e_id=False):
        nb = defaultdict(set)
        for e_id, e in enumerate(self.elements):
            for n in e:
                nb[n].add(e_id if return_e_id else tuple(e))
        return nb

    def get_element_neighbors(self):
        nb = self.get_node_neighbors()
        out = defaultdict(set)
        for e in self.elements:
            key = tuple(e)
            out[key].update(nb[e[0]])
            out[key].update(nb[e[1]])
            out[key].discard(key)
        return out

    def get_element_connected_node_ids(self, existing_ids=[], fix_node_only=False):
        if not existing_ids:
            existing_ids = range(len(self.elements))
        fixed = set(self.fix_node_ids)
        ids = set()
        for e_id in existing_ids:
            for v in self.elements[e_id]:
                if not fix_node_only or v in fixed:
                    ids.add(v)
        return list(ids)

    def get_original_shape(self, disc=1, draw_full_shape=True):
        raise NotImplementedError()

    def get_deformed_shape(self, disc=10, exagg_ratio=1.0):
        raise NotImplementedError()
