"""Engine interface shared by all stiffness back ends (the plug-in boundary).

Mirror of the reference's ``StiffnessBase`` (``src/pyconmech/frame_analysis/stiffness_base.py:8-155``):
the facade only ever talks to an object with these methods.  The tag -> section / material /
joint lookups live here because every engine needs them (``stiffness_base.py:131-156``).
"""
import warnings

from .io_base import Model


class StiffnessBase(object):
    def __init__(self, model, verbose=False, output_json=False):
        self._nodes = model.nodes
        self._elements = model.elements
        self._supports = model.supports            # dict keyed by node index, model order
        self._joints = model.joints
        self._joint_id_from_etag = model.joint_id_from_etag
        self._materials = model.materials
        self._material_id_from_etag = model.material_id_from_etag
        self._crosssecs = model.crosssecs
        self._crosssec_id_from_etag = model.crosssec_id_from_etag
        self._verbose = verbose
        self._output_json = output_json
        self._trans_tol = None
        self._rot_tol = None
        if len(self._supports) == 0:
            raise RuntimeError('there needs to be at least one support (fixed) vertex in the model!')

    @classmethod
    def from_json(cls, json_file_path=None, verbose=False):
        return cls(Model.from_json(json_file_path, verbose=verbose), verbose=verbose)

    # ---- element attributes by tag, falling back to the untagged (None) entry
    def _by_tag(self, elem_id, table, items, what, required=True):
        tag = self._elements[elem_id].elem_tag
        if tag in table:
            return items[table[tag]]
        if not required:
            return None
        warnings.warn('No {} assigned for element tag |{}|, using the default tag'.format(what, tag))
        return items[table[None]]

    def get_element_crosssec(self, elem_id):
        return self._by_tag(elem_id, self._crosssec_id_from_etag, self._crosssecs, 'cross section')

    def get_element_material(self, elem_id):
        return self._by_tag(elem_id, self._material_id_from_etag, self._materials, 'material')

    def get_element_joint(self, elem_id):
        return self._by_tag(elem_id, self._joint_id_from_etag, self._joints, 'joint', required=False)

    # ---- the interface an engine implements
    def _missing(self, *a, **k):
        raise NotImplementedError()

    set_self_weight_load = set_load = set_uniformly_distributed_loads = _missing
    set_nodal_displacement_tol = solve = _missing
    set_output_json_path = set_output_json = _missing
    get_frame_stat = get_lumped_nodal_loads = get_gravity_nodal_loads = _missing
    get_element_stiffness_matrices = get_element_local2global_rot_matrices = _missing
    get_element2dof_id_map = get_node2dof_id_map = _missing
    has_stored_result = get_solved_results = get_max_nodal_deformation = _missing
    get_nodal_deformation_tol = get_original_shape = get_deformed_shape = _missing
