"""One-call analysis from JSON-style data, mirroring the reference's
``pyconmech.frame_analysis.utils.solve_linear_elastic_from_data`` (utils.py:6-64) on the CUDA
engine, plus a batched variant for many element subsets of the same model."""
import numpy as np

from .io_base import Model, LoadCase
from .stiffness_checker import StiffnessChecker


def solve_linear_elastic_from_data(model_data, loadcase_data, verbose=True, checker_engine='cuda', device=None):
    """Same inputs and result dictionary as the reference function; the engine is the CUDA one."""
    model = Model.from_data(model_data)
    loadcase = LoadCase.from_data(loadcase_data)
    sc = StiffnessChecker(model, checker_engine=checker_engine, verbose=verbose, device=device)
    sc.set_loads(loadcase)
    sc.set_nodal_displacement_tol(trans_tol=np.inf, rot_tol=np.inf)
    success = sc.solve()
    success, nD, fR, eR = sc.get_solved_results()
    max_trans, max_rot, max_trans_vid, max_rot_vid = sc.get_max_nodal_deformation()
    sigma_max = sc.get_sigma_max_per_element()
    compliance = sc.get_compliance()
    if verbose:
        print('Solve success: {}'.format(success))
        print('Max translation deformation: {0:.5f} m at node #{1}'.format(max_trans, max_trans_vid))
        print('Max rotation deformation: {0:.5f} rad at node #{1}'.format(max_rot, max_rot_vid))
        print('Elastic energy: {} [kN m]'.format(compliance))
    return {
        'solve_success': bool(success),
        'elastic_energy': float(compliance),
        'nodal_displacement': {nid: list(nd) for nid, nd in nD.items()},
        'support_reaction': {nid: list(fr) for nid, fr in fR.items()},
        'element_reaction': {eid: [list(er[0]), list(er[1])] for eid, er in eR.items()},
        'max_trans': max_trans,
        'max_trans_nid': int(max_trans_vid),
        'sigma_max': sigma_max,
    }


def solve_many_from_data(model_data, loadcase_data, list_of_existing_ids, trans_tol=1e-3, device=None):
    """Batched variant: flags and per-subset summaries for many element subsets of one model."""
    model = Model.from_data(model_data)
    sc = StiffnessChecker(model, checker_engine='cuda', device=device)
    sc.set_loads(LoadCase.from_data(loadcase_data))
    sc.set_nodal_displacement_tol(trans_tol=trans_tol)
    flags, det = sc.solve_many(list_of_existing_ids, return_details=True)
    return {'solve_success': flags, 'max_trans': det['max_trans'].tolist(),
            'elastic_energy': det['compliance'].tolist(), 'status': det['status'].tolist()}
