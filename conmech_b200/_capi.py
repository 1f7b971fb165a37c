"""ctypes binding of libconmech_b200.so (include/conmech_b200.h).  No fallback: if the CUDA
library is missing or fails to load, importing the engine raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CONMECH_B200_LIB: another build of the same library (A/B timing of kernel variants); never a fallback
LIB_PATH = os.environ.get('CONMECH_B200_LIB') or os.path.join(_HERE, 'lib', 'libconmech_b200.so')

CM_OK = 0
ST_OK, ST_NO_SUPPORT, ST_NO_FIXED_DOF, ST_SINGULAR, ST_ZERO_LOAD = 0, 1, 2, 3, 4

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint8_p = C.POINTER(C.c_uint8)
c_uint32_p = C.POINTER(C.c_uint32)


class ModelDesc(C.Structure):
    _fields_ = [('n_nodes', C.c_int32), ('n_elems', C.c_int32), ('n_supports', C.c_int32),
                ('device', C.c_int32), ('xyz', c_double_p), ('conn', c_int32_p), ('props', c_double_p),
                ('cr', c_double_p), ('sup_node', c_int32_p), ('sup_cond', c_uint8_p), ('node_order', c_int32_p)]


class ModelInfo(C.Structure):
    _fields_ = [('n_nodes', C.c_int32), ('n_elems', C.c_int32), ('n_supports', C.c_int32),
                ('n_loadcases', C.c_int32), ('mask_words', C.c_int32), ('n_free_full', C.c_int32),
                ('half_bandwidth', C.c_int32), ('tile', C.c_int32), ('n_tile_rows_max', C.c_int32),
                ('band_tiles', C.c_int32), ('profile_flops_full', C.c_int64)]


OUT_FIELDS = ('flag', 'status', 'n_free', 'n_fixed', 'dof_map', 'node_exists', 'U', 'Rf', 'eR',
              'compliance', 'max_trans', 'max_rot', 'min_pivot', 'profile_flops', 'sigma_max', 'max_abs_sigma',
              'profile_entries', 'envelope_tiles', 'Rf_sup', 'eR_rows', 'eR_row_offsets')


class SolveOut(C.Structure):
    _fields_ = [(f, C.c_uint64) for f in OUT_FIELDS]


class SolveOutHost(C.Structure):
    _fields_ = [(f, C.c_void_p) for f in OUT_FIELDS]


# every symbol include/conmech_b200.h declares
SYMBOLS = ('cm_pack_elements', 'cm_create', 'cm_destroy', 'cm_get_info', 'cm_get_element_tables', 'cm_get_node_order',
           'cm_set_geometries', 'cm_geometry_count', 'cm_get_geometry_tables', 'cm_last_element_kernel_ms',
           'cm_solve_many_g', 'cm_solve_many_host_g', 'cm_solve_many_host_seq',
           'cm_set_loadcases', 'cm_get_element_loads', 'cm_set_tol', 'cm_set_variant', 'cm_debug_k3_phase_cycles', 'cm_debug_k2_phase_cycles',
           'cm_workspace_bytes', 'cm_solve_many', 'cm_solve_many_host', 'cm_solve_many_host_chunks', 'cm_solve_many_host_multi', 'cm_kernel_launches',
           'cm_reset_counters', 'cm_enable_timing', 'cm_last_stage_ms', 'cm_last_wave_count', 'cm_element_kernel_dev',
           'cm_element_row_offsets', 'cm_last_error', 'cm_version')

_lib = None


def load():
    """Load the library once; raise if it is not there (the product path has no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError('conmech_b200: CUDA library {} not built - run `python -m conmech_b200.build` '
                           '(there is no CPU fallback)'.format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.cm_create.argtypes = [C.POINTER(ModelDesc), C.POINTER(H)]
    lib.cm_destroy.argtypes = [H]
    lib.cm_get_info.argtypes = [H, C.POINTER(ModelInfo)]
    lib.cm_get_element_tables.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cm_get_node_order.argtypes = [H, C.c_void_p]
    lib.cm_pack_elements.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                     C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cm_set_geometries.argtypes = [H, C.c_int32, C.c_void_p]
    lib.cm_geometry_count.argtypes = [H]
    lib.cm_geometry_count.restype = C.c_int32
    lib.cm_get_geometry_tables.argtypes = [H, C.c_int32, C.c_void_p, C.c_void_p]
    lib.cm_last_element_kernel_ms.argtypes = [H, C.c_void_p]
    lib.cm_solve_many_g.argtypes = [H, C.c_uint64, C.c_uint64, C.c_int64, C.POINTER(SolveOut), C.c_uint64, C.c_size_t, C.c_uint64]
    lib.cm_solve_many_host_g.argtypes = [H, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(SolveOutHost)]
    lib.cm_solve_many_host_seq.argtypes = [H, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.POINTER(SolveOutHost)]
    lib.cm_set_loadcases.argtypes = [H, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cm_get_element_loads.argtypes = [H, C.c_int32, C.c_void_p]
    lib.cm_set_tol.argtypes = [H, C.c_double, C.c_double]
    lib.cm_set_variant.argtypes = [H, C.c_int32]
    lib.cm_debug_k3_phase_cycles.argtypes = [H, C.c_void_p]
    lib.cm_debug_k2_phase_cycles.argtypes = [H, C.c_void_p]
    lib.cm_workspace_bytes.argtypes = [H, C.c_int64]
    lib.cm_workspace_bytes.restype = C.c_size_t
    lib.cm_solve_many.argtypes = [H, C.c_uint64, C.c_int64, C.POINTER(SolveOut), C.c_uint64, C.c_size_t, C.c_uint64]
    lib.cm_solve_many_host.argtypes = [H, C.c_void_p, C.c_int64, C.POINTER(SolveOutHost)]
    lib.cm_solve_many_host_chunks.argtypes = [H, C.c_void_p, C.c_int64, C.POINTER(SolveOutHost), C.c_int64, C.c_void_p]
    lib.cm_solve_many_host_multi.argtypes = [C.POINTER(H), C.c_int32, C.c_void_p, C.c_int64, C.POINTER(SolveOutHost)]
    lib.cm_kernel_launches.argtypes = [H]
    lib.cm_kernel_launches.restype = C.c_int64
    lib.cm_reset_counters.argtypes = [H]
    lib.cm_enable_timing.argtypes = [H, C.c_int32]
    lib.cm_last_stage_ms.argtypes = [H, C.c_void_p]
    lib.cm_last_wave_count.argtypes = [H]
    lib.cm_element_row_offsets.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    lib.cm_element_kernel_dev.argtypes = [C.c_uint64] * 4 + [C.c_int64] + [C.c_uint64] * 3
    lib.cm_last_error.restype = C.c_char_p
    lib.cm_version.restype = C.c_char_p
    for name in SYMBOLS:
        getattr(lib, name)          # AttributeError if the .so is stale
        if getattr(lib, name).restype is C.c_int:
            pass
    _lib = lib
    return lib


class CmError(RuntimeError):
    pass


def check(rc):
    if rc != CM_OK:
        raise CmError('conmech_b200 error {}: {}'.format(rc, load().cm_last_error().decode()))
