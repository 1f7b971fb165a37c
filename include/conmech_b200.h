/*
 * conmech_b200 - C ABI of the B200-native partial-structure stiffness solver.
 *
 * This is the drop-in boundary for the reference's engine plug-in point
 * (reference src/pyconmech/frame_analysis/stiffness_checker.py:57-66 selects an engine
 * object; src/pyconmech/frame_analysis/stiffness_base.py:8-155 is the interface that
 * object implements; src/pyconmech_bindings/pybind_stiffness_checker.cpp:19-243 is the
 * reference's existing native binding of the same surface).  The Python host class
 * conmech_b200.CudaStiffness binds these entry points with ctypes; INTEGRATION.md shows
 * the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; every function returns 0 (CM_OK) or a negative CM_ERR_* code and
 *     leaves a message retrievable with cm_last_error() (thread-local).  No C++ exception
 *     crosses the boundary.  Numerical failure of a solve is NOT an error: it is reported
 *     per subset in `status` / `flag`, like the reference returns False
 *     (numpy_stiffness.py:238-240, 301-319, 326-336).
 *   - "dev" pointers are CUDA device addresses passed as uint64_t (what torch's
 *     Tensor.data_ptr() returns); "host" pointers are ordinary pointers.
 *   - the caller owns every input/output buffer and the workspace; the library owns only the
 *     opaque handle (model tables, load tables, small scratch) on the device it was created on.
 *   - one handle per device; a handle is not thread-safe; different handles (on the same or on
 *     different devices) may be used concurrently from different host threads.
 *   - all floating point is IEEE fp64; indices are int32; DOF numbering is 6*node + {0..5}
 *     (numpy_stiffness.py:563-570).
 */
#ifndef CONMECH_B200_H
#define CONMECH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CM_OK 0
#define CM_ERR_INVALID (-1)   /* bad argument / inconsistent model                     */
#define CM_ERR_CUDA (-2)      /* CUDA runtime error, message has the CUDA error string */
#define CM_ERR_NOMEM (-3)     /* workspace too small / allocation failed               */
#define CM_ERR_STATE (-4)     /* call order (e.g. solve before cm_set_loadcases)       */

/* per (subset, load case) status, mirrors the reference's routes to `return False`
 *  */
#define CM_ST_OK 0            /* solved; flag = max_trans < trans_tol                              */
#define CM_ST_NO_SUPPORT 1    /* no support node in the subset (numpy_stiffness.py:166-168)        */
#define CM_ST_NO_FIXED_DOF 2  /* existing supports fix zero DOFs (numpy_stiffness.py:186-189)      */
#define CM_ST_SINGULAR 3      /* zero / negative (<= -1e-14) / NaN pivot (numpy_stiffness.py:301-336) */
#define CM_ST_ZERO_LOAD 4     /* ||P|| <= 1e-14: solve skipped, U = 0, flag = 0 < tol (:267,360-363) */

typedef struct cm_handle cm_handle;

/* Frame description, host pointers.  Replaces the constructor arguments of the reference's
 * native engine (pybind_stiffness_checker.cpp:37-54: V, E, Fixities, materials) and the
 * per-element lookups of NumpyStiffness._precompute_element_stiffness_matrices
 * (numpy_stiffness.py:579-614). */
typedef struct cm_model_desc {
    int32_t n_nodes;
    int32_t n_elems;
    int32_t n_supports;
    int32_t device;          /* CUDA device ordinal                                              */
    const double *xyz;       /* [n_nodes*3] metres                                               */
    const int32_t *conn;     /* [n_elems*2] end node ids                                         */
    const double *props;     /* [n_elems*7] A, Jx, Iy, Iz, E, mu, density                        */
    const double *cr;        /* [n_elems*6] rotational end springs (x,y,z at end 0, then end 1): */
                             /*   +inf rigid, 0 released; all 0 for bending_stiff=false          */
    const int32_t *sup_node; /* [n_supports] support node ids in model (JSON) order              */
    const uint8_t *sup_cond; /* [n_supports*6] 1 = fixed, after the all-truss-neighbours         */
                             /*   rotational pin rule (numpy_stiffness.py:179-181)               */
    const int32_t *node_order; /* [n_nodes] solver ordering, position -> node id (a permutation), or  */
                             /*   NULL: the library picks the better of natural and reverse        */
                             /*   Cuthill-McKee (by envelope work of the full structure).            */
} cm_model_desc;

/* Static facts about a created model (sizes a caller needs to allocate outputs). */
typedef struct cm_model_info {
    int32_t n_nodes, n_elems, n_supports, n_loadcases;
    int32_t mask_words;        /* uint32 words per subset mask = ceil(n_elems/32)                */
    int32_t n_free_full;       /* free DOFs of the full structure                                */
    int32_t half_bandwidth;    /* of the full structure in the static solver ordering            */
    int32_t tile;              /* tile edge of the block-band storage (32)                       */
    int32_t n_tile_rows_max;   /* ceil(n_free_full / tile)                                       */
    int32_t band_tiles;        /* tiles kept per tile row (half bandwidth in tiles + 1)          */
    int64_t profile_flops_full;/* sum_j h_j^2 of the full structure in the solver ordering       */
} cm_model_info;

/* Per-element property arrays of cm_model_desc from per-tag tables (host arrays; no GPU involved):
 * the C side of the reference's per-element Python loop (numpy_stiffness.py:579-614, tag lookups
 * stiffness_base.py:131-156 done by the caller once per TAG).  sec_of / mat_of [n_elems] index into
 * sections [n_sec*4] (A, Jx, Iy, Iz) / materials [n_mat*3] (E, mu, density); joint_of [n_elems] (may be
 * NULL) indexes joints [n_joint*6] (rotational end springs x,y,z at end 0 then end 1, +inf = rigid) or
 * is -1 = no joint; bending_stiff [n_elems] (may be NULL = all true).  Writes props [n_elems*7] and
 * cr [n_elems*6] as cm_model_desc wants them. */
int cm_pack_elements(int32_t n_elems, const int32_t *sec_of, const int32_t *mat_of, const int32_t *joint_of,
                     const uint8_t *bending_stiff, int32_t n_sec, const double *sections, int32_t n_mat,
                     const double *materials, int32_t n_joint, const double *joints, double *props, double *cr);

int cm_create(const cm_model_desc *desc, cm_handle **out);
int cm_destroy(cm_handle *h);
int cm_get_info(const cm_handle *h, cm_model_info *info);

/* Element tables built on the device by the element kernel; copied to host arrays.
 * Replaces get_element_stiffness_matrices / get_element_local2global_rot_matrices /
 * get_element2dof_id_map (numpy_stiffness.py:503-513).  Any pointer may be NULL. */
int cm_get_element_tables(cm_handle *h, double *Kg /*[nE*144]*/, double *R3 /*[nE*9]*/,
                          int32_t *id_map /*[nE*12]*/);

/* Static solver ordering (position -> node id), host array [n_nodes]. */
int cm_get_node_order(const cm_handle *h, int32_t *order);

/* Geometry variants (many models of one topology, e.g. the iterates of a shape
 * optimisation).  Replaces constructing one reference engine per shape (Model.from_data io_base.py:107-144
 * + NumpyStiffness._precompute_element_stiffness_matrices numpy_stiffness.py:579-614, a Python loop over the
 * elements per model): xyz holds n_geom sets of node coordinates [n_geom*n_nodes*3] (host, metres);
 * connectivity, sections, materials, joints and supports are those of the handle.  ONE launch of the
 * element kernel builds the stiffness / rotation tables of all n_geom*n_elems elements (HBM-bound for
 * many variants); every subset of a later cm_solve_many*_g call picks its variant by index.  The
 * variants replace the handle's geometry (variant 0 need not be the original); the load cases must be
 * set again afterwards (lumped element loads depend on member lengths and axes). */
int cm_set_geometries(cm_handle *h, int32_t n_geom, const double *xyz);
int32_t cm_geometry_count(const cm_handle *h);
/* element tables of variant g, host arrays (any may be NULL) */
int cm_get_geometry_tables(cm_handle *h, int32_t g, double *Kg /*[nE*144]*/, double *R3 /*[nE*9]*/);
/* device milliseconds (CUDA events) of the most recent element-table launch (cm_create / cm_set_geometries) */
int cm_last_element_kernel_ms(const cm_handle *h, double *ms);

/* Load cases, host pointers.  Replaces set_load / set_uniformly_distributed_loads /
 * set_self_weight_load (numpy_stiffness.py:98-138); n_lc load cases are solved per subset.
 *   nodal   [n_lc*6*n_nodes]  point loads already scattered to DOFs
 *   w_udl   [n_lc*n_elems*3]  uniform line load (global axes) per element, used where has_udl
 *   has_udl [n_lc*n_elems]    may be NULL (= none)
 *   grav_on [n_lc]            self weight on/off; grav_dir [n_lc*3] (not normalised) */
int cm_set_loadcases(cm_handle *h, int32_t n_lc, const double *nodal, const double *w_udl,
                     const uint8_t *has_udl, const uint8_t *grav_on, const double *grav_dir);

/* Lumped element loads of one load case as the solver uses them, host array [nE*12]
 * (sum of UDL and self-weight parts, numpy_stiffness.py:401-430). */
int cm_get_element_loads(cm_handle *h, int32_t lc, double *q);

/* set_nodal_displacement_tol (numpy_stiffness.py:142-145). rot_tol is stored, unused. */
int cm_set_tol(cm_handle *h, double trans_tol, double rot_tol);

/* Output selection for cm_solve_many: device addresses, 0 = not wanted.
 * B = subsets, L = load cases, nV/nE = nodes/elements. */
typedef struct cm_solve_out {
    uint64_t flag;        /* u8  [B*L]      success flag (numpy_stiffness.py:441)                 */
    uint64_t status;      /* u8  [B*L]      CM_ST_*                                               */
    uint64_t n_free;      /* i32 [B]                                                              */
    uint64_t n_fixed;     /* i32 [B]                                                              */
    uint64_t dof_map;     /* i32 [B*6nV]    the reference's id_map_RO (numpy_stiffness.py:193-208) */
    uint64_t node_exists; /* u8  [B*nV]                                                           */
    uint64_t U;           /* f64 [B*L*6nV]  nodal displacements, DOF order                        */
    uint64_t Rf;          /* f64 [B*L*6nV]  fixity reactions at fixed DOFs, 0 elsewhere           */
    uint64_t eR;          /* f64 [B*L*nE*12] element end forces, local axes; 0 rows when absent   */
    uint64_t compliance;  /* f64 [B*L]                                                            */
    uint64_t max_trans;   /* f64 [B*L]                                                            */
    uint64_t max_rot;     /* f64 [B*L]                                                            */
    uint64_t min_pivot;   /* f64 [B]        smallest LDL^T pivot of the scaled matrix             */
    uint64_t profile_flops; /* f64 [B]      sum_j h_j^2 of the subset's K_mm envelope (algorithmic work  */
                          /*                of the factorisation, one multiply-add counted as 2)  */
    uint64_t sigma_max;   /* f64 [B*L*nE]   max axial normal stress per element, the reference's          */
                          /*                get_sigma_max_per_element (stiffness_checker.py:330-365: solid */
                          /*                round sections, both section moduli from Iy); 0 when absent    */
    uint64_t max_abs_sigma; /* f64 [B*L]    max over the subset's elements of |sigma_max|                  */
    uint64_t profile_entries; /* f64 [B]    sum_j (h_j + 1): entries of the scalar profile (lower, with diagonal)  */
                          /*                of the subset's K_mm; 8x this = S_K of SURVEY.md section 8d (profile)  */
    uint64_t envelope_tiles; /* f64 [B]     32x32 tiles of the block-band envelope that is factorised (8 KB each)  */
    /* Compact results - what the reference returns, without the zero padding of the dense arrays above:
     * get_solved_results() gives fR rows for the support nodes and eR rows for the existing elements only
     * (numpy_stiffness.py:375-391).  On the 1k-element frame a dense result is 127 KB per solve, the compact
     * one (U dense + these two) about 70 KB: the host entry point is bound by the device->host link. */
    uint64_t Rf_sup;      /* f64 [B*L*nS*6] fixity reactions of the support nodes in the order of cm_model_desc.sup_node */
                          /*                (the reference's fR rows); 0 at free DOFs and for absent support nodes      */
    uint64_t eR_rows;     /* f64 [rows*L*12] element end forces of the EXISTING elements only: subset b owns rows      */
                          /*                off[b] .. off[b+1] (ascending element id), stored [L][off[b+1]-off[b]][12]  */
                          /*                at eR_rows + (off[b] - off[0])*L*12                                          */
    uint64_t eR_row_offsets; /* INPUT, i64 [B+1]: off[b] = number of existing elements of the subsets before b         */
                          /*                (exclusive prefix sum of the popcounts of the masks); required with        */
                          /*                eR_rows; the device entry points want off[0] == 0                           */
} cm_solve_out;

/* Factorisation kernel variant: 1 (default) = FP64 tensor-core block-band LDL^T, dataflow
 * kernel (csrc/k3_factor_flow.cu), CTA size chosen per launch (4 warps x 4 CTAs/SM when the
 * launch fills the GPU, 8 or 16 warps per subset for few subsets of a wide band); 2 / 3 / 6 = the
 * same kernel pinned to 4 / 8 / 16-warp CTAs; 4 / 5 = 4 / 8-warp CTAs instrumented with
 * per-activity cycle counters; 7 / 8 / 9 = 2 / 3 / 1-warp CTAs (timing experiments, parity-clean);
 * 10 / 11 = 4-warp CTAs whose update-stream operands come through a shared-memory ring filled by the
 * bulk-copy engine (cp.async.bulk + mbarrier, one / two stages of two k-steps per warp, three CTAs per
 * SM), 12 = the same with 8-warp CTAs for wide bands - parity-clean, measured slower than the register
 * path at its higher occupancy, kept selectable; 13 = 4-warp CTAs that park a finished
 * accumulator tile in TENSOR MEMORY (tcgen05.alloc / st / ld, 128 columns per CTA) when M_J is not published yet
 * and run the available terms of the row's next tile meanwhile - parity-clean, measured slower;
 * 0 = scalar validation variant of the same algorithm
 * (csrc/k3_factor.cu; slow, used by the parity tests to cross-check the tensor-core path). */
int cm_set_variant(cm_handle *h, int32_t variant);

/* Development aid: after a solve with variant 4 or 5, reads and clears the 16 activity counters
 * (clock64 cycles summed over all row warps; slot meanings in csrc/k3_factor_flow.cu). */
int cm_debug_k3_phase_cycles(cm_handle *h, uint64_t *out16);
int cm_debug_k2_phase_cycles(cm_handle *h, uint64_t *out16);  /* -DK2_PROF=1 builds of the assembly kernel: cycles per phase */

/* Bytes of workspace needed to run `subsets_in_flight` subsets at once. cm_solve_many splits
 * B into waves that fit the workspace it is given. */
size_t cm_workspace_bytes(const cm_handle *h, int64_t subsets_in_flight);

/* The hot path.  Replaces NumpyStiffness.solve (numpy_stiffness.py:228-397) for B element
 * subsets at once.  masks_dev: u32 [B*mask_words], bit e%32 of word e/32 set = element e
 * exists.  Everything is enqueued on `stream` (a cudaStream_t as integer, 0 = default
 * stream); the call returns without synchronising. */
int cm_solve_many(cm_handle *h, uint64_t masks_dev, int64_t B, const cm_solve_out *out,
                  uint64_t workspace_dev, size_t workspace_bytes, uint64_t stream);

/* cm_solve_many with a geometry variant per subset: geom_dev = i32 [B] device array of variant indices
 * (0 = all subsets use variant 0; out-of-range indices are clamped - the host entry below rejects them). */
int cm_solve_many_g(cm_handle *h, uint64_t masks_dev, uint64_t geom_dev, int64_t B, const cm_solve_out *out,
                    uint64_t workspace_dev, size_t workspace_bytes, uint64_t stream);

/* Same, with HOST buffers: masks and every non-NULL output are host arrays (pinned memory lets
 * the copies overlap the kernels; pageable memory works, serialised by the driver).  The call
 * copies the masks host->device, runs the kernels wave by wave and drains every finished wave's
 * outputs to the caller's arrays on a copy stream while the next waves compute; it returns when
 * all outputs are complete.  It manages its own workspace and a bounded device staging ring
 * (a few waves of outputs, not the whole batch). */
typedef struct cm_solve_out_host {
    uint8_t *flag;
    uint8_t *status;
    int32_t *n_free;
    int32_t *n_fixed;
    int32_t *dof_map;
    uint8_t *node_exists;
    double *U;
    double *Rf;
    double *eR;
    double *compliance;
    double *max_trans;
    double *max_rot;
    double *min_pivot;
    double *profile_flops;
    double *sigma_max;
    double *max_abs_sigma;
    double *profile_entries;
    double *envelope_tiles;
    double *Rf_sup;                 /* compact results, see cm_solve_out */
    double *eR_rows;
    const int64_t *eR_row_offsets;  /* INPUT, host i64 [B+1] */
} cm_solve_out_host;

int cm_solve_many_host(cm_handle *h, const uint32_t *masks_host, int64_t B,
                       const cm_solve_out_host *out);
/* Row offsets of the compact element rows (cm_solve_out.eR_row_offsets) from a host mask matrix:
 * offsets[b] = number of existing elements of the subsets before b, offsets[B] = all rows (host i64 [B+1]). */
int cm_element_row_offsets(const uint32_t *masks_host, int64_t B, int32_t mask_words, int64_t *offsets);
/* ... with a geometry variant per subset (geom_host i32 [B], NULL = variant 0; an index outside
 * [0, cm_geometry_count) is CM_ERR_INVALID) */
int cm_solve_many_host_g(cm_handle *h, const uint32_t *masks_host, const int32_t *geom_host, int64_t B,
                         const cm_solve_out_host *out);

/* Construction sequences (callers: the sequencing loops sketched at reference
 * README.rst:39-42 and examples/notebook_demo/demo.ipynb, which call solve(order[:k]) for k = 1..N and
 * re-assemble and re-factor everything N times).  The B subsets are consecutive PREFIXES of one
 * construction sequence; the batch is cut into groups of `group` consecutive prefixes.  The solver numbers
 * the existing nodes of a subset in the handle's static ordering, so a later prefix differs from the first
 * prefix of its group (the "trunk") only at and behind the first solver row that an added element touches or
 * in front of which a new node is inserted.  The trunk is factored completely; prefix b shares the first
 * t_rows[b] 32-row tile rows of the factor (L, the inverted diagonal factors, pivots, forward solutions)
 * with it: those rows are not factored again, and the back substitution reads them from the trunk.
 * t_rows[b] = floor(first differing solver row / 32) in the trunk's numbering, 0 for the first prefix of
 * every group (host array [B]; CudaStiffness.solve_sequence computes it from the node ordering of
 * cm_get_node_order).  A prefix whose trunk cannot be used (not solvable, different Jacobi-scaling
 * decision, bad pivot inside the shared rows) is factored on its own.  Results are those of
 * cm_solve_many_host on the same masks: every row of the factor is the same arithmetic in the same
 * order, whether computed or shared. */
int cm_solve_many_host_seq(cm_handle *h, const uint32_t *masks_host, int64_t B, int32_t group,
                           const int32_t *t_rows_host, const cm_solve_out_host *out);

/* The same batch in consecutive chunks of `chunk` subsets, for a caller that is still PRODUCING the
 * masks while the first chunks are solved (the Python facade packs id lists - the reference's
 * calling convention, StiffnessChecker.solve(exist_element_ids) stiffness_checker.py:153-182 - into
 * masks chunk by chunk on another thread).  Before chunk k the call waits, without any interpreter
 * lock, until ready[k] != 0 (producer: write the chunk's masks, then set ready[k] = 1; a negative
 * value cancels the call, which then returns CM_ERR_STATE).  ready == NULL: no waiting.  Outputs
 * are written at the subset's position in the full arrays, exactly as cm_solve_many_host does. */
int cm_solve_many_host_chunks(cm_handle *h, const uint32_t *masks_host, int64_t B,
                              const cm_solve_out_host *out, int64_t chunk, const volatile int32_t *ready);

/* One process, several GPUs (reference: none - the reference is single-threaded,
 * stiffness_base.py:8-155): `handles` are n_handles handles created from the
 * SAME model on DIFFERENT devices, with the same load cases and tolerances set on each.  The batch is
 * range-partitioned over them (handle g gets the g-th contiguous share, the first B % n_handles
 * shares one subset longer), one host thread per device runs its share like cm_solve_many_host, and
 * every result is written at the subset's position in the caller's arrays - the host gather of the
 * partitioned solve; no collective, no second copy.  Returns the first failing device's error. */
int cm_solve_many_host_multi(cm_handle *const *handles, int32_t n_handles, const uint32_t *masks_host,
                             int64_t B, const cm_solve_out_host *out);

/* Counters since the last cm_reset_counters: kernels launched by this library, and device
 * milliseconds (CUDA events on the launching stream) per stage of the last cm_solve_many*
 * call when timing was enabled with cm_enable_timing(h, 1). stage: 0 dofmap+assembly,
 * 1 factorisation(+forward), 2 back substitution + recovery. */
int64_t cm_kernel_launches(const cm_handle *h);
int cm_reset_counters(cm_handle *h);
int cm_enable_timing(cm_handle *h, int32_t on);
int cm_last_stage_ms(const cm_handle *h, double *ms /*[3]*/);
/* Number of waves (K2 -> K3 -> K4 launch triples) the most recent cm_solve_many* call was cut
 * into.  Outside timing mode consecutive waves rotate over three internal streams forked
 * from and joined to the caller's stream, each with its own third of the workspace. */
int cm_last_wave_count(const cm_handle *h);

/* Stand-alone element kernel for many elements (roofline measurement of the element kernel,
 * SURVEY.md section 8d): device arrays, n elements; writes Kg [n*144], R3 [n*9]. */
int cm_element_kernel_dev(uint64_t xyz_u /*[n*3]*/, uint64_t xyz_v /*[n*3]*/, uint64_t props /*[n*7]*/,
                          uint64_t cr /*[n*6]*/, int64_t n, uint64_t Kg, uint64_t R3, uint64_t stream);

const char *cm_last_error(void);
const char *cm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CONMECH_B200_H */
