"""CPU ORACLE — test infrastructure, not product code.

A numpy/scipy restatement of the reference's numpy engine for the partial-structure
stiffness solve (reference ``src/pyconmech/frame_analysis/numpy_stiffness_assembly.py``
and ``numpy_stiffness.py``, pyconmech 0.6.0).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module; the
product path (``conmech_b200``) never does and fails loudly without its CUDA library.

Parity is PINNED: ``tests/golden/make_golden.py`` (run in the authoring container, where
``/root/reference`` exists) solved every bundled fixture and synthetic case with the real
reference and froze inputs + outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this module against those vectors (bit-for-bit on DOF maps and flags, <=1e-12 on
values - in practice the outputs are bit-identical because the same scipy calls are made
in the same order) and against the reference's own stored result
``cantilever_beam_truss_result.json`` and the textbook answers its tests hold.

Third-party arithmetic on the path: scipy's bundled SuperLU (``scipy.sparse.linalg.splu``),
reference pin ``scipy>=1.2.3`` (``setup.py:33``); this image has scipy 1.18.1.  The oracle
makes the very same call (``numpy_stiffness.py:300``).

Every function cites the reference lines it follows.  Quirks Q1-Q16 of SURVEY.md
Appendix A are reproduced on purpose.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

EPS = 1e-14   # numpy_stiffness_assembly.py:8
BIG = 1e14    # numpy_stiffness_assembly.py:9


# ----------------------------------------------------------------------------- element math
def axial_block(L, A, E, c1=np.inf, c2=np.inf):
    """2x2 bar/torsion block d*[[1,-1],[-1,1]] (numpy_stiffness_assembly.py:11-71, 74-98)."""
    if abs(c1) < EPS or abs(c2) < EPS:
        d = 0.0
    else:
        i1 = 1 / c1 if c1 < BIG else 0.0
        i2 = 1 / c2 if c2 < BIG else 0.0
        d = 1.0 / (L / (E * A) + i1 + i2)
    K = np.ones((2, 2))
    K[0, 1] = K[1, 0] = -1.0
    return K * d


def bending_coefficients(a1, a2):
    """(a, aa, a1_2, a1_3, a2_2, a2_3) for end spring ratios a1, a2
    (numpy_stiffness_assembly.py:155-189)."""
    if abs(a1) < BIG and abs(a2) < BIG:
        den = a1 * a2 + 4 * a1 + 4 * a2 + 12
        a = a1 * a2 / den
        aa = (a1 * a2 + a1 + a2) / den
        a2_2 = (a2 + 2) * a1 / den
        a2_3 = (a2 + 3) * a1 / den
        a1_2 = (a1 + 2) * a2 / den
        a1_3 = (a1 + 3) * a2 / den
    elif abs(a1) < EPS and abs(a2) > BIG:
        a, aa, a2_2, a2_3, a1_2, a1_3 = 0.0, 0.25, 0.0, 0.0, 0.5, 0.75
    elif abs(a2) < EPS and abs(a1) > BIG:
        a, aa, a1_2, a1_3, a2_2, a2_3 = 0.0, 0.25, 0.0, 0.0, 0.5, 0.75
    else:
        assert abs(a1) > EPS and abs(a2) > EPS
        i1 = 0.0 if abs(a1) > BIG else 1 / a1
        i2 = 0.0 if abs(a2) > BIG else 1 / a2
        den = 1 + 4 * i1 + 4 * i2 + 12 * i1 * i2
        a = 1 / den
        aa = (1 + i1 + i2) / den
        a2_2 = (1 + 2 * i1) / den
        a2_3 = (1 + 3 * i1) / den
        a1_2 = (1 + 2 * i2) / den
        a1_3 = (1 + 3 * i2) / den
    return a, aa, a1_2, a1_3, a2_2, a2_3


def bending_block(L, E, Iz, axis=2, c1=np.inf, c2=np.inf):
    """4x4 Euler-Bernoulli block on (v1, th1, v2, th2) with end rotational springs
    (numpy_stiffness_assembly.py:101-206).  Q2: the spring ratio is ``c*L/E*Iz`` evaluated
    left to right."""
    s = 1.0 if axis == 2 else -1.0
    assert abs(E) > EPS and abs(Iz) > EPS
    a1 = c1 * L / E * Iz
    a2 = c2 * L / E * Iz
    a, aa, a1_2, a1_3, a2_2, a2_3 = bending_coefficients(a1, a2)
    K = np.zeros((4, 4))
    K[0, :] = [12. / L**2 * aa, s * 6 / L * a2_2, -12. / L**2 * aa, s * 6 / L * a1_2]
    K[1, :] = [s * 6 / L * a2_2, 4 * a2_3, s * (-6 / L) * a2_2, 2 * a]
    K[2, :] = -K[0, :]
    K[3, :] = [s * 6 / L * a1_2, 2 * a, s * (-6 / L) * a1_2, 4 * a1_3]
    K *= E * Iz / L
    return K


def local_stiffness(L, A, Jx, Iy, Iz, E, mu, cr1, cr2):
    """12x12 local stiffness (numpy_stiffness_assembly.py:208-281).  Q1: both bending
    planes use Iz; Iy is unused."""
    G = E / (2 * (1 + mu))                       # io_base.py:303
    K = np.zeros((12, 12))
    K[np.ix_([0, 6], [0, 6])] += axial_block(L, A, E)
    K[np.ix_([3, 9], [3, 9])] += axial_block(L, Jx, G, cr1[0], cr2[0])
    K[np.ix_([1, 5, 7, 11], [1, 5, 7, 11])] += bending_block(L, E, Iz, 2, cr1[2], cr2[2])
    K[np.ix_([2, 4, 8, 10], [2, 4, 8, 10])] += bending_block(L, E, Iz, 1, cr1[1], cr2[1])
    return K


def rotation3(pu, pv):
    """3x3 global->local rotation, rows = local axes (numpy_stiffness_assembly.py:283-327).
    Q3: members parallel to global z get the mirrored frame."""
    pu = np.asarray(pu, dtype=float)
    pv = np.asarray(pv, dtype=float)
    L = np.linalg.norm(pu - pv)
    if L < EPS:
        raise ValueError('Bar length too small!: Node {} - Node {}'.format(pu, pv))
    cx = (pv[0] - pu[0]) / L
    cy = (pv[1] - pu[1]) / L
    cz = (pv[2] - pu[2]) / L
    R = np.zeros((3, 3))
    if abs(abs(cz) - 1.0) < EPS:
        R[0, 2] = -cz
        R[1, 1] = 1
        R[2, 0] = cz
    else:
        x = np.array([cx, cy, cz])
        y = -np.cross(x, [0, 0, 1.0])
        y /= np.linalg.norm(y)
        R[0, :] = x
        R[1, :] = y
        R[2, :] = np.cross(x, y)
    return R


def blockdiag4(R3):
    """12x12 blockdiag(R,R,R,R) (numpy_stiffness_assembly.py:329-343)."""
    R = np.zeros((12, 12))
    for i in range(4):
        R[3 * i:3 * i + 3, 3 * i:3 * i + 3] = R3
    return R


def lump_udl(w_G, R3, L):
    """Fixed-end lumping of a uniform line load w_G (global) to a 12-vector of global nodal
    loads (numpy_stiffness_assembly.py:393-420)."""
    w_G = np.asarray(w_G, dtype=float)
    q = np.zeros(12)
    q[0:3] = w_G * L / 2
    q[6:9] = w_G * L / 2
    w_L = R3.dot(w_G)
    M0 = np.array([0, -w_L[2] * (L**2) / 12., w_L[1] * (L**2) / 12.])
    q[3:6] = R3.T.dot(M0)
    q[9:12] = R3.T.dot(-M0)
    return q


# ----------------------------------------------------------------------------- engine
class SolveResult(object):
    """What one partial-structure solve produces, in order-free (dense) form."""
    __slots__ = ('success', 'status', 'has_result', 'n_free', 'n_fixed', 'dof_map', 'node_exists',
                 'U', 'R', 'eR', 'compliance', 'max_trans', 'max_rot', 'min_pivot')


# status codes shared with include/conmech_b200.h
ST_OK, ST_NO_SUPPORT, ST_NO_FIXED_DOF, ST_SINGULAR, ST_ZERO_LOAD = 0, 1, 2, 3, 4


class FrameOracle(object):
    """Restatement of ``NumpyStiffness`` (numpy_stiffness.py:34-634) with dense outputs.

    ``solve(ids)`` returns a :class:`SolveResult` whose ``U``/``R`` are full 6*nV vectors and
    whose ``eR`` is an (nE,12) table with NaN rows for absent elements, so comparisons never
    depend on row order.
    """

    def __init__(self, model):
        self.model = model
        self.nodes = model.nodes
        self.elements = model.elements
        self.supports = model.supports
        if len(self.supports) == 0:                      # stiffness_base.py:25
            raise RuntimeError('there needs to be at least one support (fixed) vertex in the model!')
        self.nV = len(self.nodes)
        self.nE = len(self.elements)
        self.trans_tol = 1e-3
        self.rot_tol = np.pi / 180 * 3
        self.xyz = np.array([n.point for n in self.nodes], dtype=float)
        self.conn = np.array([e.end_node_inds for e in self.elements], dtype=int)
        # numpy_stiffness.py:563-570
        self.v_id_map = np.arange(6 * self.nV).reshape(self.nV, 6)
        self.id_map = np.hstack([self.v_id_map[self.conn[:, 0]], self.v_id_map[self.conn[:, 1]]])
        self.neighbors = [[] for _ in range(self.nV)]
        for e, (u, v) in enumerate(self.conn):
            self.neighbors[u].append(e)
            self.neighbors[v].append(e)
        self.e_ind_from_tag = {}
        for e in self.elements:
            self.e_ind_from_tag.setdefault(e.elem_tag, []).append(e.elem_ind)
        self._build_element_tables()
        self.nodal_load = np.zeros(6 * self.nV)
        self.include_self_weight = False
        self.q_gravity = None          # (nE,12)
        self.q_udl = None              # (nE,12) or None when no UDL is set
        # support table after the all-truss-neighbours rotational pin (numpy_stiffness.py:172-184, Q11)
        self.sup_nodes = list(self.supports.keys())
        self.sup_cond = np.zeros((len(self.sup_nodes), 6), dtype=int)
        for i, v in enumerate(self.sup_nodes):
            c = np.array(self.supports[v].condition, dtype=int)
            if all(not self.elements[e].bending_stiff for e in self.neighbors[v]):
                c[3:] = 1
            self.sup_cond[i] = c

    # -- tag lookups (stiffness_base.py:131-156, Q12)
    def _crosssec(self, e):
        m = self.model
        tag = self.elements[e].elem_tag
        return m.crosssecs[m.crosssec_id_from_etag[tag] if tag in m.crosssec_id_from_etag
                           else m.crosssec_id_from_etag[None]]

    def _material(self, e):
        m = self.model
        tag = self.elements[e].elem_tag
        return m.materials[m.material_id_from_etag[tag] if tag in m.material_id_from_etag
                           else m.material_id_from_etag[None]]

    def _joint(self, e):
        m = self.model
        tag = self.elements[e].elem_tag
        return m.joints[m.joint_id_from_etag[tag]] if tag in m.joint_id_from_etag else None

    def _build_element_tables(self):
        """numpy_stiffness.py:579-614."""
        nE = self.nE
        self.K_G = np.zeros((nE, 12, 12))
        self.R12 = np.zeros((nE, 12, 12))
        self.R3 = np.zeros((nE, 3, 3))
        self.length = np.zeros(nE)
        for e in range(nE):
            u, v = self.conn[e]
            pu, pv = self.xyz[u], self.xyz[v]
            L = np.linalg.norm(pu - pv)
            cr1 = [np.inf] * 3
            cr2 = [np.inf] * 3
            jt = self._joint(e)
            if jt is not None:
                c = jt.c_conditions
                assert all(x is None or x == np.inf for x in c[0:3]) and \
                       all(x is None or x == np.inf for x in c[6:9]), 'We do not support translational joint now!'
                cr1 = [np.inf if x is None else x for x in c[3:6]]
                cr2 = [np.inf if x is None else x for x in c[9:12]]
            if not self.elements[e].bending_stiff:
                cr1 = [0.0] * 3
                cr2 = [0.0] * 3
            try:
                R3 = rotation3(pu, pv)
            except ValueError as err:
                raise ValueError('E#{} - {}'.format(e, err))
            cs = self._crosssec(e)
            mat = self._material(e)
            K_loc = local_stiffness(L, cs.A, cs.Jx, cs.Iy, cs.Iz, mat.E, mat.mu, cr1, cr2)
            R = blockdiag4(R3)
            self.K_G[e] = (R.T.dot(K_loc)).dot(R)
            self.R12[e] = R
            self.R3[e] = R3
            self.length[e] = L

    # -- loads (numpy_stiffness.py:98-138, 401-430)
    def set_self_weight_load(self, include_self_weight=True, gravity_direction=(0., 0., -1.)):
        self.include_self_weight = include_self_weight
        if include_self_weight:
            g = np.array(gravity_direction, dtype=float)
            assert len(g) == 3
            q = np.zeros((self.nE, 12))
            for e in range(self.nE):
                w = g * (self._material(e).density * self._crosssec(e).A)
                q[e] = lump_udl(w, self.R3[e], self.length[e])
            self.q_gravity = q

    def set_uniformly_distributed_loads(self, unif_dist_loads):
        if unif_dist_loads is None:
            self.q_udl = None
            return
        q = np.zeros((self.nE, 12))
        for tag, ul in unif_dist_loads.items():
            for e in self.e_ind_from_tag[tag]:
                q[e] = lump_udl(np.array(ul.q, dtype=float), self.R3[e], self.length[e])
        self.q_udl = q

    def set_load(self, nodal_forces):
        if nodal_forces is None:
            self.nodal_load = np.zeros(6 * self.nV)
        else:                                            # Q6: only listed nodes are overwritten
            for v, pf in nodal_forces.items():
                self.nodal_load[self.v_id_map[v]] = list(pf.force) + list(pf.moment)

    def set_loads(self, loadcase):
        """stiffness_checker.py:188-218."""
        g = loadcase.gravity_load
        if g is not None:
            self.set_self_weight_load(True, g.force)
        else:
            self.set_self_weight_load(False)
        pls = loadcase.point_loads
        self.set_load({pl.node_ind: pl for pl in pls} if pls else None)
        uls = loadcase.uniform_element_loads
        if uls:
            d = {}
            for ul in uls:
                tags = ul.elem_tags if len(ul.elem_tags) > 0 else [""]
                for t in tags:
                    d[t] = ul
            self.set_uniformly_distributed_loads(d)
        else:
            self.set_uniformly_distributed_loads(None)

    def set_nodal_displacement_tol(self, trans_tol=1e-3, rot_tol=np.inf):
        self.trans_tol = trans_tol
        self.rot_tol = rot_tol

    def load_vector(self, ids):
        """numpy_stiffness.py:250-257, 618-634: point + UDL(existing) + self weight(existing)."""
        P = np.copy(self.nodal_load)
        ids = np.asarray(ids, dtype=int)
        if self.q_udl is not None:
            tmp = np.zeros(6 * self.nV)
            np.add.at(tmp, self.id_map[ids].ravel(), self.q_udl[ids].ravel())
            P += tmp
        if self.include_self_weight:
            tmp = np.zeros(6 * self.nV)
            np.add.at(tmp, self.id_map[ids].ravel(), self.q_gravity[ids].ravel())
            P += tmp
        return P

    # -- per-subset pieces
    def dof_permutation(self, ids):
        """numpy_stiffness.py:147-219.  Returns (status, dof_map, n_free, n_fixed, node_exists,
        n_exist_fixed_nodes); dof_map is the reference's id_map_RO."""
        total = 6 * self.nV
        stat = -np.ones(total, dtype=int)
        node_exists = np.zeros(self.nV, dtype=bool)
        ids = np.asarray(ids, dtype=int)
        node_exists[self.conn[ids].ravel()] = True
        stat[self.id_map[ids].ravel()] = 0
        if not any(node_exists[v] for v in self.sup_nodes):
            return ST_NO_SUPPORT, None, 0, 0, node_exists, 0
        n_fixed = 0
        n_fix_nodes = 0
        for i, v in enumerate(self.sup_nodes):
            if node_exists[v]:
                stat[6 * v:6 * v + 6] = self.sup_cond[i]
                n_fix_nodes += 1
                n_fixed += int(self.sup_cond[i].sum())
        if n_fixed == 0:
            return ST_NO_FIXED_DOF, None, 0, 0, node_exists, n_fix_nodes     # Q10
        n_free = 6 * int(node_exists.sum()) - n_fixed
        # stable three-way partition free | fixed | absent
        dof_map = np.concatenate([np.flatnonzero(stat == 0), np.flatnonzero(stat == 1),
                                  np.flatnonzero(stat == -1)])
        return ST_OK, dof_map, n_free, n_fixed, node_exists, n_fix_nodes

    def assemble(self, ids):
        """numpy_stiffness_assembly.py:347-389: COO triplets in (e, i, j) order, entries with
        |k| <= 1e-14 dropped, duplicates summed by csc_matrix."""
        ids = np.asarray(ids, dtype=int)
        rows = np.repeat(self.id_map[ids][:, :, None], 12, axis=2).ravel()
        cols = np.repeat(self.id_map[ids][:, None, :], 12, axis=1).ravel()
        data = self.K_G[ids].ravel()
        keep = np.abs(data) > EPS
        total = 6 * self.nV
        return sp.csc_matrix((data[keep], (rows[keep], cols[keep])), shape=(total, total), dtype=float)

    def solve(self, exist_element_ids=()):
        """numpy_stiffness.py:228-397."""
        ids = list(exist_element_ids)
        if len(ids) == 0:
            ids = list(range(self.nE))
        total = 6 * self.nV
        res = SolveResult()
        res.min_pivot = np.nan
        res.compliance = 0.0
        res.max_trans = res.max_rot = np.nan
        res.U = res.R = res.eR = None
        status, dof_map, n_free, n_fixed, node_exists, _ = self.dof_permutation(ids)
        res.status, res.dof_map, res.n_free, res.n_fixed, res.node_exists = \
            status, dof_map, n_free, n_fixed, node_exists
        if status != ST_OK:
            res.success = False
            res.has_result = False
            return res

        Perm = sp.csc_matrix((np.ones(total, dtype=int), (np.arange(total), dof_map)), shape=(total, total))
        PermT = Perm.T
        K_full = self.assemble(ids)
        K_perm = (Perm.dot(K_full)).dot(PermT)
        K_mm = K_perm[0:n_free, 0:n_free]
        K_fm = K_perm[n_free:n_free + n_fixed, 0:n_free]

        P = self.load_vector(ids)
        P_m = np.zeros(n_free)
        P_f = np.zeros(n_fixed)
        U_m = np.zeros(n_free)
        if np.linalg.norm(P) > EPS:
            P_perm = Perm.dot(P)
            P_m = P_perm[0:n_free]
            P_f = P_perm[n_free:n_free + n_fixed]
            diag = K_mm.diagonal()
            if all(abs(a) > EPS for a in diag):          # Jacobi scaling, numpy_stiffness.py:283-291
                Einv = sp.diags(np.array([1 / np.sqrt(a) for a in diag]), format='csc')
            else:
                Einv = sp.eye(K_mm.shape[0], format='csc')
            K_sc = (Einv.dot(K_mm)).dot(Einv)
            try:
                LU = spla.splu(K_sc, diag_pivot_thresh=0, options=dict(SymmetricMode=True))
            except RuntimeError:
                res.status = ST_SINGULAR
                res.success = False
                res.has_result = False
                return res
            piv = LU.U.diagonal()
            res.min_pivot = float(piv.min()) if piv.size else np.nan
            if not (piv > -EPS).all():
                res.status = ST_SINGULAR
                res.success = False
                res.has_result = False
                return res
            U_m = Einv * LU.solve(Einv.dot(P_m))
            res.compliance = 0.5 * U_m.dot(P_m)
        else:
            res.status = ST_ZERO_LOAD

        R = np.zeros(total)
        R[n_free:n_free + n_fixed] = K_fm.dot(U_m) - P_f
        R = PermT.dot(R)
        U = np.zeros(total)
        U[0:n_free] = U_m
        U = PermT.dot(U)

        eR = np.full((self.nE, 12), np.nan)
        for e in ids:                                    # Q4: negated
            eR[e] = -self.R12[e].dot(self.K_G[e].dot(U[self.id_map[e]]))
        res.U, res.R, res.eR = U, R, eR
        res.has_result = True
        ex = np.flatnonzero(node_exists)
        Un = U.reshape(self.nV, 6)[ex]
        res.max_trans = float(np.max(np.linalg.norm(Un[:, 0:3], axis=1)))
        res.max_rot = float(np.max(np.linalg.norm(Un[:, 3:6], axis=1)))
        res.success = bool(res.max_trans < self.trans_tol)   # numpy_stiffness.py:441 (strict <, NaN -> False)
        return res
