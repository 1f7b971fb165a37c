"""Synthetic frames and element-subset samplers for the benchmark configurations.

These define the workloads of BASELINE.json configs 3-5 exactly as SURVEY.md §8(d) states
them, so that the CUDA path, the oracle and the reference are all driven by the same inputs.
Pure host code (numpy + ``random``); no GPU, no oracle.
"""
import random
from collections import deque

import numpy as np

from .io_base import (Model, Node, Element, Support, Joint, CrossSec, Material, LoadCase,
                      PointLoad, UniformlyDistLoad, GravityLoad)

STEEL_E = 2.1e8        # kN/m^2
STEEL_G12 = 8.076e7    # kN/m^2
STEEL_DENSITY = 78.5   # kN/m^3


def round_section(r=0.02, elem_tags=None):
    return CrossSec(A=np.pi * r**2, Jx=np.pi * r**4 / 2, Iy=np.pi * r**4 / 4, Iz=np.pi * r**4 / 4,
                    elem_tags=elem_tags, family='round', name='r{}'.format(r))


def grid_frame(nx, ny, nz, spacing=1.0, diagonals=False, joints=False, joint_seed=0, radius=0.02,
               model_name=None):
    """Space-frame grid of config 3: node id = (k*ny+j)*nx+i; members along x and
    y on every level k>0, along z between all levels; the nz=0 level is fully fixed.

    ``diagonals`` adds one face diagonal per x-z and y-z bay (non axis-aligned members).
    ``joints`` tags horizontal members per config 4: with ``random.Random(joint_seed)`` per
    horizontal element r<1/3 -> 'hinge_a', r<0.5 -> 'spring', r<0.6 -> 'both', else untagged.
    """
    nid = lambda i, j, k: (k * ny + j) * nx + i
    nodes = []
    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                nodes.append(Node([i * spacing, j * spacing, k * spacing], nid(i, j, k), k == 0))
    rng = random.Random(joint_seed)
    elements = []

    def add(u, v, horizontal):
        tag = ''
        if joints and horizontal:
            r = rng.random()
            tag = 'hinge_a' if r < 1 / 3 else 'spring' if r < 0.5 else 'both' if r < 0.6 else ''
        elements.append(Element([u, v], len(elements), tag, True))

    for k in range(nz):
        for j in range(ny):
            for i in range(nx):
                if k > 0 and i + 1 < nx:
                    add(nid(i, j, k), nid(i + 1, j, k), True)
                if k > 0 and j + 1 < ny:
                    add(nid(i, j, k), nid(i, j + 1, k), True)
                if k + 1 < nz:
                    add(nid(i, j, k), nid(i, j, k + 1), False)
                if diagonals and k + 1 < nz:
                    if i + 1 < nx:
                        add(nid(i, j, k), nid(i + 1, j, k + 1), False)
                    if j + 1 < ny:
                        add(nid(i, j, k), nid(i, j + 1, k + 1), False)
    supports = [Support([1] * 6, nid(i, j, 0)) for j in range(ny) for i in range(nx)]
    jts = []
    if joints:
        jts = [Joint([None] * 3 + [None, 0, 0] + [None] * 6, ['hinge_a']),
               Joint([None] * 3 + [None, None, 50.0] + [None] * 3 + [None, None, 50.0], ['spring']),
               Joint([None] * 3 + [None, 0, 0] + [None] * 3 + [None, 0, 0], ['both'])]
    mats = [Material(STEEL_E, STEEL_G12, 235e3, STEEL_DENSITY, elem_tags=None, family='Steel', name='S235')]
    secs = [round_section(radius)]
    name = model_name or 'grid_{}x{}x{}{}{}'.format(nx, ny, nz, '_diag' if diagonals else '', '_joints' if joints else '')
    return Model(nodes, elements, supports, jts, mats, secs, model_name=name)


def node_incidence(model):
    inc = [[] for _ in range(len(model.nodes))]
    for e in model.elements:
        u, v = e.end_node_inds
        inc[u].append(e.elem_ind)
        inc[v].append(e.elem_ind)
    return inc


def grounded_element_ids(model):
    """ids of elements with an end on a support node, ascending (facade ``fix_element_ids``)."""
    fixed = set(model.supports.keys())
    return [e.elem_ind for e in model.elements if e.end_node_inds[0] in fixed or e.end_node_inds[1] in fixed]


def bfs_sequence(model):
    """Construction sequence of config 2: breadth-first from grounded elements."""
    inc = node_incidence(model)
    frontier = deque(grounded_element_ids(model))
    chosen = set()
    order = []
    while frontier:
        e = frontier.popleft()
        if e in chosen:
            continue
        chosen.add(e)
        order.append(e)
        for v in model.elements[e].end_node_inds:
            for f in sorted(inc[v]):
                if f not in chosen:
                    frontier.append(f)
    return order


def connected_subset(model, seed, lo=0.2, hi=0.9, _cache={}):
    """Random connected element subset of config 3: ``random.Random(seed)``, target fraction
    U(lo,hi) of the elements, grown from the grounded elements by popping a random frontier
    entry.  Returns a sorted list of element ids."""
    key = id(model)
    if key not in _cache:
        _cache.clear()
        _cache[key] = (node_incidence(model), grounded_element_ids(model))
    inc, grounded = _cache[key]
    rng = random.Random(seed)
    nE = len(model.elements)
    target = max(1, int(round(rng.uniform(lo, hi) * nE)))
    frontier = list(grounded)
    in_frontier = set(frontier)
    chosen = []
    conn = [e.end_node_inds for e in model.elements]
    while frontier and len(chosen) < target:
        k = rng.randrange(len(frontier))
        frontier[k], frontier[-1] = frontier[-1], frontier[k]
        e = frontier.pop()
        chosen.append(e)
        for v in conn[e]:
            for f in inc[v]:
                if f not in in_frontier:
                    in_frontier.add(f)
                    frontier.append(f)
    chosen.sort()
    return chosen


def heightfield_subset(model, nx, ny, nz, seed):
    """Induced sub-frame of config 4: per seed a height h(i,j) in {0..nz-1}; nodes (i,j,k<=h)
    exist; every element with both ends existing is taken."""
    rng = random.Random(seed)
    h = [[rng.randrange(nz) for _ in range(nx)] for _ in range(ny)]
    def exists(n):
        i = n % nx
        j = (n // nx) % ny
        k = n // (nx * ny)
        return k <= h[j][i]
    return [e.elem_ind for e in model.elements if exists(e.end_node_inds[0]) and exists(e.end_node_inds[1])]


def shape_variant(base_points, seed, amplitude=0.1):
    """Node coordinates of geometry variant ``seed`` of a frame (many models of one topology):
    every coordinate of ``base_points`` ((nV, 3), metres) moved by U(-amplitude, amplitude), seeded."""
    rng = np.random.RandomState(seed)
    return np.asarray(base_points, dtype=float) + rng.uniform(-amplitude, amplitude, size=np.shape(base_points))


def config4_loadcases(model, seed=0):
    """The four load cases of config 4."""
    rng = random.Random(seed)
    nV = len(model.nodes)
    pls = []
    for v in range(nV):
        if rng.random() < 0.05:
            pls.append(PointLoad([rng.gauss(0, 1) for _ in range(3)], [0.0, 0.0, 0.0], v))
    g = lambda: GravityLoad([0, 0, -1])
    return [LoadCase(gravity_load=g()),
            LoadCase(point_loads=pls, gravity_load=g()),
            LoadCase(point_loads=pls),
            LoadCase(uniform_element_loads=[UniformlyDistLoad([0, 0, -0.1], [0, 0, -0.1], ['hinge_a'])],
                     gravity_load=g())]


from .masks import ids_to_mask, pack_masks, mask_to_ids   # noqa: E402,F401  (re-exported)
