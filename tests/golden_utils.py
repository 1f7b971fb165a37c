"""Helpers shared by the parity tests: golden loading, model/load reconstruction, error metric."""
import glob
import json
import os

import numpy as np

from conmech_b200 import Model, LoadCase, GravityLoad
from conmech_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
MODELS = os.path.join(GOLDEN, 'models')
REL_TOL = 1e-9        # north_star: displacements and reactions within 1e-9 relative (fp64)


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, '*.npz')))


class Golden(object):
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
        self.meta = json.loads(str(self.z['meta']))
        self.n = int(self.z['n_solves'])

    def rec(self, i):
        pre = 's{}_'.format(i)
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def extra(self, key):
        return self.z[key] if key in self.z.files else None

    def model(self):
        m = self.meta
        if 'model' in m:
            return Model.from_json(os.path.join(MODELS, m['model']))
        return syn.grid_frame(*m['grid'], diagonals=m['diagonals'], joints=m['joints'])

    def loadcase(self, model):
        m = self.meta
        if 'loadcase' in m:
            return syn.config4_loadcases(model)[m['loadcase']]
        pls = []
        uls = []
        if m.get('loads'):
            lc = LoadCase.from_json(os.path.join(MODELS, m['loads']))
            pls, uls = lc.point_loads, lc.uniform_element_loads
        g = GravityLoad(m['gravity']) if m.get('gravity') else None
        return LoadCase(pls, uls, g)


def rel_err(a, b):
    """||a-b||_inf / max(||b||_inf, tiny)."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    if b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def split_err(a, b):
    """Worst relative error over the translation/force and rotation/moment parts of a
    (...,6k)-shaped array, taken separately."""
    a = np.asarray(a, dtype=float).reshape(-1, 6)
    b = np.asarray(b, dtype=float).reshape(-1, 6)
    return max(rel_err(a[:, 0:3], b[:, 0:3]), rel_err(a[:, 3:6], b[:, 3:6]))
