"""Frame data model: nodes, elements, supports, joints, sections, materials and loads.

Host-side mirror of the reference's input format (reference
``src/pyconmech/frame_analysis/io_base.py:31-423``): same class names, the same
attribute names (the engines duck-type on them, so a reference ``Model`` can be
handed to :class:`conmech_b200.CudaStiffness` unchanged) and the same JSON
layout.  Written from scratch for this repo; nothing here
touches the GPU.

Internal units are kN, m, rad.  Coordinates are converted to metres when a model
is read (``from_data``), exactly once (reference ``io_base.py:107-113``).
"""
import json
import os
import datetime
import warnings
from collections import OrderedDict

LENGTH_SCALE_CONVERSION = {'millimeter': 1e-3, 'meter': 1.0}


def mu2G(mu, E):
    """Shear modulus from Poisson ratio (reference ``io_base.py:303``)."""
    return E / (2 * (1 + mu))


def G2mu(G, E):
    """Poisson ratio from shear modulus (reference ``io_base.py:308``)."""
    return E / (2 * G) - 1


def _first_index_per_tag(items, what):
    """tag -> index of the entry that applies to it.

    When several entries claim one tag the reference keeps ``list(set_of_ids)[0]``
    (``io_base.py:56-87``); the same expression is used here so the
    choice is identical for any id pattern.
    """
    ids_of = {}
    for i, item in enumerate(items):
        for tag in item.elem_tags:
            ids_of.setdefault(tag, set()).add(i)
    chosen = {}
    for tag, ids in ids_of.items():
        if len(ids) > 1:
            warnings.warn('{} {} entries ({}) claim element tag {!r}; using the first'.format(
                len(ids), what, sorted(ids), tag))
        chosen[tag] = list(ids)[0]
    return chosen


class Node(object):
    def __init__(self, point, node_ind, is_grounded):
        self.point = point
        self.node_ind = node_ind
        self.is_grounded = is_grounded

    @classmethod
    def from_data(cls, data):
        return cls(data['point'], data['node_ind'], data['is_grounded'])

    def to_data(self):
        return {'point': list(self.point), 'node_ind': self.node_ind, 'is_grounded': self.is_grounded}

    def __repr__(self):
        return 'Node(#{},{},Grd:{})'.format(self.node_ind, self.point, self.is_grounded)


class Element(object):
    def __init__(self, end_node_inds, elem_ind, elem_tag='', bending_stiff=True):
        assert end_node_inds[0] != end_node_inds[1], 'zero length element not allowed!'
        self.end_node_inds = end_node_inds
        self.elem_ind = elem_ind
        self.elem_tag = elem_tag
        self.bending_stiff = bending_stiff

    @classmethod
    def from_data(cls, data):
        return cls(data['end_node_inds'], data['elem_ind'], data['elem_tag'], data['bending_stiff'])

    def to_data(self):
        return {'end_node_inds': self.end_node_inds, 'elem_ind': self.elem_ind,
                'elem_tag': self.elem_tag, 'bending_stiff': self.bending_stiff}

    def __repr__(self):
        return 'Element(#{}({}),{},Bend:{})'.format(self.elem_ind, self.elem_tag, self.end_node_inds, self.bending_stiff)


class Support(object):
    """``condition``: six bool/int flags, 1 = DOF fixed (Tx,Ty,Tz,Rx,Ry,Rz)."""
    def __init__(self, condition, node_ind):
        self.condition = condition
        self.node_ind = node_ind

    @classmethod
    def from_data(cls, data):
        return cls(data['condition'], data['node_ind'])

    def to_data(self):
        return {'condition': self.condition, 'node_ind': self.node_ind}

    def __repr__(self):
        return 'Support(#{},{})'.format(self.node_ind, self.condition)


class Joint(object):
    """End-release springs of every element carrying one of ``elem_tags``.

    ``c_conditions``: 12 entries, [T(3), R(3)] at end 0 then end 1; ``None`` = rigid,
    a number = spring stiffness (kN/m, kNm/rad), 0 = released.  Only the rotational
    entries may be finite (reference ``numpy_stiffness.py:594-595``).
    """
    def __init__(self, c_conditions, elem_tags):
        assert len(c_conditions) == 12
        self.c_conditions = c_conditions
        self.elem_tags = elem_tags

    @classmethod
    def from_data(cls, data):
        return cls(data['c_conditions'], data['elem_tags'])

    def to_data(self):
        return {'c_conditions': self.c_conditions, 'elem_tags': self.elem_tags}

    def __repr__(self):
        return 'Joint(#{},{})'.format(self.elem_tags, self.c_conditions)


class CrossSec(object):
    def __init__(self, A, Jx, Iy, Iz, elem_tags=None, family='unnamed', name='unnamed'):
        self.A = A
        self.Jx = Jx
        self.Iy = Iy
        self.Iz = Iz
        self.elem_tags = elem_tags if elem_tags else [None]
        self.family = family
        self.name = name

    @classmethod
    def from_data(cls, data):
        return cls(data['A'], data['Jx'], data['Iy'], data['Iz'], data['elem_tags'],
                   data.get('family', 'unnamed'), data.get('name', 'unnamed'))

    def to_data(self):
        return {'A': self.A, 'Jx': self.Jx, 'Iy': self.Iy, 'Iz': self.Iz,
                'elem_tags': self.elem_tags, 'family': self.family, 'name': self.name}

    def __repr__(self):
        return 'CrossSec({}-{} A:{} Jx:{} Iy:{} Iz:{} tags:{})'.format(
            self.family, self.name, self.A, self.Jx, self.Iy, self.Iz, self.elem_tags)


class Material(object):
    def __init__(self, E, G12, fy, density, elem_tags=None, family='unnamed', name='unnamed',
                 type_name='ISO', G3=None):
        self.E = E
        self.G12 = G12
        self.G3 = G3 or G12
        self.mu = G2mu(G12, E)
        self.fy = fy
        self.density = density
        self.elem_tags = elem_tags if elem_tags else [None]
        self.family = family
        self.name = name
        self.type_name = type_name

    @classmethod
    def from_data(cls, data):
        return cls(data['E'], data['G12'], data['fy'], data['density'], data['elem_tags'],
                   data.get('family', 'unnamed'), data.get('name', 'unnamed'),
                   data.get('type_name', 'ISO'), data.get('G3', None))

    def to_data(self):
        return {'E': self.E, 'G12': self.G12, 'G3': self.G3, 'mu': self.mu, 'fy': self.fy,
                'density': self.density, 'elem_tags': self.elem_tags, 'family': self.family,
                'name': self.name, 'type_name': self.type_name}

    def __repr__(self):
        return 'Material({}-{} E:{} mu:{} G12:{} density:{} tags:{})'.format(
            self.family, self.name, self.E, self.mu, self.G12, self.density, self.elem_tags)


class PointLoad(object):
    def __init__(self, force, moment, node_ind, loadcase=0):
        self.force = force
        self.moment = moment
        self.node_ind = node_ind
        self.loadcase = loadcase

    @classmethod
    def from_data(cls, data):
        return cls(data['force'], data['moment'], data['node_ind'], data.get('loadcase', 0))

    def to_data(self):
        return {'force': self.force, 'moment': self.moment, 'node_ind': self.node_ind,
                'loadcase': self.loadcase}

    def __repr__(self):
        return 'PointLoad(node {} | F {} | M {} | lc#{})'.format(self.node_ind, self.force, self.moment, self.loadcase)


class UniformlyDistLoad(object):
    """Line load ``q`` (kN/m, global axes) on every element carrying one of ``elem_tags``."""
    def __init__(self, q, load, elem_tags, loadcase=0):
        self.q = q
        self.load = load
        self.elem_tags = elem_tags
        self.loadcase = loadcase

    @classmethod
    def from_data(cls, data):
        return cls(data['q'], data['load'], data['elem_tags'], data.get('loadcase', 0))

    def to_data(self):
        return {'q': self.q, 'load': self.load, 'elem_tags': self.elem_tags, 'loadcase': self.loadcase}

    def __repr__(self):
        return 'UniformlyDistLoad(tags {} | q {} | lc#{})'.format(self.elem_tags, self.q, self.loadcase)


class GravityLoad(object):
    """Self weight: per-element line load ``force * density * A`` (``force`` is not normalised)."""
    def __init__(self, force=[0, 0, -1], loadcase=0):
        self.force = force
        self.loadcase = loadcase

    @classmethod
    def from_data(cls, data):
        return cls(data['force'], data.get('loadcase', 0))

    def to_data(self):
        return {'force': self.force, 'loadcase': self.loadcase}

    def __repr__(self):
        return 'GravityLoad({}, lc#{})'.format(self.force, self.loadcase)


class LoadCase(object):
    def __init__(self, point_loads=None, uniform_element_loads=None, gravity_load=None):
        self.point_loads = point_loads or []
        self.uniform_element_loads = uniform_element_loads or []
        self.gravity_load = gravity_load

    @classmethod
    def from_json(cls, file_path, lc_ind=0):
        assert os.path.exists(file_path), 'json file path does not exist!'
        with open(file_path, 'r') as f:
            data = json.load(f)
        data = data['loadcases'] if 'loadcases' in data else data
        if 'ploads' in data:          # flat LoadCase.to_data() layout (examples/scripts/run.py:141-143)
            return cls.from_data(data)
        return cls.from_data(data[str(lc_ind)])

    @classmethod
    def from_data(cls, data):
        point_loads = [PointLoad.from_data(pl) for pl in data.get('ploads', [])]
        element_loads = [UniformlyDistLoad.from_data(el) for el in data.get('eloads', [])]
        # The reference keeps data['gravity'] as a raw dict, which its own set_loads cannot
        # consume (io_base.py:179).  Here the dict is parsed into a GravityLoad.
        g = data.get('gravity', None)
        if isinstance(g, dict):
            g = GravityLoad.from_data(g) if 'force' in g else None
        return cls(point_loads, element_loads, g)

    def to_data(self):
        return {'ploads': [pl.to_data() for pl in self.point_loads],
                'eloads': [el.to_data() for el in self.uniform_element_loads],
                'gravity': self.gravity_load.to_data() if self.gravity_load is not None else None}

    def __repr__(self):
        return 'LoadCase(#pl:{},#el:{},gravity:{})'.format(
            len(self.point_loads), len(self.uniform_element_loads), self.gravity_load)


class Model(object):
    """A frame: lists of nodes/elements plus tag-indexed joints, materials and sections.

    ``supports`` becomes a dict keyed by node index that keeps the JSON order (the
    order in which fixity-reaction rows are reported).
    """
    def __init__(self, nodes, elements, supports, joints, materials, crosssecs, unit='meter', model_name=None):
        assert unit in LENGTH_SCALE_CONVERSION
        self.model_name = model_name
        self.generate_time = str(datetime.datetime.now())
        self.unit = unit
        scale = LENGTH_SCALE_CONVERSION[unit]
        if scale != 1.0:
            warnings.warn('Model unit {}: coordinates scaled by {} to metres.'.format(unit, scale))
        for node in nodes:
            node.point = [scale * c for c in node.point]

        self.nodes = nodes
        self.elements = elements
        for i, n in enumerate(nodes):
            assert i == n.node_ind, 'node_ind must equal list position'
        for i, e in enumerate(elements):
            assert i == e.elem_ind, 'elem_ind must equal list position'

        self.supports = {}
        for s in supports:
            self.supports[s.node_ind] = s

        self.joints = joints or []
        self.joint_id_from_etag = _first_index_per_tag(self.joints, 'joint')
        assert len(materials) > 0
        self.materials = materials
        self.material_id_from_etag = _first_index_per_tag(materials, 'material')
        assert len(crosssecs) > 0
        self.crosssecs = crosssecs
        self.crosssec_id_from_etag = _first_index_per_tag(crosssecs, 'cross section')

    @property
    def node_num(self):
        return len(self.nodes)

    @property
    def element_num(self):
        return len(self.elements)

    @classmethod
    def from_json(cls, file_path, verbose=False):
        assert os.path.exists(file_path), 'json file path does not exist!'
        with open(file_path, 'r') as f:
            data = json.load(f)
        if 'model_name' not in data:
            data['model_name'] = os.path.basename(file_path).split('.json')[0]
        return cls.from_data(data, verbose)

    @classmethod
    def from_data(cls, data, verbose=False):
        scale = LENGTH_SCALE_CONVERSION[data['unit']]
        nodes = []
        for nd in data['nodes']:
            nd = dict(nd)
            nd['point'] = [scale * c for c in nd['point'][:3]]
            nodes.append(Node.from_data(nd))
        elements = [Element.from_data(e) for e in data['elements']]
        supports = [Support.from_data(s) for s in data['supports']]
        joints = [Joint.from_data(j) for j in data.get('joints', [])]
        materials = [Material.from_data(m) for m in data['materials']]
        crosssecs = [CrossSec.from_data(c) for c in data['cross_secs']]
        if 'node_num' in data:
            assert len(nodes) == data['node_num']
        if 'element_num' in data:
            assert len(elements) == data['element_num']
        nV = len(nodes)
        for e in elements:
            assert 0 <= e.end_node_inds[0] < nV and 0 <= e.end_node_inds[1] < nV, \
                'element end point id not in node_list id range!'
        if not any(n.is_grounded for n in nodes):
            warnings.warn('The structure must have at lease one grounded node!')
        if verbose:
            print('Model: {} | unit: {} | Nodes: {} | Elements: {} | Supports: {} | Joints: {} | '
                  'Materials: {} | Cross Secs: {}'.format(data['model_name'], data['unit'], len(nodes),
                  len(elements), len(supports), len(joints), len(materials), len(crosssecs)))
        # coordinates are already metres: do not pass the unit on (reference io_base.py:112,144)
        return cls(nodes, elements, supports, joints, materials, crosssecs, model_name=data['model_name'])

    def to_data(self):
        data = OrderedDict()
        data['model_name'] = self.model_name
        data['unit'] = self.unit
        data['generate_time'] = self.generate_time
        data['node_num'] = len(self.nodes)
        data['element_num'] = len(self.elements)
        data['nodes'] = [n.to_data() for n in self.nodes]
        data['elements'] = [e.to_data() for e in self.elements]
        data['supports'] = [s.to_data() for s in self.supports.values()]
        data['joints'] = [j.to_data() for j in self.joints]
        data['materials'] = [m.to_data() for m in self.materials]
        data['cross_secs'] = [cs.to_data() for cs in self.crosssecs]
        return data
