"""conmech_b200: B200-native engine for conmech's partial-structure stiffness solve.

``StiffnessChecker(model, checker_engine="cuda")`` keeps the pyconmech API; the engine is
:class:`CudaStiffness` (ctypes -> libconmech_b200.so -> sm_100a kernels).  Importing the package
needs neither the CUDA library nor a GPU; constructing an engine does, and raises without them.
"""
from .io_base import (Model, LoadCase, Node, Element, Support, Joint, CrossSec, Material,
                      PointLoad, UniformlyDistLoad, GravityLoad, mu2G, G2mu)
from .stiffness_base import StiffnessBase
from .cuda_stiffness import CudaStiffness
from .stiffness_checker import StiffnessChecker
from .masks import pack_masks, pack_masks_csr, ids_to_mask, mask_to_ids
from .sharding import shard_range, solve_many_sharded

__version__ = '0.1.0'
