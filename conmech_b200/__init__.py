"""conmech_b200: B200-native engine for conmech's partial-structure stiffness solve."""
from .io_base import (Model, LoadCase, Node, Element, Support, Joint, CrossSec, Material,
                      PointLoad, UniformlyDistLoad, GravityLoad, mu2G, G2mu)

__version__ = '0.1.0'
