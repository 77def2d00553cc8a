"""hoomd.md for the pair-force path: Integrator, nlist.Cell, pair.LJ/DPD/..., methods.ConstantVolume."""
from . import nlist, pair, methods, compute, bond, angle, dihedral  # noqa: F401
from .integrate import Integrator  # noqa: F401
