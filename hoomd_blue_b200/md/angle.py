"""hoomd.md.angle for this path (hoomd/md/angle.py): Harmonic and CosineSquared. Angles come from
Snapshot.angles (member tags a, b, c with b the vertex)."""
from .bond import Bond


class Angle(Bond):
    _types_attr = "_angle_types"
    _param_keys = ("k", "t0")


class Harmonic(Angle):
    """hoomd.md.angle.Harmonic(): V = k/2 (theta - t0)^2, params[angle type] = dict(k=, t0=)
    (hoomd/md/angle.py, hoomd/md/HarmonicAngleForceCompute.cc)."""
    _kind = 2


class CosineSquared(Angle):
    """hoomd.md.angle.CosineSquared(): V = k/2 (cos theta - cos t0)^2
    (hoomd/md/CosineSqAngleForceCompute.cc)."""
    _kind = 3
