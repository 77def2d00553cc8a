"""hoomd.md.dihedral for this path (hoomd/md/dihedral.py): Periodic. Dihedrals come from
Snapshot.dihedrals (four member tags)."""
from .bond import Bond


class Dihedral(Bond):
    _types_attr = "_dihedral_types"


class Periodic(Dihedral):
    """hoomd.md.dihedral.Periodic(): V = k/2 (1 + d cos(n phi - phi0)), params[dihedral type] =
    dict(k=, d=, n=, phi0=) (hoomd/md/dihedral.py, hoomd/md/HarmonicDihedralForceCompute.cc)."""
    _kind = 4
    _param_keys = ("k", "d", "n", "phi0")
