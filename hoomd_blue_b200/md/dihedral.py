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


class OPLS(Dihedral):
    """hoomd.md.dihedral.OPLS(): V = 1/2 [k1 (1 + cos phi) + k2 (1 - cos 2 phi) + k3 (1 + cos 3 phi)
    + k4 (1 - cos 4 phi)], params[dihedral type] = dict(k1=, k2=, k3=, k4=)
    (hoomd/md/dihedral.py, hoomd/md/OPLSDihedralForceCompute.cc)."""
    _kind = 5
    _param_keys = ("k1", "k2", "k3", "k4")
