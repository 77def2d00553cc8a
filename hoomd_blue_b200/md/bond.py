"""hoomd.md.bond for this path (hoomd/md/bond.py): Harmonic. Bonds come from Snapshot.bonds; the
neighbour list excludes bonded pairs by default (Cell(exclusions=('bond',)))."""
import ctypes as C

import numpy as np

from .. import _lib
from .._lib import check


class Bond:
    _loggables = {"energy": "scalar"}
    _kind = 0
    _param_keys = ()
    _types_attr = "_bond_types"

    def __init__(self):
        self.params = {}
        self._h = None
        self._sim = None

    def _attach(self, sim):
        if self._h is not None:
            return
        L = _lib.lib()
        h = C.c_void_p()
        check(L.hb2_bond_create(C.byref(h), sim.state._h, self._kind))
        self._h, self._sim = h, sim
        self._sync()

    def _sync(self):
        types = getattr(self._sim.state, self._types_attr, [])
        for name, prm in self.params.items():
            if name not in types:
                raise KeyError(f"type {name!r} is not in the state ({types})")
            missing = [k for k in self._param_keys if k not in prm]
            if missing:
                raise KeyError(f"bond parameters {missing} are missing for {name!r}")
            arr = (C.c_double * len(self._param_keys))(*[float(prm[k]) for k in self._param_keys])
            check(_lib.lib().hb2_bond_set_params(self._h, types.index(name), arr, len(self._param_keys)))

    def _detach(self):
        if self._h is not None:
            _lib.lib().hb2_bond_destroy(self._h)
            self._h = None

    def _get(self, what):
        if self._h is None:
            raise RuntimeError("the bond force is not attached")
        self._sync()
        n = self._sim.state.N_particles
        f = np.zeros((n, 3)) if what == "forces" else None
        e = np.zeros(n) if what == "energies" else None
        v = np.zeros((n, 6)) if what == "virials" else None
        tot = C.c_double(0)
        p = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        check(_lib.lib().hb2_bond_get(self._h, p(f), p(e), p(v), C.byref(tot)))
        return dict(forces=f, energies=e, virials=v, energy=tot.value)[what]

    @property
    def forces(self):
        return self._get("forces")

    @property
    def energies(self):
        return self._get("energies")

    @property
    def virials(self):
        if not self._sim.always_compute_pressure:
            return None
        return self._get("virials")

    @property
    def energy(self):
        return self._get("energy")


class Harmonic(Bond):
    """hoomd.md.bond.Harmonic(): V(r) = k/2 (r - r0)^2, params[bond type] = dict(k=, r0=)
    (hoomd/md/bond.py:96-131, hoomd/md/EvaluatorBondHarmonic.h)."""
    _kind = 0
    _param_keys = ("k", "r0")


class FENEWCA(Bond):
    """hoomd.md.bond.FENEWCA(): FENE spring + WCA core in r - delta, params[bond type] =
    dict(k=, r0=, epsilon=, sigma=, delta=) (hoomd/md/bond.py:134-196, hoomd/md/EvaluatorBondFENE.h).
    A bond stretched beyond r0 raises "bond out of bounds"."""
    _kind = 1
    _param_keys = ("k", "r0", "epsilon", "sigma", "delta")
