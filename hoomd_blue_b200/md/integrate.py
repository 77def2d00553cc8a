"""hoomd.md.Integrator for this path (hoomd/md/integrate.py:128-293)."""
import ctypes as C

from .. import _lib
from .._lib import check
import numpy as np

from .methods import ConstantVolume, ConstantPressure, Bussi, Langevin


class Integrator:
    """hoomd.md.Integrator(dt, forces=[pair, ...], methods=[ConstantVolume(...)]): every force
    contributes to the net force (hoomd/Integrator.cc:166-246); pair forces share one neighbour list.
    Constraints, rigid bodies, half-step hooks and integrate_rotational_dof are outside this path."""

    def __init__(self, dt, integrate_rotational_dof=False, forces=None, constraints=None,
                 methods=None, rigid=None, half_step_hook=None):
        if integrate_rotational_dof or constraints or rigid or half_step_hook:
            raise NotImplementedError("only translational ConstantVolume integration is on this path")
        self.dt = float(dt)
        self.forces = list(forces or [])
        self.methods = list(methods or [])
        self._h = None
        self._sim = None

    def _attach(self, sim):
        if self._h is not None:
            return
        from .bond import Bond
        from .pair import Pair, LJ
        from .nlist import Cell
        bonds = [f for f in self.forces if isinstance(f, Bond)]
        pairs = [f for f in self.forces if not isinstance(f, Bond)]
        if any(not isinstance(f, Pair) for f in pairs):
            raise NotImplementedError("this path carries pair and bonded forces")
        if any(f.nlist is not pairs[0].nlist for f in pairs):
            raise NotImplementedError("the pair forces of one integrator share one neighbour list on this path")
        if not pairs:
            # the net-force array belongs to the pair force: a bond-only integrator gets a pair force
            # that evaluates nothing (epsilon = 0) over a list without exclusions
            hidden = LJ(nlist=Cell(buffer=0.4, exclusions=()), default_r_cut=1.0)
            for a in sim.state.particle_types:
                for b in sim.state.particle_types:
                    hidden.params[(a, b)] = dict(epsilon=0.0, sigma=1.0)
            pairs = [hidden]
            self._hidden_pair = hidden
        if len(self.methods) > 1 or (self.methods
                                     and not isinstance(self.methods[0], (ConstantVolume, ConstantPressure, Langevin))):
            raise NotImplementedError("this path integrates with one ConstantVolume, ConstantPressure or "
                                      "Langevin method")
        L = _lib.lib()
        h = C.c_void_p()
        check(L.hb2_integrator_create(C.byref(h), sim.state._h, self.dt))
        self._h = h
        self._sim = sim
        self._pair = pairs[0]
        self._pairs = pairs
        for f in pairs:
            f._attach(sim)
            check(L.hb2_integrator_add_force(h, f._h))
        for b in bonds:
            b._attach(sim)
            check(L.hb2_integrator_add_bond(h, b._h))
        self._bonds = bonds
        # (no method: forces are computed, particles do not move — only run(0) is meaningful)
        th = self.methods[0].thermostat if self.methods else None
        if self.methods and isinstance(self.methods[0], Langevin):
            m = self.methods[0]
            gamma = np.array([m.gamma[t] for t in sim.state.particle_types], dtype=np.float64)
            check(L.hb2_integrator_set_langevin(h, m.kT, _lib.ptr(gamma, C.c_double), len(gamma)))
        elif self.methods and isinstance(self.methods[0], ConstantPressure):
            m = self.methods[0]
            from ..filter import All
            if not isinstance(m.filter, All):
                raise NotImplementedError("ConstantPressure integrates every particle on this path (filter.All)")
            S = (C.c_double * 6)(*m.S)
            dofs = (C.c_int * 6)(*[int(b) for b in m.box_dof])
            check(L.hb2_integrator_set_constant_pressure(
                h, S, m.tauS, m._couples[m.couple], dofs, int(m.rescale_all), m.gamma,
                1 if isinstance(th, Bussi) else 0, th.kT if th else 0.0, th.tau if th else 0.0))
            m._integrator = self
            check(L.hb2_integrator_set_barostat_dof(h, (C.c_double * 6)(*m._barostat_dof)))
        elif isinstance(th, Bussi):
            check(L.hb2_integrator_set_constant_volume(h, 1, th.kT, th.tau))
        else:
            check(L.hb2_integrator_set_constant_volume(h, 0, 0.0, 0.0))
        # the method's particle group (hoomd/md/methods/methods.py:145-175 takes any filter)
        self._group_tags = None
        if self.methods:
            from ..filter import All
            flt = self.methods[0].filter
            if not isinstance(flt, All):
                tags = np.ascontiguousarray(flt(sim.state), dtype=np.uint32)
                self._group_tags = tags
                buf = tags if len(tags) else np.zeros(1, dtype=np.uint32)
                check(L.hb2_integrator_set_filter_tags(h, _lib.ptr(buf, C.c_uint32), len(tags)))

    def _detach(self):
        if self._h is not None:
            _lib.lib().hb2_integrator_destroy(self._h)
            self._h = None
        for f in self.forces + ([self._hidden_pair] if getattr(self, "_hidden_pair", None) else []):
            f._detach()
            if hasattr(f, "nlist"):
                f.nlist._detach()

    def _run(self, steps, first_segment=True):
        """One segment of Simulation.run; the per-run statistics (nlist.num_builds, shortest_rebuild)
        are reset by the first segment only (System::run, hoomd/System.cc:103-110)."""
        if steps > 0 and not self.methods:
            raise NotImplementedError("an integrator without a method can only run(0) on this path")
        for b in getattr(self, "_bonds", []):
            b._sync()
        check(_lib.lib().hb2_integrator_set_reset_stats(self._h, int(bool(first_segment))))
        if self.methods and isinstance(self.methods[0], ConstantPressure):
            # parameters may have been changed between runs (npt.S = ..., npt.tauS = ...)
            m = self.methods[0]
            th = m.thermostat
            check(_lib.lib().hb2_integrator_set_constant_pressure(
                self._h, (C.c_double * 6)(*m.S), m.tauS, m._couples[m.couple],
                (C.c_int * 6)(*[int(b) for b in m.box_dof]), int(m.rescale_all), m.gamma,
                1 if isinstance(th, Bussi) else 0, th.kT if th else 0.0, th.tau if th else 0.0))
        check(_lib.lib().hb2_integrator_run(self._h, steps))
        if self.methods and isinstance(self.methods[0], ConstantPressure):
            self._sim.state._refresh_box()

    def _last_run_ms(self):
        ms = C.c_float(0)
        check(_lib.lib().hb2_integrator_last_run_ms(self._h, C.byref(ms)))
        return ms.value

    def _set_profiling(self, on):
        check(_lib.lib().hb2_integrator_set_profiling(self._h, int(bool(on))))

    def _profile(self):
        """dict(integrate_ms, nlist_ms, force_ms, steps, builds, force_launches) since profiling
        was switched on."""
        ms = (C.c_double * 3)()
        n = (C.c_uint64 * 3)()
        check(_lib.lib().hb2_integrator_get_profile(self._h, ms, n))
        return dict(integrate_ms=ms[0], nlist_ms=ms[1], force_ms=ms[2], steps=int(n[0]),
                    builds=int(n[1]), force_launches=int(n[2]))

    def _thermo(self):
        ke = C.c_double(0)
        pe = C.c_double(0)
        v = (C.c_double * 6)()
        check(_lib.lib().hb2_integrator_thermo(self._h, C.byref(ke), C.byref(pe), v))
        return ke.value, pe.value, list(v)

    @property
    def _translational_dof(self):
        return self._translational_dof_of(None)

    def _translational_dof_of(self, tags):
        """hoomd/md/IntegratorTwoStep.cc:154-189 for the query group `tags` (None: every particle): the
        dimension times the size of (method group ∩ query group), less the query group's share of the three
        centre-of-mass degrees when the (single) method conserves momentum AND acts on every particle —
        ConstantVolume does (hoomd/md/TwoStepConstantVolume.h), Langevin does not
        (hoomd/md/TwoStepLangevinBase.h) — and 0 without a method."""
        if not self.methods or self._sim is None:
            return 0
        N = self._sim.state.N_particles
        method_tags = getattr(self, "_group_tags", None)
        n_query = N if tags is None else len(tags)
        if method_tags is None:
            n_both = n_query
        elif tags is None:
            n_both = len(method_tags)
        else:
            n_both = len(np.intersect1d(method_tags, tags))
        dof = 3.0 * n_both
        if not isinstance(self.methods[0], Langevin) and (method_tags is None or len(method_tags) == N) and N:
            dof -= 3.0 * n_query / N
        return int(dof) if float(dof).is_integer() else dof

    def _kinetic_of(self, tags):
        ke = C.c_double(0)
        buf = tags if len(tags) else np.zeros(1, dtype=np.uint32)
        check(_lib.lib().hb2_integrator_kinetic_of_tags(self._h, _lib.ptr(buf, C.c_uint32), len(tags), C.byref(ke)))
        return ke.value
