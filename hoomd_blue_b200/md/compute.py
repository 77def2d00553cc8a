"""hoomd.md.compute.ThermodynamicQuantities for this path (hoomd/md/compute.py:55-192):
kinetic/potential energy, temperature, pressure of filter.All(); for any other filter the kinetic energy,
degrees of freedom and kinetic temperature of its group (what a thermostat on a sub-group uses)."""
import numpy as np


class ThermodynamicQuantities:
    _loggables = {"kinetic_energy": "scalar", "potential_energy": "scalar", "kinetic_temperature": "scalar",
                  "pressure": "scalar", "degrees_of_freedom": "scalar", "pressure_tensor": "sequence"}

    def __init__(self, filter):
        self.filter = filter
        self._sim = None

    def _attach(self, sim):
        self._sim = sim

    def _tags(self):
        from ..filter import All
        if isinstance(self.filter, All):
            return None
        return np.ascontiguousarray(self.filter(self._sim.state), dtype=np.uint32)

    def _q(self, need_virial=False):
        if self._sim is None or self._sim.operations.integrator is None:
            raise RuntimeError("ThermodynamicQuantities is not attached")
        integ = self._sim.operations.integrator
        if need_virial and not self._sim.always_compute_pressure:
            # The reference raises the pressure flag on the steps where a logger or writer asks for the
            # pressure (hoomd/md/compute.py -> ComputeThermo::computeProperties with PDataFlags
            # pressure_tensor). Here the observer switches the virial on for its own force pass over the
            # current list and off again: the trajectory is not perturbed.
            from .._lib import check, lib
            check(lib().hb2_system_set_compute_virial(self._sim.state._h, 1))
            try:
                return integ._thermo()
            finally:
                check(lib().hb2_system_set_compute_virial(self._sim.state._h, 0))
        return integ._thermo()

    @property
    def kinetic_energy(self):
        tags = self._tags()
        if tags is not None:
            return self._sim.operations.integrator._kinetic_of(tags)
        return self._q()[0]

    @property
    def potential_energy(self):
        if self._tags() is not None:
            raise NotImplementedError("potential energy of a sub-group is outside this path")
        return self._q()[1]

    @property
    def degrees_of_freedom(self):
        return self._sim.operations.integrator._translational_dof_of(self._tags())

    @property
    def kinetic_temperature(self):
        return 2.0 * self.kinetic_energy / self.degrees_of_freedom

    @property
    def pressure(self):
        """P = (2 K_trans / 3 + W) / V with W = (W_xx + W_yy + W_zz) / 3 summed over particles
        (hoomd/md/ComputeThermo.cc, 3D), including the forces' external virial (LJ tail correction)."""
        if self._tags() is not None:
            raise NotImplementedError("pressure of a sub-group is outside this path")
        ke, pe, w = self._q(need_virial=True)
        V = self._sim.state.box.volume
        return (2.0 * ke / 3.0 + (w[0] + w[3] + w[5]) / 3.0) / V

    @property
    def pressure_tensor(self):
        """(P_xx, P_xy, P_xz, P_yy, P_yz, P_zz) = (sum m v_a v_b + W_ab) / V (hoomd/md/ComputeThermo.cc:104-123,
        :236-252), the external virial (LJ tail correction) on the diagonal."""
        if self._tags() is not None:
            raise NotImplementedError("pressure tensor of a sub-group is outside this path")
        import ctypes as C
        from .._lib import check, lib
        _, _, w = self._q(need_virial=True)
        kt = (C.c_double * 6)()
        check(lib().hb2_integrator_kinetic_tensor(self._sim.operations.integrator._h, kt))
        V = self._sim.state.box.volume
        return tuple((kt[q] + w[q]) / V for q in range(6))
