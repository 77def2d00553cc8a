"""hoomd.md.compute.ThermodynamicQuantities for this path (hoomd/md/compute.py:55-192):
kinetic/potential energy, temperature, pressure of filter.All()."""
import numpy as np


class ThermodynamicQuantities:
    _loggables = {"kinetic_energy": "scalar", "potential_energy": "scalar", "kinetic_temperature": "scalar",
                  "pressure": "scalar", "degrees_of_freedom": "scalar"}

    def __init__(self, filter):
        self.filter = filter
        self._sim = None

    def _attach(self, sim):
        self._sim = sim

    def _q(self):
        if self._sim is None or self._sim.operations.integrator is None:
            raise RuntimeError("ThermodynamicQuantities is not attached")
        return self._sim.operations.integrator._thermo()

    @property
    def kinetic_energy(self):
        return self._q()[0]

    @property
    def potential_energy(self):
        return self._q()[1]

    @property
    def degrees_of_freedom(self):
        return 3 * self._sim.state.N_particles - 3

    @property
    def kinetic_temperature(self):
        return 2.0 * self.kinetic_energy / self.degrees_of_freedom

    @property
    def pressure(self):
        """P = (2 K_trans / 3 + W) / V with W = (W_xx + W_yy + W_zz) / 3 summed over particles
        (hoomd/md/ComputeThermo.cc, 3D). Needs always_compute_pressure."""
        ke, pe, w = self._q()
        V = self._sim.state.box.volume
        return (2.0 * ke / 3.0 + (w[0] + w[3] + w[5]) / 3.0) / V

    @property
    def pressure_tensor(self):
        raise NotImplementedError("pressure_tensor needs the kinetic tensor; outside this path")
