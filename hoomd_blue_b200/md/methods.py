"""hoomd.md.methods for this path (hoomd/md/methods/methods.py:85-175, thermostats.py)."""
from ..filter import All, ParticleFilter


class Thermostat:
    pass


class Bussi(Thermostat):
    """hoomd.md.methods.thermostats.Bussi(kT, tau=0.0) (hoomd/md/methods/thermostats.py)."""

    def __init__(self, kT, tau=0.0):
        self.kT = float(kT)
        self.tau = float(tau)


class thermostats:  # namespace mirror: hoomd.md.methods.thermostats.Bussi
    Thermostat = Thermostat
    Bussi = Bussi


class Method:
    pass


class ConstantVolume(Method):
    """hoomd.md.methods.ConstantVolume(filter, thermostat=None)."""

    def __init__(self, filter, thermostat=None):
        if not isinstance(filter, ParticleFilter):
            raise TypeError("filter must be a hoomd.filter.ParticleFilter")
        if thermostat is not None and not isinstance(thermostat, Bussi):
            raise NotImplementedError("thermostat must be None or thermostats.Bussi")
        self.filter = filter
        self.thermostat = thermostat


class ConstantPressure(Method):
    """hoomd.md.methods.ConstantPressure(filter, S, tauS, couple, thermostat=None, box_dof=[True, True, True,
    False, False, False], rescale_all=False, gamma=0.0) (hoomd/md/methods/methods.py:177-580,
    hoomd/md/TwoStepConstantPressure.cc): Martyna-Tobias-Klein barostat on the box lengths, optional Bussi
    thermostat and Langevin piston. `S` is a pressure or the six stress components
    [S_xx, S_yy, S_zz, S_yz, S_xz, S_xy] (constants on this path). The shear degrees of freedom need a
    triclinic box, the stress of a sub-group its own virial sums: both are outside this path, so
    box_dof[3:] must be False and the filter must select every particle."""

    _loggables = {"barostat_energy": "scalar"}
    _couples = {"none": 0, "xy": 1, "xz": 2, "yz": 3, "xyz": 4}

    def __init__(self, filter, S, tauS, couple, thermostat=None,
                 box_dof=(True, True, True, False, False, False), rescale_all=False, gamma=0.0):
        if not isinstance(filter, ParticleFilter):
            raise TypeError("filter must be a hoomd.filter.ParticleFilter")
        if thermostat is not None and not isinstance(thermostat, Bussi):
            raise NotImplementedError("thermostat must be None or thermostats.Bussi")
        self.filter = filter
        self.thermostat = thermostat
        self.S = S
        self.tauS = float(tauS)
        self.couple = couple
        self.box_dof = box_dof
        self.rescale_all = bool(rescale_all)
        self.gamma = float(gamma)
        self._integrator = None
        self._barostat_dof = (0.0,) * 6

    # hoomd/md/methods/methods.py:531-540 (_preprocess_stress): a scalar is an isotropic pressure
    @property
    def S(self):
        return self._S

    @S.setter
    def S(self, value):
        if isinstance(value, (int, float)):
            value = (value, value, value, 0, 0, 0)
        value = tuple(float(v) for v in value)
        if len(value) != 6:
            raise ValueError("S must be a scalar or have six components")
        self._S = value

    @property
    def couple(self):
        return self._couple

    @couple.setter
    def couple(self, value):
        if value not in self._couples:
            raise ValueError("Invalid coupling mode " + str(value))  # TwoStepConstantPressure::setCouple
        self._couple = str(value)

    @property
    def box_dof(self):
        return self._box_dof

    @box_dof.setter
    def box_dof(self, value):
        value = [bool(v) for v in value]
        if len(value) != 6:
            raise ValueError("box_dof has six elements [x, y, z, xy, xz, yz]")
        self._box_dof = value

    @property
    def barostat_dof(self):
        """(nu_xx, nu_xy, nu_xz, nu_yy, nu_yz, nu_zz)"""
        if self._integrator is not None and self._integrator._h is not None:
            import ctypes as C
            from .._lib import check, lib
            nu = (C.c_double * 6)()
            check(lib().hb2_integrator_get_barostat_dof(self._integrator._h, nu))
            self._barostat_dof = tuple(nu)
        return self._barostat_dof

    @barostat_dof.setter
    def barostat_dof(self, value):
        value = tuple(float(v) for v in value)
        if len(value) != 6:
            raise ValueError("barostat_dof must have length 6")
        self._barostat_dof = value
        if self._integrator is not None and self._integrator._h is not None:
            import ctypes as C
            from .._lib import check, lib
            check(lib().hb2_integrator_set_barostat_dof(self._integrator._h, (C.c_double * 6)(*value)))

    def thermalize_barostat_dof(self):
        """Random barostat momenta (TwoStepConstantPressure::thermalizeBarostatDOF); call after run(0)."""
        if self._integrator is None or self._integrator._h is None:
            raise RuntimeError("Call Simulation.run(0) before thermalize_barostat_dof")
        from .._lib import check, lib
        check(lib().hb2_integrator_thermalize_barostat_dof(self._integrator._h, self._integrator._sim.timestep))

    @property
    def barostat_energy(self):
        if self._integrator is None or self._integrator._h is None:
            raise RuntimeError("barostat_energy requires run()")
        import ctypes as C
        from .._lib import check, lib
        e = C.c_double(0)
        check(lib().hb2_integrator_barostat_energy(self._integrator._h, C.byref(e)))
        return e.value


class _TypeDict(dict):
    """per-type parameter with a default (hoomd TypeParameter, hoomd/data/typeparam.py)"""

    def __init__(self, default):
        super().__init__()
        self.default = float(default)

    def __missing__(self, key):
        return self.default


class Langevin(Method):
    """hoomd.md.methods.Langevin(filter, kT, tally_reservoir_energy=False, default_gamma=1.0)
    (hoomd/md/methods/methods.py, hoomd/md/TwoStepLangevin.cc): translational Langevin dynamics of
    any hoomd.filter; `gamma[type]` is the drag coefficient. Rotational degrees of freedom (gamma_r) and the
    reservoir-energy tally are outside this path."""

    def __init__(self, filter, kT, tally_reservoir_energy=False, default_gamma=1.0, default_gamma_r=None):
        if not isinstance(filter, ParticleFilter):
            raise TypeError("filter must be a hoomd.filter.ParticleFilter")
        if tally_reservoir_energy:
            raise NotImplementedError("tally_reservoir_energy is outside this path")
        self.filter = filter
        self.kT = float(kT)
        self.gamma = _TypeDict(default_gamma)
        self.thermostat = None
