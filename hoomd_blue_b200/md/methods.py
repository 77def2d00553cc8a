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
