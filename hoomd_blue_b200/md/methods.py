"""hoomd.md.methods for this path (hoomd/md/methods/methods.py:85-175, thermostats.py)."""
from ..filter import All


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
        if not isinstance(filter, All):
            raise NotImplementedError("this path integrates filter.All() only")
        if thermostat is not None and not isinstance(thermostat, Bussi):
            raise NotImplementedError("thermostat must be None or thermostats.Bussi")
        self.filter = filter
        self.thermostat = thermostat
