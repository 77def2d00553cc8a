"""hoomd.trigger for this path (hoomd/trigger.py): Periodic, On, Before, After. A bare int where a
trigger is expected means Periodic(int), as in the reference."""


class Trigger:

    def __call__(self, timestep):
        return self.compute(timestep)

    def next_after(self, timestep):
        """first step > timestep on which the trigger fires, or None"""
        raise NotImplementedError


class Periodic(Trigger):
    """fires when (timestep - phase) % period == 0 (hoomd/Trigger.h PeriodicTrigger)"""

    def __init__(self, period, phase=0):
        if int(period) <= 0:
            raise ValueError("period must be positive")
        self.period, self.phase = int(period), int(phase)

    def compute(self, timestep):
        return (timestep - self.phase) % self.period == 0

    def next_after(self, timestep):
        return timestep + self.period - (timestep - self.phase) % self.period


class On(Trigger):

    def __init__(self, timestep):
        self.timestep = int(timestep)

    def compute(self, timestep):
        return timestep == self.timestep

    def next_after(self, timestep):
        return self.timestep if self.timestep > timestep else None


class Before(Trigger):

    def __init__(self, timestep):
        self.timestep = int(timestep)

    def compute(self, timestep):
        return timestep < self.timestep

    def next_after(self, timestep):
        return timestep + 1 if timestep + 1 < self.timestep else None


class After(Trigger):

    def __init__(self, timestep):
        self.timestep = int(timestep)

    def compute(self, timestep):
        return timestep > self.timestep

    def next_after(self, timestep):
        return max(timestep, self.timestep) + 1


def as_trigger(t):
    if isinstance(t, Trigger):
        return t
    if isinstance(t, int):
        return Periodic(t)
    raise TypeError("trigger must be a hoomd.trigger.Trigger or an int period")
