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


class Or(Trigger):
    """fires when any of the triggers fires (hoomd.trigger.Or)"""

    def __init__(self, triggers):
        self.triggers = [as_trigger(t) for t in triggers]

    def compute(self, timestep):
        return any(t(timestep) for t in self.triggers)

    def next_after(self, timestep):
        nxt = [n for n in (t.next_after(timestep) for t in self.triggers) if n is not None]
        return min(nxt) if nxt else None


class And(Trigger):
    """fires when all of the triggers fire (hoomd.trigger.And)"""

    def __init__(self, triggers):
        self.triggers = [as_trigger(t) for t in triggers]

    def compute(self, timestep):
        return all(t(timestep) for t in self.triggers)

    def next_after(self, timestep):
        step = timestep
        for _ in range(1000000):  # leap to the latest of the children's next steps until they agree
            nxt = [t.next_after(step) for t in self.triggers]
            if any(n is None for n in nxt):
                return None
            cand = max(nxt)
            if all(t(cand) for t in self.triggers):
                return cand
            step = cand - 1 if cand - 1 > step else cand
        return None


class Not(Trigger):
    """fires when the trigger does not (hoomd.trigger.Not)"""

    def __init__(self, trigger):
        self.trigger = as_trigger(trigger)

    def compute(self, timestep):
        return not self.trigger(timestep)

    def next_after(self, timestep):
        for step in range(timestep + 1, timestep + 1000001):
            if not self.trigger(step):
                return step
        return None


def as_trigger(t):
    if isinstance(t, Trigger):
        return t
    if isinstance(t, int):
        return Periodic(t)
    raise TypeError("trigger must be a hoomd.trigger.Trigger or an int period")
