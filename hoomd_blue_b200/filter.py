"""hoomd.filter for this path: ConstantVolume integrates filter.All() (hoomd/filter/all_.py)."""


class ParticleFilter:
    pass


class All(ParticleFilter):

    def __eq__(self, other):
        return isinstance(other, All)

    def __hash__(self):
        return hash("All")
