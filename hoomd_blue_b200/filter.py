"""hoomd.filter for this path (hoomd/filter/*.py, hoomd/filter/ParticleFilter*.h): a filter is a callable
that returns the sorted tags it selects from a state; methods hand those tags to the engine
(`hb2_integrator_set_filter_tags`). All, Null, Tags, Type, SetDifference, Union, Intersection and
CustomFilter follow the reference's construction, equality and hashing rules; Rigid is outside this path
(no rigid bodies)."""
import numpy as np


class ParticleFilter:
    """Base class (hoomd/filter/filter_.py): `filter(state)` -> sorted uint32 tags."""

    def __call__(self, state):
        raise NotImplementedError

    def __str__(self):
        return f"ParticleFilter.{type(self).__name__}"


class All(ParticleFilter):
    """hoomd/filter/all_.py"""

    def __call__(self, state):
        return np.arange(state.N_particles, dtype=np.uint32)

    def __eq__(self, other):
        return isinstance(other, All)

    def __hash__(self):
        return hash("All")


class Null(ParticleFilter):
    """hoomd/filter/null.py"""

    def __call__(self, state):
        return np.zeros(0, dtype=np.uint32)

    def __eq__(self, other):
        return isinstance(other, Null)

    def __hash__(self):
        return hash("Null")


class Tags(ParticleFilter):
    """hoomd/filter/tags.py: the given tags, de-duplicated and sorted."""

    def __init__(self, tags):
        self._tags = np.ascontiguousarray(np.unique(tags), dtype=np.uint32)

    @property
    def tags(self):
        return self._tags

    def __call__(self, state):
        if len(self._tags) and int(self._tags[-1]) >= state.N_particles:
            raise ValueError("filter.Tags: tag out of range")
        return self._tags

    def __hash__(self):
        if not hasattr(self, "_id"):
            self._id = hash(self._tags.tobytes())
        return self._id

    def __eq__(self, other):
        return type(self) is type(other) and np.array_equal(self._tags, other._tags)


class Type(ParticleFilter):
    """hoomd/filter/type_.py: particles whose type name is in `types`."""

    def __init__(self, types):
        self._types = frozenset(types)

    @property
    def types(self):
        return self._types

    def __call__(self, state):
        names = list(state.particle_types)
        ids = [names.index(t) for t in self._types if t in names]
        typeid = np.asarray(state.get_snapshot().particles.typeid)
        return np.flatnonzero(np.isin(typeid, ids)).astype(np.uint32)

    def __hash__(self):
        return hash(self._types)

    def __eq__(self, other):
        return type(self) is type(other) and self._types == other._types


class _SetOperation(ParticleFilter):
    """hoomd/filter/set_.py"""
    _symmetric = True

    def __init__(self, f, g):
        if f == g:
            raise ValueError(f"Cannot use same filter for {type(self).__name__}")
        self._f = f
        self._g = g

    def __hash__(self):
        return hash(hash(self._f) + hash(self._g))

    def __eq__(self, other):
        if type(self) is not type(other):
            return False
        if self._symmetric:
            return ((self._f == other._f or self._f == other._g)
                    and (self._g == other._g or self._g == other._f))
        return self._f == other._f and self._g == other._g


class SetDifference(_SetOperation):
    """f \\ g"""
    _symmetric = False

    def __call__(self, state):
        return np.setdiff1d(self._f(state), self._g(state)).astype(np.uint32)


class Union(_SetOperation):

    def __call__(self, state):
        return np.union1d(self._f(state), self._g(state)).astype(np.uint32)


class Intersection(_SetOperation):

    def __call__(self, state):
        return np.intersect1d(self._f(state), self._g(state)).astype(np.uint32)


class CustomFilter(ParticleFilter):
    """hoomd/filter/custom.py: subclass, implement __call__(state) -> tags, __hash__ and __eq__."""
