"""A small TypeParameter (hoomd/data/typeparam.py, hoomd/data/parameterdicts.py): a dict keyed by
unordered pairs of particle type names with an optional default, validated on assignment."""


def _key(k):
    if not (isinstance(k, tuple) and len(k) == 2 and all(isinstance(s, str) for s in k)):
        raise KeyError(f"type-pair keys are 2-tuples of type names, got {k!r}")
    return tuple(sorted(k))


class TypePairDict:

    def __init__(self, name, validate, default=None):
        self._name = name
        self._validate = validate
        self._default = default
        self._d = {}
        self._on_change = None

    @property
    def default(self):
        return self._default

    @default.setter
    def default(self, v):
        self._default = self._validate(v)

    def __setitem__(self, key, value):
        keys = []
        if isinstance(key, tuple) and len(key) == 2 and any(isinstance(k, (list, tuple)) for k in key):
            # hoomd allows (['A','B'], ['A','B']) cross-product keys
            a = key[0] if isinstance(key[0], (list, tuple)) else [key[0]]
            b = key[1] if isinstance(key[1], (list, tuple)) else [key[1]]
            keys = [(x, y) for x in a for y in b]
        else:
            keys = [key]
        for k in keys:
            self._d[_key(k)] = self._validate(value)
        if self._on_change:
            self._on_change()

    def __getitem__(self, key):
        k = _key(key)
        if k in self._d:
            return self._d[k]
        if self._default is not None:
            return self._default
        raise KeyError(f"{self._name}[{key}] is not set")

    def get(self, key, types_check=True):
        return self[key]

    def to_base(self):
        return dict(self._d)
