"""hoomd.logging.Logger for this path (hoomd/logging.py): a registry of (namespace, callable) pairs.
Objects expose their loggable quantities through `_loggables` (name -> category); the namespace of a
quantity is the object's module path below the package plus its class name, as in the reference
(('md', 'compute', 'ThermodynamicQuantities', 'kinetic_energy'))."""


def _namespace(obj):
    mod = type(obj).__module__.split(".")[1:]  # drop the package name
    if mod and mod[-1] in ("simulation",):
        mod = mod[:-1]
    return tuple(mod) + (type(obj).__name__,)


class Logger:

    def __init__(self, categories=None, only_default=True):
        self.categories = set(categories) if categories is not None else None
        self._q = {}
        self._cat = {}

    def add(self, obj, quantities=None, user_name=None):
        avail = getattr(obj, "_loggables", None)
        if avail is None:
            raise ValueError(f"{type(obj).__name__} has no loggable quantities")
        names = list(avail) if quantities is None else list(quantities)
        ns = _namespace(obj) if user_name is None else _namespace(obj)[:-1] + (user_name,)
        added = []
        for q in names:
            if q not in avail:
                raise ValueError(f"{type(obj).__name__} has no loggable quantity {q!r}")
            if self.categories is not None and avail[q] not in self.categories:
                continue
            self._q[ns + (q,)] = (lambda o=obj, a=q: getattr(o, a))
            self._cat[ns + (q,)] = avail[q]
            added.append(ns + (q,))
        return added

    def __iadd__(self, obj):
        self.add(obj)
        return self

    def __setitem__(self, namespace, value):
        ns = tuple(namespace) if isinstance(namespace, (tuple, list)) else (namespace,)
        cat = "scalar"
        if isinstance(value, tuple):  # (callable, category) or (object, attribute, category)
            if callable(value[0]):
                fn = value[0]
            else:
                fn = (lambda o=value[0], a=value[1]: getattr(o, a))
            if isinstance(value[-1], str):
                cat = value[-1]
        else:
            fn = value if callable(value) else (lambda v=value: v)
        self._q[ns] = fn
        self._cat[ns] = cat

    def keys(self):
        return list(self._q)

    def remove(self, obj=None, quantities=None, user_name=None):
        """hoomd.logging.Logger.remove: by object (all or some of its quantities) or by namespaces."""
        if obj is not None:
            ns = _namespace(obj) if user_name is None else _namespace(obj)[:-1] + (user_name,)
            keys = [k for k in self._q if k[:-1] == ns and (quantities is None or k[-1] in quantities)]
        else:
            keys = [tuple(q) if isinstance(q, (tuple, list)) else (q,) for q in (quantities or [])]
        for k in keys:
            self._q.pop(k, None)
            self._cat.pop(k, None)

    def _flat(self):
        """{namespace tuple: value}: what the writers consume"""
        return {ns: fn() for ns, fn in self._q.items()}

    def log(self):
        """hoomd.logging.Logger.log: nested dict by namespace, leaves are (value, category)."""
        out = {}
        for ns, fn in self._q.items():
            d = out
            for part in ns[:-1]:
                d = d.setdefault(part, {})
            d[ns[-1]] = (fn(), self._cat.get(ns, "scalar"))
        return out
