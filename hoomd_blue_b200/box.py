"""hoomd.Box for this path (hoomd/box.py): the six box parameters with the reference's constructors and
properties. The engine takes orthorhombic boxes only (tilt factors must be zero when a state is made)."""
import numpy as np


class Box:

    def __init__(self, Lx, Ly, Lz=0, xy=0, xz=0, yz=0):
        self.Lx, self.Ly, self.Lz = float(Lx), float(Ly), float(Lz)
        self.xy, self.xz, self.yz = float(xy), float(xz), float(yz)

    @classmethod
    def cube(cls, L):
        return cls(L, L, L)

    @classmethod
    def square(cls, L):
        return cls(L, L, 0)

    @classmethod
    def from_box(cls, box):
        """Box, [Lx, Ly, Lz], [Lx, Ly, Lz, xy, xz, yz], a dict with those keys, or an object with
        Lx/Ly/Lz attributes (hoomd/box.py Box.from_box)."""
        if isinstance(box, Box):
            return cls(*box)
        if isinstance(box, dict):
            return cls(**{k: box[k] for k in ("Lx", "Ly", "Lz", "xy", "xz", "yz") if k in box})
        if hasattr(box, "Lx"):
            return cls(box.Lx, box.Ly, getattr(box, "Lz", 0), getattr(box, "xy", 0), getattr(box, "xz", 0),
                       getattr(box, "yz", 0))
        v = [float(x) for x in np.asarray(box, dtype=np.float64).reshape(-1)]
        if len(v) not in (3, 6):
            raise ValueError("a box is given by 3 lengths or by 3 lengths and 3 tilt factors")
        return cls(*v)

    def __iter__(self):
        return iter((self.Lx, self.Ly, self.Lz, self.xy, self.xz, self.yz))

    def __eq__(self, other):
        return isinstance(other, Box) and tuple(self) == tuple(other)

    def __repr__(self):
        return ("hoomd.box.Box(Lx={}, Ly={}, Lz={}, xy={}, xz={}, yz={})".format(*self))

    @property
    def L(self):
        return np.array([self.Lx, self.Ly, self.Lz])

    @L.setter
    def L(self, value):
        v = np.broadcast_to(np.asarray(value, dtype=np.float64), (3,))
        self.Lx, self.Ly, self.Lz = (float(x) for x in v)

    @property
    def tilts(self):
        return np.array([self.xy, self.xz, self.yz])

    @property
    def is2D(self):
        return self.Lz == 0

    @property
    def dimensions(self):
        return 2 if self.is2D else 3

    @property
    def volume(self):
        """volume (3D) or area (2D)"""
        return self.Lx * self.Ly * (1.0 if self.is2D else self.Lz)

    @property
    def periodic(self):
        return (True, True, True)
