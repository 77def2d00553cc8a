"""ctypes binding of the reference's GSD layer (oracle/_ref/libgsd_ref.so = hoomd/extern/gsd.c compiled
in place). Test infrastructure only."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libgsd_ref.so")
_TYPE = {np.dtype(np.uint8): 1, np.dtype(np.uint16): 2, np.dtype(np.uint32): 3, np.dtype(np.uint64): 4,
         np.dtype(np.int8): 5, np.dtype(np.int16): 6, np.dtype(np.int32): 7, np.dtype(np.int64): 8,
         np.dtype(np.float32): 9, np.dtype(np.float64): 10}
_NP = {v: k for k, v in _TYPE.items()}


class _Entry(C.Structure):  # hoomd/extern/gsd.h:157-176
    _fields_ = [("frame", C.c_uint64), ("N", C.c_uint64), ("location", C.c_int64), ("M", C.c_uint32),
                ("id", C.c_uint16), ("type", C.c_uint8), ("flags", C.c_uint8)]


def available():
    return os.path.exists(LIB)


class RefGSD:

    def __init__(self):
        self.lib = C.CDLL(LIB)
        self.lib.gsd_find_chunk.restype = C.POINTER(_Entry)
        self.lib.gsd_find_chunk.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p]
        self.lib.gsd_read_chunk.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(_Entry)]
        self.lib.gsd_write_chunk.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_uint64, C.c_uint32,
                                             C.c_uint8, C.c_void_p]
        self.lib.gsd_get_nframes.restype = C.c_uint64
        self.lib.gsd_get_nframes.argtypes = [C.c_void_p]
        self.lib.gsd_make_version.restype = C.c_uint32
        self.h = C.create_string_buffer(16384)  # struct gsd_handle, opaque (a few hundred bytes)

    @classmethod
    def create(cls, path):
        g = cls()
        rc = g.lib.gsd_create_and_open(g.h, path.encode(), b"hoomd-reference-gsd", b"hoomd",
                                       g.lib.gsd_make_version(1, 4), 1, 0)  # GSD_OPEN_READWRITE
        assert rc == 0, rc
        return g

    @classmethod
    def open(cls, path):
        g = cls()
        rc = g.lib.gsd_open(g.h, path.encode(), 2)  # GSD_OPEN_READONLY
        if rc != 0:
            raise RuntimeError(f"gsd_open failed: {rc}")
        return g

    def write(self, name, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.ndim == 2
        rc = self.lib.gsd_write_chunk(self.h, name.encode(), _TYPE[arr.dtype], arr.shape[0],
                                      arr.shape[1], 0, arr.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc

    def end_frame(self):
        assert self.lib.gsd_end_frame(self.h) == 0

    def nframes(self):
        return int(self.lib.gsd_get_nframes(self.h))

    def read(self, frame, name):
        e = self.lib.gsd_find_chunk(self.h, frame, name.encode())
        if not e:
            return None
        ent = e.contents
        out = np.zeros((ent.N, ent.M), dtype=_NP[ent.type])
        assert self.lib.gsd_read_chunk(self.h, out.ctypes.data_as(C.c_void_p), e) == 0
        return out

    def close(self):
        self.lib.gsd_close(self.h)
