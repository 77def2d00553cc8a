"""GSD files with the `hoomd` schema: the frame reader behind Simulation.create_state_from_gsd and a
single-call frame writer, for the chunks this path uses.

File format (hoomd/extern/gsd.h:126-187, GSD file version 2.0): a 256-byte header {magic, index
location, index entries allocated, name-list location, name-list size / 64, schema version, gsd version,
application[64], schema[64], reserved[80]}; an index of 32-byte entries {frame u64, N u64, location i64,
M u32, id u16, type u8, flags u8} sorted by (frame, id), unused entries zero; a name list of
NUL-terminated names (id = ordinal); chunk data anywhere else in the file.
Reader semantics (hoomd/GSDReader.cc:100-155,187-243): a chunk missing in frame f is taken from frame 0
when its N matches, else the default; box/position/velocity are float32, image int32, typeid uint32,
types an N x M int8 table of NUL-padded strings. Checked in tests/test_gsd_io.py against files written
and read by the reference's own gsd.c (compiled in place into oracle/_ref/ where /root/reference exists)
and against a committed fixture that library wrote.
"""
import struct

import numpy as np

MAGIC = 0x65DF65DF65DF65DF
HEADER = struct.Struct("<QQQQQII64s64s80s")
ENTRY = np.dtype([("frame", "<u8"), ("N", "<u8"), ("location", "<i8"), ("M", "<u4"), ("id", "<u2"),
                  ("type", "u1"), ("flags", "u1")])
TYPES = {1: np.uint8, 2: np.uint16, 3: np.uint32, 4: np.uint64, 5: np.int8, 6: np.int16, 7: np.int32,
         8: np.int64, 9: np.float32, 10: np.float64}
TYPE_ID = {np.dtype(v): k for k, v in TYPES.items()}


def make_version(major, minor):
    return (major << 16) | minor


class GSDFile:
    """Read-only view of a GSD file: index, names, chunks."""

    def __init__(self, filename):
        with open(filename, "rb") as f:
            self._buf = f.read()
        if len(self._buf) < HEADER.size:
            raise RuntimeError(f"{filename}: not a GSD file")
        (magic, idx_loc, idx_n, nl_loc, nl_n, self.schema_version, self.gsd_version, app, schema,
         _) = HEADER.unpack_from(self._buf, 0)
        if magic != MAGIC:
            raise RuntimeError(f"{filename}: not a GSD file")
        major = self.gsd_version >> 16
        if major not in (1, 2) and self.gsd_version != make_version(0, 3):
            raise RuntimeError(f"{filename}: unsupported GSD file version {major}")
        self.application = app.split(b"\0")[0].decode()
        self.schema = schema.split(b"\0")[0].decode()
        index = np.frombuffer(self._buf, dtype=ENTRY, count=idx_n, offset=idx_loc)
        used = int(np.argmax(index["location"] == 0)) if (index["location"] == 0).any() else idx_n
        self.index = index[:used]
        raw = self._buf[nl_loc:nl_loc + nl_n * 64]
        if major >= 2:
            names = raw.split(b"\0")
            self.names = [n.decode() for n in names[:names.index(b"")] if n] if b"" in names else \
                [n.decode() for n in names]
        else:
            self.names = []
            for k in range(nl_n):
                n = raw[64 * k:64 * k + 64].split(b"\0")[0]
                if not n:
                    break
                self.names.append(n.decode())
        self._id = {n: i for i, n in enumerate(self.names)}
        self.nframes = int(self.index["frame"].max()) + 1 if len(self.index) else 0

    def find(self, frame, name):
        i = self._id.get(name)
        if i is None:
            return None
        sel = np.nonzero((self.index["frame"] == frame) & (self.index["id"] == i))[0]
        return self.index[sel[-1]] if len(sel) else None

    def read(self, entry):
        dt = np.dtype(TYPES[int(entry["type"])])
        n, m = int(entry["N"]), int(entry["M"])
        a = np.frombuffer(self._buf, dtype=dt, count=n * m, offset=int(entry["location"]))
        return a.reshape(n, m).copy()

    def chunk(self, frame, name, cur_n=0):
        """hoomd/GSDReader.cc:100-155: frame f, else frame 0 (only if N matches when cur_n is given)."""
        e = self.find(frame, name)
        if e is None and frame != 0:
            e = self.find(0, name)
        if e is None or (cur_n and int(e["N"]) != cur_n):
            return None
        return self.read(e)


def read_frame(filename, frame=-1):
    """(Snapshot, step) of one frame; negative frames count from the end
    (hoomd/simulation.py:188-253, hoomd/GSDReader.cc:54-72)."""
    from .simulation import Snapshot
    g = GSDFile(filename)
    if g.schema != "hoomd":
        raise RuntimeError(f"{filename}: schema is '{g.schema}', expected 'hoomd'")
    if g.schema_version >= make_version(2, 0):
        raise RuntimeError(f"{filename}: unsupported hoomd schema version")
    nframes = g.nframes
    f = frame + nframes if frame < 0 else frame
    if f < 0 or f >= nframes:
        raise RuntimeError(f"Cannot read frame {frame} {filename} only has {nframes} frames.")
    snap = Snapshot()
    c = g.chunk(f, "configuration/step")
    step = int(c[0, 0]) if c is not None else 0
    snap.configuration.step = step
    c = g.chunk(f, "configuration/dimensions")
    snap.configuration.dimensions = int(c[0, 0]) if c is not None else 3
    c = g.chunk(f, "configuration/box")
    snap.configuration.box = [float(v) for v in c.ravel()] if c is not None else [1, 1, 1, 0, 0, 0]
    c = g.chunk(f, "particles/N")
    n = int(c[0, 0]) if c is not None else 0
    snap.particles.N = n
    c = g.chunk(f, "particles/types")
    if c is not None:
        snap.particles.types = [bytes(row.astype(np.uint8)).split(b"\0")[0].decode() for row in c]
    p = snap.particles
    for name, arr in (("typeid", p.typeid), ("mass", p.mass), ("position", p.position),
                      ("velocity", p.velocity), ("image", p.image)):
        c = g.chunk(f, "particles/" + name, cur_n=n)
        if c is not None:
            arr[...] = c.reshape(arr.shape)
    return snap, step


def write_frames(filename, frames, application="hoomd_blue_b200"):
    """Write a new GSD file holding `frames` = [(snapshot, step), ...] with the hoomd schema 1.4
    (hoomd/GSDDumpWriter.cc:265-275,480-660: box and positions as float32, every particle chunk written
    in every frame)."""
    chunks = []  # (frame, name, array2d)
    for k, (snap, step) in enumerate(frames):
        p = snap.particles
        n = int(p.N)
        tw = max(len(t) for t in p.types) + 1
        types = np.zeros((len(p.types), tw), dtype=np.int8)
        for i, t in enumerate(p.types):
            types[i, :len(t)] = np.frombuffer(t.encode(), dtype=np.int8)
        chunks += [
            (k, "configuration/step", np.array([[step]], dtype=np.uint64)),
            (k, "configuration/dimensions", np.array([[snap.configuration.dimensions]], dtype=np.uint8)),
            (k, "configuration/box", np.asarray(snap.configuration.box, dtype=np.float32).reshape(6, 1)),
            (k, "particles/N", np.array([[n]], dtype=np.uint32)),
            (k, "particles/types", types),
            (k, "particles/typeid", np.asarray(p.typeid, dtype=np.uint32).reshape(n, 1)),
            (k, "particles/mass", np.asarray(p.mass, dtype=np.float32).reshape(n, 1)),
            (k, "particles/position", np.asarray(p.position, dtype=np.float32).reshape(n, 3)),
            (k, "particles/velocity", np.asarray(p.velocity, dtype=np.float32).reshape(n, 3)),
            (k, "particles/image", np.asarray(p.image, dtype=np.int32).reshape(n, 3)),
        ]
    names = []
    for _, name, _ in chunks:
        if name not in names:
            names.append(name)
    nl_bytes = b"".join(n.encode() + b"\0" for n in names)
    nl_entries = max(16, (len(nl_bytes) + 1 + 63) // 64)
    idx_entries = max(128, len(chunks) + 1)
    idx_loc = HEADER.size
    nl_loc = idx_loc + idx_entries * ENTRY.itemsize
    data_loc = nl_loc + nl_entries * 64
    index = np.zeros(idx_entries, dtype=ENTRY)
    order = sorted(range(len(chunks)), key=lambda q: (chunks[q][0], names.index(chunks[q][1])))
    blobs = []
    loc = data_loc
    for slot, q in enumerate(order):
        frame, name, arr = chunks[q]
        arr = np.ascontiguousarray(arr)
        index[slot] = (frame, arr.shape[0], loc, arr.shape[1], names.index(name), TYPE_ID[arr.dtype], 0)
        blobs.append(arr.tobytes())
        loc += arr.nbytes
    header = HEADER.pack(MAGIC, idx_loc, idx_entries, nl_loc, nl_entries, make_version(1, 4),
                         make_version(2, 0), application.encode()[:63], b"hoomd", b"")
    with open(filename, "wb") as f:
        f.write(header)
        f.write(index.tobytes())
        f.write(nl_bytes.ljust(nl_entries * 64, b"\0"))
        for b in blobs:
            f.write(b)


def write_frame(filename, snapshot, step=0):
    write_frames(filename, [(snapshot, step)])
