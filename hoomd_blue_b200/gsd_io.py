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
    nb = g.chunk(f, "bonds/N")
    nb = int(nb[0, 0]) if nb is not None else 0
    if nb:
        snap.bonds.N = nb
        t = g.chunk(f, "bonds/types")
        snap.bonds.types = ([bytes(row).split(b"\0")[0].decode() for row in t.astype(np.uint8)]
                            if t is not None else ["bondA"])
        tid = g.chunk(f, "bonds/typeid", cur_n=nb)
        if tid is not None:
            snap.bonds.typeid[:] = tid.reshape(-1)
        grp = g.chunk(f, "bonds/group", cur_n=nb)
        if grp is None:
            raise RuntimeError(f"{filename}: bonds/N = {nb} but bonds/group is missing")
        snap.bonds.group[:] = grp.reshape(nb, 2)
    na = g.chunk(f, "angles/N")
    na = int(na[0, 0]) if na is not None else 0
    if na:
        snap.angles.N = na
        t = g.chunk(f, "angles/types")
        snap.angles.types = ([bytes(row).split(b"\0")[0].decode() for row in t.astype(np.uint8)]
                             if t is not None else ["angleA"])
        tid = g.chunk(f, "angles/typeid", cur_n=na)
        if tid is not None:
            snap.angles.typeid[:] = tid.reshape(-1)
        grp = g.chunk(f, "angles/group", cur_n=na)
        if grp is None:
            raise RuntimeError(f"{filename}: angles/N = {na} but angles/group is missing")
        snap.angles.group[:] = grp.reshape(na, 3)
    nd = g.chunk(f, "dihedrals/N")
    nd = int(nd[0, 0]) if nd is not None else 0
    if nd:
        snap.dihedrals.N = nd
        t = g.chunk(f, "dihedrals/types")
        snap.dihedrals.types = ([bytes(row).split(b"\0")[0].decode() for row in t.astype(np.uint8)]
                                if t is not None else ["dihedralA"])
        tid = g.chunk(f, "dihedrals/typeid", cur_n=nd)
        if tid is not None:
            snap.dihedrals.typeid[:] = tid.reshape(-1)
        grp = g.chunk(f, "dihedrals/group", cur_n=nd)
        if grp is None:
            raise RuntimeError(f"{filename}: dihedrals/N = {nd} but dihedrals/group is missing")
        snap.dihedrals.group[:] = grp.reshape(nd, 4)
    return snap, step


def frame_chunks(snap, step, dynamic=None, first=True, log=None):
    """The chunks of one frame of the hoomd schema 1.4 (hoomd/GSDDumpWriter.cc:265-275,480-660: box
    and positions as float32). `first` frames (frame 0 of a file) carry every chunk; later frames carry
    the configuration and the categories named in `dynamic` ('property': position; 'momentum': velocity
    and image; 'attribute': N, types, typeid, mass — hoomd/write/gsd.py:103-151), the rest is read back
    from frame 0. `log`: {name: value} written as log/<name> (float64 / int64 arrays)."""
    dynamic = set(dynamic if dynamic is not None else ("property",))
    p = snap.particles
    n = int(p.N)
    out = [("configuration/step", np.array([[step]], dtype=np.uint64))]
    if first:
        out.append(("configuration/dimensions", np.array([[snap.configuration.dimensions]], dtype=np.uint8)))
    out.append(("configuration/box", np.asarray(snap.configuration.box, dtype=np.float32).reshape(6, 1)))
    if first or "attribute" in dynamic:
        tw = max(len(t) for t in p.types) + 1
        types = np.zeros((len(p.types), tw), dtype=np.int8)
        for i, t in enumerate(p.types):
            types[i, :len(t)] = np.frombuffer(t.encode(), dtype=np.int8)
        out += [("particles/N", np.array([[n]], dtype=np.uint32)), ("particles/types", types),
                ("particles/typeid", np.asarray(p.typeid, dtype=np.uint32).reshape(n, 1)),
                ("particles/mass", np.asarray(p.mass, dtype=np.float32).reshape(n, 1))]
    bonds = getattr(snap, "bonds", None)
    if bonds is not None and bonds.N and (first or "topology" in dynamic):
        tw = max(len(t) for t in bonds.types) + 1
        btypes = np.zeros((len(bonds.types), tw), dtype=np.int8)
        for i, t in enumerate(bonds.types):
            btypes[i, :len(t)] = np.frombuffer(t.encode(), dtype=np.int8)
        out += [("bonds/N", np.array([[bonds.N]], dtype=np.uint32)), ("bonds/types", btypes),
                ("bonds/typeid", np.asarray(bonds.typeid, dtype=np.uint32).reshape(-1, 1)),
                ("bonds/group", np.asarray(bonds.group, dtype=np.int32).reshape(-1, 2))]
    angles = getattr(snap, "angles", None)
    if angles is not None and angles.N and (first or "topology" in dynamic):
        tw = max(len(t) for t in angles.types) + 1
        atypes = np.zeros((len(angles.types), tw), dtype=np.int8)
        for i, t in enumerate(angles.types):
            atypes[i, :len(t)] = np.frombuffer(t.encode(), dtype=np.int8)
        out += [("angles/N", np.array([[angles.N]], dtype=np.uint32)), ("angles/types", atypes),
                ("angles/typeid", np.asarray(angles.typeid, dtype=np.uint32).reshape(-1, 1)),
                ("angles/group", np.asarray(angles.group, dtype=np.int32).reshape(-1, 3))]
    dihedrals = getattr(snap, "dihedrals", None)
    if dihedrals is not None and dihedrals.N and (first or "topology" in dynamic):
        tw = max(len(t) for t in dihedrals.types) + 1
        dtypes = np.zeros((len(dihedrals.types), tw), dtype=np.int8)
        for i, t in enumerate(dihedrals.types):
            dtypes[i, :len(t)] = np.frombuffer(t.encode(), dtype=np.int8)
        out += [("dihedrals/N", np.array([[dihedrals.N]], dtype=np.uint32)), ("dihedrals/types", dtypes),
                ("dihedrals/typeid", np.asarray(dihedrals.typeid, dtype=np.uint32).reshape(-1, 1)),
                ("dihedrals/group", np.asarray(dihedrals.group, dtype=np.int32).reshape(-1, 4))]
    if first or "property" in dynamic:
        out.append(("particles/position", np.asarray(p.position, dtype=np.float32).reshape(n, 3)))
    if first or "momentum" in dynamic:
        out += [("particles/velocity", np.asarray(p.velocity, dtype=np.float32).reshape(n, 3)),
                ("particles/image", np.asarray(p.image, dtype=np.int32).reshape(n, 3))]
    for name, value in (log or {}).items():
        arr = np.atleast_1d(np.asarray(value))
        arr = arr.astype(np.int64 if np.issubdtype(arr.dtype, np.integer) else np.float64)
        out.append(("log/" + name, arr.reshape(arr.shape[0], -1)))
    return out


class GSDWriter:
    """Appending GSD writer (file layout of hoomd/extern/gsd.c, GSD 2.0): chunk data goes to the end of
    the file, the frame's index entries (sorted by name id) follow the earlier frames' in the
    pre-allocated index, new names are appended to the name list; a full index or name list moves to
    the end of the file with twice the room (gsd.c:expand_index / append_name). mode 'wb' creates or
    truncates, 'xb' creates, 'ab' appends to an existing file or creates it."""

    def __init__(self, filename, mode="ab", application="hoomd_blue_b200"):
        import os
        if mode not in ("wb", "xb", "ab"):
            raise ValueError("mode must be 'wb', 'xb' or 'ab'")
        if mode == "xb" and os.path.exists(filename):
            raise FileExistsError(filename)
        self.filename = filename
        if mode == "ab" and os.path.exists(filename):
            g = GSDFile(filename)
            if g.schema != "hoomd" or (g.gsd_version >> 16) != 2:
                raise RuntimeError(f"{filename}: can only append to GSD 2.x files of the hoomd schema")
            with open(filename, "rb") as f:
                (_, self.idx_loc, self.idx_entries, self.nl_loc, self.nl_entries, _, _, app, _,
                 _) = HEADER.unpack(f.read(HEADER.size))
            self.application = app.split(b"\0")[0]
            self.names = list(g.names)
            self.idx_used = len(g.index)
            self.nframes = g.nframes
        else:
            self.application = application.encode()[:63]
            self.idx_entries, self.nl_entries = 128, 128
            self.idx_loc = HEADER.size
            self.nl_loc = self.idx_loc + self.idx_entries * ENTRY.itemsize
            self.names, self.idx_used, self.nframes = [], 0, 0
            with open(filename, "wb") as f:
                f.write(self._header())
                f.write(bytes(self.idx_entries * ENTRY.itemsize + self.nl_entries * 64))

    def _header(self):
        return HEADER.pack(MAGIC, self.idx_loc, self.idx_entries, self.nl_loc, self.nl_entries,
                           make_version(1, 4), make_version(2, 0), self.application, b"hoomd", b"")

    def _nl_bytes(self):
        return b"".join(n.encode() + b"\0" for n in self.names)

    def append(self, chunks):
        """Write one frame: chunks = [(name, 2D array), ...]."""
        with open(self.filename, "r+b") as f:
            new_names = []
            for name, _ in chunks:
                if name not in self.names and name not in new_names:
                    new_names.append(name)
            if new_names:
                self.names += new_names
                nl = self._nl_bytes()
                if len(nl) + 1 > self.nl_entries * 64:
                    while len(nl) + 1 > self.nl_entries * 64:
                        self.nl_entries *= 2
                    f.seek(0, 2)
                    self.nl_loc = f.tell()
                f.seek(self.nl_loc)
                f.write(nl.ljust(self.nl_entries * 64, b"\0"))
            ids = {n: i for i, n in enumerate(self.names)}
            entries = np.zeros(len(chunks), dtype=ENTRY)
            f.seek(0, 2)
            loc = f.tell()
            order = sorted(range(len(chunks)), key=lambda q: ids[chunks[q][0]])
            for slot, q in enumerate(order):
                name, arr = chunks[q]
                arr = np.ascontiguousarray(arr)
                if arr.ndim != 2:
                    raise ValueError(f"chunk {name} must be a 2D array")
                entries[slot] = (self.nframes, arr.shape[0], loc, arr.shape[1], ids[name],
                                 TYPE_ID[arr.dtype], 0)
                f.write(arr.tobytes())
                loc += arr.nbytes
            if self.idx_used + len(entries) + 1 > self.idx_entries:
                f.seek(self.idx_loc)
                old = f.read(self.idx_used * ENTRY.itemsize)
                while self.idx_used + len(entries) + 1 > self.idx_entries:
                    self.idx_entries *= 2
                f.seek(0, 2)
                self.idx_loc = f.tell()
                f.write(old.ljust(self.idx_entries * ENTRY.itemsize, b"\0"))
            f.seek(self.idx_loc + self.idx_used * ENTRY.itemsize)
            f.write(entries.tobytes())
            self.idx_used += len(entries)
            self.nframes += 1
            f.seek(0)
            f.write(self._header())


def write_frames(filename, frames, application="hoomd_blue_b200"):
    """Write a new GSD file holding `frames` = [(snapshot, step), ...], every particle chunk written
    in every frame."""
    w = GSDWriter(filename, "wb", application)
    for snap, step in frames:
        w.append(frame_chunks(snap, step, first=True))


def write_frame(filename, snapshot, step=0):
    write_frames(filename, [(snapshot, step)])
