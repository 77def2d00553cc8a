"""Simulation / State / Snapshot / Operations for the pair-force path.

Mirrors hoomd/simulation.py:33-562 (Simulation: create_state_from_snapshot / _gsd, run, timestep,
tps, walltime, always_compute_pressure), hoomd/snapshot.py (Snapshot.configuration / .particles),
hoomd/state.py (State.get_snapshot, N_particles, box, particle_types) and hoomd/operations.py
(Operations.integrator) — only the members a pair-force simulation touches.
"""
import ctypes as C
import time

import numpy as np

from . import _lib
from . import communicator as _comm
from ._lib import check, ptr


class _Configuration:

    def __init__(self):
        self._box = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0]  # Lx, Ly, Lz, xy, xz, yz
        self.dimensions = 3
        self.step = 0

    @property
    def box(self):
        return self._box

    @box.setter
    def box(self, value):
        """a hoomd.Box, 3 lengths, or 3 lengths + 3 tilt factors; stored as a list of six floats"""
        from .box import Box
        self._box = [float(v) for v in Box.from_box(value)]


class _Particles:
    """hoomd.Snapshot.particles: setting N (re)allocates zeroed arrays (hoomd/snapshot.py)."""

    def __init__(self, pinned=False):
        self.types = ["A"]
        self._N = 0
        self._pinned = pinned
        self._alloc(0)

    def _zeros(self, shape, dtype, fill=0):
        if self._pinned and int(np.prod(shape)) > 0:
            # page-locked host memory (torch is only the allocator here): the engine DMAs such arrays
            # directly instead of going through its bounce buffer
            import torch
            tdt = {np.float32: torch.float32, np.int32: torch.int32, np.uint32: torch.int32}[dtype]
            a = torch.empty(shape, dtype=tdt, pin_memory=True).numpy().view(dtype)
            a[...] = fill
            return a
        return np.full(shape, fill, dtype=dtype)

    def _alloc(self, n):
        self.position = self._zeros((n, 3), np.float32)
        self.velocity = self._zeros((n, 3), np.float32)
        self.image = self._zeros((n, 3), np.int32)
        self.typeid = self._zeros((n,), np.uint32)
        self.mass = self._zeros((n,), np.float32, 1)

    @property
    def N(self):
        return self._N

    @N.setter
    def N(self, n):
        self._N = int(n)
        self._alloc(self._N)


class _Bonds:
    """hoomd.Snapshot.bonds: N, types, typeid (N,), group (N, 2) member tags (hoomd/snapshot.py)."""

    def __init__(self):
        self.types = []
        self._N = 0
        self._alloc(0)

    def _alloc(self, n):
        self.typeid = np.zeros((n,), dtype=np.uint32)
        self.group = np.zeros((n, 2), dtype=np.uint32)

    @property
    def N(self):
        return self._N

    @N.setter
    def N(self, n):
        self._N = int(n)
        self._alloc(self._N)


class _Angles(_Bonds):
    """hoomd.Snapshot.angles: N, types, typeid (N,), group (N, 3) member tags (a, b, c; b the vertex)."""

    def _alloc(self, n):
        self.typeid = np.zeros((n,), dtype=np.uint32)
        self.group = np.zeros((n, 3), dtype=np.uint32)


class _Dihedrals(_Bonds):
    """hoomd.Snapshot.dihedrals: N, types, typeid (N,), group (N, 4) member tags."""

    def _alloc(self, n):
        self.typeid = np.zeros((n,), dtype=np.uint32)
        self.group = np.zeros((n, 4), dtype=np.uint32)


class Snapshot:
    """hoomd.Snapshot (configuration + particles + bonds + angles + dihedrals). pinned=True (not in the reference) allocates the
    particle arrays in page-locked memory so that set_snapshot / get_snapshot copy by DMA."""

    def __init__(self, communicator=None, pinned=False):
        self.configuration = _Configuration()
        self.particles = _Particles(pinned)
        self.bonds = _Bonds()
        self.angles = _Angles()
        self.dihedrals = _Dihedrals()
        self.communicator = communicator

    def replicate(self, nx, ny, nz=1):
        """hoomd.Snapshot.replicate (hoomd/snapshot.py:424-444, SnapshotParticleData::replicate
        hoomd/ParticleData.cc:3742-3810): the box grows by (nx, ny, nz); replica j = (l * ny + m) * nz + n
        holds particle i at index j * N + i, its unwrapped fractional coordinate divided by the
        replication count and shifted by (l/nx, m/ny, n/nz); bonds / angles / dihedrals follow with
        shifted tags. Orthorhombic boxes."""
        nx, ny, nz = int(nx), int(ny), int(nz)
        if min(nx, ny, nz) < 1:
            raise ValueError("replication counts must be positive")
        box = np.asarray(self.configuration.box, dtype=np.float64)
        if np.any(box[3:] != 0):
            raise NotImplementedError("replicate: tilt factors must be zero on this path")
        p = self.particles
        n_old = int(p.N)
        reps = np.array([nx, ny, nz], dtype=np.float64)
        L, Ln = box[:3], box[:3] * reps
        old = {k: getattr(p, k).copy() for k in ("position", "velocity", "image", "typeid", "mass")}
        types = list(p.types)
        unwrapped = old["position"].astype(np.float64) + old["image"] * L
        f = (unwrapped + L / 2) / L
        p.N = n_old * nx * ny * nz
        p.types = types
        j = 0
        for l in range(nx):
            for m in range(ny):
                for n in range(nz):
                    f_new = f / reps + np.array([l, m, n]) / reps
                    q = -Ln / 2 + f_new * Ln
                    img = np.floor((q + Ln / 2) / Ln)
                    q = q - img * Ln
                    q32 = q.astype(np.float32)
                    hi = (Ln / 2).astype(np.float32)
                    over, under = q32 >= hi, q32 < -hi
                    q32 = np.where(over, q32 - Ln.astype(np.float32), q32)
                    q32 = np.where(under, q32 + Ln.astype(np.float32), q32)
                    img = img + over.astype(np.float64) - under.astype(np.float64)
                    sl = slice(j * n_old, (j + 1) * n_old)
                    p.position[sl] = q32
                    p.image[sl] = img.astype(np.int32)
                    p.velocity[sl] = old["velocity"]
                    p.typeid[sl] = old["typeid"]
                    p.mass[sl] = old["mass"]
                    j += 1
        self.configuration.box = [float(Ln[0]), float(Ln[1]), float(Ln[2]), 0.0, 0.0, 0.0]
        for grp in (self.bonds, self.angles, self.dihedrals):
            _replicate_groups(grp, nx * ny * nz, n_old)
        return self

    def wrap(self):
        """hoomd.Snapshot.wrap (hoomd/snapshot.py:446-461): particles back into the box, images updated."""
        L = np.asarray(self.configuration.box[:3], dtype=np.float64)
        p = self.particles
        q = p.position.astype(np.float64)
        img = np.floor((q + L / 2) / L)
        q32 = (q - img * L).astype(np.float32)
        hi = (L / 2).astype(np.float32)
        over = q32 >= hi
        q32 = np.where(over, q32 - L.astype(np.float32), q32)
        p.position[:] = q32
        p.image[:] = p.image + (img + over).astype(np.int32)
        return self

    @classmethod
    def from_arrays(cls, box, position, velocity=None, typeid=None, types=("A",), image=None,
                    mass=None):
        s = cls()
        L = list(np.asarray(box, dtype=np.float64).reshape(-1))
        s.configuration.box = (L + [0.0] * 6)[:6] if len(L) >= 3 else [L[0]] * 3 + [0.0] * 3
        n = len(position)
        s.particles.N = n
        s.particles.types = list(types)
        s.particles.position[:] = position
        if velocity is not None:
            s.particles.velocity[:] = velocity
        if typeid is not None:
            s.particles.typeid[:] = typeid
        if image is not None:
            s.particles.image[:] = image
        if mass is not None:
            s.particles.mass[:] = mass
        return s


def _replicate_groups(grp, n_rep, old_n_particles):
    """BondedGroupData::Snapshot::replicate (hoomd/BondedGroupData.cc:1366-1405): replica j holds the
    groups with every member tag shifted by j * old_n_particles, at indices j * old_size + i."""
    if grp.N == 0:
        return
    tid, g = grp.typeid.copy(), grp.group.astype(np.int64)
    types = list(grp.types)
    grp.N = len(tid) * n_rep
    grp.types = types
    for j in range(n_rep):
        grp.typeid[j * len(tid):(j + 1) * len(tid)] = tid
        grp.group[j * len(tid):(j + 1) * len(tid)] = g + j * old_n_particles


class _BoxView:

    def __init__(self, box6):
        self.Lx, self.Ly, self.Lz, self.xy, self.xz, self.yz = [float(v) for v in box6]

    @property
    def L(self):
        return np.array([self.Lx, self.Ly, self.Lz])

    @property
    def volume(self):
        return self.Lx * self.Ly * self.Lz


class LocalParticles:
    """One rank's particles in local order: capacity-sized arrays, the first n rows are valid."""

    def __init__(self, capacity, pinned=False):
        alloc = _Particles(pinned)._zeros
        self.capacity = int(capacity)
        self.n = 0
        self._full = (alloc((capacity,), np.uint32), alloc((capacity, 3), np.float32),
                      alloc((capacity, 3), np.float32), alloc((capacity,), np.float32),
                      alloc((capacity,), np.uint32), alloc((capacity, 3), np.int32))

    def __iter__(self):  # tag, position, velocity, mass, typeid, image
        return iter(a[:self.n] for a in self._full)


class State:
    """Particle state resident in HBM (hb2_system). With more than one rank in the process group the
    box is split into a 3D grid of domains (hoomd/state.py domain_decomposition,
    hoomd/DomainDecomposition.cc) and this object holds one domain."""

    def __init__(self, simulation, snapshot, domain_decomposition=(None, None, None), local=None):
        L = _lib.lib()
        self._simulation = simulation
        comm = simulation.device.communicator
        snapshot_types, box = None, None
        if snapshot is not None and (comm.rank == 0 or snapshot.particles.N > 0):
            snapshot_types, box = list(snapshot.particles.types), list(snapshot.configuration.box)
        if comm.num_ranks > 1:
            snapshot_types, box = _comm.broadcast_object((snapshot_types, box))
        self._types = snapshot_types
        self._box = [float(v) for v in box]
        h = C.c_void_p()
        check(L.hb2_system_create(C.byref(h), simulation.device.gpu_id, len(self._types),
                                  simulation.seed & 0xFFFF))
        self._h = h
        check(L.hb2_system_set_box(h, *self._box))
        self._grid = (1, 1, 1)
        if comm.num_ranks > 1:
            grid = (C.c_uint32 * 3)()
            if all(g is not None for g in domain_decomposition):
                for k in range(3):
                    grid[k] = int(domain_decomposition[k])
            else:
                check(L.hb2_domain_find_decomposition(comm.num_ranks, *self._box[:3], grid))
            uid = (C.c_uint8 * 128)()
            if comm.rank == 0:
                check(L.hb2_comm_unique_id(uid))
            raw = _comm.broadcast_object(bytes(uid))
            uid = (C.c_uint8 * 128).from_buffer_copy(raw)
            check(L.hb2_system_set_domain(h, grid, comm.rank, comm.num_ranks, uid))
            # peer-memory mailboxes for the one-word reductions of every step (CUDA IPC handles travel like
            # the NCCL id; the engine falls back to NCCL collectively when a peer cannot be mapped)
            ipc = (C.c_uint8 * 64)()
            check(L.hb2_domain_ipc_handle(h, ipc))
            handles = b"".join(_comm.all_gather_arrays(bytes(ipc)))
            check(L.hb2_domain_set_peers(h, handles, comm.num_ranks))
            self._grid = tuple(int(g) for g in grid)
        if local is not None:
            self._set_tagged(*local)
        else:
            self.set_snapshot(snapshot)

    @property
    def domain_decomposition(self):
        return self._grid

    def _set_tagged(self, n_global, tag, pos, vel, typeid=None, mass=None, image=None):
        L = _lib.lib()
        n = len(pos)
        tag = np.ascontiguousarray(tag, dtype=np.uint32)
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        vel = np.ascontiguousarray(vel, dtype=np.float32)
        tid = np.ascontiguousarray(typeid if typeid is not None else np.zeros(n), dtype=np.uint32)
        mass = np.ascontiguousarray(mass if mass is not None else np.ones(n), dtype=np.float32)
        img = np.ascontiguousarray(image if image is not None else np.zeros((n, 3)), dtype=np.int32)
        check(L.hb2_system_set_particles_tagged(self._h, n, int(n_global), ptr(tag, C.c_uint32),
                                                ptr(pos, C.c_float), ptr(vel, C.c_float),
                                                ptr(mass, C.c_float), ptr(tid, C.c_uint32),
                                                ptr(img, C.c_int32)))

    def set_snapshot(self, snapshot):
        L = _lib.lib()
        comm = self._simulation.device.communicator
        if comm.num_ranks > 1:
            # the reference scatters rank 0's snapshot (hoomd/ParticleData.cc initializeFromSnapshot);
            # here every rank receives it and the engine keeps the particles it owns
            p = snapshot.particles if comm.rank == 0 else None
            b = getattr(snapshot, "bonds", None) if comm.rank == 0 else None
            a = getattr(snapshot, "angles", None) if comm.rank == 0 else None
            dh = getattr(snapshot, "dihedrals", None) if comm.rank == 0 else None
            data = _comm.broadcast_object(
                None if p is None else (p.position, p.velocity, p.typeid, p.mass, p.image,
                                        None if b is None or not b.N else (list(b.types), b.typeid, b.group),
                                        None if a is None or not a.N else (list(a.types), a.typeid, a.group),
                                        None if dh is None or not dh.N else (list(dh.types), dh.typeid, dh.group)))
            pos, vel, tid, mass, img, bd, ad, dd = data
            self._set_tagged(len(pos), np.arange(len(pos), dtype=np.uint32), pos, vel, tid, mass, img)
            bonds = None
            if bd is not None:
                bonds = _Bonds()
                bonds.N = len(bd[1])
                bonds.types, bonds.typeid[:], bonds.group[:] = bd[0], bd[1], bd[2]
            self._set_bonds(bonds)
            angles = None
            if ad is not None:
                angles = _Angles()
                angles.N = len(ad[1])
                angles.types, angles.typeid[:], angles.group[:] = ad[0], ad[1], ad[2]
            self._set_angles(angles)
            dihedrals = None
            if dd is not None:
                dihedrals = _Dihedrals()
                dihedrals.N = len(dd[1])
                dihedrals.types, dihedrals.typeid[:], dihedrals.group[:] = dd[0], dd[1], dd[2]
            self._set_dihedrals(dihedrals)
            return
        p = snapshot.particles
        pos = np.ascontiguousarray(p.position, dtype=np.float32)
        vel = np.ascontiguousarray(p.velocity, dtype=np.float32)
        mass = np.ascontiguousarray(p.mass, dtype=np.float32)
        tid = np.ascontiguousarray(p.typeid, dtype=np.uint32)
        img = np.ascontiguousarray(p.image, dtype=np.int32)
        check(L.hb2_system_set_particles(self._h, pos.shape[0], ptr(pos, C.c_float),
                                         ptr(vel, C.c_float), ptr(mass, C.c_float),
                                         ptr(tid, C.c_uint32), ptr(img, C.c_int32)))
        self._set_bonds(getattr(snapshot, "bonds", None))
        self._set_angles(getattr(snapshot, "angles", None))
        self._set_dihedrals(getattr(snapshot, "dihedrals", None))

    def _set_dihedrals(self, dih):
        L = _lib.lib()
        n = 0 if dih is None else int(dih.N)
        self._dihedral_types = [] if dih is None else list(dih.types)
        self._dihedral_typeid = np.zeros(0, np.uint32) if n == 0 else np.ascontiguousarray(dih.typeid, np.uint32)
        self._dihedral_group = np.zeros((0, 4), np.uint32) if n == 0 else np.ascontiguousarray(dih.group, np.uint32)
        if n == 0 and not getattr(self, "_had_dihedrals", False):
            return
        self._had_dihedrals = True
        check(L.hb2_system_set_dihedrals(self._h, n, ptr(self._dihedral_group, C.c_uint32),
                                         ptr(self._dihedral_typeid, C.c_uint32), len(self._dihedral_types)))

    def _set_angles(self, angles):
        L = _lib.lib()
        n = 0 if angles is None else int(angles.N)
        self._angle_types = [] if angles is None else list(angles.types)
        self._angle_typeid = np.zeros(0, np.uint32) if n == 0 else np.ascontiguousarray(angles.typeid, np.uint32)
        self._angle_group = np.zeros((0, 3), np.uint32) if n == 0 else np.ascontiguousarray(angles.group, np.uint32)
        if n == 0 and not getattr(self, "_had_angles", False):
            return
        self._had_angles = True
        check(L.hb2_system_set_angles(self._h, n, ptr(self._angle_group, C.c_uint32),
                                      ptr(self._angle_typeid, C.c_uint32), len(self._angle_types)))

    def _set_bonds(self, bonds):
        """BondData of the snapshot (single GPU): the table is kept on the device keyed by tag."""
        L = _lib.lib()
        n = 0 if bonds is None else int(bonds.N)
        self._bond_types = [] if bonds is None else list(bonds.types)
        self._bond_typeid = np.zeros(0, np.uint32) if n == 0 else np.ascontiguousarray(bonds.typeid, np.uint32)
        self._bond_group = np.zeros((0, 2), np.uint32) if n == 0 else np.ascontiguousarray(bonds.group, np.uint32)
        if n == 0 and not getattr(self, "_had_bonds", False):
            return
        self._had_bonds = True
        check(L.hb2_system_set_bonds(self._h, n, ptr(self._bond_group, C.c_uint32),
                                     ptr(self._bond_typeid, C.c_uint32), len(self._bond_types)))

    def _get_local(self, pinned=False, into=None):
        """LocalParticles (tag, position, velocity, mass, typeid, image — also unpackable as that
        tuple) of the particles this rank owns, in local order. pinned=True allocates the arrays
        page-locked (direct DMA in both directions); `into` is a previous result whose storage is
        reused when large enough (page-locked allocations are expensive, and the owned count drifts
        with migration)."""
        L = _lib.lib()
        n = int(L.hb2_system_n(self._h))
        if into is not None and into.capacity >= n:
            out = into
        else:
            out = LocalParticles(n + n // 16 + 1024, pinned)
        out.n = n
        tag, pos, vel, mass, tid, img = out
        check(L.hb2_system_get_local(self._h, ptr(tag, C.c_uint32), ptr(pos, C.c_float),
                                     ptr(vel, C.c_float), ptr(mass, C.c_float), ptr(tid, C.c_uint32),
                                     ptr(img, C.c_int32)))
        return out

    def get_snapshot(self, into=None):
        """hoomd.State.get_snapshot. `into` (not in the reference) reuses an existing snapshot's
        arrays — e.g. a pinned one — instead of allocating."""
        L = _lib.lib()
        n = self.N_particles
        s = into if into is not None else Snapshot()
        s.configuration.box = list(self._box)
        if s.particles.N != n:
            s.particles.N = n
        s.particles.types = list(self._types)
        if getattr(self, "_bond_types", None):
            s.bonds.N = len(self._bond_typeid)
            s.bonds.types = list(self._bond_types)
            s.bonds.typeid[:] = self._bond_typeid
            s.bonds.group[:] = self._bond_group
        if getattr(self, "_dihedral_types", None):
            s.dihedrals.N = len(self._dihedral_typeid)
            s.dihedrals.types = list(self._dihedral_types)
            s.dihedrals.typeid[:] = self._dihedral_typeid
            s.dihedrals.group[:] = self._dihedral_group
        if getattr(self, "_angle_types", None):
            s.angles.N = len(self._angle_typeid)
            s.angles.types = list(self._angle_types)
            s.angles.typeid[:] = self._angle_typeid
            s.angles.group[:] = self._angle_group
        p = s.particles
        if self._simulation.device.communicator.num_ranks > 1:
            # gathered by tag (hoomd/ParticleData.cc takeSnapshot does the same with MPI gathers;
            # here every rank receives the global snapshot)
            tag, pos, vel, mass, tid, img = self._get_local()
            p.position[:] = _comm.gather_by_tag(n, tag, pos)
            p.velocity[:] = _comm.gather_by_tag(n, tag, vel)
            p.mass[:] = _comm.gather_by_tag(n, tag, mass)
            p.typeid[:] = _comm.gather_by_tag(n, tag, tid)
            p.image[:] = _comm.gather_by_tag(n, tag, img)
            return s
        check(L.hb2_system_get_particles(self._h, ptr(p.position, C.c_float),
                                         ptr(p.velocity, C.c_float), ptr(p.mass, C.c_float),
                                         ptr(p.typeid, C.c_uint32), ptr(p.image, C.c_int32)))
        return s

    def thermalize_particle_momenta(self, filter, kT):
        """hoomd.State.thermalize_particle_momenta (hoomd/state.py:759-800): Gaussian velocities at kT
        from every particle's own counter-based stream (keyed by tag, the simulation seed and the
        current timestep), centre-of-mass momentum removed. filter must be All()."""
        from .filter import All
        from . import rng
        if not isinstance(filter, All):
            raise NotImplementedError("this path thermalizes filter.All() only")
        snap = self.get_snapshot()
        n = snap.particles.N
        v = rng.thermal_velocities(np.arange(n, dtype=np.uint32), snap.particles.mass, float(kT),
                                   self._simulation.seed, self._simulation.timestep)
        snap.particles.velocity[:] = v.astype(np.float32)
        self.set_snapshot(snap)

    @property
    def N_particles(self):
        """Global particle count (hoomd/state.py:N_particles)."""
        return int(_lib.lib().hb2_system_n_global(self._h))

    @property
    def _N_local(self):
        return int(_lib.lib().hb2_system_n(self._h))

    @property
    def _N_ghost(self):
        return int(_lib.lib().hb2_system_n_ghost(self._h))

    @property
    def particle_types(self):
        return list(self._types)

    @property
    def box(self):
        return _BoxView(self._box)

    def _refresh_box(self):
        """Box lengths back from the engine (a ConstantPressure run rescales them)."""
        import ctypes as C
        L3 = (C.c_double * 3)()
        check(_lib.lib().hb2_system_get_box(self._h, L3))
        self._box = [L3[0], L3[1], L3[2]] + list(self._box[3:])

    def _destroy(self):
        """Free the device state now (every rank of a decomposed system calls it)."""
        if getattr(self, "_h", None):
            _lib.lib().hb2_system_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass


class Operations:
    """hoomd.Operations restricted to `integrator`, `writers` and `computes` (hoomd/operations.py:32)."""

    def __init__(self, simulation):
        self._simulation = simulation
        self._integrator = None
        self.writers = []
        self.computes = []

    def __iadd__(self, op):
        self.add(op)
        return self

    def add(self, op):
        from .write import Writer
        from .md.compute import ThermodynamicQuantities
        if isinstance(op, Writer):
            self.writers.append(op)
        elif isinstance(op, ThermodynamicQuantities):
            self.computes.append(op)
            op._attach(self._simulation)
        elif hasattr(op, "_run"):
            self.integrator = op
        else:
            raise TypeError(f"cannot add {type(op).__name__} to Operations")

    @property
    def integrator(self):
        return self._integrator

    @integrator.setter
    def integrator(self, op):
        if self._integrator is not None and self._integrator is not op:
            self._integrator._detach()
        self._integrator = op


class Simulation:
    _loggables = {"timestep": "scalar", "tps": "scalar", "walltime": "scalar"}

    def __init__(self, device, seed=None):
        from .device import GPU
        if not isinstance(device, GPU):
            raise TypeError("device must be hoomd_blue_b200.device.GPU")
        self.device = device
        self.seed = 0 if seed is None else int(seed)
        self.state = None
        self.operations = Operations(self)
        self._always_compute_pressure = False
        self._initial_timestep = 0
        self._tps = 0.0
        self._walltime = 0.0
        self._last_device_ms = 0.0

    # --- state creation (hoomd/simulation.py:188-301)
    def create_state_from_snapshot(self, snapshot, domain_decomposition=(None, None, None)):
        if self.state is not None:
            raise RuntimeError("Cannot initialize more than once\n")
        self.state = State(self, snapshot, domain_decomposition)
        check(_lib.lib().hb2_system_set_timestep(self.state._h, self._initial_timestep))
        if self._always_compute_pressure:
            check(_lib.lib().hb2_system_set_compute_virial(self.state._h, 1))

    def _create_state_from_local(self, box, types, n_global, tag, position, velocity, typeid=None,
                                 mass=None, image=None, domain_decomposition=(None, None, None)):
        """Every rank passes its own share of a large system (any superset of what it owns; the
        engine keeps the owned particles). Not in the reference's API: its snapshots always
        originate on rank 0, which does not scale to the 64 M-particle configuration."""
        if self.state is not None:
            raise RuntimeError("Cannot initialize more than once\n")
        snap = Snapshot()
        snap.configuration.box = [float(v) for v in box]
        snap.particles.types = list(types)
        self.state = State(self, snap, domain_decomposition,
                           local=(n_global, tag, position, velocity, typeid, mass, image))
        check(_lib.lib().hb2_system_set_timestep(self.state._h, self._initial_timestep))
        if self._always_compute_pressure:
            check(_lib.lib().hb2_system_set_compute_virial(self.state._h, 1))

    def create_state_from_gsd(self, filename, frame=-1, domain_decomposition=(None, None, None)):
        from . import gsd_io
        snap, step = gsd_io.read_frame(filename, frame)
        if self.state is None and self._initial_timestep == 0:
            self._initial_timestep = step
        self.create_state_from_snapshot(snap, domain_decomposition)

    @property
    def timestep(self):
        if self.state is None:
            return self._initial_timestep
        return int(_lib.lib().hb2_system_timestep(self.state._h))

    @timestep.setter
    def timestep(self, step):
        if self.state is not None:
            raise RuntimeError("State must not be set to change timestep.")  # simulation.py:99-105
        self._initial_timestep = int(step)

    @property
    def always_compute_pressure(self):
        return self._always_compute_pressure

    @always_compute_pressure.setter
    def always_compute_pressure(self, value):
        self._always_compute_pressure = bool(value)
        if self.state is not None:
            check(_lib.lib().hb2_system_set_compute_virial(self.state._h, int(bool(value))))

    @property
    def tps(self):
        return self._tps

    @property
    def walltime(self):
        return self._walltime

    def run(self, steps, write_at_start=False):
        if self.state is None:
            raise RuntimeError("Cannot run before state is set.")
        integ = self.operations.integrator
        if integ is None:
            raise RuntimeError("No integrator is set")
        integ._attach(self)
        for c in self.operations.computes:
            c._attach(self)
        writers = list(self.operations.writers)
        t0 = time.perf_counter()
        device_ms = 0.0
        if not writers:
            integ._run(int(steps))
            device_ms = integ._last_run_ms()
        else:
            # System::run (hoomd/System.cc:80-181): writers fire after the step that brings the
            # timestep to a triggered value (and before the first step with write_at_start). The
            # step loop itself stays on the device between two triggered steps.
            now = self.timestep
            end = now + int(steps)
            if write_at_start:
                for w in writers:
                    if w.trigger(now):
                        w._write(self)
            first = True
            while now < end:
                nxt = [w.trigger.next_after(now) for w in writers]
                nxt = min([n for n in nxt if n is not None] + [end])
                nxt = min(nxt, end)
                integ._run(nxt - now, first_segment=first)
                first = False
                device_ms += integ._last_run_ms()
                now = nxt
                for w in writers:
                    if w.trigger(now):
                        w._write(self)
            for w in writers:
                w.flush()
        self._walltime = time.perf_counter() - t0
        self._tps = steps / self._walltime if self._walltime > 0 else 0.0
        self._last_device_ms = device_ms
