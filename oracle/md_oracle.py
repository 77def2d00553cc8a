"""ctypes front-end of oracle/md_oracle.c (fp64, FMA-free restatement of HOOMD-blue v4.8.2's
pair-force step).  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package.

Reference citations live beside each C function in md_oracle.h / md_oracle.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libmd_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "md_oracle.c")
    if (force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_build/libmd_oracle.so"])
    return _LIB


class Box(C.Structure):
    _fields_ = [("L", C.c_double * 3), ("lo", C.c_double * 3), ("hi", C.c_double * 3),
                ("Linv", C.c_double * 3), ("periodic", C.c_int * 3)]


class RunCfg(C.Structure):
    _fields_ = [("n", C.c_uint32), ("ntypes", C.c_uint32), ("dt", C.c_double),
                ("r_buff", C.c_double), ("kT", C.c_double), ("tau", C.c_double),
                ("thermostat", C.c_int), ("potential", C.c_int), ("shift_mode", C.c_int),
                ("seed", C.c_uint16), ("timestep", C.c_uint64), ("gamma", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.orc_nlist_binned.restype = C.POINTER(C.c_uint32)
        L.orc_nlist_brute.restype = C.POINTER(C.c_uint32)
        L.orc_dpd_alpha.restype = C.c_double
        L.orc_dpd_alpha.argtypes = [C.c_uint16, C.c_uint32, C.c_uint32, C.c_uint64]
        L.orc_kinetic_energy.restype = C.c_double
        L.orc_bussi_factor.restype = C.c_double
        L.orc_bussi_factor.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                       C.c_uint64, C.c_uint16, C.c_uint32]
        L.orc_lj_tail_correction.restype = C.c_double
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_cell_adj_width.restype = C.c_uint32
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def make_box(Lxyz):
    b = Box()
    lib().orc_box_set(C.byref(b), C.c_double(Lxyz[0]), C.c_double(Lxyz[1]), C.c_double(Lxyz[2]))
    return b


def type_matrix(vals, ntypes):
    """Scalar or (nt,nt) -> flat fp64 array indexed tj*nt+ti (hoomd/Index1D.h:53)."""
    a = np.empty((ntypes, ntypes), dtype=np.float64)
    a[...] = vals
    return np.ascontiguousarray(a.reshape(-1))


def min_image(box, dx):
    dx = _f64(dx).copy()
    for row in dx.reshape(-1, 3):
        lib().orc_min_image(C.byref(box), _p(row, C.c_double))
    return dx


def cell_dims(box, width, ghost_width=None):
    dim = (C.c_uint32 * 3)()
    gw = (C.c_double * 3)(*(ghost_width if ghost_width is not None else (0, 0, 0)))
    lib().orc_cell_dims(C.byref(box), C.c_double(width), gw, dim)
    return tuple(dim)


def cell_adj(dim):
    d = (C.c_uint32 * 3)(*dim)
    w = lib().orc_cell_adj_width(d)
    adj = np.zeros((dim[0] * dim[1] * dim[2], w), dtype=np.uint32)
    lib().orc_cell_adj(d, _p(adj, C.c_uint32))
    return adj


def cell_list(box, dim, pos, Nmax, n_local=None, ghost_width=None):
    pos = _f64(pos)
    n = pos.shape[0]
    n_local = n if n_local is None else n_local
    d = (C.c_uint32 * 3)(*dim)
    gw = (C.c_double * 3)(*(ghost_width if ghost_width is not None else (0, 0, 0)))
    Nc = dim[0] * dim[1] * dim[2]
    size = np.zeros(Nc, dtype=np.uint32)
    idx = np.full((Nc, Nmax), 0xFFFFFFFF, dtype=np.uint32)
    cond = (C.c_uint32 * 3)()
    lib().orc_cell_list(C.byref(box), d, gw, _p(pos, C.c_double), C.c_uint32(n_local), C.c_uint32(n),
                        C.c_uint32(Nmax), _p(size, C.c_uint32), _p(idx, C.c_uint32), cond)
    return size, idx, tuple(cond)


def nlist(box, pos, types, r_cut, r_buff, full=True, brute=False, n_local=None, ghost_width=None):
    """Returns (n_neigh, head, nlist, dim). r_cut: flat ntypes^2 matrix."""
    pos = _f64(pos)
    types = _u32(types)
    r_cut = _f64(r_cut)
    nt = int(round(np.sqrt(r_cut.size)))
    n = pos.shape[0]
    n_local = n if n_local is None else n_local
    n_neigh = np.zeros(n_local, dtype=np.uint32)
    head = np.zeros(n_local, dtype=np.uint64)
    total = C.c_uint64(0)
    dim = (C.c_uint32 * 3)()
    if brute:
        ptr = lib().orc_nlist_brute(C.byref(box), _p(pos, C.c_double), _p(types, C.c_uint32),
                                    C.c_uint32(n_local), C.c_uint32(n), _p(r_cut, C.c_double),
                                    C.c_uint32(nt), C.c_double(r_buff), C.c_int(int(full)),
                                    _p(n_neigh, C.c_uint32), _p(head, C.c_uint64), C.byref(total))
    else:
        gw = (C.c_double * 3)(*(ghost_width if ghost_width is not None else (0, 0, 0)))
        ptr = lib().orc_nlist_binned(C.byref(box), _p(pos, C.c_double), _p(types, C.c_uint32),
                                     C.c_uint32(n_local), C.c_uint32(n), _p(r_cut, C.c_double),
                                     C.c_uint32(nt), C.c_double(r_buff), C.c_int(int(full)), gw,
                                     _p(n_neigh, C.c_uint32), _p(head, C.c_uint64), C.byref(total),
                                     dim)
    out = np.ctypeslib.as_array(ptr, shape=(max(total.value, 1),))[:total.value].copy()
    lib().orc_free(ptr)
    return n_neigh, head, out, tuple(dim)


def pair_set(n_neigh, head, nl, tags=None):
    """Canonical sorted unique (min_tag, max_tag) uint32 pair array."""
    i = np.repeat(np.arange(n_neigh.size, dtype=np.uint32), n_neigh)
    j = nl.astype(np.uint32)
    if tags is not None:
        tags = _u32(tags)
        i, j = tags[i], tags[j]
    lo, hi = np.minimum(i, j), np.maximum(i, j)
    key = np.unique(lo.astype(np.uint64) << np.uint64(32) | hi.astype(np.uint64))
    return np.stack([(key >> np.uint64(32)).astype(np.uint32),
                     (key & np.uint64(0xFFFFFFFF)).astype(np.uint32)], axis=1)


def pair_lj(box, pos, types, n_neigh, head, nl, params, r_cut, r_on=None, half=False, shift_mode=0,
            compute_virial=True):
    """params: flat (ntypes^2, 2) of (epsilon, sigma). Returns force (N,4), virial (6,N)."""
    pos = _f64(pos)
    types = _u32(types)
    params = _f64(params)
    r_cut = _f64(r_cut)
    r_on = _f64(r_on if r_on is not None else np.zeros_like(r_cut))
    nt = int(round(np.sqrt(r_cut.size)))
    n = n_neigh.size
    force = np.zeros((n, 4))
    virial = np.zeros((6, n))
    lib().orc_pair_lj(C.byref(box), _p(pos, C.c_double), _p(types, C.c_uint32), C.c_uint32(n),
                      _p(_u32(n_neigh), C.c_uint32), _p(np.ascontiguousarray(head, dtype=np.uint64), C.c_uint64),
                      _p(_u32(nl), C.c_uint32), C.c_int(int(half)), _p(params, C.c_double),
                      _p(r_cut, C.c_double), _p(r_on, C.c_double), C.c_uint32(nt),
                      C.c_int(shift_mode), C.c_int(int(compute_virial)), _p(force, C.c_double),
                      _p(virial, C.c_double))
    return force, virial


def pair_dpd(box, pos, vel, types, tags, n_neigh, head, nl, params, r_cut, seed, timestep, dt, kT,
             half=False):
    """params: flat (ntypes^2, 2) of (A, gamma)."""
    pos = _f64(pos)
    vel = _f64(vel)
    types = _u32(types)
    tags = _u32(tags)
    params = _f64(params)
    r_cut = _f64(r_cut)
    nt = int(round(np.sqrt(r_cut.size)))
    n = n_neigh.size
    force = np.zeros((n, 4))
    virial = np.zeros((6, n))
    lib().orc_pair_dpd(C.byref(box), _p(pos, C.c_double), _p(vel, C.c_double), _p(types, C.c_uint32),
                       _p(tags, C.c_uint32), C.c_uint32(n), _p(_u32(n_neigh), C.c_uint32),
                       _p(np.ascontiguousarray(head, dtype=np.uint64), C.c_uint64), _p(_u32(nl), C.c_uint32),
                       C.c_int(int(half)), _p(params, C.c_double), _p(r_cut, C.c_double),
                       C.c_uint32(nt), C.c_uint16(seed), C.c_uint64(timestep), C.c_double(dt),
                       C.c_double(kT), _p(force, C.c_double), _p(virial, C.c_double))
    return force, virial


def pair_dpdlj(box, pos, vel, types, tags, n_neigh, head, nl, params, r_cut, seed, timestep, dt, kT,
               energy_shift=False):
    """params: flat (ntypes^2, 3) of (epsilon, sigma, gamma); full list."""
    pos, vel, types, tags = _f64(pos), _f64(vel), _u32(types), _u32(tags)
    params, r_cut = _f64(params), _f64(r_cut)
    nt = int(round(np.sqrt(r_cut.size)))
    n = n_neigh.size
    force = np.zeros((n, 4))
    virial = np.zeros((6, n))
    lib().orc_pair_dpdlj(C.byref(box), _p(pos, C.c_double), _p(vel, C.c_double), _p(types, C.c_uint32),
                         _p(tags, C.c_uint32), C.c_uint32(n), _p(_u32(n_neigh), C.c_uint32),
                         _p(np.ascontiguousarray(head, dtype=np.uint64), C.c_uint64), _p(_u32(nl), C.c_uint32),
                         _p(params, C.c_double), _p(r_cut, C.c_double), C.c_uint32(nt),
                         C.c_int(int(energy_shift)), C.c_uint16(seed), C.c_uint64(timestep), C.c_double(dt),
                         C.c_double(kT), _p(force, C.c_double), _p(virial, C.c_double))
    return force, virial


def eval_lj(rsq, rcutsq, epsilon, sigma, energy_shift=False):
    f = C.c_double(0)
    e = C.c_double(0)
    ok = lib().orc_eval_lj(C.c_double(rsq), C.c_double(rcutsq), C.c_double(epsilon),
                           C.c_double(sigma), C.c_int(int(energy_shift)), C.byref(f), C.byref(e))
    return bool(ok), f.value, e.value


def eval_dpd_conservative(rsq, rcutsq, A):
    f = C.c_double(0)
    e = C.c_double(0)
    ok = lib().orc_eval_dpd_conservative(C.c_double(rsq), C.c_double(rcutsq), C.c_double(A),
                                         C.byref(f), C.byref(e))
    return bool(ok), f.value, e.value


def lj_tail_correction(params, r_cut, n_per_type, volume):
    params = _f64(params)
    r_cut = _f64(r_cut)
    npt = _u32(n_per_type)
    w = C.c_double(0)
    e = lib().orc_lj_tail_correction(_p(params, C.c_double), _p(r_cut, C.c_double),
                                     C.c_uint32(npt.size), _p(npt, C.c_uint32), C.c_double(volume),
                                     C.byref(w))
    return e, w.value


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return tuple(o)


def seed_key(rng_id, timestep, seed):
    k = (C.c_uint32 * 2)()
    lib().orc_seed_key(C.c_uint8(rng_id), C.c_uint64(timestep), C.c_uint16(seed), k)
    return tuple(k)


def dpd_alpha(seed, tag_i, tag_j, timestep):
    return lib().orc_dpd_alpha(seed, tag_i, tag_j, timestep)


def bussi_factor(K, dof, dt, kT, tau, timestep, seed, first_tag=0):
    return lib().orc_bussi_factor(K, dof, dt, kT, tau, timestep, seed, first_tag)


def kinetic_energy(vel, mass):
    vel = _f64(vel)
    mass = _f64(mass)
    return lib().orc_kinetic_energy(_p(vel, C.c_double), _p(mass, C.c_double), C.c_uint32(mass.size))


def step_one(box, pos, vel, accel, image, dt, rescale=1.0):
    """In-place on fp64 pos/vel (N,3) and int32 image (N,3)."""
    lib().orc_step_one(C.byref(box), _p(pos, C.c_double), _p(vel, C.c_double), _p(accel, C.c_double),
                       _p(image, C.c_int), C.c_uint32(pos.shape[0]), C.c_double(dt),
                       C.c_double(rescale))


def step_two(vel, accel, mass, force4, dt, rescale=1.0):
    lib().orc_step_two(_p(vel, C.c_double), _p(accel, C.c_double), _p(mass, C.c_double),
                       _p(force4, C.c_double), C.c_uint32(mass.size), C.c_double(dt),
                       C.c_double(rescale))


def filter_bond_exclusions(n_neigh, head, nl, bond_group, tags=None):
    """hoomd/md/NeighborList.cc:569-640,1448-1536: returns (n_neigh, head, nlist) with every bonded
    pair removed. The reference compacts each row in place under the old head list; the rows are packed
    here afterwards so that the result has the layout nlist() returns. bond_group: (n_bonds, 2) tags."""
    n = n_neigh.size
    tags = np.arange(n, dtype=np.uint32) if tags is None else _u32(tags)
    rtag = np.zeros(int(tags.max()) + 1 if n else 1, dtype=np.uint32)
    rtag[tags] = np.arange(n, dtype=np.uint32)
    nn = np.array(n_neigh, dtype=np.uint32)
    out = np.array(nl, dtype=np.uint32)
    bg = _u32(np.asarray(bond_group).reshape(-1, 2))
    head = np.ascontiguousarray(head, dtype=np.uint64)
    lib().orc_filter_bond_exclusions(C.c_uint32(n), _p(nn, C.c_uint32), _p(head, C.c_uint64),
                                     _p(out, C.c_uint32), _p(tags, C.c_uint32), _p(rtag, C.c_uint32),
                                     _p(bg, C.c_uint32), C.c_uint32(bg.shape[0]))
    head2 = np.zeros(n, dtype=np.uint64)
    if n:
        head2[1:] = np.cumsum(nn[:-1], dtype=np.uint64)
    rows = np.repeat(np.arange(n), nn)
    within = np.arange(int(nn.sum())) - np.repeat(head2, nn).astype(np.int64)
    packed = out[(head[rows].astype(np.int64) + within)] if len(rows) else out[:0]
    return nn, head2, packed


def angle(kind, box, pos, angle_group, angle_type, params, tags=None):
    """kind 'harmonic' | 'cosinesq'; params (k, t0) per angle type; returns (force N x 4, virial 6 x N)."""
    pos = _f64(pos)
    n = pos.shape[0]
    tags = np.arange(n, dtype=np.uint32) if tags is None else _u32(tags)
    rtag = np.zeros(int(tags.max()) + 1 if n else 1, dtype=np.uint32)
    rtag[tags] = np.arange(n, dtype=np.uint32)
    ag = _u32(np.asarray(angle_group).reshape(-1, 3))
    at = _u32(angle_type)
    params = _f64(np.asarray(params).reshape(-1, 2))
    force = np.zeros((n, 4))
    virial = np.zeros((6, n))
    rc = lib().orc_angle(C.c_int({"harmonic": 0, "cosinesq": 1}[kind]), C.byref(box), _p(pos, C.c_double),
                         C.c_uint32(n), _p(rtag, C.c_uint32), _p(ag, C.c_uint32), _p(at, C.c_uint32),
                         C.c_uint32(ag.shape[0]), _p(params, C.c_double), _p(force, C.c_double),
                         _p(virial, C.c_double))
    if rc:
        raise RuntimeError(f"Error: angle {rc - 1} is incomplete.")
    return force, virial


def dihedral_opls(box, pos, group, types, params, tags=None):
    """hoomd.md.dihedral.OPLS; params (k1, k2, k3, k4) per type."""
    return dihedral_periodic(box, pos, group, types, params, tags, _kind=1)


def dihedral_periodic(box, pos, group, types, params, tags=None, _kind=0):
    """hoomd.md.dihedral.Periodic; params (k, d, n, phi0) per type; returns (force N x 4, virial 6 x N)."""
    pos = _f64(pos)
    n = pos.shape[0]
    tags = np.arange(n, dtype=np.uint32) if tags is None else _u32(tags)
    rtag = np.zeros(int(tags.max()) + 1 if n else 1, dtype=np.uint32)
    rtag[tags] = np.arange(n, dtype=np.uint32)
    dg = _u32(np.asarray(group).reshape(-1, 4))
    dt = _u32(types)
    params = _f64(np.asarray(params).reshape(-1, 4))
    force = np.zeros((n, 4))
    virial = np.zeros((6, n))
    rc = lib().orc_dihedral(C.c_int(_kind), C.byref(box), _p(pos, C.c_double), C.c_uint32(n),
                            _p(rtag, C.c_uint32), _p(dg, C.c_uint32), _p(dt, C.c_uint32),
                            C.c_uint32(dg.shape[0]), _p(params, C.c_double), _p(force, C.c_double),
                            _p(virial, C.c_double))
    if rc:
        raise RuntimeError(f"Error: dihedral {rc - 1} is incomplete.")
    return force, virial


def bond_fenewca(box, pos, bond_group, bond_type, params, tags=None):
    """PotentialBond<EvaluatorBondFENE>; params (k, r0, epsilon, sigma, delta) per bond type."""
    return bond_harmonic(box, pos, bond_group, bond_type, params, tags, _fn="orc_bond_fenewca", _np=5)


def bond_harmonic(box, pos, bond_group, bond_type, params, tags=None, _fn="orc_bond_harmonic", _np=2):
    """PotentialBond<EvaluatorBondHarmonic>: returns (force N x 4 incl. energy, virial 6 x N)."""
    pos = _f64(pos)
    n = pos.shape[0]
    tags = np.arange(n, dtype=np.uint32) if tags is None else _u32(tags)
    rtag = np.zeros(int(tags.max()) + 1 if n else 1, dtype=np.uint32)
    rtag[tags] = np.arange(n, dtype=np.uint32)
    bg = _u32(np.asarray(bond_group).reshape(-1, 2))
    bt = _u32(bond_type)
    params = _f64(np.asarray(params).reshape(-1, _np))
    force = np.zeros((n, 4))
    virial = np.zeros((6, n))
    rc = getattr(lib(), _fn)(C.byref(box), _p(pos, C.c_double), C.c_uint32(n), _p(rtag, C.c_uint32),
                                 _p(bg, C.c_uint32), _p(bt, C.c_uint32), C.c_uint32(bg.shape[0]),
                                 _p(params, C.c_double), _p(force, C.c_double), _p(virial, C.c_double))
    if rc > 0:
        raise RuntimeError(f"Error: bond {rc - 1} is incomplete.")
    if rc < 0:
        raise RuntimeError(f"bond.fene: bond out of bounds (bond {-rc - 1})")
    return force, virial


def step_two_langevin(vel, accel, mass, force4, types, tags, dt, kT, gamma, seed, timestep):
    """TwoStepLangevin::integrateStepTwo in place on fp64 vel/accel (N,3); gamma per type."""
    types = _u32(types)
    tags = _u32(tags)
    gamma = _f64(gamma)
    lib().orc_step_two_langevin(_p(vel, C.c_double), _p(accel, C.c_double), _p(_f64(mass), C.c_double),
                                _p(force4, C.c_double), _p(types, C.c_uint32), _p(tags, C.c_uint32),
                                C.c_uint32(types.size), C.c_double(dt), C.c_double(kT),
                                _p(gamma, C.c_double), C.c_uint16(seed), C.c_uint64(timestep))


def distance_check(box, pos, last_pos, types, rcut_max, r_buff):
    pos = _f64(pos)
    last_pos = _f64(last_pos)
    types = _u32(types)
    rcut_max = _f64(rcut_max)
    return bool(lib().orc_distance_check(C.byref(box), _p(pos, C.c_double), _p(last_pos, C.c_double),
                                         _p(types, C.c_uint32), C.c_uint32(types.size),
                                         _p(rcut_max, C.c_double), C.c_double(r_buff)))


def run(box, pos, vel, accel, image, mass, types, tags, params, r_cut, nsteps, dt, r_buff,
        potential="lj", shift_mode=0, thermostat=None, kT=0.0, tau=0.0, seed=0, timestep=0, gamma=1.0):
    """Whole step loop (half list) in place; returns dict(builds, n_neigh_mean, force, timestep)."""
    n = pos.shape[0]
    r_cut = _f64(r_cut)
    params = _f64(params)
    nt = int(round(np.sqrt(r_cut.size)))
    cfg = RunCfg(n, nt, dt, r_buff, kT, tau, {"bussi": 1, "langevin": 2}.get(thermostat, 0),
                 1 if potential == "dpd" else 0, shift_mode, seed, timestep, gamma)
    force = np.zeros((n, 4))
    nn = C.c_double(0)
    builds = lib().orc_run(C.byref(box), C.byref(cfg), _p(pos, C.c_double), _p(vel, C.c_double),
                           _p(accel, C.c_double), _p(image, C.c_int), _p(_f64(mass), C.c_double),
                           _p(_u32(types), C.c_uint32), _p(_u32(tags), C.c_uint32),
                           _p(params, C.c_double), _p(r_cut, C.c_double), C.c_uint64(nsteps),
                           _p(force, C.c_double), C.byref(nn))
    return dict(builds=builds, n_neigh_mean=nn.value, force=force, timestep=cfg.timestep)
