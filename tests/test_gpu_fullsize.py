"""Parity at BASELINE.json's full sizes.

configs[1] LJ N=1,048,576, configs[2] Kob-Andersen N=4,000,000 and configs[4] DPD N=8,000,000 are each
compared at run(0), AT THE CONFIGURED SIZE, with the reference itself (the HOOMD-blue CPU build under
baseline/_ref, run in a subprocess by tests/ref_run0.py; the oracle port when that build did not travel
to the box, and the assertion message says which): the neighbour-pair set bit for bit (count, per-tag
degree and an order-independent 64-bit digest of the (min, max) tag pairs), per-particle forces to 1e-5,
total energy and the virial sums to 1e-6. On top of that the size-independent properties: an independent
fp64 recomputation from the device's own pair list (numpy), a k-d tree check of sampled rows, momentum
conservation (including the DPD random force whose per-pair Philox draw is symmetric in the tags),
bitwise run-to-run reproducibility, and bounded NVE drift.
"""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pair_digest import degree, pair_digest  # noqa: E402

FORCE_RTOL = 1e-5   # north_star: per-particle forces within 1e-5 relative (norm-wise, floored at the rms)
TOTAL_RTOL = 1e-6   # north_star: total energy and virial within 1e-6 relative


def reference_run0(tmp_path, L, pos, vel, typeid, types, kind, table, mode="none", buffer=0.4,
                   dt=0.005, kT=1.0, seed=1, timestep=0):
    """run(0) of the reference on this frame (tests/ref_run0.py). table: {(a, b): (p0, p1, r_cut)}."""
    N = len(pos)
    keys = list(table)
    inp = dict(L=np.asarray([L, L, L], dtype=np.float64), pos=pos, vel=vel,
               typeid=np.zeros(N, np.uint32) if typeid is None else typeid, types=np.array(list(types)),
               kind=kind, pairs_a=np.array([k[0] for k in keys]), pairs_b=np.array([k[1] for k in keys]),
               p0=np.array([table[k][0] for k in keys], dtype=np.float64),
               p1=np.array([table[k][1] for k in keys], dtype=np.float64),
               r_cut=np.array([table[k][2] for k in keys], dtype=np.float64), mode=mode, buffer=buffer,
               dt=dt, kT=kT, seed=seed, timestep=timestep)
    fin, fout = str(tmp_path / "ref_in.npz"), str(tmp_path / "ref_out.npz")
    np.savez(fin, **inp)
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_run0.py")
    r = subprocess.run([sys.executable, script, fin, fout], capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, r.stderr[-2000:]
    out = dict(np.load(fout))
    os.remove(fin)
    os.remove(fout)
    return out


def assert_matches_reference(cell, pot, ref, N, virial=True):
    """Pair set bit-exact, forces 1e-5, energy and virial 1e-6 against `ref` (reference_run0)."""
    src = str(ref["source"])
    pl = cell.pair_list
    n, h, x = pair_digest(pl)
    assert n == int(ref["n_pairs"]), (src, n, int(ref["n_pairs"]))
    assert np.array_equal(degree(pl, N), ref["degree"]), src
    assert (h, x) == (int(ref["pair_hash"]), int(ref["pair_xor"])), src
    assert tuple(cell.dimensions) == tuple(int(d) for d in ref["dims"]), src
    del pl
    err = force_err(pot.forces, ref["force"])
    assert err <= FORCE_RTOL, (src, err)
    e_ref = float(ref["energy"])
    assert abs(pot.energy - e_ref) <= TOTAL_RTOL * abs(e_ref), (src, pot.energy, e_ref)
    if virial:
        vs, vs_ref = pot.virials.sum(axis=0), ref["virial_sum"]
        scale = np.max(np.abs(vs_ref[[0, 3, 5]]))
        assert np.max(np.abs(vs - vs_ref)) <= TOTAL_RTOL * scale, (src, vs, vs_ref)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hoomd():
    import hoomd_blue_b200 as h
    return h


def fcc(ncell, rho, kT, seed, jitter=0.0):
    a = (4.0 / rho) ** (1.0 / 3.0)
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) + 0.25
    g = np.arange(ncell)
    cells = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 1, 3)
    L = np.float32(ncell * a)
    rng = np.random.default_rng(seed)
    pos = ((cells + base[None]) * a).reshape(-1, 3) - float(L) / 2
    if jitter:
        pos = pos + rng.uniform(-jitter, jitter, size=pos.shape)
    pos = pos.astype(np.float32)
    vel = rng.normal(0, np.sqrt(kT), size=pos.shape).astype(np.float32)
    vel -= vel.mean(axis=0, dtype=np.float64).astype(np.float32)
    return pos, vel, L


def min_image(d, L):
    return d - L * np.rint(d / L)


def recompute_lj(pos, pl, L, eps, sig, rc, shift, chunk=8_000_000):
    """fp64 energy and per-particle forces from a half pair list (eps, sig, rc per pair arrays or
    scalars)."""
    N = len(pos)
    f = np.zeros((N, 3))
    e_tot = 0.0
    p64 = pos.astype(np.float64)
    for s in range(0, len(pl), chunk):
        i, j = pl[s:s + chunk, 0].astype(np.int64), pl[s:s + chunk, 1].astype(np.int64)
        d = min_image(p64[i] - p64[j], L)
        r2 = np.einsum("ij,ij->i", d, d)
        ep = eps[s:s + chunk] if np.ndim(eps) else eps
        sg = sig[s:s + chunk] if np.ndim(sig) else sig
        rcc = rc[s:s + chunk] if np.ndim(rc) else rc
        inside = r2 < rcc * rcc
        s6 = (sg * sg / r2) ** 3
        e = 4 * ep * (s6 * s6 - s6)
        if shift:
            c6 = (sg / rcc) ** 6
            e = e - 4 * ep * (c6 * c6 - c6)
        fdr = np.where(inside, 4 * ep * (12 * s6 * s6 - 6 * s6) / r2, 0.0)
        e_tot += float(np.where(inside, e, 0.0).sum())
        for c in range(3):
            w = fdr * d[:, c]
            f[:, c] += np.bincount(i, weights=w, minlength=N) - np.bincount(j, weights=w, minlength=N)
    return e_tot, f


def force_err(f, f_ref):
    n_ref = np.linalg.norm(f_ref, axis=1)
    floor = np.sqrt(np.mean(n_ref ** 2))
    return np.max(np.linalg.norm(f - f_ref, axis=1) / np.maximum(n_ref, floor))


def lj_system(hoomd, pos, vel, L, typeid=None, types=("A",), params=None, r_cut=None, mode="none",
              buffer=0.4, dt=0.005, thermostat=None, seed=3):
    sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=seed)
    sim.create_state_from_snapshot(hoomd.Snapshot.from_arrays([L, L, L, 0, 0, 0], pos, vel, typeid, types))
    cell = hoomd.md.nlist.Cell(buffer=buffer)
    lj = hoomd.md.pair.LJ(nlist=cell, mode=mode)
    for k, (eps, sig) in params.items():
        lj.params[k] = dict(epsilon=eps, sigma=sig)
        lj.r_cut[k] = r_cut[k]
    cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All(), thermostat=thermostat)
    sim.operations.integrator = hoomd.md.Integrator(dt=dt, methods=[cv], forces=[lj])
    return sim, cell, lj


def test_c2_lj_1m(hoomd, tmp_path):
    rho, kT = 0.9, 1.5
    pos, vel, L = fcc(64, rho, kT, seed=1)
    N = len(pos)
    assert N == 1_048_576
    th = hoomd.md.methods.thermostats.Bussi(kT=kT)
    sim, cell, lj = lj_system(hoomd, pos, vel, L, params={("A", "A"): (1.0, 1.0)},
                              r_cut={("A", "A"): 2.5}, thermostat=th)
    sim.run(300)  # melt
    snap = sim.state.get_snapshot()
    p, v = snap.particles.position.copy(), snap.particles.velocity.copy()
    Lf = float(L)

    def fresh():
        s, c, l = lj_system(hoomd, p, v, L, params={("A", "A"): (1.0, 1.0)}, r_cut={("A", "A"): 2.5})
        s.always_compute_pressure = True  # (both copies: the virial variant is another instantiation)
        return s, c, l

    s1, c1, l1 = fresh()
    s1.run(0)
    # the configured size against the reference itself: pair set bit-exact, forces, energy, virial
    ref = reference_run0(tmp_path, Lf, p, v, None, ("A",), "lj", {("A", "A"): (1.0, 1.0, 2.5)})
    assert_matches_reference(c1, l1, ref, N)
    pl = c1.pair_list
    # every unordered pair once, no self pairs, degrees add up
    assert (pl[:, 0] != pl[:, 1]).all()
    key = pl.min(axis=1).astype(np.uint64) << np.uint64(32) | pl.max(axis=1).astype(np.uint64)
    assert len(np.unique(key)) == len(key)
    deg = np.bincount(pl[:, 0], minlength=N) + np.bincount(pl[:, 1], minlength=N)
    assert abs(deg.mean() - 4.0 / 3.0 * np.pi * 2.9 ** 3 * rho) < 2.0 and deg.max() < 160
    # sampled rows against an independent spatial search (scipy k-d tree, periodic, fp64)
    from scipy.spatial import cKDTree
    p64 = p.astype(np.float64)
    tree = cKDTree(np.mod(p64 + Lf / 2, Lf), boxsize=Lf)
    rng = np.random.default_rng(0)
    sample = rng.choice(N, 400, replace=False)
    sel = np.isin(pl[:, 0], sample) | np.isin(pl[:, 1], sample)
    sub = pl[sel]
    for t in sample[:400]:
        have = np.sort(np.concatenate([sub[sub[:, 0] == t, 1], sub[sub[:, 1] == t, 0]]))
        want = np.array(sorted(set(tree.query_ball_point(np.mod(p64[t] + Lf / 2, Lf), 2.9)) - {int(t)}))
        if not np.array_equal(have, want):
            # only a pair within rounding of r_list may differ between the two fp64 formulations
            diff = np.setxor1d(have, want)
            d = min_image(p64[diff] - p64[t], Lf)
            assert np.all(np.abs(np.linalg.norm(d, axis=1) - 2.9) < 1e-9), (t, diff)
    # independent fp64 energy and forces from the list
    e_ref, f_ref = recompute_lj(p, pl, Lf, 1.0, 1.0, 2.5, shift=False)
    f = l1.forces
    assert force_err(f, f_ref) <= 1e-5
    assert abs(l1.energy - e_ref) <= 1e-6 * abs(e_ref)
    assert np.abs(l1.energies.sum() - e_ref) <= 1e-6 * abs(e_ref)
    # momentum conservation: the sum of all forces vanishes (relative to the sum of magnitudes)
    assert np.abs(f.sum(axis=0)).max() <= 1e-7 * np.abs(f).sum()
    # bitwise reproducibility of the list and of the forces
    s2, c2, l2 = fresh()
    s2.run(0)
    assert hashlib.sha256(c2.pair_list.tobytes()).hexdigest() == hashlib.sha256(pl.tobytes()).hexdigest()
    assert np.array_equal(l2.forces, f)
    # NVE: bounded drift, all particles still in the box, list rebuilt on the reference's criterion
    ke0, pe0, _ = s1.operations.integrator._thermo()
    s1.run(500)
    ke1, pe1, _ = s1.operations.integrator._thermo()
    assert abs((ke1 + pe1) - (ke0 + pe0)) / N < 5e-4
    assert 4 <= 500 / c1.num_builds <= 10
    q = s1.state.get_snapshot().particles.position
    assert np.all(q >= -Lf / 2) and np.all(q < Lf / 2) and np.isfinite(q).all()
    s2.run(500)
    assert np.array_equal(s2.state.get_snapshot().particles.position, q)  # deterministic trajectory


KA = {("A", "A"): (1.0, 1.0), ("A", "B"): (1.5, 0.8), ("B", "B"): (0.5, 0.88)}
KA_RC = {("A", "A"): 2.5, ("A", "B"): 2.0, ("B", "B"): 2.2}


def ka_system(hoomd, ncell, seed):
    pos, vel, L = fcc(ncell, 1.2, 1.0, seed=seed, jitter=0.05)
    N = len(pos)
    rng = np.random.default_rng(seed + 1)
    typeid = (rng.permutation(N) < N // 5).astype(np.uint32)  # 20 % B
    sim, cell, lj = lj_system(hoomd, pos, vel, L, typeid, ("A", "B"), KA, KA_RC, mode="shift",
                              buffer=0.3, dt=0.0025)
    return sim, cell, lj, pos, typeid, float(L)


def test_c3_kob_andersen(hoomd, tmp_path):
    # full fp64 recomputation with per-type-pair parameters on a 500k system of the same state point
    sim, cell, lj, pos, typeid, L = ka_system(hoomd, 50, seed=4)
    sim.run(0)
    pl = cell.pair_list
    ti, tj = typeid[pl[:, 0]], typeid[pl[:, 1]]
    code = ti + tj  # 0 AA, 1 AB, 2 BB
    eps = np.array([1.0, 1.5, 0.5])[code]
    sig = np.array([1.0, 0.8, 0.88])[code]
    rc = np.array([2.5, 2.0, 2.2])[code]
    d = min_image(pos[pl[:, 0]].astype(np.float64) - pos[pl[:, 1]].astype(np.float64), L)
    r = np.linalg.norm(d, axis=1)
    assert np.all(r <= rc + 0.3)          # the two-type stencil: per-type-pair r_list
    e_ref, f_ref = recompute_lj(pos, pl, L, eps, sig, rc, shift=True)
    assert force_err(lj.forces, f_ref) <= 1e-5
    assert abs(lj.energy - e_ref) <= 1e-6 * abs(e_ref)
    # the configured size: 4 M particles
    sim, cell, lj, pos, typeid, L = ka_system(hoomd, 100, seed=5)
    N = len(pos)
    assert N == 4_000_000
    sim.always_compute_pressure = True
    sim.run(0)
    ref = reference_run0(tmp_path, L, pos, np.zeros_like(pos), typeid, ("A", "B"), "lj",
                         {k: (KA[k][0], KA[k][1], KA_RC[k]) for k in KA}, mode="shift", buffer=0.3,
                         dt=0.0025)
    assert_matches_reference(cell, lj, ref, N)
    del ref
    f = lj.forces
    assert np.abs(f.sum(axis=0)).max() <= 1e-7 * np.abs(f).sum()
    st = cell._stats()
    assert 60 < st["mean_neighbors"] < 110
    ke0, pe0, _ = sim.operations.integrator._thermo()
    sim.run(200)
    ke1, pe1, _ = sim.operations.integrator._thermo()
    assert abs((ke1 + pe1) - (ke0 + pe0)) / N < 5e-4
    assert cell.num_builds >= 1


def test_c5_dpd_8m(hoomd, tmp_path):
    n_side, rho = 200, 3.0
    N = n_side ** 3
    assert N == 8_000_000
    L = np.float32((N / rho) ** (1.0 / 3.0))
    rng = np.random.default_rng(9)
    g = (np.arange(n_side) + 0.5) * (float(L) / n_side) - float(L) / 2
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    pos = (np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
           + rng.uniform(-0.3, 0.3, size=(N, 3))).astype(np.float32)
    half = np.float32(float(L) / 2)
    pos = np.where(pos >= half, pos - L, pos)
    pos = np.where(pos < -half, pos + L, pos).astype(np.float32)
    vel = np.zeros_like(pos)

    def make(seed):
        sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=seed)
        sim.create_state_from_snapshot(hoomd.Snapshot.from_arrays([L, L, L, 0, 0, 0], pos, vel))
        sim.always_compute_pressure = True  # (every copy: the virial variant is another instantiation)
        cell = hoomd.md.nlist.Cell(buffer=0.4)
        dpd = hoomd.md.pair.DPD(nlist=cell, kT=1.0, default_r_cut=1.0)
        dpd.params[("A", "A")] = dict(A=25.0, gamma=4.5)
        cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())
        sim.operations.integrator = hoomd.md.Integrator(dt=0.01, methods=[cv], forces=[dpd])
        return sim, cell, dpd

    s1, c1, d1 = make(5)
    s1.run(0)
    # the configured size against the reference: pair set, and forces INCLUDING the random force
    # (same seed, timestep and tags => the same per-pair Philox draws)
    ref = reference_run0(tmp_path, float(L), pos, vel, None, ("A",), "dpd", {("A", "A"): (25.0, 4.5, 1.0)},
                         buffer=0.4, dt=0.01, kT=1.0, seed=5)
    assert_matches_reference(c1, d1, ref, N)
    del ref
    f1 = d1.forces
    # alpha_ij = alpha_ji (Philox counter = (max tag, min tag)): random forces cancel pairwise
    assert np.abs(f1.sum(axis=0)).max() <= 1e-7 * np.abs(f1).sum()
    s2, c2, d2 = make(5)
    s2.run(0)
    assert np.array_equal(d2.forces, f1)          # bit-reproducible random forces
    s3, c3, d3 = make(6)
    s3.run(0)
    f3 = d3.forces
    assert not np.array_equal(f3, f1)             # ... that depend on the seed
    assert abs(d3.energy - d1.energy) <= 1e-9 * abs(d1.energy)   # conservative energy does not
    # the thermostat drives the cold start towards kT = 1 and keeps momentum at zero
    s1.run(150)
    ke, pe, _ = s1.operations.integrator._thermo()
    T = 2 * ke / (3 * N - 3)
    assert 0.8 < T < 1.2, T
    v = s1.state.get_snapshot().particles.velocity.astype(np.float64)
    assert np.abs(v.sum(axis=0)).max() / N < 1e-5
    s2.run(150)
    assert np.array_equal(s2.state.get_snapshot().particles.velocity.astype(np.float64), v)
