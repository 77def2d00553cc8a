"""Pins the CPU oracle (oracle/md_oracle.c) against (a) the reference's own golden vectors for this
path and (b) outputs of the reference itself (HOOMD-blue v4.8.2 CPU build, fp64) committed under
tests/golden/ by tests/golden/make_golden.py.  CPU only.
"""
import hashlib
import json
import os

import numpy as np
import pytest


@pytest.fixture(scope="module")
def refvec(golden_dir):
    with open(os.path.join(golden_dir, "ref_vectors.json")) as f:
        return json.load(f)


def test_philox_kat(oracle, refvec):
    # hoomd/extern/random123/tests/kat_vectors:27-29
    assert len(refvec["philox4x32_10"]) == 3
    for kat in refvec["philox4x32_10"]:
        out = oracle.philox4x32_10([int(x, 16) for x in kat["ctr"]], [int(x, 16) for x in kat["key"]])
        assert [f"{x:08x}" for x in out] == kat["out"]


def test_seed_key_kat(oracle, refvec):
    # hoomd/test/random_numbers_test.cc:456-462
    k = refvec["seed_kat"]
    key = oracle.seed_key(k["id"], int(k["timestep"], 16), int(k["seed"], 16))
    assert list(key) == [int(x, 16) for x in k["key"]]


def test_true_pair_list(oracle, refvec):
    # hoomd/md/pytest/test_nlist.py:221-222,353-372 with hoomd/conftest.py:203-262 (n=7, a=1),
    # r_cut 1.1, buffer 0
    n = 7
    r = np.arange(-n / 2, n / 2)
    x, y, z = np.meshgrid(r, r, r)
    pos = (np.vstack((x.flatten(), y.flatten(), z.flatten())).T + 0.5) * 1.0
    box = oracle.make_box([7.0, 7.0, 7.0])
    truth = set(frozenset(p) for p in refvec["true_pair_list"])
    assert len(truth) == 1029
    for full in (True, False):
        for brute in (False, True):
            nn, head, nl, dim = oracle.nlist(box, pos, np.zeros(len(pos), np.uint32), [1.1], 0.0,
                                             full=full, brute=brute)
            pairs = oracle.pair_set(nn, head, nl)
            assert len(pairs) == 1029
            assert set(frozenset((int(a), int(b))) for a, b in pairs) == truth
            if full:
                assert nn.sum() == 2 * 1029
            if not brute:
                assert dim == (6, 6, 6)


def test_forces_and_energies_json(oracle, refvec):
    # hoomd/md/pytest/test_potential.py:994-1051: r = 0.75 and 1.5, r_cut 2.5, rtol 5e-6
    fe = refvec["forces_and_energies"]
    for p, F, E in zip(fe["LJ"]["params"], fe["LJ"]["forces"], fe["LJ"]["energies"]):
        for d, f_ref, e_ref in zip((0.75, 1.5), F, E):
            ok, fdivr, eng = oracle.eval_lj(d * d, 2.5 * 2.5, p["epsilon"], p["sigma"])
            assert ok
            assert np.isclose(-d * fdivr, f_ref, rtol=5e-6, atol=1e-6 if f_ref == 0 else 0)
            assert np.isclose(eng, e_ref, rtol=5e-6, atol=1e-6 if e_ref == 0 else 0)
    dc = fe["DPDConservative"]
    for p, F, E in zip(dc["params"], dc["forces"], dc["energies"]):
        for d, f_ref, e_ref in zip((0.75, 1.5), F, E):
            ok, fdivr, eng = oracle.eval_dpd_conservative(d * d, 2.5 * 2.5, p["A"])
            assert ok
            assert np.isclose(-d * fdivr, f_ref, rtol=5e-6)
            assert np.isclose(eng, e_ref, rtol=5e-6)


def test_lj_shift_and_cutoff(oracle):
    # hoomd/md/pytest/test_potential.py:780 (energy shifting): E_shift(r) = E(r) - E(r_cut)
    ok, f0, e0 = oracle.eval_lj(1.2 ** 2, 2.5 ** 2, 1.0, 1.0, energy_shift=False)
    ok, f1, e1 = oracle.eval_lj(1.2 ** 2, 2.5 ** 2, 1.0, 1.0, energy_shift=True)
    _, _, ec = oracle.eval_lj(2.5 ** 2 * (1 - 1e-15), 2.5 ** 2, 1.0, 1.0)
    assert f0 == f1 and np.isclose(e1, e0 - ec, rtol=1e-12)
    assert not oracle.eval_lj(2.5 ** 2, 2.5 ** 2, 1.0, 1.0)[0]       # rsq < rcutsq is strict
    assert not oracle.eval_lj(1.0, 2.5 ** 2, 0.0, 1.0)[0]            # lj1 == 0 gate


def _type_params(pdict, names, cols):
    nt = len(names)
    out = np.zeros((nt, nt, cols))
    for (a, b), v in pdict.items():
        i, j = names.index(a), names.index(b)
        out[i, j] = out[j, i] = v
    return out.reshape(nt * nt, cols)


def _check_frame(oracle, g, prefix, params, r_cut, buffer, shift_mode, r_on=None):
    box = oracle.make_box(g["L"].astype(np.float64))
    pos = g["pos"].astype(np.float64)
    types = g["typeid"]
    for full in (False, True):
        nn, head, nl, dim = oracle.nlist(box, pos, types, r_cut, buffer, full=full)
        pairs = oracle.pair_set(nn, head, nl)
        assert pairs.shape == g[f"{prefix}_pairs"].shape
        assert np.array_equal(pairs, g[f"{prefix}_pairs"])          # bit-exact pair set
        assert tuple(g[f"{prefix}_dims"]) == dim
        force, virial = oracle.pair_lj(box, pos, types, nn, head, nl, params, r_cut, r_on=r_on,
                                       half=not full, shift_mode=shift_mode)
        # half list: same summation order as the reference -> near bit-level agreement
        tol = 1e-13 if not full else 1e-11
        scale = np.abs(g[f"{prefix}_force"]).max()
        assert np.abs(force[:, :3] - g[f"{prefix}_force"]).max() <= tol * scale
        assert np.allclose(force[:, 3], g[f"{prefix}_energy"], rtol=1e-10, atol=1e-10)
        assert np.allclose(virial.T, g[f"{prefix}_virial"], rtol=1e-9, atol=1e-9)
        assert np.isclose(force[:, 3].sum(), float(g[f"{prefix}_total_energy"]), rtol=1e-12)


def test_lj_2k_frame_vs_reference(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "lj_2k.npz"))
    params = np.array([[1.0, 1.0]])
    _check_frame(oracle, g, "none", params, [2.5], 0.4, 0)
    _check_frame(oracle, g, "shift", params, [2.5], 0.4, 1)
    _check_frame(oracle, g, "xplor", params, [2.5], 0.4, 2, r_on=[2.0])


def test_ka_2k_frame_vs_reference(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "ka_2k.npz"))
    names = ["A", "B"]
    params = _type_params({("A", "A"): (1.0, 1.0), ("A", "B"): (1.5, 0.8), ("B", "B"): (0.5, 0.88)},
                          names, 2)
    r_cut = _type_params({("A", "A"): 2.5, ("A", "B"): 2.0, ("B", "B"): 2.2}, names, 1).ravel()
    _check_frame(oracle, g, "shift", params, r_cut, 0.3, 1)


def test_dpd_3k_vs_reference(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "dpd_3k.npz"))
    box = oracle.make_box(g["L"].astype(np.float64))
    pos = g["pos"].astype(np.float64)
    vel = g["vel"].astype(np.float64)
    N = len(pos)
    types = g["typeid"]
    tags = np.arange(N, dtype=np.uint32)
    params = np.array([[float(g["A"]), float(g["gamma"])]])
    for full in (False, True):
        nn, head, nl, _ = oracle.nlist(box, pos, types, [1.0], 0.4, full=full)
        assert np.array_equal(oracle.pair_set(nn, head, nl), g["pairs"])
        for n, ts in enumerate(g["timesteps"]):
            force, virial = oracle.pair_dpd(box, pos, vel, types, tags, nn, head, nl, params, [1.0],
                                            int(g["seed"]), int(ts), float(g["dt"]), float(g["kT"]),
                                            half=not full)
            scale = np.abs(g[f"force_{n}"]).max()
            # random + drag forces included: pins the Philox key/counter layout and u01 mapping
            assert np.abs(force[:, :3] - g[f"force_{n}"]).max() <= 1e-11 * scale
            assert np.allclose(force[:, 3], g[f"energy_{n}"], rtol=1e-10, atol=1e-12)
            assert np.allclose(virial.T, g[f"virial_{n}"], rtol=1e-9, atol=1e-10)
    # the three timesteps must give different random forces
    assert not np.allclose(g["force_0"], g["force_1"])


def _run_traj(oracle, g, thermostat, nsteps=20):
    box = oracle.make_box(g["L"].astype(np.float64))
    pos = g["pos"].astype(np.float64).copy()
    vel = g["vel"].astype(np.float64).copy()
    img = g["image"].astype(np.int32).copy()
    N = len(pos)
    mass = np.ones(N)
    types = g["typeid"]
    tags = np.arange(N, dtype=np.uint32)
    params = np.array([[1.0, 1.0]])
    # prepRun: accel = F/m at t0 (hoomd/md/IntegratorTwoStep.cc:262-275)
    nn, head, nl, _ = oracle.nlist(box, pos, types, [2.5], 0.4, full=False)
    f0, _ = oracle.pair_lj(box, pos, types, nn, head, nl, params, [2.5], half=True)
    accel = np.ascontiguousarray(f0[:, :3] / mass[:, None])
    r = oracle.run(box, pos, vel, accel, img, mass, types, tags, params, [2.5], nsteps, 0.005, 0.4,
                   thermostat=thermostat, kT=1.5, tau=0.0, seed=11)  # Bussi default tau=0 (hoomd/md/methods/thermostats.py:319)
    return pos, vel, img, r


def test_nve_trajectory_vs_reference(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "lj_2k.npz"))
    pos, vel, img, r = _run_traj(oracle, g, None)
    assert np.abs(pos - g["nve_pos"]).max() < 1e-10
    assert np.abs(vel - g["nve_vel"]).max() < 1e-10
    assert np.array_equal(img, g["nve_image"])
    assert 1 <= r["builds"] <= 5    # (hoomd resets num_builds on every run() call: not comparable)
    ke = oracle.kinetic_energy(vel, np.ones(len(vel)))
    assert np.isclose(ke, g["nve_ke"][-1], rtol=1e-11)
    assert np.isclose(r["force"][:, 3].sum(), g["nve_pe"][-1], rtol=1e-11)


def test_bussi_trajectory_vs_reference(oracle, golden_dir):
    # pins the Bussi rescale factor incl. HOOMD's Normal/Gamma draws (hoomd/md/Thermostat.h:357-447)
    g = np.load(os.path.join(golden_dir, "lj_2k.npz"))
    pos, vel, img, r = _run_traj(oracle, g, "bussi")
    assert np.abs(pos - g["bussi_pos"]).max() < 1e-9
    assert np.abs(vel - g["bussi_vel"]).max() < 1e-9
    ke = oracle.kinetic_energy(vel, np.ones(len(vel)))
    assert np.isclose(ke, g["bussi_ke"][-1], rtol=1e-10)
    assert not np.allclose(g["bussi_vel"], g["nve_vel"], atol=1e-6)


def test_c1_64k_vs_reference(oracle, golden_dir):
    path = os.path.join(golden_dir, "c1_64k.npz")
    if not os.path.exists(path):
        pytest.skip("c1_64k.npz not generated")
    g = np.load(path)
    box = oracle.make_box(g["L"].astype(np.float64))
    pos = g["pos"].astype(np.float64)
    types = np.zeros(len(pos), np.uint32)
    nn, head, nl, dim = oracle.nlist(box, pos, types, [2.5], 0.4, full=False)
    pairs = oracle.pair_set(nn, head, nl)
    assert len(pairs) == int(g["n_pairs"])
    assert hashlib.sha256(pairs.tobytes()).hexdigest() == str(g["pairs_sha256"])
    assert tuple(g["dims"]) == dim
    force, virial = oracle.pair_lj(box, pos, types, nn, head, nl, np.array([[1.0, 1.0]]), [2.5],
                                   half=True)
    assert np.abs(force[:, :3] - g["force"]).max() <= 1e-12 * np.abs(g["force"]).max()
    assert np.isclose(force[:, 3].sum(), float(g["total_energy"]), rtol=1e-12)
    assert np.allclose(virial.sum(axis=1), g["virial_sum"], rtol=1e-10)


def test_cell_list_restatement(oracle):
    # hoomd/test/test_cell_list.cc:33-148 (dimensions), :155-238 (adjacency), :245-505 (small system)
    box = oracle.make_box([10.0, 10.0, 10.0])
    assert oracle.cell_dims(box, 1.0) == (10, 10, 10)
    assert oracle.cell_dims(box, 3.0) == (3, 3, 3)
    assert oracle.cell_dims(box, 30.0) == (1, 1, 1)
    adj = oracle.cell_adj((5, 5, 5))
    assert adj.shape == (125, 27)
    assert np.all(np.diff(adj.astype(np.int64), axis=1) > 0)          # sorted, unique
    # cell (0,0,0): neighbours wrap to index 4 in every direction
    want = sorted((k % 5 * 5 + j % 5) * 5 + i % 5 for k in (-1, 0, 1) for j in (-1, 0, 1)
                  for i in (-1, 0, 1))
    assert list(adj[0]) == want
    assert oracle.cell_adj((2, 2, 2)).shape == (8, 8)                # de-duplicated rows
    pos = np.array([[-4.9, -4.9, -4.9], [4.9, 4.9, 4.9], [0.0, 0.0, 0.0], [0.1, 0.1, 0.1]])
    size, idx, cond = oracle.cell_list(box, (3, 3, 3), pos, 4)
    assert cond == (0, 0, 0)
    assert size[0] == 1 and size[26] == 1 and size[13] == 2
    assert list(idx[13][:2]) == [2, 3]                               # ascending particle index
    _, _, cond = oracle.cell_list(box, (3, 3, 3), np.array([[np.nan, 0, 0]]), 4)
    assert cond[1] == 1                                              # NaN position flagged
    _, _, cond = oracle.cell_list(box, (3, 3, 3), np.array([[5.1, 0, 0]]), 4)
    assert cond[2] == 1                                              # out of box flagged
    _, _, cond = oracle.cell_list(box, (3, 3, 3), np.zeros((6, 3)), 4)
    assert cond[0] == 6                                              # overflow reports needed Nmax


def test_min_image_and_wrap(oracle):
    # hoomd/md/test/test_neighborlist.cc:33-326 exercises pairs across each periodic face
    box = oracle.make_box([20.0, 40.0, 60.0])
    d = oracle.min_image(box, np.array([[19.0, -39.0, 59.0], [10.0, 20.0, 30.0],
                                        [-10.0, -20.0, -30.0], [9.9, 0, 0]]))
    assert np.allclose(d[0], [-1.0, 1.0, -1.0])
    assert np.allclose(d[1], [-10.0, -20.0, -30.0])                  # >= hi wraps
    assert np.allclose(d[2], [-10.0, -20.0, -30.0])                  # == lo stays
    assert np.allclose(d[3], [9.9, 0, 0])
    pos = np.array([[10.0, -20.5, 30.0]])
    vel = np.zeros((1, 3))
    acc = np.zeros((1, 3))
    img = np.zeros((1, 3), dtype=np.int32)
    oracle.step_one(box, pos, vel, acc, img, 0.005)
    assert np.allclose(pos, [[-10.0, 19.5, -30.0]]) and img.tolist() == [[1, -1, 1]]


def test_distance_check(oracle):
    box = oracle.make_box([10.0, 10.0, 10.0])
    last = np.zeros((2, 3))
    pos = np.array([[0.0, 0.0, 0.0], [0.19, 0.0, 0.0]])
    t = np.zeros(2, np.uint32)
    assert not oracle.distance_check(box, pos, last, t, [2.5], 0.4)
    pos[1, 0] = 0.2000001
    assert oracle.distance_check(box, pos, last, t, [2.5], 0.4)     # >= (buffer/2)^2
    pos[1] = [4.95, 0, 0]
    last[1] = [-4.95, 0, 0]
    assert not oracle.distance_check(box, pos, last, t, [2.5], 0.4)  # min image of displacement


def test_langevin_step_two(oracle):
    """TwoStepLangevin::integrateStepTwo (hoomd/md/TwoStepLangevin.cc:213-292). The reference's own
    tests of the method check attributes only (hoomd/md/pytest/test_methods.py:66-75, :345-358), so the
    restatement is pinned through what it shares with the pinned paths — Philox (KAT), Seed/Counter and
    the uniform draw (DPD fixture) — plus the identities of the update rule and the
    fluctuation-dissipation statistic. PARITY OF THE LANGEVIN NOISE STREAM IS NOT PINNED by a
    reference-generated trajectory."""
    rng = np.random.default_rng(3)
    n = 4096
    vel = rng.normal(0, 1.0, size=(n, 3))
    mass = rng.uniform(0.5, 2.0, size=n)
    force = np.zeros((n, 4))
    force[:, :3] = rng.normal(0, 3.0, size=(n, 3))
    types = (np.arange(n) % 2).astype(np.uint32)
    tags = rng.permutation(n).astype(np.uint32)
    # gamma = 0: no drag and no noise -> the NVE half step
    v0, a0 = vel.copy(), np.zeros((n, 3))
    oracle.step_two_langevin(v0, a0, mass, force, types, tags, 0.005, 1.5, [0.0, 0.0], 7, 10)
    v1, a1 = vel.copy(), np.zeros((n, 3))
    oracle.step_two(v1, a1, mass, force, 0.005)
    assert np.array_equal(v0, v1) and np.array_equal(a0, a1)
    # the noise is keyed by (timestep, seed, tag) only: independent of memory order, new every step,
    # uniform in [-1, 1] scaled by sqrt(6 gamma kT / dt), gamma per type
    gamma = np.array([0.5, 2.0])
    v2, a2 = vel.copy(), np.zeros((n, 3))
    oracle.step_two_langevin(v2, a2, mass, force, types, tags, 0.005, 1.5, gamma, 7, 10)
    perm = rng.permutation(n)
    v3, a3 = vel[perm].copy(), np.zeros((n, 3))
    oracle.step_two_langevin(v3, a3, mass[perm], force[perm], types[perm], tags[perm], 0.005, 1.5, gamma,
                             7, 10)
    assert np.array_equal(v3, v2[perm])
    g = gamma[types][:, None]
    r = (a2 * mass[:, None] - force[:, :3] + g * vel) / np.sqrt(6.0 * g * 1.5 / 0.005)
    assert np.all(np.abs(r) <= 1.0 + 1e-12)
    assert abs(r.mean()) < 0.02 and abs(r.var() - 1.0 / 3.0) < 0.01
    key = np.array([0, 0], dtype=np.uint32)
    key_buf = np.zeros(2, dtype=np.uint32)
    import ctypes as C
    oracle.lib().orc_seed_key(C.c_uint8(25), C.c_uint64(10), C.c_uint16(7),
                              key_buf.ctypes.data_as(C.POINTER(C.c_uint32)))
    ctr = np.array([0, 0, 0, tags[5]], dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    oracle.lib().orc_philox4x32_10(ctr.ctypes.data_as(C.POINTER(C.c_uint32)),
                                   key_buf.ctypes.data_as(C.POINTER(C.c_uint32)),
                                   out.ctypes.data_as(C.POINTER(C.c_uint32)))
    u = (int(out[0]) << 32 | int(out[1])) * 2.0 ** -64 + 2.0 ** -65
    assert np.isclose(r[5, 0], -1.0 + 2.0 * u, rtol=0, atol=1e-12)  # first draw = x component
    v4, a4 = vel.copy(), np.zeros((n, 3))
    oracle.step_two_langevin(v4, a4, mass, force, types, tags, 0.005, 1.5, gamma, 7, 11)
    assert not np.allclose(a4, a2)
    # fluctuation-dissipation: free particles relax to <m v^2> = kT per component
    vel = np.zeros((n, 3))
    acc = np.zeros((n, 3))
    zero = np.zeros((n, 4))
    dt, kT = 0.01, 1.5
    kes = []
    for t in range(1500):
        vel += 0.5 * dt * acc                      # step one's velocity half (positions play no role)
        oracle.step_two_langevin(vel, acc, mass, zero, types, tags, dt, kT, [2.0, 4.0], 7, t)
        if t >= 500:
            kes.append(oracle.kinetic_energy(vel, mass))
    assert abs(np.mean(kes) / (1.5 * n * kT) - 1.0) < 0.02


def test_host_rng_and_thermalize(oracle):
    """hoomd_blue_b200/rng.py (numpy Philox, Seed/Counter, Box-Muller) against the oracle's Philox and a
    scalar restatement of hoomd/RandomNumbers.h:63-87,361-374, and thermal_velocities
    (hoomd/ParticleGroup.cc:690-806): tag-keyed, zero total momentum, <m v^2> = kT per component.
    THE THERMALIZE STREAM IS NOT PINNED by reference output (RNG id 35 from hoomd/RNGIdentifiers.h:60)."""
    import ctypes as C
    from hoomd_blue_b200 import rng
    g = np.random.default_rng(1)
    ctr = g.integers(0, 2 ** 32, size=(64, 4), dtype=np.uint64).astype(np.uint32)
    key = g.integers(0, 2 ** 32, size=2, dtype=np.uint64).astype(np.uint32)
    got = rng.philox4x32_10(ctr, key)
    out = np.zeros(4, dtype=np.uint32)
    for i in range(len(ctr)):
        c = np.ascontiguousarray(ctr[i])
        oracle.lib().orc_philox4x32_10(c.ctypes.data_as(C.POINTER(C.c_uint32)),
                                       key.ctypes.data_as(C.POINTER(C.c_uint32)),
                                       out.ctypes.data_as(C.POINTER(C.c_uint32)))
        assert np.array_equal(got[i], out)
    kb = np.zeros(2, dtype=np.uint32)
    oracle.lib().orc_seed_key(C.c_uint8(35), C.c_uint64((1 << 35) + 17), C.c_uint16(7),
                              kb.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert np.array_equal(rng.seed_key(35, (1 << 35) + 17, 7), kb)
    tags = np.array([0, 5, 123456, 2 ** 31 + 3], dtype=np.uint32)
    nd = rng.normal_draws(35, 99, 7, tags, 3)
    for a, tag in enumerate(tags):
        for d in range(3):
            c = np.array([d, 0, 0, tag], dtype=np.uint32)
            k2 = rng.seed_key(35, 99, 7)
            oracle.lib().orc_philox4x32_10(c.ctypes.data_as(C.POINTER(C.c_uint32)),
                                           k2.ctypes.data_as(C.POINTER(C.c_uint32)),
                                           out.ctypes.data_as(C.POINTER(C.c_uint32)))
            u0 = (int(out[0]) << 32) | int(out[1])
            u1 = (int(out[2]) << 32) | int(out[3])
            s0 = u0 - (1 << 64) if u0 >= (1 << 63) else u0
            x = np.sin(np.pi * (s0 * 2.0 ** -63 + 2.0 ** -64)) * np.sqrt(-2.0 * np.log(u1 * 2.0 ** -64 + 2.0 ** -65))
            assert np.isclose(nd[a, d], x, rtol=1e-13, atol=1e-15)
    n = 20000
    mass = g.uniform(0.5, 3.0, n)
    tg = g.permutation(n).astype(np.uint32)
    v = rng.thermal_velocities(tg, mass, 1.5, 7, 0)
    assert np.abs((mass[:, None] * v).sum(axis=0)).max() < 1e-9
    assert abs((mass[:, None] * v ** 2).mean() / 1.5 - 1.0) < 0.02
    perm = g.permutation(n)
    assert np.allclose(rng.thermal_velocities(tg[perm], mass[perm], 1.5, 7, 0), v[perm], atol=1e-12)
    assert not np.allclose(rng.thermal_velocities(tg, mass, 1.5, 7, 1), v)


def test_langevin_noise_stream_pinned_by_the_reference(oracle, golden_dir):
    """tests/golden/langevin_2k.npz was produced by the reference itself (tests/golden/make_langevin.py,
    HOOMD-blue CPU build under baseline/_ref): Langevin(kT = 1.5, gamma = 2) on the 2k LJ frame, seed 11.
    The oracle's restatement (orc_step_two_langevin inside orc run) must reproduce the reference's
    velocities after ONE step to fp64 round-off — that pins the Philox key / counter layout, the three
    draws per particle and their order — and track it over 20 steps."""
    import os
    fx = dict(np.load(os.path.join(golden_dir, "lj_2k.npz")))
    ref = dict(np.load(os.path.join(golden_dir, "langevin_2k.npz")))
    n = len(fx["pos"])
    box = oracle.make_box([float(v) for v in fx["L"]])
    L = fx["L"].astype(np.float64)

    def run(steps):
        pos, vel = fx["pos"].astype(np.float64), fx["vel"].astype(np.float64)
        img = np.zeros((n, 3), dtype=np.int32)
        mass, types, tags = np.ones(n), np.zeros(n, dtype=np.uint32), np.arange(n, dtype=np.uint32)
        nn, head, nl, _ = oracle.nlist(box, pos, types, [2.5], 0.4, full=True)
        f0, _ = oracle.pair_lj(box, pos, types, nn, head, nl, [[1.0, 1.0]], [2.5], shift_mode=0)
        accel = np.ascontiguousarray(f0[:, :3] / mass[:, None])
        oracle.run(box, pos, vel, accel, img, mass, types, tags, [[1.0, 1.0]], [2.5], steps, 0.005, 0.4,
                   thermostat="langevin", kT=float(ref["kT"]), seed=int(ref["seed"]), gamma=float(ref["gamma"]))
        return pos, vel

    pos, vel = run(1)
    d = pos - ref["pos_1"]
    d -= L * np.round(d / L)
    assert np.max(np.abs(d)) < 1e-12
    assert np.max(np.abs(vel - ref["vel_1"])) < 1e-11
    pos, vel = run(20)
    d = pos - ref["pos_20"]
    d -= L * np.round(d / L)
    assert np.max(np.abs(d)) < 1e-9 and np.max(np.abs(vel - ref["vel_20"])) < 1e-8


def test_thermalize_pinned_by_the_reference(golden_dir):
    """State.thermalize_particle_momenta of the reference (seed 7, timestep 0, kT 1.5, unit masses)
    against the host implementation in hoomd_blue_b200/rng.py (hoomd/ParticleData / hoomd/State.cc
    thermalizeParticleMomenta: Philox keyed by tag, Box-Muller normals, centre-of-mass momentum removed)."""
    import os
    from hoomd_blue_b200 import rng
    ref = dict(np.load(os.path.join(golden_dir, "langevin_2k.npz")))
    n = len(ref["thermal_vel"])
    got = rng.thermal_velocities(np.arange(n, dtype=np.uint32), np.ones(n, dtype=np.float32), 1.5,
                                 int(ref["thermal_seed"]), 0)
    assert np.max(np.abs(got - ref["thermal_vel"])) < 1e-12


def test_npt_barostat_replay(golden_dir):
    """The barostat restatement (oracle/npt_oracle.py) replays the reference's NPH trajectory: from the
    recorded pressure, kinetic energy and box of every step it reproduces nu and the box lengths of the next
    (hoomd/md/TwoStepConstantPressure.cc advanceBarostat / updatePropagator / integrateStepOne)."""
    from oracle import npt_oracle as npt
    fx = np.load(os.path.join(golden_dir, "npt_2k.npz"))
    N = len(fx["a_pos"])
    ndof, dt, tauS, S3 = 3 * N - 3, 0.005, 0.5, [2.0] * 3
    box, nu6, P, ke = fx["a_box"], fx["a_nu"], fx["a_P"], fx["a_ke"]
    for k in range(len(P) - 1):
        nu = nu6[k][[0, 3, 5]]
        V = float(np.prod(box[k]))
        half = npt.advance_barostat(nu, [P[k]] * 3, ke[k], V, S3, ndof, tauS, dt)
        _, exp_r, _ = npt.propagator(half, dt)
        Lnew = exp_r * box[k]
        np.testing.assert_allclose(Lnew, box[k + 1], rtol=1e-13)
        full = npt.advance_barostat(half, [P[k + 1]] * 3, ke[k + 1], float(np.prod(Lnew)), S3, ndof, tauS, dt)
        np.testing.assert_allclose(full, nu6[k + 1][[0, 3, 5]], rtol=1e-10, atol=1e-14)
    assert abs(npt.barostat_energy(nu6[-1], ndof, tauS) - float(fx["a_benergy"])) <= 1e-10 * abs(float(fx["a_benergy"]))
    # the couplings that drop out when a box degree of freedom is frozen
    assert npt.relevant_couple("xyz", (True, True, False)) == "xy"
    assert npt.relevant_couple("xy", (True, False, True)) == "none"
    assert npt.relevant_couple("yz", (False, True, True)) == "yz"
