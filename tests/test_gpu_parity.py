"""GPU parity tests: the CUDA path (through the C ABI / operator surface) against the committed
reference fixtures (tests/golden, made by the reference itself) and against the CPU oracle on seeded
inputs. Bars (BASELINE.json north_star): neighbour-pair set bit-exact; per-particle force within 1e-5
relative (norm-wise, floored at the rms force); total energy and virial within 1e-6 relative.
"""
import hashlib
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FORCE_RTOL = 1e-5
TOTAL_RTOL = 1e-6


@pytest.fixture(scope="module")
def hoomd():
    import hoomd_blue_b200 as h
    return h


def canon_pairs(pl):
    pl = np.asarray(pl, dtype=np.uint32).reshape(-1, 2)
    lo, hi = pl.min(axis=1), pl.max(axis=1)
    key = np.unique(lo.astype(np.uint64) << np.uint64(32) | hi.astype(np.uint64))
    return np.stack([(key >> np.uint64(32)).astype(np.uint32),
                     (key & np.uint64(0xFFFFFFFF)).astype(np.uint32)], axis=1)


def force_err(f, f_ref):
    """max over particles of |f - f_ref| / max(|f_ref|, rms|f_ref|)"""
    n_ref = np.linalg.norm(f_ref, axis=1)
    floor = np.sqrt(np.mean(n_ref ** 2))
    return np.max(np.linalg.norm(f - f_ref, axis=1) / np.maximum(n_ref, floor))


def make_sim(hoomd, L, pos, vel=None, typeid=None, types=("A",), image=None, seed=1, timestep=None):
    sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=seed)
    if timestep:
        sim.timestep = timestep
    snap = hoomd.Snapshot.from_arrays(list(L) + [0, 0, 0], pos, vel, typeid, types, image)
    sim.create_state_from_snapshot(snap)
    return sim


def lj_sim(hoomd, fx, params, r_cut, buffer, mode, r_on=None, types=("A",), seed=1, thermostat=None,
           dt=0.005, tpp=None):
    sim = make_sim(hoomd, fx["L"], fx["pos"], fx["vel"], fx["typeid"] if "typeid" in fx else None,
                   types, fx["image"] if "image" in fx else None, seed)
    sim.always_compute_pressure = True
    cell = hoomd.md.nlist.Cell(buffer=buffer)
    lj = hoomd.md.pair.LJ(nlist=cell, mode=mode)
    for (a, b), (eps, sig) in params.items():
        lj.params[(a, b)] = dict(epsilon=eps, sigma=sig)
        lj.r_cut[(a, b)] = r_cut[(a, b)]
        if r_on is not None:
            lj.r_on[(a, b)] = r_on[(a, b)]
    if tpp:
        lj.set_threads_per_particle(tpp)
    cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All(), thermostat=thermostat)
    sim.operations.integrator = hoomd.md.Integrator(dt=dt, methods=[cv], forces=[lj])
    return sim, cell, lj


def check_forces(lj, fx, prefix):
    f = lj.forces
    assert force_err(f, fx[f"{prefix}force"]) <= FORCE_RTOL
    e_tot = float(fx[f"{prefix}total_energy"])
    assert abs(lj.energy - e_tot) <= TOTAL_RTOL * abs(e_tot)
    # per-particle energies: absolute error relative to the rms energy
    e = lj.energies
    e_ref = fx[f"{prefix}energy"]
    assert np.max(np.abs(e - e_ref)) <= 1e-5 * np.sqrt(np.mean(e_ref ** 2))
    v = lj.virials
    v_ref = fx[f"{prefix}virial"]
    vs, vs_ref = v.sum(axis=0), v_ref.sum(axis=0)
    # the diagonal sums carry the pressure; off-diagonals average to ~0, so compare against the
    # diagonal scale
    scale = np.max(np.abs(vs_ref[[0, 3, 5]]))
    assert np.max(np.abs(vs[[0, 3, 5]] - vs_ref[[0, 3, 5]])) <= TOTAL_RTOL * scale
    assert np.max(np.abs(vs - vs_ref)) <= 1e-5 * scale
    assert np.max(np.abs(v - v_ref)) <= 1e-5 * np.sqrt(np.mean(v_ref ** 2)) * 10


# ---------------------------------------------------------------- reference golden vectors ----
def test_true_pair_list(hoomd, golden_dir):
    """hoomd/md/pytest/test_nlist.py:221-222,353-372 with hoomd/conftest.py:203-262."""
    with open(os.path.join(golden_dir, "ref_vectors.json")) as f:
        truth = set(frozenset(p) for p in json.load(f)["true_pair_list"])
    n = 7
    r = np.arange(-n / 2, n / 2)
    x, y, z = np.meshgrid(r, r, r)
    pos = (np.vstack((x.flatten(), y.flatten(), z.flatten())).T + 0.5) * 1.0
    fx = dict(L=[7.0, 7.0, 7.0], pos=pos.astype(np.float32), vel=np.zeros_like(pos, dtype=np.float32))
    sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 1.1}, 0.0, "none")
    sim.run(0)
    pairs = cell.pair_list
    assert len(pairs) == 1029
    assert set(frozenset((int(a), int(b))) for a, b in pairs) == truth
    assert cell.dimensions == (6, 6, 6)
    assert cell.num_builds == 1


def test_forces_and_energies_json(hoomd, golden_dir):
    """hoomd/md/pytest/test_potential.py:952-1051: two particles at r = 0.75 and 1.5, rtol 5e-6."""
    with open(os.path.join(golden_dir, "ref_vectors.json")) as f:
        fe = json.load(f)["forces_and_energies"]
    for name, cls_name, keys in (("LJ", "LJ", ("epsilon", "sigma")),
                                 ("DPDConservative", "DPDConservative", ("A",)),
                                 ("Gaussian", "Gaussian", ("epsilon", "sigma")),
                                 ("Yukawa", "Yukawa", ("epsilon", "kappa")),
                                 ("LJ1208", "LJ1208", ("epsilon", "sigma")),
                                 ("LJ0804", "LJ0804", ("epsilon", "sigma")),
                                 ("ForceShiftedLJ", "ForceShiftedLJ", ("epsilon", "sigma")),
                                 ("Mie", "Mie", ("epsilon", "sigma", "n", "m")),
                                 ("Morse", "Morse", ("D0", "alpha", "r0")),
                                 ("Buckingham", "Buckingham", ("A", "rho", "C")),
                                 ("ExpandedLJ", "ExpandedLJ", ("epsilon", "sigma", "delta")),
                                 ("ExpandedGaussian", "ExpandedGaussian", ("epsilon", "sigma", "delta")),
                                 ("LJGauss", "LJGauss", ("epsilon", "sigma", "r0")),
                                 ("TWF", "TWF", ("epsilon", "sigma", "alpha")),
                                 ("ExpandedMie", "ExpandedMie", ("epsilon", "sigma", "n", "m", "delta")),
                                 ("OPP", "OPP", ("C1", "C2", "eta1", "eta2", "k", "phi")),
                                 ("Fourier", "Fourier", ("a", "b")),
                                 ("Table", "Table", ("r_min", "U", "F"))):
        tab = fe[name]
        for p, F, E in zip(tab["params"], tab["forces"], tab["energies"]):
            for d, f_ref, e_ref in zip((0.75, 1.5), F, E):
                pos = np.array([[0, 0, 0.1], [0, 0, 0.1 + d]], dtype=np.float32)
                sim = make_sim(hoomd, [20.0, 20.0, 20.0], pos, np.zeros_like(pos))
                cell = hoomd.md.nlist.Cell(buffer=0.4)
                cls = getattr(hoomd.md.pair, cls_name)
                pot = cls(nlist=cell, default_r_cut=2.5)
                pot.params[("A", "A")] = {k: p[k] for k in keys}
                cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())
                sim.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[cv], forces=[pot])
                sim.run(0)
                d32 = float(pos[1, 2]) - float(pos[0, 2])  # the fp32 separation actually simulated
                f = pot.forces
                # the table holds the force on particle 0 along r_hat = +z
                # (test_potential.py:1044-1048: sim_forces[0] == forces[i] * r)
                assert abs(d32 - d) < 1e-6
                if isinstance(f_ref, str) or isinstance(e_ref, str):
                    # 'infinity' / 'neg_infinity' entries (hoomd/md/pytest/test_potential.py:1017-1027):
                    # overlapping TWF particles, ExpandedLJ at r = delta
                    want = {"infinity": np.inf, "neg_infinity": -np.inf}
                    assert f[0, 2] == want[f_ref] and f[1, 2] == -want[f_ref] and pot.energy == want[e_ref]
                    continue
                # steep tables (ExpandedLJ reaches 6e8 at delta = 1): fp32 evaluates r - delta to ~1e-7
                rtol = 2e-5 if abs(f_ref) < 1e4 else 2e-4
                assert np.isclose(f[0, 2], f_ref, rtol=rtol, atol=1e-6 if f_ref == 0 else 0), (name, p, d)
                assert np.isclose(f[1, 2], -f_ref, rtol=rtol, atol=1e-6 if f_ref == 0 else 0)
                assert np.isclose(pot.energy, e_ref, rtol=rtol, atol=1e-6 if e_ref == 0 else 0), (name, p, d)


@pytest.mark.parametrize("mode", ["none", "shift", "xplor"])
def test_lj_2k(hoomd, golden_dir, mode):
    fx = dict(np.load(os.path.join(golden_dir, "lj_2k.npz")))
    r_on = {("A", "A"): 2.0} if mode == "xplor" else None
    sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, mode, r_on)
    sim.run(0)
    assert np.array_equal(canon_pairs(cell.pair_list), fx[f"{mode}_pairs"])
    assert cell.dimensions == tuple(int(d) for d in fx[f"{mode}_dims"])
    check_forces(lj, fx, f"{mode}_")


@pytest.mark.parametrize("tpp", [4, 8, 16, 32])
def test_lj_2k_threads_per_particle(hoomd, golden_dir, tpp):
    fx = dict(np.load(os.path.join(golden_dir, "lj_2k.npz")))
    sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "none",
                           tpp=tpp)
    sim.run(0)
    check_forces(lj, fx, "none_")


KA_PARAMS = {("A", "A"): (1.0, 1.0), ("A", "B"): (1.5, 0.8), ("B", "B"): (0.5, 0.88)}
KA_RC = {("A", "A"): 2.5, ("A", "B"): 2.0, ("B", "B"): 2.2}


def test_ka_2k(hoomd, golden_dir):
    """Kob-Andersen 80:20: per-type-pair parameters and r_cut, mode shift."""
    fx = dict(np.load(os.path.join(golden_dir, "ka_2k.npz")))
    sim, cell, lj = lj_sim(hoomd, fx, KA_PARAMS, KA_RC, 0.3, "shift", types=("A", "B"))
    sim.run(0)
    assert np.array_equal(canon_pairs(cell.pair_list), fx["shift_pairs"])
    check_forces(lj, fx, "shift_")


def test_dpd_3k(hoomd, golden_dir):
    """DPD with the per-pair Philox draw at three timesteps (one beyond 2^32)."""
    fx = dict(np.load(os.path.join(golden_dir, "dpd_3k.npz")))
    for n, ts in enumerate(fx["timesteps"]):
        sim = make_sim(hoomd, fx["L"], fx["pos"], fx["vel"], seed=int(fx["seed"]), timestep=int(ts))
        sim.always_compute_pressure = True
        cell = hoomd.md.nlist.Cell(buffer=float(fx["buffer"]))
        dpd = hoomd.md.pair.DPD(nlist=cell, kT=float(fx["kT"]), default_r_cut=float(fx["r_cut"]))
        dpd.params[("A", "A")] = dict(A=float(fx["A"]), gamma=float(fx["gamma"]))
        cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())
        sim.operations.integrator = hoomd.md.Integrator(dt=float(fx["dt"]), methods=[cv],
                                                        forces=[dpd])
        sim.run(0)
        if n == 0:
            assert np.array_equal(canon_pairs(cell.pair_list), fx["pairs"])
        assert force_err(dpd.forces, fx[f"force_{n}"]) <= FORCE_RTOL
        e_ref = fx[f"energy_{n}"]
        assert abs(dpd.energy - e_ref.sum()) <= TOTAL_RTOL * abs(e_ref.sum())
        vs, vs_ref = dpd.virials.sum(axis=0), fx[f"virial_{n}"].sum(axis=0)
        scale = np.max(np.abs(vs_ref[[0, 3, 5]]))
        assert np.max(np.abs(vs[[0, 3, 5]] - vs_ref[[0, 3, 5]])) <= TOTAL_RTOL * scale


def test_c1_64k(hoomd, golden_dir):
    """BASELINE.json configs[0]: LJ liquid N=64,000 — pair set by sha256, forces, totals."""
    fx = dict(np.load(os.path.join(golden_dir, "c1_64k.npz")))
    fx["typeid"] = np.zeros(len(fx["pos"]), dtype=np.uint32)
    sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "none")
    sim.run(0)
    pairs = canon_pairs(cell.pair_list)
    assert len(pairs) == int(fx["n_pairs"])
    assert np.array_equal(pairs[:64], fx["pairs_head"])
    assert hashlib.sha256(pairs.tobytes()).hexdigest() == str(fx["pairs_sha256"])
    assert cell.dimensions == tuple(int(d) for d in fx["dims"])
    assert force_err(lj.forces, fx["force"]) <= FORCE_RTOL
    e_tot = float(fx["total_energy"])
    assert abs(lj.energy - e_tot) <= TOTAL_RTOL * abs(e_tot)
    vs, vs_ref = lj.virials.sum(axis=0), fx["virial_sum"]
    scale = np.max(np.abs(vs_ref[[0, 3, 5]]))
    assert np.max(np.abs(vs[[0, 3, 5]] - vs_ref[[0, 3, 5]])) <= TOTAL_RTOL * scale


# ------------------------------------------------------------------------- trajectories ----
def test_nve_trajectory_lj_2k(hoomd, golden_dir):
    """20 NVE steps against the reference's fp64 trajectory: fp32 state tracks it to ~1e-4, energy
    is conserved as well as the reference conserves it."""
    fx = dict(np.load(os.path.join(golden_dir, "lj_2k.npz")))
    sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "none",
                           seed=11)
    thermo = hoomd.md.compute.ThermodynamicQuantities(filter=hoomd.filter.All())
    thermo._attach(sim)
    sim.run(0)
    ke, pe = [thermo.kinetic_energy], [thermo.potential_energy]
    for _ in range(20):
        sim.run(1)
        ke.append(thermo.kinetic_energy)
        pe.append(thermo.potential_energy)
    ke, pe = np.array(ke), np.array(pe)
    assert sim.timestep == 20
    assert np.allclose(ke, fx["nve_ke"], rtol=2e-5)
    assert np.allclose(pe, fx["nve_pe"], rtol=2e-5)
    e, e_ref = ke + pe, fx["nve_ke"] + fx["nve_pe"]
    assert np.max(np.abs(e - e[0])) <= 2.0 * np.max(np.abs(e_ref - e_ref[0])) + 1e-6 * abs(e_ref[0])
    s = sim.state.get_snapshot()
    L = fx["L"].astype(np.float64)
    d = s.particles.position.astype(np.float64) - fx["nve_pos"]
    d -= L * np.round(d / L)
    assert np.max(np.abs(d)) < 2e-4
    assert np.max(np.abs(s.particles.velocity - fx["nve_vel"])) < 2e-3
    unwrapped = s.particles.position + s.particles.image * L
    unwrapped_ref = fx["nve_pos"] + fx["nve_image"] * L
    assert np.max(np.abs(unwrapped - unwrapped_ref)) < 2e-4
    assert cell.num_builds == int(fx["nve_builds"])


def test_bussi_trajectory_lj_2k(hoomd, golden_dir):
    """20 steps with the Bussi thermostat: the rescale factor is a function of (KE, timestep, seed)
    only, so the kinetic-energy trace must follow the reference's."""
    fx = dict(np.load(os.path.join(golden_dir, "lj_2k.npz")))
    th = hoomd.md.methods.thermostats.Bussi(kT=1.5)
    sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "none",
                           seed=11, thermostat=th)
    thermo = hoomd.md.compute.ThermodynamicQuantities(filter=hoomd.filter.All())
    thermo._attach(sim)
    sim.run(0)
    ke = [thermo.kinetic_energy]
    for _ in range(20):
        sim.run(1)
        ke.append(thermo.kinetic_energy)
    assert np.allclose(np.array(ke), fx["bussi_ke"], rtol=5e-5)
    s = sim.state.get_snapshot()
    L = fx["L"].astype(np.float64)
    d = s.particles.position.astype(np.float64) - fx["bussi_pos"]
    d -= L * np.round(d / L)
    assert np.max(np.abs(d)) < 2e-4


def test_dpd_trajectory(hoomd, golden_dir):
    fx = dict(np.load(os.path.join(golden_dir, "dpd_3k.npz")))
    sim = make_sim(hoomd, fx["L"], fx["pos"], fx["vel"], seed=int(fx["seed"]))
    cell = hoomd.md.nlist.Cell(buffer=float(fx["buffer"]))
    dpd = hoomd.md.pair.DPD(nlist=cell, kT=float(fx["kT"]), default_r_cut=float(fx["r_cut"]))
    dpd.params[("A", "A")] = dict(A=float(fx["A"]), gamma=float(fx["gamma"]))
    cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())
    sim.operations.integrator = hoomd.md.Integrator(dt=float(fx["dt"]), methods=[cv], forces=[dpd])
    sim.run(10)
    s = sim.state.get_snapshot()
    L = fx["L"].astype(np.float64)
    d = s.particles.position.astype(np.float64) - fx["traj_pos"]
    d -= L * np.round(d / L)
    assert np.max(np.abs(d)) < 1e-4
    assert np.max(np.abs(s.particles.velocity - fx["traj_vel"])) < 2e-3


def test_langevin_trajectory_vs_oracle(hoomd, oracle, golden_dir):
    """hoomd.md.methods.Langevin (hoomd/md/TwoStepLangevin.cc) against the oracle's fp64 restatement on
    the 2k LJ frame: the noise is a function of (timestep, seed, tag), so the fp32 trajectory tracks the
    fp64 one like the NVE trajectory does."""
    fx = dict(np.load(os.path.join(golden_dir, "lj_2k.npz")))
    n = len(fx["pos"])
    sim = make_sim(hoomd, fx["L"], fx["pos"], fx["vel"], seed=11)
    cell = hoomd.md.nlist.Cell(buffer=0.4)
    lj = hoomd.md.pair.LJ(nlist=cell, default_r_cut=2.5)
    lj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
    lv = hoomd.md.methods.Langevin(filter=hoomd.filter.All(), kT=1.5, default_gamma=2.0)
    sim.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[lv], forces=[lj])
    sim.run(20)
    s = sim.state.get_snapshot()

    box = oracle.make_box([float(v) for v in fx["L"]])
    pos, vel = fx["pos"].astype(np.float64), fx["vel"].astype(np.float64)
    img = np.zeros((n, 3), dtype=np.int32)
    mass, types, tags = np.ones(n), np.zeros(n, dtype=np.uint32), np.arange(n, dtype=np.uint32)
    nn, head, nl, _ = oracle.nlist(box, pos, types, [2.5], 0.4, full=True)
    f0, _ = oracle.pair_lj(box, pos, types, nn, head, nl, [[1.0, 1.0]], [2.5], shift_mode=0)
    accel = np.ascontiguousarray(f0[:, :3] / mass[:, None])
    oracle.run(box, pos, vel, accel, img, mass, types, tags, [[1.0, 1.0]], [2.5], 20, 0.005, 0.4,
               thermostat="langevin", kT=1.5, seed=11, gamma=2.0)
    L = fx["L"].astype(np.float64)
    d = s.particles.position.astype(np.float64) - pos
    d -= L * np.round(d / L)
    assert np.max(np.abs(d)) < 2e-4
    assert np.max(np.abs(s.particles.velocity - vel)) < 2e-3
    # and it is not the NVE trajectory
    assert np.max(np.abs(s.particles.velocity - fx["nve_vel"])) > 0.05
    # the REFERENCE's own Langevin trajectory from the same frame and seed (tests/golden/langevin_2k.npz,
    # made by tests/golden/make_langevin.py): same noise stream, fp32 vs fp64 arithmetic
    ref = dict(np.load(os.path.join(golden_dir, "langevin_2k.npz")))
    d = s.particles.position.astype(np.float64) - ref["pos_20"]
    d -= L * np.round(d / L)
    assert np.max(np.abs(d)) < 2e-4
    assert np.max(np.abs(s.particles.velocity - ref["vel_20"])) < 2e-3
    sim1 = make_sim(hoomd, fx["L"], fx["pos"], fx["vel"], seed=11)
    cell1 = hoomd.md.nlist.Cell(buffer=0.4)
    lj1 = hoomd.md.pair.LJ(nlist=cell1, default_r_cut=2.5)
    lj1.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
    lv1 = hoomd.md.methods.Langevin(filter=hoomd.filter.All(), kT=1.5, default_gamma=2.0)
    sim1.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[lv1], forces=[lj1])
    sim1.run(1)
    s1 = sim1.state.get_snapshot()
    # one step: the random kick is ~0.1 per component; agreement to 1e-5 means the same draws
    assert np.max(np.abs(s1.particles.velocity - ref["vel_1"])) < 1e-5


def test_thermalize_particle_momenta(hoomd):
    """hoomd.State.thermalize_particle_momenta: velocities at kT with zero total momentum, a function of
    (seed, timestep, tag) only."""
    from hoomd_blue_b200 import rng
    pos, vel, L = random_liquid(16, 0.8, seed=3)
    mass = np.random.default_rng(4).uniform(0.5, 2.0, len(pos)).astype(np.float32)
    snap = hoomd.Snapshot.from_arrays(list(L) + [0, 0, 0], pos, vel, mass=mass)
    sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=7)
    sim.create_state_from_snapshot(snap)
    sim.state.thermalize_particle_momenta(filter=hoomd.filter.All(), kT=1.5)
    s = sim.state.get_snapshot()
    v = s.particles.velocity.astype(np.float64)
    assert np.array_equal(s.particles.position, pos) and np.array_equal(s.particles.mass, mass)
    assert np.abs((mass[:, None] * v).sum(axis=0)).max() < 1e-3
    assert abs((mass[:, None] * v ** 2).mean() / 1.5 - 1.0) < 0.05
    want = rng.thermal_velocities(np.arange(len(pos), dtype=np.uint32), mass, 1.5, 7, 0).astype(np.float32)
    assert np.array_equal(s.particles.velocity, want)


def test_langevin_temperature(hoomd):
    """Fluctuation-dissipation on the device: a dilute, nearly ideal gas started at rest reaches the
    set temperature (per-type drag coefficients)."""
    pos, vel, L = random_liquid(24, 0.05, seed=5)
    typeid = (np.arange(len(pos)) % 2).astype(np.uint32)
    sim = make_sim(hoomd, L, pos, np.zeros_like(vel), typeid=typeid, types=("A", "B"), seed=3)
    cell = hoomd.md.nlist.Cell(buffer=0.4)
    lj = hoomd.md.pair.LJ(nlist=cell, default_r_cut=2.5)
    lj.params[(["A", "B"], ["A", "B"])] = dict(epsilon=0.0, sigma=1.0)
    lv = hoomd.md.methods.Langevin(filter=hoomd.filter.All(), kT=2.0)
    lv.gamma["A"] = 2.0
    lv.gamma["B"] = 5.0
    sim.operations.integrator = hoomd.md.Integrator(dt=0.01, methods=[lv], forces=[lj])
    thermo = hoomd.md.compute.ThermodynamicQuantities(filter=hoomd.filter.All())
    thermo._attach(sim)
    sim.run(600)
    temps = []
    for _ in range(40):
        sim.run(25)
        temps.append(thermo.kinetic_temperature)
    assert abs(np.mean(temps) / 2.0 - 1.0) < 0.02


def test_dpdlj(hoomd, oracle, golden_dir):
    """hoomd.md.pair.DPDLJ: the conservative energies of the reference's table
    (forces_and_energies.json; the reference skips its force check because of the random force,
    hoomd/md/pytest/test_potential.py:1019-1021), and conservative + drag + random force against the
    oracle on a liquid with moving particles, at a timestep beyond 2^32."""
    with open(os.path.join(golden_dir, "ref_vectors.json")) as f:
        tab = json.load(f)["forces_and_energies"]["DPDLJ"]
    for p, E in zip(tab["params"], tab["energies"]):
        for d, e_ref in zip((0.75, 1.5), E):
            pos = np.array([[0, 0, 0.1], [0, 0, 0.1 + d]], dtype=np.float32)
            sim = make_sim(hoomd, [20.0, 20.0, 20.0], pos, np.zeros_like(pos))
            pot = hoomd.md.pair.DPDLJ(nlist=hoomd.md.nlist.Cell(buffer=0.4), kT=1.0, default_r_cut=2.5)
            pot.params[("A", "A")] = p
            cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())
            sim.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[cv], forces=[pot])
            sim.run(0)
            assert np.isclose(pot.energy, e_ref, rtol=2e-5, atol=1e-6 if e_ref == 0 else 0), (p, d)
    pos, vel, L = random_liquid(16, 0.8, seed=21, jitter=0.15)
    n = len(pos)
    step = (1 << 33) + 5
    sim = make_sim(hoomd, L, pos, vel, seed=9, timestep=step)
    sim.always_compute_pressure = True
    cell = hoomd.md.nlist.Cell(buffer=0.3)
    pot = hoomd.md.pair.DPDLJ(nlist=cell, kT=1.3, default_r_cut=2.0, mode="shift")
    pot.params[("A", "A")] = dict(epsilon=1.0, sigma=0.9, gamma=4.5)
    cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())
    sim.operations.integrator = hoomd.md.Integrator(dt=0.004, methods=[cv], forces=[pot])
    sim.run(0)
    box = oracle.make_box([float(v) for v in L])
    types, tags = np.zeros(n, dtype=np.uint32), np.arange(n, dtype=np.uint32)
    nn, head, nl, _ = oracle.nlist(box, pos, types, [2.0], 0.3, full=True)
    f_ref, v_ref = oracle.pair_dpdlj(box, pos, vel, types, tags, nn, head, nl, [[1.0, 0.9, 4.5]], [2.0],
                                     9, step, 0.004, 1.3, energy_shift=True)
    # the random force is O(sqrt(6 kT gamma / dt)) ~ 90 per pair: fp32 alpha carries 2^-24 of it
    assert force_err(pot.forces, f_ref[:, :3]) <= 3e-5
    e_ref = f_ref[:, 3].sum()
    assert abs(pot.energy - e_ref) <= TOTAL_RTOL * abs(e_ref)
    assert np.abs(pot.forces.sum(axis=0)).max() <= 1e-5 * np.abs(pot.forces).sum()  # third law incl. noise
    f_cons, _ = oracle.pair_lj(box, pos, types, nn, head, nl, [[1.0, 0.9]], [2.0], shift_mode=1)
    assert force_err(pot.forces, f_cons[:, :3]) > 1e-2  # the thermostat forces are really there
    # virial is the conservative part only (hoomd/md/PotentialPairDPDThermo.h:227-236)
    vs, vs_ref = pot.virials.sum(axis=0), v_ref.sum(axis=1)
    assert np.max(np.abs(vs - vs_ref)) <= 1e-5 * np.max(np.abs(vs_ref))


# ------------------------------------------------------------------ oracle on seeded inputs ----
def random_liquid(n_side, rho, seed, jitter=0.1):
    rng = np.random.default_rng(seed)
    N = n_side ** 3
    L = np.float32((N / rho) ** (1 / 3))
    g = (np.arange(n_side) + 0.5) * (float(L) / n_side) - float(L) / 2
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    pos = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    pos += rng.uniform(-jitter, jitter, size=pos.shape)
    pos = pos.astype(np.float32)
    Lf = float(L)
    pos[pos >= Lf / 2] -= L
    pos[pos < -Lf / 2] += L
    pos = np.clip(pos, -Lf / 2, np.nextafter(np.float32(Lf / 2), np.float32(0)))
    vel = rng.normal(0, 1.0, size=pos.shape).astype(np.float32)
    return pos, vel, np.array([L, L, L], dtype=np.float32)


@pytest.mark.parametrize("n_side,rho", [(30, 0.9), (20, 0.4), (48, 0.9)])
def test_oracle_random_liquid(hoomd, oracle, n_side, rho):
    pos, vel, L = random_liquid(n_side, rho, seed=n_side)
    fx = dict(L=L, pos=pos, vel=vel)
    sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "shift")
    sim.run(0)
    box = oracle.make_box([float(v) for v in L])
    types = np.zeros(len(pos), dtype=np.uint32)
    nn, head, nl, dim = oracle.nlist(box, pos, types, [2.5], 0.4, full=True)
    assert np.array_equal(canon_pairs(cell.pair_list), oracle.pair_set(nn, head, nl))
    assert cell.dimensions == dim
    f_ref, v_ref = oracle.pair_lj(box, pos, types, nn, head, nl, [[1.0, 1.0]], [2.5], shift_mode=1)
    assert force_err(lj.forces, f_ref[:, :3]) <= FORCE_RTOL
    e_ref = f_ref[:, 3].sum()
    assert abs(lj.energy - e_ref) <= TOTAL_RTOL * abs(e_ref)
    vs, vs_ref = lj.virials.sum(axis=0), v_ref.sum(axis=1)
    scale = np.max(np.abs(vs_ref[[0, 3, 5]]))
    assert np.max(np.abs(vs[[0, 3, 5]] - vs_ref[[0, 3, 5]])) <= TOTAL_RTOL * scale


def random_gas(n, L, seed):
    rng = np.random.default_rng(seed)
    L = np.asarray(L, dtype=np.float32)
    pos = (rng.random((n, 3)) - 0.5) * L.astype(np.float64)
    pos = np.clip(pos.astype(np.float32), -L / 2, np.nextafter(L / 2, np.float32(0)))
    return pos, L


@pytest.mark.parametrize("L,n,kind", [
    ((14.3, 23.1, 31.7), 7000, "one"),      # orthorhombic, cell grid 4 x 7 x 10
    ((6.5, 20.0, 9.1), 900, "one"),         # 2 cells along x: both images of a neighbour in one stencil
    ((5.9, 6.1, 40.0), 1000, "one"),        # 2 x 2 x 13
    ((17.0, 12.5, 21.3), 3500, "two"),      # per-type-pair r_cut, one pair switched off (r_cut = 0)
    ((6.2, 19.0, 8.8), 800, "two"),
])
def test_oracle_orthorhombic_boxes(hoomd, oracle, L, n, kind):
    """Non-cubic boxes, grids with fewer than three cells along an axis, and per-type-pair cut-offs
    including a disabled pair — the shapes hoomd/md/pytest/test_nlist.py and test_potential.py sweep —
    on a random ideal gas (close pairs included: ulp-level agreement of steep forces is part of it)."""
    pos, Lf = random_gas(n, L, seed=int(sum(L) * 10) + n)
    # push overlapping pairs apart a little: r^-13 forces above 1e12 say nothing about the kernel
    if kind == "one":
        types, names = np.zeros(n, dtype=np.uint32), ("A",)
        r_cut = {("A", "A"): 2.5}
        params = {("A", "A"): (1.0, 0.3)}
    else:
        types, names = (np.arange(n) % 2).astype(np.uint32), ("A", "B")
        r_cut = {("A", "A"): 2.5, ("A", "B"): 0.0, ("B", "B"): 1.8}
        params = {("A", "A"): (1.0, 0.3), ("A", "B"): (2.0, 0.25), ("B", "B"): (0.5, 0.35)}
    fx = dict(L=Lf, pos=pos, vel=np.zeros_like(pos), typeid=types)
    sim, cell, lj = lj_sim(hoomd, fx, params, r_cut, 0.4, "shift", types=names)
    sim.run(0)
    box = oracle.make_box([float(v) for v in Lf])
    nt = len(names)
    rc = np.zeros((nt, nt))
    pr = np.zeros((nt * nt, 2))
    for (a, b), v in r_cut.items():
        i, j = names.index(a), names.index(b)
        rc[i, j] = rc[j, i] = v
        pr[i * nt + j] = pr[j * nt + i] = params[(a, b)]
    nn, head, nl, dim = oracle.nlist(box, pos, types, rc.ravel(), 0.4, full=True)
    assert cell.dimensions == dim
    assert np.array_equal(canon_pairs(cell.pair_list), oracle.pair_set(nn, head, nl))
    f_ref, v_ref = oracle.pair_lj(box, pos, types, nn, head, nl, pr, rc.ravel(), shift_mode=1)
    assert force_err(lj.forces, f_ref[:, :3]) <= FORCE_RTOL
    e_ref = f_ref[:, 3].sum()
    assert abs(lj.energy - e_ref) <= TOTAL_RTOL * max(abs(e_ref), np.abs(f_ref[:, 3]).sum() * 1e-3)
    if kind == "two":
        # the disabled pair contributes nothing: A-B pairs are absent from the list altogether
        pl = cell.pair_list
        assert not np.any(types[pl[:, 0]] != types[pl[:, 1]])


def test_edge_cases(hoomd):
    """box too small, NaN position, particle out of the box, empty cells / sparse gas."""
    pos = np.array([[0, 0, 0], [1, 0, 0]], dtype=np.float32)
    fx = dict(L=[5.0, 5.0, 5.0], pos=pos, vel=np.zeros_like(pos))
    sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "none")
    with pytest.raises(RuntimeError, match="too small"):
        sim.run(0)
    bad = pos.copy()
    bad[1, 0] = np.nan
    fx = dict(L=[20.0, 20.0, 20.0], pos=bad, vel=np.zeros_like(pos))
    sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "none")
    with pytest.raises(RuntimeError, match="NaN"):
        sim.run(0)
    bad = pos.copy()
    bad[1, 0] = 11.0
    fx = dict(L=[20.0, 20.0, 20.0], pos=bad, vel=np.zeros_like(pos))
    sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "none")
    with pytest.raises(RuntimeError, match="no longer in the simulation box"):
        sim.run(0)
    # two particles across the periodic face, far fewer particles than cells
    p2 = np.array([[9.9, 0, 0], [-9.9, 0, 0], [0, 5, 5]], dtype=np.float32)
    fx = dict(L=[20.0, 20.0, 20.0], pos=p2, vel=np.zeros_like(p2))
    sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "none")
    sim.run(0)
    assert [sorted(p) for p in cell.pair_list.tolist()] == [[0, 1]]
    f = lj.forces
    assert np.allclose(f[0], -f[1]) and np.allclose(f[2], 0) and f[0, 0] != 0
    assert lj.virials is not None
    sim.always_compute_pressure = False
    assert lj.virials is None


def test_nve_drift_1e5(hoomd, golden_dir):
    """NVE energy drift over 1e5 steps (BASELINE.json: "comparable NVE energy drift over 1e5 steps"):
    the 2,048-particle reference frame, LJ shift, dt = 0.005, against the fp64 CPU oracle's trace
    (tests/golden/make_drift.py). Trajectories diverge chaotically; the drift statistic is what
    compares. The oracle's first sample is skipped (its harness starts from zero accelerations)."""
    fx = dict(np.load(os.path.join(golden_dir, "lj_2k.npz")))
    ref = np.load(os.path.join(golden_dir, "nve_drift_2k.npz"))
    N = len(fx["pos"])
    e_ref = (ref["ke"] + ref["pe"]) / N
    every = int(ref["every"])
    sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "shift")
    sim.run(0)
    integ = sim.operations.integrator
    e = []
    builds = 0
    for _ in range(len(e_ref) - 1):
        sim.run(every)
        builds += cell.num_builds
        ke, pe, _ = integ._thermo()
        e.append((ke + pe) / N)
    e = np.array(e)
    drift = e[-1] - e[0]
    drift_ref = e_ref[-1] - e_ref[1]
    fluct, fluct_ref = np.std(np.diff(e)), np.std(np.diff(e_ref[1:]))
    print(f"NVE 1e5 steps: drift/N gpu {drift:.2e} ref {drift_ref:.2e}; chunk-to-chunk sigma gpu "
          f"{fluct:.2e} ref {fluct_ref:.2e}; builds gpu {builds} ref {int(ref['builds'])}")
    # same short-time energy noise (the integrator and the truncated force are the same) ...
    assert fluct <= 2.0 * fluct_ref
    # ... and a long-time drift of the same order as the fp64 reference's (|ref| ~ 5e-4 per particle)
    assert abs(drift) <= max(5.0 * abs(drift_ref), 3e-3)
    # same rebuild cadence (identical criterion, different chaotic trajectory)
    assert abs(builds - int(ref["builds"])) <= 0.02 * int(ref["builds"])


def test_nve_drift_1e5_reference_8k(hoomd, golden_dir):
    """SURVEY.md 8(d) gate 4 as specified: NVE (thermostat=None) over 1e5 steps at N = 8,000 against the
    trace of the REFERENCE ITSELF from the same frame (tests/golden/nve_drift_8k.npz, made by
    tests/golden/make_drift_ref.py with the HOOMD-blue CPU build under baseline/_ref). Trajectories
    diverge chaotically; what compares is (a) the chunk-to-chunk energy noise — the integrator and the
    truncated, shifted force are the same — (b) the long-time drift, measured between the means of the
    first and the last 20 samples so that the noise averages out, and (c) the rebuild cadence."""
    ref = dict(np.load(os.path.join(golden_dir, "nve_drift_8k.npz")))
    N = len(ref["pos"])
    assert N == 8000
    L = float(ref["L"])
    fx = dict(L=np.array([L, L, L], dtype=np.float32), pos=ref["pos"], vel=ref["vel"])
    sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "shift")
    sim.run(0)
    integ = sim.operations.integrator
    every = int(ref["every"])
    e_ref = (ref["ke"] + ref["pe"]) / N
    ke, pe, _ = integ._thermo()
    e = [(ke + pe) / N]
    builds = 0
    for _ in range(len(e_ref) - 1):
        sim.run(every)
        builds += cell.num_builds
        ke, pe, _ = integ._thermo()
        e.append((ke + pe) / N)
    e = np.array(e)
    # same starting energy (same frame, fp32 vs fp64 force pass)
    assert abs(e[0] - e_ref[0]) <= 1e-6 * abs(e_ref[0])
    fluct, fluct_ref = np.std(np.diff(e)), np.std(np.diff(e_ref))
    drift = e[-20:].mean() - e[:20].mean()
    drift_ref = e_ref[-20:].mean() - e_ref[:20].mean()
    noise_ref = np.std(e_ref - np.polyval(np.polyfit(np.arange(len(e_ref)), e_ref, 1), np.arange(len(e_ref))))
    print(f"NVE 1e5 steps N=8000: drift/N gpu {drift:.2e} reference {drift_ref:.2e} (reference noise "
          f"{noise_ref:.2e}); chunk-to-chunk sigma gpu {fluct:.2e} reference {fluct_ref:.2e}; builds gpu "
          f"{builds} reference {int(ref['builds'])}")
    assert fluct <= 1.5 * fluct_ref
    # kT = 1.5: a drift of 1e-3 per particle is 4.4e-4 of the kinetic energy per particle over 1e5 steps;
    # the fp64 reference itself wanders by its noise (~1e-4)
    assert abs(drift) <= 1e-3
    # Rebuild cadence. The reference run carries HOOMD's default ParticleSorter (hoomd/tune/sorter.py:44,
    # every 200 steps), whose sort signal forces a list update (hoomd/md/NeighborList.cc:222): each forced
    # update restarts the displacement clock at a random phase of a natural interval, i.e. adds half a build
    # on average. This engine re-orders at every rebuild and has no separate sorter, so the comparison is
    # against the reference's count less those extra half builds.
    steps = every * (len(e_ref) - 1)
    natural_ref = int(ref["builds"]) - 0.5 * (steps // 200)
    print(f"reference builds less the sorter's forced updates: {natural_ref:.0f}")
    assert abs(builds - natural_ref) <= 0.01 * natural_ref


def test_snapshot_roundtrip_pageable_and_pinned(hoomd):
    """set_snapshot / get_snapshot are exact copies in tag order, through the bounce buffer (pageable
    arrays) and by direct DMA (page-locked arrays), before and after the particles were re-ordered."""
    pos, vel, L = random_liquid(16, 0.8, seed=2)
    N = len(pos)
    rng = np.random.default_rng(1)
    img = rng.integers(-3, 4, size=(N, 3)).astype(np.int32)
    tid = rng.integers(0, 2, size=N).astype(np.uint32)
    mass = rng.uniform(0.5, 2.0, size=N).astype(np.float32)
    sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=1)
    snap = hoomd.Snapshot.from_arrays(list(L) + [0, 0, 0], pos, vel, tid, ("A", "B"), img, mass)
    sim.create_state_from_snapshot(snap)
    for pinned in (False, True):
        out = sim.state.get_snapshot(into=hoomd.Snapshot(pinned=pinned))
        for a, b in ((out.particles.position, pos), (out.particles.velocity, vel),
                     (out.particles.image, img), (out.particles.typeid, tid), (out.particles.mass, mass)):
            assert np.array_equal(a, b)
        sim.state.set_snapshot(out)
    cell = hoomd.md.nlist.Cell(buffer=0.4)
    lj = hoomd.md.pair.LJ(nlist=cell, default_r_cut=2.5)
    for k in (("A", "A"), ("A", "B"), ("B", "B")):
        lj.params[k] = dict(epsilon=1.0, sigma=1.0)
    cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())
    sim.operations.integrator = hoomd.md.Integrator(dt=0.0, methods=[cv], forces=[lj])
    sim.run(3)  # dt = 0: the cell sort permutes the arrays, positions stay put
    out = sim.state.get_snapshot(into=hoomd.Snapshot(pinned=True))
    assert np.array_equal(out.particles.position, pos)
    assert np.array_equal(out.particles.typeid, tid)
    assert np.array_equal(out.particles.image, img)


@pytest.mark.parametrize("cls_name,params", [
    ("Gaussian", dict(epsilon=1.5, sigma=0.7)), ("Yukawa", dict(epsilon=2.0, kappa=2.5)),
    ("LJ1208", dict(epsilon=1.0, sigma=1.0)), ("LJ0804", dict(epsilon=1.0, sigma=1.0)),
    ("ForceShiftedLJ", dict(epsilon=1.0, sigma=1.0)), ("Mie", dict(epsilon=1.0, sigma=1.0, n=14, m=7)),
    ("Morse", dict(D0=1.2, alpha=3.5, r0=1.1)), ("LJ", dict(epsilon=1.0, sigma=1.0)),
    ("Buckingham", dict(A=400.0, rho=0.25, C=1.0)), ("ExpandedLJ", dict(epsilon=1.0, sigma=0.8, delta=0.2)),
    ("ExpandedGaussian", dict(epsilon=1.5, sigma=0.6, delta=0.3)),
    ("LJGauss", dict(epsilon=1.0, sigma=0.2, r0=1.5)), ("TWF", dict(epsilon=1.0, sigma=0.45, alpha=20.0)),
    ("ExpandedMie", dict(epsilon=1.0, sigma=0.8, n=13, m=7, delta=0.15)),
    ("OPP", dict(C1=1.0, C2=1.0, eta1=15.0, eta2=3.0, k=3.0, phi=0.5))])
def test_evaluator_family_consistency(hoomd, cls_name, params):
    """Every evaluator through the kernel template on a liquid: F = -dE/dr (central differences of the
    total energy along the force of a few particles), Newton's third law (sum of forces = 0), and
    mode='shift' changing energies by exactly the pair count inside r_cut times V(r_cut)
    (hoomd/md/pytest/test_potential.py:901-949 checks the first on two particles). r_cut = 3 keeps the
    force jump at the cut-off (a kink in E that central differences would smear) negligible."""
    pos, vel, L = random_liquid(12, 0.7, seed=8, jitter=0.2)
    N = len(pos)
    fx = dict(L=L, pos=pos, vel=vel)

    def build(p, mode):
        sim = make_sim(hoomd, fx["L"], p, fx["vel"])
        cell = hoomd.md.nlist.Cell(buffer=0.3)
        pot = getattr(hoomd.md.pair, cls_name)(nlist=cell, default_r_cut=3.0, mode=mode)
        pot.params[("A", "A")] = params
        cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())
        sim.operations.integrator = hoomd.md.Integrator(dt=0.001, methods=[cv], forces=[pot])
        sim.run(0)
        return sim, cell, pot

    sim, cell, pot = build(pos, "none")
    f = pot.forces
    e0 = pot.energy
    assert np.abs(f.sum(axis=0)).max() <= 1e-6 * np.abs(f).sum()
    # energy shift: V(r_cut) per pair inside the cut-off
    pl = cell.pair_list
    d = pos[pl[:, 0]].astype(np.float64) - pos[pl[:, 1]].astype(np.float64)
    d -= float(L[0]) * np.rint(d / float(L[0]))
    n_in = int((np.einsum("ij,ij->i", d, d) < 9.0).sum())
    sim_s, cell_s, pot_s = build(pos, "shift")
    assert np.allclose(pot_s.forces, f, rtol=1e-6, atol=1e-6)
    two = np.array([[0, 0, 0], [0, 0, np.float32(3.0) - np.float32(1e-4)]], dtype=np.float32)
    sim2 = make_sim(hoomd, [20.0, 20.0, 20.0], two, np.zeros_like(two))
    c2 = hoomd.md.nlist.Cell(buffer=0.3)
    p2 = getattr(hoomd.md.pair, cls_name)(nlist=c2, default_r_cut=3.0)
    p2.params[("A", "A")] = params
    sim2.operations.integrator = hoomd.md.Integrator(
        dt=0.001, methods=[hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())], forces=[p2])
    sim2.run(0)
    v_cut = p2.energy  # V just inside r_cut
    if cls_name != "ForceShiftedLJ":
        assert abs((e0 - pot_s.energy) - n_in * v_cut) <= 2e-3 * max(abs(n_in * v_cut), 1e-3) + 1e-4 * abs(e0)
    # F = -dE/dx by central differences on three particles (fp32 energies: generous tolerance)
    h = np.float32(2.0 ** -7)
    scale = np.sqrt(np.mean(f ** 2))
    for i in (0, N // 2, N - 1):
        for c in range(3):
            pp, pm = pos.copy(), pos.copy()
            pp[i, c] += h
            pm[i, c] -= h
            # per-particle energies differenced in fp64: the unchanged particles cancel exactly, so the
            # total's fp32 rounding (1e-7 of a large sum) does not swamp the derivative
            ep = build(pp, "shift")[2].energies
            em = build(pm, "shift")[2].energies
            step = float(pp[i, c]) - float(pm[i, c])
            assert abs(-(ep - em).sum() / step - f[i, c]) <= 2e-2 * max(abs(f[i, c]), scale)


# ------------------------------------------------ net force, tail correction, observers ----
def _two_particle_sim(hoomd, d, L=20.0, vel=None):
    pos = np.array([[-d / 2, 0, 0.1], [d / 2, 0, 0.1]], dtype=np.float32)
    vel = np.zeros_like(pos) if vel is None else np.asarray(vel, dtype=np.float32)
    return make_sim(hoomd, [L, L, L], pos, vel)


def test_tail_corrections(hoomd, oracle):
    """hoomd/md/pytest/test_potential.py:1209-1290 (test_tail_corrections), same numbers: the energy and
    pressure corrections of the LJ potential against the closed forms, and against the oracle's
    restatement of PotentialPair::computeTailCorrection (hoomd/md/PotentialPair.h:255-341)."""
    sigma, epsilon, d_pair, r_cut = 1.0, 0.5, 1.5, 2.0
    v1 = np.array([0.46168675, -0.21020661, 0.21240303])
    out = {}
    for tail in (True, False):
        sim = _two_particle_sim(hoomd, d_pair, vel=[v1, -v1])
        cell = hoomd.md.nlist.Cell(buffer=0.4)
        lj = hoomd.md.pair.LJ(nlist=cell, default_r_cut=r_cut, mode="none", tail_correction=tail)
        lj.params[("A", "A")] = dict(sigma=sigma, epsilon=epsilon)
        lang = hoomd.md.methods.Langevin(filter=hoomd.filter.All(), kT=1.0)
        sim.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[lang], forces=[lj])
        sim.always_compute_pressure = True
        thermo = hoomd.md.compute.ThermodynamicQuantities(filter=hoomd.filter.All())
        sim.operations.computes.append(thermo)
        sim.run(0)
        out[tail] = (thermo.potential_energy, thermo.pressure, lj.energy, lj.additional_energy,
                     lj.additional_virial, sim.state.box.volume)
    N, volume = 2, out[True][5]
    rho = N / volume
    lj1, lj2 = 4 * epsilon * sigma ** 12, 4 * epsilon * sigma ** 6
    e_pair = 4 * epsilon * ((sigma / d_pair) ** 12 - (sigma / d_pair) ** 6)
    dE = 2 * N * np.pi * rho * (lj1 / 9 / r_cut ** 9 - lj2 / 3 / r_cut ** 3)
    dP = 4 / 6 * rho ** 2 * np.pi * (lj1 * 4 / 3 / r_cut ** 9 - lj2 * 2 / r_cut ** 3)
    ke = 0.5 * (v1 @ v1 + v1 @ v1)
    ljf = 24 * epsilon / d_pair * (2 * (sigma / d_pair) ** 12 - (sigma / d_pair) ** 6)
    p_plain = (2 * ke / 3 + ljf * d_pair / 3) / volume
    assert np.isclose(out[False][0], e_pair, rtol=1e-6)
    assert np.isclose(out[True][0], e_pair + dE, rtol=1e-6)
    assert np.isclose(out[False][1], p_plain, rtol=1e-6)
    assert np.isclose(out[True][1], p_plain + dP, rtol=1e-6)
    assert dE < 0 and dP < 0
    # Force.energy includes the external energy, additional_* expose it (hoomd/md/force.py:113-156)
    assert np.isclose(out[True][2], e_pair + dE, rtol=1e-6) and np.isclose(out[True][3], dE, rtol=1e-12)
    assert out[False][3] == 0.0
    e_orc, w_orc = oracle.lj_tail_correction([[epsilon, sigma]], [r_cut], [N], volume)
    assert np.isclose(out[True][3], e_orc, rtol=1e-12)
    av = np.asarray(out[True][4])
    assert np.allclose(av[[0, 3, 5]], w_orc, rtol=1e-12) and np.all(av[[1, 2, 4]] == 0)
    assert np.isclose(w_orc, volume * dP, rtol=1e-12)


def test_tail_correction_two_types(hoomd, oracle):
    """Per-type-pair densities in the correction (KA mixture)."""
    fx = dict(np.load(os.path.join(GOLDEN_DIR, "ka_2k.npz")))
    sim, cell, lj = lj_sim(hoomd, fx, KA_PARAMS, KA_RC, 0.3, "none", types=("A", "B"))
    lj.tail_correction = True
    sim.run(0)
    n_per_type = np.bincount(fx["typeid"], minlength=2)
    volume = float(np.prod(fx["L"].astype(np.float64)))
    order = [("A", "A"), ("A", "B"), ("A", "B"), ("B", "B")]  # tj * ntypes + ti
    e_orc, w_orc = oracle.lj_tail_correction([list(KA_PARAMS[k]) for k in order], [KA_RC[k] for k in order],
                                             n_per_type, volume)
    assert np.isclose(lj.additional_energy, e_orc, rtol=1e-12)
    assert np.allclose(np.asarray(lj.additional_virial)[[0, 3, 5]], w_orc, rtol=1e-12)


def test_net_force_of_two_pair_forces(hoomd, oracle):
    """Integrator::computeNetForce (hoomd/Integrator.cc:166-246): LJ + Yukawa on ONE neighbour list whose
    r_cut matrix is the maximum over both (hoomd/md/NeighborList.cc:452-529). Each force keeps its own
    per-particle arrays; the trajectory follows the sum; the list holds the larger cut-off."""
    fx = dict(np.load(os.path.join(GOLDEN_DIR, "lj_2k.npz")))
    N = len(fx["pos"])

    def build(which):
        sim = make_sim(hoomd, fx["L"], fx["pos"], fx["vel"], seed=11)
        sim.always_compute_pressure = True
        cell = hoomd.md.nlist.Cell(buffer=0.4)
        forces = []
        if "lj" in which:
            lj = hoomd.md.pair.LJ(nlist=cell, default_r_cut=2.5)
            lj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
            forces.append(lj)
        if "yuk" in which:
            yk = hoomd.md.pair.Yukawa(nlist=cell, default_r_cut=3.0)
            yk.params[("A", "A")] = dict(epsilon=0.7, kappa=1.3)
            forces.append(yk)
        cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())
        sim.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[cv], forces=forces)
        thermo = hoomd.md.compute.ThermodynamicQuantities(filter=hoomd.filter.All())
        thermo._attach(sim)
        return sim, cell, forces, thermo

    s_lj, c_lj, (lj0,), _ = build(("lj",))
    s_yk, c_yk, (yk0,), _ = build(("yuk",))
    s_both, c_both, (lj, yk), thermo = build(("lj", "yuk"))
    for s in (s_lj, s_yk, s_both):
        s.run(0)
    # the shared list carries max(r_cut) + buffer: the Yukawa-only list
    assert np.array_equal(canon_pairs(c_both.pair_list), canon_pairs(c_yk.pair_list))
    assert len(c_both.pair_list) > len(c_lj.pair_list)
    # per-force results are those of the force alone (summation order differs: longer rows)
    frms = np.sqrt(np.mean(np.sum(lj0.forces ** 2, axis=1)))
    assert np.max(np.linalg.norm(lj.forces - lj0.forces, axis=1)) <= 1e-5 * frms
    assert np.array_equal(yk.forces, yk0.forces)
    assert abs(lj.energy - lj0.energy) <= 1e-6 * abs(lj0.energy)
    # the oracle's LJ forces on the oracle's list with the larger cut-off
    box = oracle.make_box([float(v) for v in fx["L"]])
    types = np.zeros(N, dtype=np.uint32)
    nn, head, nl, _ = oracle.nlist(box, fx["pos"], types, [3.0], 0.4, full=True)
    f_orc, _ = oracle.pair_lj(box, fx["pos"], types, nn, head, nl, [[1.0, 1.0]], [2.5])
    assert force_err(lj.forces, f_orc[:, :3]) <= FORCE_RTOL
    # thermodynamics see the sum
    assert np.isclose(thermo.potential_energy, lj.energy + yk.energy, rtol=1e-9)
    p_sum = thermo.pressure
    # the trajectory integrates the NET force: v(dt) = v(0) + dt/2 (F_net(0) + F_net(dt)) / m exactly
    # (velocity Verlet, unit masses), with both forces' own getters giving F_net at either end
    ke0 = thermo.kinetic_energy
    f_net0 = lj.forces + yk.forces
    s_both.run(1)
    f_net1 = lj.forces + yk.forces
    v1 = s_both.state.get_snapshot().particles.velocity.astype(np.float64)
    v0 = fx["vel"].astype(np.float64)
    want = 0.5 * 0.005 * (f_net0 + f_net1)
    assert np.max(np.linalg.norm((v1 - v0) - want, axis=1)) <= 2e-5 * np.max(np.linalg.norm(want, axis=1)) + 1e-6
    assert np.isfinite(p_sum) and thermo.kinetic_energy != ke0
    # both forces stay valid through rebuilds: 50 steps conserve energy
    e0 = thermo.kinetic_energy + thermo.potential_energy
    s_both.run(50)
    e1 = thermo.kinetic_energy + thermo.potential_energy
    assert abs(e1 - e0) / N < 2e-3
    assert c_both.num_builds >= 1


def test_pressure_without_always_compute_pressure(hoomd, golden_dir):
    """ThermodynamicQuantities.pressure must not silently drop the virial when the simulation does not
    keep the pressure flag on (the reference raises the flag on the steps where it is logged): same value
    as with always_compute_pressure, and asking for it does not perturb the trajectory."""
    fx = dict(np.load(os.path.join(golden_dir, "lj_2k.npz")))

    def build(always):
        sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "none", seed=11)
        sim.always_compute_pressure = always
        thermo = hoomd.md.compute.ThermodynamicQuantities(filter=hoomd.filter.All())
        thermo._attach(sim)
        return sim, lj, thermo

    s_on, _, t_on = build(True)
    s_off, lj_off, t_off = build(False)
    s_plain, _, _ = build(False)
    for s in (s_on, s_off, s_plain):
        s.run(10)
    p_on, p_off = t_on.pressure, t_off.pressure
    assert np.isclose(p_off, p_on, rtol=1e-6) and abs(p_on) > 0.1
    assert lj_off.virials is None  # the flag is off again for the force's own getters
    assert t_off.degrees_of_freedom == 3 * len(fx["pos"]) - 3
    s_off.run(10)
    s_plain.run(10)
    assert np.array_equal(s_off.state.get_snapshot().particles.position,
                          s_plain.state.get_snapshot().particles.position)


def test_num_builds_covers_the_whole_run(hoomd, golden_dir):
    """nlist.num_builds / shortest_rebuild cover one Simulation.run (System::run resets them once,
    hoomd/System.cc:103-110), whether or not a writer splits the run into segments."""
    fx = dict(np.load(os.path.join(golden_dir, "lj_2k.npz")))
    sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "none", seed=11)
    sim.run(40)
    want = (cell.num_builds, cell.shortest_rebuild)
    sim2, cell2, lj2 = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "none", seed=11)
    import io
    logger = hoomd.logging.Logger(categories=["scalar"])
    logger.add(lj2, quantities=["energy"])
    sim2.operations.writers.append(hoomd.write.Table(hoomd.trigger.Periodic(1), logger, io.StringIO("")))
    sim2.run(40)
    assert (cell2.num_builds, cell2.shortest_rebuild) == want and want[0] >= 2


def test_nlist_r_cut_and_local_arrays(hoomd, oracle, golden_dir):
    """hoomd.md.nlist.NeighborList.r_cut (hoomd/md/nlist.py:127-131): the list's own cut-off widens the
    list beyond the pair force's (maximum over both, hoomd/md/NeighborList.cc:452-529) without changing
    the forces; cpu/gpu_local_nlist_arrays (hoomd/md/nlist.py:154-268) expose head_list / n_neigh / nlist
    of that build in local indices."""
    import torch
    fx = dict(np.load(os.path.join(golden_dir, "lj_2k.npz")))
    N = len(fx["pos"])
    sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "none")
    cell.r_cut[("A", "A")] = 3.0
    sim.run(0)
    box = oracle.make_box([float(v) for v in fx["L"]])
    types = np.zeros(N, dtype=np.uint32)
    nn, head, nl, _ = oracle.nlist(box, fx["pos"], types, [3.0], 0.4, full=True)
    assert np.array_equal(canon_pairs(cell.pair_list), oracle.pair_set(nn, head, nl))
    check_forces(lj, fx, "none_")
    # local arrays: rows in local (cell-ordered) indices; map to tags through the local snapshot
    loc_tag = np.asarray(list(sim.state._get_local())[0])
    with cell.cpu_local_nlist_arrays as data:
        assert not data.half_nlist and len(data.n_neigh) == N
        i = np.repeat(np.arange(N), data.n_neigh)
        k = np.concatenate([np.arange(h, h + n) for h, n in zip(data.head_list, data.n_neigh)])
        j = data.nlist[k.astype(np.int64)]
        pl = np.stack([loc_tag[i], loc_tag[j]], axis=1)
        assert np.array_equal(canon_pairs(pl), oracle.pair_set(nn, head, nl))
        assert int(data.n_neigh.sum()) == 2 * len(oracle.pair_set(nn, head, nl))
    with cell.gpu_local_nlist_arrays as gdata:
        d_nn = torch.as_tensor(gdata.n_neigh, device="cuda")
        assert np.array_equal(d_nn.cpu().numpy().view(np.uint32), data.n_neigh)
