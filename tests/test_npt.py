"""hoomd.md.methods.ConstantPressure (hoomd/md/methods/methods.py:177-580, hoomd/md/TwoStepConstantPressure.cc):
host-side API checks on the CPU; on the GPU, trajectories against fixtures the reference itself produced
(tests/golden/npt_2k.npz, tests/golden/make_npt_ref.py): box lengths, barostat momenta, pressure and kinetic
energy after every step, final state, barostat energy. The barostat arithmetic alone is replayed to
round-off by tests/test_oracle_golden.py::test_npt_barostat_replay."""
import os

import numpy as np
import pytest

CASES = {
    "a": dict(S=2.0, tauS=0.5, couple="xyz", thermostat=None, gamma=0.0,
              box_dof=[True, True, True, False, False, False], tail=False, thermalize=False, steps=40, seed=21),
    "b": dict(S=[3.0, 4.0, 5.0, 0, 0, 0], tauS=1.0, couple="none", thermostat=1.5, gamma=0.5,
              box_dof=[True, True, True, False, False, False], tail=False, thermalize=False, steps=30, seed=22),
    "c": dict(S=1.0, tauS=0.8, couple="xy", thermostat=None, gamma=0.0,
              box_dof=[True, True, False, False, False, False], tail=True, thermalize=True, steps=30, seed=23),
}


def test_constant_pressure_api():
    """Argument handling as the reference's ParameterDict (scalar S -> six components, couple names, six
    box_dof flags, barostat_dof of length six) — no device needed."""
    import hoomd_blue_b200 as hoomd
    M = hoomd.md.methods
    npt = M.ConstantPressure(filter=hoomd.filter.All(), S=2, tauS=1.0, couple="xyz")
    assert npt.S == (2.0, 2.0, 2.0, 0.0, 0.0, 0.0)
    assert npt.box_dof == [True, True, True, False, False, False] and npt.rescale_all is False
    assert npt.gamma == 0.0 and npt.thermostat is None and npt.barostat_dof == (0.0,) * 6
    npt.S = [1, 2, 3, 0, 0, 0]
    assert npt.S == (1.0, 2.0, 3.0, 0.0, 0.0, 0.0)
    npt.barostat_dof = (0.1, 0, 0, 0.2, 0, 0.3)
    assert npt.barostat_dof == (0.1, 0.0, 0.0, 0.2, 0.0, 0.3)
    with pytest.raises(ValueError):
        npt.couple = "zz"
    with pytest.raises(ValueError):
        npt.S = [1, 2, 3]
    with pytest.raises(ValueError):
        npt.barostat_dof = (1, 2, 3)
    with pytest.raises(RuntimeError):
        npt.thermalize_barostat_dof()  # before run(0), like the reference
    with pytest.raises(TypeError):
        M.ConstantPressure(filter=None, S=1, tauS=1, couple="none")


def _sim(hoomd, fx, c):
    sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=c["seed"])
    L = fx["L"]
    sim.create_state_from_snapshot(hoomd.Snapshot.from_arrays(list(L) + [0, 0, 0], fx["pos"], fx["vel"], None,
                                                              ("A",), fx["image"]))
    cell = hoomd.md.nlist.Cell(buffer=0.4)
    lj = hoomd.md.pair.LJ(nlist=cell, tail_correction=c["tail"])
    lj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
    lj.r_cut[("A", "A")] = 2.5
    th = hoomd.md.methods.thermostats.Bussi(kT=c["thermostat"]) if c["thermostat"] else None
    npt = hoomd.md.methods.ConstantPressure(filter=hoomd.filter.All(), S=c["S"], tauS=c["tauS"], couple=c["couple"],
                                            thermostat=th, box_dof=c["box_dof"], gamma=c["gamma"])
    sim.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[npt], forces=[lj])
    tq = hoomd.md.compute.ThermodynamicQuantities(filter=hoomd.filter.All())
    sim.operations.computes.append(tq)
    sim.always_compute_pressure = True
    return sim, npt, tq, cell


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_constant_pressure_vs_reference(golden_dir, name):
    import hoomd_blue_b200 as hoomd
    fx = np.load(os.path.join(golden_dir, "lj_2k.npz"))
    ref = np.load(os.path.join(golden_dir, "npt_2k.npz"))
    c = CASES[name]
    sim, npt, tq, cell = _sim(hoomd, fx, c)
    sim.run(0)
    if c["thermalize"]:
        npt.thermalize_barostat_dof()
    box, nu, P, ke, builds, Pt = [], [], [], [], 0, []

    def rec():
        b = sim.state.box
        box.append([b.Lx, b.Ly, b.Lz])
        nu.append(list(npt.barostat_dof))
        P.append(tq.pressure)
        ke.append(tq.kinetic_energy)
        Pt.append(list(tq.pressure_tensor))

    rec()
    for _ in range(c["steps"]):
        sim.run(1)
        builds += cell.num_builds
        rec()
    box, nu, P, ke = np.array(box), np.array(nu), np.array(P), np.array(ke)
    r = {k: ref[f"{name}_{k}"] for k in ("box", "nu", "P", "ke", "pos", "vel", "image", "benergy", "builds")}
    # thermalize_barostat_dof draws the reference's numbers (host Philox, RNG id 27)
    np.testing.assert_allclose(nu[0], r["nu"][0], rtol=1e-12, atol=1e-15)
    # fp32 particle state against the fp64 reference: the box follows to 1e-6, the momenta to 1e-4 of their scale
    np.testing.assert_allclose(box, r["box"], rtol=2e-6)
    scale = np.max(np.abs(r["nu"]))
    assert np.max(np.abs(nu - r["nu"])) <= 2e-4 * scale
    np.testing.assert_allclose(P, r["P"], rtol=0, atol=2e-4 * np.max(np.abs(r["P"])))
    np.testing.assert_allclose(ke, r["ke"], rtol=5e-5)
    # ThermodynamicQuantities.pressure_tensor (kinetic tensor + virial sums; off-diagonals are small numbers
    # made of large cancelling terms: tolerance on the scale of the diagonal)
    np.testing.assert_allclose(np.array(Pt), ref[f"{name}_Pt"], rtol=0, atol=2e-4 * np.max(np.abs(r["P"])))
    assert abs(npt.barostat_energy - float(r["benergy"])) <= 5e-4 * abs(float(r["benergy"]))
    assert abs(builds - int(r["builds"])) <= 1
    s = sim.state.get_snapshot()
    Lf = r["box"][-1]
    np.testing.assert_allclose(np.array(s.configuration.box[:3]), Lf, rtol=2e-6)
    d = s.particles.position.astype(np.float64) - r["pos"]
    d -= Lf * np.round(d / Lf)
    assert np.max(np.abs(d)) < 2e-3
    assert np.max(np.abs(s.particles.velocity - r["vel"])) < 5e-3
    # unwrapped coordinates agree too (images are updated when the rescaled box wraps a particle)
    un = s.particles.position.astype(np.float64) + s.particles.image * Lf - (r["pos"] + r["image"] * Lf)
    assert np.max(np.abs(un)) < 2e-3

    # one run(steps) is the same trajectory as steps x run(1), bit for bit
    sim2, npt2, _, _ = _sim(hoomd, fx, c)
    sim2.run(0)
    if c["thermalize"]:
        npt2.thermalize_barostat_dof()
    sim2.run(c["steps"])
    s2 = sim2.state.get_snapshot()
    assert np.array_equal(s2.particles.position, s.particles.position)
    assert np.array_equal(s2.particles.velocity, s.particles.velocity)
    assert tuple(npt2.barostat_dof) == tuple(npt.barostat_dof)
    assert list(s2.configuration.box) == list(s.configuration.box)


@pytest.mark.gpu
def test_constant_pressure_reaches_set_point(golden_dir):
    """NPT with a Bussi thermostat relaxes the box until the time-averaged pressure is the set point; the
    neighbour list follows the shrinking / growing box (rebuilds through the lambda-scaled distance check)."""
    import hoomd_blue_b200 as hoomd
    fx = np.load(os.path.join(golden_dir, "lj_2k.npz"))
    c = dict(CASES["b"], S=4.0, couple="xyz", gamma=0.0, tauS=0.5, seed=5)
    sim, npt, tq, cell = _sim(hoomd, fx, c)
    sim.run(6000)
    Ps, Vs = [], []
    for _ in range(40):
        sim.run(100)
        Ps.append(tq.pressure)
        Vs.append(sim.state.box.volume)
    assert abs(np.mean(Ps) - 4.0) < 0.25, np.mean(Ps)
    assert 2.0 * tq.kinetic_energy / tq.degrees_of_freedom == pytest.approx(1.5, rel=0.05)
    assert np.std(Vs) / np.mean(Vs) < 0.05
    # the shear degrees of freedom need a triclinic box: refused, not ignored
    sim3, _, _, _ = _sim(hoomd, fx, dict(c, box_dof=[True, True, True, True, False, False]))
    with pytest.raises(Exception, match="triclinic"):
        sim3.run(0)
