"""Particle filters for the integration methods (hoomd/filter/*.py, hoomd/md/methods/methods.py:145-175):
host-side semantics on the CPU; on the GPU, trajectories of ConstantVolume / Bussi / Langevin acting on a
sub-group against fixtures produced by the reference itself (tests/golden/filter_2k.npz,
tests/golden/make_filter_ref.py)."""
import os

import numpy as np
import pytest


class _State:
    N_particles = 10
    particle_types = ["A", "B"]

    def get_snapshot(self):
        class P:
            typeid = np.array([0, 1, 0, 1, 0, 1, 0, 1, 0, 1], dtype=np.uint32)

        class S:
            particles = P

        return S


def test_filter_semantics():
    """Construction, selection, equality and hashing as hoomd/filter/{all_,null,tags,type_,set_}.py."""
    from hoomd_blue_b200 import filter as F
    st = _State()
    assert np.array_equal(F.All()(st), np.arange(10))
    assert len(F.Null()(st)) == 0
    t = F.Tags([5, 1, 1, 3])
    assert np.array_equal(t(st), [1, 3, 5]) and t(st).dtype == np.uint32
    assert t == F.Tags([1, 3, 5]) and hash(t) == hash(F.Tags([3, 5, 1])) and t != F.Tags([1, 3])
    assert np.array_equal(F.Type(["B"])(st), [1, 3, 5, 7, 9])
    assert F.Type(["A", "B"]) == F.Type(["B", "A"]) and F.Type(["A"]) != F.Type(["B"])
    a, b = F.Tags([0, 1, 2, 3]), F.Tags([2, 3, 4])
    assert np.array_equal(F.SetDifference(a, b)(st), [0, 1])
    assert np.array_equal(F.Union(a, b)(st), [0, 1, 2, 3, 4])
    assert np.array_equal(F.Intersection(a, b)(st), [2, 3])
    assert F.Union(a, b) == F.Union(b, a) and F.SetDifference(a, b) != F.SetDifference(b, a)
    assert hash(F.Union(a, b)) == hash(F.Union(b, a))
    with pytest.raises(ValueError):
        F.Union(a, F.Tags([0, 1, 2, 3]))
    with pytest.raises(ValueError):
        F.Tags([10])(st)
    assert F.All() == F.All() and F.All() != F.Null()


def _sim(hoomd, fx, method):
    sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=11)
    L = fx["L"]
    sim.create_state_from_snapshot(hoomd.Snapshot.from_arrays(list(L) + [0, 0, 0], fx["pos"], fx["vel"], None,
                                                              ("A",), fx["image"]))
    cell = hoomd.md.nlist.Cell(buffer=0.4)
    lj = hoomd.md.pair.LJ(nlist=cell)
    lj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
    lj.r_cut[("A", "A")] = 2.5
    sim.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[method], forces=[lj])
    return sim


def _compare(sim, fx, ref, prefix, frozen, tol_pos, tol_vel):
    s = sim.state.get_snapshot()
    L = fx["L"].astype(np.float64)
    d = s.particles.position.astype(np.float64) - ref[prefix + "pos"]
    d -= L * np.round(d / L)
    assert np.max(np.abs(d)) < tol_pos
    assert np.max(np.abs(s.particles.velocity - ref[prefix + "vel"])) < tol_vel
    # particles outside the group keep their state bit for bit
    assert np.array_equal(s.particles.position[frozen], fx["pos"][frozen])
    assert np.array_equal(s.particles.velocity[frozen], fx["vel"][frozen])
    moved = np.setdiff1d(np.arange(len(fx["pos"])), frozen)
    assert np.all(np.any(s.particles.position[moved] != fx["pos"][moved], axis=1))


@pytest.mark.gpu
def test_bussi_on_a_tag_group(golden_dir):
    import hoomd_blue_b200 as hoomd
    fx = dict(np.load(os.path.join(golden_dir, "lj_2k.npz")))
    ref = dict(np.load(os.path.join(golden_dir, "filter_2k.npz")))
    N = len(fx["pos"])
    flt = hoomd.filter.Tags(np.arange(0, N, 2))
    cv = hoomd.md.methods.ConstantVolume(filter=flt, thermostat=hoomd.md.methods.thermostats.Bussi(kT=1.5))
    sim = _sim(hoomd, fx, cv)
    th = hoomd.md.compute.ThermodynamicQuantities(filter=flt)
    th_all = hoomd.md.compute.ThermodynamicQuantities(filter=hoomd.filter.All())
    th._attach(sim)
    th_all._attach(sim)
    sim.run(0)
    ke = [th.kinetic_energy]
    for _ in range(20):
        sim.run(1)
        ke.append(th.kinetic_energy)
    assert th.degrees_of_freedom == float(ref["bussi_dof"]) == 3 * (N // 2)
    assert th_all.degrees_of_freedom == float(ref["bussi_dof_all"])
    # the thermostat rescales with the GROUP's kinetic energy and degrees of freedom: the trace follows the
    # reference's (same Philox stream keyed by timestep and seed)
    assert np.allclose(ke, ref["bussi_ke"], rtol=5e-5)
    _compare(sim, fx, ref, "bussi_", np.arange(1, N, 2), 2e-4, 2e-3)


@pytest.mark.gpu
def test_nve_on_a_set_difference(golden_dir):
    import hoomd_blue_b200 as hoomd
    fx = dict(np.load(os.path.join(golden_dir, "lj_2k.npz")))
    ref = dict(np.load(os.path.join(golden_dir, "filter_2k.npz")))
    flt = hoomd.filter.SetDifference(hoomd.filter.All(), hoomd.filter.Tags(np.arange(100)))
    sim = _sim(hoomd, fx, hoomd.md.methods.ConstantVolume(filter=flt))
    th = hoomd.md.compute.ThermodynamicQuantities(filter=flt)
    th._attach(sim)
    sim.run(20)
    assert th.degrees_of_freedom == float(ref["nve_dof"])
    assert abs(th.kinetic_energy - float(ref["nve_ke"])) <= 5e-5 * float(ref["nve_ke"])
    _compare(sim, fx, ref, "nve_", np.arange(100), 2e-4, 2e-3)
    # the same run cut into segments (stand-alone kernels at the seams) is the same trajectory
    sim2 = _sim(hoomd, fx, hoomd.md.methods.ConstantVolume(filter=flt))
    for n in (7, 1, 12):
        sim2.run(n)
    a, b = sim.state.get_snapshot(), sim2.state.get_snapshot()
    assert np.array_equal(a.particles.position, b.particles.position)
    assert np.array_equal(a.particles.velocity, b.particles.velocity)


@pytest.mark.gpu
def test_langevin_on_a_tag_group(golden_dir):
    import hoomd_blue_b200 as hoomd
    fx = dict(np.load(os.path.join(golden_dir, "lj_2k.npz")))
    ref = dict(np.load(os.path.join(golden_dir, "filter_2k.npz")))
    N = len(fx["pos"])
    flt = hoomd.filter.Tags(np.arange(1, N, 2))
    sim = _sim(hoomd, fx, hoomd.md.methods.Langevin(filter=flt, kT=1.0))
    th = hoomd.md.compute.ThermodynamicQuantities(filter=flt)
    th._attach(sim)
    sim.run(5)
    assert th.degrees_of_freedom == float(ref["lang_dof"])
    # the noise is keyed by (timestep, seed, tag): the reference's draws, so the trajectories agree closely
    _compare(sim, fx, ref, "lang_", np.arange(0, N, 2), 1e-4, 1e-3)
