"""Reference trajectories of integration methods on a sub-group of the particles (hoomd.filter.*):
tests/golden/filter_2k.npz, made with the HOOMD-blue 4.8.2 CPU build under baseline/_ref from the lj_2k
frame (tests/golden/lj_2k.npz).

    python tests/golden/make_filter_ref.py

  bussi_*   ConstantVolume(filter=Tags(even tags), thermostat=Bussi(kT=1.5)), 20 steps: final positions and
            velocities, the group's kinetic energy and degrees of freedom after every step
  nve_*     ConstantVolume(filter=SetDifference(All, Tags(0..99))), 20 steps
  lang_*    Langevin(filter=Tags(odd tags), kT=1.0, gamma=1), 5 steps
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.environ.get("HOOMD_REF_BUILD", os.path.join(os.path.dirname(os.path.dirname(HERE)), "baseline", "_ref"))
sys.path.insert(0, BUILD)
import hoomd  # noqa: E402  (before numpy: libstdc++ load order)
import hoomd.md  # noqa: E402
import numpy as np  # noqa: E402


def make(fx, seed):
    sim = hoomd.Simulation(device=hoomd.device.CPU(), seed=seed)
    snap = hoomd.Snapshot()
    L = fx["L"]
    snap.configuration.box = [float(L[0]), float(L[1]), float(L[2]), 0, 0, 0]
    snap.particles.N = len(fx["pos"])
    snap.particles.types = ["A"]
    snap.particles.position[:] = fx["pos"]
    snap.particles.velocity[:] = fx["vel"]
    snap.particles.image[:] = fx["image"]
    sim.create_state_from_snapshot(snap)
    cell = hoomd.md.nlist.Cell(buffer=0.4)
    lj = hoomd.md.pair.LJ(nlist=cell)
    lj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
    lj.r_cut[("A", "A")] = 2.5
    return sim, lj


def final(sim):
    s = sim.state.get_snapshot()
    return np.array(s.particles.position), np.array(s.particles.velocity), np.array(s.particles.image, dtype=np.int32)


def main():
    fx = dict(np.load(os.path.join(HERE, "lj_2k.npz")))
    N = len(fx["pos"])
    out = {}
    # (A) Bussi on the even tags
    sim, lj = make(fx, 11)
    flt = hoomd.filter.Tags(list(range(0, N, 2)))
    cv = hoomd.md.methods.ConstantVolume(filter=flt, thermostat=hoomd.md.methods.thermostats.Bussi(kT=1.5))
    sim.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[cv], forces=[lj])
    th = hoomd.md.compute.ThermodynamicQuantities(filter=flt)
    th_all = hoomd.md.compute.ThermodynamicQuantities(filter=hoomd.filter.All())
    sim.operations.computes.extend([th, th_all])
    sim.run(0)
    ke, dof = [th.kinetic_energy], th.degrees_of_freedom
    for _ in range(20):
        sim.run(1)
        ke.append(th.kinetic_energy)
    out["bussi_pos"], out["bussi_vel"], out["bussi_image"] = final(sim)
    out["bussi_ke"] = np.array(ke)
    out["bussi_dof"] = np.array(dof)
    out["bussi_dof_all"] = np.array(th_all.degrees_of_freedom)
    # (B) NVE on everything but the first 100 tags
    sim, lj = make(fx, 11)
    flt = hoomd.filter.SetDifference(hoomd.filter.All(), hoomd.filter.Tags(list(range(100))))
    cv = hoomd.md.methods.ConstantVolume(filter=flt)
    sim.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[cv], forces=[lj])
    th = hoomd.md.compute.ThermodynamicQuantities(filter=flt)
    sim.operations.computes.append(th)
    sim.run(20)
    out["nve_pos"], out["nve_vel"], out["nve_image"] = final(sim)
    out["nve_ke"] = np.array(th.kinetic_energy)
    out["nve_dof"] = np.array(th.degrees_of_freedom)
    # (C) Langevin on the odd tags
    sim, lj = make(fx, 11)
    flt = hoomd.filter.Tags(list(range(1, N, 2)))
    lg = hoomd.md.methods.Langevin(filter=flt, kT=1.0)
    sim.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[lg], forces=[lj])
    th = hoomd.md.compute.ThermodynamicQuantities(filter=flt)
    sim.operations.computes.append(th)
    sim.run(5)
    out["lang_pos"], out["lang_vel"], out["lang_image"] = final(sim)
    out["lang_dof"] = np.array(th.degrees_of_freedom)
    np.savez_compressed(os.path.join(HERE, "filter_2k.npz"), **out)
    print("filter_2k: dof", out["bussi_dof"], out["bussi_dof_all"], out["nve_dof"], out["lang_dof"], "ke", out["bussi_ke"][[0, -1]])


if __name__ == "__main__":
    main()
