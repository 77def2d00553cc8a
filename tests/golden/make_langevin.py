"""Langevin trajectory and thermalised momenta from the REFERENCE ITSELF (baseline/_ref): pins the noise
streams that the reference's own tests never check numerically.

    python tests/golden/make_langevin.py        # writes tests/golden/langevin_2k.npz

lj_2k.npz frame (N = 2048), LJ(eps = sigma = 1, r_cut 2.5, mode none), Cell(buffer 0.4),
methods.Langevin(kT = 1.5, default_gamma = 2.0), dt = 0.005, Simulation(seed = 11): positions and
velocities after 1 and 20 steps. And State.thermalize_particle_momenta(kT = 1.5) with seed 7 at timestep 0.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.environ.get("HOOMD_REF_BUILD", os.path.join(ROOT, "baseline", "_ref")))
import hoomd  # noqa: E402
import hoomd.md  # noqa: E402,F401
import numpy as np  # noqa: E402

fx = np.load(os.path.join(ROOT, "tests", "golden", "lj_2k.npz"))
N = len(fx["pos"])


def make(seed):
    sim = hoomd.Simulation(device=hoomd.device.CPU(), seed=seed)
    snap = hoomd.Snapshot()
    snap.configuration.box = [float(v) for v in fx["L"]] + [0, 0, 0]
    snap.particles.N = N
    snap.particles.types = ["A"]
    snap.particles.position[:] = fx["pos"]
    snap.particles.velocity[:] = fx["vel"]
    sim.create_state_from_snapshot(snap)
    return sim


sim = make(11)
cell = hoomd.md.nlist.Cell(buffer=0.4)
lj = hoomd.md.pair.LJ(nlist=cell, default_r_cut=2.5)
lj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
lv = hoomd.md.methods.Langevin(filter=hoomd.filter.All(), kT=1.5, default_gamma=2.0)
sim.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[lv], forces=[lj])
out = {}
sim.run(1)
s = sim.state.get_snapshot()
out["pos_1"], out["vel_1"] = np.array(s.particles.position), np.array(s.particles.velocity)
sim.run(19)
s = sim.state.get_snapshot()
out["pos_20"], out["vel_20"] = np.array(s.particles.position), np.array(s.particles.velocity)

th = make(7)
th.state.thermalize_particle_momenta(filter=hoomd.filter.All(), kT=1.5)
out["thermal_vel"] = np.array(th.state.get_snapshot().particles.velocity)
np.savez(os.path.join(ROOT, "tests", "golden", "langevin_2k.npz"), kT=1.5, gamma=2.0, dt=0.005, seed=11,
         thermal_seed=7, source=f"HOOMD-blue {hoomd.version.version} CPU fp64 (baseline/_ref)", **out)
print("written", {k: v.shape for k, v in out.items()})
