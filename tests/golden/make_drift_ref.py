"""NVE energy trace over 1e5 steps of an 8,000-particle LJ liquid computed by the REFERENCE ITSELF
(HOOMD-blue 4.8.2 CPU build under baseline/_ref, made by baseline/build_ref.sh) — SURVEY.md 8(d) gate 4.

    python tests/golden/make_drift_ref.py        # ~30 min on one core; writes tests/golden/nve_drift_8k.npz

The frame: 20^3 simple-cubic lattice at rho = 0.9, momenta thermalised at kT = 1.5 (seed 7), melted for
2,000 steps under Bussi on the reference, rounded to float32 like a GSD frame and re-wrapped. From that
frame the reference runs ConstantVolume(thermostat=None), LJ(eps = sigma = 1, r_cut = 2.5, mode 'shift'),
Cell(buffer = 0.4), dt = 0.005 and logs kinetic and potential energy every 1,000 steps; the GPU test
starts from the same frame.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.environ.get("HOOMD_REF_BUILD", os.path.join(ROOT, "baseline", "_ref")))
import hoomd  # noqa: E402  (before numpy)
import hoomd.md  # noqa: E402,F401
import numpy as np  # noqa: E402

N_SIDE, RHO, KT, DT, BUF, RC, EVERY, TOTAL = 20, 0.9, 1.5, 0.005, 0.4, 2.5, 1000, 100_000


def ops(sim, thermostat):
    cell = hoomd.md.nlist.Cell(buffer=BUF)
    lj = hoomd.md.pair.LJ(nlist=cell, default_r_cut=RC, mode="shift")
    lj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
    cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All(), thermostat=thermostat)
    sim.operations.integrator = hoomd.md.Integrator(dt=DT, methods=[cv], forces=[lj])
    thermo = hoomd.md.compute.ThermodynamicQuantities(filter=hoomd.filter.All())
    sim.operations.computes.append(thermo)
    return cell, thermo


def main():
    N = N_SIDE ** 3
    a = RHO ** (-1.0 / 3.0)
    L = np.float32(N_SIDE * a)
    g = (np.arange(N_SIDE) + 0.5) * a - float(L) / 2
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    snap = hoomd.Snapshot()
    snap.configuration.box = [float(L)] * 3 + [0, 0, 0]
    snap.particles.N = N
    snap.particles.types = ["A"]
    snap.particles.position[:] = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    sim = hoomd.Simulation(device=hoomd.device.CPU(), seed=7)
    sim.create_state_from_snapshot(snap)
    sim.state.thermalize_particle_momenta(filter=hoomd.filter.All(), kT=KT)
    ops(sim, hoomd.md.methods.thermostats.Bussi(kT=KT))
    sim.run(2000)
    s = sim.state.get_snapshot()
    Lf = np.float32(s.configuration.box[0]).astype(np.float64)
    pos = np.asarray(s.particles.position, dtype=np.float32).astype(np.float64)
    pos = np.where(pos >= Lf / 2, pos - Lf, pos)
    pos = np.where(pos < -Lf / 2, pos + Lf, pos).astype(np.float32)
    vel = np.asarray(s.particles.velocity, dtype=np.float32)

    snap2 = hoomd.Snapshot()
    snap2.configuration.box = [float(np.float32(Lf))] * 3 + [0, 0, 0]
    snap2.particles.N = N
    snap2.particles.types = ["A"]
    snap2.particles.position[:] = pos
    snap2.particles.velocity[:] = vel
    nve = hoomd.Simulation(device=hoomd.device.CPU(), seed=7)
    nve.create_state_from_snapshot(snap2)
    cell, thermo = ops(nve, None)
    nve.run(0)
    ke, pe = [thermo.kinetic_energy], [thermo.potential_energy]
    builds = 0
    for t in range(EVERY, TOTAL + 1, EVERY):
        nve.run(EVERY)
        builds += cell.num_builds
        ke.append(thermo.kinetic_energy)
        pe.append(thermo.potential_energy)
        if t % 10000 == 0:
            print(t, (ke[-1] + pe[-1] - ke[0] - pe[0]) / N, flush=True)
    np.savez(os.path.join(ROOT, "tests", "golden", "nve_drift_8k.npz"), L=np.float32(Lf), pos=pos, vel=vel,
             ke=np.array(ke), pe=np.array(pe), every=EVERY, dt=DT, buffer=BUF, r_cut=RC, builds=builds,
             mode="shift", source=f"HOOMD-blue {hoomd.version.version} CPU fp64 (baseline/_ref)")


if __name__ == "__main__":
    main()
