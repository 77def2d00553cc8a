"""One single-core run of the REFERENCE (HOOMD-blue 4.8.2 CPU build under baseline/_ref) on the C1
workload: the reference-equilibrated LJ liquid N = 64,000 (tests/golden/c1_64k.npz), Cell(buffer=0.4),
pair.LJ(r_cut=2.5), ConstantVolume + Bussi(kT=1.5), dt = 0.005 — BASELINE.json configs[0].

    python baseline/ref_worker.py <warmup steps> <timed steps>

Prints one JSON line {"seconds", "steps", "N", "tps"}. bench.py starts one of these per host core
(HOOMD's MD CPU path is single-threaded per rank and this image has no MPI) and adds the throughputs up.
The reference is imported before numpy (libstdc++ load order).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.environ.get("HOOMD_REF_BUILD", os.path.join(ROOT, "baseline", "_ref")))
import hoomd  # noqa: E402
import hoomd.md  # noqa: E402,F401
import numpy as np  # noqa: E402


def main():
    warmup, steps = int(sys.argv[1]), int(sys.argv[2])
    fx = np.load(os.path.join(ROOT, "tests", "golden", "c1_64k.npz"))
    N = len(fx["pos"])
    sim = hoomd.Simulation(device=hoomd.device.CPU(), seed=7)
    snap = hoomd.Snapshot()
    L = [float(v) for v in fx["L"]]
    snap.configuration.box = [L[0], L[1], L[2], 0, 0, 0]
    snap.particles.N = N
    snap.particles.types = ["A"]
    snap.particles.position[:] = fx["pos"]
    snap.particles.velocity[:] = fx["vel"]
    sim.create_state_from_snapshot(snap)
    cell = hoomd.md.nlist.Cell(buffer=0.4)
    lj = hoomd.md.pair.LJ(nlist=cell, default_r_cut=2.5)
    lj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
    cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All(),
                                         thermostat=hoomd.md.methods.thermostats.Bussi(kT=1.5))
    sim.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[cv], forces=[lj])
    sim.run(warmup)
    t0 = time.perf_counter()
    sim.run(steps)
    dt = time.perf_counter() - t0
    print(json.dumps({"seconds": dt, "steps": steps, "N": N, "tps": steps / dt,
                      "version": hoomd.version.version}))


if __name__ == "__main__":
    main()
