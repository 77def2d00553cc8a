"""Runs the UNMODIFIED reference (HOOMD-blue 4.8.2 CPU build, baseline/_ref) twice on the lj_2k frame — once
with its own hoomd.md.pair.LJ, once with hoomd_b200_plugin.pair.LJ (this repository's CUDA path behind the
reference's ForceCompute interface, plugin/hoomd_b200_plugin) — and compares forces, energy, pressure and an
NVE trajectory. Launched as a subprocess by tests/test_plugin.py (the reference must be imported before numpy)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
sys.path.insert(0, os.path.join(ROOT, "plugin"))
import hoomd  # noqa: E402
import hoomd.md  # noqa: E402
import hoomd_b200_plugin  # noqa: E402
import numpy as np  # noqa: E402


def make(fx, which, mode="none"):
    sim = hoomd.Simulation(device=hoomd.device.CPU(), seed=3)
    snap = hoomd.Snapshot()
    L = fx["L"]
    snap.configuration.box = [float(L[0]), float(L[1]), float(L[2]), 0, 0, 0]
    snap.particles.N = len(fx["pos"])
    snap.particles.types = ["A"]
    snap.particles.position[:] = fx["pos"]
    snap.particles.velocity[:] = fx["vel"]
    snap.particles.image[:] = fx["image"]
    sim.create_state_from_snapshot(snap)
    if which == "reference":
        lj = hoomd.md.pair.LJ(nlist=hoomd.md.nlist.Cell(buffer=0.4), mode=mode)
        lj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
        lj.r_cut[("A", "A")] = 2.5
    else:
        lj = hoomd_b200_plugin.pair.LJ(default_r_cut=2.5, buffer=0.4, mode=mode)
        lj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
    nve = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())
    sim.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[nve], forces=[lj])
    tq = hoomd.md.compute.ThermodynamicQuantities(filter=hoomd.filter.All())
    sim.operations.computes.append(tq)
    sim.always_compute_pressure = True
    return sim, lj, tq


def main():
    fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "lj_2k.npz")))
    for mode in ("none", "shift"):
        res = {}
        for which in ("reference", "b200"):
            sim, lj, tq = make(fx, which, mode)
            sim.run(0)
            r = dict(f=np.array(lj.forces), e=np.array(lj.energies), E=lj.energy, P=tq.pressure,
                     w=np.array(lj.virials))
            sim.run(20)
            s = sim.state.get_snapshot()
            r.update(pos=np.array(s.particles.position), vel=np.array(s.particles.velocity), ke=tq.kinetic_energy,
                     pe=tq.potential_energy)
            if which == "b200":
                r["launches"] = lj.gpu_launches
            res[which] = r
        a, b = res["reference"], res["b200"]
        nrm = np.linalg.norm(a["f"], axis=1)
        ferr = np.max(np.linalg.norm(a["f"] - b["f"], axis=1) / np.maximum(nrm, np.sqrt(np.mean(nrm ** 2))))
        assert ferr <= 1e-5, ferr
        assert abs(a["E"] - b["E"]) <= 1e-6 * abs(a["E"]), (a["E"], b["E"])
        assert np.max(np.abs(a["e"] - b["e"])) <= 1e-5 * np.max(np.abs(a["e"]))
        assert abs(a["P"] - b["P"]) <= 1e-5 * abs(a["P"]), (a["P"], b["P"])
        assert np.max(np.abs(a["w"] - b["w"])) <= 2e-5 * np.max(np.abs(a["w"]))
        L = fx["L"].astype(np.float64)
        d = a["pos"] - b["pos"]
        d -= L * np.round(d / L)
        assert np.max(np.abs(d)) < 1e-4, np.max(np.abs(d))
        assert np.max(np.abs(a["vel"] - b["vel"])) < 1e-3
        assert abs(a["ke"] - b["ke"]) <= 1e-5 * abs(a["ke"]) and abs(a["pe"] - b["pe"]) <= 1e-5 * abs(a["pe"])
        assert b["launches"] > 0
        print(f"mode={mode}: force_err={ferr:.2e} dE={abs(a['E'] - b['E']) / abs(a['E']):.1e} "
              f"max|dx| after 20 NVE steps={np.max(np.abs(d)):.1e} launches={b['launches']}")
    print("plugin ok")


if __name__ == "__main__":
    main()
