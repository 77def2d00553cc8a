"""Reference trajectories of hoomd.md.methods.ConstantPressure: tests/golden/npt_2k.npz, made with the
HOOMD-blue 4.8.2 CPU build under baseline/_ref (baseline/build_ref.sh) from the lj_2k frame.

    python tests/golden/make_npt_ref.py

  a_*  NPH: S = 2.0, tauS = 0.5, couple 'xyz', 40 steps
  b_*  NPT: Bussi(kT = 1.5), S = [3, 4, 5, 0, 0, 0], tauS = 1.0, couple 'none', Langevin piston gamma = 0.5
       (pins the barostat's noise stream), 30 steps
  c_*  couple 'xy', box_dof x and y only, LJ with tail_correction (external virial), barostat momenta drawn
       by thermalize_barostat_dof after run(0), 30 steps
Per case: box lengths, barostat_dof, pressure, pressure tensor (xx, xy, xz, yy, yz, zz) and kinetic energy after every step (index 0: after run(0)),
final positions / velocities / images, barostat_energy at the end.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.environ.get("HOOMD_REF_BUILD", os.path.join(os.path.dirname(os.path.dirname(HERE)), "baseline", "_ref"))
sys.path.insert(0, BUILD)
import hoomd  # noqa: E402  (before numpy: libstdc++ load order)
import hoomd.md  # noqa: E402
import numpy as np  # noqa: E402

CASES = {
    "a": dict(S=2.0, tauS=0.5, couple="xyz", thermostat=None, gamma=0.0,
              box_dof=[True, True, True, False, False, False], tail=False, thermalize=False, steps=40, seed=21),
    "b": dict(S=[3.0, 4.0, 5.0, 0, 0, 0], tauS=1.0, couple="none", thermostat=1.5, gamma=0.5,
              box_dof=[True, True, True, False, False, False], tail=False, thermalize=False, steps=30, seed=22),
    "c": dict(S=1.0, tauS=0.8, couple="xy", thermostat=None, gamma=0.0,
              box_dof=[True, True, False, False, False, False], tail=True, thermalize=True, steps=30, seed=23),
}


def run_case(fx, c):
    sim = hoomd.Simulation(device=hoomd.device.CPU(), seed=c["seed"])
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
    s = sim.state.get_snapshot()
    return dict(box=np.array(box), nu=np.array(nu), P=np.array(P), ke=np.array(ke), Pt=np.array(Pt),
                pos=np.array(s.particles.position), vel=np.array(s.particles.velocity),
                image=np.array(s.particles.image, dtype=np.int32), benergy=np.array(npt.barostat_energy),
                builds=np.array(builds), pe=np.array(tq.potential_energy))


def main():
    fx = dict(np.load(os.path.join(HERE, "lj_2k.npz")))
    out = {}
    for name, c in CASES.items():
        r = run_case(fx, c)
        for k, v in r.items():
            out[f"{name}_{k}"] = v
        print(name, "L", r["box"][0], "->", r["box"][-1], "nu", r["nu"][-1][[0, 3, 5]], "P", r["P"][0], r["P"][-1],
              "builds", r["builds"])
    np.savez_compressed(os.path.join(HERE, "npt_2k.npz"), **out)


if __name__ == "__main__":
    main()
