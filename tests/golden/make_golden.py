"""Generate tests/golden/*.npz from the REFERENCE ITSELF (HOOMD-blue v4.8.2 CPU build, fp64).

Needs the reference build that `baseline/build_ref.sh` produces from /root/reference into the
git-ignored baseline/_ref (SURVEY.md §8(c) recipe; override the location with HOOMD_REF_BUILD). Tests
read only the committed .npz/.json files; `tests/ref_run0.py` re-runs the reference on a frame when the
build travelled to the GPU box. (The fixtures committed in round 1 were made with the survey's
pre-session build of the same sources; the rebuilt reference reproduces them bit for bit: lj_2k
none_force differs by 0.0 from a fresh run(0).)

    python tests/golden/make_golden.py [--only NAME]

Fixtures (inputs are float32-exact like a GSD frame: hoomd/GSDReader.cc:194,235):
  ref_vectors.json    the reference's own KATs for this path, extracted verbatim as data:
                      true_pair_list.json, forces_and_energies.json[LJ, DPDConservative],
                      Philox4x32-10 kat_vectors:27-29, Seed/Counter KATs random_numbers_test.cc:456-533
  lj_2k.npz           LJ liquid N=2048 rho=0.9 (melted with Bussi kT=1.5): half pair list as tags,
                      per-particle force/energy/virial for mode none / shift / xplor, NVE and Bussi
                      20-step trajectories
  ka_2k.npz           Kob-Andersen 80:20 N=2048 rho=1.2, per-type-pair params, mode shift
  dpd_3k.npz          DPD N=3000 rho=3 A=25 gamma=4.5 kT=1 seed=5 at several timesteps
  c1_64k.npz          BASELINE.json configs[0]: LJ N=64,000 melted 1000 steps; pair-set sha256,
                      forces, totals
"""
import argparse
import hashlib
import json
import os
import sys

BUILD = os.environ.get("HOOMD_REF_BUILD",
                       os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                                    "baseline", "_ref"))
sys.path.insert(0, BUILD)
# hoomd must be imported BEFORE numpy in this container (libstdc++ load order; SURVEY.md §4 note)
import hoomd  # noqa: E402
import hoomd.md  # noqa: E402
import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def f32(a):
    return np.asarray(a, dtype=np.float32)


def make_sim(box_L, pos, vel, types, typeid, seed, image=None, mass=None):
    sim = hoomd.Simulation(device=hoomd.device.CPU(), seed=seed)
    snap = hoomd.Snapshot()
    snap.configuration.box = [float(box_L[0]), float(box_L[1]), float(box_L[2]), 0, 0, 0]
    snap.particles.N = len(pos)
    snap.particles.types = list(types)
    snap.particles.position[:] = pos
    snap.particles.velocity[:] = vel
    snap.particles.typeid[:] = typeid
    if image is not None:
        snap.particles.image[:] = image
    if mass is not None:
        snap.particles.mass[:] = mass
    sim.create_state_from_snapshot(snap)
    return sim


def frame32(sim):
    """Snapshot rounded to float32 and re-wrapped so every coordinate is in [-L/2, L/2)."""
    s = sim.state.get_snapshot()
    L = f32(s.configuration.box[:3]).astype(np.float64)
    pos = f32(s.particles.position).astype(np.float64)
    img = np.array(s.particles.image, dtype=np.int32)
    for d in range(3):
        hi = pos[:, d] >= L[d] / 2
        pos[hi, d] -= L[d]
        img[hi, d] += 1
        lo = pos[:, d] < -L[d] / 2
        pos[lo, d] += L[d]
        img[lo, d] -= 1
    pos = f32(pos)
    assert np.all(pos.astype(np.float64) < L / 2) and np.all(pos.astype(np.float64) >= -L / 2)
    return dict(L=f32(L), pos=pos, vel=f32(s.particles.velocity), image=img,
                typeid=np.array(s.particles.typeid, dtype=np.uint32))


def canon_pairs(pl):
    pl = np.asarray(pl, dtype=np.uint32).reshape(-1, 2)
    lo, hi = pl.min(axis=1), pl.max(axis=1)
    key = np.unique(lo.astype(np.uint64) << np.uint64(32) | hi.astype(np.uint64))
    return np.stack([(key >> np.uint64(32)).astype(np.uint32),
                     (key & np.uint64(0xFFFFFFFF)).astype(np.uint32)], axis=1)


def fcc(ncell, rho):
    a = (4.0 / rho) ** (1.0 / 3.0)
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) + 0.25
    g = np.arange(ncell)
    cells = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 1, 3)
    pos = ((cells + base[None]) * a).reshape(-1, 3)
    L = np.float32(ncell * a)
    return f32(pos - float(L) / 2), f32([L, L, L])


def eval_lj(sim_args, params, r_cut, buffer, mode, r_on=None, types=("A",)):
    sim = make_sim(**sim_args)
    sim.always_compute_pressure = True
    cell = hoomd.md.nlist.Cell(buffer=buffer)
    lj = hoomd.md.pair.LJ(nlist=cell, mode=mode)
    for (a, b), (eps, sig) in params.items():
        lj.params[(a, b)] = dict(epsilon=eps, sigma=sig)
        lj.r_cut[(a, b)] = r_cut[(a, b)]
        if r_on is not None:
            lj.r_on[(a, b)] = r_on[(a, b)]
    nve = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())
    sim.operations.integrator = hoomd.md.Integrator(dt=0.005, methods=[nve], forces=[lj])
    sim.run(0)
    return dict(pairs=canon_pairs(cell.pair_list), force=np.array(lj.forces),
                energy=np.array(lj.energies), virial=np.array(lj.virials), total_energy=lj.energy,
                dims=np.array(cell.dimensions))


def gen_ref_vectors():
    out = {}
    with open(f"{REF}/hoomd/md/pytest/true_pair_list.json") as f:
        out["true_pair_list"] = json.load(f)
    with open(f"{REF}/hoomd/md/pytest/forces_and_energies.json") as f:
        fe = json.load(f)
    out["forces_and_energies"] = {k: fe[k] for k in ("LJ", "DPDConservative", "DPD")}
    kats = []
    with open(f"{REF}/hoomd/extern/random123/tests/kat_vectors") as f:
        for line in f:
            t = line.split()
            if len(t) == 12 and t[0] == "philox4x32" and t[1] == "10":
                kats.append(dict(ctr=t[2:6], key=t[6:8], out=t[8:12]))
    out["philox4x32_10"] = kats
    # hoomd/test/random_numbers_test.cc:456-496
    out["seed_kat"] = dict(id=1, timestep="0xabcdef1234567890", seed="0x5eed",
                           key=["0x015eed12", "0x34567890"])
    with open(f"{HERE}/ref_vectors.json", "w") as f:
        json.dump(out, f)
    print("ref_vectors.json", {k: len(v) for k, v in out.items()})


def melt(pos, L, typeid, types, seed, kT, steps, params, r_cut, buffer=0.4, dt=0.005, mode="none"):
    rng = np.random.default_rng(seed)
    vel = f32(rng.normal(0, np.sqrt(kT), size=pos.shape))
    vel -= vel.mean(axis=0, dtype=np.float64).astype(np.float32)
    sim = make_sim(L, pos, vel, types, typeid, seed)
    cell = hoomd.md.nlist.Cell(buffer=buffer)
    lj = hoomd.md.pair.LJ(nlist=cell, mode=mode)
    for (a, b), (eps, sig) in params.items():
        lj.params[(a, b)] = dict(epsilon=eps, sigma=sig)
        lj.r_cut[(a, b)] = r_cut[(a, b)]
    cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All(),
                                         thermostat=hoomd.md.methods.thermostats.Bussi(kT=kT))
    sim.operations.integrator = hoomd.md.Integrator(dt=dt, methods=[cv], forces=[lj])
    sim.run(steps)
    return frame32(sim)


def trajectory(fr, types, params, r_cut, nsteps, dt, seed, thermostat_kT=None, buffer=0.4):
    sim = make_sim(fr["L"], fr["pos"], fr["vel"], types, fr["typeid"], seed, image=fr["image"])
    cell = hoomd.md.nlist.Cell(buffer=buffer)
    lj = hoomd.md.pair.LJ(nlist=cell)
    for (a, b), (eps, sig) in params.items():
        lj.params[(a, b)] = dict(epsilon=eps, sigma=sig)
        lj.r_cut[(a, b)] = r_cut[(a, b)]
    th = None if thermostat_kT is None else hoomd.md.methods.thermostats.Bussi(kT=thermostat_kT)
    cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All(), thermostat=th)
    sim.operations.integrator = hoomd.md.Integrator(dt=dt, methods=[cv], forces=[lj])
    thermo = hoomd.md.compute.ThermodynamicQuantities(filter=hoomd.filter.All())
    sim.operations.computes.append(thermo)
    ke, pe = [], []
    sim.run(0)
    ke.append(thermo.kinetic_energy)
    pe.append(thermo.potential_energy)
    for _ in range(nsteps):
        sim.run(1)
        ke.append(thermo.kinetic_energy)
        pe.append(thermo.potential_energy)
    s = sim.state.get_snapshot()
    return dict(pos=np.array(s.particles.position), vel=np.array(s.particles.velocity),
                image=np.array(s.particles.image, dtype=np.int32), ke=np.array(ke), pe=np.array(pe),
                builds=cell.num_builds)


def gen_lj_2k():
    pos, L = fcc(8, 0.9)
    N = len(pos)
    types = ["A"]
    params = {("A", "A"): (1.0, 1.0)}
    rc = {("A", "A"): 2.5}
    fr = melt(pos, L, np.zeros(N, dtype=np.uint32), types, 11, 1.5, 600, params, rc)
    args = dict(box_L=fr["L"], pos=fr["pos"], vel=fr["vel"], types=types, typeid=fr["typeid"],
                seed=11, image=fr["image"])
    out = dict(L=fr["L"], pos=fr["pos"], vel=fr["vel"], image=fr["image"], typeid=fr["typeid"])
    for mode, ron in (("none", None), ("shift", None), ("xplor", {("A", "A"): 2.0})):
        r = eval_lj(args, params, rc, 0.4, mode, r_on=ron)
        for k, v in r.items():
            out[f"{mode}_{k}"] = v
    t = trajectory(fr, types, params, rc, 20, 0.005, 11)
    for k, v in t.items():
        out[f"nve_{k}"] = v
    t = trajectory(fr, types, params, rc, 20, 0.005, 11, thermostat_kT=1.5)
    for k, v in t.items():
        out[f"bussi_{k}"] = v
    np.savez_compressed(f"{HERE}/lj_2k.npz", **out)
    print("lj_2k", N, "pairs", len(out["none_pairs"]), "E", out["none_total_energy"])


KA_PARAMS = {("A", "A"): (1.0, 1.0), ("A", "B"): (1.5, 0.8), ("B", "B"): (0.5, 0.88)}
KA_RC = {("A", "A"): 2.5, ("A", "B"): 2.0, ("B", "B"): 2.2}


def gen_ka_2k():
    pos, L = fcc(8, 1.2)
    N = len(pos)
    rng = np.random.default_rng(3)
    typeid = np.zeros(N, dtype=np.uint32)
    typeid[rng.permutation(N)[: N // 5]] = 1
    types = ["A", "B"]
    fr = melt(pos, L, typeid, types, 3, 1.0, 600, KA_PARAMS, KA_RC, buffer=0.3, dt=0.0025,
              mode="shift")
    args = dict(box_L=fr["L"], pos=fr["pos"], vel=fr["vel"], types=types, typeid=fr["typeid"],
                seed=3, image=fr["image"])
    out = dict(L=fr["L"], pos=fr["pos"], vel=fr["vel"], image=fr["image"], typeid=fr["typeid"])
    r = eval_lj(args, KA_PARAMS, KA_RC, 0.3, "shift")
    for k, v in r.items():
        out[f"shift_{k}"] = v
    np.savez_compressed(f"{HERE}/ka_2k.npz", **out)
    print("ka_2k", N, "pairs", len(out["shift_pairs"]), "E", out["shift_total_energy"])


def gen_dpd_3k():
    rng = np.random.default_rng(5)
    N, L = 3000, np.float32(10.0)
    pos = f32(rng.uniform(-5, 5, size=(N, 3)))
    pos[pos >= 5.0] = np.float32(-5.0)
    vel = f32(rng.normal(0, 1.0, size=(N, 3)))
    vel -= vel.mean(axis=0, dtype=np.float64).astype(np.float32)
    out = dict(L=f32([L, L, L]), pos=pos, vel=vel, typeid=np.zeros(N, dtype=np.uint32), seed=5,
               A=25.0, gamma=4.5, kT=1.0, r_cut=1.0, dt=0.01, buffer=0.4)

    def dpd_sim(ts):
        sim = hoomd.Simulation(device=hoomd.device.CPU(), seed=5)
        if ts:
            sim.timestep = ts
        snap = hoomd.Snapshot()
        snap.configuration.box = [10, 10, 10, 0, 0, 0]
        snap.particles.N = N
        snap.particles.types = ["A"]
        snap.particles.position[:] = pos
        snap.particles.velocity[:] = vel
        sim.create_state_from_snapshot(snap)
        sim.always_compute_pressure = True
        cell = hoomd.md.nlist.Cell(buffer=0.4)
        dpd = hoomd.md.pair.DPD(nlist=cell, kT=1.0, default_r_cut=1.0)
        dpd.params[("A", "A")] = dict(A=25.0, gamma=4.5)
        cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())
        sim.operations.integrator = hoomd.md.Integrator(dt=0.01, methods=[cv], forces=[dpd])
        return sim, cell, dpd

    steps = [0, 7, 4294967296 + 12345]
    out["timesteps"] = np.array(steps, dtype=np.uint64)
    for n, ts in enumerate(steps):
        sim, cell, dpd = dpd_sim(ts)
        sim.run(0)
        out[f"force_{n}"] = np.array(dpd.forces)
        out[f"energy_{n}"] = np.array(dpd.energies)
        out[f"virial_{n}"] = np.array(dpd.virials)
        if n == 0:
            out["pairs"] = canon_pairs(cell.pair_list)
    # short trajectory (10 steps) from timestep 0
    sim, cell, dpd = dpd_sim(0)
    sim.run(10)
    s = sim.state.get_snapshot()
    out["traj_pos"] = np.array(s.particles.position)
    out["traj_vel"] = np.array(s.particles.velocity)
    np.savez_compressed(f"{HERE}/dpd_3k.npz", **out)
    print("dpd_3k pairs", len(out["pairs"]))


def gen_c1_64k():
    n, rho = 40, 0.9
    N = n ** 3
    L = np.float32((N / rho) ** (1 / 3))
    g = (np.arange(n) + 0.5) * (float(L) / n) - float(L) / 2
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    pos = f32(np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1))
    types = ["A"]
    params = {("A", "A"): (1.0, 1.0)}
    rc = {("A", "A"): 2.5}
    fr = melt(pos, f32([L, L, L]), np.zeros(N, dtype=np.uint32), types, 7, 1.5, 1000, params, rc)
    args = dict(box_L=fr["L"], pos=fr["pos"], vel=fr["vel"], types=types, typeid=fr["typeid"],
                seed=7, image=fr["image"])
    r = eval_lj(args, params, rc, 0.4, "none")
    pairs = r["pairs"]
    out = dict(L=fr["L"], pos=fr["pos"], vel=fr["vel"], image=fr["image"],
               n_pairs=len(pairs), pairs_sha256=hashlib.sha256(pairs.tobytes()).hexdigest(),
               pairs_head=pairs[:64], force=r["force"], energy=f32(r["energy"]),
               total_energy=r["total_energy"], virial_sum=r["virial"].sum(axis=0), dims=r["dims"])
    np.savez_compressed(f"{HERE}/c1_64k.npz", **out)
    print("c1_64k pairs", len(pairs), out["pairs_sha256"], "E", r["total_energy"])


GEN = dict(ref_vectors=gen_ref_vectors, lj_2k=gen_lj_2k, ka_2k=gen_ka_2k, dpd_3k=gen_dpd_3k,
           c1_64k=gen_c1_64k)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    for name, fn in GEN.items():
        if a.only in (None, name):
            fn()
