"""run(0) of the REFERENCE (HOOMD-blue 4.8.2 CPU build under baseline/_ref, made by
baseline/build_ref.sh) on a frame handed over as .npz: pair set, per-particle forces, energy, virial.

Runs as a subprocess of the gpu tests (`python tests/ref_run0.py in.npz out.npz`): the reference must
be imported before numpy (libstdc++ load order), which a pytest process cannot
guarantee. When the reference build is not there, the oracle port (oracle/md_oracle.c) does the same
job and the result says so (`source`).

Input keys: L(3) pos(N,3) vel(N,3) typeid(N) types(list) kind('lj'|'dpd') pairs_a/pairs_b (type names
per parameter row) p0 p1 (epsilon/sigma or A/gamma) r_cut (per row) mode buffer dt kT seed timestep.
Output: n_pairs, pair_hash (order-independent 64-bit sum of mixed (min,max) tag pairs), pair_xor, degree(N),
force(N,3), energy, virial_sum(6), dims(3), source.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.environ.get("HOOMD_REF_BUILD", os.path.join(ROOT, "baseline", "_ref"))
_ref = None
if os.path.isdir(os.path.join(REF_DIR, "hoomd")) and not os.environ.get("HB2_FORCE_PORT"):
    sys.path.insert(0, REF_DIR)
    try:
        import hoomd as _ref  # noqa: E402  (before numpy)
        import hoomd.md  # noqa: E402,F401
    except Exception as ex:  # pragma: no cover - depends on the box
        sys.stderr.write(f"reference import failed ({ex}); using the oracle port\n")
        _ref = None
import numpy as np  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from pair_digest import degree, pair_digest  # noqa: E402


def run_reference(fx):
    hoomd = _ref
    N = len(fx["pos"])
    sim = hoomd.Simulation(device=hoomd.device.CPU(), seed=int(fx["seed"]))
    snap = hoomd.Snapshot()
    L = [float(v) for v in fx["L"]]
    snap.configuration.box = [L[0], L[1], L[2], 0, 0, 0]
    snap.particles.N = N
    types = [str(t) for t in fx["types"]]
    snap.particles.types = types
    snap.particles.position[:] = fx["pos"]
    snap.particles.velocity[:] = fx["vel"]
    snap.particles.typeid[:] = fx["typeid"]
    if int(fx["timestep"]):
        sim.timestep = int(fx["timestep"])  # allowed before the state exists (hoomd/simulation.py)
    sim.create_state_from_snapshot(snap)
    sim.always_compute_pressure = True
    cell = hoomd.md.nlist.Cell(buffer=float(fx["buffer"]))
    kind = str(fx["kind"])
    if kind == "lj":
        pot = hoomd.md.pair.LJ(nlist=cell, mode=str(fx["mode"]))
        for a, b, p0, p1, rc in zip(fx["pairs_a"], fx["pairs_b"], fx["p0"], fx["p1"], fx["r_cut"]):
            pot.params[(str(a), str(b))] = dict(epsilon=float(p0), sigma=float(p1))
            pot.r_cut[(str(a), str(b))] = float(rc)
    elif kind == "dpd":
        pot = hoomd.md.pair.DPD(nlist=cell, kT=float(fx["kT"]), default_r_cut=float(fx["r_cut"][0]))
        for a, b, p0, p1, rc in zip(fx["pairs_a"], fx["pairs_b"], fx["p0"], fx["p1"], fx["r_cut"]):
            pot.params[(str(a), str(b))] = dict(A=float(p0), gamma=float(p1))
            pot.r_cut[(str(a), str(b))] = float(rc)
    else:
        raise ValueError(kind)
    cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())
    sim.operations.integrator = hoomd.md.Integrator(dt=float(fx["dt"]), methods=[cv], forces=[pot])
    sim.run(0)
    pl = cell.pair_list
    n, h, x = pair_digest(pl)
    out = dict(n_pairs=n, pair_hash=np.uint64(h), pair_xor=np.uint64(x), degree=degree(pl, N),
               force=np.asarray(pot.forces, dtype=np.float64), energy=float(pot.energy),
               virial_sum=np.asarray(pot.virials, dtype=np.float64).sum(axis=0),
               dims=np.asarray(cell.dimensions, dtype=np.int64), source="reference")
    return out


def run_port(fx):
    sys.path.insert(0, ROOT)
    from oracle import md_oracle as orc
    orc.lib()
    N = len(fx["pos"])
    types = [str(t) for t in fx["types"]]
    nt = len(types)
    idx = {t: i for i, t in enumerate(types)}
    p0 = np.zeros((nt, nt))
    p1 = np.zeros((nt, nt))
    rc = np.zeros((nt, nt))
    for a, b, v0, v1, r in zip(fx["pairs_a"], fx["pairs_b"], fx["p0"], fx["p1"], fx["r_cut"]):
        i, j = idx[str(a)], idx[str(b)]
        p0[i, j] = p0[j, i] = float(v0)
        p1[i, j] = p1[j, i] = float(v1)
        rc[i, j] = rc[j, i] = float(r)
    box = orc.make_box([float(v) for v in fx["L"]])
    pos = np.asarray(fx["pos"], dtype=np.float32)
    tid = np.asarray(fx["typeid"], dtype=np.uint32)
    nn, head, nl, dim = orc.nlist(box, pos, tid, rc.reshape(-1).tolist(), float(fx["buffer"]), full=True)
    # half view of the full list: every unordered pair once (tags == indices here)
    i = np.repeat(np.arange(N, dtype=np.uint32), nn)
    keep = i < nl
    pl = np.stack([i[keep], nl[keep].astype(np.uint32)], axis=1)
    del i, keep
    n, h, x = pair_digest(pl)
    kind = str(fx["kind"])
    if kind == "lj":
        mode = {"none": 0, "shift": 1, "xplor": 2}[str(fx["mode"])]
        params = np.stack([p0.reshape(-1), p1.reshape(-1)], axis=1)
        f4, vir = orc.pair_lj(box, pos, tid, nn, head, nl, params, rc.reshape(-1), shift_mode=mode)
    else:
        params = np.stack([p0.reshape(-1), p1.reshape(-1)], axis=1)
        f4, vir = orc.pair_dpd(box, pos, np.asarray(fx["vel"], dtype=np.float32), tid,
                               np.arange(N, dtype=np.uint32), nn, head, nl, params, rc.reshape(-1),
                               int(fx["seed"]), int(fx["timestep"]), float(fx["dt"]), float(fx["kT"]))
    f4 = np.asarray(f4, dtype=np.float64).reshape(N, 4)
    vir = np.asarray(vir, dtype=np.float64).reshape(6, -1)[:, :N]
    return dict(n_pairs=n, pair_hash=np.uint64(h), pair_xor=np.uint64(x), degree=degree(pl, N),
                force=f4[:, :3].copy(), energy=float(f4[:, 3].sum()), virial_sum=vir.sum(axis=1),
                dims=np.asarray(dim, dtype=np.int64), source="port")


def main():
    fx = dict(np.load(sys.argv[1], allow_pickle=False))
    out = run_reference(fx) if _ref is not None else run_port(fx)
    np.savez(sys.argv[2], **out)


if __name__ == "__main__":
    main()
