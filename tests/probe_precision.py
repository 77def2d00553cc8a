"""Prints measured parity errors on the golden fixtures (diagnostic; not a test)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hoomd_blue_b200 as hoomd
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_gpu_parity import lj_sim, force_err, canon_pairs
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
for name in ("lj_2k", "c1_64k"):
    fx = dict(np.load(os.path.join(G, name + ".npz")))
    fx["typeid"] = np.zeros(len(fx["pos"]), dtype=np.uint32)
    pre = "none_" if name == "lj_2k" else ""
    for tpp in (4, 8, 32):
        sim, cell, lj = lj_sim(hoomd, fx, {("A", "A"): (1.0, 1.0)}, {("A", "A"): 2.5}, 0.4, "none", tpp=tpp)
        sim.run(0)
        f = lj.forces; fr = fx[pre + "force"]
        nr = np.linalg.norm(fr, axis=1); rms = np.sqrt(np.mean(nr**2))
        e = np.linalg.norm(f - fr, axis=1) / np.maximum(nr, rms)
        et = float(fx[pre + "total_energy"])
        print(name, "tpp", tpp, "force err max %.2e mean %.2e p99 %.2e rms_force %.2f" % (e.max(), e.mean(), np.quantile(e, 0.99), rms),
              "E rel %.2e" % (abs(lj.energy - et) / abs(et)), "stats", cell._stats())
