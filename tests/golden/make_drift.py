"""NVE energy trace over 1e5 steps of the 2,048-particle LJ liquid (tests/golden/lj_2k.npz, a frame
made by the reference) computed with the CPU oracle port (fp64, pinned to the reference's 20-step
trajectories in tests/test_oracle_golden.py). The reference itself would need ~25 min per 1e5 steps at
N=8,000; the port reproduces its arithmetic and runs 2,048 particles in ~7 min.

    python tests/golden/make_drift.py        # writes tests/golden/nve_drift_2k.npz

The oracle's step loop takes the accelerations as given; this harness starts them at zero, so the
first half kick is missing and the first sample (t = 0) sits 7e-3 per particle below the rest. The
dynamics after that are ordinary NVE; consumers skip sample 0.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import md_oracle as orc  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "lj_2k.npz"))
L = [float(v) for v in g["L"]]
box = orc.make_box(L)
pos = g["pos"].astype(np.float64).copy()
vel = g["vel"].astype(np.float64).copy()
N = len(pos)
accel = np.zeros_like(pos)
image = g["image"].astype(np.int32).copy()
mass = np.ones(N)
types = np.zeros(N, dtype=np.uint32)
tags = np.arange(N, dtype=np.uint32)
DT, BUF, RC, CHUNK, TOTAL = 0.005, 0.4, 2.5, 1000, 100_000


def energy(r):
    ke = 0.5 * (mass[:, None] * vel ** 2).sum()
    return ke, r["force"][:, 3].sum()


r = orc.run(box, pos, vel, accel, image, mass, types, tags, [[1.0, 1.0]], [RC], 0, DT, BUF, shift_mode=1)
ke, pe = [], []
k, p = energy(r)
ke.append(k)
pe.append(p)
t = 0
builds = 0
while t < TOTAL:
    r = orc.run(box, pos, vel, accel, image, mass, types, tags, [[1.0, 1.0]], [RC], CHUNK, DT, BUF,
                shift_mode=1, timestep=t)
    t += CHUNK
    builds += r["builds"]
    k, p = energy(r)
    ke.append(k)
    pe.append(p)
    if t % 10000 == 0:
        print(t, (ke[-1] + pe[-1] - ke[0] - pe[0]) / N, flush=True)
np.savez(os.path.join(ROOT, "tests", "golden", "nve_drift_2k.npz"), ke=np.array(ke), pe=np.array(pe),
         every=CHUNK, dt=DT, buffer=BUF, r_cut=RC, builds=builds, mode="shift")
