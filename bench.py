#!/usr/bin/env python
"""bench.py — LJ particle-timesteps/s of the B200 pair-force step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one MD timestep (step one -> nlist check/rebuild -> pair forces -> step two) of the LJ
liquid of BASELINE.json configs[1]: N = 1,048,576 (FCC 64^3 cells melted to a liquid), rho = 0.9,
kT = 1.5, r_cut = 2.5, buffer = 0.4, Cell nlist + ConstantVolume with the Bussi thermostat, dt = 0.005.
Inputs are synthetic (seeded lattice + Maxwell velocities, equilibrated on the device before timing).

Prints ONE JSON line (see the task contract): value = particle-steps/s with the state resident in HBM;
e2e = the same metric through the public API with HOST snapshots in and out inside the timed region;
roofline = the pair-force kernel against the measured HBM peak; cpu_baseline = the reference's own CPU
implementation (HOOMD-blue built by baseline/build_ref.sh into baseline/_ref, one single-core process per
host core; the oracle port when that build is absent) timed on this box's host cores on a bounded sample.
`--impl reference` is that CPU arm alone. With N > 1 GPUs the line also carries `parity` (the decomposed
C1 frame against the reference fixture) and `strong_baseline` (the same 64 M system on ONE GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

_emit = print  # replaced in __main__ by a writer to the saved stdout descriptor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RHO, KT, RCUT, BUFFER, DT = 0.9, 1.5, 2.5, 0.4, 0.005
METRIC = "LJ particle-timesteps/s"
UNIT = "particle-steps/s"


def fcc_liquid(ncell, seed=1):
    a = (4.0 / RHO) ** (1.0 / 3.0)
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) + 0.25
    g = np.arange(ncell)
    cells = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 1, 3)
    L = np.float32(ncell * a)
    pos = ((cells + base[None]) * a).reshape(-1, 3) - float(L) / 2
    pos = pos.astype(np.float32)
    rng = np.random.default_rng(seed)
    vel = rng.normal(0, np.sqrt(KT), size=pos.shape).astype(np.float32)
    vel -= vel.mean(axis=0, dtype=np.float64).astype(np.float32)
    return pos, vel, L


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def roofline_pair(prof, nbar, N, hbm, peak_src, world=1):
    t = prof["force_ms"] / max(prof["force_launches"], 1) * 1e-3
    alg = (4.0 * nbar + 44.0) * N
    # traffic = dram bytes per launch from an `ncu --set full` capture of THIS workload (C2 on one GPU;
    # profiles/pair_force_summary.json). It is a recorded measurement, not taken in this run, and it
    # does not apply to other particle counts: null there.
    traffic, traffic_src = None, None
    summ = os.path.join(ROOT, "profiles", "pair_force_summary.json")
    if world == 1 and N == 1048576 and os.path.exists(summ):
        with open(summ) as f:
            js = json.load(f)
        traffic = js.get("dram_bytes_per_launch")
        traffic_src = "ncu capture " + js.get("capture", "profiles/pair_force_summary.json")
    # (algorithmic bytes follow SURVEY 8(d): 4 B per listed neighbour + 44 B per particle; the shipped kernel
    # reads 2 B per neighbour from the tile-shared list, so its measured traffic is below that figure)
    return {"kernel": "pair_force_tile_kernel<EvaluatorPairLJ>", "bound": "hbm",
            "achieved": alg / t / 1e9, "peak": hbm, "unit": "GB/s", "frac": alg / t / 1e9 / hbm,
            "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": alg,
            "launch_ms": t * 1e3, "peak_source": peak_src}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""

    def __init__(self, index=0):
        self.rows = []
        self.nvml_rows = None
        self.proc = None
        self.index = index

    def _nvml_loop(self):
        import pynvml as nv
        bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.nvml_rows.append((sm, [n for n, b in bits.items() if r & b]))
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        # NVML in-process, one sample every 5 ms (nvidia-smi -lms needs ~0.3 s to start, longer than the timed
        # region at C2); the nvidia-smi loop below is the fallback where pynvml is missing
        self.nvml_rows = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists ordinals
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[self.index])
            self._h = nv.nvmlDeviceGetHandleByIndex(idx)
            self._max = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
            self.nvml_rows = []
            self._stop = threading.Event()
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            self.proc = None
            return
        except Exception:
            self.nvml_rows = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml_rows is not None:
            self._stop.set()
            self.t.join(timeout=1)
            sm = [r[0] for r in self.nvml_rows]
            reasons = sorted({n for r in self.nvml_rows for n in r[1]})
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self._max), "reasons": reasons,
                    "samples": len(sm), "source": "nvml, 5 ms period"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def replicated_c1(repl, grid, coord):
    """BASELINE.json configs[3]: the equilibrated C1 frame (tests/golden/c1_64k.npz, made by the
    reference) replicated repl = (nx, ny, nz) times like hoomd.State.replicate (hoomd/state.py:580-609).
    Returns only the replicas that overlap the domain `coord` of `grid` (tags are global)."""
    fx = np.load(os.path.join(ROOT, "tests", "golden", "c1_64k.npz"))
    p0, v0, L0 = fx["pos"].astype(np.float64), fx["vel"], float(fx["L"][0])
    n0 = len(p0)
    L = np.array([np.float32(L0 * r) for r in repl], dtype=np.float64)
    tags, pos, vel = [], [], []
    for i in range(repl[0]):
        for j in range(repl[1]):
            for k in range(repl[2]):
                idx = (i, j, k)
                keep = True
                for d in range(3):
                    # replica slab [idx, idx+1) * L0 against the domain [coord, coord+1) * L / grid
                    lo, hi = idx[d] * L0, (idx[d] + 1) * L0
                    dlo, dhi = coord[d] * L[d] / grid[d], (coord[d] + 1) * L[d] / grid[d]
                    keep = keep and hi > dlo - 1e-3 and lo < dhi + 1e-3
                if not keep:
                    continue
                shift = (np.array(idx) + 0.5) * L0 - L / 2
                pos.append((p0 + shift).astype(np.float32))
                vel.append(v0)
                r = (i * repl[1] + j) * repl[2] + k
                tags.append(np.arange(n0, dtype=np.uint32) + np.uint32(r * n0))
    n_global = n0 * repl[0] * repl[1] * repl[2]
    pos = np.concatenate(pos)
    # a replica edge can round onto the box face
    hi = np.nextafter(np.float32(L / 2), np.float32(0)).astype(np.float32)
    pos = np.minimum(np.maximum(pos, (-L / 2).astype(np.float32)), hi)
    return n_global, np.concatenate(tags), pos, np.concatenate(vel), L


def attach_ops(hoomd, sim, thermostat=True):
    cell = hoomd.md.nlist.Cell(buffer=BUFFER)
    lj = hoomd.md.pair.LJ(nlist=cell, default_r_cut=RCUT)
    lj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
    th = hoomd.md.methods.thermostats.Bussi(kT=KT) if thermostat else None
    cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All(), thermostat=th)
    sim.operations.integrator = hoomd.md.Integrator(dt=DT, methods=[cv], forces=[lj])
    return sim, cell, lj


def build_sim(hoomd, pos, vel, L, device_id=0, seed=1, thermostat=True):
    sim = hoomd.Simulation(device=hoomd.device.GPU(gpu_id=device_id), seed=seed)
    sim.create_state_from_snapshot(hoomd.Snapshot.from_arrays([L, L, L, 0, 0, 0], pos, vel))
    return attach_ops(hoomd, sim, thermostat)


def oracle_run(snapshot_pos, snapshot_vel, L, nsteps, seed=1):
    """The CPU oracle port on one core (half list, like hoomd.device.CPU). Returns seconds, builds."""
    from oracle import md_oracle as orc
    N = len(snapshot_pos)
    box = orc.make_box([float(L)] * 3)
    pos = snapshot_pos.astype(np.float64).copy()
    vel = snapshot_vel.astype(np.float64).copy()
    accel = np.zeros_like(pos)
    image = np.zeros((N, 3), dtype=np.int32)
    mass = np.ones(N)
    types = np.zeros(N, dtype=np.uint32)
    tags = np.arange(N, dtype=np.uint32)
    # one untimed step-0 call warms the accelerations (prepRun) and page-faults the arrays
    orc.run(box, pos, vel, accel, image, mass, types, tags, [[1.0, 1.0]], [RCUT], 0, DT, BUFFER,
            thermostat="bussi", kT=KT, tau=0.0, seed=seed)
    t0 = time.perf_counter()
    r = orc.run(box, pos, vel, accel, image, mass, types, tags, [[1.0, 1.0]], [RCUT], nsteps, DT,
                BUFFER, thermostat="bussi", kT=KT, tau=0.0, seed=seed)
    return time.perf_counter() - t0, r["builds"]


def c1_sample():
    """The reference-generated equilibrated C1 liquid (N=64,000, same rho/kT as the benchmark state):
    the bounded CPU sample."""
    fx = np.load(os.path.join(ROOT, "tests", "golden", "c1_64k.npz"))
    return fx["pos"], fx["vel"], fx["L"][0]


def cpu_baseline(pos, vel, L, nsteps, replicas):
    """`replicas` independent single-core oracle-port runs concurrently; aggregate throughput."""
    out = [None] * replicas

    def work(i):
        out[i] = oracle_run(pos, vel, L, nsteps)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(replicas)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    wall = time.perf_counter() - t0
    # each replica times its own steps; the aggregate uses the slowest replica's clock
    slowest = max(o[0] for o in out)
    return replicas * nsteps * len(pos) / slowest, wall


def reference_available():
    return os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "hoomd")) and not os.environ.get("HB2_FORCE_PORT")


def reference_cpu(warmup, nsteps, replicas):
    """The reference's own CPU path (hoomd.device.CPU of the build under baseline/_ref) on the C1
    workload: one single-core process per replica (HOOMD's MD CPU path is single-threaded per rank and
    this image has no MPI: SURVEY.md 8(d)), all running concurrently; aggregate particle-steps/s on the
    slowest replica's clock. Returns (value, wall seconds, N, version) or None when a worker fails."""
    worker = os.path.join(ROOT, "baseline", "ref_worker.py")
    t0 = time.perf_counter()
    procs = [subprocess.Popen([sys.executable, worker, str(warmup), str(nsteps)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for _ in range(replicas)]
    outs = [p.communicate() for p in procs]
    wall = time.perf_counter() - t0
    res = []
    for p, (so, se) in zip(procs, outs):
        if p.returncode != 0:
            sys.stderr.write("reference worker failed:\n" + se[-2000:] + "\n")
            return None
        res.append(json.loads(so.strip().splitlines()[-1]))
    slowest = max(r["seconds"] for r in res)
    return replicas * nsteps * res[0]["N"] / slowest, wall, res[0]["N"], res[0]["version"]


def cpu_arm(warmup, nsteps):
    """(value, cpu_baseline dict, N) of the CPU arm on this box's host cores: the reference when its
    build travelled with the repository, else the oracle port."""
    cores = os.cpu_count() or 1
    if reference_available():
        r = reference_cpu(warmup, nsteps, cores)
        if r is not None:
            value, wall, n, ver = r
            return value, {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                           "sample": f"{cores} independent single-core hoomd.device.CPU processes "
                                     f"(HOOMD-blue {ver}, fp64, baseline/_ref) x {nsteps} steps after "
                                     f"{warmup} warm-up steps of the N={n} equilibrated liquid of "
                                     "configs[0] (same rho, kT, r_cut, buffer, Bussi as the 1M workload; "
                                     "no MPI in this image: throughput-equivalent, not strong-scaled; "
                                     f"{wall:.1f} s wall)"}, n
    pos, vel, L = c1_sample()
    value, wall = cpu_baseline(pos, vel, L, nsteps, cores)
    return value, {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{cores} independent single-core oracle-port replicas x {nsteps} steps of "
                             f"the N={len(pos)} equilibrated liquid of configs[0] (same rho, kT, r_cut, "
                             f"buffer; baseline/_ref not present; {wall:.1f} s wall)"}, len(pos)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # a bounded sample of the workload: the C1 liquid (configs[0]) advanced `steps` steps after `warmup`
    # warm-up steps per host core; HB2_REF_STEPS caps the timed steps (CPU: ~9 steps/s per core)
    nsteps = max(1, min(args.steps, int(os.environ.get("HB2_REF_STEPS", "1000"))))
    warmup = max(0, min(args.warmup, 200))
    value, cb, n = cpu_arm(warmup, nsteps)
    cores = cb["cores"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": nsteps, "warmup": warmup, "ms_per_step": 1e3 * n * cores / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"LJ liquid N={n} rho=0.9 kT=1.5 r_cut=2.5 buffer=0.4, Cell nlist + "
                               "ConstantVolume/Bussi, dt=0.005 (BASELINE.json configs[0], the reference's "
                               f"CPU-runnable case) x {cores} single-core replicas"},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(json.dumps(line))


def parity_check(hoomd, dist, rank, world, local):
    """The reference-generated C1 frame (tests/golden/c1_64k.npz) on the same device(s) and
    decomposition as the benchmark: pair set against the reference's sha256, forces, energy."""
    import hashlib
    fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "c1_64k.npz")))
    sim = hoomd.Simulation(device=hoomd.device.GPU(gpu_id=local), seed=7)
    sim.create_state_from_snapshot(hoomd.Snapshot.from_arrays(list(fx["L"]) + [0, 0, 0], fx["pos"], fx["vel"],
                                                              None, ("A",), fx["image"]))
    sim, cell, lj = attach_ops(hoomd, sim, thermostat=False)
    sim.run(0)
    pl = np.asarray(cell.pair_list, dtype=np.uint32).reshape(-1, 2)   # collective with N > 1
    f = lj.forces
    e = lj.energy
    grid = list(sim.state.domain_decomposition)
    out = None
    if rank == 0:
        lo, hi = pl.min(axis=1), pl.max(axis=1)
        key = np.unique(lo.astype(np.uint64) << np.uint64(32) | hi.astype(np.uint64))
        pairs = np.stack([(key >> np.uint64(32)).astype(np.uint32),
                          (key & np.uint64(0xFFFFFFFF)).astype(np.uint32)], axis=1)
        n_ref = np.linalg.norm(fx["force"], axis=1)
        floor = np.sqrt(np.mean(n_ref ** 2))
        ferr = float(np.max(np.linalg.norm(f - fx["force"], axis=1) / np.maximum(n_ref, floor)))
        e_ref = float(fx["total_energy"])
        out = {"frame": "tests/golden/c1_64k.npz (reference-generated, N=64000)", "decomposition": grid,
               "pairs_listed": int(len(pl)), "pairs_unique": int(len(pairs)),
               "pairs_reference": int(fx["n_pairs"]),
               "pair_set_sha256_match": hashlib.sha256(pairs.tobytes()).hexdigest() == str(fx["pairs_sha256"]),
               "force_rel_err_max": ferr, "energy_rel_err": abs(e - e_ref) / abs(e_ref),
               "pass": bool(len(pl) == len(pairs) == int(fx["n_pairs"])
                            and hashlib.sha256(pairs.tobytes()).hexdigest() == str(fx["pairs_sha256"])
                            and ferr <= 1e-5 and abs(e - e_ref) <= 1e-6 * abs(e_ref))}
    sim.operations.integrator._detach()
    sim.state._destroy()
    return out


def strong_baseline(hoomd, local, repl, steps, warmup):
    """The N > 1 workload (configs[3], 64 M particles) on ONE B200 — it fits in 180 GB — so that the
    multi-GPU numbers have a same-system denominator."""
    from hoomd_blue_b200.communicator import Communicator
    one = Communicator()
    one.rank, one.num_ranks = 0, 1
    N, tags, pos, vel, L = replicated_c1(repl, (1, 1, 1), (0, 0, 0))
    sim = hoomd.Simulation(device=hoomd.device.GPU(gpu_id=local, communicator=one), seed=1)
    sim._create_state_from_local([float(v) for v in L] + [0, 0, 0], ["A"], N, tags, pos, vel)
    del tags, pos, vel
    sim, cell, lj = attach_ops(hoomd, sim)
    sim.run(50)
    sim.run(warmup)
    sim.run(steps)
    ms = sim._last_device_ms
    out = {"n_gpus": 1, "N": int(N), "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
           "value": N * steps / (ms * 1e-3), "unit": UNIT,
           "what": "the same 64 M-particle system, undivided, on rank 0's GPU (device-timed)"}
    sim.operations.integrator._detach()
    sim.state._destroy()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import hoomd_blue_b200 as hoomd
    from hoomd_blue_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # Workload. 1 GPU: configs[1], the 1,048,576-particle liquid (FCC lattice melted on the device).
    # N > 1 GPUs: configs[3], the equilibrated 64,000-particle reference frame replicated 10x10x10 =
    # 64 M particles in ONE box, spatially decomposed over the N GPUs (strong scaling: total work
    # fixed). HB2_BENCH_REPL=nx,ny,nz overrides the replication (e.g. a per-GPU-fixed series).
    repl_env = os.environ.get("HB2_BENCH_REPL", "")
    if world == 1 and not repl_env:
        ncell = int(os.environ.get("HB2_BENCH_NCELL", "64"))
        pos, vel, L = fcc_liquid(ncell, seed=1)
        N = len(pos)
        sim, cell, lj = build_sim(hoomd, pos, vel, L, device_id=local, seed=1)
        equil = int(os.environ.get("HB2_BENCH_EQUIL", "2000"))
        workload = (f"LJ liquid N={N} rho=0.9 kT=1.5 r_cut=2.5 buffer=0.4, Cell nlist + "
                    "ConstantVolume/Bussi, dt=0.005 (BASELINE.json configs[1])")
        scaling, grid = "weak", (1, 1, 1)
    else:
        repl = tuple(int(x) for x in (repl_env or "10,10,10").split(","))
        fx = np.load(os.path.join(ROOT, "tests", "golden", "c1_64k.npz"))
        L0 = float(fx["L"][0])
        Lbox = [float(np.float32(L0 * r)) for r in repl]
        g = (__import__("ctypes").c_uint32 * 3)()
        _lib.check(_lib.lib().hb2_domain_find_decomposition(world, *Lbox, g))
        grid = tuple(int(x) for x in g)
        coord = (rank % grid[0], (rank // grid[0]) % grid[1], rank // (grid[0] * grid[1]))
        N, tags, pos, vel, _ = replicated_c1(repl, grid, coord)
        sim = hoomd.Simulation(device=hoomd.device.GPU(gpu_id=local), seed=1)
        sim._create_state_from_local(Lbox + [0, 0, 0], ["A"], N, tags, pos, vel,
                                     domain_decomposition=grid if world > 1 else (None, None, None))
        del tags, pos, vel
        sim, cell, lj = attach_ops(hoomd, sim)
        equil = int(os.environ.get("HB2_BENCH_EQUIL", "50"))
        workload = (f"LJ liquid N={N} (reference-equilibrated 64,000-particle frame replicated "
                    f"{repl[0]}x{repl[1]}x{repl[2]}) rho=0.9 kT=1.5 r_cut=2.5 buffer=0.4, Cell nlist + "
                    f"ConstantVolume/Bussi, dt=0.005, domain decomposition {grid[0]}x{grid[1]}x{grid[2]} "
                    "with NCCL ghost exchange (BASELINE.json configs[3])")
        scaling = "strong"
    tpp = int(os.environ.get("HB2_TPP", "0"))
    if tpp:
        lj.set_threads_per_particle(tpp)
    # melt / equilibrate (untimed)
    sim.run(equil)
    stats = cell._stats()
    n_local, n_ghost = sim.state._N_local, sim.state._N_ghost

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # clocks and throttle reasons are sampled from the warm-up to the end of the e2e loop: the timed
    # region alone is shorter than nvidia-smi's sampling period at C2
    sampler = ClockSampler(local)
    sampler.start()
    sim.run(args.warmup)
    L0c = _lib.lib().hb2_launch_count()
    barrier()
    sim.run(args.steps)
    ms = sim._last_device_ms  # CUDA events on the engine's stream around the K steps
    barrier()
    launches = int(_lib.lib().hb2_launch_count() - L0c)
    builds = cell.num_builds
    ms = max_over_ranks(ms)
    value = N * args.steps / (ms * 1e-3)

    # ---- per-phase device times of the same loop (separate, event-instrumented run: the events
    # would perturb the timed region above)
    integ = sim.operations.integrator
    P = max(20, min(args.steps, 300))
    integ._set_profiling(True)
    sim.run(P)
    prof = integ._profile()
    prof_run_ms = max_over_ranks(sim._last_device_ms) / P
    integ._set_profiling(False)
    pstats = cell._stats()
    if os.environ.get("HB2_BENCH_DEBUG"):
        for _ in range(3):
            sim.run(args.steps)
            if rank == 0:
                print("debug run ms/step", sim._last_device_ms / args.steps, "builds", cell.num_builds,
                      file=sys.stderr)

    # ---- e2e: host arrays in -> run(S) -> host arrays + energy out, all inside the timed region,
    # through the public API (1 GPU: Snapshot in / out; N GPUs: each rank's own particles in / out)
    S = int(os.environ.get("HB2_E2E_BLOCK", "100"))
    calls = max(1, min(5, args.steps // S if args.steps >= S else 1))
    if world == 1:
        # the host copy of the state lives in page-locked memory (contract: inputs come from pinned
        # host memory), so both directions are plain DMA
        snap = sim.state.get_snapshot(into=hoomd.Snapshot(pinned=True))
    else:
        loc = sim.state._get_local(pinned=True)
    barrier()
    t0 = time.perf_counter()
    dbg = os.environ.get("HB2_BENCH_DEBUG")
    for _ in range(calls):
        ta = time.perf_counter()
        if world == 1:
            sim.state.set_snapshot(snap)
            tb = time.perf_counter()
            sim.run(S)
            tc = time.perf_counter()
            snap = sim.state.get_snapshot(into=snap)
        else:
            tag_, pos_, vel_, mass_, tid_, img_ = loc
            sim.state._set_tagged(N, tag_, pos_, vel_, tid_, mass_, img_)
            tb = time.perf_counter()
            sim.run(S)
            tc = time.perf_counter()
            loc = sim.state._get_local(pinned=True, into=loc)
        td = time.perf_counter()
        e = lj.energy
        if dbg and rank == 0:
            print("e2e call: set %.1f ms, run %.1f ms (device %.1f ms), get %.1f ms, energy %.1f ms"
                  % ((tb - ta) * 1e3, (tc - tb) * 1e3, sim._last_device_ms, (td - tc) * 1e3,
                     (time.perf_counter() - td) * 1e3), file=sys.stderr)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = N * S * calls / e2e_s
    clocks = sampler.stop()
    clocks["window"] = "warm-up + timed steps + instrumented steps + e2e loop"
    parity = parity_check(hoomd, dist if world > 1 else None, rank, world, local)
    strong = None
    if world > 1 and (os.environ.get("HB2_BENCH_STRONG", "" if repl_env else "1") == "1"):
        # free this rank's share of the decomposed system before the single-GPU run of the same system
        integ = None
        sim.operations.integrator._detach()
        sim.state._destroy()
        barrier()
        if rank == 0:
            strong = strong_baseline(hoomd, local, repl, min(args.steps, 100), min(args.warmup, 20))
        barrier()
    h2d = N * (12 + 12 + 4 + 4 + 12) + (0 if world == 1 else 4 * N)
    d2h = h2d + 8

    line = None
    if rank == 0:
        hbm, peak_src = peaks()
        nbar, k = stats["mean_neighbors"], args.steps / max(builds, 1)
        bytes_step = (4 * nbar + 44) + 168 + 32 + (4 * nbar + 168) / k
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload,
                       "l2": "working set per GPU (nlist rows %.0f MB at 2 B per entry + state) exceeds the "
                             "126 MB L2; no flush" % (n_local * stats["pitch"] * 2 / 1e6),
                       "particles_per_gpu_rank0": n_local, "ghosts_rank0": n_ghost,
                       "decomposition": list(grid),
                       "mean_neighbors": nbar, "steps_per_rebuild": k, "nlist_builds": builds,
                       "threads_per_particle": tpp or 8},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / S,
                    "d2h_bytes_per_step": d2h / S, "steps_per_call": S, "calls": calls,
                    "what": ("state.set_snapshot(pinned host snapshot) + sim.run(S) + state.get_snapshot() + energy"
                             if world == 1 else
                             "per rank: state._set_tagged(pinned host arrays) + sim.run(S) + "
                             "state._get_local() + energy (collective)")},
            "gpu_launches": launches,
            "parity": parity,
            "strong_baseline": strong,
            # dominant kernel: pair_force_kernel. Algorithmic bytes per launch = (4 nbar + 44) N_local
            # (neighbour row + own position + count + head + force write)
            "roofline": roofline_pair(prof, pstats["mean_neighbors"], n_local, hbm, peak_src, world),
            "phases": {"force_ms_per_step": prof["force_ms"] / max(prof["force_launches"], 1),
                       "nlist_ms_per_build": prof["nlist_ms"] / max(prof["builds"], 1),
                       "nlist_ms_per_step": prof["nlist_ms"] / max(prof["steps"], 1),
                       "integrate_ms_per_step": prof["integrate_ms"] / max(prof["steps"], 1),
                       "profiled_steps": prof["steps"], "profiled_builds": prof["builds"],
                       "profiled_run_ms_per_step": prof_run_ms,
                       "note": "rank 0; nlist includes migration + ghost exchange, force includes the "
                               "per-step ghost refresh" if world > 1 else "single GPU"},
            "roofline_step": {"bound": "hbm", "achieved": bytes_step * value / world / 1e9,
                              "peak": hbm, "unit": "GB/s",
                              "frac": bytes_step * value / world / 1e9 / hbm,
                              "bytes_per_particle_step": bytes_step, "peak_source": peak_src},
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if os.environ.get("HB2_BENCH_CPU", "1") != "0" and world == 1:
            # bounded: ~10-30 s of CPU work (100 steps of the 64,000-particle liquid per core)
            _, line["cpu_baseline"], _ = cpu_arm(10, int(os.environ.get("HB2_CPU_STEPS", "100")))
        _emit(json.dumps(line))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    a = ap.parse_args()
    # The contract is ONE JSON line on stdout. Libraries write banners to file descriptor 1 from C
    # (NCCL prints its version there at communicator creation), so fd 1 is pointed at stderr for the
    # whole run and the JSON line goes to the saved descriptor.
    sys.stdout.flush()
    _json_fd = os.dup(1)
    os.dup2(2, 1)
    _emit = lambda text: os.write(_json_fd, (text + "\n").encode())  # noqa: E731
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
