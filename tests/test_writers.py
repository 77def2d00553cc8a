"""Triggers, the logger, the Table writer and the appending GSD trajectory writer (SURVEY 8f rank 1:
hoomd/write/gsd.py, hoomd/GSDDumpWriter.cc, hoomd/write/table.py, hoomd/trigger.py). The files are read
back by this package's reader and — where oracle/_ref/libgsd_ref.so exists — by the reference's own
gsd.c."""
import io
import os

import numpy as np
import pytest

from tests import gsd_ref
from tests.test_gsd_io import _snapshot


def test_triggers():
    from hoomd_blue_b200 import trigger as T
    p = T.Periodic(100, phase=30)
    assert [t for t in range(0, 400) if p(t)] == [30, 130, 230, 330]
    assert p.next_after(30) == 130 and p.next_after(29) == 30 and p.next_after(131) == 230
    assert T.as_trigger(7).period == 7
    assert T.On(5)(5) and not T.On(5)(6) and T.On(5).next_after(5) is None and T.On(5).next_after(1) == 5
    assert T.Before(3)(2) and not T.Before(3)(3) and T.After(3)(4) and not T.After(3)(3)
    assert T.After(10).next_after(2) == 11 and T.Before(10).next_after(9) is None
    both = T.And([T.Periodic(6), T.Periodic(4)])
    assert [t for t in range(0, 40) if both(t)] == [0, 12, 24, 36]
    assert both.next_after(0) == 12 and both.next_after(12) == 24 and both.next_after(13) == 24
    either = T.Or([T.Periodic(6), T.On(7)])
    assert [t for t in range(0, 20) if either(t)] == [0, 6, 7, 12, 18] and either.next_after(6) == 7
    late = T.And([T.Periodic(5), T.After(12)])
    assert late.next_after(0) == 15 and not late(10)
    assert T.Not(T.Periodic(2)).next_after(2) == 3 and T.Not(T.Periodic(2))(3)
    assert T.And([T.On(3), T.On(4)]).next_after(0) is None
    with pytest.raises(ValueError):
        T.Periodic(0)
    with pytest.raises(TypeError):
        T.as_trigger("x")


def test_box():
    import hoomd_blue_b200 as hoomd
    b = hoomd.Box.cube(5)
    assert list(b) == [5, 5, 5, 0, 0, 0] and b.volume == 125 and not b.is2D and b.dimensions == 3
    assert hoomd.Box.square(4).is2D and hoomd.Box.square(4).volume == 16
    assert hoomd.Box.from_box([1, 2, 3]) == hoomd.Box(1, 2, 3)
    assert hoomd.Box.from_box(dict(Lx=1, Ly=2, Lz=3, xy=0.5)).xy == 0.5
    assert hoomd.Box.from_box(hoomd.Box(1, 2, 3, 0.1, 0.2, 0.3)).tilts.tolist() == [0.1, 0.2, 0.3]
    b.L = 7
    assert b.L.tolist() == [7, 7, 7]
    s = hoomd.Snapshot()
    s.configuration.box = hoomd.Box(3, 4, 5)
    assert list(s.configuration.box) == [3, 4, 5, 0, 0, 0]
    with pytest.raises(ValueError):
        hoomd.Box.from_box([1, 2])


def test_snapshot_replicate_and_wrap():
    """hoomd.Snapshot.replicate / wrap (hoomd/ParticleData.cc:3742-3810, hoomd/BondedGroupData.cc:1366-1405)."""
    from hoomd_blue_b200 import Snapshot
    s = _snapshot(30, 5)
    L = np.array(s.configuration.box[:3])
    s.bonds.N, s.bonds.types = 3, ["b"]
    s.bonds.group[:] = [(0, 1), (1, 2), (28, 29)]
    s.angles.N, s.angles.types = 1, ["a"]
    s.angles.group[:] = [(0, 1, 2)]
    pos0 = s.particles.position.astype(np.float64) + s.particles.image * L
    vel0, tid0 = s.particles.velocity.copy(), s.particles.typeid.copy()
    s.replicate(2, 3, 1)
    assert s.particles.N == 180 and s.configuration.box[:3] == list(L * [2, 3, 1])
    Ln = L * [2, 3, 1]
    un = s.particles.position.astype(np.float64) + s.particles.image * Ln
    j = 0
    for l in range(2):
        for m in range(3):
            blk = slice(30 * j, 30 * (j + 1))
            # replica (l, m, 0): the unwrapped original, moved to the replica's cell of the new box
            want = pos0 + L / 2 - Ln / 2 + L * [l, m, 0]
            assert np.allclose(un[blk], want, atol=2e-5)
            assert np.array_equal(s.particles.velocity[blk], vel0) and np.array_equal(s.particles.typeid[blk], tid0)
            assert np.array_equal(s.bonds.group[3 * j:3 * j + 3], np.array([(0, 1), (1, 2), (28, 29)]) + 30 * j)
            assert np.array_equal(s.angles.group[j], np.array([0, 1, 2]) + 30 * j)
            j += 1
    assert s.bonds.N == 18 and s.angles.N == 6 and s.bonds.types == ["b"]
    assert np.all(s.particles.position < (Ln / 2).astype(np.float32))
    assert np.all(s.particles.position >= -(Ln / 2).astype(np.float32))
    # wrap
    t = Snapshot.from_arrays([4, 4, 4, 0, 0, 0], np.array([[2.5, -2.0, 0.0], [-6.1, 1.9, 7.9]], dtype=np.float32))
    t.wrap()
    assert np.allclose(t.particles.position, [[-1.5, -2.0, 0.0], [1.9, 1.9, -0.1]], atol=1e-6)
    assert np.array_equal(t.particles.image, [[1, 0, 0], [-2, 0, 2]])
    with pytest.raises(ValueError):
        t.replicate(0, 1, 1)


def test_gsd_append_writer(tmp_path):
    """300 appended frames (the 128-entry index and the name list both have to move), dynamic
    categories, log chunks; frame 0 fallback on read."""
    from hoomd_blue_b200 import gsd_io
    path = str(tmp_path / "traj.gsd")
    base = _snapshot(40, 1)
    w = gsd_io.GSDWriter(path, "wb")
    frames = []
    for k in range(300):
        s = _snapshot(40, 1)
        s.particles.position[:] = base.particles.position + np.float32(0.001 * k)
        s.particles.velocity[:] = base.particles.velocity * np.float32(k)
        dyn = {"property"} if k % 2 else {"property", "momentum"}
        log = {f"md/q{k % 70}/value": 0.5 * k, "Simulation/timestep": 10 * k}
        w.append(gsd_io.frame_chunks(s, 10 * k, dynamic=dyn, first=(k == 0), log=log))
        frames.append(s)
    g = gsd_io.GSDFile(path)
    assert g.nframes == 300 and len(g.names) > 70
    for k in (0, 1, 2, 127, 128, 299):
        s, step = gsd_io.read_frame(path, k)
        assert step == 10 * k
        assert np.array_equal(s.particles.position, frames[k].particles.position)
        want_vel = frames[k].particles.velocity if k % 2 == 0 else frames[0].particles.velocity
        assert np.array_equal(s.particles.velocity, want_vel)
        assert np.array_equal(s.particles.mass, base.particles.mass) and s.particles.types == base.particles.types
        assert g.chunk(k, f"log/md/q{k % 70}/value")[0, 0] == 0.5 * k
        assert g.chunk(k, "log/Simulation/timestep").dtype == np.int64
    # append to the existing file
    w2 = gsd_io.GSDWriter(path, "ab")
    assert w2.nframes == 300
    w2.append(gsd_io.frame_chunks(frames[5], 7777, dynamic={"property"}, first=False))
    assert gsd_io.read_frame(path, -1)[1] == 7777 and gsd_io.GSDFile(path).nframes == 301
    with pytest.raises(FileExistsError):
        gsd_io.GSDWriter(path, "xb")
    if gsd_ref.available():
        r = gsd_ref.RefGSD.open(path)
        assert r.nframes() == 301
        for k in (0, 129, 300):
            assert r.read(k, "configuration/step")[0, 0] == (7777 if k == 300 else 10 * k)
        assert np.array_equal(r.read(299, "particles/position"), frames[299].particles.position)
        assert r.read(299, "particles/velocity") is None  # odd frame: not dynamic
        assert r.read(298, "log/md/q18/value")[0, 0] == 149.0
        r.close()


class _Thermo:
    _loggables = {"kinetic_energy": "scalar", "pressure": "scalar", "tensor": "sequence"}
    kinetic_energy = 1.25
    pressure = -3.0
    tensor = (1.0, 2.0)


_Thermo.__module__ = "hoomd_blue_b200.md.compute"


def test_logger_and_table():
    from hoomd_blue_b200 import logging, write, device, Simulation
    lg = logging.Logger(categories=["scalar"])
    th = _Thermo()
    added = lg.add(th)
    assert added == [("md", "compute", "_Thermo", "kinetic_energy"), ("md", "compute", "_Thermo", "pressure")]
    lg[("custom", "answer")] = (lambda: 42, "scalar")
    with pytest.raises(ValueError):
        lg.add(th, quantities=["nope"])
    out = io.StringIO()

    class _Sim:
        class device:
            class communicator:
                rank = 0
    tb = write.Table(trigger=10, logger=lg, output=out, max_header_len=2)
    tb._write(_Sim)
    th.kinetic_energy = 2.5
    tb._write(_Sim)
    nested = lg.log()
    assert nested["md"]["compute"]["_Thermo"]["pressure"] == (-3.0, "scalar")
    assert nested["custom"]["answer"] == (42, "scalar")
    lg2 = logging.Logger()
    lg2.add(th)
    assert lg2.log()["md"]["compute"]["_Thermo"]["tensor"] == ((1.0, 2.0), "sequence")
    lg2.remove(th, quantities=["tensor"])
    assert "tensor" not in lg2.log()["md"]["compute"]["_Thermo"]
    lines = out.getvalue().splitlines()
    assert lines[0] == "_Thermo.kinetic_energy _Thermo.pressure custom.answer"
    assert lines[1] == "1.25 -3 42" and lines[2] == "2.5 -3 42"


@pytest.mark.gpu
def test_periodic_gsd_writer_in_run(tmp_path):
    """A run with writers is the same trajectory as one without; frames land on the triggered steps
    with the logged thermodynamics of those steps."""
    import hoomd_blue_b200 as hoomd
    from tests.test_gpu_parity import random_liquid

    def make():
        pos, vel, L = random_liquid(12, 0.8, seed=2)
        sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=4)
        sim.create_state_from_snapshot(hoomd.Snapshot.from_arrays(list(L) + [0, 0, 0], pos, vel))
        cell = hoomd.md.nlist.Cell(buffer=0.4)
        lj = hoomd.md.pair.LJ(nlist=cell, default_r_cut=2.5)
        lj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
        cv = hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())
        sim.operations.integrator = hoomd.md.Integrator(dt=0.002, methods=[cv], forces=[lj])
        return sim, lj

    ref, _ = make()
    ref.run(230)
    sim, lj = make()
    thermo = hoomd.md.compute.ThermodynamicQuantities(filter=hoomd.filter.All())
    sim.operations += thermo
    lg = hoomd.logging.Logger(categories=["scalar"])
    lg.add(thermo, quantities=["kinetic_energy", "potential_energy"])
    lg.add(sim, quantities=["timestep"])
    path = str(tmp_path / "traj.gsd")
    sim.operations.writers.append(hoomd.write.GSD(trigger=hoomd.trigger.Periodic(50), filename=path,
                                                  mode="wb", dynamic=["property", "momentum"], logger=lg))
    table = io.StringIO()
    sim.operations += hoomd.write.Table(trigger=100, logger=lg, output=table)
    sim.run(230, write_at_start=True)
    assert sim.timestep == 230
    a, b = sim.state.get_snapshot(), ref.state.get_snapshot()
    assert np.array_equal(a.particles.position, b.particles.position)
    assert np.array_equal(a.particles.velocity, b.particles.velocity)
    g = hoomd.gsd_io.GSDFile(path)
    assert g.nframes == 5
    assert [int(g.chunk(k, "configuration/step")[0, 0]) for k in range(5)] == [0, 50, 100, 150, 200]
    assert [int(g.chunk(k, "log/Simulation/timestep")[0, 0]) for k in range(5)] == [0, 50, 100, 150, 200]
    e = [g.chunk(k, "log/md/compute/ThermodynamicQuantities/kinetic_energy")[0, 0]
         + g.chunk(k, "log/md/compute/ThermodynamicQuantities/potential_energy")[0, 0] for k in range(5)]
    assert np.ptp(e) < 2e-2 * abs(e[0])  # NVE from a jittered lattice: the logged total energy is conserved
    s4, step4 = hoomd.gsd_io.read_frame(path, 4)
    chk, _ = make()
    chk.run(200)
    c = chk.state.get_snapshot()
    assert step4 == 200 and np.array_equal(s4.particles.position, c.particles.position)
    ke_chk, pe_chk, _ = chk.operations.integrator._thermo()
    assert g.chunk(4, "log/md/compute/ThermodynamicQuantities/kinetic_energy")[0, 0] == ke_chk
    assert g.chunk(4, "log/md/compute/ThermodynamicQuantities/potential_energy")[0, 0] == pe_chk
    assert np.array_equal(s4.particles.image, c.particles.image)
    rows = table.getvalue().splitlines()
    assert len(rows) == 1 + 3 and rows[0].split()[-1] == "Simulation.timestep"
    assert [r.split()[-1] for r in rows[1:]] == ["0", "100", "200"]
    # restart from the trajectory's last frame
    sim2 = hoomd.Simulation(device=hoomd.device.GPU(), seed=4)
    sim2.create_state_from_gsd(path)
    assert sim2.timestep == 200 and sim2.state.N_particles == 12 ** 3


@pytest.mark.gpu
def test_molecular_run_with_writers_equals_plain_run(tmp_path):
    """The composition INTEGRATION.md shows (pair + FENE bonds + angles, bond / 1-3 exclusions, the
    Langevin method, a trajectory writer with logged thermodynamics): interrupting the run for the writer
    and its logger must not change the trajectory (the BD force carried into the next step one survives
    the getters' recomputation of the pair force)."""
    import hoomd_blue_b200 as hoomd
    from tests.test_bonds import lattice_chains

    def make():
        pos, group, L = lattice_chains(16, 8, 8, 8, 1.0, seed=9, jitter=0.05)
        n = len(pos)
        bonded = set(map(tuple, group.tolist()))
        angles = np.array([(i - 1, i, i + 1) for i in range(1, n - 1)
                           if (i - 1, i) in bonded and (i, i + 1) in bonded], dtype=np.uint32)
        snap = hoomd.Snapshot.from_arrays(list(L) + [0, 0, 0], pos, np.zeros_like(pos))
        snap.bonds.N, snap.bonds.types = len(group), ["polymer"]
        snap.bonds.group[:] = group
        snap.angles.N, snap.angles.types = len(angles), ["backbone"]
        snap.angles.group[:] = angles
        sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=12)
        sim.create_state_from_snapshot(snap)
        cell = hoomd.md.nlist.Cell(buffer=0.3, exclusions=("bond", "1-3"))
        lj = hoomd.md.pair.LJ(nlist=cell, default_r_cut=2.0 ** (1.0 / 6.0), mode="shift")
        lj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
        fene = hoomd.md.bond.FENEWCA()
        fene.params["polymer"] = dict(k=30.0, r0=1.5, epsilon=1.0, sigma=1.0, delta=0.0)
        bend = hoomd.md.angle.Harmonic()
        bend.params["backbone"] = dict(k=5.0, t0=2.6)
        lv = hoomd.md.methods.Langevin(filter=hoomd.filter.All(), kT=1.0)
        sim.operations.integrator = hoomd.md.Integrator(dt=0.002, methods=[lv], forces=[lj, fene, bend])
        return sim

    plain = make()
    plain.run(300)
    sim = make()
    thermo = hoomd.md.compute.ThermodynamicQuantities(filter=hoomd.filter.All())
    sim.operations += thermo
    lg = hoomd.logging.Logger(categories=["scalar"])
    lg.add(thermo, quantities=["kinetic_temperature", "potential_energy"])
    path = str(tmp_path / "mol.gsd")
    sim.operations.writers.append(hoomd.write.GSD(trigger=hoomd.trigger.Periodic(70), filename=path, mode="wb",
                                                  dynamic=["property", "momentum"], logger=lg))
    sim.run(300)
    a, b = sim.state.get_snapshot(), plain.state.get_snapshot()
    assert np.array_equal(a.particles.position, b.particles.position)
    assert np.array_equal(a.particles.velocity, b.particles.velocity)
    g = hoomd.gsd_io.GSDFile(path)
    assert g.nframes == 4 and int(g.chunk(0, "bonds/N")[0, 0]) == a.bonds.N
    temps = [g.chunk(k, "log/md/compute/ThermodynamicQuantities/kinetic_temperature")[0, 0] for k in range(4)]
    assert 0.2 < temps[-1] < 2.0
