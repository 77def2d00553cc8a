"""GSD (hoomd schema) reader/writer of the path's input format against the reference's own GSD layer:
a committed fixture that hoomd/extern/gsd.c wrote (tests/golden/make_gsd_fixture.py), and — where
oracle/_ref/libgsd_ref.so exists (built in place from /root/reference by `make -C oracle ref`; it
travels to the GPU box) — live cross-checks in both directions."""
import os

import numpy as np
import pytest

from tests import gsd_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "ref_two_frames.gsd")
EXP = os.path.join(ROOT, "tests", "golden", "ref_two_frames_expect.npz")


def test_read_reference_written_fixture():
    from hoomd_blue_b200 import gsd_io
    ex = np.load(EXP)
    g = gsd_io.GSDFile(FIX)
    assert (g.schema, g.nframes, g.gsd_version >> 16) == ("hoomd", 2, 2)
    s0, step0 = gsd_io.read_frame(FIX, 0)
    assert step0 == 1000 and s0.particles.types == ["A", "Poly"] and s0.particles.N == 50
    assert s0.configuration.box == [8.0, 9.0, 10.0, 0.0, 0.0, 0.0]
    assert np.array_equal(s0.particles.position, ex["pos0"])
    assert np.array_equal(s0.particles.velocity, ex["vel0"])
    assert np.array_equal(s0.particles.image, ex["img0"])
    assert np.array_equal(s0.particles.typeid, ex["typeid"])
    assert np.array_equal(s0.particles.mass, ex["mass"])
    # frame 1 holds only the dynamic chunks: the rest falls back to frame 0 (hoomd/GSDReader.cc:100-155)
    for fr in (1, -1):
        s1, step1 = gsd_io.read_frame(FIX, fr)
        assert step1 == 2000 and s1.particles.types == ["A", "Poly"]
        assert np.array_equal(s1.particles.position, ex["pos0"] * np.float32(0.5))
        assert np.array_equal(s1.particles.velocity, -ex["vel0"])
        assert np.array_equal(s1.particles.image, ex["img0"] + 1)
        assert np.array_equal(s1.particles.mass, ex["mass"])
        assert np.array_equal(s1.particles.typeid, ex["typeid"])
    with pytest.raises(RuntimeError, match="only has 2 frames"):
        gsd_io.read_frame(FIX, 2)


def _snapshot(n, seed):
    from hoomd_blue_b200 import Snapshot
    rng = np.random.default_rng(seed)
    s = Snapshot.from_arrays([7, 8, 9, 0, 0, 0], (rng.random((n, 3)) - 0.5).astype(np.float32) * 6,
                             rng.normal(size=(n, 3)).astype(np.float32), rng.integers(0, 3, n),
                             ("A", "B", "Long_type_name"), rng.integers(-5, 6, (n, 3)),
                             rng.uniform(0.5, 3, n).astype(np.float32))
    return s


def _same(a, b):
    pa, pb = a.particles, b.particles
    return (pa.types == pb.types and a.configuration.box == b.configuration.box
            and all(np.array_equal(getattr(pa, k), getattr(pb, k))
                    for k in ("position", "velocity", "image", "typeid", "mass")))


def test_write_read_roundtrip(tmp_path):
    from hoomd_blue_b200 import gsd_io
    a, b = _snapshot(300, 1), _snapshot(300, 2)
    path = str(tmp_path / "t.gsd")
    gsd_io.write_frames(path, [(a, 5), (b, 10 ** 12)])
    ra, sa = gsd_io.read_frame(path, 0)
    rb, sb = gsd_io.read_frame(path, -1)
    assert (sa, sb) == (5, 10 ** 12) and _same(ra, a) and _same(rb, b)
    with open(path, "r+b") as f:
        f.write(b"\0" * 8)
    with pytest.raises(RuntimeError, match="not a GSD file"):
        gsd_io.read_frame(path, 0)


@pytest.mark.skipif(not gsd_ref.available(), reason="oracle/_ref/libgsd_ref.so not built")
def test_against_reference_gsd_layer(tmp_path):
    from hoomd_blue_b200 import gsd_io
    # our writer -> the reference's reader
    a, b = _snapshot(257, 3), _snapshot(257, 4)
    path = str(tmp_path / "ours.gsd")
    gsd_io.write_frames(path, [(a, 7), (b, 8)])
    g = gsd_ref.RefGSD.open(path)
    assert g.nframes() == 2
    for fr, (s, step) in enumerate([(a, 7), (b, 8)]):
        assert int(g.read(fr, "configuration/step")[0, 0]) == step
        assert np.array_equal(g.read(fr, "particles/position"), s.particles.position)
        assert np.array_equal(g.read(fr, "particles/velocity"), s.particles.velocity)
        assert np.array_equal(g.read(fr, "particles/image"), s.particles.image)
        assert np.array_equal(g.read(fr, "particles/typeid").ravel(), s.particles.typeid)
        assert np.array_equal(g.read(fr, "particles/mass").ravel(), s.particles.mass)
        assert np.array_equal(g.read(fr, "configuration/box").ravel(),
                              np.asarray(s.configuration.box, dtype=np.float32))
        t = g.read(fr, "particles/types")
        assert [bytes(r.astype(np.uint8)).split(b"\0")[0].decode() for r in t] == s.particles.types
    assert g.read(0, "particles/charge") is None
    g.close()
    # the reference's writer -> our reader (a frame without velocities: the default stays)
    path = str(tmp_path / "ref.gsd")
    w = gsd_ref.RefGSD.create(path)
    w.write("configuration/step", np.array([[3]], dtype=np.uint64))
    w.write("configuration/box", np.array([4, 4, 4, 0, 0, 0], dtype=np.float32).reshape(6, 1))
    w.write("particles/N", np.array([[257]], dtype=np.uint32))
    w.write("particles/position", a.particles.position)
    w.end_frame()
    w.close()
    s, step = gsd_io.read_frame(path, 0)
    assert step == 3 and s.particles.N == 257 and s.particles.types == ["A"]
    assert np.array_equal(s.particles.position, a.particles.position)
    assert not s.particles.velocity.any() and (s.particles.mass == 1).all()


@pytest.mark.gpu
def test_create_state_from_gsd():
    import hoomd_blue_b200 as hoomd
    ex = np.load(EXP)
    sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=1)
    sim.create_state_from_gsd(FIX, frame=-1)
    assert sim.timestep == 2000
    snap = sim.state.get_snapshot()
    assert snap.particles.types == ["A", "Poly"]
    assert np.array_equal(snap.particles.position, ex["pos0"] * np.float32(0.5))
    assert np.array_equal(snap.particles.typeid, ex["typeid"])
    assert np.array_equal(snap.particles.image, ex["img0"] + 1)
    # hoomd.write.GSD.write(state, filename) and back
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "dump.gsd")
        hoomd.write.GSD.write(sim.state, out)
        back, step = hoomd.gsd_io.read_frame(out)
        assert step == 2000 and _same(back, snap)
