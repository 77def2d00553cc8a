"""Writes tests/golden/ref_two_frames.gsd with the REFERENCE's own GSD layer (hoomd/extern/gsd.c,
compiled in place into oracle/_ref/libgsd_ref.so by `make -C oracle ref`), laid out the way
hoomd/GSDDumpWriter.cc writes a hoomd-schema trajectory: frame 0 carries every chunk, frame 1 only the
dynamic ones (step, position, velocity, image), so a reader must fall back to frame 0 for the rest.

    python tests/golden/make_gsd_fixture.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.gsd_ref import RefGSD  # noqa: E402


def main():
    rng = np.random.default_rng(12)
    n = 50
    out = os.path.join(ROOT, "tests", "golden", "ref_two_frames.gsd")
    if os.path.exists(out):
        os.remove(out)
    g = RefGSD.create(out)
    pos0 = (rng.random((n, 3)) - 0.5).astype(np.float32) * np.float32(8)
    vel0 = rng.normal(size=(n, 3)).astype(np.float32)
    img0 = rng.integers(-2, 3, size=(n, 3)).astype(np.int32)
    typeid = (np.arange(n) % 3 == 0).astype(np.uint32)
    mass = rng.uniform(1, 2, n).astype(np.float32)
    types = np.zeros((2, 5), dtype=np.int8)
    types[0, :1] = np.frombuffer(b"A", dtype=np.int8)
    types[1, :4] = np.frombuffer(b"Poly", dtype=np.int8)
    g.write("configuration/step", np.array([[1000]], dtype=np.uint64))
    g.write("configuration/dimensions", np.array([[3]], dtype=np.uint8))
    g.write("configuration/box", np.array([8, 9, 10, 0, 0, 0], dtype=np.float32).reshape(6, 1))
    g.write("particles/N", np.array([[n]], dtype=np.uint32))
    g.write("particles/types", types)
    g.write("particles/typeid", typeid.reshape(n, 1))
    g.write("particles/mass", mass.reshape(n, 1))
    g.write("particles/position", pos0)
    g.write("particles/velocity", vel0)
    g.write("particles/image", img0)
    g.end_frame()
    g.write("configuration/step", np.array([[2000]], dtype=np.uint64))
    g.write("particles/position", (pos0 * np.float32(0.5)).astype(np.float32))
    g.write("particles/velocity", (-vel0).astype(np.float32))
    g.write("particles/image", (img0 + 1).astype(np.int32))
    g.end_frame()
    g.close()
    np.savez(os.path.join(ROOT, "tests", "golden", "ref_two_frames_expect.npz"), pos0=pos0, vel0=vel0,
             img0=img0, typeid=typeid, mass=mass)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
