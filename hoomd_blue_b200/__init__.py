"""hoomd_blue_b200 — B200-native short-range MD pair-force step behind HOOMD-blue's operator surface.

Only the path BASELINE.json:north_star names is implemented (cell binning -> neighbour list ->
PotentialPair<Evaluator> -> ConstantVolume half steps); the Python classes mirror the reference's
names, constructor arguments and properties for that path so that a script written for
``hoomd`` runs with ``import hoomd_blue_b200 as hoomd``. All compute happens in
libhoomd_b200.so (hand-written sm_100a CUDA behind the C ABI of include/hoomd_b200.h); nothing here
falls back to the CPU.
"""
from . import _lib  # noqa: F401
from . import device, filter, md  # noqa: F401
from .simulation import Simulation, Snapshot, State, Operations  # noqa: F401
from . import gsd_io  # noqa: F401
from . import write, trigger, logging, box  # noqa: F401
from .box import Box  # noqa: F401

__all__ = ["Simulation", "Snapshot", "device", "filter", "md", "gsd_io", "write", "trigger", "logging", "Box"]
version = "0.1.0+hoomd4.8.2-path"
