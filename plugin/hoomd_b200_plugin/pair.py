"""hoomd_b200_plugin.pair.LJ: the operator a user swaps in for hoomd.md.pair.LJ (hoomd/md/pair/pair.py:163-240).
The neighbour list is the library's own (hoomd.md.nlist.Cell semantics: buffer, exclusions none), so the
constructor takes `buffer` instead of an nlist object."""
import os

import hoomd
from hoomd.md import force

from . import _b200_plugin

_DEFAULT_LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                            "hoomd_blue_b200", "libhoomd_b200.so")
_MODES = {"none": 0, "shift": 1, "xplor": 2}


class LJ(force.Force):
    """LJ pair force evaluated on the GPU by libhoomd_b200.so (cell list -> neighbour list -> pair kernel).

    Args:
        default_r_cut (float): cut-off of every type pair unless `r_cut[(a, b)]` says otherwise.
        buffer (float): neighbour-list buffer (hoomd.md.nlist.Cell(buffer)).
        mode (str): 'none', 'shift' or 'xplor'.
        gpu_id (int): CUDA ordinal.
        library (str): path of libhoomd_b200.so (default: the in-tree build).
    """

    def __init__(self, default_r_cut, buffer=0.4, mode="none", gpu_id=0, library=None):
        super().__init__()
        if mode not in _MODES:
            raise ValueError("mode must be 'none', 'shift' or 'xplor'")
        self._b200 = dict(r_cut_default=float(default_r_cut), buffer=float(buffer), mode=_MODES[mode],
                          gpu_id=int(gpu_id), library=library or _DEFAULT_LIB)
        self.params = {}
        self.r_cut = {}

    def _attach_hook(self):
        if not isinstance(self._simulation.device, hoomd.device.CPU):
            raise RuntimeError("hoomd_b200_plugin drives its own GPU state: attach it to a hoomd.device.CPU simulation")
        b = self._b200
        self._cpp_obj = _b200_plugin.PairLJB200(self._simulation.state._cpp_sys_def, b["library"], b["gpu_id"],
                                                b["buffer"], b["mode"])
        types = self._simulation.state.particle_types
        for i, a in enumerate(types):
            for j, c in enumerate(types):
                if j < i:
                    continue
                p = self.params.get((a, c), self.params.get((c, a)))
                if p is None:
                    raise ValueError(f"params[({a!r}, {c!r})] is not set")
                rc = self.r_cut.get((a, c), self.r_cut.get((c, a), b["r_cut_default"]))
                self._cpp_obj.setParams(i, j, float(p["epsilon"]), float(p["sigma"]), float(rc))

    @property
    def num_builds(self):
        return self._cpp_obj.getNumBuilds()

    @property
    def gpu_launches(self):
        return self._cpp_obj.getLaunchCount()
