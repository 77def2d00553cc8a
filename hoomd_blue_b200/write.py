"""hoomd.write for this path: GSD.write, the one-shot frame dump (hoomd/write/gsd.py:215-253). The
periodic trajectory writer (triggers, dynamic property selection, logger chunks) is outside the path."""
from . import gsd_io


class GSD:

    @staticmethod
    def write(state, filename, filter=None, mode="wb", logger=None):
        if mode not in ("wb", "xb"):
            raise ValueError("mode must be 'wb' or 'xb'")
        if logger is not None:
            raise NotImplementedError("logged quantities in GSD files are outside this path")
        import os
        if mode == "xb" and os.path.exists(filename):
            raise FileExistsError(filename)
        snap = state.get_snapshot()
        if state._simulation.device.communicator.rank == 0:
            gsd_io.write_frame(filename, snap, state._simulation.timestep)
