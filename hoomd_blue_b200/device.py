"""hoomd.device for this path (hoomd/device.py). Only GPU exists: there is no CPU implementation."""
from . import _lib
from .communicator import Communicator


class Device:
    pass


class GPU(Device):
    """hoomd.device.GPU(gpu_id=...) — selects the CUDA ordinal the state lives on."""

    def __init__(self, gpu_ids=None, gpu_id=None, **kwargs):
        if gpu_ids is not None and gpu_id is None:
            gpu_id = gpu_ids[0] if hasattr(gpu_ids, "__len__") else gpu_ids
        self.communicator = kwargs.get("communicator") or Communicator()
        if gpu_id is None and self.communicator.num_ranks > 1:
            import os
            gpu_id = int(os.environ.get("LOCAL_RANK", self.communicator.rank))
        self.gpu_id = int(gpu_id or 0)
        _lib.lib()  # fail loudly now if the CUDA library is not built


class CPU(Device):
    """Not available: this package accelerates the GPU path only."""

    def __init__(self, *args, **kwargs):
        raise RuntimeError("hoomd_blue_b200 has no CPU execution path; use device.GPU() "
                           "(the CPU implementation of this path is the reference itself)")


auto_select = GPU
