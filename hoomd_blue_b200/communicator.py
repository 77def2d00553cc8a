"""hoomd.communicator for this path (hoomd/communicator.py): one process per GPU.

The reference partitions ranks with MPI; here the process group is torch.distributed's (NCCL on GPU
boxes, gloo in CPU tests) and is used ONLY for plumbing: bootstrapping the engine's own NCCL
communicator (the 128-byte unique id) and gathering per-particle results by tag for the Python
properties. Ghost exchange, migration and the per-step reductions run inside libhoomd_b200.so on the
engine's stream (csrc/comm.cu)."""
import numpy as np


def _dist():
    try:
        import torch.distributed as dist
    except Exception:  # pragma: no cover - torch is part of the image
        return None
    return dist if dist.is_available() and dist.is_initialized() else None


class Communicator:
    """hoomd.communicator.Communicator: rank / num_ranks of the (single) partition."""

    def __init__(self):
        d = _dist()
        self.rank = d.get_rank() if d else 0
        self.num_ranks = d.get_world_size() if d else 1

    def barrier(self):
        d = _dist()
        if d:
            d.barrier()


def broadcast_object(obj, src=0):
    d = _dist()
    if d is None:
        return obj
    box = [obj]
    d.broadcast_object_list(box, src=src)
    return box[0]


def all_gather_arrays(arr):
    """List (one entry per rank) of the numpy arrays each rank passed."""
    d = _dist()
    if d is None:
        return [arr]
    out = [None] * d.get_world_size()
    d.all_gather_object(out, arr)
    return out


def gather_by_tag(n_global, tags, values):
    """Scatter every rank's (tags, values) rows into one tag-ordered array (all ranks get it)."""
    parts = all_gather_arrays((np.asarray(tags), np.asarray(values)))
    shape = (int(n_global),) + tuple(values.shape[1:])
    out = np.zeros(shape, dtype=values.dtype)
    for t, v in parts:
        out[t] = v
    return out
