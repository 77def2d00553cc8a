"""hoomd.md.nlist for this path (hoomd/md/nlist.py:87-418): Cell."""
import ctypes as C

import numpy as np

from .. import _lib
from .. import communicator as _comm
from .._lib import check, ptr
from .typeparam import TypePairDict


class _DeviceArray:
    """A view of device memory through __cuda_array_interface__ (version 3): CuPy, PyTorch
    (torch.as_tensor) and Numba wrap it without a copy, like the reference's HOOMDGPUArray. The list is
    the library's: treat it as read-only (the interface's read-only flag stays off because PyTorch refuses
    to wrap arrays that set it)."""

    def __init__(self, pointer, shape, typestr="<i4", read_only=False):  # int32 view of the uint32 data
        self.__cuda_array_interface__ = dict(shape=tuple(int(s) for s in shape), typestr=typestr,
                                             data=(int(pointer), bool(read_only)), version=3, strides=None)
        self.shape = tuple(int(s) for s in shape)


class LocalNeighborListData:
    """hoomd.md.data.LocalNeighborListData{CPU,GPU} for a full list (hoomd/md/nlist.py:154-268,
    hoomd/md/NeighborList.h:662-722): `head_list[i]` is the offset of particle i's row in `nlist`,
    `n_neigh[i]` its length; indices are local (rank-local, cell-ordered) particle indices — the
    state's local snapshot gives the tags. `half_nlist` is False. Use as a context manager."""

    def __init__(self, nl, on_device):
        L = _lib.lib()
        d_nn, d_nl = C.c_void_p(), C.c_void_p()
        pitch, rows = C.c_uint32(0), C.c_uint32(0)
        check(L.hb2_nlist_device_arrays(nl._h, C.byref(d_nn), C.byref(d_nl), C.byref(pitch), C.byref(rows)))
        self.half_nlist = False
        n, p = rows.value, pitch.value
        if on_device:
            self.n_neigh = _DeviceArray(d_nn.value, (n,))
            self.nlist = _DeviceArray(d_nl.value, (n * p,))
            self._head = None
        else:
            import torch  # device memory plumbing only
            dev = torch.device("cuda", nl._sim.device.gpu_id)
            self.n_neigh = torch.as_tensor(_DeviceArray(d_nn.value, (n,)), device=dev).cpu().numpy().view(np.uint32)
            self.nlist = torch.as_tensor(_DeviceArray(d_nl.value, (n * p,)), device=dev).cpu().numpy().view(np.uint32)
            self.n_neigh.flags.writeable = self.nlist.flags.writeable = False
        self.head_list = np.arange(n, dtype=np.uint64) * np.uint64(p)
        self.head_list.flags.writeable = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def _nonneg(v):
    v = float(v)
    if v < 0:
        raise ValueError("r_cut must be non-negative")
    return v


class NeighborList:
    """Base class (hoomd/md/nlist.py:87-323)."""

    def __init__(self, buffer, exclusions, rebuild_check_delay, check_dist, mesh, default_r_cut):
        unknown = set(exclusions or ()) - {"bond", "angle", "1-3"}
        if unknown:
            raise NotImplementedError(f"exclusions {sorted(unknown)} are outside this path "
                                      "(supported: 'bond', 'angle', '1-3')")
        if mesh is not None:
            raise NotImplementedError("mesh neighbour lists are outside this path")
        self.buffer = float(buffer)
        self.exclusions = tuple(exclusions or ())
        self.rebuild_check_delay = int(rebuild_check_delay)
        self.check_dist = bool(check_dist)
        self.default_r_cut = float(default_r_cut)
        # hoomd/md/nlist.py:127-131: the list's own r_cut per unordered type pair; the effective cut-off
        # is the maximum over this matrix and the pair forces attached to the list
        self.r_cut = TypePairDict("r_cut", lambda v: _nonneg(v), float(default_r_cut))
        self.r_cut._on_change = self._sync_r_cut
        self._h = None
        self._sim = None

    @property
    def _attached(self):
        return self._h is not None

    def _attach(self, sim):
        if self._h is not None:
            if self._sim is not sim:
                raise RuntimeError("nlist is attached to another simulation")
            return
        h = C.c_void_p()
        check(_lib.lib().hb2_nlist_create(C.byref(h), sim.state._h, self.buffer,
                                          self.rebuild_check_delay, int(self.check_dist)))
        self._h = h
        self._sim = sim
        self._sync_r_cut()
        check(_lib.lib().hb2_nlist_set_exclusions(h, int("bond" in self.exclusions)))
        extra = self._extra_exclusion_pairs(sim.state)
        if len(extra):
            check(_lib.lib().hb2_nlist_set_exclusion_pairs(h, len(extra), ptr(extra, C.c_uint32)))

    def _sync_r_cut(self):
        if self._h is None:
            return
        types = self._sim.state.particle_types
        for i, a in enumerate(types):
            for j, b in enumerate(types):
                if j >= i:
                    check(_lib.lib().hb2_nlist_set_r_cut(self._h, i, j, self.r_cut[(a, b)]))

    @property
    def cpu_local_nlist_arrays(self):
        """Host copies of the list of the last build (hoomd/md/nlist.py:154-209)."""
        self._need()
        return LocalNeighborListData(self, on_device=False)

    @property
    def gpu_local_nlist_arrays(self):
        """Zero-copy device views (__cuda_array_interface__) of the list of the last build
        (hoomd/md/nlist.py:211-268); valid until the next rebuild."""
        self._need()
        return LocalNeighborListData(self, on_device=True)

    def _extra_exclusion_pairs(self, state):
        """Pairs (by tag) that 'angle' and '1-3' add (hoomd/md/NeighborList.cc:881-909: the two end
        particles of every angle; :1048-1120: every two particles bonded to a common one)."""
        pairs = []
        if "angle" in self.exclusions and len(getattr(state, "_angle_group", ())):
            g = np.asarray(state._angle_group)
            pairs.append(g[:, [0, 2]])
        if "1-3" in self.exclusions and len(getattr(state, "_bond_group", ())):
            g = np.asarray(state._bond_group, dtype=np.int64)
            both = np.concatenate([g, g[:, ::-1]])          # (centre, partner) for both directions
            both = both[np.argsort(both[:, 0], kind="stable")]
            # two entries k apart with the same centre are two partners of that centre
            k = 1
            while k < len(both):
                same = both[:-k, 0] == both[k:, 0]
                if not same.any():
                    break
                pairs.append(np.stack([both[:-k, 1][same], both[k:, 1][same]], axis=1))
                k += 1
        if not pairs:
            return np.zeros((0, 2), dtype=np.uint32)
        allp = np.concatenate(pairs).astype(np.int64)
        allp = allp[allp[:, 0] != allp[:, 1]]
        allp = np.unique(np.sort(allp, axis=1), axis=0)   # addExclusion ignores duplicates
        return np.ascontiguousarray(allp, dtype=np.uint32)

    def _detach(self):
        if self._h is not None:
            _lib.lib().hb2_nlist_destroy(self._h)
            self._h = None
            self._sim = None

    def _need(self):
        if self._h is None:
            raise RuntimeError("nlist is not attached (run the simulation first)")  # DataAccessError

    @property
    def local_pair_list(self):
        """Tag pairs listed on this rank; a pair that straddles two domains is reported by the
        owner of its smaller tag only, so the ranks' lists are disjoint
        (hoomd/md/nlist.py:270-291)."""
        self._need()
        L = _lib.lib()
        n = C.c_uint64(0)
        check(L.hb2_nlist_pair_count(self._h, C.byref(n)))
        out = np.zeros((max(n.value, 1), 2), dtype=np.uint32)
        check(L.hb2_nlist_pair_list(self._h, ptr(out, C.c_uint32), out.shape[0], C.byref(n)))
        return out[:n.value]

    @property
    def pair_list(self):
        """(n, 2) uint32 tag pairs, each unordered pair once (hoomd/md/nlist.py:293-305). With a
        domain decomposition the ranks' lists are concatenated (the reference gathers to rank 0;
        here every rank receives the list)."""
        local = self.local_pair_list
        if self._sim.device.communicator.num_ranks > 1:
            return np.concatenate(_comm.all_gather_arrays(local), axis=0)
        return local

    @property
    def num_builds(self):
        self._need()
        return int(_lib.lib().hb2_nlist_num_builds(self._h))

    @property
    def shortest_rebuild(self):
        self._need()
        return int(_lib.lib().hb2_nlist_shortest_rebuild(self._h))

    def _stats(self):
        self._need()
        mean = C.c_double(0)
        mx = C.c_uint32(0)
        pitch = C.c_uint32(0)
        check(_lib.lib().hb2_nlist_stats(self._h, C.byref(mean), C.byref(mx), C.byref(pitch)))
        return dict(mean_neighbors=mean.value, max_neighbors=mx.value, pitch=pitch.value)


class Cell(NeighborList):
    """hoomd.md.nlist.Cell (hoomd/md/nlist.py:326-418). `deterministic` is accepted and always
    honoured: binning is a key sort and rows are written in ascending particle index."""

    def __init__(self, buffer, exclusions=("bond",), rebuild_check_delay=1, check_dist=True,
                 deterministic=False, mesh=None, default_r_cut=0.0):
        super().__init__(buffer, exclusions, rebuild_check_delay, check_dist, mesh, default_r_cut)
        self.deterministic = bool(deterministic)

    @property
    def dimensions(self):
        self._need()
        d = (C.c_uint32 * 3)()
        check(_lib.lib().hb2_nlist_dimensions(self._h, d))
        return tuple(int(x) for x in d)

    @property
    def allocated_particles_per_cell(self):
        # the cell list is an offset table into the cell-ordered particle arrays: nothing is padded
        return 0
