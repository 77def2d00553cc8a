"""Host-side logic of the domain decomposition (no GPU): the grid choice, particle placement and the
staged ghost plan of the C-ABI library, exercised across two real processes over gloo. The exchange
itself is mirrored with torch.distributed point-to-point calls in the order the engine issues its NCCL
calls (above first, then below; receive from below, then from above), and the result is checked with
the CPU oracle: every owned particle must find its complete neighbour set among owned + ghost
particles, and the union of the ranks' lists must be the undivided system's pair set.
"""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _lib():
    from hoomd_blue_b200 import _lib
    return _lib


def test_find_decomposition_matches_reference_choices():
    """hoomd/DomainDecomposition.cc:327-409 on a cube: 1x1x2, 1x2x2, 2x2x2; slabs follow the long axis."""
    L = _lib().lib()
    g = (C.c_uint32 * 3)()
    for n, want in [(1, (1, 1, 1)), (2, (1, 1, 2)), (4, (1, 2, 2)), (8, (2, 2, 2)), (3, (1, 1, 3)),
                    (6, (1, 2, 3))]:
        assert L.hb2_domain_find_decomposition(n, 10.0, 10.0, 10.0, g) == 0
        assert tuple(g) == want, (n, tuple(g))
    assert L.hb2_domain_find_decomposition(2, 40.0, 10.0, 10.0, g) == 0
    assert tuple(g) == (2, 1, 1)
    assert L.hb2_domain_find_decomposition(4, 40.0, 10.0, 10.0, g) == 0
    assert tuple(g) == (4, 1, 1)
    assert L.hb2_domain_find_decomposition(0, 1.0, 1.0, 1.0, g) != 0


def _place(box, grid, pos):
    lib = _lib()
    out = np.zeros(len(pos), dtype=np.uint32)
    g = (C.c_uint32 * 3)(*grid)
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    lib.check(lib.lib().hb2_domain_place(C.byref(box), g, lib.ptr(pos, C.c_float), len(pos),
                                         lib.ptr(out, C.c_uint32)))
    return out


def _ghost_mask(box, grid, coord, r_ghost, pos):
    lib = _lib()
    out = np.zeros(len(pos), dtype=np.uint8)
    g = (C.c_uint32 * 3)(*grid)
    c = (C.c_uint32 * 3)(*coord)
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    lib.check(lib.lib().hb2_domain_ghost_mask(C.byref(box), g, c, r_ghost, lib.ptr(pos, C.c_float),
                                              len(pos), out.ctypes.data_as(C.POINTER(C.c_uint8))))
    return out


def _box(L):
    lib = _lib()
    b = lib.Box()
    lib.check(lib.lib().hb2_box_make(C.byref(b), *[float(x) for x in L]))
    return b


def test_place_particles_partition():
    """Every particle has exactly one owner; faces belong to the upper domain (f in [k/n, (k+1)/n))."""
    rng = np.random.default_rng(5)
    L = [12.0, 9.0, 15.0]
    box = _box(L)
    pos = (rng.random((5000, 3)) - 0.5) * np.array(L)
    pos[0] = [-6.0, -4.5, -7.5]   # the lower corner belongs to rank 0
    pos[1] = [0.0, 0.0, 0.0]      # the centre plane belongs to the upper half
    pos = pos.astype(np.float32)
    for grid in [(1, 1, 2), (1, 2, 2), (2, 2, 2), (3, 1, 2)]:
        r = _place(box, grid, pos)
        assert r.max() < grid[0] * grid[1] * grid[2]
        assert r[0] == 0
        f = (pos.astype(np.float64) + np.array(L) / 2) / np.array(L)
        c = np.minimum((f * np.array(grid)).astype(int), np.array(grid) - 1)
        want = (c[:, 2] * grid[1] + c[:, 1]) * grid[0] + c[:, 0]
        assert np.array_equal(r, want.astype(np.uint32))
    bad = np.array([[7.0, 0, 0]], dtype=np.float32)
    with pytest.raises(_lib().Hb2Error):
        _place(box, (1, 1, 2), bad)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, grid, q):
    try:
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from oracle import md_oracle as orc
        from hoomd_blue_b200.communicator import Communicator, gather_by_tag, all_gather_arrays

        comm = Communicator()
        assert (comm.rank, comm.num_ranks) == (rank, world)
        rng = np.random.default_rng(11)  # same system on every rank
        n_side, rho, r_cut, buf = 12, 0.8, 2.0, 0.3
        N = n_side ** 3
        Lx = (N / rho) ** (1 / 3)
        L = [np.float32(Lx), np.float32(Lx), np.float32(Lx * 1.0)]
        g = (np.arange(n_side) + 0.5) * (float(L[0]) / n_side) - float(L[0]) / 2
        x, y, z = np.meshgrid(g, g, g, indexing="ij")
        pos = (np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
               + rng.uniform(-0.35, 0.35, size=(N, 3))).astype(np.float32)
        half = np.float32(float(L[0]) / 2)
        pos = np.where(pos >= half, pos - L[0], pos)
        pos = np.where(pos < -half, pos + L[0], pos).astype(np.float32)
        box = _box(L)
        r_list = r_cut + buf
        coord = (rank % grid[0], (rank // grid[0]) % grid[1], rank // (grid[0] * grid[1]))

        def nbr(d, up):
            c = list(coord)
            c[d] = (c[d] + (1 if up else grid[d] - 1)) % grid[d]
            return (c[2] * grid[1] + c[1]) * grid[0] + c[0]

        owner = _place(box, grid, pos)
        mine = np.nonzero(owner == rank)[0]
        tags = mine.astype(np.uint32)
        p = pos[mine]
        n_owned = len(p)
        # staged ghost exchange, dimension by dimension, forwarding earlier ghosts
        for d in range(3):
            if grid[d] == 1:
                continue
            m = _ghost_mask(box, grid, coord, r_list, p)
            up = np.nonzero(m & (1 << (2 * d + 1)))[0]
            dn = np.nonzero(m & (1 << (2 * d)))[0]
            sends = [(nbr(d, True), up), (nbr(d, False), dn)]   # above first, then below
            recvs = [nbr(d, False), nbr(d, True)]               # from below, then from above
            got = []
            for k in range(2):
                dst, idx = sends[k]
                payload = [np.concatenate([tags[idx].astype(np.float64)[:, None],
                                           p[idx].astype(np.float64)], axis=1)]
                # two ranks: the peer is both neighbours; pair the k-th send with the k-th receive
                box_ = [None]
                if rank < dst or (rank == dst):
                    dist.send_object_list(payload, dst=dst)
                    dist.recv_object_list(box_, src=recvs[k])
                else:
                    dist.recv_object_list(box_, src=recvs[k])
                    dist.send_object_list(payload, dst=dst)
                got.append(box_[0])
            for a in got:
                tags = np.concatenate([tags, a[:, 0].astype(np.uint32)])
                p = np.concatenate([p, a[:, 1:].astype(np.float32)])
        assert len(np.unique(tags)) == len(tags), "a particle arrived twice"

        # oracle on the undivided system and on this rank's owned + ghost set (global box)
        obox = orc.make_box([float(v) for v in L])
        types = np.zeros(N, dtype=np.uint32)
        nn, head, nl, _ = orc.nlist(obox, pos, types, [r_cut], buf, full=True)
        nn_l, head_l, nl_l, _ = orc.nlist(obox, p, np.zeros(len(p), dtype=np.uint32), [r_cut], buf,
                                          full=True)
        local_pairs = []
        for i in range(n_owned):
            want = np.sort(nl[head[i0]:head[i0] + nn[i0]]) if (i0 := int(tags[i])) >= 0 else None
            have = np.sort(tags[nl_l[head_l[i]:head_l[i] + nn_l[i]]])
            assert np.array_equal(want, have), f"rank {rank}: neighbours of tag {i0} differ"
            for j in have:
                jl = int(np.nonzero(tags == j)[0][0])
                if (jl < n_owned and i0 < j) or (jl >= n_owned and i0 < j):
                    local_pairs.append((i0, int(j)))
        lp = np.array(local_pairs, dtype=np.uint32).reshape(-1, 2)
        allp = np.concatenate(all_gather_arrays(lp), axis=0)
        key = np.unique(allp[:, 0].astype(np.uint64) << np.uint64(32) | allp[:, 1].astype(np.uint64))
        assert len(key) == len(allp), "a pair was reported twice"
        ref = orc.pair_set(nn, head, nl)
        ref_key = ref[:, 0].astype(np.uint64) << np.uint64(32) | ref[:, 1].astype(np.uint64)
        assert np.array_equal(key, np.sort(ref_key))
        # gather_by_tag puts every rank's rows back in tag order
        back = gather_by_tag(N, tags[:n_owned], p[:n_owned])
        assert np.array_equal(back, pos)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", n_owned, len(p) - n_owned))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + traceback.format_exc(), 0, 0))


@pytest.mark.parametrize("grid", [(1, 1, 2), (2, 1, 1)])
def test_two_rank_ghost_plan_over_gloo(grid):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, grid, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] == "ok", r[1]
    assert sum(r[2] for r in res) == 12 ** 3
    assert all(r[3] > 0 for r in res)


def test_fill_holes_plan():
    """The in-place compaction of a migration stage (csrc/comm.cu fill_holes_kernel), restated: with the m
    leaving indices in ascending order, the r-th stayer of the tail [N - m, N) moves into the r-th hole below
    N - m. Afterwards [0, N - m) holds exactly the stayers, each once, and the plan is a pure function of the
    leaver set."""
    import numpy as np
    rng = np.random.default_rng(3)
    for N, m in ((1, 0), (1, 1), (10, 3), (1000, 37), (1000, 999), (64, 64), (5000, 1)):
        leave = np.sort(rng.choice(N, size=m, replace=False)).astype(np.int64)
        is_leaver = np.zeros(N, dtype=bool)
        is_leaver[leave] = True
        data = np.arange(N) * 10 + 7  # a payload that names its original slot
        out = data.copy()
        n_stay = N - m
        h = int(np.searchsorted(leave, n_stay, side="left"))  # holes below the new end
        moved = 0
        for k in range(m):  # one thread per tail slot, as the kernel
            t = n_stay + k
            if is_leaver[t]:
                continue
            r = k - (int(np.searchsorted(leave, t, side="left")) - h)
            assert 0 <= r < h
            out[leave[r]] = data[t]
            moved += 1
        assert moved == h
        assert sorted(out[:n_stay].tolist()) == sorted(data[~is_leaver].tolist())
