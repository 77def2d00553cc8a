"""Layer 1 of the C ABI (the stateless kernel drivers a HOOMD `...GPU` class would call) on
caller-owned device arrays — here torch CUDA tensors used purely as device memory — against the CPU
oracle: cell order -> neighbour list -> pair forces -> ConstantVolume half steps."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _p(t):
    return C.c_void_p(t.data_ptr())


def test_stateless_drivers_against_oracle(oracle):
    import torch
    from hoomd_blue_b200 import _lib
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    n_side, rho, r_cut, buf = 24, 0.85, 2.5, 0.4
    N = n_side ** 3
    Lb = float(np.float32((N / rho) ** (1 / 3)))
    g = (np.arange(n_side) + 0.5) * (Lb / n_side) - Lb / 2
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    pos = (np.stack([x.ravel(), y.ravel(), z.ravel()], 1) + rng.uniform(-0.2, 0.2, (N, 3))).astype(np.float32)
    pos = np.clip(pos, -Lb / 2, np.nextafter(np.float32(Lb / 2), np.float32(0)))
    box = _lib.Box()
    _lib.check(L.hb2_box_make(C.byref(box), Lb, Lb, Lb))
    d = int(Lb / (r_cut + buf))
    dim = (C.c_uint32 * 3)(d, d, d)
    Nc = d ** 3

    pos4 = np.zeros((N + 1, 4), dtype=np.float32)
    pos4[:N, :3] = pos
    d_pos = torch.from_numpy(pos4).to(dev)
    d_order = torch.zeros(N, dtype=torch.int32, device=dev)
    d_cs = torch.zeros(Nc, dtype=torch.int32, device=dev)
    d_ce = torch.zeros(Nc, dtype=torch.int32, device=dev)
    d_cond = torch.zeros(4, dtype=torch.int32, device=dev)
    nbytes = C.c_size_t(0)
    _lib.check(L.hb2_gpu_compute_cell_order(None, N, C.byref(box), dim, None, None, None, None, None,
                                            C.byref(nbytes), None))
    d_tmp = torch.zeros(nbytes.value, dtype=torch.uint8, device=dev)
    _lib.check(L.hb2_gpu_compute_cell_order(_p(d_pos), N, C.byref(box), dim, _p(d_order), _p(d_cs),
                                            _p(d_ce), _p(d_cond), _p(d_tmp), C.byref(nbytes), None))
    torch.cuda.synchronize()
    order = d_order.cpu().numpy().astype(np.int64)
    assert sorted(order.tolist()) == list(range(N))
    assert d_cond.cpu().numpy()[1:3].tolist() == [0, 0]
    cs, ce = d_cs.cpu().numpy(), d_ce.cpu().numpy()
    assert (ce - cs).sum() == N and (ce - cs).max() < 64

    # the caller applies the order (what SFCPackTuner does), then builds the list
    d_pos_s = torch.cat([d_pos[:N][torch.from_numpy(order).to(dev)], d_pos[N:]]).contiguous()
    pitch = 160
    d_nlist = torch.zeros(N * pitch, dtype=torch.int32, device=dev)
    d_nn = torch.zeros(N, dtype=torch.int32, device=dev)
    d_last = torch.zeros((N, 4), dtype=torch.float32, device=dev)
    rl = np.array([(r_cut + buf) ** 2], dtype=np.float64)
    d_rl = torch.from_numpy(rl).to(dev)
    d_cond.zero_()
    _lib.check(L.hb2_gpu_compute_nlist_binned(_p(d_pos_s), N, N, C.byref(box), dim, _p(d_cs), _p(d_ce),
                                              _p(d_rl), 1, _p(d_nlist), pitch, _p(d_nn), _p(d_last),
                                              _p(d_cond), None))
    torch.cuda.synchronize()
    nn = d_nn.cpu().numpy()
    assert d_cond.cpu().numpy()[0] == nn.max() <= pitch
    nl = d_nlist.cpu().numpy().reshape(N, pitch)
    ps = d_pos_s.cpu().numpy()[:N, :3]
    obox = oracle.make_box([Lb] * 3)
    types = np.zeros(N, dtype=np.uint32)
    onn, ohead, onl, _ = oracle.nlist(obox, ps, types, [r_cut], buf, full=True)
    assert np.array_equal(nn, onn)
    for i in rng.integers(0, N, 200):
        assert np.array_equal(nl[i, :nn[i]], np.sort(onl[ohead[i]:ohead[i] + onn[i]]))
        assert (nl[i, nn[i]:(nn[i] + 31) // 32 * 32] == N).all()      # padded with the dummy index
    assert np.array_equal(d_last.cpu().numpy()[:, :3], ps)

    # pair forces
    prm = (C.c_float * _lib.PARAM_FLOATS)()
    _lib.check(L.hb2_pack_params(_lib.EVAL["lj"], (C.c_double * 2)(1.0, 1.0), 2, prm))
    d_prm = torch.tensor(list(prm), dtype=torch.float32, device=dev)
    d_rc = torch.tensor([r_cut ** 2], dtype=torch.float32, device=dev)
    d_ro = torch.zeros(1, dtype=torch.float32, device=dev)
    d_f = torch.zeros((N, 4), dtype=torch.float32, device=dev)
    d_v = torch.zeros(6 * N, dtype=torch.float32, device=dev)
    _lib.check(L.hb2_gpu_compute_pair_forces(0, 0, _p(d_pos_s), None, None, N, C.byref(box), _p(d_nn),
                                             _p(d_nlist), pitch, _p(d_prm), _p(d_rc), _p(d_ro), 1,
                                             _p(d_f), _p(d_v), N, 0, 0, 0.0, 0.0,
                                             r_cut + 2 * buf, 8, None))
    torch.cuda.synchronize()
    f = d_f.cpu().numpy().astype(np.float64)
    f_ref, v_ref = oracle.pair_lj(obox, ps, types, onn, ohead, onl, [[1.0, 1.0]], [r_cut])
    nref = np.linalg.norm(f_ref[:, :3], axis=1)
    err = np.max(np.linalg.norm(f[:, :3] - f_ref[:, :3], axis=1) / np.maximum(nref, np.sqrt(np.mean(nref ** 2))))
    assert err <= 1e-5
    assert abs(f[:, 3].sum() - f_ref[:, 3].sum()) <= 1e-6 * abs(f_ref[:, 3].sum())
    vs = d_v.cpu().numpy().astype(np.float64).reshape(6, N).sum(axis=1)
    assert np.max(np.abs(vs[[0, 3, 5]] - v_ref.sum(axis=1)[[0, 3, 5]])) <= 1e-6 * np.abs(v_ref.sum(axis=1)[[0, 3, 5]]).max()

    # half steps (hoomd/md/TwoStepConstantVolume.cc:33-79, :234-260) against plain numpy fp32
    vel4 = np.zeros((N, 4), dtype=np.float32)
    vel4[:, :3] = rng.normal(0, 1, (N, 3))
    vel4[:, 3] = 1.0
    d_vel = torch.from_numpy(vel4).to(dev)
    d_img = torch.zeros(3 * N, dtype=torch.int32, device=dev)
    d_flag = torch.zeros(1, dtype=torch.int32, device=dev)
    dt = np.float32(0.005)
    p0 = d_pos_s.cpu().numpy()[:N].copy()
    f32 = d_f.cpu().numpy()
    _lib.check(L.hb2_gpu_nve_step_one(_p(d_pos_s), _p(d_vel), _p(d_f), _p(d_img), N, C.byref(box), dt,
                                      None, _p(d_last), 0.2, _p(d_flag), None))
    torch.cuda.synchronize()
    v_half = vel4[:, :3] + np.float32(0.5) * dt * f32[:, :3]
    p1 = p0[:, :3] + dt * v_half
    wrapped = np.where(p1 >= np.float32(Lb / 2), p1 - np.float32(Lb), np.where(p1 < np.float32(-Lb / 2), p1 + np.float32(Lb), p1))
    assert np.allclose(d_vel.cpu().numpy()[:, :3], v_half, rtol=1e-6, atol=1e-5)
    assert np.allclose(d_pos_s.cpu().numpy()[:N, :3], wrapped, rtol=0, atol=1e-5)
    assert int(d_flag.item()) == 0
    d_ke = torch.zeros(1, dtype=torch.float64, device=dev)
    _lib.check(L.hb2_gpu_nve_step_two(_p(d_vel), _p(d_f), N, dt, None, _p(d_ke), None))
    torch.cuda.synchronize()
    v1 = v_half + np.float32(0.5) * dt * f32[:, :3]
    assert np.allclose(d_vel.cpu().numpy()[:, :3], v1, rtol=1e-6, atol=1e-5)
    assert abs(float(d_ke.item()) - 0.5 * (v1.astype(np.float64) ** 2).sum()) <= 1e-6 * float(d_ke.item())
