"""Bonds and bond exclusions: hoomd.md.bond.Harmonic, Snapshot.bonds and
nlist.Cell(exclusions=('bond',)) against the oracle's restatement (hoomd/md/PotentialBond.h:176-312,
hoomd/md/EvaluatorBondHarmonic.h, hoomd/md/NeighborList.cc:569-640,1448-1536), which is pinned to the
reference's own vectors (hoomd/md/pytest/test_bond.py:17-37, two particles 0.969 apart in a box of 5)."""
import os

import numpy as np
import pytest

from tests import gsd_ref

# (k, r0, force on particle 0 along x, total energy) hoomd/md/pytest/test_bond.py:17-37
REF_VECTORS = [(30.0, 1.6, -18.9300, 5.9724), (25.0, 1.7, -18.2750, 6.6795), (20.0, 1.8, -16.6200, 6.9056)]


# (params, force on particle 0 along x, total energy) hoomd/md/pytest/test_bond.py:38-58
FENE_VECTORS = [(dict(k=30.0, r0=1.6, epsilon=0.9, sigma=1.1, delta=-0.5), 282.296, 70.5638),
                (dict(k=25.0, r0=1.7, epsilon=1.0, sigma=1.0, delta=-0.5), 146.288, 49.2476),
                (dict(k=20.0, r0=1.8, epsilon=1.1, sigma=0.9, delta=-0.5), 88.8238, 35.3135)]
FENE_KEYS = ("k", "r0", "epsilon", "sigma", "delta")


# (kind, k, t0, force, energy) hoomd/md/pytest/test_angle.py:14-57: three particles at 60 degrees
ANGLE_VECTORS = [("harmonic", 3.0, np.pi / 2, -1.5708, 0.4112), ("harmonic", 10.0, np.pi / 4, 2.6180, 0.3427),
                 ("harmonic", 5.0, np.pi / 6, 2.6180, 0.6854), ("cosinesq", 3.0, np.pi / 2, -1.29904, 0.375),
                 ("cosinesq", 10.0, np.pi / 4, 1.7936, 0.214466), ("cosinesq", 5.0, np.pi / 6, 1.58494, 0.334936)]


def triplet(d=1.0, theta_deg=60):
    # hoomd/md/pytest/test_angle.py:69-107 triplet_snapshot_factory
    th = theta_deg * np.pi / 180
    p = np.array([[-d * np.sin(th / 2), d * np.cos(th / 2), 0.0], [0.0, 0.0, 0.0],
                  [d * np.sin(th / 2), d * np.cos(th / 2), 0.0]])
    p[:, 2] += 0.1
    return p.astype(np.float32)


def check_angle_vectors(forces, energy, force, e_ref):
    th = 60 * np.pi / 180
    fa = force * np.array([np.cos(th / 2), np.sin(th / 2), 0])
    assert energy == pytest.approx(e_ref, rel=1e-2)
    np.testing.assert_allclose(forces[0], fa, rtol=1e-2, atol=1e-5)
    np.testing.assert_allclose(forces[1], [0, -1 * force, 0], rtol=1e-2, atol=1e-5)
    np.testing.assert_allclose(forces[2], [-fa[0], fa[1], fa[2]], rtol=1e-2, atol=1e-5)


@pytest.mark.parametrize("kind,k,t0,force,energy", ANGLE_VECTORS)
def test_oracle_angle_reference_vectors(oracle, kind, k, t0, force, energy):
    box = oracle.make_box([20.0, 20.0, 20.0])
    f, v = oracle.angle(kind, box, triplet(), [[0, 1, 2]], [0], [[k, t0]])
    check_angle_vectors(f[:, :3], f[:, 3].sum(), force, energy)
    assert np.abs(f[:, :3].sum(axis=0)).max() < 1e-12          # no net force
    pos = triplet().astype(np.float64)
    assert np.abs(np.cross(pos, f[:, :3]).sum(axis=0)).max() < 1e-6  # no net torque
    # across the periodic boundary: same forces for a translated, wrapped triplet
    shifted = (pos + [9.7, 9.8, 0]) 
    shifted = (shifted + 10.0) % 20.0 - 10.0
    f2, _ = oracle.angle(kind, box, shifted, [[0, 1, 2]], [0], [[k, t0]])
    np.testing.assert_allclose(f2, f, atol=1e-6)


# (k, d, n, phi0, force, energy) hoomd/md/pytest/test_dihedral.py:14-35: four particles, phi = 45 degrees
DIHEDRAL_VECTORS = [(3.0, -1, 2, np.pi / 2, 0, 3), (10.0, 1, 1, np.pi / 4, 5.0, 5.0),
                    (5.0, 1, 3, np.pi / 6, 1.9411, 0.0852)]


# (k1, k2, k3, k4, force, energy) hoomd/md/pytest/test_dihedral.py:36-56
OPLS_VECTORS = [(1.0, 1.5, 0.5, 0.75, -0.616117, 2.42678), (0.5, 2.5, 1.5, 1.0, -0.732233, 2.89645),
                (2.0, 1.0, 0.25, 3.5, -0.0277282, 5.74372)]


def quadruplet(d=1.0, phi_deg=45):
    # hoomd/md/pytest/test_dihedral.py:69-104 dihedral_snapshot_factory
    ph = phi_deg * np.pi / 180
    return np.array([[0.0, d * np.cos(ph / 2), d * np.sin(ph / 2) + 0.1], [0.0, 0.0, 0.1], [d, 0.0, 0.1],
                     [d, d * np.cos(ph / 2), -d * np.sin(ph / 2) + 0.1]], dtype=np.float32)


def check_dihedral_vectors(forces, energy, force, e_ref):
    ph = 45 * np.pi / 180
    fa = force * np.array([0, np.sin(-ph / 2), np.cos(-ph / 2)])
    assert energy == pytest.approx(e_ref, rel=1e-2, abs=1e-5)
    np.testing.assert_allclose(forces[0], fa, rtol=1e-2, atol=1e-5)
    np.testing.assert_allclose(forces[1], -1 * fa, rtol=1e-2, atol=1e-5)
    np.testing.assert_allclose(forces[2], [0, -1 * fa[1], fa[2]], rtol=1e-2, atol=1e-5)
    np.testing.assert_allclose(forces[3], [0, fa[1], -1 * fa[2]], rtol=1e-2, atol=1e-5)


@pytest.mark.parametrize("k,d,n,phi0,force,energy", DIHEDRAL_VECTORS)
def test_oracle_dihedral_reference_vectors(oracle, k, d, n, phi0, force, energy):
    box = oracle.make_box([20.0, 20.0, 20.0])
    f, v = oracle.dihedral_periodic(box, quadruplet(), [[0, 1, 2, 3]], [0], [[k, d, n, phi0]])
    check_dihedral_vectors(f[:, :3], f[:, 3].sum(), force, energy)
    assert np.abs(f[:, :3].sum(axis=0)).max() < 1e-12
    pos = quadruplet().astype(np.float64)
    assert np.abs(np.cross(pos, f[:, :3]).sum(axis=0)).max() < 1e-6
    # F = -dE/dx by central differences of the oracle's energy, every particle and component
    h = 1e-6
    for i in range(4):
        for c in range(3):
            pp, pm = pos.copy(), pos.copy()
            pp[i, c] += h
            pm[i, c] -= h
            ep = oracle.dihedral_periodic(box, pp, [[0, 1, 2, 3]], [0], [[k, d, n, phi0]])[0][:, 3].sum()
            em = oracle.dihedral_periodic(box, pm, [[0, 1, 2, 3]], [0], [[k, d, n, phi0]])[0][:, 3].sum()
            assert -(ep - em) / (2 * h) == pytest.approx(f[i, c], rel=1e-4, abs=1e-6)


@pytest.mark.parametrize("k1,k2,k3,k4,force,energy", OPLS_VECTORS)
def test_oracle_opls_reference_vectors(oracle, k1, k2, k3, k4, force, energy):
    box = oracle.make_box([20.0, 20.0, 20.0])
    prm = [[k1, k2, k3, k4]]
    f, v = oracle.dihedral_opls(box, quadruplet(), [[0, 1, 2, 3]], [0], prm)
    check_dihedral_vectors(f[:, :3], f[:, 3].sum(), force, energy)
    assert np.abs(f[:, :3].sum(axis=0)).max() < 1e-12
    pos = quadruplet().astype(np.float64)
    h = 1e-6
    for i in range(4):
        for c in range(3):
            pp, pm = pos.copy(), pos.copy()
            pp[i, c] += h
            pm[i, c] -= h
            ep = oracle.dihedral_opls(box, pp, [[0, 1, 2, 3]], [0], prm)[0][:, 3].sum()
            em = oracle.dihedral_opls(box, pm, [[0, 1, 2, 3]], [0], prm)[0][:, 3].sum()
            assert -(ep - em) / (2 * h) == pytest.approx(f[i, c], rel=1e-4, abs=1e-6)
    # the closed form: V = 1/2 [k1 (1 + cos) + k2 (1 - cos 2) + k3 (1 + cos 3) + k4 (1 - cos 4)] at 45 deg
    ph = np.pi / 4
    want = 0.5 * (k1 * (1 + np.cos(ph)) + k2 * (1 - np.cos(2 * ph)) + k3 * (1 + np.cos(3 * ph))
                  + k4 * (1 - np.cos(4 * ph)))
    assert f[:, 3].sum() == pytest.approx(want, rel=1e-6)


def two_particles(d=0.969):
    # hoomd/conftest.py:168-196 two_particle_snapshot_factory(d=0.969, L=5)
    return np.array([[-d / 2, 0, 0.1], [d / 2, 0, 0.1]], dtype=np.float32)


@pytest.mark.parametrize("k,r0,force,energy", REF_VECTORS)
def test_oracle_harmonic_reference_vectors(oracle, k, r0, force, energy):
    box = oracle.make_box([5.0, 5.0, 5.0])
    f, v = oracle.bond_harmonic(box, two_particles(), [[0, 1]], [0], [[k, r0]])
    assert f[:, 3].sum() == pytest.approx(energy, rel=1e-2)   # the reference's own tolerances
    np.testing.assert_allclose(f[0, :3], [force, 0, 0], rtol=1e-2, atol=1e-5)
    np.testing.assert_allclose(f[1, :3], [-force, 0, 0], rtol=1e-2, atol=1e-5)
    # and exactly: F = k (r0 / r - 1) dx, E = k/2 (r - r0)^2, W_xx = 1/2 F_x dx split over both members
    r = float(np.float32(0.969 / 2)) * 2
    assert f[0, 0] == pytest.approx(-k * (r0 / r - 1.0) * r, rel=1e-12)
    assert f[:, 3].sum() == pytest.approx(0.5 * k * (r - r0) ** 2, rel=1e-12)
    assert v[0].sum() == pytest.approx(k * (r0 / r - 1.0) * r * r, rel=1e-12)
    # a bond across the periodic boundary uses the minimum image
    far = np.array([[-2.4, 0, 0], [2.4, 0, 0]], dtype=np.float32)
    f2, _ = oracle.bond_harmonic(box, far, [[0, 1]], [0], [[k, r0]])
    assert f2[0, 0] == pytest.approx(+k * (r0 / 0.2 - 1.0) * 0.2, rel=1e-5)


@pytest.mark.parametrize("prm,force,energy", FENE_VECTORS)
def test_oracle_fenewca_reference_vectors(oracle, prm, force, energy):
    box = oracle.make_box([5.0, 5.0, 5.0])
    f, v = oracle.bond_fenewca(box, two_particles(), [[0, 1]], [0], [[prm[k] for k in FENE_KEYS]])
    assert f[:, 3].sum() == pytest.approx(energy, rel=1e-2)
    np.testing.assert_allclose(f[0, :3], [force, 0, 0], rtol=1e-2, atol=1e-5)
    np.testing.assert_allclose(f[1, :3], [-force, 0, 0], rtol=1e-2, atol=1e-5)
    # F = -dE/dr by central differences of the oracle's own energy
    h = 1e-5
    e = [oracle.bond_fenewca(box, np.array([[-d / 2, 0, 0.1], [d / 2, 0, 0.1]]), [[0, 1]], [0],
                             [[prm[k] for k in FENE_KEYS]])[0][:, 3].sum() for d in (0.969 - h, 0.969 + h)]
    r = float(np.float32(0.969 / 2)) * 2
    assert -f[1, 0] == pytest.approx((e[1] - e[0]) / (2 * h), rel=1e-4)  # force on 0 = +dE/dr (pushes apart)
    assert r > 0
    # stretched beyond r0 (in r - delta): "bond out of bounds"
    with pytest.raises(RuntimeError, match="out of bounds"):
        oracle.bond_fenewca(box, two_particles(prm["r0"] + prm["delta"] + 0.01), [[0, 1]], [0],
                            [[prm[k] for k in FENE_KEYS]])


def chains(n_chains, length, rho, seed, bond=0.97):
    """random-walk bead-spring chains in a cubic box, wrapped"""
    rng = np.random.default_rng(seed)
    n = n_chains * length
    L = np.float32((n / rho) ** (1 / 3))
    pos = np.zeros((n, 3))
    group = []
    for c in range(n_chains):
        p = rng.uniform(-L / 2, L / 2, 3)
        for k in range(length):
            i = c * length + k
            pos[i] = p
            if k:
                group.append((i - 1, i))
            step = rng.normal(size=3)
            p = p + bond * step / np.linalg.norm(step)
    pos = (pos + L / 2) % float(L) - L / 2
    pos = np.clip(pos.astype(np.float32), -float(L) / 2, np.nextafter(np.float32(float(L) / 2), np.float32(0)))
    return pos, np.array(group, dtype=np.uint32), np.array([L, L, L], dtype=np.float32)


def lattice_chains(nx, ny, nz, length, a, seed, jitter=0.05):
    """chains laid along x on a cubic lattice (no overlaps): a gentle start for dynamics"""
    assert nx % length == 0
    rng = np.random.default_rng(seed)
    L = np.array([nx * a, ny * a, nz * a], dtype=np.float32)
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    order = np.lexsort((ix.ravel(), iy.ravel(), iz.ravel()))  # x fastest
    pos = (np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1)[order] + 0.5) * a - L.astype(np.float64) / 2
    pos += rng.uniform(-jitter, jitter, size=pos.shape)
    n = len(pos)
    idx = np.arange(n)
    group = np.stack([idx[:-1], idx[1:]], axis=1)[(idx[1:] % length) != 0]
    pos = np.clip(pos.astype(np.float32), -L / 2, np.nextafter(L / 2, np.float32(0)))
    return pos, group.astype(np.uint32), L


def test_oracle_exclusion_filter(oracle):
    """filtered list == full list minus the bonded pairs that were in it"""
    pos, group, L = chains(40, 10, 0.6, seed=1)
    box = oracle.make_box([float(v) for v in L])
    types = np.zeros(len(pos), dtype=np.uint32)
    nn, head, nl, _ = oracle.nlist(box, pos, types, [2.0], 0.3, full=True)
    full = set(map(tuple, oracle.pair_set(nn, head, nl)))
    nn2, head2, nl2 = oracle.filter_bond_exclusions(nn, head, nl, group)
    filt = set(map(tuple, oracle.pair_set(nn2, head2, nl2)))
    bonded = set((int(min(a, b)), int(max(a, b))) for a, b in group)
    assert filt == full - bonded and len(full & bonded) == len(group)  # bond length 0.97 < r_list
    assert nn.sum() - nn2.sum() == 2 * len(group)


def test_extra_exclusion_pairs():
    """'angle' and '1-3' exclusion pairs against a literal restatement of the reference's loops
    (hoomd/md/NeighborList.cc:881-909 and :1048-1120)."""
    from hoomd_blue_b200.md.nlist import Cell
    rng = np.random.default_rng(11)
    n = 60
    bonds = set()
    while len(bonds) < 80:
        a, b = rng.integers(0, n, 2)
        if a != b and (a, b) not in bonds and (b, a) not in bonds:
            bonds.add((int(a), int(b)))
    bonds = np.array(sorted(bonds), dtype=np.uint32)
    angles = rng.integers(0, n, (25, 3)).astype(np.uint32)
    angles = angles[(angles[:, 0] != angles[:, 2])]

    class _State:
        _bond_group, _angle_group = bonds, angles

    # reference: per-atom partner lists, then every two partners of a common atom
    partners = {}
    for a, b in bonds.tolist():
        partners.setdefault(a, []).append(b)
        partners.setdefault(b, []).append(a)
    want13 = set()
    for c, p in partners.items():
        for j in range(len(p)):
            for k in range(j + 1, len(p)):
                if p[j] != p[k]:
                    want13.add((min(p[j], p[k]), max(p[j], p[k])))
    want_angle = set((int(min(a, c)), int(max(a, c))) for a, _, c in angles.tolist())
    got = lambda ex: set(map(tuple, Cell(buffer=0.4, exclusions=ex)._extra_exclusion_pairs(_State).tolist()))  # noqa: E731
    assert got(("bond", "1-3")) == want13
    assert got(("angle",)) == want_angle
    assert got(("bond", "angle", "1-3")) == want13 | want_angle
    assert got(("bond",)) == set()
    with pytest.raises(NotImplementedError):
        Cell(buffer=0.4, exclusions=("bond", "dihedral"))


def test_gsd_bonds_roundtrip(tmp_path):
    from hoomd_blue_b200 import gsd_io, Snapshot
    pos, group, L = chains(5, 6, 0.5, seed=2)
    s = Snapshot.from_arrays(list(L) + [0, 0, 0], pos)
    s.bonds.N = len(group)
    s.bonds.types = ["polymer", "x"]
    s.bonds.group[:] = group
    s.bonds.typeid[:] = np.arange(len(group)) % 2
    path = str(tmp_path / "b.gsd")
    gsd_io.write_frame(path, s, 5)
    r, step = gsd_io.read_frame(path)
    assert step == 5 and r.bonds.N == len(group) and r.bonds.types == ["polymer", "x"]
    assert np.array_equal(r.bonds.group, group) and np.array_equal(r.bonds.typeid, s.bonds.typeid)
    if gsd_ref.available():
        g = gsd_ref.RefGSD.open(path)
        assert np.array_equal(g.read(0, "bonds/group"), group.astype(np.int32))
        assert g.read(0, "bonds/N")[0, 0] == len(group)
        g.close()


# ------------------------------------------------------------------------------------ GPU ----
@pytest.mark.gpu
@pytest.mark.parametrize("k,r0,force,energy", REF_VECTORS)
def test_harmonic_reference_vectors(k, r0, force, energy):
    """hoomd/md/pytest/test_bond.py:113-144 as written there (a bond-only integrator)."""
    import hoomd_blue_b200 as hoomd
    snap = hoomd.Snapshot.from_arrays([5, 5, 5, 0, 0, 0], two_particles())
    snap.bonds.N = 1
    snap.bonds.types = ["A-A"]
    snap.bonds.typeid[0] = 0
    snap.bonds.group[0] = (0, 1)
    sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=1)
    sim.create_state_from_snapshot(snap)
    potential = hoomd.md.bond.Harmonic()
    potential.params["A-A"] = dict(k=k, r0=r0)
    sim.operations.integrator = hoomd.md.Integrator(dt=0.005, forces=[potential])
    sim.run(0)
    assert sum(potential.energies) == pytest.approx(energy, rel=1e-2)
    np.testing.assert_allclose(potential.forces[0], [force, 0.0, 0.0], rtol=1e-2, atol=1e-5)
    np.testing.assert_allclose(potential.forces[1], [-1 * force, 0.0, 0.0], rtol=1e-2, atol=1e-5)
    r = float(np.float32(0.969 / 2)) * 2
    assert potential.forces[0, 0] == pytest.approx(-k * (r0 / r - 1.0) * r, rel=1e-6)
    assert potential.energy == pytest.approx(0.5 * k * (r - r0) ** 2, rel=1e-6)
    back = sim.state.get_snapshot()
    assert back.bonds.N == 1 and back.bonds.types == ["A-A"] and tuple(back.bonds.group[0]) == (0, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("prm,force,energy", FENE_VECTORS)
def test_fenewca_reference_vectors(oracle, prm, force, energy):
    import hoomd_blue_b200 as hoomd

    def make(d):
        snap = hoomd.Snapshot.from_arrays([5, 5, 5, 0, 0, 0], two_particles(d))
        snap.bonds.N = 1
        snap.bonds.types = ["A-A"]
        snap.bonds.group[0] = (0, 1)
        sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=1)
        sim.create_state_from_snapshot(snap)
        potential = hoomd.md.bond.FENEWCA()
        potential.params["A-A"] = prm
        sim.operations.integrator = hoomd.md.Integrator(dt=0.005, forces=[potential])
        return sim, potential

    sim, potential = make(0.969)
    sim.run(0)
    assert sum(potential.energies) == pytest.approx(energy, rel=1e-2)
    np.testing.assert_allclose(potential.forces[0], [force, 0.0, 0.0], rtol=1e-2, atol=1e-5)
    np.testing.assert_allclose(potential.forces[1], [-1 * force, 0.0, 0.0], rtol=1e-2, atol=1e-5)
    box = oracle.make_box([5.0, 5.0, 5.0])
    f_ref, _ = oracle.bond_fenewca(box, two_particles(), [[0, 1]], [0], [[prm[k] for k in FENE_KEYS]])
    np.testing.assert_allclose(potential.forces, f_ref[:, :3], rtol=2e-5, atol=1e-5)
    assert potential.energy == pytest.approx(f_ref[:, 3].sum(), rel=2e-5)
    sim, potential = make(prm["r0"] + prm["delta"] + 0.01)
    with pytest.raises(RuntimeError, match="out of bounds"):
        sim.run(0)
        potential.forces


@pytest.mark.gpu
@pytest.mark.parametrize("kind,k,t0,force,energy", ANGLE_VECTORS)
def test_angle_reference_vectors(oracle, kind, k, t0, force, energy):
    """hoomd/md/pytest/test_angle.py:137-175 as written there."""
    import hoomd_blue_b200 as hoomd
    snap = hoomd.Snapshot.from_arrays([20, 20, 20, 0, 0, 0], triplet())
    snap.angles.N = 1
    snap.angles.types = ["A-A-A"]
    snap.angles.typeid[0] = 0
    snap.angles.group[0] = (0, 1, 2)
    sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=1)
    sim.always_compute_pressure = True
    sim.create_state_from_snapshot(snap)
    cls = hoomd.md.angle.Harmonic if kind == "harmonic" else hoomd.md.angle.CosineSquared
    potential = cls()
    potential.params["A-A-A"] = dict(k=k, t0=t0)
    sim.operations.integrator = hoomd.md.Integrator(dt=0.005, forces=[potential])
    sim.run(0)
    check_angle_vectors(potential.forces, potential.energy, force, energy)
    box = oracle.make_box([20.0, 20.0, 20.0])
    f_ref, v_ref = oracle.angle(kind, box, triplet(), [[0, 1, 2]], [0], [[k, t0]])
    np.testing.assert_allclose(potential.forces, f_ref[:, :3], rtol=2e-5, atol=2e-6)
    assert potential.energy == pytest.approx(f_ref[:, 3].sum(), rel=2e-5)
    np.testing.assert_allclose(potential.virials, v_ref.T, rtol=2e-5, atol=2e-6)
    back = sim.state.get_snapshot()
    assert back.angles.N == 1 and tuple(back.angles.group[0]) == (0, 1, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("k,d,n,phi0,force,energy", DIHEDRAL_VECTORS)
def test_dihedral_reference_vectors(oracle, k, d, n, phi0, force, energy):
    """hoomd/md/pytest/test_dihedral.py:138-177 as written there, plus the oracle to 2e-5."""
    import hoomd_blue_b200 as hoomd
    snap = hoomd.Snapshot.from_arrays([20, 20, 20, 0, 0, 0], quadruplet())
    snap.dihedrals.N = 1
    snap.dihedrals.types = ["A-A-A-A"]
    snap.dihedrals.group[0] = (0, 1, 2, 3)
    sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=1)
    sim.always_compute_pressure = True
    sim.create_state_from_snapshot(snap)
    potential = hoomd.md.dihedral.Periodic()
    potential.params["A-A-A-A"] = dict(k=k, d=d, n=n, phi0=phi0)
    sim.operations.integrator = hoomd.md.Integrator(dt=0.005, forces=[potential])
    sim.run(0)
    check_dihedral_vectors(potential.forces, sum(potential.energies), force, energy)
    box = oracle.make_box([20.0, 20.0, 20.0])
    f_ref, v_ref = oracle.dihedral_periodic(box, quadruplet(), [[0, 1, 2, 3]], [0], [[k, d, n, phi0]])
    np.testing.assert_allclose(potential.forces, f_ref[:, :3], rtol=2e-5, atol=2e-6)
    assert potential.energy == pytest.approx(f_ref[:, 3].sum(), rel=2e-5, abs=1e-6)
    np.testing.assert_allclose(potential.virials, v_ref.T, rtol=2e-5, atol=2e-6)
    assert tuple(sim.state.get_snapshot().dihedrals.group[0]) == (0, 1, 2, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("k1,k2,k3,k4,force,energy", OPLS_VECTORS)
def test_opls_reference_vectors(oracle, k1, k2, k3, k4, force, energy):
    import hoomd_blue_b200 as hoomd
    snap = hoomd.Snapshot.from_arrays([20, 20, 20, 0, 0, 0], quadruplet())
    snap.dihedrals.N = 1
    snap.dihedrals.types = ["A-A-A-A"]
    snap.dihedrals.group[0] = (0, 1, 2, 3)
    sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=1)
    sim.create_state_from_snapshot(snap)
    potential = hoomd.md.dihedral.OPLS()
    potential.params["A-A-A-A"] = dict(k1=k1, k2=k2, k3=k3, k4=k4)
    sim.operations.integrator = hoomd.md.Integrator(dt=0.005, forces=[potential])
    sim.run(0)
    check_dihedral_vectors(potential.forces, sum(potential.energies), force, energy)
    box = oracle.make_box([20.0, 20.0, 20.0])
    f_ref, _ = oracle.dihedral_opls(box, quadruplet(), [[0, 1, 2, 3]], [0], [[k1, k2, k3, k4]])
    np.testing.assert_allclose(potential.forces, f_ref[:, :3], rtol=2e-5, atol=2e-6)
    assert potential.energy == pytest.approx(f_ref[:, 3].sum(), rel=2e-5)


@pytest.mark.gpu
def test_semiflexible_chains_vs_oracle(oracle):
    """bonds + angles + WCA on lattice chains: every force against the oracle, the net force conserves
    energy through rebuilds (the angle table follows the particles through rtag)."""
    import hoomd_blue_b200 as hoomd
    from tests.test_gpu_parity import force_err, FORCE_RTOL
    pos, group, L = lattice_chains(32, 12, 12, 16, 1.0, seed=6, jitter=0.08)
    n = len(pos)
    bonded = set(map(tuple, group.tolist()))
    angles = np.array([(i - 1, i, i + 1) for i in range(1, n - 1)
                       if (i - 1, i) in bonded and (i, i + 1) in bonded], dtype=np.uint32)
    rng = np.random.default_rng(7)
    vel0 = rng.normal(0, 0.8, size=pos.shape).astype(np.float32)
    vel0 -= vel0.mean(axis=0)
    snap = hoomd.Snapshot.from_arrays(list(L) + [0, 0, 0], pos, vel0)
    snap.bonds.N = len(group)
    snap.bonds.types = ["polymer"]
    snap.bonds.group[:] = group
    snap.angles.N = len(angles)
    snap.angles.types = ["stiff", "soft"]
    snap.angles.group[:] = angles
    snap.angles.typeid[:] = np.arange(len(angles)) % 2
    dihedrals = np.array([(i - 1, i, i + 1, i + 2) for i in range(1, n - 2)
                          if (i - 1, i) in bonded and (i, i + 1) in bonded and (i + 1, i + 2) in bonded],
                         dtype=np.uint32)
    snap.dihedrals.N = len(dihedrals)
    snap.dihedrals.types = ["torsion"]
    snap.dihedrals.group[:] = dihedrals
    sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=2)
    sim.create_state_from_snapshot(snap)
    cell = hoomd.md.nlist.Cell(buffer=0.3)
    lj = hoomd.md.pair.LJ(nlist=cell, default_r_cut=2.0 ** (1.0 / 6.0), mode="shift")
    lj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
    hb = hoomd.md.bond.Harmonic()
    hb.params["polymer"] = dict(k=200.0, r0=0.95)
    ha = hoomd.md.angle.Harmonic()
    ha.params["stiff"] = dict(k=20.0, t0=np.pi)
    ha.params["soft"] = dict(k=5.0, t0=2.5)
    hd = hoomd.md.dihedral.Periodic()
    hd.params["torsion"] = dict(k=2.0, d=1, n=3, phi0=0.4)
    sim.operations.integrator = hoomd.md.Integrator(
        dt=0.001, methods=[hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())],
        forces=[lj, hb, ha, hd])
    sim.run(0)
    box = oracle.make_box([float(v) for v in L])
    a_ref, _ = oracle.angle("harmonic", box, pos, angles, snap.angles.typeid, [[20.0, np.pi], [5.0, 2.5]])
    assert force_err(ha.forces, a_ref[:, :3]) <= FORCE_RTOL
    assert abs(ha.energy - a_ref[:, 3].sum()) <= 1e-5 * abs(a_ref[:, 3].sum())
    b_ref, _ = oracle.bond_harmonic(box, pos, group, np.zeros(len(group), np.uint32), [[200.0, 0.95]])
    d_ref, _ = oracle.dihedral_periodic(box, pos, dihedrals, np.zeros(len(dihedrals), np.uint32),
                                        [[2.0, 1, 3, 0.4]])
    # nearly straight chains: the torsion forces are ill-conditioned (1 / |a x b|^2), compare to 1e-4
    assert force_err(hd.forces, d_ref[:, :3]) <= 1e-4
    assert abs(hd.energy - d_ref[:, 3].sum()) <= 1e-5 * abs(d_ref[:, 3].sum())
    ke, pe, _ = sim.operations.integrator._thermo()
    e_all = lj.energy + b_ref[:, 3].sum() + a_ref[:, 3].sum() + d_ref[:, 3].sum()
    assert abs(pe - e_all) <= 1e-5 * abs(pe)
    # exclusions=('bond', 'angle'): the chain ends of every angle leave the list too; for linear chains
    # ('bond', '1-3') names the same pairs
    from tests.test_gpu_parity import canon_pairs
    types = np.zeros(n, dtype=np.uint32)
    rc = 2.5  # the chain ends of an angle are ~2 apart: inside this list
    nn, head, nl, _ = oracle.nlist(box, pos, types, [rc], 0.3, full=True)
    ex = np.concatenate([group, angles[:, [0, 2]]])
    nn_x, head_x, nl_x = oracle.filter_bond_exclusions(nn, head, nl, ex)
    want = oracle.pair_set(nn_x, head_x, nl_x)
    nn_b, head_b, nl_b = oracle.filter_bond_exclusions(nn, head, nl, group)
    assert len(want) == len(oracle.pair_set(nn_b, head_b, nl_b)) - len(angles)
    for excl in (("bond", "angle"), ("bond", "1-3")):
        s2 = hoomd.Simulation(device=hoomd.device.GPU(), seed=2)
        s2.create_state_from_snapshot(snap)
        c2 = hoomd.md.nlist.Cell(buffer=0.3, exclusions=excl)
        l2 = hoomd.md.pair.LJ(nlist=c2, default_r_cut=rc, mode="shift")
        l2.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
        s2.operations.integrator = hoomd.md.Integrator(
            dt=0.001, methods=[hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())], forces=[l2])
        s2.run(0)
        assert np.array_equal(canon_pairs(c2.pair_list), want), excl
    # dynamics without the torsion (it is singular for the straight triplets these chains start from;
    # test_zigzag_chains_nve integrates it): energy is conserved through rebuilds
    sim.operations.integrator = hoomd.md.Integrator(
        dt=0.001, methods=[hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())], forces=[lj, hb, ha])
    sim.run(500)
    e = []
    for _ in range(10):
        sim.run(200)
        ke, pe, _ = sim.operations.integrator._thermo()
        e.append(ke + pe)
    assert cell.num_builds >= 1 and ke > 0.3 * n
    assert np.ptp(e) <= 2e-3 * ke


@pytest.mark.gpu
def test_zigzag_chains_nve():
    """bonds + angles + periodic dihedrals on bent (zig-zag) chains: the net force conserves energy."""
    import hoomd_blue_b200 as hoomd
    rng = np.random.default_rng(3)
    nc, ln, b0, th = 216, 8, 1.0, 1.911
    L = np.float32(6 * 9.0)
    pos, bonds, angles, dih = [], [], [], []
    for c in range(nc):
        o = (np.array([c % 6, (c // 6) % 6, c // 36]) + 0.5) * 9.0 - float(L) / 2
        for k in range(ln):
            i = c * ln + k
            pos.append(o + [k * b0 * np.sin(th / 2), (k % 2) * b0 * np.cos(th / 2), 0.0])
            if k >= 1:
                bonds.append((i - 1, i))
            if k >= 2:
                angles.append((i - 2, i - 1, i))
            if k >= 3:
                dih.append((i - 3, i - 2, i - 1, i))
    pos = np.array(pos) + rng.normal(0, 0.03, (nc * ln, 3))
    pos = (pos + float(L) / 2) % float(L) - float(L) / 2   # chains of the last column cross the box face
    pos = np.clip(pos.astype(np.float32), -L / 2, np.nextafter(L / 2, np.float32(0)))
    vel = rng.normal(0, 0.5, pos.shape).astype(np.float32)
    snap = hoomd.Snapshot.from_arrays([L, L, L, 0, 0, 0], pos, vel)
    for grp, data, name in ((snap.bonds, bonds, "b"), (snap.angles, angles, "a"), (snap.dihedrals, dih, "d")):
        grp.N = len(data)
        grp.types = [name]
        grp.group[:] = data
    sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=2)
    sim.create_state_from_snapshot(snap)
    hb = hoomd.md.bond.Harmonic()
    hb.params["b"] = dict(k=300.0, r0=1.0)
    ha = hoomd.md.angle.Harmonic()
    ha.params["a"] = dict(k=30.0, t0=th)
    hd = hoomd.md.dihedral.Periodic()
    hd.params["d"] = dict(k=3.0, d=1, n=3, phi0=0.0)
    sim.operations.integrator = hoomd.md.Integrator(
        dt=0.001, methods=[hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())], forces=[hb, ha, hd])
    sim.run(200)
    e = []
    for _ in range(10):
        sim.run(200)
        ke, pe, _ = sim.operations.integrator._thermo()
        e.append(ke + pe)
    assert hd.energy > 0 and ke > 0.05 * len(pos)
    assert np.ptp(e) <= 2e-3 * ke


@pytest.mark.gpu
def test_polymer_melt_vs_oracle(oracle):
    """Kremer-Grest-like melt: WCA pair force over a list that excludes the bonded pairs + harmonic
    bonds. Pair set bit-exact against the oracle's filtered list, per-force and net results in tolerance,
    NVE with both forces conserves energy, and the trajectory survives re-ordering at rebuilds."""
    import hoomd_blue_b200 as hoomd
    from tests.test_gpu_parity import canon_pairs, force_err, FORCE_RTOL, TOTAL_RTOL
    pos, group, L = chains(400, 16, 0.7, seed=3)
    n = len(pos)
    rc = 2.0 ** (1.0 / 6.0)

    def make(exclusions):
        snap = hoomd.Snapshot.from_arrays(list(L) + [0, 0, 0], pos, np.zeros_like(pos))
        snap.bonds.N = len(group)
        snap.bonds.types = ["polymer"]
        snap.bonds.group[:] = group
        sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=2)
        sim.always_compute_pressure = True
        sim.create_state_from_snapshot(snap)
        cell = hoomd.md.nlist.Cell(buffer=0.3, exclusions=exclusions)
        lj = hoomd.md.pair.LJ(nlist=cell, default_r_cut=rc, mode="shift")
        lj.params[("A", "A")] = dict(epsilon=1.0, sigma=0.8)
        hb = hoomd.md.bond.Harmonic()
        hb.params["polymer"] = dict(k=200.0, r0=0.9)
        sim.operations.integrator = hoomd.md.Integrator(
            dt=0.001, methods=[hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())], forces=[lj, hb])
        return sim, cell, lj, hb

    sim, cell, lj, hb = make(("bond",))
    sim.run(0)
    box = oracle.make_box([float(v) for v in L])
    types = np.zeros(n, dtype=np.uint32)
    nn, head, nl, _ = oracle.nlist(box, pos, types, [rc], 0.3, full=True)
    nn_x, head_x, nl_x = oracle.filter_bond_exclusions(nn, head, nl, group)
    assert np.array_equal(canon_pairs(cell.pair_list), oracle.pair_set(nn_x, head_x, nl_x))
    f_ref, v_ref = oracle.pair_lj(box, pos, types, nn_x, head_x, nl_x, [[1.0, 0.8]], [rc], shift_mode=1)
    assert force_err(lj.forces, f_ref[:, :3]) <= FORCE_RTOL
    assert abs(lj.energy - f_ref[:, 3].sum()) <= TOTAL_RTOL * abs(f_ref[:, 3].sum())
    b_ref, bv_ref = oracle.bond_harmonic(box, pos, group, np.zeros(len(group), np.uint32), [[200.0, 0.9]])
    assert force_err(hb.forces, b_ref[:, :3]) <= FORCE_RTOL
    assert abs(hb.energy - b_ref[:, 3].sum()) <= TOTAL_RTOL * abs(b_ref[:, 3].sum())
    assert np.allclose(hb.energies, b_ref[:, 3], rtol=1e-5, atol=1e-6)
    vs, vs_ref = hb.virials.sum(axis=0), bv_ref.sum(axis=1)
    assert np.max(np.abs(vs - vs_ref)) <= 1e-5 * np.max(np.abs(vs_ref))
    # thermodynamics see the sum of both forces
    ke, pe, w = sim.operations.integrator._thermo()
    e_ref = f_ref[:, 3].sum() + b_ref[:, 3].sum()
    assert abs(pe - e_ref) <= 1e-6 * abs(e_ref)
    w_ref = v_ref.sum(axis=1) + vs_ref
    assert np.max(np.abs(np.array(w)[[0, 3, 5]] - w_ref[[0, 3, 5]])) <= 1e-5 * np.max(np.abs(w_ref))
    # without exclusions the bonded pairs come back
    sim0, cell0, lj0, hb0 = make(())
    sim0.run(0)
    assert len(cell0.pair_list) == len(cell.pair_list) + len(group)
    # NVE with the net force, from a gentle start: energy is conserved through rebuilds
    pos, group, L = lattice_chains(32, 12, 12, 16, 1.0, seed=4)
    rng = np.random.default_rng(5)
    vel0 = rng.normal(0, 1.0, size=pos.shape).astype(np.float32)
    vel0 -= vel0.mean(axis=0)
    snap = hoomd.Snapshot.from_arrays(list(L) + [0, 0, 0], pos, vel0)
    snap.bonds.N = len(group)
    snap.bonds.types = ["polymer"]
    snap.bonds.group[:] = group
    sim = hoomd.Simulation(device=hoomd.device.GPU(), seed=2)
    sim.create_state_from_snapshot(snap)
    cell = hoomd.md.nlist.Cell(buffer=0.3)
    lj = hoomd.md.pair.LJ(nlist=cell, default_r_cut=rc, mode="shift")
    lj.params[("A", "A")] = dict(epsilon=1.0, sigma=1.0)
    hb = hoomd.md.bond.Harmonic()
    hb.params["polymer"] = dict(k=200.0, r0=0.95)
    sim.operations.integrator = hoomd.md.Integrator(
        dt=0.001, methods=[hoomd.md.methods.ConstantVolume(filter=hoomd.filter.All())], forces=[lj, hb])
    sim.run(500)
    e = []
    for _ in range(10):
        sim.run(200)
        ke, pe, _ = sim.operations.integrator._thermo()
        e.append(ke + pe)
    assert cell.num_builds >= 1 and ke > 0.5 * len(pos)
    assert np.ptp(e) <= 2e-3 * ke
    # bonds are still intact after particles were re-ordered: partner lookup goes through rtag
    s = sim.state.get_snapshot()
    d = s.particles.position[group[:, 0]].astype(np.float64) - s.particles.position[group[:, 1]]
    d -= L * np.round(d / L)
    assert np.max(np.linalg.norm(d, axis=1)) < 1.6
