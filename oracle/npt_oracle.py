"""CPU restatement (numpy, fp64) of the barostat of hoomd.md.methods.ConstantPressure for the orthorhombic
box degrees of freedom. TEST INFRASTRUCTURE: imported by tests/ only, never by the product.

Follows hoomd/md/TwoStepConstantPressure.cc (HOOMD-blue 4.8.2): advanceBarostat :860-1000 (noise-free part and
the Langevin piston's decay), updatePropagator :716-858 (diagonal elements), integrateStepOne :213-262 (box
update), getBarostatEnergy :693-714. Pinned by tests/test_oracle_golden.py::test_npt_barostat_replay against
the reference-generated trajectory tests/golden/npt_2k.npz (case a: nu and box lengths to round-off).
"""
import numpy as np

F_COEFF = [1.0, 1.0 / 6.0, 1.0 / 120.0, 1.0 / 5040.0, 1.0 / 362880.0, 1.0 / 39916800.0]


def barostat_mass(ndof, tauS, T=1.0, d=3):
    return (ndof + d) / d * T * tauS * tauS


def relevant_couple(couple, flags):
    """couple in {'none','xy','xz','yz','xyz'}, flags = (x, y, z) booleans (getRelevantCouplings :128-170)"""
    x, y, z = flags
    if not x:
        couple = {"xyz": "yz", "xy": "none", "xz": "none"}.get(couple, couple)
    if not y:
        couple = {"xyz": "xz", "yz": "none", "xy": "none"}.get(couple, couple)
    if not z:
        couple = {"xyz": "xy", "yz": "none", "xz": "none"}.get(couple, couple)
    return couple


def advance_barostat(nu, P_diag3, ke, V, S3, ndof, tauS, dt, couple="xyz", flags=(True, True, True), T=1.0,
                     gamma=0.0, noise3=(0.0, 0.0, 0.0)):
    """half a barostat step: nu = (nu_xx, nu_yy, nu_zz); P_diag3 the diagonal of the pressure tensor"""
    W = barostat_mass(ndof, tauS, T)
    mtk_term = 2.0 * ke * (0.5 * dt / ndof / W)
    Pxx, Pyy, Pzz = P_diag3
    c = relevant_couple(couple, flags)
    z = list(noise3)
    if c == "none":
        Pd, Rd = [Pxx, Pyy, Pzz], [z[0], z[1], z[2]]
    elif c == "xy":
        a = 0.5 * (Pxx + Pyy)
        Pd, Rd = [a, a, Pzz], [z[0], z[0], z[1]]
    elif c == "xz":
        a = 0.5 * (Pxx + Pzz)
        Pd, Rd = [a, Pyy, a], [z[0], z[1], z[0]]
    elif c == "yz":
        a = 0.5 * (Pyy + Pzz)
        Pd, Rd = [Pxx, a, a], [z[0], z[1], z[1]]
    else:
        a = (1.0 / 3.0) * (Pxx + Pyy + Pzz)
        Pd, Rd = [a, a, a], [z[0], z[0], z[0]]
    e = np.exp(-gamma * dt / 2.0)
    coeff = np.sqrt(T * (1.0 - e * e) / W)
    out = list(nu)
    for d in range(3):
        if flags[d]:
            out[d] = out[d] * e + coeff * Rd[d]
            out[d] += 0.5 * dt * V / W * (Pd[d] - S3[d]) + mtk_term
    return np.array(out)


def propagator(nu, dt):
    """diagonals of exp_v, exp_r, exp_r_int"""
    nu = np.asarray(nu, dtype=np.float64)
    arg_r = 0.5 * nu * dt
    f_r = np.zeros(3)
    term = np.ones(3)
    for i in range(6):
        f_r += F_COEFF[i] * term
        term = term * arg_r * arg_r
    exp_v = np.exp(2.0 * (-0.25 * nu) * dt)
    exp_r = np.exp(nu * dt)
    exp_r_int = dt * np.exp(0.5 * nu * dt) * f_r
    return exp_v, exp_r, exp_r_int


def mtk_factor(nu, ndof, dt):
    return np.exp(-0.5 * dt * float(np.sum(nu)) / ndof)


def barostat_energy(nu6, ndof, tauS, T=1.0):
    return barostat_mass(ndof, tauS, T) * float(np.sum(np.asarray(nu6) ** 2)) / 2.0
