"""hoomd.md.pair for this path (hoomd/md/pair/pair.py): Pair base, LJ, DPD, DPDConservative and the
pair_plugin example potential (example_plugins/pair_plugin/pair.py)."""
import ctypes as C

import numpy as np

from .. import _lib
from .. import communicator as _comm
from .._lib import check, ptr
from .typeparam import TypePairDict


def _nonneg(v):
    v = float(v)
    if v < 0:
        raise ValueError("expected a non-negative real")
    return v


class Force:
    _loggables = {"energy": "scalar"}
    """hoomd.md.force.Force properties (hoomd/md/force.py:77-156), tag-ordered."""

    _h = None
    _sim = None

    def _need(self):
        if self._h is None:
            raise RuntimeError("force is not attached (run the simulation first)")

    def _decomposed(self):
        return self._sim.device.communicator.num_ranks > 1

    def _local(self, want_virial=False):
        """(tags, forces, energies, virials) of the particles this rank owns."""
        n = self._sim.state._N_local
        tag = np.zeros(n, dtype=np.uint32)
        f = np.zeros((n, 3), dtype=np.float64)
        e = np.zeros(n, dtype=np.float64)
        v = np.zeros((n, 6), dtype=np.float64) if want_virial else None
        check(_lib.lib().hb2_pair_get_local(self._h, ptr(tag, C.c_uint32), ptr(f, C.c_double),
                                            ptr(e, C.c_double),
                                            ptr(v, C.c_double) if want_virial else None))
        return tag, f, e, v

    @property
    def forces(self):
        self._need()
        if self._decomposed():
            tag, f, _, _ = self._local()
            return _comm.gather_by_tag(self._sim.state.N_particles, tag, f)
        out = np.zeros((self._sim.state.N_particles, 3), dtype=np.float64)
        check(_lib.lib().hb2_pair_get_forces(self._h, ptr(out, C.c_double)))
        return out

    @property
    def energies(self):
        self._need()
        if self._decomposed():
            tag, _, e, _ = self._local()
            return _comm.gather_by_tag(self._sim.state.N_particles, tag, e)
        out = np.zeros(self._sim.state.N_particles, dtype=np.float64)
        check(_lib.lib().hb2_pair_get_energies(self._h, ptr(out, C.c_double)))
        return out

    @property
    def virials(self):
        """(N, 6) or None when the pressure flag is off (hoomd/ForceCompute.cc:491-496)."""
        self._need()
        if not self._sim.always_compute_pressure:
            return None
        if self._decomposed():
            tag, _, _, v = self._local(want_virial=True)
            return _comm.gather_by_tag(self._sim.state.N_particles, tag, v)
        out = np.zeros((self._sim.state.N_particles, 6), dtype=np.float64)
        check(_lib.lib().hb2_pair_get_virials(self._h, ptr(out, C.c_double)))
        return out

    @property
    def energy(self):
        self._need()
        e = C.c_double(0)
        check(_lib.lib().hb2_pair_get_energy(self._h, C.byref(e)))
        return e.value

    @property
    def additional_energy(self):
        self._need()
        e = C.c_double(0)
        v = (C.c_double * 6)()
        check(_lib.lib().hb2_pair_get_additional(self._h, C.byref(e), v))
        return e.value

    @property
    def additional_virial(self):
        self._need()
        e = C.c_double(0)
        v = (C.c_double * 6)()
        check(_lib.lib().hb2_pair_get_additional(self._h, C.byref(e), v))
        return np.array(list(v))


class Pair(Force):
    """hoomd.md.pair.Pair(nlist, default_r_cut=None, default_r_on=0., mode='none')."""

    _accepted_modes = ("none", "shift", "xplor")
    _evaluator = None  # name in _lib.EVAL
    _param_keys = ()

    def __init__(self, nlist, default_r_cut=None, default_r_on=0.0, mode="none"):
        if mode not in self._accepted_modes:
            raise ValueError(f"mode must be one of {self._accepted_modes}")
        self.nlist = nlist
        self.mode = mode
        self.r_cut = TypePairDict("r_cut", _nonneg, None if default_r_cut is None else default_r_cut)
        self.r_on = TypePairDict("r_on", _nonneg, default_r_on)
        self.params = TypePairDict("params", self._validate_params)
        self.r_cut._on_change = self.r_on._on_change = self.params._on_change = self._sync
        self._tpp = None

    def _validate_params(self, v):
        missing = [k for k in self._param_keys if k not in v]
        if missing:
            raise ValueError(f"missing parameters {missing}")
        return {k: float(v[k]) for k in self._param_keys}

    # --- attach: create the C++ object and push every type-pair parameter
    # (hoomd/md/pair/pair.py:126-144)
    def _attach(self, sim):
        if self._h is not None:
            return
        self.nlist._attach(sim)
        h = C.c_void_p()
        check(_lib.lib().hb2_pair_create(C.byref(h), sim.state._h, self.nlist._h,
                                         _lib.EVAL[self._evaluator], _lib.SHIFT[self.mode]))
        self._h = h
        self._sim = sim
        self._sync()
        self._attach_extra()

    def _attach_extra(self):
        pass

    def _detach(self):
        if self._h is not None:
            _lib.lib().hb2_pair_destroy(self._h)
            self._h = None
            self._sim = None

    def _sync(self):
        if self._h is None:
            return
        L = _lib.lib()
        types = self._sim.state.particle_types
        for i, a in enumerate(types):
            for j, b in enumerate(types):
                if j < i:
                    continue
                p = self.params[(a, b)]
                arr = (C.c_double * len(self._param_keys))(*[p[k] for k in self._param_keys])
                check(L.hb2_pair_set_params(self._h, i, j, arr, len(self._param_keys)))
                check(L.hb2_pair_set_r_cut(self._h, i, j, self.r_cut[(a, b)]))
                check(L.hb2_pair_set_r_on(self._h, i, j, self.r_on[(a, b)]))
        if self._tpp:
            check(L.hb2_pair_set_threads_per_particle(self._h, self._tpp))

    @property
    def kernel_parameters(self):
        """Autotuner knob of the reference (block size, threads per particle); here tpp only."""
        return {"threads_per_particle": self._tpp or 8}

    def set_threads_per_particle(self, tpp):
        self._tpp = int(tpp)
        if self._h is not None:
            check(_lib.lib().hb2_pair_set_threads_per_particle(self._h, self._tpp))


class LJ(Pair):
    """hoomd.md.pair.LJ(nlist, default_r_cut=None, default_r_on=0., mode='none',
    tail_correction=False); params {epsilon, sigma} (hoomd/md/pair/pair.py:163-260)."""
    _evaluator = "lj"
    _param_keys = ("epsilon", "sigma")

    def __init__(self, nlist, default_r_cut=None, default_r_on=0.0, mode="none",
                 tail_correction=False):
        super().__init__(nlist, default_r_cut, default_r_on, mode)
        self.tail_correction = bool(tail_correction)

    def _attach_extra(self):
        check(_lib.lib().hb2_pair_set_tail_correction(self._h, int(self.tail_correction)))


class DPD(Pair):
    """hoomd.md.pair.DPD(nlist, kT, default_r_cut=None); params {A, gamma}
    (hoomd/md/pair/pair.py:676-779). mode is fixed to 'none'."""
    _accepted_modes = ("none",)
    _evaluator = "dpd"
    _param_keys = ("A", "gamma")

    def __init__(self, nlist, kT, default_r_cut=None):
        super().__init__(nlist, default_r_cut, 0.0, "none")
        self.kT = float(kT)

    def _attach_extra(self):
        check(_lib.lib().hb2_pair_set_kT(self._h, self.kT))


class DPDConservative(Pair):
    """hoomd.md.pair.DPDConservative(nlist, default_r_cut=None); params {A}."""
    _accepted_modes = ("none",)
    _evaluator = "dpd_conservative"
    _param_keys = ("A",)

    def __init__(self, nlist, default_r_cut=None):
        super().__init__(nlist, default_r_cut, 0.0, "none")


class ExamplePair(Pair):
    """The pair_plugin potential (example_plugins/pair_plugin/pair.py): params {k, sigma},
    modes none/shift/xplor."""
    _evaluator = "example_pair"
    _param_keys = ("k", "sigma")


# ---- further potentials through the same evaluator template (hoomd/md/pair/pair.py)
class Gaussian(Pair):
    """hoomd.md.pair.Gaussian: V = epsilon exp(-r^2 / (2 sigma^2)); params {epsilon, sigma}
    (hoomd/md/pair/pair.py Gaussian, hoomd/md/EvaluatorPairGauss.h)."""
    _evaluator = "gauss"
    _param_keys = ("epsilon", "sigma")


Gauss = Gaussian  # the pre-v4 name


class Yukawa(Pair):
    """hoomd.md.pair.Yukawa: V = epsilon exp(-kappa r) / r; params {epsilon, kappa}."""
    _evaluator = "yukawa"
    _param_keys = ("epsilon", "kappa")


class LJ1208(Pair):
    """hoomd.md.pair.LJ1208: V = 4 epsilon [(sigma/r)^12 - (sigma/r)^8]; params {epsilon, sigma}."""
    _evaluator = "lj1208"
    _param_keys = ("epsilon", "sigma")


class LJ0804(Pair):
    """hoomd.md.pair.LJ0804: V = 4 epsilon [(sigma/r)^8 - (sigma/r)^4]; params {epsilon, sigma}."""
    _evaluator = "lj0804"
    _param_keys = ("epsilon", "sigma")


class ForceShiftedLJ(Pair):
    """hoomd.md.pair.ForceShiftedLJ: LJ with the force shifted to zero at r_cut; params {epsilon,
    sigma}; modes none / shift."""
    _accepted_modes = ("none", "shift")
    _evaluator = "force_shifted_lj"
    _param_keys = ("epsilon", "sigma")


class Mie(Pair):
    """hoomd.md.pair.Mie: the n-m potential; params {epsilon, sigma, n, m}."""
    _evaluator = "mie"
    _param_keys = ("epsilon", "sigma", "n", "m")


class Morse(Pair):
    """hoomd.md.pair.Morse: V = D0 [exp(-2 alpha (r - r0)) - 2 exp(-alpha (r - r0))]; params {D0, alpha,
    r0}."""
    _evaluator = "morse"
    _param_keys = ("D0", "alpha", "r0")


class Buckingham(Pair):
    """hoomd.md.pair.Buckingham: V = A exp(-r / rho) - C / r^6; params {A, rho, C}."""
    _evaluator = "buckingham"
    _param_keys = ("A", "rho", "C")


class ExpandedLJ(Pair):
    """hoomd.md.pair.ExpandedLJ: LJ in r - delta; params {epsilon, sigma, delta}."""
    _evaluator = "expanded_lj"
    _param_keys = ("epsilon", "sigma", "delta")


class ExpandedGaussian(Pair):
    """hoomd.md.pair.ExpandedGaussian: V = eps exp(-(r - delta)^2 / (2 sigma^2)); params {epsilon, sigma,
    delta}."""
    _evaluator = "expanded_gaussian"
    _param_keys = ("epsilon", "sigma", "delta")


class LJGauss(Pair):
    """hoomd.md.pair.LJGauss: V = 1/r^12 - 2/r^6 - eps exp(-(r - r0)^2 / (2 sigma^2)); params {epsilon,
    sigma, r0}."""
    _evaluator = "lj_gauss"
    _param_keys = ("epsilon", "sigma", "r0")


class TWF(Pair):
    """hoomd.md.pair.TWF (ten Wolde-Frenkel); params {epsilon, sigma, alpha}. Overlapping pairs
    (r <= sigma) have infinite energy and force, as in the reference."""
    _evaluator = "twf"
    _param_keys = ("epsilon", "sigma", "alpha")


class DPDLJ(Pair):
    """hoomd.md.pair.DPDLJ(nlist, kT, default_r_cut=None, mode='none'): LJ conservative force with the
    DPD thermostat; params {epsilon, sigma, gamma} (hoomd/md/pair/pair.py:837-935); modes none / shift."""
    _accepted_modes = ("none", "shift")
    _evaluator = "dpdlj"
    _param_keys = ("epsilon", "sigma", "gamma")

    def __init__(self, nlist, kT, default_r_cut=None, mode="none"):
        super().__init__(nlist, default_r_cut, 0.0, mode)
        self.kT = float(kT)

    def _attach_extra(self):
        check(_lib.lib().hb2_pair_set_kT(self._h, self.kT))


class ExpandedMie(Pair):
    """hoomd.md.pair.ExpandedMie: Mie n-m potential in r - delta; params {epsilon, sigma, n, m, delta}
    (hoomd/md/pair/pair.py:1222-1300)."""
    _evaluator = "expanded_mie"
    _param_keys = ("epsilon", "sigma", "n", "m", "delta")


class OPP(Pair):
    """hoomd.md.pair.OPP: V = C1 r^-eta1 + C2 r^-eta2 cos(k r - phi); params {C1, C2, eta1, eta2, k, phi}
    (hoomd/md/pair/pair.py:1690-1760)."""
    _evaluator = "opp"
    _param_keys = ("C1", "C2", "eta1", "eta2", "k", "phi")


class Fourier(Pair):
    """hoomd.md.pair.Fourier: V = r^-12 + r^-2 sum_n [a_n cos(n pi r / r_cut) + b_n sin(n pi r / r_cut)];
    params dict(a=[a2, a3, a4], b=[b2, b3, b4]); modes 'none' and 'xplor' (hoomd/md/pair/pair.py:1614-1680)."""
    _accepted_modes = ("none", "xplor")
    _evaluator = "fourier"
    _param_keys = ("a2", "a3", "a4", "b2", "b3", "b4")

    def _validate_params(self, v):
        if "a" not in v or "b" not in v:
            raise ValueError("missing parameters: a=[a2, a3, a4], b=[b2, b3, b4]")
        a, b = [float(x) for x in v["a"]], [float(x) for x in v["b"]]
        if len(a) != 3 or len(b) != 3:
            raise ValueError("a and b have three elements each")
        return dict(a=tuple(a), b=tuple(b), a2=a[0], a3=a[1], a4=a[2], b2=b[0], b3=b[1], b4=b[2])


class Table(Pair):
    """hoomd.md.pair.Table(nlist, default_r_cut=None): tabulated potential, params {r_min, U, F} with
    U and F = -dU/dr sampled at r_min + i (r_cut - r_min) / len(U), linearly interpolated
    (hoomd/md/pair/pair.py Table, hoomd/md/EvaluatorPairTable.h). No energy shift modes."""
    _accepted_modes = ("none",)
    _evaluator = "table"
    _param_keys = ("r_min", "U", "F")

    def __init__(self, nlist, default_r_cut=None):
        super().__init__(nlist, default_r_cut, 0.0, "none")

    def _validate_params(self, v):
        missing = [k for k in self._param_keys if k not in v]
        if missing:
            raise ValueError(f"missing parameters {missing}")
        U = np.ascontiguousarray(v["U"], dtype=np.float64).reshape(-1)
        F = np.ascontiguousarray(v["F"], dtype=np.float64).reshape(-1)
        if len(U) != len(F):
            raise ValueError("The length of V and F arrays must be equal.")
        if len(U) == 0:
            raise ValueError("The length of V and F must not be zero.")
        return dict(r_min=float(v["r_min"]), U=U, F=F)

    def _sync(self):
        if self._h is None:
            return
        L = _lib.lib()
        types = self._sim.state.particle_types
        for i, a in enumerate(types):
            for j, b in enumerate(types):
                if j < i:
                    continue
                p = self.params[(a, b)]
                check(L.hb2_pair_set_table(self._h, i, j, p["r_min"], len(p["U"]),
                                           _lib.ptr(p["U"], C.c_double), _lib.ptr(p["F"], C.c_double)))
                check(L.hb2_pair_set_r_cut(self._h, i, j, self.r_cut[(a, b)]))
