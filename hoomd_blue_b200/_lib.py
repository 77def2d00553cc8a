"""ctypes binding of libhoomd_b200.so (include/hoomd_b200.h). There is no fallback: if the CUDA
library is missing or no device is visible the operators raise."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, os.environ.get("HB2_LIB", "libhoomd_b200.so"))

# every symbol include/hoomd_b200.h declares (tests check the exports against the header)
(OK, ERR_INVALID, ERR_CUDA, ERR_NAN_POSITION, ERR_OUT_OF_BOX, ERR_BOX_TOO_SMALL, ERR_NLIST_OVERFLOW,
 ERR_NO_DEVICE, ERR_COMM) = range(9)
EVAL = {"lj": 0, "dpd": 1, "dpd_conservative": 2, "example_pair": 3, "gauss": 4, "yukawa": 5,
        "lj1208": 6, "lj0804": 7, "force_shifted_lj": 8, "mie": 9, "morse": 10,
        "buckingham": 11, "expanded_lj": 12, "expanded_gaussian": 13, "lj_gauss": 14, "twf": 15, "dpdlj": 16,
        "table": 17, "expanded_mie": 18, "opp": 19, "fourier": 20}
SHIFT = {"none": 0, "shift": 1, "xplor": 2}
PARAM_FLOATS = 8  # HB2_PARAM_FLOATS (include/hoomd_b200.h)


class Box(C.Structure):
    _fields_ = [("L", C.c_double * 3), ("lo", C.c_double * 3), ("hi", C.c_double * 3),
                ("Linv", C.c_double * 3), ("Lf", C.c_float * 3), ("lof", C.c_float * 3),
                ("hif", C.c_float * 3), ("Linvf", C.c_float * 3), ("periodic", C.c_int32 * 3),
                ("_pad", C.c_int32)]


class Hb2Error(RuntimeError):

    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


_lib = None

_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_vp = C.c_void_p


def _declare(L):
    L.hb2_last_error.restype = C.c_char_p
    L.hb2_abi_version.restype = C.c_int
    L.hb2_launch_count.restype = C.c_uint64
    L.hb2_box_make.argtypes = [C.POINTER(Box), C.c_double, C.c_double, C.c_double]
    L.hb2_pack_params.argtypes = [C.c_int, _f64p, C.c_uint32, _f32p]
    # layer 1
    L.hb2_gpu_compute_cell_order.argtypes = [_vp, C.c_uint32, C.POINTER(Box), _u32p, _vp, _vp, _vp,
                                             _vp, _vp, C.POINTER(C.c_size_t), _vp]
    L.hb2_gpu_compute_nlist_binned.argtypes = [_vp, C.c_uint32, C.c_uint32, C.POINTER(Box), _u32p,
                                               _vp, _vp, _vp, C.c_uint32, _vp, C.c_uint32, _vp, _vp,
                                               _vp, _vp]
    L.hb2_gpu_compute_pair_forces.argtypes = [C.c_int, C.c_int, _vp, _vp, _vp, C.c_uint32,
                                              C.POINTER(Box), _vp, _vp, C.c_uint32, _vp, _vp, _vp,
                                              C.c_uint32, _vp, _vp, C.c_size_t, C.c_uint32,
                                              C.c_uint64, C.c_float, C.c_float, C.c_float, C.c_uint32, _vp]
    L.hb2_gpu_nve_step_one.argtypes = [_vp, _vp, _vp, _vp, C.c_uint32, C.POINTER(Box), C.c_float,
                                       _vp, _vp, C.c_float, _vp, _vp]
    L.hb2_gpu_nve_step_two.argtypes = [_vp, _vp, C.c_uint32, C.c_float, _vp, _vp, _vp]
    # layer 2
    L.hb2_system_create.argtypes = [C.POINTER(_vp), C.c_int, C.c_uint32, C.c_uint32]
    L.hb2_system_destroy.argtypes = [_vp]
    L.hb2_system_destroy.restype = None
    L.hb2_system_set_box.argtypes = [_vp] + [C.c_double] * 6
    L.hb2_system_set_particles.argtypes = [_vp, C.c_uint32, _f32p, _f32p, _f32p, _u32p, _i32p]
    L.hb2_system_get_particles.argtypes = [_vp, _f32p, _f32p, _f32p, _u32p, _i32p]
    L.hb2_system_n.argtypes = [_vp]
    L.hb2_system_n.restype = C.c_uint32
    L.hb2_comm_unique_id.argtypes = [_vp]
    L.hb2_system_set_domain.argtypes = [_vp, _u32p, C.c_uint32, C.c_uint32, _vp]
    L.hb2_domain_find_decomposition.argtypes = [C.c_uint32, C.c_double, C.c_double, C.c_double, _u32p]
    L.hb2_domain_place.argtypes = [C.POINTER(Box), _u32p, _f32p, C.c_uint32, _u32p]
    L.hb2_domain_ghost_mask.argtypes = [C.POINTER(Box), _u32p, _u32p, C.c_double, _f32p, C.c_uint32,
                                        C.POINTER(C.c_uint8)]
    L.hb2_system_set_particles_tagged.argtypes = [_vp, C.c_uint32, C.c_uint64, _u32p, _f32p, _f32p,
                                                  _f32p, _u32p, _i32p]
    L.hb2_system_get_local.argtypes = [_vp, _u32p, _f32p, _f32p, _f32p, _u32p, _i32p]
    L.hb2_system_n_global.argtypes = [_vp]
    L.hb2_system_n_global.restype = C.c_uint64
    L.hb2_system_n_ghost.argtypes = [_vp]
    L.hb2_system_n_ghost.restype = C.c_uint32
    L.hb2_pair_get_local.argtypes = [_vp, _u32p, _f64p, _f64p, _f64p]
    L.hb2_system_timestep.argtypes = [_vp]
    L.hb2_system_timestep.restype = C.c_uint64
    L.hb2_system_set_timestep.argtypes = [_vp, C.c_uint64]
    L.hb2_system_set_compute_virial.argtypes = [_vp, C.c_int]
    L.hb2_system_synchronize.argtypes = [_vp]
    L.hb2_nlist_create.argtypes = [C.POINTER(_vp), _vp, C.c_double, C.c_uint32, C.c_int]
    L.hb2_nlist_destroy.argtypes = [_vp]
    L.hb2_nlist_destroy.restype = None
    L.hb2_nlist_set_buffer.argtypes = [_vp, C.c_double]
    L.hb2_nlist_compute.argtypes = [_vp]
    L.hb2_nlist_num_builds.argtypes = [_vp]
    L.hb2_nlist_set_r_cut.argtypes = [_vp, C.c_uint32, C.c_uint32, C.c_double]
    L.hb2_nlist_device_arrays.argtypes = [_vp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.hb2_nlist_num_builds.restype = C.c_uint64
    L.hb2_nlist_shortest_rebuild.argtypes = [_vp]
    L.hb2_nlist_shortest_rebuild.restype = C.c_uint64
    L.hb2_nlist_dimensions.argtypes = [_vp, _u32p]
    L.hb2_nlist_pair_count.argtypes = [_vp, C.POINTER(C.c_uint64)]
    L.hb2_nlist_pair_list.argtypes = [_vp, _u32p, C.c_uint64, C.POINTER(C.c_uint64)]
    L.hb2_nlist_stats.argtypes = [_vp, _f64p, _u32p, _u32p]
    L.hb2_pair_create.argtypes = [C.POINTER(_vp), _vp, _vp, C.c_int, C.c_int]
    L.hb2_pair_destroy.argtypes = [_vp]
    L.hb2_pair_destroy.restype = None
    L.hb2_pair_set_params.argtypes = [_vp, C.c_uint32, C.c_uint32, _f64p, C.c_uint32]
    L.hb2_pair_set_r_cut.argtypes = [_vp, C.c_uint32, C.c_uint32, C.c_double]
    L.hb2_pair_set_table.argtypes = [_vp, C.c_uint32, C.c_uint32, C.c_double, C.c_uint32, _f64p, _f64p]
    L.hb2_pair_set_r_on.argtypes = [_vp, C.c_uint32, C.c_uint32, C.c_double]
    L.hb2_pair_set_kT.argtypes = [_vp, C.c_double]
    L.hb2_pair_set_threads_per_particle.argtypes = [_vp, C.c_uint32]
    L.hb2_pair_compute.argtypes = [_vp, C.c_uint64]
    L.hb2_pair_get_forces.argtypes = [_vp, _f64p]
    L.hb2_pair_get_energies.argtypes = [_vp, _f64p]
    L.hb2_pair_get_virials.argtypes = [_vp, _f64p]
    L.hb2_pair_get_energy.argtypes = [_vp, _f64p]
    L.hb2_pair_set_tail_correction.argtypes = [_vp, C.c_int]
    L.hb2_pair_get_additional.argtypes = [_vp, _f64p, _f64p]
    L.hb2_integrator_create.argtypes = [C.POINTER(_vp), _vp, C.c_double]
    L.hb2_integrator_destroy.argtypes = [_vp]
    L.hb2_integrator_destroy.restype = None
    L.hb2_integrator_add_force.argtypes = [_vp, _vp]
    L.hb2_integrator_set_reset_stats.argtypes = [_vp, C.c_int]
    L.hb2_integrator_set_constant_volume.argtypes = [_vp, C.c_int, C.c_double, C.c_double]
    L.hb2_system_set_bonds.argtypes = [_vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32]
    L.hb2_system_set_angles.argtypes = [_vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32]
    L.hb2_system_set_dihedrals.argtypes = [_vp, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32]
    L.hb2_nlist_set_exclusions.argtypes = [_vp, C.c_int]
    L.hb2_nlist_set_exclusion_pairs.argtypes = [_vp, C.c_uint32, C.POINTER(C.c_uint32)]
    L.hb2_bond_create.argtypes = [C.POINTER(_vp), _vp, C.c_int]
    L.hb2_bond_destroy.argtypes = [_vp]
    L.hb2_bond_destroy.restype = None
    L.hb2_bond_set_params.argtypes = [_vp, C.c_uint32, _f64p, C.c_uint32]
    L.hb2_bond_get.argtypes = [_vp, _f64p, _f64p, _f64p, _f64p]
    L.hb2_integrator_add_bond.argtypes = [_vp, _vp]
    L.hb2_integrator_set_langevin.argtypes = [_vp, C.c_double, _f64p, C.c_uint32]
    L.hb2_integrator_set_constant_pressure.argtypes = [_vp, _f64p, C.c_double, C.c_int, C.POINTER(C.c_int),
                                                       C.c_int, C.c_double, C.c_int, C.c_double, C.c_double]
    L.hb2_integrator_get_barostat_dof.argtypes = [_vp, _f64p]
    L.hb2_integrator_set_barostat_dof.argtypes = [_vp, _f64p]
    L.hb2_integrator_thermalize_barostat_dof.argtypes = [_vp, C.c_uint64]
    L.hb2_integrator_barostat_energy.argtypes = [_vp, _f64p]
    L.hb2_system_get_box.argtypes = [_vp, _f64p]
    L.hb2_domain_ipc_handle.argtypes = [_vp, C.c_void_p]
    L.hb2_domain_set_peers.argtypes = [_vp, C.c_void_p, C.c_uint32]
    L.hb2_domain_p2p_active.argtypes = [_vp]
    L.hb2_integrator_kinetic_tensor.argtypes = [_vp, _f64p]
    L.hb2_integrator_set_filter_tags.argtypes = [_vp, C.POINTER(C.c_uint32), C.c_uint64]
    L.hb2_integrator_kinetic_of_tags.argtypes = [_vp, C.POINTER(C.c_uint32), C.c_uint64, C.POINTER(C.c_double)]
    L.hb2_integrator_run.argtypes = [_vp, C.c_uint64]
    L.hb2_integrator_thermo.argtypes = [_vp, _f64p, _f64p, _f64p]
    L.hb2_integrator_last_run_ms.argtypes = [_vp, _f32p]
    L.hb2_integrator_set_profiling.argtypes = [_vp, C.c_int]
    L.hb2_integrator_get_profile.argtypes = [_vp, _f64p, C.POINTER(C.c_uint64)]


def lib():
    """The loaded library. Raises if it was not built (python -m hoomd_blue_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m hoomd_blue_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)")
        L = C.CDLL(LIB_PATH)
        _declare(L)
        if L.hb2_abi_version() != 2:
            raise ImportError("libhoomd_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(status):
    if status != OK:
        raise Hb2Error(status, lib().hb2_last_error().decode())


def ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))
