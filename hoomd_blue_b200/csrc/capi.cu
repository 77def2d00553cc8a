// capi.cu — layer 1 of the C ABI: stateless kernel drivers over caller-owned device arrays
// (include/hoomd_b200.h). These are what a HOOMD `...GPU` class calls in place of the reference's
// gpu_compute_cell_list / gpu_compute_nlist_binned / gpu_compute_pair_forces<evaluator> /
// gpu_nvt_rescale_step_one/two drivers; the engine (engine.cu) calls the same launchers.
#include "common.cuh"
#include "kernels.cuh"

#include <cmath>
#include <vector>

using namespace hb2;

extern "C" hb2_status hb2_gpu_compute_cell_order(const float* d_pos4, uint32_t N, const hb2_box* box,
                                                 const uint32_t dim[3], uint32_t* d_order,
                                                 uint32_t* d_cell_start, uint32_t* d_cell_end,
                                                 uint32_t* d_conditions, void* d_tmp,
                                                 size_t* tmp_bytes, void* stream)
    {
    if (!box || !dim || !tmp_bytes)
        {
        set_error("hb2_gpu_compute_cell_order: null argument");
        return HB2_ERR_INVALID;
        }
    const uint64_t Nc = (uint64_t)dim[0] * dim[1] * dim[2];
    if (Nc == 0 || Nc >= (1ull << 31))
        {
        set_error("hb2_gpu_compute_cell_order: bad cell grid");
        return HB2_ERR_INVALID;
        }
    if (!d_tmp)
        {
        *tmp_bytes = cell_order_tmp_bytes(N, (uint32_t)Nc);
        return HB2_OK;
        }
    if (!d_pos4 || !d_order || !d_cell_start || !d_cell_end || !d_conditions)
        {
        set_error("hb2_gpu_compute_cell_order: null device pointer");
        return HB2_ERR_INVALID;
        }
    return compute_cell_order((const float4*)d_pos4, N, N, *box, periodic_geom(dim), d_order,
                              d_cell_start, d_cell_end, d_conditions, d_tmp, *tmp_bytes,
                              (cudaStream_t)stream, 0);
    }

extern "C" hb2_status hb2_gpu_compute_nlist_binned(float* d_pos4, uint32_t N, uint32_t N_total,
                                                   const hb2_box* box, const uint32_t dim[3],
                                                   const uint32_t* d_cell_start,
                                                   const uint32_t* d_cell_end,
                                                   const double* d_r_listsq, uint32_t ntypes,
                                                   uint32_t* d_nlist, uint32_t pitch,
                                                   uint32_t* d_n_neigh, float* d_last_pos4,
                                                   uint32_t* d_conditions, void* stream)
    {
    if (!d_pos4 || !box || !dim || !d_cell_start || !d_cell_end || !d_r_listsq || !d_nlist || !d_n_neigh
        || !d_conditions || ntypes == 0)
        {
        set_error("hb2_gpu_compute_nlist_binned: null argument");
        return HB2_ERR_INVALID;
        }
    if (pitch % 32u != 0)
        {
        set_error("hb2_gpu_compute_nlist_binned: pitch must be a multiple of 32");
        return HB2_ERR_INVALID;
        }
    // the far-away dummy particle that pads the rows lives at index N_total
    const float4 far_pos = make_float4(1.0e18f, 1.0e18f, 1.0e18f, 0.f);
    HB2_CUDA(cudaMemcpyAsync((float4*)d_pos4 + N_total, &far_pos, sizeof(float4), cudaMemcpyHostToDevice,
                             (cudaStream_t)stream));
    // the guard band and the one-type fast path need the host copy of the r_list^2 matrix
    std::vector<double> h((size_t)ntypes * ntypes);
    HB2_CUDA(cudaMemcpyAsync(h.data(), d_r_listsq, sizeof(double) * h.size(), cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
    HB2_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    double rl_min = 0.0;
    for (double v : h)
        if (v > 0.0 && (rl_min == 0.0 || std::sqrt(v) < rl_min))
            rl_min = std::sqrt(v);
    NlistBuildArgs a;
    a.pos = (const float4*)d_pos4;
    a.N = N;
    a.geom = periodic_geom(dim);
    a.cell_start = d_cell_start;
    a.cell_end = d_cell_end;
    a.r_listsq = d_r_listsq;
    a.ntypes = ntypes;
    a.r_listsq_single = h[0];
    a.r_list_min = rl_min;
    a.Lmax = (float)std::max(box->L[0], std::max(box->L[1], box->L[2]));
    a.nlist = d_nlist;
    a.pitch = pitch;
    a.n_neigh = d_n_neigh;
    a.last_pos = (float4*)d_last_pos4;
    a.conditions = d_conditions;
    a.dummy_index = N_total;
    return launch_nlist_build(a, *box, (cudaStream_t)stream);
    }

extern "C" hb2_status hb2_gpu_compute_pair_forces(hb2_evaluator ev, hb2_shift_mode mode,
                                                  const float* d_pos4, const float* d_vel4,
                                                  const uint32_t* d_tag, uint32_t N,
                                                  const hb2_box* box, const uint32_t* d_n_neigh,
                                                  const uint32_t* d_nlist, uint32_t pitch,
                                                  const float* d_params, const float* d_rcutsq,
                                                  const float* d_ronsq, uint32_t ntypes,
                                                  float* d_force4, float* d_virial,
                                                  size_t virial_pitch, uint32_t seed,
                                                  uint64_t timestep, float dt, float kT,
                                                  float face_margin,
                                                  uint32_t threads_per_particle, void* stream)
    {
    if (!d_pos4 || !box || !d_n_neigh || !d_nlist || !d_params || !d_rcutsq || !d_ronsq || !d_force4
        || ntypes == 0)
        {
        set_error("hb2_gpu_compute_pair_forces: null argument");
        return HB2_ERR_INVALID;
        }
    PairArgs a;
    a.pos = (const float4*)d_pos4;
    a.vel = (const float4*)d_vel4;
    a.tag = d_tag;
    a.N = N;
    a.n_neigh = d_n_neigh;
    a.nlist = d_nlist;
    a.pitch = pitch;
    a.params = d_params;
    a.rcutsq = d_rcutsq;
    a.ronsq = d_ronsq;
    float h[HB2_PARAM_FLOATS + 2];
    HB2_CUDA(cudaMemcpyAsync(h, d_params, HB2_PARAM_FLOATS * sizeof(float), cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
    HB2_CUDA(cudaMemcpyAsync(h + HB2_PARAM_FLOATS, d_rcutsq, sizeof(float), cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
    HB2_CUDA(cudaMemcpyAsync(h + HB2_PARAM_FLOATS + 1, d_ronsq, sizeof(float), cudaMemcpyDeviceToHost,
                             (cudaStream_t)stream));
    HB2_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    for (int q = 0; q < HB2_PARAM_FLOATS; q++)
        a.params0[q] = h[q];
    a.rcutsq0 = h[HB2_PARAM_FLOATS];
    a.ronsq0 = h[HB2_PARAM_FLOATS + 1];
    a.ntypes = ntypes;
    a.force = (float4*)d_force4;
    a.virial = d_virial;
    a.virial_pitch = virial_pitch;
    a.seed = seed;
    a.d_timestep = nullptr;
    a.timestep = timestep;
    a.dt = dt;
    a.kT = kT;
    a.face_margin = face_margin;
    a.tpp = threads_per_particle ? threads_per_particle : 8;
    return launch_pair_forces(ev, mode, a, *box, (cudaStream_t)stream);
    }

extern "C" hb2_status hb2_gpu_nve_step_one(float* d_pos4, float* d_vel4, const float* d_force4,
                                           int32_t* d_image3, uint32_t N, const hb2_box* box,
                                           float dt, const double* d_rescale,
                                           const float* d_last_pos4, float maxshift,
                                           uint32_t* d_flag, void* stream)
    {
    if (!d_pos4 || !d_vel4 || !d_force4 || !d_image3 || !box || (d_last_pos4 && !d_flag))
        {
        set_error("hb2_gpu_nve_step_one: null argument");
        return HB2_ERR_INVALID;
        }
    return launch_step_one((float4*)d_pos4, (float4*)d_vel4, (const float4*)d_force4, d_image3, N,
                           *box, dt, d_rescale, (const float4*)d_last_pos4, maxshift, d_flag,
                           (cudaStream_t)stream);
    }

extern "C" hb2_status hb2_gpu_nve_step_two(float* d_vel4, const float* d_force4, uint32_t N, float dt,
                                           const double* d_rescale, double* d_ke_out, void* stream)
    {
    if (!d_vel4 || !d_force4)
        {
        set_error("hb2_gpu_nve_step_two: null argument");
        return HB2_ERR_INVALID;
        }
    if (!d_ke_out)
        return launch_step_two((float4*)d_vel4, (const float4*)d_force4, N, dt, d_rescale, nullptr,
                               nullptr, nullptr, (cudaStream_t)stream);
    // scratch for the deterministic reduction (layer 1 is stateless: allocate stream-ordered)
    double* partial = nullptr;
    uint32_t* ticket = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    HB2_CUDA(cudaMallocAsync((void**)&partial, sizeof(double) * (step_two_blocks(N) + 1), st));
    HB2_CUDA(cudaMallocAsync((void**)&ticket, sizeof(uint32_t), st));
    HB2_CUDA(cudaMemsetAsync(ticket, 0, sizeof(uint32_t), st));
    hb2_status s = launch_step_two((float4*)d_vel4, (const float4*)d_force4, N, dt, d_rescale,
                                   partial, ticket, d_ke_out, st);
    cudaFreeAsync(partial, st);
    cudaFreeAsync(ticket, st);
    return s;
    }
