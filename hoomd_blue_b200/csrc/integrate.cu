// integrate.cu — ConstantVolume velocity-Verlet half steps, the neighbour-list distance check folded
// into step one, deterministic fp64 reductions (kinetic energy, potential energy, virial) and the
// Bussi thermostat factor computed on the device (no host round trip inside the step loop).
//
// Replaces TwoStepConstantVolume::integrateStepOne/Two (hoomd/md/TwoStepConstantVolume.cc:7-80,
// :213-261; GPU kernels hoomd/md/TwoStepConstantVolumeGPU.cu:35-86,166-198),
// NeighborList::distanceCheck (hoomd/md/NeighborList.cc:1220-1289; GPU kernel
// hoomd/md/NeighborListGPU.cu:42-84), ComputeThermo's kinetic/potential sums
// (hoomd/md/ComputeThermo.cc:139-153,213-220) and BussiThermostat (hoomd/md/Thermostat.h:357-447).
#include "common.cuh"
#include "kernels.cuh"
#include "rng.cuh"
#include "reduce.cuh"

namespace hb2
    {
constexpr int INT_BLOCK = 256;

// v += 1/2 a dt; v *= s; r += v dt; wrap (x, y, z; >= hi -> -L, image+1; < lo -> +L, image-1:
// hoomd/BoxDim.h:407-474). The acceleration is F/m of the force array left by the previous step
// (what step two stored as `accel`, hoomd/md/TwoStepConstantVolume.cc:247-250). Then the distance
// check of the coming nlist->compute: |minImage(r - r_last)|^2 >= maxshift^2 sets the flag.
__global__ void __launch_bounds__(INT_BLOCK)
    step_one_kernel(float4* __restrict__ pos, float4* __restrict__ vel,
                    const float4* __restrict__ force, int* __restrict__ image, uint32_t N, BoxF box,
                    float dt, const double* __restrict__ d_rescale,
                    const float4* __restrict__ last_pos, float maxshiftsq,
                    uint32_t* __restrict__ d_flag)
    {
    const uint32_t i = blockIdx.x * INT_BLOCK + threadIdx.x;
    if (i >= N)
        return;
    const float s = d_rescale ? (float)(*d_rescale) : 1.0f;
    float4 p = pos[i];
    float4 v = vel[i];
    const float4 f = force[i];
    const float minv = 1.0f / v.w;
    const float hdt = 0.5f * dt;
    v.x = (v.x + hdt * (f.x * minv)) * s;
    v.y = (v.y + hdt * (f.y * minv)) * s;
    v.z = (v.z + hdt * (f.z * minv)) * s;
    p.x += dt * v.x;
    p.y += dt * v.y;
    p.z += dt * v.z;
    int ix = 0, iy = 0, iz = 0;
    if (p.x >= box.hi[0])
        {
        p.x -= box.L[0];
        ix = 1;
        }
    else if (p.x < box.lo[0])
        {
        p.x += box.L[0];
        ix = -1;
        }
    if (p.y >= box.hi[1])
        {
        p.y -= box.L[1];
        iy = 1;
        }
    else if (p.y < box.lo[1])
        {
        p.y += box.L[1];
        iy = -1;
        }
    if (p.z >= box.hi[2])
        {
        p.z -= box.L[2];
        iz = 1;
        }
    else if (p.z < box.lo[2])
        {
        p.z += box.L[2];
        iz = -1;
        }
    pos[i] = p;
    vel[i] = v;
    if (ix | iy | iz) // image traffic only for the few particles that crossed a face
        {
        image[3 * i + 0] += ix;
        image[3 * i + 1] += iy;
        image[3 * i + 2] += iz;
        }
    if (last_pos)
        {
        const float4 q = last_pos[i];
        const float dx = min_image_f(p.x - q.x, box.L[0], box.hi[0]);
        const float dy = min_image_f(p.y - q.y, box.L[1], box.hi[1]);
        const float dz = min_image_f(p.z - q.z, box.L[2], box.hi[2]);
        if (dx * dx + dy * dy + dz * dz >= maxshiftsq)
            *d_flag = 1u; // benign race: every writer stores the same value
        }
    }

// a = F/m; v *= s; v += 1/2 dt a (hoomd/md/TwoStepConstantVolume.cc:234-260), optionally followed
// by KE = 1/2 sum m v^2 in fp64 (hoomd/md/ComputeThermo.cc:139-153).
// When `bussi.rescale` is set, the thread that finishes the KE sum also draws the Bussi factor of the
// NEXT step (hoomd/md/Thermostat.h:357-447) — no extra launch, no host round trip.
template<bool KE>
__global__ void __launch_bounds__(INT_BLOCK)
    step_two_kernel(float4* __restrict__ vel, const float4* __restrict__ force, uint32_t N, float dt,
                    const double* __restrict__ d_rescale, double* __restrict__ partial,
                    uint32_t* __restrict__ ticket, double* __restrict__ ke_out, BussiNext bussi)
    {
    const uint32_t i = blockIdx.x * INT_BLOCK + threadIdx.x;
    double ke[1] = {0.0};
    if (i < N)
        {
        const float s = d_rescale ? (float)(*d_rescale) : 1.0f;
        float4 v = vel[i];
        const float4 f = force[i];
        const float minv = 1.0f / v.w;
        const float hdt = 0.5f * dt;
        v.x = v.x * s + hdt * (f.x * minv);
        v.y = v.y * s + hdt * (f.y * minv);
        v.z = v.z * s + hdt * (f.z * minv);
        vel[i] = v;
        if (KE)
            ke[0] = 0.5 * (double)v.w
                    * ((double)v.x * (double)v.x + (double)v.y * (double)v.y
                       + (double)v.z * (double)v.z);
        }
    if (KE)
        {
        const bool fin = grid_sum<1, INT_BLOCK>(ke, partial, ticket, ke_out);
        if (fin && bussi.rescale)
            *bussi.rescale = bussi_factor(*ke_out, bussi.dof, bussi.dt, bussi.kT, bussi.tau,
                                          bussi.timestep, (uint16_t)bussi.seed, bussi.first_tag);
        }
    }

// TwoStepLangevin::integrateStepTwo (hoomd/md/TwoStepLangevin.cc:213-292, translational, 3D):
// a = (F + r sqrt(6 gamma kT / dt) - gamma v) / m with r uniform in [-1, 1] from
// RandomGenerator(Seed(TwoStepLangevin, timestep, seed), Counter(tag)); v += 1/2 dt a. The three
// draws are three Philox calls keyed by the particle's TAG, so the noise does not depend on the
// memory order or the domain decomposition. gc[2 t] = gamma, gc[2 t + 1] = sqrt(6 gamma kT / dt).
// The acceleration the reference stores for the next step one includes the BD force
// (TwoStepLangevin.cc:281-284); here step one derives it from a force array, so the total force
// F + F_bd is written to `total` for it.
__global__ void __launch_bounds__(INT_BLOCK)
    step_two_langevin_kernel(float4* __restrict__ vel, const float4* __restrict__ force,
                             float4* __restrict__ total, const float4* __restrict__ pos,
                             const uint32_t* __restrict__ tag, uint32_t N, float dt,
                             const float* __restrict__ gc, uint32_t seed, uint64_t timestep)
    {
    const uint32_t i = blockIdx.x * INT_BLOCK + threadIdx.x;
    if (i >= N)
        return;
    float4 v = vel[i];
    const float4 f = force[i];
    const uint32_t t = __float_as_uint(pos[i].w);
    const float gamma = gc[2 * t], coeff = gc[2 * t + 1];
    Rng g(RNG_ID_LANGEVIN, timestep, (uint16_t)seed, tag[i]);
    const float rx = (float)(-1.0 + 2.0 * rng_canonical(g));
    const float ry = (float)(-1.0 + 2.0 * rng_canonical(g));
    const float rz = (float)(-1.0 + 2.0 * rng_canonical(g));
    const float minv = 1.0f / v.w;
    const float hdt = 0.5f * dt;
    const float tx = f.x + (rx * coeff - gamma * v.x);
    const float ty = f.y + (ry * coeff - gamma * v.y);
    const float tz = f.z + (rz * coeff - gamma * v.z);
    v.x += hdt * (tx * minv);
    v.y += hdt * (ty * minv);
    v.z += hdt * (tz * minv);
    vel[i] = v;
    total[i] = make_float4(tx, ty, tz, f.w);
    }

// net force += one more force (Integrator::computeNetForce, hoomd/Integrator.cc:166-246)
__global__ void __launch_bounds__(INT_BLOCK)
    accumulate_force_kernel(float4* __restrict__ dst, float* __restrict__ dst_virial, size_t dst_pitch,
                            const float4* __restrict__ src, const float* __restrict__ src_virial,
                            size_t src_pitch, uint32_t N, const uint32_t* __restrict__ skip_flag)
    {
    if (skip_flag && *skip_flag)
        return;
    const uint32_t i = blockIdx.x * INT_BLOCK + threadIdx.x;
    if (i >= N)
        return;
    float4 d = dst[i];
    const float4 s = src[i];
    d.x += s.x;
    d.y += s.y;
    d.z += s.z;
    d.w += s.w;
    dst[i] = d;
    if (dst_virial && src_virial)
        {
#pragma unroll
        for (int q = 0; q < 6; q++)
            dst_virial[q * dst_pitch + i] += src_virial[q * src_pitch + i];
        }
    }

hb2_status launch_accumulate_force(float4* dst, float* dst_virial, size_t dst_pitch, const float4* src,
                                   const float* src_virial, size_t src_pitch, uint32_t N,
                                   const uint32_t* skip_flag, cudaStream_t stream)
    {
    if (N == 0)
        return HB2_OK;
    accumulate_force_kernel<<<(N + INT_BLOCK - 1) / INT_BLOCK, INT_BLOCK, 0, stream>>>(
        dst, dst_virial, dst_pitch, src, src_virial, src_pitch, N, skip_flag);
    HB2_CHECK_LAUNCH("accumulate_force_kernel");
    return HB2_OK;
    }

hb2_status launch_step_two_langevin(float4* vel, const float4* force, float4* total, const float4* pos,
                                    const uint32_t* tag, uint32_t N, float dt, const float* d_gamma_coeff,
                                    uint32_t seed, uint64_t timestep, cudaStream_t stream)
    {
    if (N == 0)
        return HB2_OK;
    step_two_langevin_kernel<<<(N + INT_BLOCK - 1) / INT_BLOCK, INT_BLOCK, 0, stream>>>(
        vel, force, total, pos, tag, N, dt, d_gamma_coeff, seed, timestep);
    HB2_CHECK_LAUNCH("step_two_langevin_kernel");
    return HB2_OK;
    }

// KE of the current velocities only (prepRun / thermo queries)
__global__ void __launch_bounds__(INT_BLOCK)
    kinetic_kernel(const float4* __restrict__ vel, uint32_t N, double* __restrict__ partial,
                   uint32_t* __restrict__ ticket, double* __restrict__ ke_out)
    {
    const uint32_t i = blockIdx.x * INT_BLOCK + threadIdx.x;
    double ke[1] = {0.0};
    if (i < N)
        {
        const float4 v = vel[i];
        ke[0] = 0.5 * (double)v.w
                * ((double)v.x * (double)v.x + (double)v.y * (double)v.y + (double)v.z * (double)v.z);
        }
    grid_sum<1, INT_BLOCK>(ke, partial, ticket, ke_out);
    }

// sum of per-particle potential energy (force.w, hoomd/ForceCompute.cc:213-220) and the six
// per-particle virial components
__global__ void __launch_bounds__(INT_BLOCK)
    energy_kernel(const float4* __restrict__ force, const float* __restrict__ virial,
                  size_t virial_pitch, uint32_t N, double* __restrict__ partial,
                  uint32_t* __restrict__ ticket, double* __restrict__ out7)
    {
    const uint32_t i = blockIdx.x * INT_BLOCK + threadIdx.x;
    double v[7] = {0, 0, 0, 0, 0, 0, 0};
    if (i < N)
        {
        v[0] = (double)force[i].w;
        if (virial)
            for (int q = 0; q < 6; q++)
                v[1 + q] = (double)virial[q * virial_pitch + i];
        }
    grid_sum<7, INT_BLOCK>(v, partial, ticket, out7);
    }

__global__ void bussi_kernel(const double* __restrict__ d_ke, double dof, double dt, double kT,
                             double tau, uint64_t timestep, uint32_t seed, uint32_t first_tag,
                             double* __restrict__ d_rescale)
    {
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *d_rescale = bussi_factor(*d_ke, dof, dt, kT, tau, timestep, (uint16_t)seed, first_tag);
    }

__global__ void advance_timestep_kernel(uint64_t* d_timestep)
    {
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *d_timestep += 1;
    }

uint32_t step_two_blocks(uint32_t N)
    {
    return (N + INT_BLOCK - 1) / INT_BLOCK;
    }
uint32_t reduce_blocks(uint32_t N)
    {
    return (N + INT_BLOCK - 1) / INT_BLOCK;
    }

hb2_status launch_step_one(float4* pos, float4* vel, const float4* force, int* image, uint32_t N,
                           const hb2_box& box, float dt, const double* d_rescale,
                           const float4* last_pos, float maxshift, uint32_t* d_flag,
                           cudaStream_t stream)
    {
    if (N == 0)
        return HB2_OK;
    step_one_kernel<<<(N + INT_BLOCK - 1) / INT_BLOCK, INT_BLOCK, 0, stream>>>(
        pos, vel, force, image, N, boxf(box), dt, d_rescale, last_pos, maxshift * maxshift, d_flag);
    HB2_CHECK_LAUNCH("step_one_kernel");
    return HB2_OK;
    }

hb2_status launch_step_two(float4* vel, const float4* force, uint32_t N, float dt,
                           const double* d_rescale, double* d_ke_partial, uint32_t* d_ticket,
                           double* d_ke_out, cudaStream_t stream, const BussiNext* bussi)
    {
    if (N == 0)
        return HB2_OK;
    const uint32_t blocks = step_two_blocks(N);
    BussiNext b;
    if (bussi)
        b = *bussi;
    if (d_ke_out)
        step_two_kernel<true><<<blocks, INT_BLOCK, 0, stream>>>(vel, force, N, dt, d_rescale,
                                                                d_ke_partial, d_ticket, d_ke_out, b);
    else
        step_two_kernel<false><<<blocks, INT_BLOCK, 0, stream>>>(vel, force, N, dt, d_rescale,
                                                                 nullptr, nullptr, nullptr, b);
    HB2_CHECK_LAUNCH("step_two_kernel");
    return HB2_OK;
    }

hb2_status launch_kinetic(const float4* vel, uint32_t N, double* d_partial, uint32_t* d_ticket,
                          double* d_ke_out, cudaStream_t stream)
    {
    if (N == 0)
        return HB2_OK;
    kinetic_kernel<<<reduce_blocks(N), INT_BLOCK, 0, stream>>>(vel, N, d_partial, d_ticket, d_ke_out);
    HB2_CHECK_LAUNCH("kinetic_kernel");
    return HB2_OK;
    }

hb2_status launch_reduce_energy(const float4* force, const float* virial, size_t virial_pitch,
                                uint32_t N, double* d_partial, uint32_t* d_ticket, double* d_out7,
                                cudaStream_t stream)
    {
    if (N == 0)
        return HB2_OK;
    energy_kernel<<<reduce_blocks(N), INT_BLOCK, 0, stream>>>(force, virial, virial_pitch, N,
                                                              d_partial, d_ticket, d_out7);
    HB2_CHECK_LAUNCH("energy_kernel");
    return HB2_OK;
    }

hb2_status launch_bussi(const double* d_ke, double dof, double dt, double kT, double tau,
                        uint64_t timestep, uint32_t seed, uint32_t first_tag, double* d_rescale,
                        cudaStream_t stream)
    {
    bussi_kernel<<<1, 32, 0, stream>>>(d_ke, dof, dt, kT, tau, timestep, seed, first_tag, d_rescale);
    HB2_CHECK_LAUNCH("bussi_kernel");
    return HB2_OK;
    }

hb2_status launch_advance_timestep(uint64_t* d_timestep, cudaStream_t stream)
    {
    advance_timestep_kernel<<<1, 32, 0, stream>>>(d_timestep);
    HB2_CHECK_LAUNCH("advance_timestep_kernel");
    return HB2_OK;
    }

double bussi_factor_host(double K, double dof, double dt, double kT, double tau, uint64_t timestep,
                         uint32_t seed, uint32_t first_tag)
    {
    return bussi_factor(K, dof, dt, kT, tau, timestep, (uint16_t)seed, first_tag);
    }
    } // namespace hb2
