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
#include <algorithm>
#include <cstdlib>

namespace hb2
    {
constexpr int INT_BLOCK = 256;

// The half kicks and the drift with explicit roundings: the stand-alone kernels and the fused
// step-two/step-one kernel must produce the same bits (a run cut into segments by a writer is the same
// trajectory as an uninterrupted one), whatever the compiler would contract.
HB2_DEVICE bool in_group(const GroupMask& g, const uint32_t i)
    {
    return !g.member || g.member[g.tag[i]] != 0;
    }
HB2_DEVICE float kick(const float v, const float f, const float minv, const float hdt)
    {
    return __fmaf_rn(hdt, __fmul_rn(f, minv), v);
    }
HB2_DEVICE float drift(const float p, const float v, const float dt)
    {
    return __fmaf_rn(dt, v, p);
    }

// v += 1/2 a dt; v *= s; r += v dt; wrap (x, y, z; >= hi -> -L, image+1; < lo -> +L, image-1:
// hoomd/BoxDim.h:407-474). The acceleration is F/m of the force array left by the previous step
// (what step two stored as `accel`, hoomd/md/TwoStepConstantVolume.cc:247-250). Then the distance
// check of the coming nlist->compute: |minImage(r - r_last)|^2 >= maxshift^2 sets the flag.
__global__ void __launch_bounds__(INT_BLOCK)
    step_one_kernel(float4* __restrict__ pos, float4* __restrict__ vel,
                    const float4* __restrict__ force, int* __restrict__ image, uint32_t N, BoxF box,
                    float dt, const double* __restrict__ d_rescale,
                    const float4* __restrict__ last_pos, float maxshiftsq,
                    uint32_t* __restrict__ d_flag, GroupMask group)
    {
    const uint32_t i = blockIdx.x * INT_BLOCK + threadIdx.x;
    if (i >= N || !in_group(group, i)) // particles outside the method's group do not move
        return;
    const float s = d_rescale ? (float)(*d_rescale) : 1.0f;
    float4 p = pos[i];
    float4 v = vel[i];
    const float4 f = force[i];
    const float minv = 1.0f / v.w;
    const float hdt = 0.5f * dt;
    v.x = __fmul_rn(kick(v.x, f.x, minv, hdt), s);
    v.y = __fmul_rn(kick(v.y, f.y, minv, hdt), s);
    v.z = __fmul_rn(kick(v.z, f.z, minv, hdt), s);
    p.x = drift(p.x, v.x, dt);
    p.y = drift(p.y, v.y, dt);
    p.z = drift(p.z, v.z, dt);
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
                    uint32_t* __restrict__ ticket, double* __restrict__ ke_out, BussiNext bussi, GroupMask group)
    {
    const uint32_t i = blockIdx.x * INT_BLOCK + threadIdx.x;
    double ke[1] = {0.0};
    if (i < N && in_group(group, i))
        {
        const float s = d_rescale ? (float)(*d_rescale) : 1.0f;
        float4 v = vel[i];
        const float4 f = force[i];
        const float minv = 1.0f / v.w;
        const float hdt = 0.5f * dt;
        v.x = kick(__fmul_rn(v.x, s), f.x, minv, hdt);
        v.y = kick(__fmul_rn(v.y, s), f.y, minv, hdt);
        v.z = kick(__fmul_rn(v.z, s), f.z, minv, hdt);
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

// ---- step two of timestep t fused with step one of timestep t + 1 (the same force array serves both:
// what step two stores as the acceleration is what the next step one reads, hoomd/md/
// TwoStepConstantVolume.cc:247-250): one launch, and the velocities written by the first half are read
// back by the second out of L2 — 96 instead of 144 bytes of HBM traffic per particle while the velocity
// and force arrays fit in L2.
// With the Bussi thermostat the factor of step t + 1 depends on the kinetic energy of ALL particles after
// step two: the kernel is then persistent (one co-resident wave of blocks, each walking "virtual blocks" of
// INT_BLOCK consecutive particles), the deterministic KE sum is the one of step_two_kernel (same virtual
// blocks, same trees: bit-identical), the block that completes it draws the factor and publishes a
// generation word, and every block waits for that word between the halves.
HB2_DEVICE uint32_t ld_volatile_u32(const uint32_t* p)
    {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
    }

template<bool THERMO>
__global__ void __launch_bounds__(INT_BLOCK)
    step_two_one_kernel(float4* __restrict__ pos, float4* __restrict__ vel, const float4* __restrict__ force,
                        int* __restrict__ image, uint32_t N, BoxF box, float dt,
                        const float4* __restrict__ last_pos, float maxshiftsq, uint32_t* __restrict__ d_flag,
                        double* __restrict__ partial, uint32_t* __restrict__ ticket,
                        double* __restrict__ ke_out, BussiNext bussi, uint32_t* __restrict__ d_gen,
                        uint32_t epoch, GroupMask group)
    {
    __shared__ double s_w[4][INT_BLOCK / 32];
    __shared__ bool s_last;
    const uint32_t nvb = (N + INT_BLOCK - 1) / INT_BLOCK;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float hdt = 0.5f * dt;
    // ---- step two of timestep t (no rescale: the factor is applied in step one). FU virtual blocks per
    // trip: their loads are all issued before the first is used
    constexpr int FU = 4;
    for (uint32_t vb0 = blockIdx.x; vb0 < nvb; vb0 += FU * gridDim.x)
        {
        float4 v[FU], f[FU];
        uint32_t idx[FU];
#pragma unroll
        for (int u = 0; u < FU; u++)
            {
            const uint32_t vb = vb0 + u * gridDim.x;
            idx[u] = vb < nvb ? vb * INT_BLOCK + threadIdx.x : N;
            if (idx[u] < N && !in_group(group, idx[u]))
                idx[u] = N;
            if (idx[u] < N)
                {
                v[u] = vel[idx[u]];
                f[u] = force[idx[u]];
                }
            }
        double ke[FU];
#pragma unroll
        for (int u = 0; u < FU; u++)
            {
            ke[u] = 0.0;
            if (idx[u] < N)
                {
                const float minv = 1.0f / v[u].w;
                v[u].x = kick(v[u].x, f[u].x, minv, hdt);
                v[u].y = kick(v[u].y, f[u].y, minv, hdt);
                v[u].z = kick(v[u].z, f[u].z, minv, hdt);
                vel[idx[u]] = v[u];
                if (THERMO)
                    ke[u] = 0.5 * (double)v[u].w
                            * ((double)v[u].x * (double)v[u].x + (double)v[u].y * (double)v[u].y
                               + (double)v[u].z * (double)v[u].z);
                }
            }
        if (THERMO)
            {
            // the block stage of grid_sum, for each of the FU virtual blocks
#pragma unroll
            for (int u = 0; u < FU; u++)
                {
#pragma unroll
                for (int off = 16; off > 0; off >>= 1)
                    ke[u] += __shfl_xor_sync(0xffffffffu, ke[u], off);
                if (lane == 0)
                    s_w[u][warp] = ke[u];
                }
            __syncthreads();
            if (threadIdx.x < FU)
                {
                const uint32_t vb = vb0 + threadIdx.x * gridDim.x;
                if (vb < nvb)
                    {
                    double t = 0.0;
                    for (int w = 0; w < INT_BLOCK / 32; w++)
                        t += s_w[threadIdx.x][w];
                    partial[vb] = t;
                    }
                }
            __syncthreads();
            }
        }
    float s = 1.0f;
    if (THERMO)
        {
        if (threadIdx.x == 0)
            {
            __threadfence();
            s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
            }
        __syncthreads();
        if (s_last)
            {
            // the grid stage of grid_sum over the nvb partials
            __threadfence();
            double t = 0.0;
            for (uint32_t b = threadIdx.x; b < nvb; b += INT_BLOCK)
                t += partial[b];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
                t += __shfl_xor_sync(0xffffffffu, t, off);
            if (lane == 0)
                s_w[0][warp] = t;
            __syncthreads();
            if (threadIdx.x == 0)
                {
                double tot = 0.0;
                for (int w = 0; w < INT_BLOCK / 32; w++)
                    tot += s_w[0][w];
                *ke_out = tot;
                *ticket = 0u;
                *bussi.rescale = bussi_factor(tot, bussi.dof, bussi.dt, bussi.kT, bussi.tau, bussi.timestep,
                                              (uint16_t)bussi.seed, bussi.first_tag);
                __threadfence();
                atomicExch(d_gen, epoch);
                }
            }
        if (threadIdx.x == 0)
            {
            while (ld_volatile_u32(d_gen) != epoch)
                __nanosleep(64);
            __threadfence();
            }
        __syncthreads();
        s = (float)__ldcg(bussi.rescale);
        }
    // ---- step one of timestep t + 1 (step_one_kernel), two virtual blocks per trip
    constexpr int GU = 2;
    for (uint32_t vb0 = blockIdx.x; vb0 < nvb; vb0 += GU * gridDim.x)
        {
        float4 pp[GU], vv[GU], ff[GU], qq[GU];
        uint32_t idx[GU];
#pragma unroll
        for (int u = 0; u < GU; u++)
            {
            const uint32_t vb = vb0 + u * gridDim.x;
            idx[u] = vb < nvb ? vb * INT_BLOCK + threadIdx.x : N;
            if (idx[u] < N && !in_group(group, idx[u]))
                idx[u] = N;
            if (idx[u] < N)
                {
                pp[u] = pos[idx[u]];
                vv[u] = vel[idx[u]];
                ff[u] = force[idx[u]];
                if (last_pos)
                    qq[u] = last_pos[idx[u]];
                }
            }
#pragma unroll
        for (int u = 0; u < GU; u++)
            {
            const uint32_t i = idx[u];
            if (i >= N)
                continue;
            float4 p = pp[u], v = vv[u];
            const float4 f = ff[u];
            const float minv = 1.0f / v.w;
            v.x = __fmul_rn(kick(v.x, f.x, minv, hdt), s);
            v.y = __fmul_rn(kick(v.y, f.y, minv, hdt), s);
            v.z = __fmul_rn(kick(v.z, f.z, minv, hdt), s);
            p.x = drift(p.x, v.x, dt);
            p.y = drift(p.y, v.y, dt);
            p.z = drift(p.z, v.z, dt);
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
            if (ix | iy | iz)
                {
                image[3 * i + 0] += ix;
                image[3 * i + 1] += iy;
                image[3 * i + 2] += iz;
                }
            if (last_pos)
                {
                const float4 q = qq[u];
                const float dx = min_image_f(p.x - q.x, box.L[0], box.hi[0]);
                const float dy = min_image_f(p.y - q.y, box.L[1], box.hi[1]);
                const float dz = min_image_f(p.z - q.z, box.L[2], box.hi[2]);
                if (dx * dx + dy * dy + dz * dz >= maxshiftsq)
                    *d_flag = 1u;
                }
            }
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
                             const float* __restrict__ gc, uint32_t seed, uint64_t timestep,
                             const uint8_t* __restrict__ member)
    {
    const uint32_t i = blockIdx.x * INT_BLOCK + threadIdx.x;
    if (i >= N)
        return;
    if (member && !member[tag[i]])
        {
        total[i] = force[i]; // not integrated; the array stays defined
        return;
        }
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

// ---- ConstantPressure (MTK barostat, orthorhombic box degrees of freedom) -------------------------
// TwoStepConstantPressure::integrateStepOne (hoomd/md/TwoStepConstantPressure.cc:213-395; GPU kernels
// hoomd/md/TwoStepConstantPressureGPU.cu): v += 1/2 dt a; v <- exp_v v; v *= s; r <- exp_r r + exp_r_int v;
// wrap into the NEW box; then the neighbour list's distance check with the homogeneous dilation taken
// out (hoomd/md/NeighborList.cc:1236-1260): |minImage(r - lambda r_last)|^2 >= maxsq[type]. The
// propagator matrices are upper triangular; the host hands the diagonals (tilt is outside this path).
// Particles outside the group are only rescaled, and only with rescale_all.
__global__ void __launch_bounds__(INT_BLOCK)
    step_one_npt_kernel(float4* __restrict__ pos, float4* __restrict__ vel, const float4* __restrict__ force,
                        int* __restrict__ image, uint32_t N, BoxF box, NptFactors k,
                        const double* __restrict__ d_rescale, const float4* __restrict__ last_pos,
                        uint32_t* __restrict__ d_flag, GroupMask group)
    {
    const uint32_t i = blockIdx.x * INT_BLOCK + threadIdx.x;
    if (i >= N)
        return;
    const bool member = in_group(group, i);
    if (!member && !k.rescale_all)
        return;
    float4 p = pos[i];
    p.x *= k.exp_r[0];
    p.y *= k.exp_r[1];
    p.z *= k.exp_r[2];
    if (member)
        {
        const float s = (d_rescale ? (float)(*d_rescale) : 1.0f) * k.mtk;
        float4 v = vel[i];
        const float4 f = force[i];
        const float minv = 1.0f / v.w;
        const float hdt = 0.5f * k.dt;
        v.x = kick(v.x, f.x, minv, hdt) * k.exp_v[0] * s;
        v.y = kick(v.y, f.y, minv, hdt) * k.exp_v[1] * s;
        v.z = kick(v.z, f.z, minv, hdt) * k.exp_v[2] * s;
        p.x = __fmaf_rn(k.exp_r_int[0], v.x, p.x);
        p.y = __fmaf_rn(k.exp_r_int[1], v.y, p.y);
        p.z = __fmaf_rn(k.exp_r_int[2], v.z, p.z);
        vel[i] = v;
        }
    int ix = 0, iy = 0, iz = 0;
    // a rescaled coordinate can sit a rounding error outside the rescaled box as well as a drift outside
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
    if (ix | iy | iz)
        {
        image[3 * i + 0] += ix;
        image[3 * i + 1] += iy;
        image[3 * i + 2] += iz;
        }
    if (last_pos)
        {
        const float4 q = last_pos[i];
        const uint32_t t = __float_as_uint(p.w);
        const float dx = min_image_f(p.x - k.lambda[0] * q.x, box.L[0], box.hi[0]);
        const float dy = min_image_f(p.y - k.lambda[1] * q.y, box.L[1], box.hi[1]);
        const float dz = min_image_f(p.z - k.lambda[2] * q.z, box.L[2], box.hi[2]);
        if (dx * dx + dy * dy + dz * dz >= k.maxsq[t < NPT_MAX_TYPES ? t : NPT_MAX_TYPES - 1])
            *d_flag = 1u;
        }
    }

// TwoStepConstantPressure::integrateStepTwo (hoomd/md/TwoStepConstantPressure.cc:583-640):
// v *= s; v <- exp_v v; v += 1/2 dt a; then the kinetic part of the pressure tensor, sum m v_a v_b over
// the group in fp64 (hoomd/md/ComputeThermo.cc:104-123), out6 = xx, xy, xz, yy, yz, zz.
__global__ void __launch_bounds__(INT_BLOCK)
    step_two_npt_kernel(float4* __restrict__ vel, const float4* __restrict__ force, uint32_t N, NptFactors k,
                        double* __restrict__ partial, uint32_t* __restrict__ ticket, double* __restrict__ out6,
                        GroupMask group)
    {
    const uint32_t i = blockIdx.x * INT_BLOCK + threadIdx.x;
    double kt[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (i < N && in_group(group, i))
        {
        float4 v = vel[i];
        if (k.dt != 0.0f) // dt = 0: the tensor of the velocities as they are (thermo at the start of a run)
            {
            const float4 f = force[i];
            const float minv = 1.0f / v.w;
            const float hdt = 0.5f * k.dt;
            v.x = kick(v.x * k.mtk * k.exp_v[0], f.x, minv, hdt);
            v.y = kick(v.y * k.mtk * k.exp_v[1], f.y, minv, hdt);
            v.z = kick(v.z * k.mtk * k.exp_v[2], f.z, minv, hdt);
            vel[i] = v;
            }
        const double m = (double)v.w, x = (double)v.x, y = (double)v.y, z = (double)v.z;
        kt[0] = m * (x * x);
        kt[1] = m * (x * y);
        kt[2] = m * (x * z);
        kt[3] = m * (y * y);
        kt[4] = m * (y * z);
        kt[5] = m * (z * z);
        }
    grid_sum<6, INT_BLOCK>(kt, partial, ticket, out6);
    }

hb2_status launch_step_one_npt(float4* pos, float4* vel, const float4* force, int* image, uint32_t N,
                               const hb2_box& box, const NptFactors& k, const double* d_rescale,
                               const float4* last_pos, uint32_t* d_flag, cudaStream_t stream, GroupMask group)
    {
    if (N == 0)
        return HB2_OK;
    step_one_npt_kernel<<<(N + INT_BLOCK - 1) / INT_BLOCK, INT_BLOCK, 0, stream>>>(
        pos, vel, force, image, N, boxf(box), k, d_rescale, last_pos, d_flag, group);
    HB2_CHECK_LAUNCH("step_one_npt_kernel");
    return HB2_OK;
    }

hb2_status launch_step_two_npt(float4* vel, const float4* force, uint32_t N, const NptFactors& k,
                               double* d_partial, uint32_t* d_ticket, double* d_out6, cudaStream_t stream,
                               GroupMask group)
    {
    if (N == 0)
        return HB2_OK;
    step_two_npt_kernel<<<(N + INT_BLOCK - 1) / INT_BLOCK, INT_BLOCK, 0, stream>>>(vel, force, N, k, d_partial,
                                                                                   d_ticket, d_out6, group);
    HB2_CHECK_LAUNCH("step_two_npt_kernel");
    return HB2_OK;
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
                                    uint32_t seed, uint64_t timestep, cudaStream_t stream, const uint8_t* member)
    {
    if (N == 0)
        return HB2_OK;
    step_two_langevin_kernel<<<(N + INT_BLOCK - 1) / INT_BLOCK, INT_BLOCK, 0, stream>>>(
        vel, force, total, pos, tag, N, dt, d_gamma_coeff, seed, timestep, member);
    HB2_CHECK_LAUNCH("step_two_langevin_kernel");
    return HB2_OK;
    }

// KE of the current velocities only (prepRun / thermo queries)
__global__ void __launch_bounds__(INT_BLOCK)
    kinetic_kernel(const float4* __restrict__ vel, uint32_t N, double* __restrict__ partial,
                   uint32_t* __restrict__ ticket, double* __restrict__ ke_out, GroupMask group)
    {
    const uint32_t i = blockIdx.x * INT_BLOCK + threadIdx.x;
    double ke[1] = {0.0};
    if (i < N && in_group(group, i))
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
                           cudaStream_t stream, GroupMask group)
    {
    if (N == 0)
        return HB2_OK;
    step_one_kernel<<<(N + INT_BLOCK - 1) / INT_BLOCK, INT_BLOCK, 0, stream>>>(
        pos, vel, force, image, N, boxf(box), dt, d_rescale, last_pos, maxshift * maxshift, d_flag, group);
    HB2_CHECK_LAUNCH("step_one_kernel");
    return HB2_OK;
    }

hb2_status launch_step_two(float4* vel, const float4* force, uint32_t N, float dt,
                           const double* d_rescale, double* d_ke_partial, uint32_t* d_ticket,
                           double* d_ke_out, cudaStream_t stream, const BussiNext* bussi, GroupMask group)
    {
    if (N == 0)
        return HB2_OK;
    const uint32_t blocks = step_two_blocks(N);
    BussiNext b;
    if (bussi)
        b = *bussi;
    if (d_ke_out)
        step_two_kernel<true><<<blocks, INT_BLOCK, 0, stream>>>(vel, force, N, dt, d_rescale,
                                                                d_ke_partial, d_ticket, d_ke_out, b, group);
    else
        step_two_kernel<false><<<blocks, INT_BLOCK, 0, stream>>>(vel, force, N, dt, d_rescale,
                                                                 nullptr, nullptr, nullptr, b, group);
    HB2_CHECK_LAUNCH("step_two_kernel");
    return HB2_OK;
    }

// Step two of timestep t and step one of timestep t + 1 in one launch. bussi (with bussi->rescale set):
// the KE after step two goes to d_ke_out and the factor drawn from it is applied in the step-one half;
// d_gen is a zero-initialised device word, epoch a value that differs from launch to launch.
hb2_status launch_step_two_one(float4* pos, float4* vel, const float4* force, int* image, uint32_t N,
                               const hb2_box& box, float dt, const float4* last_pos, float maxshift,
                               uint32_t* d_flag, double* d_ke_partial, uint32_t* d_ticket, double* d_ke_out,
                               const BussiNext* bussi, uint32_t* d_gen, uint32_t epoch, cudaStream_t stream,
                               bool* fused, GroupMask group)
    {
    if (fused)
        *fused = true;
    if (N == 0)
        return HB2_OK;
    const uint32_t nvb = (N + INT_BLOCK - 1) / INT_BLOCK;
    if (bussi && bussi->rescale)
        {
        // one co-resident wave: the blocks wait for each other between the halves
        static int wave = 0;
        if (wave == 0)
            {
            int dev = 0, sms = 0, per_sm = 0;
            HB2_CUDA(cudaGetDevice(&dev));
            HB2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            HB2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step_two_one_kernel<true>, INT_BLOCK, 0));
            wave = sms * per_sm;
            if (wave <= 0)
                {
                set_error("step_two_one_kernel: no resident blocks");
                return HB2_ERR_INVALID;
                }
            }
        // A cooperative launch: the runtime guarantees that the whole grid is resident together (the blocks
        // wait for each other between the halves) or refuses — then *fused is cleared and the caller uses
        // the two stand-alone kernels.
        float4* a_pos = pos;
        float4* a_vel = vel;
        const float4* a_force = force;
        int* a_image = image;
        uint32_t a_N = N;
        BoxF a_box = boxf(box);
        float a_dt = dt;
        const float4* a_last = last_pos;
        float a_ms = maxshift * maxshift;
        uint32_t* a_flag = d_flag;
        double* a_part = d_ke_partial;
        uint32_t* a_ticket = d_ticket;
        double* a_ke = d_ke_out;
        BussiNext a_b = *bussi;
        uint32_t* a_gen = d_gen;
        uint32_t a_epoch = epoch;
        GroupMask a_group = group;
        void* args[] = {&a_pos, &a_vel, &a_force, &a_image, &a_N, &a_box, &a_dt, &a_last, &a_ms, &a_flag,
                        &a_part, &a_ticket, &a_ke, &a_b, &a_gen, &a_epoch, &a_group};
        // HB2_FUSE_COOP=0: a plain launch of the same one-wave grid (A/B of the cooperative launch's cost; the
        // blocks still all become resident on a stream that runs nothing else, but nothing guarantees it)
        static const bool coop = !(getenv("HB2_FUSE_COOP") && atoi(getenv("HB2_FUSE_COOP")) == 0);
        const cudaError_t e = coop ? cudaLaunchCooperativeKernel((const void*)step_two_one_kernel<true>,
                                                                 dim3(std::min<uint32_t>(nvb, (uint32_t)wave)),
                                                                 dim3(INT_BLOCK), args, 0, stream)
                                   : cudaLaunchKernel((const void*)step_two_one_kernel<true>,
                                                      dim3(std::min<uint32_t>(nvb, (uint32_t)wave)),
                                                      dim3(INT_BLOCK), args, 0, stream);
        if (e != cudaSuccess)
            {
            cudaGetLastError();
            if (fused)
                *fused = false;
            return HB2_OK;
            }
        count_launch(1);
        if (fused)
            *fused = true;
        return HB2_OK;
        }
    else
        {
        BussiNext none;
        step_two_one_kernel<false><<<nvb, INT_BLOCK, 0, stream>>>(pos, vel, force, image, N, boxf(box), dt, last_pos,
                                                                  maxshift * maxshift, d_flag, nullptr, nullptr,
                                                                  nullptr, none, nullptr, 0u, group);
        }
    HB2_CHECK_LAUNCH("step_two_one_kernel");
    return HB2_OK;
    }

hb2_status launch_kinetic(const float4* vel, uint32_t N, double* d_partial, uint32_t* d_ticket,
                          double* d_ke_out, cudaStream_t stream, GroupMask group)
    {
    if (N == 0)
        return HB2_OK;
    kinetic_kernel<<<reduce_blocks(N), INT_BLOCK, 0, stream>>>(vel, N, d_partial, d_ticket, d_ke_out, group);
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

// n successive standard-normal draws of RandomGenerator(Seed(id, timestep, seed), Counter(first_tag)) on
// the host (the barostat's noise, hoomd/md/TwoStepConstantPressure.cc:907-945, and its thermalisation)
void rng_normals_host(uint8_t id, uint64_t timestep, uint32_t seed, uint32_t first_tag, int n, double* out)
    {
    Rng g(id, timestep, (uint16_t)seed, first_tag);
    for (int k = 0; k < n; k++)
        out[k] = rng_normal(g);
    }

double bussi_factor_host(double K, double dof, double dt, double kT, double tau, uint64_t timestep,
                         uint32_t seed, uint32_t first_tag)
    {
    return bussi_factor(K, dof, dt, kT, tau, timestep, (uint16_t)seed, first_tag);
    }
    } // namespace hb2
