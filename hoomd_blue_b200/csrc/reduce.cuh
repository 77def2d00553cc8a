// reduce.cuh — deterministic fp64 grid reduction shared by the integration kernels and the fused
// force + step-two kernel.
#pragma once
#include "common.cuh"

namespace hb2
    {
// Deterministic block -> grid fp64 sum: every block writes its partial, the last block to finish
// (ticket) adds the partials in index order with a fixed tree.
// Returns true in the one thread that wrote the final sums (so a caller can chain scalar work).
template<int NV, int BLOCK>
HB2_DEVICE bool grid_sum(double (&v)[NV], double* __restrict__ partial, uint32_t* __restrict__ ticket,
                         double* __restrict__ out)
    {
    __shared__ double s_w[NV][BLOCK / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; q++)
        {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
            v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
        if (lane == 0)
            s_w[q][warp] = v[q];
        }
    __syncthreads();
    if (threadIdx.x == 0)
        {
#pragma unroll
        for (int q = 0; q < NV; q++)
            {
            double t = 0.0;
            for (int w = 0; w < BLOCK / 32; w++)
                t += s_w[q][w];
            partial[(size_t)q * gridDim.x + blockIdx.x] = t;
            }
        __threadfence();
        const uint32_t done = atomicAdd(ticket, 1u);
        s_last = (done == gridDim.x - 1);
        }
    __syncthreads();
    if (s_last)
        {
        __threadfence();
#pragma unroll
        for (int q = 0; q < NV; q++)
            {
            double t = 0.0;
            for (uint32_t b = threadIdx.x; b < gridDim.x; b += BLOCK)
                t += partial[(size_t)q * gridDim.x + b];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
                t += __shfl_xor_sync(0xffffffffu, t, off);
            if (lane == 0)
                s_w[q][warp] = t;
            }
        __syncthreads();
        if (threadIdx.x == 0)
            {
#pragma unroll
            for (int q = 0; q < NV; q++)
                {
                double t = 0.0;
                for (int w = 0; w < BLOCK / 32; w++)
                    t += s_w[q][w];
                out[q] = t;
                }
            *ticket = 0u;
            return true;
            }
        }
    return false;
    }

    } // namespace hb2
