// common.cuh — shared device/host helpers of the B200 pair-force path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "../../include/hoomd_b200.h"

#define HB2_DEVICE __device__ __forceinline__
#define HB2_HOSTDEVICE __host__ __device__ __forceinline__

namespace hb2
    {
constexpr int NUM_SMS = 148; // B200: 2 dies x 74 SMs

void set_error(const std::string& msg);
hb2_status cuda_fail(cudaError_t e, const char* what);
extern uint64_t g_launches;
inline void count_launch(int n = 1)
    {
    g_launches += n;
    }

#define HB2_CUDA(call)                                   \
    do                                                   \
        {                                                \
        cudaError_t _e = (call);                         \
        if (_e != cudaSuccess)                           \
            return ::hb2::cuda_fail(_e, #call);          \
        } while (0)

#define HB2_CHECK_LAUNCH(name)                           \
    do                                                   \
        {                                                \
        cudaError_t _e = cudaPeekAtLastError();          \
        if (_e != cudaSuccess)                           \
            return ::hb2::cuda_fail(_e, name);           \
        ::hb2::count_launch();                           \
        } while (0)

// fp32 box passed by value to kernels (the hot loops); fp64 box for the exact paths.
struct BoxF
    {
    float L[3], lo[3], hi[3], Linv[3];
    };
struct BoxD
    {
    double L[3], lo[3], hi[3], Linv[3];
    };
inline BoxF boxf(const hb2_box& b)
    {
    BoxF o;
    for (int d = 0; d < 3; d++)
        {
        o.L[d] = b.Lf[d];
        o.lo[d] = b.lof[d];
        o.hi[d] = b.hif[d];
        o.Linv[d] = b.Linvf[d];
        }
    return o;
    }
inline BoxD boxd(const hb2_box& b)
    {
    BoxD o;
    for (int d = 0; d < 3; d++)
        {
        o.L[d] = b.L[d];
        o.lo[d] = b.lo[d];
        o.hi[d] = b.hi[d];
        o.Linv[d] = b.Linv[d];
        }
    return o;
    }

// Exact restatement of the CPU minimum image (hoomd/BoxDim.h:343-389, zero tilt, periodic) in
// fp64 with explicit round-to-nearest intrinsics so nvcc cannot contract to FMA.
HB2_DEVICE void min_image_exact(const BoxD& b, double& x, double& y, double& z)
    {
    if (z >= b.hi[2])
        z = __dadd_rn(z, -b.L[2]);
    else if (z < b.lo[2])
        z = __dadd_rn(z, b.L[2]);
    if (y >= b.hi[1])
        {
        int i = (int)__dadd_rn(__dmul_rn(y, b.Linv[1]), 0.5);
        y = __dadd_rn(y, -__dmul_rn((double)i, b.L[1]));
        }
    else if (y < b.lo[1])
        {
        int i = (int)__dadd_rn(__dmul_rn(-y, b.Linv[1]), 0.5);
        y = __dadd_rn(y, __dmul_rn((double)i, b.L[1]));
        }
    if (x >= b.hi[0])
        {
        int i = (int)__dadd_rn(__dmul_rn(x, b.Linv[0]), 0.5);
        x = __dadd_rn(x, -__dmul_rn((double)i, b.L[0]));
        }
    else if (x < b.lo[0])
        {
        int i = (int)__dadd_rn(__dmul_rn(-x, b.Linv[0]), 0.5);
        x = __dadd_rn(x, __dmul_rn((double)i, b.L[0]));
        }
    }

// |minImage(pi - pj)|^2 exactly as hoomd/md/NeighborListBinned.cc:139-143 computes it on the CPU
// build: fp64, x*x + y*y + z*z left to right, no FMA.
HB2_DEVICE double dist_sq_exact(const BoxD& b, const float4& pi, const float4& pj)
    {
    double x = __dadd_rn((double)pi.x, -(double)pj.x);
    double y = __dadd_rn((double)pi.y, -(double)pj.y);
    double z = __dadd_rn((double)pi.z, -(double)pj.z);
    min_image_exact(b, x, y, z);
    return __dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z));
    }

// fp32 minimum image for one component when |d| < 1.5 L (always true for wrapped positions).
HB2_DEVICE float min_image_f(float d, float L, float hi)
    {
    if (d >= hi)
        d -= L;
    else if (d < -hi)
        d += L;
    return d;
    }

HB2_DEVICE float4 ldg4(const float4* p)
    {
    return __ldg(p);
    }

    } // namespace hb2
