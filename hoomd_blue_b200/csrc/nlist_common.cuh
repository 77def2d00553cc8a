// nlist_common.cuh — stencil geometry and warp helpers shared by the neighbour-list build kernels
// (nlist.cu: per-particle rows, the layer-1 driver; nlist_quad.cu: quad rows, the engine's list).
#pragma once
#include "common.cuh"
#include <math.h>

namespace hb2
    {
constexpr float NL_FAR = 1.0e18f;

struct NlConst
    {
    BoxF bf;
    BoxD bd;
    float band; // relative guard band of the fp32 distance test
    };

// k-th (ascending) member of the sorted, de-duplicated periodic neighbourhood {c-1, c, c+1} of
// coordinate c on a ring of size dim, and the image shift (in units of L) that brings that cell next
// to c. Returns the neighbourhood size.
HB2_DEVICE int ring3(int c, int dim, int k, int& coord, int& shift)
    {
    shift = 0;
    if (dim == 1)
        {
        coord = 0;
        return 1;
        }
    if (dim == 2)
        {
        coord = k & 1;
        return 2;
        }
    if (c == 0)
        {
        coord = (k == 2) ? dim - 1 : k;
        shift = (k == 2) ? -1 : 0;
        }
    else if (c == dim - 1)
        {
        coord = (k == 0) ? 0 : dim - 3 + k;
        shift = (k == 0) ? 1 : 0;
        }
    else
        coord = c - 1 + k;
    return 3;
    }

// The same along a dimension that may be decomposed over n ranks (this rank at grid coordinate rc):
// the local grid is then open — interior home cells 1..dim-2 always have both neighbours — and only
// the ghost-cell layers that lie across the GLOBAL periodic boundary carry an image shift.
HB2_DEVICE int stencil3(int c, int dim, int n, int rc, int k, int& coord, int& shift)
    {
    if (n > 1)
        {
        coord = c - 1 + k;
        shift = 0;
        if (coord == 0 && rc == 0)
            shift = -1;
        else if (coord == dim - 1 && rc == n - 1)
            shift = 1;
        return 3;
        }
    return ring3(c, dim, k, coord, shift);
    }

HB2_DEVICE uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t& total)
    {
    uint32_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1)
        {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off)
            incl += t;
        }
    total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - v;
    }


// position of the q-th (0-based) set bit of m (m < 2^24 and popc(m) > q): binary search on popc
HB2_DEVICE uint32_t nth_set_bit(uint32_t m, uint32_t q)
    {
    uint32_t t = 0, c;
    c = __popc(m & 0xFFFu);
    if (q >= c)
        {
        q -= c;
        t = 12;
        m >>= 12;
        }
    c = __popc(m & 0x3Fu);
    if (q >= c)
        {
        q -= c;
        t += 6;
        m >>= 6;
        }
    c = __popc(m & 0x7u);
    if (q >= c)
        {
        q -= c;
        t += 3;
        m >>= 3;
        }
    c = m & 1u;
    if (q >= c)
        {
        q -= c;
        t += 1;
        m >>= 1;
        }
    c = m & 1u;
    if (q >= c)
        t += 1;
    return t;
    }

// fp32 error model of rsq (DESIGN.md "guard band"): each component of the separation carries at most
// 2^-23 Lmax absolute error (one rounding of a difference of magnitude <= Lmax, one of the image
// shift), so rsq carries <= 2 sqrt(3) r 2^-23 Lmax + 2^-22 rsq. The band is twice that, relative to
// the smallest r_list in use. Pairs inside the band are decided in fp64 like the CPU build.
inline NlConst nl_make_const(const hb2_box& box, double r_list_min, float Lmax)
    {
    NlConst cst;
    cst.bf = boxf(box);
    cst.bd = boxd(box);
    const double rl_min = r_list_min > 0 ? r_list_min : 1.0;
    double band = 2.0 * (2.0 * 1.7320508 * ldexp(1.0, -23) * (double)Lmax / rl_min + ldexp(1.0, -22));
    if (band > 0.25)
        band = 0.25;
    cst.band = (float)band;
    return cst;
    }
    } // namespace hb2
