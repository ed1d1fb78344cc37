// bc5_core.cuh -- BC5 (two BC4 halves: R and G) block encoder, one lane per 4x4 block.
//
// Same results as rgbcx::encode_bc5(pDst, pixels, 0, 1, 4) (/root/reference/extern/bc7enc_rdo/rgbcx.cpp:2913 ->
// encode_bc4 :2608-2728), the BC5 branch of vierkant::bcn::compress (src/texture_block_compression.cpp:131).
// Integer only: endpoints = (max, min) of the channel, 3-bit selectors from seven thresholds on 14*(v - min) + 4.
// 64 B in, 16 B out per block: 128-bit coalesced loads, one 128-bit store per lane.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VKT_BC5_FN __host__ __device__ __forceinline__
#else
#define VKT_BC5_FN inline
#endif

namespace vkt
{

// Selector arithmetic.  rgbcx counts how many of the thresholds delta * {13, 11, 9, 7, 5, 3, 1} the value x = 14 (v - min) + 4
// reaches (rgbcx.cpp:2655-2683) and translates the count through {1, 7, 6, 5, 4, 3, 2, 0}.  x >= delta (2m - 1) is
// x + delta >= 2 delta m, so the count is min(7, (x + delta) / (2 delta)) -- one division instead of seven compares -- and the
// division is a multiplication by a rounded-up 21-bit reciprocal: exact for every numerator below 2^12 and divisor below 2^9
// (here x + delta <= 3829, 2 delta <= 510; all 255 x 3830 cases are checked in tests/test_host_emul.py), and the product stays
// below 2^32.  Seven instructions per texel and channel instead of about twenty: the kernel was bound by them, not by HBM.
VKT_BC5_FN uint32_t bc4_reciprocal(uint32_t delta) { return ((1u << 21) + 2u * delta - 1u) / (2u * delta); }// ceil(2^21 / (2 delta)), delta >= 1
VKT_BC5_FN uint32_t bc4_count(uint32_t x_plus_delta, uint32_t recip)
{
    const uint32_t q = (x_plus_delta * recip) >> 21;
    return q < 7u ? q : 7u;
}

// px[i] = packed RGBA texel i; channel = 0 (R) or 1 (G).  Returns the 8-byte BC4 block, little endian.
VKT_BC5_FN uint64_t bc4_encode_channel(const uint32_t px[16], int channel)
{
    uint32_t mn = 255, mx = 0;
#pragma unroll
    for(int i = 0; i < 16; ++i)
    {
        const uint32_t v = (px[i] >> (8 * channel)) & 255u;
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
    }
    const uint64_t blk = uint64_t(mx) | (uint64_t(mn) << 8);
    if(mx == mn) { return blk; }
    const uint32_t delta = mx - mn;
    const uint32_t recip = bc4_reciprocal(delta);
    const uint32_t bias = 4u + delta - mn * 14u;// x + delta = 14 v + bias (wraps like the int arithmetic of the reference: the sum is in range)
    uint32_t lo = 0, hi = 0;// selectors of texels 0..7 / 8..15, three bits each
#pragma unroll
    for(int i = 0; i < 16; ++i)
    {
        const uint32_t v = (px[i] >> (8 * channel)) & 255u;
        const uint32_t k = bc4_count(v * 14u + bias, recip);
        const uint32_t s = (0x02345671u >> (4u * k)) & 7u;// translation {1,7,6,5,4,3,2,0}: one nibble per count
        if(i < 8) { lo |= s << (3 * i); }
        else { hi |= s << (3 * (i - 8)); }
    }
    return blk | (uint64_t(lo) << 16) | (uint64_t(hi) << 40);
}

#if defined(__CUDACC__)
// Both channels of a block at once, R in the low and G in the high 16 bits of a word: one PRMT per texel unpacks both, the
// minimum / maximum are 16x2 SIMD (VIMNMX.U16x2) and 14 v + bias is one IMAD for the pair (no carry: at most 3829 per half).
// Same arithmetic as bc4_encode_channel, which stays the reference for the host build and for the tests.
__device__ __forceinline__ uint4 bc5_encode_block_rg(const uint32_t px[16])
{
    uint32_t rg[16];
    uint32_t mn = 0x00FF00FFu, mx = 0u;
#pragma unroll
    for(int i = 0; i < 16; ++i)
    {
        rg[i] = __byte_perm(px[i], 0u, 0x4140);// r | g << 16
        mn = __vminu2(mn, rg[i]), mx = __vmaxu2(mx, rg[i]);
    }
    const uint32_t mn_r = mn & 0xFFFFu, mn_g = mn >> 16, mx_r = mx & 0xFFFFu, mx_g = mx >> 16;
    const uint32_t d_r = mx_r - mn_r, d_g = mx_g - mn_g;
    // a flat channel keeps all selectors 0 (rgbcx.cpp:2631-2640): a reciprocal of 0 makes every count 0 and the translation is
    // bypassed with the channel's mask below
    const uint32_t rc_r = d_r ? bc4_reciprocal(d_r) << 11 : 0u, rc_g = d_g ? bc4_reciprocal(d_g) << 11 : 0u;// mulhi(x, recip << 11) == x * recip >> 21
    // 14 v + bias per half = 14 (v - min) + 4 + delta, in [4 + delta, 3829].  The two biases may be negative on their own; added as
    // ONE 32-bit number (wrap-around arithmetic) the sum 14 rg + bias is exactly X_r + 65536 X_g with both X in range, so no half
    // borrows from the other
    const uint32_t bias = (4u + d_r - mn_r * 14u) + ((4u + d_g - mn_g * 14u) << 16);
    uint32_t lo_r = 0, hi_r = 0, lo_g = 0, hi_g = 0;
#pragma unroll
    for(int i = 0; i < 16; ++i)
    {
        const uint32_t x = rg[i] * 14u + bias;
        const uint32_t k_r = min(__umulhi(x & 0xFFFFu, rc_r), 7u), k_g = min(__umulhi(x >> 16, rc_g), 7u);
        const uint32_t s_r = (0x02345671u >> (4u * k_r)) & 7u, s_g = (0x02345671u >> (4u * k_g)) & 7u;
        if(i < 8) { lo_r |= s_r << (3 * i), lo_g |= s_g << (3 * i); }
        else { hi_r |= s_r << (3 * (i - 8)), hi_g |= s_g << (3 * (i - 8)); }
    }
    if(!d_r) { lo_r = hi_r = 0; }
    if(!d_g) { lo_g = hi_g = 0; }
    const uint64_t r = uint64_t(mx_r) | (uint64_t(mn_r) << 8) | (uint64_t(lo_r) << 16) | (uint64_t(hi_r) << 40);
    const uint64_t g = uint64_t(mx_g) | (uint64_t(mn_g) << 8) | (uint64_t(lo_g) << 16) | (uint64_t(hi_g) << 40);
    return make_uint4(uint32_t(r), uint32_t(r >> 32), uint32_t(g), uint32_t(g >> 32));
}

__global__ void __launch_bounds__(256) bc5_encode_kernel(const uint8_t *__restrict__ img, uint32_t blocks_x, uint32_t num_blocks,
                                                         uint32_t comps, uint32_t stride, uint4 *__restrict__ out)
{
    const uint32_t b = blockIdx.x * 256 + threadIdx.x;
    if(b >= num_blocks) { return; }
    const uint32_t bx = b % blocks_x, by = b / blocks_x;
    uint32_t px[16];
    const bool vec16 = (comps == 4) && ((stride & 15u) == 0) && ((reinterpret_cast<uintptr_t>(img) & 15u) == 0);
    if(vec16)
    {
#pragma unroll
        for(int y = 0; y < 4; ++y)
        {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(img + size_t(by * 4 + y) * stride) + bx);
            px[4 * y + 0] = v.x, px[4 * y + 1] = v.y, px[4 * y + 2] = v.z, px[4 * y + 3] = v.w;
        }
    }
    else
    {
#pragma unroll
        for(int y = 0; y < 4; ++y)
        {
            const uint8_t *row = img + size_t(by * 4 + y) * stride + size_t(bx) * 4 * comps;
#pragma unroll
            for(int x = 0; x < 4; ++x) { px[4 * y + x] = uint32_t(row[x * comps]) | (uint32_t(row[x * comps + 1]) << 8); }
        }
    }
    out[b] = bc5_encode_block_rg(px);
}
#endif

}// namespace vkt
