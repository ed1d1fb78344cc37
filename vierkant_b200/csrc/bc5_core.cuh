// bc5_core.cuh -- BC5 (two BC4 halves: R and G) block encoder, one lane per 4x4 block.
//
// Same results as rgbcx::encode_bc5(pDst, pixels, 0, 1, 4) (/root/reference/extern/bc7enc_rdo/rgbcx.cpp:2913 ->
// encode_bc4 :2608-2728), the BC5 branch of vierkant::bcn::compress (src/texture_block_compression.cpp:131).
// Integer only: endpoints = (max, min) of the channel, 3-bit selectors from seven thresholds on 14*(v - min) + 4.
// The kernel is HBM-bound (64 B in, 16 B out per block): 128-bit coalesced loads, one 128-bit store per lane.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VKT_BC5_FN __host__ __device__ __forceinline__
#else
#define VKT_BC5_FN inline
#endif

namespace vkt
{

// px[i] = packed RGBA texel i; channel = 0 (R) or 1 (G).  Returns the 8-byte BC4 block, little endian.
VKT_BC5_FN uint64_t bc4_encode_channel(const uint32_t px[16], int channel)
{
    uint32_t v[16];
    uint32_t mn = 255, mx = 0;
#pragma unroll
    for(int i = 0; i < 16; ++i)
    {
        v[i] = (px[i] >> (8 * channel)) & 255u;
        mn = v[i] < mn ? v[i] : mn;
        mx = v[i] > mx ? v[i] : mx;
    }
    uint64_t blk = uint64_t(mx) | (uint64_t(mn) << 8);
    if(mx == mn) { return blk; }
    const int delta = int(mx - mn);
    const int bias = 4 - int(mn) * 14;
    uint64_t sel = 0;
#pragma unroll
    for(int i = 0; i < 16; ++i)
    {
        const int x = int(v[i]) * 14 + bias;
        // number of thresholds delta*{13,11,9,7,5,3,1} reached (rgbcx.cpp:2655-2683)
        const int k = (x >= delta * 13) + (x >= delta * 11) + (x >= delta * 9) + (x >= delta * 7) + (x >= delta * 5) + (x >= delta * 3) + (x >= delta);
        // translation {1,7,6,5,4,3,2,0}: k == 0 -> 1, k == 7 -> 0, else 8 - k
        const uint32_t s = (k == 0) ? 1u : (k == 7) ? 0u : uint32_t(8 - k);
        sel |= uint64_t(s) << (3 * i);
    }
    return blk | (sel << 16);
}

#if defined(__CUDACC__)
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
    const uint64_t r = bc4_encode_channel(px, 0), g = bc4_encode_channel(px, 1);
    out[b] = make_uint4(uint32_t(r), uint32_t(r >> 32), uint32_t(g), uint32_t(g >> 32));
}
#endif

}// namespace vkt
