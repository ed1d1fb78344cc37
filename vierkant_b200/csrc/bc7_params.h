// bc7_params.h -- validation / preprocessing of the public vkt_bc7_params into the kernel's Bc7KernelParams.
// Host-only, header-only; shared by the C-ABI shim (bcn_cuda.cu) and the host emulation test harness.
#pragma once
#include "../../include/vierkant_bcn_cuda.h"
#include "bc7_core.cuh"

inline void vkt_bc7_params_init_inline(vkt_bc7_params *p)
{
    // bc7enc_compress_block_params_init, /root/reference/extern/bc7enc_rdo/bc7enc.h:95-113
    *p = vkt_bc7_params{};
    p->mode_mask = 0xFFFFFFFFu;
    p->max_partitions = 64;
    p->weights[0] = 128, p->weights[1] = 64, p->weights[2] = 16, p->weights[3] = 32;
    p->uber_level = 0;
    p->perceptual = 1;
    p->try_least_squares = 1;
    p->mode17_partition_estimation_filterbank = 1;
    p->pbit1_weight = 1.0f;
    p->mode1_error_weight = p->mode5_error_weight = p->mode6_error_weight = p->mode7_error_weight = 1.0f;
    p->low_frequency_partition_weight = 1.0f;
}

#ifdef VKT_BCN_DEFINE_PARAMS_INIT
extern "C" void vkt_bc7_params_init(vkt_bc7_params *p) { vkt_bc7_params_init_inline(p); }
#endif

namespace vkt
{

// Returns VKT_BCN_OK or VKT_BCN_ERR_INVALID.  Weight setup follows bc7enc.cpp:2409-2420 (float constant expressions
// evaluated in float, truncating conversion) -- compile with -ffp-contract=off.
inline int bc7_prepare_params(const vkt_bc7_params *p, Bc7KernelParams *k)
{
    if(p->uber_level > 4) { return VKT_BCN_ERR_INVALID; }
    // The knobs bc7enc_rdo's RDO post-processor drives (vierkant never sets them): served by the extended kernel variant.
    k->force_selectors = p->force_selectors != 0;
    k->quant_mode6 = p->quant_mode6_endpoints != 0;
    k->low_freq_weight = p->low_frequency_partition_weight;
    k->forced_sel = 0;
    k->m6_reduced = nullptr;// per device: set by the launcher
    if(k->force_selectors)
    {
        // A forced selector indexes the palette of whichever mode is being tried (bc7enc.cpp:701-707); beyond that palette the
        // reference reads an uninitialised colour.  Every selector must therefore exist in every mode the mask lets run.
        uint32_t limit = 16;
        if((p->mode_mask & (1u << 1)) && p->max_partitions > 0) { limit = 8; }
        if(p->mode_mask & ((1u << 5) | (1u << 7))) { limit = 4; }
        for(int i = 0; i < 16; ++i)
        {
            if(p->selectors[i] >= limit) { return VKT_BCN_ERR_INVALID; }
            k->forced_sel |= (uint64_t) p->selectors[i] << (4 * i);
        }
    }
    // (uint64_t)((double)err * weight + .5f), bc7enc.cpp:1819, is only defined for a finite, non-negative product in range
    if(!(p->low_frequency_partition_weight >= 0.0f && p->low_frequency_partition_weight <= 65536.0f)) { return VKT_BCN_ERR_INVALID; }
    k->ext = (k->force_selectors || k->quant_mode6 || p->low_frequency_partition_weight != 1.0f) ? 1u : 0u;
    const bool alpha_modes = (p->mode_mask & ((1u << 5) | (1u << 6) | (1u << 7))) != 0;
    // an opaque block needs mode 6, or mode 1 WITH partitions to try: with mode 6 masked out and max_partitions == 0 the reference
    // skips mode 1 as well (bc7enc.cpp:2336) and encodes an uninitialised result -- undefined there, refused here (found by the
    // parameter fuzz of tests/test_bc7_gpu.py)
    const bool opaque_modes = (p->mode_mask & (1u << 6)) != 0 || ((p->mode_mask & (1u << 1)) != 0 && p->max_partitions > 0);
    if(!alpha_modes || !opaque_modes) { return VKT_BCN_ERR_INVALID; }// the reference asserts (bc7enc.cpp:2141,2295)
    k->mode_mask = p->mode_mask;
    k->max_partitions = p->max_partitions;
    if(p->perceptual)
    {
        const float pr_weight = (.5f / (1.0f - .2126f)) * (.5f / (1.0f - .2126f));
        const float pb_weight = (.5f / (1.0f - .0722f)) * (.5f / (1.0f - .0722f));
        k->w[0] = (uint32_t) (int) (p->weights[0] * 4.0f);
        k->w[1] = (uint32_t) (int) (p->weights[1] * 4.0f * pr_weight);
        k->w[2] = (uint32_t) (int) (p->weights[2] * 4.0f * pb_weight);
        k->w[3] = p->weights[3] * 4;
    }
    else
    {
        for(int i = 0; i < 4; ++i) { k->w[i] = p->weights[i]; }
    }
    {
        // largest possible per-texel error: |dl| <= 510, |dcr| <= 803, |dcb| <= 947 (YCbCr ranges of 8-bit colours, >> 8),
        // |da| <= 255; linear metric: 255 per channel.  Below 2^28 the kernel may fuse error and selector in one key.
        const uint64_t b = p->perceptual ? (uint64_t) k->w[0] * 510 * 510 + (uint64_t) k->w[1] * 803 * 803 + (uint64_t) k->w[2] * 947 * 947 +
                                                   (uint64_t) k->w[3] * 255 * 255
                                         : ((uint64_t) k->w[0] + k->w[1] + k->w[2] + k->w[3]) * 255 * 255;
        k->key28 = (b < (1ull << 28)) ? 1u : 0u;
    }
    for(int i = 0; i < 4; ++i) { k->w16[i] = k->w[i] * 16u; }// only read when key28 (then w * 16 * d^2 < 2^32)
    k->uber_level = p->uber_level;
    k->try_least_squares = p->try_least_squares != 0;
    k->filterbank = p->mode17_partition_estimation_filterbank != 0;
    k->force_alpha = p->force_alpha != 0;
    k->bias_mode1_pbits = p->bias_mode1_pbits != 0;
    k->pbit1_weight = p->pbit1_weight;
    k->mode1_w = p->mode1_error_weight;
    k->mode5_w = p->mode5_error_weight;
    k->mode6_w = p->mode6_error_weight;
    k->mode7_w = p->mode7_error_weight;
    return VKT_BCN_OK;
}

}// namespace vkt
