// tests/host_emul/emul.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the device-side block search (vierkant_b200/csrc/bc7_core.cuh) as plain host C++ (-ffp-contract=off) so the
// search logic can be diffed against the reference on a machine without a GPU.  Not linked into the product library;
// the product has no CPU path.  Warp ballots degenerate to per-lane predicates here, which is the only behavioural
// difference from the GPU build (it affects work skipping, never results).
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../vierkant_b200/csrc/bc7_core.cuh"
#include "../../vierkant_b200/csrc/bc7_params.h"
#include "../../vierkant_b200/csrc/resize_axis.h"

// the uber-free instantiation whenever the parameters allow it, as the CUDA dispatch does
template<bool PERC, bool KEY28, bool ALPHA>
static void enc(const vkt::Bc7Tables &tables, const vkt::Bc7KernelParams &kp, vkt::Lane<1> lane, uint32_t blk[4])
{
    if(kp.uber_level == 0) { vkt::encode_block<PERC, KEY28, ALPHA, false, 1>(tables, kp, lane, blk); }
    else { vkt::encode_block<PERC, KEY28, ALPHA, true, 1>(tables, kp, lane, blk); }
}

extern "C" {

// number of (max, ly, hy) cells whose compile-time uber selector map differs from the reference's float expression
int emul_uber_map_mismatches()
{
    const vkt::UberMaps um = vkt::make_uber_maps();
    int bad = 0;
    for(int k = 0; k < 3; ++k)
    {
        const int max_sel = (k == 0) ? 3 : (k == 1) ? 7 : 15;
        const uint64_t live = (max_sel == 15) ? ~0ull : ((1ull << (4 * (max_sel + 1))) - 1ull);
        for(int ly = -2; ly <= 1; ++ly)
        {
            for(int hy = max_sel - 1; hy <= max_sel + 2; ++hy)
            {
                bad += (um.m[k][ly + 2][hy - (max_sel - 1)] & live) != (vkt::bc7_uber_map_reference(max_sel, ly, hy) & live);
            }
        }
    }
    return bad;
}

int emul_bc7_encode_blocks(const uint8_t *px, uint64_t num_blocks, const vkt_bc7_params *params, uint8_t *out, int threads)
{
    static vkt::Bc7Tables tables;
    static bool once = (vkt::bc7_tables_build(&tables), true);
    (void) once;
    vkt::Bc7KernelParams kp;
    vkt_bc7_params def;
    if(!params)
    {
        vkt_bc7_params_init(&def);
        params = &def;
    }
    int rc = vkt::bc7_prepare_params(params, &kp);
    if(rc) { return rc; }
    const bool perceptual = params->perceptual != 0;
    auto work = [&](uint64_t b0, uint64_t b1) {
        vkt::Texel column[16];
        uint32_t blk[4];
        for(uint64_t b = b0; b < b1; ++b)
        {
            for(int i = 0; i < 16; ++i) { memcpy(&column[i].px, px + 64 * b + 4 * i, 4); }
            vkt::Lane<1> lane{column};
            uint32_t raw[16];
            memcpy(raw, px + 64 * b, 64);
            const bool alpha = vkt::block_has_alpha(kp, raw);
            const int sel = (perceptual ? 4 : 0) | (kp.key28 ? 2 : 0) | (alpha ? 1 : 0);
            switch(sel)
            {
                case 7: enc<true, true, true>(tables, kp, lane, blk); break;
                case 6: enc<true, true, false>(tables, kp, lane, blk); break;
                case 5: enc<true, false, true>(tables, kp, lane, blk); break;
                case 4: enc<true, false, false>(tables, kp, lane, blk); break;
                case 3: enc<false, true, true>(tables, kp, lane, blk); break;
                case 2: enc<false, true, false>(tables, kp, lane, blk); break;
                case 1: enc<false, false, true>(tables, kp, lane, blk); break;
                default: enc<false, false, false>(tables, kp, lane, blk); break;
            }
            memcpy(out + 16 * b, blk, 16);
        }
    };
    if(threads <= 1) { work(0, num_blocks); }
    else
    {
        std::vector<std::thread> pool;
        for(int t = 0; t < threads; ++t)
        {
            pool.emplace_back(work, num_blocks * t / threads, num_blocks * (t + 1) / threads);
        }
        for(auto &t: pool) { t.join(); }
    }
    return 0;
}

// The product's host-built tap lists (resize_axis.h) evaluated on the CPU in the order the two CUDA passes use.
int emul_resize_u8(const uint8_t *in, uint32_t w, uint32_t h, uint32_t comps, uint8_t *out, uint32_t ow, uint32_t oh)
{
    vkt::ResizeAxis ax, ay;
    ax.build((int) w, (int) ow);
    ay.build((int) h, (int) oh);
    const size_t row = size_t(ow) * comps;
    std::vector<float> band(size_t(h) * row, 0.0f);
    for(uint32_t r = 0; r < h; ++r)
    {
        for(uint32_t x = 0; x < ow; ++x)
        {
            for(uint32_t c = 0; c < comps; ++c)
            {
                float acc = 0.0f;
                for(int t = ax.start[x]; t < ax.start[x + 1]; ++t)
                {
                    acc = acc + (((float) in[(size_t(r) * w + size_t(ax.idx[size_t(t)])) * comps + c]) / 255.0f) * ax.coef[size_t(t)];
                }
                band[size_t(r) * row + size_t(x) * comps + c] = acc;
            }
        }
    }
    for(uint32_t y = 0; y < oh; ++y)
    {
        for(size_t i = 0; i < row; ++i)
        {
            float acc = 0.0f;
            for(int t = ay.start[y]; t < ay.start[y + 1]; ++t) { acc = acc + band[size_t(ay.idx[size_t(t)]) * row + i] * ay.coef[size_t(t)]; }
            float f = acc < 0.0f ? 0.0f : (acc > 1.0f ? 1.0f : acc);
            const float s = f * 255.0f;
            const int q = (int) s;
            out[size_t(y) * row + i] = (uint8_t) (q + ((s - (float) q) >= 0.5f ? 1 : 0));
        }
    }
    return 0;
}
}
