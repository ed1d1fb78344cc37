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

extern "C" {

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
                case 7: vkt::encode_block<true, true, true, 1>(tables, kp, lane, blk); break;
                case 6: vkt::encode_block<true, true, false, 1>(tables, kp, lane, blk); break;
                case 5: vkt::encode_block<true, false, true, 1>(tables, kp, lane, blk); break;
                case 4: vkt::encode_block<true, false, false, 1>(tables, kp, lane, blk); break;
                case 3: vkt::encode_block<false, true, true, 1>(tables, kp, lane, blk); break;
                case 2: vkt::encode_block<false, true, false, 1>(tables, kp, lane, blk); break;
                case 1: vkt::encode_block<false, false, true, 1>(tables, kp, lane, blk); break;
                default: vkt::encode_block<false, false, false, 1>(tables, kp, lane, blk); break;
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
}
