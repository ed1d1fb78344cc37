// tools/probe/key_map.cpp -- analysis only (CPU): the key iteration (best of the first 14 scanned partitions, 15 = the block is already
// exact) of every 4x4 block of a raw RGBA8 image, one byte per block in raster order -- the launch order of a level, so 256 consecutive
// bytes are one CTA of the opaque kernel.  Input for tools/bin_order.py --anneal.  Host build of the device code (bc7_core.cuh):
//   g++ -std=c++17 -O2 -ffp-contract=off -DVKT_BCN_DEFINE_PARAMS_INIT -Iinclude -o /tmp/key_map tools/probe/key_map.cpp vierkant_b200/csrc/bc7_tables.cpp -lpthread
//   /tmp/key_map level.raw <width> <height> keys.bin        (level.raw: a stbir-filtered level, e.g. oracle resize of a synth texture)
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <thread>
#include "../../vierkant_b200/csrc/bc7_core.cuh"
#include "../../vierkant_b200/csrc/bc7_params.h"
int main(int argc, char **argv)
{
    const int W = atoi(argv[2]), H = atoi(argv[3]);
    std::vector<uint8_t> img(size_t(W) * H * 4);
    FILE *f = fopen(argv[1], "rb"); if(fread(img.data(), 1, img.size(), f) != img.size()) return 1; fclose(f);
    static vkt::Bc7Tables tables; vkt::bc7_tables_build(&tables);
    vkt_bc7_params p; vkt_bc7_params_init_inline(&p);
    vkt::Bc7KernelParams kp; if(vkt::bc7_prepare_params(&p, &kp)) return 2;
    kp.opt7 = &tables.opt7[0][0];
    const int bx = W / 4, by = H / 4;
    std::vector<uint8_t> keys(size_t(bx) * by);
    auto work = [&](int y0, int y1) {
        for(int y = y0; y < y1; ++y) for(int x = 0; x < bx; ++x)
        {
            vkt::Texel col[16];
            for(int i = 0; i < 16; ++i) memcpy(&col[i].px, &img[(size_t(y * 4 + i / 4) * W + x * 4 + i % 4) * 4], 4);
            vkt::Lane<1> L{col};
            vkt::prepare_lane<1>(L);
            uint64_t best = vkt::kNoErr; int best_it = 0;
            for(int it = 0; it < 14; ++it)
            {
                if(best == 0) break;
                const uint32_t part = tables.order[it];
                const uint64_t e = vkt::estimate_pair<false, true, vkt::kKvKey28, true, 1>(tables, kp, L, part, best);
                if(e < best) { best = e, best_it = it; }
            }
            keys[size_t(y) * bx + x] = uint8_t(best == 0 ? 15 : best_it);
        }
    };
    std::vector<std::thread> th; const int T = 16;
    for(int t = 0; t < T; ++t) th.emplace_back(work, by * t / T, by * (t + 1) / T);
    for(auto &t: th) t.join();
    f = fopen(argv[4], "wb"); fwrite(keys.data(), 1, keys.size(), f); fclose(f);
    return 0;
}
