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

#include "../../vierkant_b200/csrc/bc5_core.cuh"
#include "../../vierkant_b200/csrc/bc7_core.cuh"
#include "../../vierkant_b200/csrc/bc7_params.h"
#include "../../vierkant_b200/csrc/chain_plan.h"
#include "../../vierkant_b200/csrc/resize_axis.h"
#include "../../vierkant_b200/csrc/resize_strip.h"

// the uber-free instantiation whenever the parameters allow it, as the CUDA dispatch does
template<bool PERC, int KV, bool ALPHA>
static void enc(const vkt::Bc7Tables &tables, const vkt::Bc7KernelParams &kp, vkt::Lane<1> lane, uint32_t blk[4])
{
    if(kp.uber_level == 0 && KV == vkt::kKvKey28) { vkt::encode_block<PERC, KV, ALPHA, false, 1>(tables, kp, lane, blk); }
    else { vkt::encode_block<PERC, KV, ALPHA, true, 1>(tables, kp, lane, blk); }
}

extern "C" {

// The multi-device plan of one compress() chain (chain_plan.h) for device g of G, with the real stbir tap ranges.
// level_height: the chain's level heights (level 0 = rounded base height); src_height: the source image's height.
// Outputs (per level l < 16): own block rows, needed pixel rows, and for l == 0 the source rows the need range reads.
// Returns sliced-level count M in the low byte, participating devices in the next byte, or -1 if need[] does not cover
// the taps of the next level (the property the plan must guarantee).
int emul_chain_plan(const uint32_t *level_height, uint32_t num_levels, uint32_t src_height, uint32_t G, uint32_t g, uint32_t *own0, uint32_t *own1,
                    uint32_t *need0, uint32_t *need1, uint32_t *src0, uint32_t *src1)
{
    const vkt::ChainSplit split = vkt::chain_split(level_height, num_levels, G);
    if(g >= split.devices) { return int(split.sliced | (split.devices << 8)); }
    std::vector<vkt::ResizeAxis> ax(split.sliced);
    std::vector<std::vector<int>> first(split.sliced), last(split.sliced);
    std::vector<vkt::RowTaps> taps(split.sliced);
    for(uint32_t l = 0; l < split.sliced; ++l)
    {
        const int in = int(l ? level_height[l - 1] : src_height), out = int(level_height[l]);
        ax[l].build(in, out);
        first[l].resize(size_t(out)), last[l].resize(size_t(out));
        for(int o = 0; o < out; ++o)
        {
            int lo = in - 1, hi = 0;
            for(int t = ax[l].start[size_t(o)]; t < ax[l].start[size_t(o) + 1]; ++t) { lo = std::min(lo, ax[l].idx[size_t(t)]), hi = std::max(hi, ax[l].idx[size_t(t)]); }
            if(ax[l].start[size_t(o)] == ax[l].start[size_t(o) + 1]) { lo = hi = 0; }
            first[l][size_t(o)] = lo, last[l][size_t(o)] = hi;
        }
        taps[l] = {first[l].data(), last[l].data()};
    }
    const vkt::DeviceRows r = vkt::device_rows(split, level_height, taps.data(), g);
    for(uint32_t l = 0; l < split.sliced; ++l)
    {
        own0[l] = r.own[l].first, own1[l] = r.own[l].second, need0[l] = r.need[l].first, need1[l] = r.need[l].second;
        if(l + 1 < split.sliced)
        {
            for(uint32_t y = r.need[l + 1].first; y < r.need[l + 1].second; ++y)
            {
                if(uint32_t(first[l + 1][y]) < r.need[l].first || uint32_t(last[l + 1][y]) >= r.need[l].second) { return -1; }
            }
        }
    }
    if(split.sliced)
    {
        const auto s = vkt::source_rows(taps[0], r.need[0].first, r.need[0].second, src_height);
        *src0 = s.first, *src1 = s.second;
    }
    return int(split.sliced | (split.devices << 8));
}

// number of places where a tap list of ResizeAxis(in, out) steps back to a smaller input sample, or where the first /
// last taps of consecutive outputs do (the vertical CUDA pass walks input rows in ascending order and relies on both)
int emul_resize_axis_order_violations(int in, int out)
{
    vkt::ResizeAxis a;
    a.build(in, out);
    int bad = 0, prev_first = -1, prev_last = -1;
    for(int o = 0; o < out; ++o)
    {
        const int t0 = a.start[size_t(o)], t1 = a.start[size_t(o) + 1];
        for(int t = t0 + 1; t < t1; ++t) { bad += a.idx[size_t(t)] < a.idx[size_t(t) - 1]; }
        if(t1 > t0)
        {
            bad += a.idx[size_t(t0)] < prev_first;
            bad += a.idx[size_t(t1) - 1] < prev_last;
            prev_first = a.idx[size_t(t0)], prev_last = a.idx[size_t(t1) - 1];
        }
    }
    return bad;
}

// number of u8 values whose division-free decode (resize_decode_u8) differs from the reference's v / 255.0f
int emul_resize_decode_mismatches()
{
    int bad = 0;
    for(uint32_t v = 0; v < 256; ++v)
    {
        volatile float want = (float) v / 255.0f;
        bad += vkt::resize_decode_u8(v) != want;
    }
    return bad;
}

// the fused pass's two-instruction decode (resize_decode_split) against v / 255.0f, and its float(v) from the 2^23 + v bit pattern
int emul_resize_decode_split_mismatches()
{
    int bad = 0;
    for(uint32_t v = 0; v < 256; ++v)
    {
        volatile float want = (float) v / 255.0f;
        const uint32_t bits = 0x4B000000u | v;
        float m;
        std::memcpy(&m, &bits, sizeof(m));
        volatile float fv = m - 8388608.0f;
        bad += (fv != (float) v) || (vkt::resize_decode_split(fv) != want);
    }
    return bad;
}

// the conversion-free encode (resize_encode_u8) against (int)((double) s + 0.5) for every `stride`-th float in [0, 255]
// (stride 1: all 1.13e9 of them, a few seconds)
long long emul_resize_encode_mismatches(uint32_t stride)
{
    long long bad = 0;
    const float top = 255.0f;
    uint32_t hi;
    std::memcpy(&hi, &top, sizeof(hi));
    for(uint64_t b = 0; b <= hi; b += stride ? stride : 1)
    {
        const uint32_t bits = (uint32_t) b;
        float s;
        std::memcpy(&s, &bits, sizeof(s));
        bad += vkt::resize_encode_u8(s) != (uint32_t) (int) ((double) s + 0.5);
    }
    return bad;
}

// BC4 selector count: the reciprocal multiplication of bc5_core.cuh against the seven threshold compares of rgbcx
// (rgbcx.cpp:2655-2683), for every delta and every reachable numerator.  Returns the number of disagreements.
int emul_bc4_count_mismatches()
{
    int bad = 0;
    for(uint32_t delta = 1; delta <= 255; ++delta)
    {
        const uint32_t recip = vkt::bc4_reciprocal(delta);
        for(uint32_t x = 0; x <= 14u * 255u + 4u; ++x)
        {
            const int xi = int(x), d = int(delta);
            const uint32_t want = uint32_t((xi >= d * 13) + (xi >= d * 11) + (xi >= d * 9) + (xi >= d * 7) + (xi >= d * 5) + (xi >= d * 3) + (xi >= d));
            bad += vkt::bc4_count(x + delta, recip) != want;
        }
    }
    return bad;
}

// the device-side BC5 block encoder as host code: tiles (n, 16, 4) -> (n, 16) bytes
void emul_bc5_encode_blocks(const uint8_t *px, uint64_t num_blocks, uint8_t *out)
{
    for(uint64_t b = 0; b < num_blocks; ++b)
    {
        uint32_t t[16];
        memcpy(t, px + 64 * b, 64);
        const uint64_t r = vkt::bc4_encode_channel(t, 0), g = vkt::bc4_encode_channel(t, 1);
        memcpy(out + 16 * b, &r, 8);
        memcpy(out + 16 * b + 8, &g, 8);
    }
}

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
    static uint8_t m6_reduced[vkt::kBc7M6ReducedBytes];
    static bool once6 = (vkt::bc7_m6_reduced_build(m6_reduced), true);
    (void) once6;
    kp.m6_reduced = m6_reduced;
    kp.opt7 = &tables.opt7[0][0];
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
            int sel = (perceptual ? 4 : 0) | (kp.key28 ? 2 : 0) | (alpha ? 1 : 0);
            if(kp.ext) { sel = 8 | (perceptual ? 2 : 0) | (alpha ? 1 : 0); }// the extended variant, as the CUDA dispatch picks it
            switch(sel)
            {
                case 11: enc<true, vkt::kKvExt, true>(tables, kp, lane, blk); break;
                case 10: enc<true, vkt::kKvExt, false>(tables, kp, lane, blk); break;
                case 9: enc<false, vkt::kKvExt, true>(tables, kp, lane, blk); break;
                case 8: enc<false, vkt::kKvExt, false>(tables, kp, lane, blk); break;
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

// The resize strip kernels (resize_strip.h: the body the CUDA kernels run per thread) executed thread by thread on the host:
// RGBA image w x h -> ow x oh (ow == w, oh == h or w == 2 ow, h == 2 oh), output rows [y0, y1) in strips of `strip` rows, exactly as
// resize_device launches them (one call per row range).  Rows outside [y0, y1) of `out` are left untouched.
// Returns 0, or -1 when the axes are not regular / uniform (the product would take another path).
int emul_resize_strip(const uint8_t *in, uint32_t w, uint32_t h, uint8_t *out, uint32_t ow, uint32_t oh, uint32_t y0, uint32_t y1, int strip)
{
    vkt::ResizeAxis ax, ay;
    ax.build((int) w, (int) ow);
    ay.build((int) h, (int) oh);
    const int S = (w == ow) ? 1 : 2, T = (S == 1) ? 3 : 8, NC = (S == 1) ? 4 : 2;
    if((w != ow && w != 2 * ow) || (h != oh && h != 2 * oh) || (w == ow) != (h == oh) || ow % uint32_t(NC) || w % 4u || strip <= 0 || y1 > oh) { return -1; }
    vkt::FusedCoef cx = {}, cy = {};
    for(const vkt::ResizeAxis *a: {&ax, &ay})
    {
        const int in_n = a->in_size, out_n = a->out_size;
        for(int o = 0; o < out_n; ++o)
        {
            if(a->start[size_t(o) + 1] - a->start[size_t(o)] != T) { return -1; }
            for(int t = 0; t < T; ++t)
            {
                const int v = S * o - (T - S) / 2 + t, want = v < 0 ? 0 : (v >= in_n ? in_n - 1 : v);
                if(a->idx[size_t(a->start[size_t(o)] + t)] != want) { return -1; }
                if(memcmp(&a->coef[size_t(a->start[size_t(o)] + t)], &a->coef[size_t(a->start[0] + t)], sizeof(float)) != 0) { return -1; }
            }
        }
        for(int t = 0; t < T; ++t) { (a == &ax ? cx : cy).c[t] = a->coef[size_t(a->start[0] + t)]; }
    }
    const int groups = int(ow) / NC, strips = (int(y1 - y0) + strip - 1) / strip;
    // (a CTA is 128 threads: run the surplus threads of the last CTA as well -- they must return without touching anything)
    const int threads = (groups + 127) / 128 * 128;
    for(int by = 0; by < strips; ++by)
    {
        for(int k = 0; k < threads; ++k)
        {
            if(S == 1) { vkt::resize_strip_thread<1, 3, 4>(in, int(w), int(h), int(ow), int(y0), int(y1), strip, cx, cy, out, k, by); }
            else { vkt::resize_strip_thread<2, 8, 2>(in, int(w), int(h), int(ow), int(y0), int(y1), strip, cx, cy, out, k, by); }
        }
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
