// bc7_tables.cpp -- host-side construction of vkt::Bc7Tables (runs once per context, then uploaded to the GPU).
//
// Same arithmetic as bc7enc_compress_block_init() (/root/reference/extern/bc7enc_rdo/bc7enc.cpp:124-285): float
// midpoints via IEEE division by 255.0f, exhaustive search for the optimal single-colour endpoint pairs of mode 1
// (6 bits + shared p-bit, selector 2 of 8) and mode 7 (5 bits + two p-bits, selector 1 of 4).
// Build with -ffp-contract=off (the midpoints are float expressions).
#include "bc7_tables.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace vkt
{

namespace
{
constexpr uint16_t k_part2[64] = {
        0xCCCC, 0x8888, 0xEEEE, 0xECC8, 0xC880, 0xFEEC, 0xFEC8, 0xEC80, 0xC800, 0xFFEC, 0xFE80, 0xE800, 0xFFE8,
        0xFF00, 0xFFF0, 0xF000, 0xF710, 0x008E, 0x7100, 0x08CE, 0x008C, 0x7310, 0x3100, 0x8CCE, 0x088C, 0x3110,
        0x6666, 0x366C, 0x17E8, 0x0FF0, 0x718E, 0x399C, 0xAAAA, 0xF0F0, 0x5A5A, 0x33CC, 0x3C3C, 0x55AA, 0x9696,
        0xA55A, 0x73CE, 0x13C8, 0x324C, 0x3BDC, 0x6996, 0xC33C, 0x9966, 0x0660, 0x0272, 0x04E4, 0x4E40, 0x2720,
        0xC936, 0x936C, 0x39C6, 0x639C, 0x9336, 0x9CC6, 0x817E, 0xE718, 0xCCF0, 0x0FCC, 0x7744, 0xEE22};
constexpr uint8_t k_anchor2[64] = {15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 2,  8,  2,  2, 8,
                                   8,  15, 2,  8,  2,  2,  8,  8,  2,  2,  15, 15, 6,  8,  2,  8,  15, 15, 2,  8,  2, 2,
                                   2,  15, 15, 6,  6,  2,  6,  8,  15, 15, 2,  2,  15, 15, 15, 15, 15, 2,  2,  15};
constexpr uint8_t k_order[64] = {0,  13, 1,  2,  15, 14, 10, 16, 3,  23, 26, 6,  7,  21, 19, 29, 8,  4,  9,  20, 5,  31,
                                 22, 17, 18, 11, 12, 30, 24, 25, 28, 27, 32, 33, 34, 45, 46, 51, 49, 50, 48, 38, 39, 37,
                                 53, 52, 54, 36, 57, 58, 55, 41, 40, 42, 43, 59, 44, 56, 47, 35, 60, 63, 62, 61};
constexpr uint32_t bit(int x) { return 1u << x; }
constexpr uint32_t k_all = 0xFFFFFFFFu;
constexpr uint32_t k_pred[35] = {k_all, k_all, k_all, k_all, k_all,
                                 bit(1) | bit(2) | bit(8), bit(1) | bit(3) | bit(7), k_all, k_all,
                                 bit(2) | bit(8) | bit(16), bit(7) | bit(3) | bit(15), k_all,
                                 bit(8) | bit(14) | bit(16), bit(7) | bit(14) | bit(15), k_all, k_all, k_all, k_all,
                                 bit(14) | bit(15), bit(16) | bit(22) | bit(14), bit(17) | bit(24) | bit(14),
                                 bit(2) | bit(14) | bit(15) | bit(1), k_all,
                                 bit(1) | bit(3) | bit(14) | bit(16) | bit(22), k_all,
                                 bit(1) | bit(2) | bit(15) | bit(17) | bit(24), bit(1) | bit(3) | bit(22), k_all, k_all,
                                 k_all, bit(14) | bit(15) | bit(16) | bit(17), k_all, k_all,
                                 bit(1) | bit(2) | bit(3) | bit(27) | bit(4) | bit(24),
                                 bit(14) | bit(15) | bit(16) | bit(11) | bit(17) | bit(27)};
constexpr float k_w2x[4][4] = {{0.000000f, 0.000000f, 1.000000f, 0.000000f},
                               {0.107666f, 0.220459f, 0.451416f, 0.328125f},
                               {0.451416f, 0.220459f, 0.107666f, 0.671875f},
                               {1.000000f, 0.000000f, 0.000000f, 1.000000f}};
constexpr float k_w3x[8][4] = {{0.000000f, 0.000000f, 1.000000f, 0.000000f}, {0.019775f, 0.120850f, 0.738525f, 0.140625f},
                               {0.079102f, 0.202148f, 0.516602f, 0.281250f}, {0.177979f, 0.243896f, 0.334229f, 0.421875f},
                               {0.334229f, 0.243896f, 0.177979f, 0.578125f}, {0.516602f, 0.202148f, 0.079102f, 0.718750f},
                               {0.738525f, 0.120850f, 0.019775f, 0.859375f}, {1.000000f, 0.000000f, 0.000000f, 1.000000f}};
constexpr float k_w4x[16][4] = {
        {0.000000f, 0.000000f, 1.000000f, 0.000000f}, {0.003906f, 0.058594f, 0.878906f, 0.062500f},
        {0.019775f, 0.120850f, 0.738525f, 0.140625f}, {0.041260f, 0.161865f, 0.635010f, 0.203125f},
        {0.070557f, 0.195068f, 0.539307f, 0.265625f}, {0.107666f, 0.220459f, 0.451416f, 0.328125f},
        {0.165039f, 0.241211f, 0.352539f, 0.406250f}, {0.219727f, 0.249023f, 0.282227f, 0.468750f},
        {0.282227f, 0.249023f, 0.219727f, 0.531250f}, {0.352539f, 0.241211f, 0.165039f, 0.593750f},
        {0.451416f, 0.220459f, 0.107666f, 0.671875f}, {0.539307f, 0.195068f, 0.070557f, 0.734375f},
        {0.635010f, 0.161865f, 0.041260f, 0.796875f}, {0.738525f, 0.120850f, 0.019775f, 0.859375f},
        {0.878906f, 0.058594f, 0.003906f, 0.937500f}, {1.000000f, 0.000000f, 0.000000f, 1.000000f}};

// value of an endpoint of `bits` bits (p-bit already appended) replicated to 8 bits
inline uint32_t to8(uint32_t q, uint32_t bits)
{
    uint32_t v = q << (8 - bits);
    return v | (v >> bits);
}

// midpoint between the reconstruction levels of bin i and bin i+1 for a (bits-1)+p quantiser
inline float midpoint(uint32_t i, uint32_t p, uint32_t bins, uint32_t bits, bool has_p)
{
    if(i == bins - 1) { return 1.0f; }
    const uint32_t ql = has_p ? ((i << 1) | p) : i;
    const uint32_t qh = has_p ? ((std::min(bins - 1, i + 1) << 1) | p) : std::min(bins - 1, i + 1);
    const float lo = to8(ql, bits) / 255.0f;
    const float hi = to8(qh, bits) / 255.0f;
    return (lo + hi) / 2.0f;
}
}// namespace

uint64_t bc7_uber_map_reference(int max_sel, int ly, int hy)
{
    uint64_t map = 0;
    for(int sel = 0; sel <= max_sel; ++sel)
    {
        // (int)clampf(floorf((float)max_selector * ((float)sel - (float)ly) / ((float)hy - (float)ly) + .5f), 0, (float)max_selector)
        volatile float num = static_cast<float>(max_sel) * (static_cast<float>(sel) - static_cast<float>(ly));
        volatile float den = static_cast<float>(hy) - static_cast<float>(ly);
        volatile float q = num / den;
        volatile float r = q + .5f;
        float v = std::floor(r);
        v = v < 0.0f ? 0.0f : (v > static_cast<float>(max_sel) ? static_cast<float>(max_sel) : v);
        map |= static_cast<uint64_t>(static_cast<int>(v)) << (4 * sel);
    }
    return map;
}

void bc7_tables_build(Bc7Tables *t)
{
    std::memset(t, 0, sizeof(*t));
    std::memcpy(t->part2, k_part2, sizeof(k_part2));
    std::memcpy(t->anchor2, k_anchor2, sizeof(k_anchor2));
    std::memcpy(t->order, k_order, sizeof(k_order));
    std::memcpy(t->pred, k_pred, sizeof(k_pred));
    std::memcpy(t->w2x, k_w2x, sizeof(k_w2x));
    std::memcpy(t->w3x, k_w3x, sizeof(k_w3x));
    std::memcpy(t->w4x, k_w4x, sizeof(k_w4x));
    for(uint32_t p = 0; p < 64; ++p)
    {
        uint32_t n = 0;
        for(uint32_t sub = 0; sub < 2; ++sub)
        {
            for(uint32_t i = 0; i < 16; ++i)
            {
                if(((k_part2[p] >> i) & 1u) == sub)
                {
                    t->est_perm[p] |= static_cast<uint64_t>(i) << (4 * n);
                    t->est_idx[p][n++] = static_cast<uint8_t>(i);
                }
            }
            if(sub == 0) { t->est_n0[p] = static_cast<uint8_t>(n); }
        }
    }

    for(uint32_t v = 0; v < 256; ++v) { t->unit8[v] = static_cast<float>(v) / 255.0f; }
    for(uint32_t p = 0; p < 2; ++p)
    {
        for(uint32_t i = 0; i < 32; ++i) { t->mid7[i][p] = midpoint(i, p, 32, 6, true); }
        for(uint32_t i = 0; i < 64; ++i) { t->mid1[i][p] = midpoint(i, p, 64, 7, true); }
    }
    for(uint32_t i = 0; i < 128; ++i) { t->mid5[i] = midpoint(i, 0, 128, 7, false); }

    // mode 1: selector 2 of the 3-bit weights {0,9,18,...}: k = (lo*(64-18) + hi*18 + 32) >> 6, first strict minimum wins
    for(int c = 0; c < 256; ++c)
    {
        for(uint32_t p = 0; p < 2; ++p)
        {
            uint32_t best_err = 0xFFFF, best_l = 0, best_h = 0;
            for(uint32_t l = 0; l < 64; ++l)
            {
                const uint32_t low = to8((l << 1) | p, 7);
                for(uint32_t h = 0; h < 64; ++h)
                {
                    const uint32_t high = to8((h << 1) | p, 7);
                    const int k = static_cast<int>((low * 46 + high * 18 + 32) >> 6);
                    const uint32_t err = static_cast<uint32_t>((k - c) * (k - c));
                    if(err < best_err) { best_err = err, best_l = l, best_h = h; }
                }
            }
            t->opt1[c][p] = best_err | (best_l << 16) | (best_h << 24);
        }
        // mode 7: selector 1 of the 2-bit weights {0,21,43,64}
        for(uint32_t hp = 0; hp < 2; ++hp)
        {
            for(uint32_t lp = 0; lp < 2; ++lp)
            {
                uint32_t best_err = 0xFFFF, best_l = 0, best_h = 0;
                for(uint32_t l = 0; l < 32; ++l)
                {
                    const uint32_t low = to8((l << 1) | lp, 6);
                    for(uint32_t h = 0; h < 32; ++h)
                    {
                        const uint32_t high = to8((h << 1) | hp, 6);
                        const int k = static_cast<int>((low * 43 + high * 21 + 32) >> 6);
                        const uint32_t err = static_cast<uint32_t>((k - c) * (k - c));
                        if(err < best_err) { best_err = err, best_l = l, best_h = h; }
                    }
                }
                t->opt7[c][hp * 2 + lp] = best_err | (best_l << 16) | (best_h << 24);
            }
        }
    }
}

// g_mode6_reduced_quant, bc7enc.cpp:188-211: for every 11-bit endpoint position the nearest of 64 evenly spread 7-bit
// levels, with the p-bit appended before the comparison; float arithmetic, first strict minimum wins.
void bc7_m6_reduced_build(uint8_t *out)
{
    for(uint32_t p = 0; p < 2; ++p)
    {
        for(uint32_t i = 0; i < 2048; ++i)
        {
            const float target = static_cast<float>(i) / 2047.0f;
            float best = 1e+9f;
            int level = 0;
            for(int j = 0; j < 64; ++j)
            {
                const int q = (j * 127 + 31) / 63;
                const float d = std::fabs(static_cast<float>((q << 1) + static_cast<int>(p)) / 255.0f - target);
                if(d < best) { best = d, level = q; }
            }
            out[i * 2 + p] = static_cast<uint8_t>(level);
        }
    }
}

}// namespace vkt
