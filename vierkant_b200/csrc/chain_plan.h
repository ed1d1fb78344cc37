// chain_plan.h -- who computes which rows of which level of one vierkant::bcn::compress() chain (host logic, no CUDA).
//
// Reference: the level loop of vierkant::bcn::compress (/root/reference/src/texture_block_compression.cpp:99-146) runs
// every level on one machine; SURVEY.md 8e shards it by independent block rows.  With G devices:
//   * levels [0, M) are "sliced": device g encodes block rows [rows*g/G, rows*(g+1)/G) of each and produces, by
//     resizing, exactly the pixel rows those need -- its own rows plus whatever its rows of the next level read
//     through the filter taps (a halo that grows by ~5 rows per level).  need[] is computed from the real tap ranges.
//   * slicing stops where the halo would outgrow the slice (level height < 128 rows per device, or fewer than 4 block
//     rows per device): the small levels [M, L) are finished by device 0 from level M-1, which the other devices hand
//     over through pinned host memory.
//   * images too small to slice at all (M would be 0) are done by device 0 alone.
// The same header is compiled into tests/host_emul so the plan is checked on machines without a GPU.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <utility>
#include <vector>

namespace vkt
{

struct ChainSplit
{
    uint32_t levels = 0;// L
    uint32_t sliced = 0;// M: levels [0, M) are split by block rows over `devices`
    uint32_t devices = 1;// devices that take part (1 if the image is too small to slice)
    bool tail() const { return sliced < levels; }
};

inline ChainSplit chain_split(const uint32_t *level_height, uint32_t num_levels, uint32_t G)
{
    ChainSplit s;
    s.levels = num_levels, s.sliced = num_levels, s.devices = G ? G : 1;
    if(s.devices > 1)
    {
        uint32_t m = 0;
        while(m < num_levels && level_height[m] / 4 >= s.devices * 4 && level_height[m] >= 128u * s.devices) { ++m; }
        if(m == 0) { s.devices = 1; }
        else { s.sliced = m; }
    }
    return s;
}

// first_in / last_in of one vertical axis: per output row the smallest / largest input row its taps read
struct RowTaps
{
    const int *first_in = nullptr, *last_in = nullptr;
};

struct DeviceRows
{
    std::vector<std::pair<uint32_t, uint32_t>> own; // per sliced level: block rows [first, second) this device encodes
    std::vector<std::pair<uint32_t, uint32_t>> need;// per sliced level: pixel rows [first, second) it has to produce
};

// taps[l] maps rows of level l-1 (l == 0: the source image) to rows of level l, for l < split.sliced
inline DeviceRows device_rows(const ChainSplit &split, const uint32_t *level_height, const RowTaps *taps, uint32_t g)
{
    DeviceRows r;
    const uint32_t M = split.sliced;
    r.own.resize(M), r.need.resize(M);
    for(uint32_t l = 0; l < M; ++l)
    {
        const uint32_t rows = level_height[l] / 4;
        r.own[l] = {uint32_t(uint64_t(rows) * g / split.devices), uint32_t(uint64_t(rows) * (g + 1) / split.devices)};
    }
    for(uint32_t l = M; l-- > 0;)
    {
        r.need[l] = {r.own[l].first * 4, r.own[l].second * 4};
        if(l + 1 < M)
        {
            uint32_t lo = r.need[l].first, hi = r.need[l].second;
            for(uint32_t y = r.need[l + 1].first; y < r.need[l + 1].second; ++y)
            {
                lo = std::min(lo, uint32_t(taps[l + 1].first_in[y])), hi = std::max(hi, uint32_t(taps[l + 1].last_in[y]) + 1u);
            }
            r.need[l] = {lo, std::min(hi, level_height[l])};
        }
    }
    return r;
}

// source rows [first, second) read by the resize of output rows [y0, y1) of level 0
inline std::pair<uint32_t, uint32_t> source_rows(const RowTaps &t0, uint32_t y0, uint32_t y1, uint32_t src_height)
{
    uint32_t lo = src_height, hi = 0;
    for(uint32_t y = y0; y < y1; ++y) { lo = std::min(lo, uint32_t(t0.first_in[y])), hi = std::max(hi, uint32_t(t0.last_in[y]) + 1u); }
    return {std::min(lo, src_height), std::min(hi, src_height)};
}

}// namespace vkt
