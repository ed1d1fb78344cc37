// bc7_tables.h -- constant data of the BC7 encoder, shared by the CUDA kernels and the host-side table builder.
//
// One POD struct that the host fills once per context (bc7_tables_build, bc7_tables.cpp) and uploads to HBM; every
// CTA copies it into shared memory so that lane-divergent lookups (optimal single-colour endpoints, quantiser
// midpoints, least-squares weights, per-lane partition masks) are bank-parallel instead of serialising in the
// constant cache.
//
// Reference: the tables bc7enc_compress_block_init() builds (/root/reference/extern/bc7enc_rdo/bc7enc.cpp:124-285) and
// the static tables at bc7enc.cpp:48-105, 1714-1751, 1765-1775.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace vkt
{

struct Bc7Tables
{
    // optimal single-colour endpoints, packed err | lo << 16 | hi << 24
    uint32_t opt1[256][2];// [colour][pbit]              bc7enc.cpp:213-240
    float mid1[64][2];    // mode-1 6-bit+p quantiser midpoints   bc7enc.cpp:150-169
    float mid7[32][2];    // mode-7 5-bit+p                       bc7enc.cpp:129-148
    float mid5[128];      // mode-5 7-bit                         bc7enc.cpp:171-186
    // least-squares tuples {w*w, (1-w)*w, (1-w)*(1-w), w} -- the decimal literals of bc7enc.cpp:52-57 (data!)
    float w2x[4][4];
    float w3x[8][4];
    float w4x[16][4];
    uint32_t pred[36];   // filterbank predictors bc7enc.cpp:1714-1751 (+1 pad)
    uint16_t part2[64];  // two-subset partition masks, bit i = subset of texel i (bc7enc.cpp:60-70 packed)
    uint8_t anchor2[64]; // bc7enc.cpp:94
    uint8_t order[64];   // partition scan order bc7enc.cpp:1765-1775
    // estimator work lists: texel indices of subset 0 (ascending) followed by those of subset 1, and |subset 0|
    uint64_t est_perm[64];// the same list, one nibble per texel
    uint8_t est_idx[64][16];
    uint8_t est_n0[64];
    float unit8[256];// v / 255.0f, the correctly rounded quotient find_optimal_solution divides out per component (bc7enc.cpp:1040-1042)
    // LAST: only the alpha kernels (mode 7) read it; the opaque kernels copy the struct up to here into shared memory
    uint32_t opt7[256][4];// [colour][hi_p * 2 + lo_p]   bc7enc.cpp:242-282
};

static_assert(sizeof(Bc7Tables) % 16 == 0, "copied to shared memory as uint4");

// Filled on the host (product code, C++): vierkant_b200/csrc/bc7_tables.cpp
void bc7_tables_build(Bc7Tables *t);
// g_mode6_reduced_quant[2048][2] (bc7enc.cpp:188-211; only read with quant_mode6_endpoints): out[value * 2 + p], 4096 bytes.
// Kept outside Bc7Tables: it is not copied to shared memory.
constexpr size_t kBc7M6ReducedBytes = 2048 * 2;
void bc7_m6_reduced_build(uint8_t *out);
// The reference's float expression for the uber-level selector rescaling (bc7enc.cpp:1399), for checking the
// integer-generated table in bc7_core.cuh: nibble s of the result is the rescaled selector.
uint64_t bc7_uber_map_reference(int max_sel, int ly, int hy);

}// namespace vkt
