/* oracle/bc5_oracle.c -- TEST INFRASTRUCTURE ONLY (see bc7_oracle.h for the rules).
 *
 * CPU restatement of rgbcx::encode_bc5(pDst, pPixels, 0, 1, 4) (/root/reference/extern/bc7enc_rdo/rgbcx.cpp:2913-2919),
 * the BC5 branch of vierkant::bcn::compress (src/texture_block_compression.cpp:131): two BC4 blocks, red then green,
 * each by rgbcx::encode_bc4 (rgbcx.cpp:2608-2728): endpoints (max, min), 3-bit selectors from seven thresholds on
 * 14 * v + (4 - 14 * min) against delta * {13, 11, 9, 7, 5, 3, 1}, translated by {1, 7, 6, 5, 4, 3, 2, 0}.
 * Parity status: PINNED against the unmodified reference (oracle/_ref, ref_bc5_encode_blocks).
 */
#include <stdint.h>
#include <string.h>

#include "bc7_oracle.h"

static void bc4_channel(const uint8_t *px, int stride, uint8_t *dst)
{
    uint32_t mn = 255, mx = 0;
    for(int i = 0; i < 16; ++i)
    {
        const uint32_t v = px[i * stride];
        if(v < mn) { mn = v; }
        if(v > mx) { mx = v; }
    }
    dst[0] = (uint8_t) mx;
    dst[1] = (uint8_t) mn;
    if(mx == mn)
    {
        memset(dst + 2, 0, 6);
        return;
    }
    static const uint32_t tran[8] = {1, 7, 6, 5, 4, 3, 2, 0};
    const int delta = (int) (mx - mn);
    const int bias = 4 - (int) mn * 14;
    uint64_t bits = 0;
    for(int i = 0; i < 16; ++i)
    {
        const int v = px[i * stride] * 14 + bias;
        int n = 0;
        for(int t = 13; t >= 1; t -= 2) { n += (v >= delta * t); }
        bits |= (uint64_t) tran[n] << (3 * i);
    }
    for(int i = 0; i < 6; ++i) { dst[2 + i] = (uint8_t) (bits >> (8 * i)); }
}

void port_bc5_encode_blocks(const uint8_t *px, uint64_t num_blocks, uint8_t *out)
{
    for(uint64_t b = 0; b < num_blocks; ++b)
    {
        bc4_channel(px + 64 * b + 0, 4, out + 16 * b);
        bc4_channel(px + 64 * b + 1, 4, out + 16 * b + 8);
    }
}
