/* oracle/compress_oracle.c -- TEST INFRASTRUCTURE ONLY (see bc7_oracle.h for the rules).
 *
 * CPU restatement of vierkant::bcn::compress (/root/reference/src/texture_block_compression.cpp:64-154): round the size
 * up to a multiple of 4 (:80-81), level count (:84-86), per level: resize from the PREVIOUS level's image (:101, level 0
 * included), cut into 4x4 tiles with get_block (:39-60; alpha := 255 for 3-component images), encode every tile
 * (:116-136), next size = round4(max(d / 2, 1)) (:141-146).
 * Parity status: PINNED against the unmodified reference (oracle/_ref, ref_compress).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "bc7_oracle.h"

static uint32_t round4(uint32_t v) { return (v + 3u) & ~3u; }

uint32_t port_compress_num_levels(uint32_t width, uint32_t height, int mipmaps)
{
    const uint32_t w = round4(width), h = round4(height);
    const int max_levels = (int) fmax(0.0, (double) (int) (log2((double) (w > h ? w : h)) - 2)) + 1; /* :84-85 */
    return mipmaps ? (uint32_t) max_levels : 1u;
}

/* level_blocks[l] must hold (w_l / 4) * (h_l / 4) * 16 bytes; mode 0 = BC5, 1 = BC7. Returns the level count. */
uint32_t port_compress(const uint8_t *img, uint32_t width, uint32_t height, uint32_t comps, uint32_t mode, int mipmaps,
                       const port_bc7_params *params, uint8_t *const *level_blocks, int threads)
{
    uint32_t w = round4(width), h = round4(height);
    const uint32_t levels = port_compress_num_levels(width, height, mipmaps);
    uint8_t *prev = (uint8_t *) malloc((size_t) width * height * comps);
    memcpy(prev, img, (size_t) width * height * comps);
    uint32_t pw = width, ph = height;
    for(uint32_t l = 0; l < levels; ++l)
    {
        uint8_t *cur = (uint8_t *) malloc((size_t) w * h * comps);
        port_resize_u8(prev, pw, ph, comps, cur, w, h);
        free(prev);
        prev = cur, pw = w, ph = h;
        const uint32_t bx = w / 4, by = h / 4;
        uint8_t *tiles = (uint8_t *) malloc((size_t) bx * by * 64);
        for(uint32_t y = 0; y < h; ++y)
        {
            for(uint32_t x = 0; x < w; ++x)
            {
                const uint8_t *s = cur + ((size_t) y * w + x) * comps;
                uint8_t *d = tiles + ((size_t) (y / 4) * bx + x / 4) * 64 + ((y & 3) * 4 + (x & 3)) * 4;
                d[0] = s[0], d[1] = s[1], d[2] = s[2], d[3] = comps == 4 ? s[3] : 255;
            }
        }
        if(mode == 1) { port_bc7_encode_blocks(tiles, (uint64_t) bx * by, params, level_blocks[l], threads); }
        else { port_bc5_encode_blocks(tiles, (uint64_t) bx * by, level_blocks[l]); }
        free(tiles);
        const uint32_t hw = w / 2 > 1 ? w / 2 : 1, hh = h / 2 > 1 ? h / 2 : 1;
        w = round4(hw), h = round4(hh);
    }
    free(prev);
    return levels;
}
