// dropin_test_shim.cpp -- TEST INFRASTRUCTURE: exposes the drop-in vierkant::bcn::compress() (the TU in this directory,
// compiled against the reference's own headers) through a small C interface so that tests/test_dropin_gpu.py can run
// the reference's tests/TestCompressionBC7.cpp cases against it and compare the bytes with the reference build.
#include <cstdint>
#include <cstring>

#include <crocore/Image.hpp>
#include <vierkant/texture_block_compression.hpp>

#include "texture_block_compression_batch.hpp"

extern "C" {

struct dropin_result
{
    vierkant::bcn::compress_result_t r;
    bool image_ok = false;
};

dropin_result *dropin_compress(const uint8_t *img_data, uint32_t w, uint32_t h, uint32_t comps, uint32_t mode, int mipmaps)
{
    auto img = crocore::Image_<uint8_t>::create(const_cast<uint8_t *>(img_data), w, h, comps, true);
    vierkant::bcn::compress_info_t info = {};
    info.mode = static_cast<vierkant::bcn::CompressionMode>(mode);
    info.image = img;
    info.generate_mipmaps = mipmaps != 0;
    auto *ret = new dropin_result;
    ret->r = vierkant::bcn::compress(info);
    ret->image_ok = static_cast<bool>(info.image);
    return ret;
}
// two images through the batch overload; returns result `which` (0 / 1)
dropin_result *dropin_compress_pair(const uint8_t *a, uint32_t aw, uint32_t ah, const uint8_t *b, uint32_t bw, uint32_t bh, uint32_t comps,
                                    uint32_t mode, int mipmaps, int which)
{
    vierkant::bcn::compress_info_t infos[2] = {};
    infos[0].image = crocore::Image_<uint8_t>::create(const_cast<uint8_t *>(a), aw, ah, comps, true);
    infos[1].image = crocore::Image_<uint8_t>::create(const_cast<uint8_t *>(b), bw, bh, comps, true);
    for(auto &i: infos) { i.mode = static_cast<vierkant::bcn::CompressionMode>(mode), i.generate_mipmaps = mipmaps != 0; }
    auto results = vierkant::bcn::compress(std::span<const vierkant::bcn::compress_info_t>(infos, 2));
    auto *ret = new dropin_result;
    ret->r = std::move(results[which ? 1 : 0]);
    ret->image_ok = true;
    return ret;
}
uint32_t dropin_result_num_levels(const dropin_result *r) { return static_cast<uint32_t>(r->r.levels.size()); }
uint32_t dropin_result_base_width(const dropin_result *r) { return r->r.base_width; }
uint32_t dropin_result_base_height(const dropin_result *r) { return r->r.base_height; }
uint32_t dropin_result_mode(const dropin_result *r) { return r->r.mode; }
int64_t dropin_result_duration_ms(const dropin_result *r) { return r->r.duration.count(); }
uint64_t dropin_result_level_blocks(const dropin_result *r, uint32_t l) { return r->r.levels[l].size(); }
const void *dropin_result_level_data(const dropin_result *r, uint32_t l) { return r->r.levels[l].data(); }
void dropin_result_free(dropin_result *r) { delete r; }
}
