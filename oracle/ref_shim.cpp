// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// Thin C-ABI around the UNMODIFIED reference sources, compiled in place from /root/reference by
// oracle/Makefile into oracle/_ref/libvkt_ref.so.  It exists so that tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference leg can run the reference's own hot path:
//   vierkant::bcn::compress()          src/texture_block_compression.cpp:64-154
//   bc7enc_compress_block()            extern/bc7enc_rdo/bc7enc.cpp:2402-2438
//   crocore::Image_<uint8_t>::resize   extern/crocore/src/Image.cpp:239-247  (stbir_resize_uint8)
//   rgbcx::encode_bc5                  extern/bc7enc_rdo/rgbcx.cpp:2913
//   bc7decomp::unpack_bc7              extern/bc7enc_rdo/bc7decomp.cpp:594
// No reference source is copied into this repository; this file only calls the reference's public symbols.
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "bc7decomp.h"
#include "bc7enc.h"
#include "rgbcx.h"
#include <crocore/Image.hpp>
#include <crocore/ThreadPoolClassic.hpp>
#include <vierkant/texture_block_compression.hpp>

extern "C" {

// Field-by-field mirror of bc7enc_compress_block_params (bc7enc.h:14-75); never memcpy'd across the ABI.
struct ref_bc7_params
{
    uint32_t mode_mask;
    uint32_t max_partitions;
    uint32_t weights[4];
    uint32_t uber_level;
    uint32_t perceptual;
    uint32_t try_least_squares;
    uint32_t mode17_partition_estimation_filterbank;
    uint32_t force_alpha;
    uint32_t force_selectors;
    uint8_t selectors[16];
    uint32_t quant_mode6_endpoints;
    uint32_t bias_mode1_pbits;
    float pbit1_weight;
    float mode1_error_weight;
    float mode5_error_weight;
    float mode6_error_weight;
    float mode7_error_weight;
    float low_frequency_partition_weight;
};

static void to_native(const ref_bc7_params *s, bc7enc_compress_block_params *d)
{
    d->clear();
    bc7enc_compress_block_params_init(d);
    if(!s) { return; }
    d->m_mode_mask = s->mode_mask;
    d->m_max_partitions = s->max_partitions;
    for(int i = 0; i < 4; ++i) { d->m_weights[i] = s->weights[i]; }
    d->m_uber_level = s->uber_level;
    d->m_perceptual = s->perceptual != 0;
    d->m_try_least_squares = s->try_least_squares != 0;
    d->m_mode17_partition_estimation_filterbank = s->mode17_partition_estimation_filterbank != 0;
    d->m_force_alpha = s->force_alpha != 0;
    d->m_force_selectors = s->force_selectors != 0;
    memcpy(d->m_selectors, s->selectors, 16);
    d->m_quant_mode6_endpoints = s->quant_mode6_endpoints != 0;
    d->m_bias_mode1_pbits = s->bias_mode1_pbits != 0;
    d->m_pbit1_weight = s->pbit1_weight;
    d->m_mode1_error_weight = s->mode1_error_weight;
    d->m_mode5_error_weight = s->mode5_error_weight;
    d->m_mode6_error_weight = s->mode6_error_weight;
    d->m_mode7_error_weight = s->mode7_error_weight;
    d->m_low_frequency_partition_weight = s->low_frequency_partition_weight;
}

void ref_bc7_params_init(ref_bc7_params *p)
{
    bc7enc_compress_block_params n;
    n.clear();
    bc7enc_compress_block_params_init(&n);
    memset(p, 0, sizeof(*p));
    p->mode_mask = n.m_mode_mask;
    p->max_partitions = n.m_max_partitions;
    for(int i = 0; i < 4; ++i) { p->weights[i] = n.m_weights[i]; }
    p->uber_level = n.m_uber_level;
    p->perceptual = n.m_perceptual;
    p->try_least_squares = n.m_try_least_squares;
    p->mode17_partition_estimation_filterbank = n.m_mode17_partition_estimation_filterbank;
    p->pbit1_weight = n.m_pbit1_weight;
    p->mode1_error_weight = n.m_mode1_error_weight;
    p->mode5_error_weight = n.m_mode5_error_weight;
    p->mode6_error_weight = n.m_mode6_error_weight;
    p->mode7_error_weight = n.m_mode7_error_weight;
    p->low_frequency_partition_weight = n.m_low_frequency_partition_weight;
}

static void ensure_init()
{
    static bool once = [] {
        rgbcx::init(rgbcx::bc1_approx_mode::cBC1Ideal);
        bc7enc_compress_block_init();
        return true;
    }();
    (void) once;
}

// "direct" loop: blocks are 64-byte RGBA8 4x4 tiles, already gathered.  threads<=1 -> inline.
void ref_bc7_encode_blocks(const uint8_t *px, uint64_t num_blocks, const ref_bc7_params *params, uint8_t *out,
                           int threads)
{
    ensure_init();
    bc7enc_compress_block_params p;
    to_native(params, &p);
    auto work = [&](uint64_t b0, uint64_t b1) {
        for(uint64_t b = b0; b < b1; ++b) { bc7enc_compress_block(out + 16 * b, px + 64 * b, &p); }
    };
    if(threads <= 1) { work(0, num_blocks); }
    else
    {
        std::vector<std::thread> pool;
        std::atomic<uint64_t> next{0};
        const uint64_t chunk = 256;
        for(int t = 0; t < threads; ++t)
        {
            pool.emplace_back([&] {
                for(;;)
                {
                    uint64_t b0 = next.fetch_add(chunk);
                    if(b0 >= num_blocks) { break; }
                    work(b0, std::min(num_blocks, b0 + chunk));
                }
            });
        }
        for(auto &t: pool) { t.join(); }
    }
}

void ref_bc5_encode_blocks(const uint8_t *px, uint64_t num_blocks, uint8_t *out)
{
    ensure_init();
    for(uint64_t b = 0; b < num_blocks; ++b) { rgbcx::encode_bc5(out + 16 * b, px + 64 * b, 0, 1, 4); }
}

void ref_bc7_unpack_blocks(const uint8_t *blocks, uint64_t num_blocks, uint8_t *px)
{
    for(uint64_t b = 0; b < num_blocks; ++b)
    {
        bc7decomp::unpack_bc7(blocks + 16 * b, reinterpret_cast<bc7decomp::color_rgba *>(px + 64 * b));
    }
}

// crocore::Image_<uint8_t>::resize == stbir_resize_uint8 with default filters
void ref_resize_u8(const uint8_t *in, uint32_t w, uint32_t h, uint32_t comps, uint8_t *out, uint32_t ow, uint32_t oh)
{
    auto img = crocore::Image_<uint8_t>::create(const_cast<uint8_t *>(in), w, h, comps, true);
    auto res = std::dynamic_pointer_cast<crocore::Image_<uint8_t>>(img->resize(ow, oh));
    memcpy(out, res->data(), size_t(ow) * oh * comps);
}

struct ref_result
{
    vierkant::bcn::compress_result_t r;
};

// vierkant::bcn::compress() itself.  threads==0 -> no delegate (inline), else ThreadPoolClassic(threads) delegate
// exactly as model::compress_textures does (src/model/model_loading.cpp:110-118).
ref_result *ref_compress(const uint8_t *img_data, uint32_t w, uint32_t h, uint32_t comps, uint32_t mode, int mipmaps,
                         int threads)
{
    auto img = crocore::Image_<uint8_t>::create(const_cast<uint8_t *>(img_data), w, h, comps, true);
    vierkant::bcn::compress_info_t info = {};
    info.mode = static_cast<vierkant::bcn::CompressionMode>(mode);
    info.image = img;
    info.generate_mipmaps = mipmaps != 0;
    auto *ret = new ref_result;
    if(threads > 0)
    {
        crocore::ThreadPoolClassic pool(static_cast<size_t>(threads));
        info.delegate_fn = [&pool](auto fn) { return pool.post(fn); };
        ret->r = vierkant::bcn::compress(info);
    }
    else { ret->r = vierkant::bcn::compress(info); }
    return ret;
}
uint32_t ref_result_num_levels(const ref_result *r) { return static_cast<uint32_t>(r->r.levels.size()); }
uint32_t ref_result_base_width(const ref_result *r) { return r->r.base_width; }
uint32_t ref_result_base_height(const ref_result *r) { return r->r.base_height; }
uint32_t ref_result_mode(const ref_result *r) { return r->r.mode; }
int64_t ref_result_duration_ms(const ref_result *r) { return r->r.duration.count(); }
uint64_t ref_result_level_blocks(const ref_result *r, uint32_t l) { return r->r.levels[l].size(); }
const void *ref_result_level_data(const ref_result *r, uint32_t l) { return r->r.levels[l].data(); }
void ref_result_free(ref_result *r) { delete r; }

unsigned ref_hardware_concurrency() { return std::thread::hardware_concurrency(); }
}
