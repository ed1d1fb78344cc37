// texture_block_compression_cuda.cpp -- drop-in translation unit for vierkant: vierkant::bcn::compress() on B200.
//
// Replaces /root/reference/src/texture_block_compression.cpp when vierkant is configured with -DVIERKANT_BCN_CUDA=ON
// (integration/vierkant_bcn_cuda.cmake).  Same signature, same compress_info_t / compress_result_t contract
// (include/vierkant/texture_block_compression.hpp:15-52), identical block bytes and mip chains; the work itself is done
// by libvierkant_bcn_cuda through its C ABI (include/vierkant_bcn_cuda.h).  There is no CPU fallback: if no CUDA device
// is usable, compress() throws std::runtime_error.
//
// Contract notes:
//   * level count / sizes: vkt_bcn_cuda_compress_plan == texture_block_compression.cpp:80-86,141-146
//   * the image must be a crocore::Image_<uint8_t> with >= 3 components (the reference asserts, :68-69); 3-component
//     images get alpha = 255 (get_block, :39-60)
//   * delegate_fn is accepted and ignored: the GPU does the fan-out the reference delegates to a thread pool (:107-139)
//   * duration is the whole call in milliseconds, never 0: the reference's own test asserts duration > 0 ms
//     (tests/TestCompressionBC7.cpp:48) and a GPU call can finish in less than one
#include <algorithm>
#include <array>
#include <chrono>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include <vierkant/texture_block_compression.hpp>

#include "texture_block_compression_batch.hpp"
#include "vierkant_bcn_cuda.h"

namespace vierkant::bcn
{

namespace
{
//! process-wide encoder context over every visible CUDA device (the role of the reference's init_helper_t, :29-37)
struct cuda_context_t
{
    vkt_bcn_ctx *ctx = nullptr;
    cuda_context_t()
    {
        // VIERKANT_BCN_CUDA_DEVICES="0,2": restrict the encoder to these CUDA ordinals (default: every visible device)
        std::vector<int> devices;
        if(const char *e = std::getenv("VIERKANT_BCN_CUDA_DEVICES"))
        {
            for(const char *q = e; *q;)
            {
                char *end = nullptr;
                const long v = std::strtol(q, &end, 10);
                if(end == q) { break; }
                devices.push_back(int(v));
                q = (*end == ',') ? end + 1 : end;
            }
        }
        if(vkt_bcn_cuda_create(&ctx, devices.empty() ? nullptr : devices.data(), int(devices.size())) != VKT_BCN_OK)
        {
            throw std::runtime_error(std::string("vierkant::bcn (CUDA): ") + vkt_bcn_cuda_last_error(nullptr));
        }
    }
    ~cuda_context_t() { vkt_bcn_cuda_destroy(ctx); }
    cuda_context_t(const cuda_context_t &) = delete;
    cuda_context_t &operator=(const cuda_context_t &) = delete;
};
cuda_context_t &shared_context()
{
    static cuda_context_t context;// thread-safe magic static, as in the reference (:66)
    return context;
}
}// namespace

compress_result_t compress(const compress_info_t &compress_info)
{
    cuda_context_t &context = shared_context();

    auto image = std::dynamic_pointer_cast<const crocore::Image_<uint8_t>>(compress_info.image);
    if(!image || image->num_components() < 3)
    {
        throw std::invalid_argument("vierkant::bcn::compress: expected an 8-bit image with 3 or 4 components");
    }
    auto start_time = std::chrono::steady_clock::now();

    vkt_bcn_plan plan = {};
    if(vkt_bcn_cuda_compress_plan(image->width(), image->height(), compress_info.generate_mipmaps ? 1 : 0, &plan) != VKT_BCN_OK)
    {
        throw std::invalid_argument("vierkant::bcn::compress: empty image");
    }
    compress_result_t ret = {};
    ret.mode = compress_info.mode;
    ret.base_width = plan.base_width;
    ret.base_height = plan.base_height;
    ret.levels.resize(plan.num_levels);
    static_assert(sizeof(block_t) == 16, "block_t is the 16-byte BCn block");

    // The levels' storage is allocated by this callback, which the library calls once the whole chain is queued on the GPU:
    // std::vector::resize maps and zero-fills the pages while the upload and the kernels already run (the reference sizes
    // the vectors up front, src/texture_block_compression.cpp:88-96 -- in front of the first upload that is pure latency).
    auto alloc_level = [](void *user, uint32_t level, size_t bytes) -> void * {
        auto &levels = *static_cast<std::vector<std::vector<block_t>> *>(user);
        try
        {
            levels[level].resize(bytes / sizeof(block_t));
        } catch(...)// (no exception may cross the C ABI, least of all on the library's helper thread: null = out of memory)
        {
            return nullptr;
        }
        return levels[level].data();
    };
    // bc7enc_compress_block_params_init() defaults, as the reference uses them (:73-74); NULL selects them
    const uint32_t mode = compress_info.mode == BC7 ? VKT_BCN_MODE_BC7 : VKT_BCN_MODE_BC5;
    const int rc = vkt_bcn_cuda_compress_alloc(context.ctx, mode, static_cast<const uint8_t *>(image->data()), image->width(), image->height(),
                                               image->num_components(), compress_info.generate_mipmaps ? 1 : 0, nullptr, +alloc_level, &ret.levels);
    if(rc != VKT_BCN_OK)
    {
        throw std::runtime_error(std::string("vierkant::bcn::compress (CUDA): ") + vkt_bcn_cuda_last_error(context.ctx));
    }
    auto elapsed = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - start_time);
    ret.duration = std::max(elapsed, std::chrono::milliseconds(1));
    return ret;
}

std::vector<compress_result_t> compress(std::span<const compress_info_t> compress_infos)
{
    cuda_context_t &context = shared_context();
    auto start_time = std::chrono::steady_clock::now();
    const size_t n = compress_infos.size();
    std::vector<compress_result_t> results(n);
    std::vector<std::array<void *, 16>> level_ptrs(n);
    std::vector<vkt_bcn_source> sources(n);
    bool mipmaps = n ? compress_infos[0].generate_mipmaps : false;
    for(size_t i = 0; i < n; ++i)
    {
        const compress_info_t &info = compress_infos[i];
        auto image = std::dynamic_pointer_cast<const crocore::Image_<uint8_t>>(info.image);
        if(!image || image->num_components() < 3)
        {
            throw std::invalid_argument("vierkant::bcn::compress: expected 8-bit images with 3 or 4 components");
        }
        if(info.generate_mipmaps != mipmaps) { throw std::invalid_argument("vierkant::bcn::compress: one mipmap setting per batch"); }
        vkt_bcn_plan plan = {};
        if(vkt_bcn_cuda_compress_plan(image->width(), image->height(), mipmaps ? 1 : 0, &plan) != VKT_BCN_OK)
        {
            throw std::invalid_argument("vierkant::bcn::compress: empty image");
        }
        compress_result_t &ret = results[i];
        ret.mode = info.mode;
        ret.base_width = plan.base_width;
        ret.base_height = plan.base_height;
        ret.levels.resize(plan.num_levels);
        level_ptrs[i].fill(nullptr);
        for(uint32_t l = 0; l < plan.num_levels; ++l)
        {
            ret.levels[l].resize(plan.level_num_blocks[l]);
            level_ptrs[i][l] = ret.levels[l].data();
        }
        sources[i] = {static_cast<const uint8_t *>(image->data()), image->width(), image->height(), image->num_components(),
                      info.mode == BC7 ? uint32_t(VKT_BCN_MODE_BC7) : uint32_t(VKT_BCN_MODE_BC5), level_ptrs[i].data()};
    }
    if(vkt_bcn_cuda_compress_batch(context.ctx, sources.data(), uint32_t(n), mipmaps ? 1 : 0, nullptr) != VKT_BCN_OK)
    {
        throw std::runtime_error(std::string("vierkant::bcn::compress (CUDA, batch): ") + vkt_bcn_cuda_last_error(context.ctx));
    }
    // model::compress_textures SUMS the results' durations for its "avg. Mpx/s" log (src/model/model_loading.cpp:119,135-138).
    // The textures of a batch overlap on the device, so each result carries its pixel share of the batch's wall time -- the
    // sum is the batch's time (never 0 per result: the reference's test asserts duration > 0 ms, tests/TestCompressionBC7.cpp:48).
    const double elapsed_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start_time).count();
    double total_px = 0.0;
    for(const auto &r: results) { total_px += double(r.base_width) * r.base_height; }
    for(auto &r: results)
    {
        const double share = total_px > 0.0 ? double(r.base_width) * r.base_height / total_px : 0.0;
        r.duration = std::max(std::chrono::milliseconds(static_cast<int64_t>(elapsed_ms * share + 0.5)), std::chrono::milliseconds(1));
    }
    return results;
}

}// namespace vierkant::bcn
