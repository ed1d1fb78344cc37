// bcn_cuda.cu -- libvierkant_bcn_cuda: CUDA kernels (sm_100a) + the C ABI declared in include/vierkant_bcn_cuda.h.
//
// Replaces the per-block CPU loop of vierkant::bcn::compress() (/root/reference/src/texture_block_compression.cpp:107-139)
// and, for the whole-chain entry point, its stbir resize calls (:101).  There is no CPU fallback in this file: every
// entry point either runs the kernels or returns an error code.
//
// Build (see __graft_entry__.build / vierkant_b200/build.py):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true
//        -Xcompiler -fPIC,-ffp-contract=off -shared -o libvierkant_bcn_cuda.so bcn_cuda.cu bc7_tables.cpp
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#define VKT_BCN_DEFINE_PARAMS_INIT
#include "bc5_core.cuh"
#include "bc7_core.cuh"
#include "bc7_params.h"

namespace vkt
{

// ------------------------------------------------------------------------------------------------ texel staging
// One lane == one block.  A warp reads 32 horizontally adjacent blocks: for RGBA8 each of the 4 texel rows is one
// fully coalesced 512-byte request of 128-bit loads.
template<int NT>
__device__ __forceinline__ void load_block_texels(const uint8_t *__restrict__ img, uint32_t comps, uint32_t stride, bool vec16,
                                                  uint32_t bx, uint32_t by, Texel *col)
{
    if(vec16)
    {
#pragma unroll
        for(int y = 0; y < 4; ++y)
        {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(img + size_t(by * 4 + y) * stride) + bx);
            col[(4 * y + 0) * NT].px = v.x, col[(4 * y + 1) * NT].px = v.y, col[(4 * y + 2) * NT].px = v.z, col[(4 * y + 3) * NT].px = v.w;
        }
    }
    else
    {
        for(int y = 0; y < 4; ++y)
        {
            const uint8_t *row = img + size_t(by * 4 + y) * stride + size_t(bx) * 4 * comps;
            for(int x = 0; x < 4; ++x)
            {
                const uint8_t *t = row + x * comps;
                const uint32_t a = (comps == 4) ? t[3] : 255u;// get_block: alpha := 255 for 3-component images
                col[(4 * y + x) * NT].px = pack4(t[0], t[1], t[2], a);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ BC7 kernels
#ifndef VKT_BC7_THREADS
#define VKT_BC7_THREADS 256
#endif
#ifndef VKT_BC7_CTAS
#define VKT_BC7_CTAS 3
#endif
#ifndef VKT_BC7_CTAS_ALPHA
#define VKT_BC7_CTAS_ALPHA 2
#endif
constexpr int kBc7Threads = VKT_BC7_THREADS;
// Opaque blocks: 3 CTAs/SM = 3 x (64 KB lane columns + tables) of shared memory at <= 80 registers per thread.
// Alpha blocks carry a fourth channel through every stage: 2 CTAs/SM at <= 128 registers (no spills) is faster.
constexpr int kBc7CtasPerSm = VKT_BC7_CTAS, kBc7CtasPerSmAlpha = VKT_BC7_CTAS_ALPHA;

// Stage 0: split the level's blocks into an opaque and an alpha work list (the dispatch of bc7enc_compress_block,
// bc7enc.cpp:2422-2437), so that every warp of the encode kernels holds blocks of one kind.  counts[0] = #opaque,
// counts[1] = #alpha.  List order is irrelevant to the output (block b always lands in out[b]); warp-aggregated
// atomics keep it nearly sorted, so the encode kernels' texel loads stay coalesced.
__global__ void __launch_bounds__(256) bc7_classify_kernel(const uint8_t *__restrict__ img, uint32_t blocks_x, uint32_t num_blocks,
                                                            uint32_t stride, int vec16, uint32_t *__restrict__ counts,
                                                            uint32_t *__restrict__ list_opaque, uint32_t *__restrict__ list_alpha)
{
    const uint32_t b = blockIdx.x * 256 + threadIdx.x;
    const bool valid = b < num_blocks;
    bool alpha = false;
    if(valid)
    {
        const uint32_t bx = b % blocks_x, by = b / blocks_x;
        uint32_t and_all = 0xFFFFFFFFu;
        if(vec16)
        {
#pragma unroll
            for(int y = 0; y < 4; ++y)
            {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(img + size_t(by * 4 + y) * stride) + bx);
                and_all &= v.x & v.y & v.z & v.w;
            }
        }
        else
        {
            for(int y = 0; y < 4; ++y)
            {
                const uint8_t *row = img + size_t(by * 4 + y) * stride + size_t(bx) * 16;
                for(int x = 0; x < 4; ++x) { and_all &= uint32_t(row[4 * x + 3]) << 24; }
            }
        }
        alpha = (and_all >> 24) != 255u;
    }
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t m_alpha = __ballot_sync(0xFFFFFFFFu, valid && alpha), m_opaque = __ballot_sync(0xFFFFFFFFu, valid && !alpha);
    uint32_t base_a = 0, base_o = 0;
    if(lane == 0)
    {
        if(m_opaque) { base_o = atomicAdd(&counts[0], __popc(m_opaque)); }
        if(m_alpha) { base_a = atomicAdd(&counts[1], __popc(m_alpha)); }
    }
    base_o = __shfl_sync(0xFFFFFFFFu, base_o, 0), base_a = __shfl_sync(0xFFFFFFFFu, base_a, 0);
    const uint32_t below = (1u << lane) - 1u;
    if(valid)
    {
        if(alpha) { list_alpha[base_a + __popc(m_alpha & below)] = b; }
        else { list_opaque[base_o + __popc(m_opaque & below)] = b; }
    }
}

// One lane == one block of the work list (list == nullptr: every block of the level, in order).
template<bool PERC, bool KEY28, bool ALPHA, int NT>
__global__ void __launch_bounds__(NT, ALPHA ? kBc7CtasPerSmAlpha : kBc7CtasPerSm)
        bc7_encode_kernel(const uint8_t *__restrict__ img, uint32_t blocks_x, uint32_t num_blocks, uint32_t comps, uint32_t stride, int vec16,
                          const Bc7KernelParams P, const Bc7Tables *__restrict__ g_tables, const uint32_t *__restrict__ list,
                          const uint32_t *__restrict__ count, uint4 *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const uint32_t n = count ? __ldg(count) : num_blocks;
    if(blockIdx.x * NT >= n) { return; }// the grid is sized for the whole level; lists are usually shorter
    Bc7Tables &s_tables = *reinterpret_cast<Bc7Tables *>(s_raw);
    Texel *s_lane = reinterpret_cast<Texel *>(s_raw + sizeof(Bc7Tables));// one 16-record column per lane (see Lane<>)
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(g_tables);
        uint4 *dst = reinterpret_cast<uint4 *>(&s_tables);
        for(uint32_t i = threadIdx.x; i < sizeof(Bc7Tables) / 16; i += NT) { dst[i] = __ldg(src + i); }
    }
    const uint32_t i = blockIdx.x * NT + threadIdx.x;
    const uint32_t ii = min(i, n - 1);// out-of-range lanes redo the last block (keeps warps converged), no store
    const uint32_t b = list ? __ldg(list + ii) : ii;
    Lane<NT> lane{s_lane + threadIdx.x};
    load_block_texels<NT>(img, comps, stride, vec16 != 0, b % blocks_x, b / blocks_x, lane.p);
    __syncthreads();
    uint32_t blk[4];
    encode_block<PERC, KEY28, ALPHA, NT>(s_tables, P, lane, blk);
    if(i < n) { out[b] = make_uint4(blk[0], blk[1], blk[2], blk[3]); }
}

constexpr size_t kBc7SmemBytes = sizeof(Bc7Tables) + size_t(kBc7Threads) * 16 * sizeof(Texel);

template<bool PERC, bool KEY28, bool ALPHA>
static cudaError_t bc7_kernel_attribute()
{
    return cudaFuncSetAttribute(bc7_encode_kernel<PERC, KEY28, ALPHA, kBc7Threads>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kBc7SmemBytes));
}
// > 48 KB of dynamic shared memory needs an explicit opt-in per kernel (and per device: called from context creation)
static cudaError_t bc7_kernel_attributes()
{
    cudaError_t e = bc7_kernel_attribute<true, true, false>();
    if(e == cudaSuccess) { e = bc7_kernel_attribute<true, true, true>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<true, false, false>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<true, false, true>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<false, true, false>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<false, true, true>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<false, false, false>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<false, false, true>(); }
    return e;
}

// ------------------------------------------------------------------------------------------------ context
}// namespace vkt
struct vkt_axis_cache;// resize_core.cuh
namespace vkt
{
struct DeviceSlot
{
    vkt_axis_cache *axis_cache = nullptr;
    int device = -1;
    cudaStream_t stream = nullptr;
    Bc7Tables *d_tables = nullptr;
    void *d_in = nullptr, *d_out = nullptr, *d_tmp = nullptr;
    size_t in_cap = 0, out_cap = 0, tmp_cap = 0;
    std::mutex mtx;
};

}// namespace vkt

struct vkt_bcn_ctx
{
    std::vector<vkt::DeviceSlot *> slots;
    std::string last_error;
    std::mutex err_mtx;
    vkt_bcn_stats stats{};
    std::mutex stats_mtx;
};

namespace vkt
{
static thread_local std::string t_create_error;

static int fail(vkt_bcn_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if(ctx)
    {
        std::lock_guard<std::mutex> g(ctx->err_mtx);
        ctx->last_error = buf;
    }
    else { t_create_error = buf; }
    return code;
}

#define VKT_CUDA(ctx, call)                                                                                                    \
    do {                                                                                                                       \
        cudaError_t e_ = (call);                                                                                               \
        if(e_ != cudaSuccess)                                                                                                  \
        {                                                                                                                      \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? VKT_BCN_ERR_OOM : VKT_BCN_ERR_CUDA, "%s failed: %s", #call,   \
                        cudaGetErrorString(e_));                                                                               \
        }                                                                                                                      \
    } while(0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int ensure(vkt_bcn_ctx *ctx, void **ptr, size_t *cap, size_t need)
{
    if(*cap >= need) { return VKT_BCN_OK; }
    if(*ptr) { VKT_CUDA(ctx, cudaFree(*ptr)); }
    *ptr = nullptr, *cap = 0;
    VKT_CUDA(ctx, cudaMalloc(ptr, need));
    *cap = need;
    return VKT_BCN_OK;
}

static void count(vkt_bcn_ctx *ctx, uint64_t launches, uint64_t h2d, uint64_t d2h)
{
    std::lock_guard<std::mutex> g(ctx->stats_mtx);
    ctx->stats.kernel_launches += launches;
    ctx->stats.h2d_bytes += h2d;
    ctx->stats.d2h_bytes += d2h;
}

static int check_image(vkt_bcn_ctx *ctx, const void *px, uint32_t w, uint32_t h, uint32_t comps, uint32_t *stride, const void *out)
{
    if(!ctx) { return VKT_BCN_ERR_INVALID; }
    if(!px || !out) { return fail(ctx, VKT_BCN_ERR_INVALID, "null buffer"); }
    if(w == 0 || h == 0 || (w & 3) || (h & 3)) { return fail(ctx, VKT_BCN_ERR_INVALID, "width/height must be non-zero multiples of 4 (got %ux%u)", w, h); }
    if(comps != 3 && comps != 4) { return fail(ctx, VKT_BCN_ERR_INVALID, "comps must be 3 or 4 (got %u)", comps); }
    if(*stride == 0) { *stride = w * comps; }
    if(*stride < w * comps) { return fail(ctx, VKT_BCN_ERR_INVALID, "row stride %u smaller than a row", *stride); }
    return VKT_BCN_OK;
}

// launch the BC7 kernel on device-resident data (current device must be the slot's)
static int launch_bc7(vkt_bcn_ctx *ctx, DeviceSlot *s, const void *d_px, uint32_t w, uint32_t h, uint32_t comps, uint32_t stride,
                      const vkt_bc7_params *params, void *d_out, cudaStream_t stream)
{
    vkt_bc7_params def;
    if(!params)
    {
        vkt_bc7_params_init_inline(&def);
        params = &def;
    }
    Bc7KernelParams kp;
    const int rc = bc7_prepare_params(params, &kp);
    if(rc == VKT_BCN_ERR_UNSUPPORTED)
    {
        return fail(ctx, rc, "unsupported bc7 parameters (force_selectors / quant_mode6_endpoints / low_frequency_partition_weight != 1)");
    }
    if(rc) { return fail(ctx, rc, "invalid bc7 parameters (mode_mask must enable mode 6 or 1 and one of 5/6/7; uber_level <= 4)"); }
    const uint32_t bx = w / 4, nblocks = bx * (h / 4);
    const int vec16 = (comps == 4) && ((stride & 15u) == 0) && ((reinterpret_cast<uintptr_t>(d_px) & 15u) == 0);
    const uint32_t grid = (nblocks + kBc7Threads - 1) / kBc7Threads;
    const uint8_t *px = static_cast<const uint8_t *>(d_px);
    uint4 *outp = static_cast<uint4 *>(d_out);
    auto encode = [&](bool alpha, const uint32_t *list, const uint32_t *cnt) {
        auto go = [&](auto kernel) {
            kernel<<<grid, kBc7Threads, kBc7SmemBytes, stream>>>(px, bx, nblocks, comps, stride, vec16, kp, s->d_tables, list, cnt, outp);
        };
        const int sel = (params->perceptual ? 4 : 0) | (kp.key28 ? 2 : 0) | (alpha ? 1 : 0);
        switch(sel)
        {
            case 7: go(bc7_encode_kernel<true, true, true, kBc7Threads>); break;
            case 6: go(bc7_encode_kernel<true, true, false, kBc7Threads>); break;
            case 5: go(bc7_encode_kernel<true, false, true, kBc7Threads>); break;
            case 4: go(bc7_encode_kernel<true, false, false, kBc7Threads>); break;
            case 3: go(bc7_encode_kernel<false, true, true, kBc7Threads>); break;
            case 2: go(bc7_encode_kernel<false, true, false, kBc7Threads>); break;
            case 1: go(bc7_encode_kernel<false, false, true, kBc7Threads>); break;
            default: go(bc7_encode_kernel<false, false, false, kBc7Threads>); break;
        }
    };
    if(kp.force_alpha)
    {
        encode(true, nullptr, nullptr);
        count(ctx, 1, 0, 0);
    }
    else if(comps == 3)
    {
        encode(false, nullptr, nullptr);// get_block injects alpha = 255: every block is opaque
        count(ctx, 1, 0, 0);
    }
    else
    {
        // stream-ordered scratch: [counts: 2 x u32, padded to 256 B][opaque list][alpha list]
        uint8_t *scratch = nullptr;
        const size_t list_bytes = align_up(size_t(nblocks) * sizeof(uint32_t), 256);
        VKT_CUDA(ctx, cudaMallocAsync(reinterpret_cast<void **>(&scratch), 256 + 2 * list_bytes, stream));
        uint32_t *counts = reinterpret_cast<uint32_t *>(scratch);
        uint32_t *list_o = reinterpret_cast<uint32_t *>(scratch + 256), *list_a = reinterpret_cast<uint32_t *>(scratch + 256 + list_bytes);
        VKT_CUDA(ctx, cudaMemsetAsync(counts, 0, 8, stream));
        bc7_classify_kernel<<<(nblocks + 255) / 256, 256, 0, stream>>>(px, bx, nblocks, stride, vec16, counts, list_o, list_a);
        encode(false, list_o, counts);
        encode(true, list_a, counts + 1);
        const cudaError_t e = cudaGetLastError();
        VKT_CUDA(ctx, cudaFreeAsync(scratch, stream));
        VKT_CUDA(ctx, e);
        count(ctx, 3, 0, 0);
    }
    VKT_CUDA(ctx, cudaGetLastError());
    return VKT_BCN_OK;
}

static int launch_bc5(vkt_bcn_ctx *ctx, DeviceSlot *s, const void *d_px, uint32_t w, uint32_t h, uint32_t comps, uint32_t stride,
                      void *d_out, cudaStream_t stream)
{
    const uint32_t bx = w / 4, nblocks = bx * (h / 4);
    const uint32_t grid = (nblocks + 255) / 256;
    bc5_encode_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint8_t *>(d_px), bx, nblocks, comps, stride, static_cast<uint4 *>(d_out));
    VKT_CUDA(ctx, cudaGetLastError());
    count(ctx, 1, 0, 0);
    return VKT_BCN_OK;
}

// Encode block rows [row0, row1) of one host image on one slot: async H2D, kernel, async D2H on the slot's stream.
// The caller synchronises the stream.  Pinned host memory makes both copies truly asynchronous.
static int encode_rows_async(vkt_bcn_ctx *ctx, DeviceSlot *s, uint32_t mode, const vkt_bcn_image &img, uint32_t row0, uint32_t row1,
                             const vkt_bc7_params *params, size_t in_off, size_t out_off)
{
    const uint32_t stride = img.row_stride_bytes ? img.row_stride_bytes : img.width * img.comps;
    const uint32_t rows = row1 - row0;
    if(rows == 0) { return VKT_BCN_OK; }
    const size_t row_bytes = size_t(img.width) * img.comps;
    const size_t in_bytes = size_t(rows) * 4 * row_bytes;
    const size_t out_bytes = size_t(rows) * (img.width / 4) * 16;
    uint8_t *d_in = static_cast<uint8_t *>(s->d_in) + in_off;
    uint8_t *d_out = static_cast<uint8_t *>(s->d_out) + out_off;
    const uint8_t *h_in = img.pixels + size_t(row0) * 4 * stride;
    // device copy is tightly packed
    VKT_CUDA(ctx, cudaMemcpy2DAsync(d_in, row_bytes, h_in, stride, row_bytes, size_t(rows) * 4, cudaMemcpyHostToDevice, s->stream));
    int rc;
    if(mode == VKT_BCN_MODE_BC7)
    {
        rc = launch_bc7(ctx, s, d_in, img.width, rows * 4, img.comps, uint32_t(row_bytes), params, d_out, s->stream);
    }
    else { rc = launch_bc5(ctx, s, d_in, img.width, rows * 4, img.comps, uint32_t(row_bytes), d_out, s->stream); }
    if(rc) { return rc; }
    VKT_CUDA(ctx, cudaMemcpyAsync(static_cast<uint8_t *>(img.out_blocks) + size_t(row0) * (img.width / 4) * 16, d_out, out_bytes,
                                  cudaMemcpyDeviceToHost, s->stream));
    count(ctx, 0, in_bytes, out_bytes);
    return VKT_BCN_OK;
}

}// namespace vkt

#include "resize_core.cuh"

using namespace vkt;

// ================================================================================================ C ABI
extern "C" {

int vkt_bcn_cuda_device_count(void)
{
    int n = 0;
    if(cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int vkt_bcn_cuda_create(vkt_bcn_ctx **out_ctx, const int *devices, int num_devices)
{
    if(!out_ctx) { return VKT_BCN_ERR_INVALID; }
    *out_ctx = nullptr;
    const int visible = vkt_bcn_cuda_device_count();
    if(visible <= 0) { return fail(nullptr, VKT_BCN_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)"); }
    if(num_devices <= 0)
    {
        num_devices = visible;
        devices = nullptr;
    }
    auto *ctx = new(std::nothrow) vkt_bcn_ctx;
    if(!ctx) { return VKT_BCN_ERR_OOM; }
    Bc7Tables host_tables;
    bc7_tables_build(&host_tables);
    for(int i = 0; i < num_devices; ++i)
    {
        const int dev = devices ? devices[i] : i;
        if(dev < 0 || dev >= visible)
        {
            vkt_bcn_cuda_destroy(ctx);
            return fail(nullptr, VKT_BCN_ERR_INVALID, "device ordinal %d out of range (%d visible)", dev, visible);
        }
        auto *s = new DeviceSlot;
        s->device = dev;
        ctx->slots.push_back(s);
        cudaError_t e = cudaSetDevice(dev);
        if(e == cudaSuccess) { e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking); }
        if(e == cudaSuccess) { e = cudaMalloc(reinterpret_cast<void **>(&s->d_tables), sizeof(Bc7Tables)); }
        if(e == cudaSuccess) { e = cudaMemcpy(s->d_tables, &host_tables, sizeof(Bc7Tables), cudaMemcpyHostToDevice); }
        if(e == cudaSuccess) { e = bc7_kernel_attributes(); }
        if(e == cudaSuccess)
        {
            // the per-launch work lists come from the stream-ordered allocator: keep freed blocks cached in the pool
            // instead of returning them to the driver at every synchronisation
            cudaMemPool_t pool = nullptr;
            e = cudaDeviceGetDefaultMemPool(&pool, dev);
            uint64_t keep = ~0ull;
            if(e == cudaSuccess) { e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep); }
        }
        if(e != cudaSuccess)
        {
            fail(nullptr, VKT_BCN_ERR_CUDA, "device %d initialisation failed: %s", dev, cudaGetErrorString(e));
            vkt_bcn_cuda_destroy(ctx);
            return VKT_BCN_ERR_CUDA;
        }
    }
    *out_ctx = ctx;
    return VKT_BCN_OK;
}

void vkt_bcn_cuda_destroy(vkt_bcn_ctx *ctx)
{
    if(!ctx) { return; }
    for(auto *s: ctx->slots)
    {
        if(cudaSetDevice(s->device) == cudaSuccess)
        {
            if(s->stream)
            {
                cudaStreamSynchronize(s->stream);
                cudaStreamDestroy(s->stream);
            }
            cudaFree(s->d_tables);
            cudaFree(s->d_in);
            cudaFree(s->d_out);
            cudaFree(s->d_tmp);
            delete s->axis_cache;// frees the cached resize tables
        }
        delete s;
    }
    delete ctx;
}

int vkt_bcn_cuda_num_devices(const vkt_bcn_ctx *ctx) { return ctx ? int(ctx->slots.size()) : 0; }

const char *vkt_bcn_cuda_last_error(const vkt_bcn_ctx *ctx)
{
    if(!ctx) { return t_create_error.c_str(); }
    return ctx->last_error.c_str();
}

int vkt_bcn_cuda_get_stats(const vkt_bcn_ctx *ctx, vkt_bcn_stats *out)
{
    if(!ctx || !out) { return VKT_BCN_ERR_INVALID; }
    auto *c = const_cast<vkt_bcn_ctx *>(ctx);
    std::lock_guard<std::mutex> g(c->stats_mtx);
    *out = c->stats;
    return VKT_BCN_OK;
}

int vkt_bcn_cuda_encode_batch(vkt_bcn_ctx *ctx, uint32_t mode, const vkt_bcn_image *images, uint32_t num_images,
                              const vkt_bc7_params *params)
{
    if(!ctx) { return VKT_BCN_ERR_INVALID; }
    if(mode != VKT_BCN_MODE_BC7 && mode != VKT_BCN_MODE_BC5) { return fail(ctx, VKT_BCN_ERR_INVALID, "unknown mode %u", mode); }
    if(num_images && !images) { return fail(ctx, VKT_BCN_ERR_INVALID, "null image array"); }
    for(uint32_t i = 0; i < num_images; ++i)
    {
        uint32_t stride = images[i].row_stride_bytes;
        const int rc = check_image(ctx, images[i].pixels, images[i].width, images[i].height, images[i].comps, &stride, images[i].out_blocks);
        if(rc) { return rc; }
    }
    if(mode == VKT_BCN_MODE_BC7)
    {
        // validate parameters before any work is queued
        vkt_bc7_params def;
        vkt_bc7_params_init_inline(&def);
        Bc7KernelParams kp;
        const int rc = bc7_prepare_params(params ? params : &def, &kp);
        if(rc) { return fail(ctx, rc, rc == VKT_BCN_ERR_UNSUPPORTED ? "unsupported bc7 parameters" : "invalid bc7 parameters"); }
    }
    const uint32_t G = uint32_t(ctx->slots.size());
    // Partition: every image's block rows are split evenly over the G devices; images with fewer
    // block rows than devices go to one device, rotating.  Per slot, all its pieces are queued back to back on its
    // stream (H2D -> kernel -> D2H per piece), then every stream is synchronised once.
    struct Piece
    {
        uint32_t img, row0, row1;
        size_t in_off, out_off;
    };
    std::vector<std::vector<Piece>> plan(G);
    std::vector<size_t> in_need(G, 0), out_need(G, 0);
    uint32_t rr = 0;
    for(uint32_t i = 0; i < num_images; ++i)
    {
        const uint32_t rows = images[i].height / 4;
        const size_t row_in = size_t(images[i].width) * images[i].comps * 4, row_out = size_t(images[i].width / 4) * 16;
        if(rows < G * 4)
        {
            const uint32_t g = rr++ % G;
            plan[g].push_back({i, 0, rows, in_need[g], out_need[g]});
            in_need[g] += align_up(rows * row_in, 256), out_need[g] += align_up(rows * row_out, 256);
        }
        else
        {
            for(uint32_t g = 0; g < G; ++g)
            {
                const uint32_t r0 = uint32_t(uint64_t(rows) * g / G), r1 = uint32_t(uint64_t(rows) * (g + 1) / G);
                plan[g].push_back({i, r0, r1, in_need[g], out_need[g]});
                in_need[g] += align_up((r1 - r0) * row_in, 256), out_need[g] += align_up((r1 - r0) * row_out, 256);
            }
        }
    }
    int rc = VKT_BCN_OK;
    std::vector<std::unique_lock<std::mutex>> locks;
    for(uint32_t g = 0; g < G; ++g) { locks.emplace_back(ctx->slots[g]->mtx); }
    for(uint32_t g = 0; g < G && !rc; ++g)
    {
        DeviceSlot *s = ctx->slots[g];
        if(plan[g].empty()) { continue; }
        VKT_CUDA(ctx, cudaSetDevice(s->device));
        if((rc = ensure(ctx, &s->d_in, &s->in_cap, in_need[g]))) { break; }
        if((rc = ensure(ctx, &s->d_out, &s->out_cap, out_need[g]))) { break; }
    }
    // queue piece by piece, round-robin over devices so that all PCIe links start early
    size_t max_pieces = 0;
    for(auto &p: plan) { max_pieces = std::max(max_pieces, p.size()); }
    for(size_t k = 0; k < max_pieces && !rc; ++k)
    {
        for(uint32_t g = 0; g < G && !rc; ++g)
        {
            if(k >= plan[g].size()) { continue; }
            DeviceSlot *s = ctx->slots[g];
            const Piece &p = plan[g][k];
            VKT_CUDA(ctx, cudaSetDevice(s->device));
            rc = encode_rows_async(ctx, s, mode, images[p.img], p.row0, p.row1, params, p.in_off, p.out_off);
        }
    }
    for(uint32_t g = 0; g < G; ++g)
    {
        DeviceSlot *s = ctx->slots[g];
        if(plan[g].empty()) { continue; }
        if(cudaSetDevice(s->device) != cudaSuccess) { continue; }
        const cudaError_t e = cudaStreamSynchronize(s->stream);
        if(e != cudaSuccess && !rc) { rc = fail(ctx, VKT_BCN_ERR_CUDA, "stream synchronize failed: %s", cudaGetErrorString(e)); }
    }
    return rc;
}

int vkt_bcn_cuda_encode_bc7(vkt_bcn_ctx *ctx, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t comps,
                            uint32_t row_stride_bytes, const vkt_bc7_params *params, void *out_blocks)
{
    const vkt_bcn_image img{pixels, width, height, comps, row_stride_bytes, out_blocks};
    return vkt_bcn_cuda_encode_batch(ctx, VKT_BCN_MODE_BC7, &img, 1, params);
}

int vkt_bcn_cuda_encode_bc5(vkt_bcn_ctx *ctx, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t comps,
                            uint32_t row_stride_bytes, void *out_blocks)
{
    const vkt_bcn_image img{pixels, width, height, comps, row_stride_bytes, out_blocks};
    return vkt_bcn_cuda_encode_batch(ctx, VKT_BCN_MODE_BC5, &img, 1, nullptr);
}

static int device_entry(vkt_bcn_ctx *ctx, int slot, uint32_t mode, const void *d_pixels, uint32_t width, uint32_t height, uint32_t comps,
                        uint32_t stride, const vkt_bc7_params *params, void *d_out, void *cuda_stream)
{
    int rc = check_image(ctx, d_pixels, width, height, comps, &stride, d_out);
    if(rc) { return rc; }
    if(slot < 0 || slot >= int(ctx->slots.size())) { return fail(ctx, VKT_BCN_ERR_INVALID, "slot %d out of range", slot); }
    DeviceSlot *s = ctx->slots[size_t(slot)];
    std::lock_guard<std::mutex> g(s->mtx);
    VKT_CUDA(ctx, cudaSetDevice(s->device));
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : s->stream;
    rc = (mode == VKT_BCN_MODE_BC7) ? launch_bc7(ctx, s, d_pixels, width, height, comps, stride, params, d_out, st)
                                    : launch_bc5(ctx, s, d_pixels, width, height, comps, stride, d_out, st);
    if(rc) { return rc; }
    if(!cuda_stream) { VKT_CUDA(ctx, cudaStreamSynchronize(st)); }
    return VKT_BCN_OK;
}

int vkt_bcn_cuda_encode_bc7_device(vkt_bcn_ctx *ctx, int slot, const void *d_pixels, uint32_t width, uint32_t height, uint32_t comps,
                                   uint32_t row_stride_bytes, const vkt_bc7_params *params, void *d_out_blocks, void *cuda_stream)
{
    return device_entry(ctx, slot, VKT_BCN_MODE_BC7, d_pixels, width, height, comps, row_stride_bytes, params, d_out_blocks, cuda_stream);
}

int vkt_bcn_cuda_encode_bc5_device(vkt_bcn_ctx *ctx, int slot, const void *d_pixels, uint32_t width, uint32_t height, uint32_t comps,
                                   uint32_t row_stride_bytes, void *d_out_blocks, void *cuda_stream)
{
    return device_entry(ctx, slot, VKT_BCN_MODE_BC5, d_pixels, width, height, comps, row_stride_bytes, nullptr, d_out_blocks, cuda_stream);
}

// texture_block_compression.cpp:80-86,141-146
int vkt_bcn_cuda_compress_plan(uint32_t width, uint32_t height, int generate_mipmaps, vkt_bcn_plan *plan)
{
    if(!plan || width == 0 || height == 0) { return VKT_BCN_ERR_INVALID; }
    auto round4 = [](uint32_t v) { return (v + 3u) & ~3u; };
    uint32_t w = round4(width), h = round4(height);
    const uint32_t m = std::max(w, h);
    uint32_t lg = 0;// floor(log2(m)), exact in integers (the reference truncates log2(double) - 2)
    while((2u << lg) <= m) { ++lg; }
    const uint32_t max_levels = uint32_t(std::max<int32_t>(0, int32_t(lg) - 2)) + 1;
    std::memset(plan, 0, sizeof(*plan));
    plan->base_width = w, plan->base_height = h;
    plan->num_levels = generate_mipmaps ? max_levels : 1;
    for(uint32_t l = 0; l < plan->num_levels; ++l)
    {
        plan->level_width[l] = w, plan->level_height[l] = h;
        plan->level_num_blocks[l] = uint64_t(w / 4) * (h / 4);
        w = round4(std::max<uint32_t>(w / 2, 1)), h = round4(std::max<uint32_t>(h / 2, 1));
    }
    return VKT_BCN_OK;
}

int vkt_bcn_cuda_resize_u8(vkt_bcn_ctx *ctx, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t comps,
                           uint8_t *out_pixels, uint32_t out_width, uint32_t out_height)
{
    if(!ctx) { return VKT_BCN_ERR_INVALID; }
    return resize_host(ctx, pixels, width, height, comps, out_pixels, out_width, out_height);
}

int vkt_bcn_cuda_compress(vkt_bcn_ctx *ctx, uint32_t mode, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t comps,
                          int generate_mipmaps, const vkt_bc7_params *params, void *const *level_blocks)
{
    if(!ctx) { return VKT_BCN_ERR_INVALID; }
    return compress_chain(ctx, mode, pixels, width, height, comps, generate_mipmaps, params, level_blocks);
}

}// extern "C"
