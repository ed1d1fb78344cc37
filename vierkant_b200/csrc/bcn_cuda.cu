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
#include <atomic>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#define VKT_BCN_DEFINE_PARAMS_INIT
#include "bc5_core.cuh"
#include "bc7_core.cuh"
#include "bc7_params.h"
#include "host_copy.h"

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
#ifndef VKT_BC7_THREADS_ALPHA
#define VKT_BC7_THREADS_ALPHA 256
#endif
constexpr int kBc7Threads = VKT_BC7_THREADS, kBc7ThreadsAlpha = VKT_BC7_THREADS_ALPHA;
__host__ __device__ constexpr int bc7_threads(bool alpha) { return alpha ? kBc7ThreadsAlpha : kBc7Threads; }
// Opaque blocks: 3 CTAs/SM = 3 x (64 KB lane columns + tables) of shared memory at <= 80 registers per thread.
// Alpha blocks carry a fourth channel through every stage: 2 CTAs/SM at <= 128 registers (no spills) is faster.
// So are the opaque kernels that carry the uber-level stages (selector-search dominated: +2.4 % at uber level 4).
#ifndef VKT_BC7_CTAS_UBER
#define VKT_BC7_CTAS_UBER 2
#endif
constexpr int kBc7CtasPerSm = VKT_BC7_CTAS, kBc7CtasPerSmAlpha = VKT_BC7_CTAS_ALPHA, kBc7CtasPerSmUber = VKT_BC7_CTAS_UBER;

// Up to 16 images (the levels of a mip chain, the textures of a material, or row slices of them) are encoded by ONE
// launch: the block index space of the launch is the concatenation of the images' block arrays.  A tail of tiny mip
// levels otherwise costs one full block latency (~130 us) per level, serialised on the stream.
constexpr uint32_t kBc7MaxImages = 16;
struct Bc7Image
{
    const uint8_t *img;
    uint4 *out;
    uint32_t blocks_x, first_block;// first_block: position of the image's block 0 in the launch's index space
    uint32_t comps, stride;
    int vec16;
};
struct Bc7Batch
{
    Bc7Image im[kBc7MaxImages];
    uint32_t num_images, total_blocks;
};

__device__ __forceinline__ uint32_t batch_find(const Bc7Batch &B, uint32_t g)
{
    uint32_t k = 0;
#pragma unroll
    for(uint32_t i = 1; i < kBc7MaxImages; ++i) { k += (i < B.num_images && g >= B.im[i].first_block) ? 1u : 0u; }
    return k;
}

// Stage 0: split the launch's blocks into an opaque and an alpha work list (the dispatch of bc7enc_compress_block,
// bc7enc.cpp:2422-2437), so that every warp of the encode kernels holds blocks of one kind.  counts[0] = #opaque,
// counts[1] = #alpha.  List order is irrelevant to the output (a block always lands in its own slot); warp-aggregated
// atomics keep it nearly sorted, so the encode kernels' texel loads stay coalesced.
__global__ void __launch_bounds__(256) bc7_classify_kernel(const __grid_constant__ Bc7Batch B, uint32_t *__restrict__ counts,
                                                            uint32_t *__restrict__ list_opaque, uint32_t *__restrict__ list_alpha)
{
    const uint32_t g = blockIdx.x * 256 + threadIdx.x;
    const bool valid = g < B.total_blocks;
    bool alpha = false;
    if(valid)
    {
        const Bc7Image &I = B.im[batch_find(B, g)];
        const uint32_t b = g - I.first_block;
        const uint32_t bx = b % I.blocks_x, by = b / I.blocks_x;
        uint32_t and_all = 0xFFFFFFFFu;
        if(I.comps == 4)// 3-component images are opaque by construction (get_block injects alpha = 255)
        {
            if(I.vec16)
            {
#pragma unroll
                for(int y = 0; y < 4; ++y)
                {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(I.img + size_t(by * 4 + y) * I.stride) + bx);
                    and_all &= v.x & v.y & v.z & v.w;
                }
            }
            else
            {
                for(int y = 0; y < 4; ++y)
                {
                    const uint8_t *row = I.img + size_t(by * 4 + y) * I.stride + size_t(bx) * 16;
                    for(int x = 0; x < 4; ++x) { and_all &= uint32_t(row[4 * x + 3]) << 24; }
                }
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
        if(alpha) { list_alpha[base_a + __popc(m_alpha & below)] = g; }
        else { list_opaque[base_o + __popc(m_opaque & below)] = g; }
    }
}

// (the mode-7 single-colour table at the end of Bc7Tables is read from global memory through Bc7KernelParams::opt7: 16 loads per
// mode-7 cell, and 4 KB less shared memory per CTA)
__host__ __device__ constexpr size_t bc7_smem_table_bytes(bool) { return offsetof(Bc7Tables, opt7); }
static_assert(offsetof(Bc7Tables, opt7) % 16 == 0, "table prefix is copied as uint4");

// One lane == one block of the work list (list == nullptr: every block of the launch, in order).
// UBER == false: the search without the uber-level stages (launched when uber_level == 0).
template<bool PERC, int KV, bool ALPHA, bool UBER, int NT>
__global__ void __launch_bounds__(NT, ALPHA ? kBc7CtasPerSmAlpha : (UBER ? kBc7CtasPerSmUber : kBc7CtasPerSm))
        bc7_encode_kernel(const __grid_constant__ Bc7Batch B, const Bc7KernelParams P, const Bc7Tables *__restrict__ g_tables,
                          const uint32_t *__restrict__ list, const uint32_t *__restrict__ count)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const uint32_t n = count ? __ldg(count) : B.total_blocks;
    if(blockIdx.x * NT >= n) { return; }// the grid is sized for the whole launch; lists are usually shorter
    // shared memory: [tables (opaque kernels: without the mode-7 table at their end)][lane columns][exchange area]
    constexpr size_t kTab = bc7_smem_table_bytes(ALPHA);
    Bc7Tables &s_tables = *reinterpret_cast<Bc7Tables *>(s_raw);
    Texel *s_lane = reinterpret_cast<Texel *>(s_raw + kTab);// one 16-record column per lane (see Lane<>)
    CtaScratch<NT> *s_scratch = reinterpret_cast<CtaScratch<NT> *>(s_lane + size_t(NT) * 16);
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(g_tables);
        uint4 *dst = reinterpret_cast<uint4 *>(&s_tables);
        for(uint32_t i = threadIdx.x; i < kTab / 16; i += NT) { dst[i] = __ldg(src + i); }
        if(threadIdx.x < 192) { (&s_scratch->cnt[0][0])[threadIdx.x] = 0u; }// cnt[8][16] and cnt2[8][8] are adjacent
    }
    const uint32_t i = blockIdx.x * NT + threadIdx.x;
    const uint32_t ii = min(i, n - 1);// out-of-range lanes redo the last block (keeps warps converged), no store
    const uint32_t g = list ? __ldg(list + ii) : ii;
    const Bc7Image &I = B.im[batch_find(B, g)];
    const uint32_t b = g - I.first_block;
    Lane<NT> lane{s_lane + threadIdx.x};
    load_block_texels<NT>(I.img, I.comps, I.stride, I.vec16 != 0, b % I.blocks_x, b / I.blocks_x, lane.p);
    __syncthreads();
    uint32_t blk[4];
    // the lane may finish another block of the CTA than the one it loaded (encode_block regroups them): the block's index in
    // the launch travels with it
    const uint32_t done = encode_block<PERC, KV, ALPHA, UBER, NT>(s_tables, P, lane, blk, (i < n) ? g : 0xFFFFFFFFu);
    if(done != 0xFFFFFFFFu)
    {
        const Bc7Image &J = B.im[batch_find(B, done)];
        J.out[done - J.first_block] = make_uint4(blk[0], blk[1], blk[2], blk[3]);
    }
}

constexpr size_t bc7_smem_bytes(bool alpha)
{
    return bc7_smem_table_bytes(alpha) + size_t(bc7_threads(alpha)) * 16 * sizeof(Texel) + (alpha ? sizeof(CtaScratch<kBc7ThreadsAlpha>) : sizeof(CtaScratch<kBc7Threads>));
}

template<bool PERC, int KV, bool ALPHA, bool UBER>
static cudaError_t bc7_kernel_attribute()
{
    return cudaFuncSetAttribute(bc7_encode_kernel<PERC, KV, ALPHA, UBER, bc7_threads(ALPHA)>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                int(bc7_smem_bytes(ALPHA)));
}
// > 48 KB of dynamic shared memory needs an explicit opt-in per kernel (and per device: called from context creation)
static cudaError_t bc7_kernel_attributes()
{
    cudaError_t e = bc7_kernel_attribute<true, true, false, false>();
    if(e == cudaSuccess) { e = bc7_kernel_attribute<true, true, true, false>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<true, true, false, true>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<true, true, true, true>(); }
#ifndef VKT_BC7_DEV_DEFAULT_VARIANTS_ONLY
    if(e == cudaSuccess) { e = bc7_kernel_attribute<false, true, false, false>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<false, true, true, false>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<true, false, false, true>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<true, false, true, true>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<false, true, false, true>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<false, true, true, true>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<false, false, false, true>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<false, false, true, true>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<true, kKvExt, false, true>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<true, kKvExt, true, true>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<false, kKvExt, false, true>(); }
    if(e == cudaSuccess) { e = bc7_kernel_attribute<false, kKvExt, true, true>(); }
#endif
    return e;
}

// ------------------------------------------------------------------------------------------------ issue-rate probe
// Measurement support: the ALU roofline's denominator.  Eight independent chains per thread, half the
// instructions on the FMA pipe (IMAD), half on the ALU pipe (LOP3 / IADD3) -- the mix at which an SM sub-partition can
// issue one warp instruction per clock.  32 lane-operations per warp instruction.
__global__ void __launch_bounds__(256) alu_probe_kernel(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t a[8], b[8];
#pragma unroll
    for(int k = 0; k < 8; ++k) { a[k] = seed + threadIdx.x * 8u + k, b[k] = seed * 3u + k; }
#pragma unroll 1
    for(int i = 0; i < iters; ++i)
    {
#pragma unroll
        for(int r = 0; r < 4; ++r)
        {
#pragma unroll
            for(int k = 0; k < 8; ++k)
            {
                a[k] = a[k] * 0x9E3779B1u + b[k];// IMAD
                b[k] = b[k] ^ a[k];// LOP3
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for(int k = 0; k < 8; ++k) { acc += a[k] ^ b[k]; }
    if(acc == 0x12345u) { out[blockIdx.x * 256 + threadIdx.x] = acc; }
}
constexpr int kProbeOpsPerIter = 4 * 8 * 2;// per thread and iteration: 4 rounds x 8 chains x (IMAD + LOP3)

// ------------------------------------------------------------------------------------------------ context
}// namespace vkt
struct vkt_axis_cache;// resize_core.cuh
namespace vkt
{
struct DeviceSlot
{
    vkt_axis_cache *axis_cache = nullptr;
    int device = -1;
    cudaStream_t stream = nullptr, stream2 = nullptr, stream3 = nullptr;// 2: uploads / second batch lane, 3: downloads
    cudaStream_t stream4 = nullptr, stream5 = nullptr, stream6 = nullptr;// encode lanes of the pipelined chain (round robin)
    Bc7Tables *d_tables = nullptr;
    void *d_in = nullptr, *d_out = nullptr, *d_tmp = nullptr;
    size_t in_cap = 0, out_cap = 0, tmp_cap = 0;
    std::vector<cudaEvent_t> event_pool;// ordering events of the pipelined chain (guarded by mtx like the buffers)
    size_t events_used = 0;
    // pageable caller memory (host_copy.h): pinned staging for the source rows this device reads and for its blocks, and the
    // copies from the block staging into the caller's buffers that chain_wait() still owes (each once its event has fired)
    void *h_in = nullptr, *h_out = nullptr;
    size_t h_in_cap = 0, h_out_cap = 0;
    struct PendingCopy
    {
        cudaEvent_t ready;
        const void *src;
        void *dst;
        size_t bytes;
        uint32_t level = 0;// of a compress() chain (a deferred destination is resolved from level + offset)
        size_t offset = 0;
    };
    std::vector<PendingCopy> pending;
    bool src_in_place = false;// the running chain reads its source where the caller keeps it (device memory of this device)
    std::mutex mtx;
};

}// namespace vkt

struct vkt_bcn_ctx
{
    std::vector<vkt::DeviceSlot *> slots;
    std::vector<vkt::DeviceSlot *> slots2;// further sets of streams / buffers per device (lanes; entry k: device k % G), made on first use
    std::mutex lanes_mtx;                  // guards slots2 (its size and its growth); never held while a slot mutex is waited for
    vkt::Bc7Tables host_tables;
    std::string last_error;
    std::mutex err_mtx;
    vkt_bcn_stats stats{};
    std::mutex stats_mtx;
    void *h_stage = nullptr;// pinned gather buffer of the multi-device chain (used with every slot mutex held)
    std::unique_ptr<vkt::CopyPool> copy_pool;// made on first use (a caller with pageable buffers)
    std::mutex copy_pool_mtx;
    size_t stage_cap = 0;
    std::unique_lock<std::mutex> shard_lock;// slot 0's mutex, held from compress_shard_begin to compress_shard_end
    std::atomic<bool> shard_open{false};
};

namespace vkt
{
static thread_local std::string t_create_error;
static thread_local std::string t_last_error;// vkt_bcn_cuda_last_error's answer outlives other threads' failures

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

static int ensure_pinned(vkt_bcn_ctx *ctx, void **ptr, size_t *cap, size_t need)
{
    if(*cap >= need) { return VKT_BCN_OK; }
    if(*ptr) { VKT_CUDA(ctx, cudaFreeHost(*ptr)); }
    *ptr = nullptr, *cap = 0;
    VKT_CUDA(ctx, cudaHostAlloc(ptr, need, cudaHostAllocPortable));
    *cap = need;
    return VKT_BCN_OK;
}

// Host memory CUDA does not know (malloc, std::vector): copies from / to it would be staged by the driver on the calling thread.
static bool is_pageable_host(const void *p)
{
    cudaPointerAttributes a;
    if(cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// device memory of device `dev` (an allocation of this process or an imported one)?
static bool device_pointer_on(const void *p, int dev)
{
    cudaPointerAttributes a;
    if(cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice && a.device == dev;
}

static CopyPool &copy_pool(vkt_bcn_ctx *ctx)
{
    std::lock_guard<std::mutex> g(ctx->copy_pool_mtx);
    if(!ctx->copy_pool)
    {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        unsigned workers = std::min(7u, std::max(1u, hw / 2));
        if(const char *e = getenv("VKT_BCN_COPY_THREADS")) { workers = unsigned(std::max(1, atoi(e))) - 1u; }// (tuning; 1 = caller only)
        ctx->copy_pool = std::make_unique<CopyPool>(workers);
    }
    return *ctx->copy_pool;
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

// One device-resident image (or row slice) of a launch.
struct DevImage
{
    const void *d_px;
    uint32_t w, h, comps, stride;
    void *d_out;
};

// launch the BC7 kernels on device-resident data (current device must be the slot's): all images in as few launches as
// possible (kBc7MaxImages per launch)
static int launch_bc7_batch(vkt_bcn_ctx *ctx, DeviceSlot *s, const DevImage *images, uint32_t num_images, const vkt_bc7_params *params,
                            cudaStream_t stream)
{
    vkt_bc7_params def;
    if(!params)
    {
        vkt_bc7_params_init_inline(&def);
        params = &def;
    }
    Bc7KernelParams kp;
    const int rc = bc7_prepare_params(params, &kp);
    if(rc)
    {
        return fail(ctx, rc, "invalid bc7 parameters (mode_mask must enable mode 6 or 1 and one of 5/6/7; uber_level <= 4; forced selectors "
                             "must exist in every enabled mode's palette; 0 <= low_frequency_partition_weight <= 65536)");
    }
    kp.m6_reduced = reinterpret_cast<const uint8_t *>(s->d_tables + 1);
    kp.opt7 = reinterpret_cast<const uint32_t *>(reinterpret_cast<const char *>(s->d_tables) + offsetof(Bc7Tables, opt7));
    for(uint32_t first = 0; first < num_images; first += kBc7MaxImages)
    {
        Bc7Batch B = {};
        B.num_images = std::min<uint32_t>(kBc7MaxImages, num_images - first);
        uint64_t total = 0;
        bool any4 = false;
        for(uint32_t k = 0; k < B.num_images; ++k)
        {
            const DevImage &d = images[first + k];
            Bc7Image &I = B.im[k];
            I.img = static_cast<const uint8_t *>(d.d_px), I.out = static_cast<uint4 *>(d.d_out);
            I.blocks_x = d.w / 4, I.first_block = uint32_t(total), I.comps = d.comps, I.stride = d.stride;
            I.vec16 = (d.comps == 4) && ((d.stride & 15u) == 0) && ((reinterpret_cast<uintptr_t>(d.d_px) & 15u) == 0);
            total += uint64_t(d.w / 4) * (d.h / 4);
            any4 = any4 || d.comps == 4;
        }
        if(total == 0) { continue; }
        if(total > 0xFFFFFFF0ull) { return fail(ctx, VKT_BCN_ERR_INVALID, "more than 2^32 blocks in one launch"); }
        B.total_blocks = uint32_t(total);
        auto encode = [&](bool alpha, const uint32_t *list, const uint32_t *cnt) {
            const uint32_t nt = uint32_t(bc7_threads(alpha)), grid = (B.total_blocks + nt - 1) / nt;
            auto go = [&](auto kernel) { kernel<<<grid, nt, bc7_smem_bytes(alpha), stream>>>(B, kp, s->d_tables, list, cnt); };
            // uber-free kernels exist for the 28-bit-key variants (every sane weight set); the wide-error ones always carry the stages
            int sel = ((kp.uber_level == 0 && kp.key28) ? 8 : 0) | (params->perceptual ? 4 : 0) | (kp.key28 ? 2 : 0) | (alpha ? 1 : 0);
            // forced selectors / reduced mode-6 quantisation / low-frequency partition weight: the extended variant (wide errors, uber stages)
            if(kp.ext) { sel = 16 | (params->perceptual ? 2 : 0) | (alpha ? 1 : 0); }
            switch(sel)
            {
#ifndef VKT_BC7_DEV_DEFAULT_VARIANTS_ONLY
                case 19: go(bc7_encode_kernel<true, kKvExt, true, true, kBc7ThreadsAlpha>); break;
                case 18: go(bc7_encode_kernel<true, kKvExt, false, true, kBc7Threads>); break;
                case 17: go(bc7_encode_kernel<false, kKvExt, true, true, kBc7ThreadsAlpha>); break;
                case 16: go(bc7_encode_kernel<false, kKvExt, false, true, kBc7Threads>); break;
#endif
                case 15: go(bc7_encode_kernel<true, true, true, false, kBc7ThreadsAlpha>); break;
                case 14: go(bc7_encode_kernel<true, true, false, false, kBc7Threads>); break;
                case 7: go(bc7_encode_kernel<true, true, true, true, kBc7ThreadsAlpha>); break;
                case 6: go(bc7_encode_kernel<true, true, false, true, kBc7Threads>); break;
#ifndef VKT_BC7_DEV_DEFAULT_VARIANTS_ONLY// (tuning builds compile the perceptual 28-bit-key kernels only: defaults and uber levels)
                case 11: go(bc7_encode_kernel<false, true, true, false, kBc7ThreadsAlpha>); break;
                case 10: go(bc7_encode_kernel<false, true, false, false, kBc7Threads>); break;
                case 5: go(bc7_encode_kernel<true, false, true, true, kBc7ThreadsAlpha>); break;
                case 4: go(bc7_encode_kernel<true, false, false, true, kBc7Threads>); break;
                case 3: go(bc7_encode_kernel<false, true, true, true, kBc7ThreadsAlpha>); break;
                case 2: go(bc7_encode_kernel<false, true, false, true, kBc7Threads>); break;
                case 1: go(bc7_encode_kernel<false, false, true, true, kBc7ThreadsAlpha>); break;
                default: go(bc7_encode_kernel<false, false, false, true, kBc7Threads>); break;
#else
                default: break;
#endif
            }
        };
        if(kp.force_alpha)
        {
            encode(true, nullptr, nullptr);
            count(ctx, 1, 0, 0);
        }
        else if(!any4)
        {
            encode(false, nullptr, nullptr);// get_block injects alpha = 255: every block is opaque
            count(ctx, 1, 0, 0);
        }
        else
        {
            // stream-ordered scratch: [counts: 2 x u32, padded to 256 B][opaque list][alpha list]
            uint8_t *scratch = nullptr;
            const size_t list_bytes = align_up(size_t(B.total_blocks) * sizeof(uint32_t), 256);
            VKT_CUDA(ctx, cudaMallocAsync(reinterpret_cast<void **>(&scratch), 256 + 2 * list_bytes, stream));
            uint32_t *counts = reinterpret_cast<uint32_t *>(scratch);
            uint32_t *list_o = reinterpret_cast<uint32_t *>(scratch + 256), *list_a = reinterpret_cast<uint32_t *>(scratch + 256 + list_bytes);
            cudaError_t e = cudaMemsetAsync(counts, 0, 8, stream);
            if(e == cudaSuccess)
            {
                bc7_classify_kernel<<<(B.total_blocks + 255) / 256, 256, 0, stream>>>(B, counts, list_o, list_a);
                encode(false, list_o, counts);
                encode(true, list_a, counts + 1);
                e = cudaGetLastError();
            }
            VKT_CUDA(ctx, cudaFreeAsync(scratch, stream));// (also when something above failed: the scratch never leaks)
            VKT_CUDA(ctx, e);
            count(ctx, 3, 0, 0);
        }
        VKT_CUDA(ctx, cudaGetLastError());
    }
    return VKT_BCN_OK;
}

static int launch_bc7(vkt_bcn_ctx *ctx, DeviceSlot *s, const void *d_px, uint32_t w, uint32_t h, uint32_t comps, uint32_t stride,
                      const vkt_bc7_params *params, void *d_out, cudaStream_t stream)
{
    const DevImage d{d_px, w, h, comps, stride, d_out};
    return launch_bc7_batch(ctx, s, &d, 1, params, stream);
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

}// namespace vkt

// streams, tables and kernel attributes of one slot (a device, or a device's second lane)
static cudaError_t init_slot(vkt::DeviceSlot *s, const vkt::Bc7Tables &host_tables)
{
    using namespace vkt;
    const int dev = s->device;
        cudaError_t e = cudaSetDevice(dev);
    // `stream` carries the (cheap) resize kernels of the pipelined chain: highest priority, so that their CTAs are placed
    // ahead of the pending CTAs of the long encode kernels running on the other lanes
    int prio_lo = 0, prio_hi = 0;
    if(e == cudaSuccess) { e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi); }
    if(e == cudaSuccess) { e = cudaStreamCreateWithPriority(&s->stream, cudaStreamNonBlocking, prio_hi); }
    if(e == cudaSuccess) { e = cudaStreamCreateWithFlags(&s->stream2, cudaStreamNonBlocking); }
    if(e == cudaSuccess) { e = cudaStreamCreateWithFlags(&s->stream3, cudaStreamNonBlocking); }
    if(e == cudaSuccess) { e = cudaStreamCreateWithFlags(&s->stream4, cudaStreamNonBlocking); }
    if(e == cudaSuccess) { e = cudaStreamCreateWithFlags(&s->stream5, cudaStreamNonBlocking); }
    if(e == cudaSuccess) { e = cudaStreamCreateWithFlags(&s->stream6, cudaStreamNonBlocking); }
    // [Bc7Tables][reduced mode-6 quantiser table] in one allocation (the second one stays in global memory: kKvExt only)
    if(e == cudaSuccess) { e = cudaMalloc(reinterpret_cast<void **>(&s->d_tables), sizeof(Bc7Tables) + kBc7M6ReducedBytes); }
    if(e == cudaSuccess) { e = cudaMemcpy(s->d_tables, &host_tables, sizeof(Bc7Tables), cudaMemcpyHostToDevice); }
    if(e == cudaSuccess)
    {
        std::vector<uint8_t> m6(kBc7M6ReducedBytes);
        bc7_m6_reduced_build(m6.data());
        e = cudaMemcpy(s->d_tables + 1, m6.data(), kBc7M6ReducedBytes, cudaMemcpyHostToDevice);
    }
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
    return e;
}

static void destroy_slot(vkt::DeviceSlot *s);

#include "resize_core.cuh"

using namespace vkt;

// streams, events, buffers and cached tables of one slot (also a partly initialised one: every member starts out null)
static void destroy_slot(vkt::DeviceSlot *s)
{
    if(cudaSetDevice(s->device) == cudaSuccess)
    {
        for(cudaStream_t st: {s->stream, s->stream3, s->stream4, s->stream5, s->stream6, s->stream2})
        {
            if(st)
            {
                cudaStreamSynchronize(st);
                cudaStreamDestroy(st);
            }
        }
        for(cudaEvent_t ev: s->event_pool) { cudaEventDestroy(ev); }
        cudaFree(s->d_tables);
        cudaFree(s->d_in);
        cudaFree(s->d_out);
        cudaFree(s->d_tmp);
        if(s->h_in) { cudaFreeHost(s->h_in); }
        if(s->h_out) { cudaFreeHost(s->h_out); }
        delete s->axis_cache;// frees the cached resize tables
    }
    delete s;
}

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
    bc7_tables_build(&ctx->host_tables);
    {
        // the compile-time uber-level selector maps (bc7_core.cuh) against the reference's float expression
        const UberMaps um = make_uber_maps();
        for(int k = 0; k < 3; ++k)
        {
            const int max_sel = (k == 0) ? 3 : (k == 1) ? 7 : 15;
            const uint64_t live = (max_sel == 15) ? ~0ull : ((1ull << (4 * (max_sel + 1))) - 1ull);
            for(int ly = -2; ly <= 1; ++ly)
            {
                for(int hy = max_sel - 1; hy <= max_sel + 2; ++hy)
                {
                    if((um.m[k][ly + 2][hy - (max_sel - 1)] & live) != (bc7_uber_map_reference(max_sel, ly, hy) & live))
                    {
                        vkt_bcn_cuda_destroy(ctx);
                        return fail(nullptr, VKT_BCN_ERR_INVALID, "internal: uber selector map mismatch (max %d, ly %d, hy %d)", max_sel, ly, hy);
                    }
                }
            }
        }
    }
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
        const cudaError_t e = init_slot(s, ctx->host_tables);
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
    if(ctx->h_stage) { cudaFreeHost(ctx->h_stage); }
    std::vector<vkt::DeviceSlot *> all(ctx->slots);
    all.insert(all.end(), ctx->slots2.begin(), ctx->slots2.end());
    for(auto *s: all) { destroy_slot(s); }
    delete ctx;
}

int vkt_bcn_cuda_num_devices(const vkt_bcn_ctx *ctx) { return ctx ? int(ctx->slots.size()) : 0; }

const char *vkt_bcn_cuda_last_error(const vkt_bcn_ctx *ctx)
{
    if(!ctx) { return t_create_error.c_str(); }
    {
        // copied under the lock into a per-thread buffer: another thread's failure may reassign ctx->last_error at any time
        auto *c = const_cast<vkt_bcn_ctx *>(ctx);
        std::lock_guard<std::mutex> g(c->err_mtx);
        t_last_error = c->last_error;
    }
    return t_last_error.c_str();
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
    // block rows than devices go to one device, rotating.  Per device the pieces are cut into groups (<= 16 images,
    // ~1 M blocks): a group is H2D copies -> ONE encode launch over all its pieces -> D2H copies, and consecutive
    // groups alternate between the slot's two streams so that the copies of one overlap the kernels of the other.
    struct Piece
    {
        uint32_t img, row0, row1;
        size_t in_off, out_off;
    };
    static const uint64_t kBandBlocks = getenv("VKT_BCN_BAND_BLOCKS") ? strtoull(getenv("VKT_BCN_BAND_BLOCKS"), nullptr, 10) : (1u << 18);// (tuning; measured 2^20 / 2^18 / 2^17 / 2^16: 3.58 / 2.56 / 2.76 / 2.84 ms for a pinned 4096^2 level)
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
                // a device's share of a large image is cut into row bands of about kBandBlocks blocks: each band is a group of
                // its own (upload -> launch -> download on alternating streams), so the transfers of one band run under the
                // kernels of its neighbours instead of before and after one big launch
                const uint64_t blocks_per_row = images[i].width / 4;
                const uint32_t band_rows = uint32_t(std::max<uint64_t>(4, (kBandBlocks + blocks_per_row - 1) / blocks_per_row));
                for(uint32_t a = r0; a < r1; a += band_rows)
                {
                    const uint32_t b = std::min(r1, a + band_rows);
                    plan[g].push_back({i, a, b, in_need[g], out_need[g]});
                    in_need[g] += align_up((b - a) * row_in, 256), out_need[g] += align_up((b - a) * row_out, 256);
                }
            }
        }
    }
    int rc = VKT_BCN_OK;
    std::vector<char> pageable_in(num_images), pageable_out(num_images);
    for(uint32_t i = 0; i < num_images; ++i) { pageable_in[i] = is_pageable_host(images[i].pixels), pageable_out[i] = is_pageable_host(images[i].out_blocks); }
    std::vector<std::unique_lock<std::mutex>> locks;
    for(uint32_t g = 0; g < G; ++g) { locks.emplace_back(ctx->slots[g]->mtx); }
    for(uint32_t g = 0; g < G && !rc; ++g)
    {
        DeviceSlot *s = ctx->slots[g];
        if(plan[g].empty()) { continue; }
        if(cudaSetDevice(s->device) != cudaSuccess)
        {
            rc = fail(ctx, VKT_BCN_ERR_CUDA, "cudaSetDevice(%d) failed", s->device);
            break;
        }
        if((rc = ensure(ctx, &s->d_in, &s->in_cap, in_need[g]))) { break; }
        if((rc = ensure(ctx, &s->d_out, &s->out_cap, out_need[g]))) { break; }
        // pageable images / block buffers are staged through pinned mirrors of d_in / d_out (see host_copy.h)
        bool page_in = false, page_out = false;
        for(const Piece &p: plan[g]) { page_in = page_in || pageable_in[p.img], page_out = page_out || pageable_out[p.img]; }
        if(page_in && (rc = ensure_pinned(ctx, &s->h_in, &s->h_in_cap, in_need[g]))) { break; }
        if(page_out && (rc = ensure_pinned(ctx, &s->h_out, &s->h_out_cap, out_need[g]))) { break; }
        s->pending.clear(), s->events_used = 0;
    }
    // queue group by group, round-robin over devices so that all PCIe links start early.  A failure inside queue_group
    // only ends the queueing: the synchronise loop below still runs, so no copy into / out of the caller's buffers is in
    // flight when this call reports the error.
    static const uint64_t kGroupBlocks = getenv("VKT_BCN_GROUP_BLOCKS") ? strtoull(getenv("VKT_BCN_GROUP_BLOCKS"), nullptr, 10) : kBandBlocks;// (tuning)
    std::vector<size_t> next(G, 0), group_no(G, 0);
    auto queue_group = [&](uint32_t g) -> int {
        DeviceSlot *s = ctx->slots[g];
        VKT_CUDA(ctx, cudaSetDevice(s->device));
        cudaStream_t st = (group_no[g]++ & 1) ? s->stream2 : s->stream;
        const size_t first = next[g];
        uint64_t blocks = 0;
        std::vector<DevImage> dev;
        while(next[g] < plan[g].size() && dev.size() < kBc7MaxImages && (dev.empty() || blocks < kGroupBlocks))
        {
            const Piece &p = plan[g][next[g]++];
            const vkt_bcn_image &img = images[p.img];
            const uint32_t stride = img.row_stride_bytes ? img.row_stride_bytes : img.width * img.comps;
            const uint32_t rows = p.row1 - p.row0;
            const size_t row_bytes = size_t(img.width) * img.comps;
            uint8_t *d_in = static_cast<uint8_t *>(s->d_in) + p.in_off;
            // the device copy is tightly packed
            const uint8_t *from = img.pixels + size_t(p.row0) * 4 * stride;
            if(pageable_in[p.img])
            {
                uint8_t *staged = static_cast<uint8_t *>(s->h_in) + p.in_off;
                if(stride == row_bytes) { copy_pool(ctx).copy(staged, from, size_t(rows) * 4 * row_bytes); }
                else
                {
                    for(size_t r = 0; r < size_t(rows) * 4; ++r) { memcpy(staged + r * row_bytes, from + r * stride, row_bytes); }
                }
                VKT_CUDA(ctx, cudaMemcpyAsync(d_in, staged, size_t(rows) * 4 * row_bytes, cudaMemcpyHostToDevice, st));
            }
            else
            {
                VKT_CUDA(ctx, cudaMemcpy2DAsync(d_in, row_bytes, from, stride, row_bytes, size_t(rows) * 4, cudaMemcpyDefault, st));
            }
            dev.push_back({d_in, img.width, rows * 4, img.comps, uint32_t(row_bytes), static_cast<uint8_t *>(s->d_out) + p.out_off});
            blocks += uint64_t(rows) * (img.width / 4);
            count(ctx, 0, size_t(rows) * 4 * row_bytes, 0);
        }
        int r = VKT_BCN_OK;
        if(mode == VKT_BCN_MODE_BC7) { r = launch_bc7_batch(ctx, s, dev.data(), uint32_t(dev.size()), params, st); }
        else
        {
            for(const DevImage &d: dev)
            {
                if((r = launch_bc5(ctx, s, d.d_px, d.w, d.h, d.comps, d.stride, d.d_out, st))) { break; }
            }
        }
        for(size_t k = first; k < next[g] && !r; ++k)
        {
            const Piece &p = plan[g][k];
            const vkt_bcn_image &img = images[p.img];
            const size_t row_blk = size_t(img.width / 4) * 16, bytes = size_t(p.row1 - p.row0) * row_blk;
            uint8_t *user = static_cast<uint8_t *>(img.out_blocks) + size_t(p.row0) * row_blk;
            if(pageable_out[p.img])
            {
                uint8_t *staged = static_cast<uint8_t *>(s->h_out) + p.out_off;
                VKT_CUDA(ctx, cudaMemcpyAsync(staged, static_cast<uint8_t *>(s->d_out) + p.out_off, bytes, cudaMemcpyDeviceToHost, st));
                if(s->events_used == s->event_pool.size())
                {
                    cudaEvent_t fresh;
                    VKT_CUDA(ctx, cudaEventCreateWithFlags(&fresh, cudaEventDisableTiming));
                    s->event_pool.push_back(fresh);
                }
                cudaEvent_t landed = s->event_pool[s->events_used++];
                VKT_CUDA(ctx, cudaEventRecord(landed, st));
                s->pending.push_back({landed, staged, user, bytes, 0u, size_t(0)});
            }
            else { VKT_CUDA(ctx, cudaMemcpyAsync(user, static_cast<uint8_t *>(s->d_out) + p.out_off, bytes, cudaMemcpyDefault, st)); }
            count(ctx, 0, 0, bytes);
        }
        return r;
    };
    bool more = !rc;
    while(more && !rc)
    {
        more = false;
        for(uint32_t g = 0; g < G && !rc; ++g)
        {
            if(next[g] >= plan[g].size()) { continue; }
            rc = queue_group(g);
            more = more || next[g] < plan[g].size();
        }
    }
    for(uint32_t g = 0; g < G; ++g)
    {
        DeviceSlot *s = ctx->slots[g];
        if(plan[g].empty()) { continue; }
        if(cudaSetDevice(s->device) != cudaSuccess) { continue; }
        cudaError_t e = cudaSuccess;
        for(const DeviceSlot::PendingCopy &pc: s->pending)// blocks staged for pageable destinations, each as soon as it has landed
        {
            const cudaError_t e1 = cudaEventSynchronize(pc.ready);
            if(e1 == cudaSuccess && !rc) { copy_pool(ctx).copy(pc.dst, pc.src, pc.bytes); }
            else if(e1 != cudaSuccess && e == cudaSuccess) { e = e1; }
        }
        s->pending.clear();
        const cudaError_t e0 = cudaStreamSynchronize(s->stream), e2 = cudaStreamSynchronize(s->stream2);
        if(e == cudaSuccess) { e = (e0 != cudaSuccess) ? e0 : e2; }
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

int vkt_bcn_cuda_encode_batch_device(vkt_bcn_ctx *ctx, int slot, uint32_t mode, const vkt_bcn_image *images, uint32_t num_images,
                                     const vkt_bc7_params *params, void *cuda_stream)
{
    if(!ctx) { return VKT_BCN_ERR_INVALID; }
    if(mode != VKT_BCN_MODE_BC7 && mode != VKT_BCN_MODE_BC5) { return fail(ctx, VKT_BCN_ERR_INVALID, "unknown mode %u", mode); }
    if(num_images && !images) { return fail(ctx, VKT_BCN_ERR_INVALID, "null image array"); }
    if(slot < 0 || slot >= int(ctx->slots.size())) { return fail(ctx, VKT_BCN_ERR_INVALID, "slot %d out of range", slot); }
    std::vector<DevImage> dev;
    for(uint32_t i = 0; i < num_images; ++i)
    {
        uint32_t stride = images[i].row_stride_bytes;
        const int rc = check_image(ctx, images[i].pixels, images[i].width, images[i].height, images[i].comps, &stride, images[i].out_blocks);
        if(rc) { return rc; }
        dev.push_back({images[i].pixels, images[i].width, images[i].height, images[i].comps, stride, images[i].out_blocks});
    }
    DeviceSlot *s = ctx->slots[size_t(slot)];
    std::lock_guard<std::mutex> g(s->mtx);
    VKT_CUDA(ctx, cudaSetDevice(s->device));
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : s->stream;
    int rc = VKT_BCN_OK;
    if(mode == VKT_BCN_MODE_BC7) { rc = launch_bc7_batch(ctx, s, dev.data(), uint32_t(dev.size()), params, st); }
    else
    {
        for(const DevImage &d: dev)
        {
            if((rc = launch_bc5(ctx, s, d.d_px, d.w, d.h, d.comps, d.stride, d.d_out, st))) { break; }
        }
    }
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

int vkt_bcn_cuda_measure_issue_peak(vkt_bcn_ctx *ctx, int slot, double *lane_ops_per_second)
{
    if(!ctx || !lane_ops_per_second) { return VKT_BCN_ERR_INVALID; }
    if(slot < 0 || slot >= int(ctx->slots.size())) { return fail(ctx, VKT_BCN_ERR_INVALID, "slot %d out of range", slot); }
    DeviceSlot *s = ctx->slots[size_t(slot)];
    std::lock_guard<std::mutex> g(s->mtx);
    VKT_CUDA(ctx, cudaSetDevice(s->device));
    cudaDeviceProp prop;
    VKT_CUDA(ctx, cudaGetDeviceProperties(&prop, s->device));
    const int grid = prop.multiProcessorCount * 8, iters = 4096;
    int rc = ensure(ctx, &s->d_tmp, &s->tmp_cap, size_t(grid) * 256 * sizeof(uint32_t));
    if(rc) { return rc; }
    cudaEvent_t e0, e1;
    VKT_CUDA(ctx, cudaEventCreate(&e0));
    VKT_CUDA(ctx, cudaEventCreate(&e1));
    float best = 0.0f;
    for(int rep = 0; rep < 4; ++rep)
    {
        VKT_CUDA(ctx, cudaEventRecord(e0, s->stream));
        alu_probe_kernel<<<grid, 256, 0, s->stream>>>(static_cast<uint32_t *>(s->d_tmp), iters, 17u + rep);
        VKT_CUDA(ctx, cudaEventRecord(e1, s->stream));
        VKT_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.0f;
        VKT_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if(rep > 0 && (best == 0.0f || ms < best)) { best = ms; }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *lane_ops_per_second = double(grid) * 256.0 * iters * kProbeOpsPerIter / (double(best) * 1e-3);
    return VKT_BCN_OK;
}

int vkt_bcn_cuda_resize_u8(vkt_bcn_ctx *ctx, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t comps,
                           uint8_t *out_pixels, uint32_t out_width, uint32_t out_height)
{
    if(!ctx) { return VKT_BCN_ERR_INVALID; }
    return resize_host(ctx, pixels, width, height, comps, out_pixels, out_width, out_height);
}

int vkt_bcn_cuda_compress_batch(vkt_bcn_ctx *ctx, const vkt_bcn_source *sources, uint32_t num_sources, int generate_mipmaps,
                                const vkt_bc7_params *params)
{
    if(!ctx) { return VKT_BCN_ERR_INVALID; }
    if(num_sources && !sources) { return fail(ctx, VKT_BCN_ERR_INVALID, "null source list"); }
    return compress_many(ctx, sources, num_sources, generate_mipmaps, params);
}

int vkt_bcn_cuda_compress(vkt_bcn_ctx *ctx, uint32_t mode, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t comps,
                          int generate_mipmaps, const vkt_bc7_params *params, void *const *level_blocks)
{
    if(!ctx) { return VKT_BCN_ERR_INVALID; }
    return compress_chain(ctx, mode, pixels, width, height, comps, generate_mipmaps, params, level_blocks);
}

int vkt_bcn_cuda_compress_alloc(vkt_bcn_ctx *ctx, uint32_t mode, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t comps,
                                int generate_mipmaps, const vkt_bc7_params *params, vkt_bcn_alloc_fn alloc_level, void *user)
{
    if(!ctx) { return VKT_BCN_ERR_INVALID; }
    return compress_chain_alloc(ctx, mode, pixels, width, height, comps, generate_mipmaps, params, alloc_level, user);
}

// ---- one chain split over several processes (one GPU each) ------------------------------------------------------------
int vkt_bcn_cuda_compress_shard_plan(uint32_t width, uint32_t height, int generate_mipmaps, uint32_t world, vkt_bcn_shard_plan *out)
{
    vkt_bcn_plan plan;
    if(!out || world == 0 || vkt_bcn_cuda_compress_plan(width, height, generate_mipmaps, &plan)) { return VKT_BCN_ERR_INVALID; }
    const ChainSplit split = chain_split(plan.level_height, plan.num_levels, world);
    std::memset(out, 0, sizeof(*out));
    out->num_levels = plan.num_levels, out->sliced_levels = split.devices > 1 ? split.sliced : 0, out->workers = split.devices;
    out->handover_bytes = split.tail() ? uint64_t(plan.level_width[split.sliced - 1]) * plan.level_height[split.sliced - 1] * 4u : 0u;
    return VKT_BCN_OK;
}

int vkt_bcn_cuda_compress_shard_rows(uint32_t width, uint32_t height, int generate_mipmaps, uint32_t rank, uint32_t world, uint32_t level,
                                     uint32_t *first_block_row, uint32_t *end_block_row)
{
    vkt_bcn_plan plan;
    if(!first_block_row || !end_block_row || world == 0 || rank >= world || vkt_bcn_cuda_compress_plan(width, height, generate_mipmaps, &plan))
    {
        return VKT_BCN_ERR_INVALID;
    }
    if(level >= plan.num_levels) { return VKT_BCN_ERR_INVALID; }
    const ChainSplit split = chain_split(plan.level_height, plan.num_levels, world);
    const uint32_t rows = plan.level_height[level] / 4;
    if(split.devices > 1 && level < split.sliced)
    {
        *first_block_row = uint32_t(uint64_t(rows) * rank / split.devices), *end_block_row = uint32_t(uint64_t(rows) * (rank + 1) / split.devices);
    }
    else { *first_block_row = rank == 0 ? 0u : rows, *end_block_row = rows; }
    return VKT_BCN_OK;
}

int vkt_bcn_cuda_compress_shard_begin(vkt_bcn_ctx *ctx, uint32_t mode, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t comps,
                                      int generate_mipmaps, const vkt_bc7_params *params, uint32_t rank, uint32_t world,
                                      void *const *level_blocks, void *handover)
{
    if(!ctx) { return VKT_BCN_ERR_INVALID; }
    if(ctx->slots.size() != 1) { return fail(ctx, VKT_BCN_ERR_INVALID, "a process shard runs on a single-device context (%zu devices)", ctx->slots.size()); }
    // (checked before the mutex is touched: the thread that opened a shard still holds it)
    if(ctx->shard_open) { return fail(ctx, VKT_BCN_ERR_INVALID, "compress_shard_begin without the compress_shard_end of the previous one"); }
    std::unique_lock<std::mutex> lock(ctx->slots[0]->mtx);
    ChainShard sh;
    sh.rank = rank, sh.world = world, sh.handover = static_cast<uint8_t *>(handover), sh.phase = 1;
    int rc = chain_enqueue(ctx, ctx->slots, mode, pixels, width, height, comps, generate_mipmaps, params, level_blocks, nullptr, &sh);
    if(!rc && sh.handed_over)
    {
        // the caller's barrier must mean "every worker's rows are in the hand-over buffer"
        const cudaError_t e = cudaEventSynchronize(sh.handed_over);
        if(e != cudaSuccess) { rc = fail(ctx, VKT_BCN_ERR_CUDA, "hand-over copy failed: %s", cudaGetErrorString(e)); }
    }
    if(rc)
    {
        chain_wait(ctx, ctx->slots);// nothing of a failed call stays in flight
        return rc;
    }
    ctx->shard_lock = std::move(lock), ctx->shard_open = true;
    return VKT_BCN_OK;
}

int vkt_bcn_cuda_compress_shard_end(vkt_bcn_ctx *ctx, uint32_t mode, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t comps,
                                    int generate_mipmaps, const vkt_bc7_params *params, uint32_t rank, uint32_t world,
                                    void *const *level_blocks, void *handover)
{
    if(!ctx) { return VKT_BCN_ERR_INVALID; }
    if(!ctx->shard_open || !ctx->shard_lock.owns_lock()) { return fail(ctx, VKT_BCN_ERR_INVALID, "compress_shard_end without compress_shard_begin"); }
    ChainShard sh;
    sh.rank = rank, sh.world = world, sh.handover = static_cast<uint8_t *>(handover), sh.phase = 2;
    int rc = chain_enqueue(ctx, ctx->slots, mode, pixels, width, height, comps, generate_mipmaps, params, level_blocks, nullptr, &sh);
    const int rw = chain_wait(ctx, ctx->slots);
    if(!rc) { rc = rw; }
    ctx->shard_open = false;
    ctx->shard_lock.unlock();
    ctx->shard_lock = std::unique_lock<std::mutex>();
    return rc;
}

int vkt_bcn_cuda_import_external_fd(vkt_bcn_ctx *ctx, int slot, int fd, uint64_t bytes, void **d_ptr, void **external_handle)
{
    if(!ctx || !d_ptr || !external_handle || fd < 0 || !bytes) { return ctx ? fail(ctx, VKT_BCN_ERR_INVALID, "bad external memory arguments") : VKT_BCN_ERR_INVALID; }
    if(slot < 0 || slot >= int(ctx->slots.size())) { return fail(ctx, VKT_BCN_ERR_INVALID, "slot %d out of range", slot); }
    *d_ptr = nullptr, *external_handle = nullptr;
    VKT_CUDA(ctx, cudaSetDevice(ctx->slots[size_t(slot)]->device));
    cudaExternalMemoryHandleDesc hd = {};
    hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
    hd.handle.fd = fd;
    hd.size = bytes;
    cudaExternalMemory_t ext = nullptr;
    VKT_CUDA(ctx, cudaImportExternalMemory(&ext, &hd));
    cudaExternalMemoryBufferDesc bd = {};
    bd.offset = 0, bd.size = bytes;
    void *p = nullptr;
    const cudaError_t e = cudaExternalMemoryGetMappedBuffer(&p, ext, &bd);
    if(e != cudaSuccess)
    {
        cudaDestroyExternalMemory(ext);
        return fail(ctx, VKT_BCN_ERR_CUDA, "cudaExternalMemoryGetMappedBuffer failed: %s", cudaGetErrorString(e));
    }
    *d_ptr = p, *external_handle = ext;
    return VKT_BCN_OK;
}

int vkt_bcn_cuda_release_external(vkt_bcn_ctx *ctx, int slot, void *d_ptr, void *external_handle)
{
    if(!ctx || !external_handle) { return VKT_BCN_ERR_INVALID; }
    if(slot < 0 || slot >= int(ctx->slots.size())) { return fail(ctx, VKT_BCN_ERR_INVALID, "slot %d out of range", slot); }
    VKT_CUDA(ctx, cudaSetDevice(ctx->slots[size_t(slot)]->device));
    if(d_ptr) { VKT_CUDA(ctx, cudaFree(d_ptr)); }// the mapping of cudaExternalMemoryGetMappedBuffer
    VKT_CUDA(ctx, cudaDestroyExternalMemory(static_cast<cudaExternalMemory_t>(external_handle)));
    return VKT_BCN_OK;
}

int vkt_bcn_cuda_host_register(vkt_bcn_ctx *ctx, void *ptr, size_t bytes)
{
    if(!ctx || !ptr || !bytes) { return VKT_BCN_ERR_INVALID; }
    VKT_CUDA(ctx, cudaSetDevice(ctx->slots[0]->device));
    VKT_CUDA(ctx, cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return VKT_BCN_OK;
}

int vkt_bcn_cuda_host_unregister(vkt_bcn_ctx *ctx, void *ptr)
{
    if(!ctx || !ptr) { return VKT_BCN_ERR_INVALID; }
    VKT_CUDA(ctx, cudaHostUnregister(ptr));
    return VKT_BCN_OK;
}

}// extern "C"
