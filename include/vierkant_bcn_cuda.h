/* vierkant_bcn_cuda.h -- C ABI of libvierkant_bcn_cuda, the B200 (sm_100a) drop-in for vierkant's BCn block encoders.
 *
 * The reference has no FFI for this path: vierkant::bcn::compress() (src/texture_block_compression.cpp:64-154) calls
 * the vendored bc7enc_rdo functions directly, one 4x4 block at a time:
 *     bc7enc_compress_block_init()                          extern/bc7enc_rdo/bc7enc.h:116   (call site :34)
 *     bc7enc_compress_block(pBlock, pPixelsRGBA, &params)   extern/bc7enc_rdo/bc7enc.h:121   (call site :132)
 *     rgbcx::encode_bc5(pDst, pPixels, 0, 1, 4)             extern/bc7enc_rdo/rgbcx.h:266    (call site :131)
 *     crocore::Image_<uint8_t>::resize -> stbir_resize_uint8  extern/crocore/src/Image.cpp:239-247 (call site :101)
 * The entry points below are what a binding for that path binds instead: they take a whole level image (or a whole
 * batch of level images) per call, because the unit a GPU wants is "all blocks of the level", not one block.
 * INTEGRATION.md shows the C++20 wrapper that keeps vierkant::bcn::compress()'s signature on top of this ABI.
 *
 * Conventions
 *   - plain C types only; the caller owns every buffer; all calls are synchronous unless a stream is passed
 *   - return value: VKT_BCN_OK (0) or a negative VKT_BCN_ERR_* code; vkt_bcn_cuda_last_error() gives the message
 *   - there is NO CPU fallback: without a usable CUDA device every call fails with VKT_BCN_ERR_NO_DEVICE
 *   - output blocks are bit-identical to the reference's (same inputs, same parameters); block order is row-major,
 *     blocks[bx + by * (width / 4)], 16 bytes each (vierkant::bcn::block_t, texture_block_compression.hpp:22-25)
 *   - a context may be used from several host threads at once: on a single-device context concurrent vkt_bcn_cuda_compress /
 *     _compress_alloc calls run on up to four lanes of the device (own streams and buffers each), so a loader that fans its
 *     textures out over threads overlaps their uploads, kernels and downloads; every other call is serialised per device slot
 */
#ifndef VIERKANT_BCN_CUDA_H
#define VIERKANT_BCN_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VKT_BCN_OK 0
#define VKT_BCN_ERR_INVALID (-1)     /* bad argument (null pointer, size not a multiple of 4, comps not 3/4, ...) */
#define VKT_BCN_ERR_UNSUPPORTED (-2) /* reserved (no parameter combination is refused as unsupported any more) */
#define VKT_BCN_ERR_CUDA (-3)        /* CUDA runtime failure, see vkt_bcn_cuda_last_error */
#define VKT_BCN_ERR_NO_DEVICE (-4)   /* no CUDA device / driver: the library has no CPU path */
#define VKT_BCN_ERR_OOM (-5)

/* == vierkant::bcn::CompressionMode (include/vierkant/texture_block_compression.hpp:15-19) */
#define VKT_BCN_MODE_BC5 0u
#define VKT_BCN_MODE_BC7 1u

typedef struct vkt_bcn_ctx vkt_bcn_ctx;

/* Field-by-field mirror of bc7enc_compress_block_params (extern/bc7enc_rdo/bc7enc.h:14-75).  Never memcpy the C++
 * struct across the ABI (it contains bools and padding); vkt_bc7_params_init() = bc7enc_compress_block_params_init()
 * (bc7enc.h:95-113), which is what vierkant uses (texture_block_compression.cpp:73-74).
 *
 * Every field is honoured.  force_selectors / selectors, quant_mode6_endpoints and low_frequency_partition_weight != 1.0
 * (only driven by bc7enc_rdo's RDO post-processor, which vierkant does not compile, src/CMakeLists.txt:11) run on a
 * separate, slower kernel variant that keeps the estimator's early-outs (bc7enc.cpp:1536,1810), which such a weight makes
 * observable.  VKT_BCN_ERR_INVALID: uber_level > 4; a mode_mask without an opaque (6 | 1) or an alpha (5 | 6 | 7) mode (the
 * reference asserts), or with mode 1 as the only opaque mode and max_partitions == 0 (the reference then encodes an uninitialised result); with force_selectors, a selector that does not exist in the palette of every enabled mode (16 / 8 / 4
 * entries for mode 6 / 1 / 5 and 7 -- the reference would read an uninitialised colour); a low-frequency weight that is
 * negative, above 65536 or NaN (the reference's float -> uint64 conversion is undefined there). */
typedef struct vkt_bc7_params
{
    uint32_t mode_mask;
    uint32_t max_partitions; /* 0..64 */
    uint32_t weights[4];
    uint32_t uber_level; /* 0..4 */
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
} vkt_bc7_params;

void vkt_bc7_params_init(vkt_bc7_params *p);

/* One level image and where its blocks go.  width and height must be multiples of 4 (vierkant::bcn::compress rounds
 * up and resizes before encoding, texture_block_compression.cpp:80-81,101); comps is 3 or 4 (3 => alpha := 255, as
 * get_block does, texture_block_compression.cpp:39-60); row_stride_bytes 0 means tightly packed. */
typedef struct vkt_bcn_image
{
    const uint8_t *pixels;
    uint32_t width, height, comps, row_stride_bytes;
    void *out_blocks; /* (width / 4) * (height / 4) * 16 bytes */
} vkt_bcn_image;

/* Number of usable CUDA devices (0 if none; never fails). */
int vkt_bcn_cuda_device_count(void);

/* Create a context over `num_devices` CUDA device ordinals (devices == NULL: ordinals 0..num_devices-1; num_devices
 * <= 0: every visible device).  Builds the encoder tables (the job of bc7enc_compress_block_init / rgbcx::init) and
 * uploads them once per device. */
int vkt_bcn_cuda_create(vkt_bcn_ctx **out_ctx, const int *devices, int num_devices);
void vkt_bcn_cuda_destroy(vkt_bcn_ctx *ctx);
int vkt_bcn_cuda_num_devices(const vkt_bcn_ctx *ctx);

/* Message of the last failure on this context (ctx == NULL: of the last failed create on this thread). */
const char *vkt_bcn_cuda_last_error(const vkt_bcn_ctx *ctx);

/* Replaces the per-block bc7enc_compress_block loop of one level (texture_block_compression.cpp:107-139).
 * Host buffers in, host buffers out; block rows are split across the context's devices. params == NULL: defaults. */
int vkt_bcn_cuda_encode_bc7(vkt_bcn_ctx *ctx, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t comps,
                            uint32_t row_stride_bytes, const vkt_bc7_params *params, void *out_blocks);

/* Same for the BC5 branch (rgbcx::encode_bc5(pBlock, pixels, 0, 1, 4), texture_block_compression.cpp:131). */
int vkt_bcn_cuda_encode_bc5(vkt_bcn_ctx *ctx, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t comps,
                            uint32_t row_stride_bytes, void *out_blocks);

/* All levels of a mip chain, or all textures of a model, in one call (what model::compress_textures loops over,
 * src/model/model_loading.cpp:96-118): images are distributed over the devices, copies and kernels are pipelined on
 * per-device streams, results are gathered with async device-to-host copies.  mode is VKT_BCN_MODE_*. */
int vkt_bcn_cuda_encode_batch(vkt_bcn_ctx *ctx, uint32_t mode, const vkt_bcn_image *images, uint32_t num_images,
                              const vkt_bc7_params *params);

/* Kernel-only variants on device-resident buffers of device slot `slot` (0 <= slot < num_devices): d_pixels is
 * RGBA8/RGB8 in HBM, d_out_blocks receives 16 B per block.  cuda_stream is a cudaStream_t (NULL = the slot's own
 * stream, synchronised before return; otherwise the launch is asynchronous on that stream). */
int vkt_bcn_cuda_encode_bc7_device(vkt_bcn_ctx *ctx, int slot, const void *d_pixels, uint32_t width, uint32_t height,
                                   uint32_t comps, uint32_t row_stride_bytes, const vkt_bc7_params *params,
                                   void *d_out_blocks, void *cuda_stream);
int vkt_bcn_cuda_encode_bc5_device(vkt_bcn_ctx *ctx, int slot, const void *d_pixels, uint32_t width, uint32_t height,
                                   uint32_t comps, uint32_t row_stride_bytes, void *d_out_blocks, void *cuda_stream);

/* Several device-resident images (all levels of a chain, all textures of a material) in one go: every 16 images share
 * ONE set of kernel launches, so a tail of tiny mip levels costs nothing extra.  images[i].pixels / .out_blocks are
 * device pointers on slot `slot`; stream semantics as above. */
int vkt_bcn_cuda_encode_batch_device(vkt_bcn_ctx *ctx, int slot, uint32_t mode, const vkt_bcn_image *images,
                                     uint32_t num_images, const vkt_bc7_params *params, void *cuda_stream);

/* crocore::Image_<uint8_t>::resize == stbir_resize_uint8 with its defaults (extern/crocore/src/Image.cpp:239-247),
 * bit-exact, on the GPU.  Host in / host out, tightly packed, comps 1..4. */
int vkt_bcn_cuda_resize_u8(vkt_bcn_ctx *ctx, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t comps,
                           uint8_t *out_pixels, uint32_t out_width, uint32_t out_height);

/* The whole of vierkant::bcn::compress() (texture_block_compression.cpp:64-154): round the size up to a multiple of 4,
 * resize level 0 and every further level from the previous one (on the GPU, stbir-exact), encode every level.
 * Call vkt_bcn_cuda_compress_plan first to get the level count and per-level block counts, then pass
 * level_blocks[l] buffers of 16 * num_blocks[l] bytes.  Only the source image crosses PCIe on the way in.
 * `pixels` and every `level_blocks[l]` may be host memory (pinned for full speed, pageable works) or CUDA device memory
 * (unified addressing decides per pointer): with device destinations -- for instance the VkBuffer vierkant uploads from,
 * exported with VK_KHR_external_memory_fd and mapped by cudaImportExternalMemory / cudaExternalMemoryGetMappedBuffer --
 * the blocks never touch the host (what create_compressed_texture's staging copy costs today,
 * src/model/model_loading.cpp:483-488).  Same for vkt_bcn_cuda_compress_batch. */
typedef struct vkt_bcn_plan
{
    uint32_t base_width, base_height, num_levels;
    uint32_t level_width[16], level_height[16];
    uint64_t level_num_blocks[16];
} vkt_bcn_plan;
int vkt_bcn_cuda_compress_plan(uint32_t width, uint32_t height, int generate_mipmaps, vkt_bcn_plan *plan);
int vkt_bcn_cuda_compress(vkt_bcn_ctx *ctx, uint32_t mode, const uint8_t *pixels, uint32_t width, uint32_t height,
                          uint32_t comps, int generate_mipmaps, const vkt_bc7_params *params, void *const *level_blocks);

/* vkt_bcn_cuda_compress for a caller that allocates its result per call -- vierkant::bcn::compress() returns a fresh
 * compress_result_t whose levels are std::vector<block_t> (include/vierkant/texture_block_compression.hpp:27-44;
 * src/texture_block_compression.cpp:88-96 sizes them before the first block is encoded).  Here alloc_level(user, level, bytes) is
 * called -- once per level, level 0 first, never concurrently, from a helper thread of the library -- while the calling thread
 * queues the chain and the GPU already works, so the allocation and zero fill of the result (22 MB of fresh pages for a 4096^2
 * chain) overlap the upload and the kernels instead of preceding them.  It returns host memory of `bytes` bytes that stays valid until the call returns; the blocks are handed over
 * as their downloads land.  A null return aborts the call with VKT_BCN_ERR_OOM (nothing is left in flight). */
typedef void *(*vkt_bcn_alloc_fn)(void *user, uint32_t level, size_t bytes);
int vkt_bcn_cuda_compress_alloc(vkt_bcn_ctx *ctx, uint32_t mode, const uint8_t *pixels, uint32_t width, uint32_t height,
                                uint32_t comps, int generate_mipmaps, const vkt_bc7_params *params, vkt_bcn_alloc_fn alloc_level,
                                void *user);

/* Several textures, each with its whole chain, in one call -- what model::compress_textures does texture by texture
 * (src/model/model_loading.cpp:96-118).  Every device of the context works on two textures at a time
 * (two sets of streams and buffers), so the upload and resizes of the next texture run under the encode kernels of the
 * current one and no device idles between textures; textures go round-robin over the devices.  Results are identical
 * to calling vkt_bcn_cuda_compress for each texture.  level_blocks[i] is the level_blocks argument of that call for
 * texture i (vkt_bcn_cuda_compress_plan(width, height, generate_mipmaps) gives its sizes). */
typedef struct vkt_bcn_source
{
    const uint8_t *pixels;          /* tightly packed, row-major */
    uint32_t width, height, comps;  /* comps: 3 or 4 */
    uint32_t mode;                  /* VKT_BCN_MODE_* (a material mixes BC7 colour maps and BC5 normal maps) */
    void *const *level_blocks;
} vkt_bcn_source;
int vkt_bcn_cuda_compress_batch(vkt_bcn_ctx *ctx, const vkt_bcn_source *sources, uint32_t num_sources, int generate_mipmaps,
                                const vkt_bc7_params *params);

/* ONE chain split over several PROCESSES, one GPU each (north_star: "independent block rows and mip levels
 * shard across the 8 GPUs of one box by simple partitioning, with no NCCL collective, and results are gathered back over
 * pinned async copies").  The reference shards a level into batches of four block rows for its thread pool
 * (src/texture_block_compression.cpp:107-139); here worker `rank` of `world` (a single-device context each) takes the
 * block rows [rows * rank / world, rows * (rank + 1) / world) of every level that is large enough to slice, uploads only the
 * source rows those reach through the filter taps (own rows + a halo of a few rows per level, recomputed instead of
 * exchanged), and writes its blocks at their final position inside level_blocks[l] -- so when `pixels` and `level_blocks`
 * are one buffer shared by the workers (POSIX shared memory registered with vkt_bcn_cuda_host_register, or any pinned
 * mapping), every GPU's copy engine gathers straight into the caller's contiguous level arrays.  Levels too small to slice
 * (fewer than 128 pixel rows or 4 block rows per worker) are finished by worker 0; it continues the chain from the last
 * sliced level, whose rows the other workers leave in `handover` (host memory all workers see, `handover_bytes` large):
 *     every worker:  vkt_bcn_cuda_compress_shard_begin(...)   queue own slices; returns once own rows are in `handover`
 *     the caller:    a barrier over the workers (any mechanism; no data moves)
 *     every worker:  vkt_bcn_cuda_compress_shard_end(...)     worker 0 queues the small levels; all wait for their work
 * The union of the workers' blocks is byte-identical to vkt_bcn_cuda_compress on one device.  `pixels` / `level_blocks[l]`
 * may be host or device memory as for vkt_bcn_cuda_compress (a device-resident source on the worker's own GPU is read in
 * place).  begin and end must be called from the same host thread with the same arguments; no other call may be made on
 * the context in between. */
typedef struct vkt_bcn_shard_plan
{
    uint32_t num_levels;
    uint32_t sliced_levels;  /* levels [0, sliced_levels) are split by block rows over `workers` (0: worker 0 does everything) */
    uint32_t workers;        /* workers that take part (1 if the image is too small to slice) */
    uint64_t handover_bytes; /* size of the hand-over buffer (0: none needed) */
} vkt_bcn_shard_plan;
int vkt_bcn_cuda_compress_shard_plan(uint32_t width, uint32_t height, int generate_mipmaps, uint32_t world, vkt_bcn_shard_plan *out);
/* block rows [*first_block_row, *end_block_row) of `level` that worker `rank` encodes (empty for levels it has no part in) */
int vkt_bcn_cuda_compress_shard_rows(uint32_t width, uint32_t height, int generate_mipmaps, uint32_t rank, uint32_t world,
                                     uint32_t level, uint32_t *first_block_row, uint32_t *end_block_row);
int vkt_bcn_cuda_compress_shard_begin(vkt_bcn_ctx *ctx, uint32_t mode, const uint8_t *pixels, uint32_t width, uint32_t height,
                                      uint32_t comps, int generate_mipmaps, const vkt_bc7_params *params, uint32_t rank,
                                      uint32_t world, void *const *level_blocks, void *handover);
int vkt_bcn_cuda_compress_shard_end(vkt_bcn_ctx *ctx, uint32_t mode, const uint8_t *pixels, uint32_t width, uint32_t height,
                                    uint32_t comps, int generate_mipmaps, const vkt_bc7_params *params, uint32_t rank,
                                    uint32_t world, void *const *level_blocks, void *handover);

/* Page-lock caller memory (cudaHostRegister, portable) so that copies from / to it run asynchronously at full link speed:
 * what a loader does once for a long-lived decode buffer or for the shared result buffer of a sharded chain.  Pageable
 * memory works everywhere without this (the library stages it), only slower. */
int vkt_bcn_cuda_host_register(vkt_bcn_ctx *ctx, void *ptr, size_t bytes);
int vkt_bcn_cuda_host_unregister(vkt_bcn_ctx *ctx, void *ptr);

/* The Vulkan hand-off: vierkant copies every level from the compress_result_t into a host-visible staging
 * VkBuffer and uploads from there (create_compressed_texture, src/model/model_loading.cpp:460-494).  With the staging buffer's
 * memory exported as an opaque fd (VK_KHR_external_memory_fd, vkGetMemoryFdKHR) this call maps it into the context's device
 * `slot` (cudaImportExternalMemory + cudaExternalMemoryGetMappedBuffer); *d_ptr can then be passed as a level_blocks[l]
 * destination (plus an offset) to vkt_bcn_cuda_compress / _batch / _shard_*, and the blocks land in the Vulkan buffer without
 * touching the host.  CUDA takes ownership of `fd` on success.  Release with vkt_bcn_cuda_release_external before the Vulkan
 * memory is freed.  INTEGRATION.md section 5 shows both sides. */
int vkt_bcn_cuda_import_external_fd(vkt_bcn_ctx *ctx, int slot, int fd, uint64_t bytes, void **d_ptr, void **external_handle);
int vkt_bcn_cuda_release_external(vkt_bcn_ctx *ctx, int slot, void *d_ptr, void *external_handle);

/* Counters for the measurement harness: kernels launched / bytes copied by this context since creation. */
typedef struct vkt_bcn_stats
{
    uint64_t kernel_launches;
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
} vkt_bcn_stats;
int vkt_bcn_cuda_get_stats(const vkt_bcn_ctx *ctx, vkt_bcn_stats *out);


/* Measurement support: sustained issue rate of the device's integer pipes in lane-operations per second (a short
 * probe kernel with a balanced FMA-pipe / ALU-pipe mix), the denominator of the ALU roofline bench.py reports. */
int vkt_bcn_cuda_measure_issue_peak(vkt_bcn_ctx *ctx, int slot, double *lane_ops_per_second);

#ifdef __cplusplus
}
#endif
#endif
