/* oracle/bc7_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port" oracle) of the reference's BC7 hot path:
 *   bc7enc_compress_block()  /root/reference/extern/bc7enc_rdo/bc7enc.cpp:2402-2438 and everything below it.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may link or call this.
 * The product library (libvierkant_bcn_cuda) never does; it has no CPU path at all.
 *
 * Parity status: PINNED.  The restatement is checked byte-for-byte against the unmodified reference compiled in
 * place (oracle/_ref/libvkt_ref.so) by tests/test_oracle_pinning.py, and against the committed golden vectors in
 * tests/golden/ that were generated from that reference (tests/golden/make_golden.py).
 */
#ifndef VKT_BC7_ORACLE_H
#define VKT_BC7_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Field-by-field mirror of bc7enc_compress_block_params (bc7enc.h:14-75); same layout as oracle/ref_shim.cpp. */
typedef struct port_bc7_params
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
} port_bc7_params;

void port_bc7_params_init(port_bc7_params *p);

/* One 4x4 RGBA8 tile (64 bytes, row-major, R first) -> one 16-byte BC7 block.  Returns 1 if the alpha path ran. */
int port_bc7_encode_block(const uint8_t *rgba64, const port_bc7_params *params, uint8_t *out16);

/* Loop over pre-gathered tiles; threads <= 1 runs inline. */
void port_bc7_encode_blocks(const uint8_t *px, uint64_t num_blocks, const port_bc7_params *params, uint8_t *out,
                            int threads);

/* Decoder for the modes bc7enc emits (1, 5, 6, 7) plus the remaining BC7 modes; used for the PSNR fallback metric. */
void port_bc7_unpack_blocks(const uint8_t *blocks, uint64_t num_blocks, uint8_t *px);

/* Implemented in bc5_oracle.c / stbir_oracle.c / compress_oracle.c */
void port_bc5_encode_blocks(const uint8_t *px, uint64_t num_blocks, uint8_t *out);
void port_resize_u8(const uint8_t *in, uint32_t w, uint32_t h, uint32_t comps, uint8_t *out, uint32_t ow, uint32_t oh);
uint32_t port_compress_num_levels(uint32_t width, uint32_t height, int mipmaps);
uint32_t port_compress(const uint8_t *img, uint32_t width, uint32_t height, uint32_t comps, uint32_t mode, int mipmaps,
                       const port_bc7_params *params, uint8_t *const *level_blocks, int threads);

#ifdef __cplusplus
}
#endif
#endif
