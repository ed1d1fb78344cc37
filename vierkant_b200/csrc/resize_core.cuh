// resize_core.cuh -- placeholder until the stbir-exact GPU resize lands in this round.
#pragma once
namespace vkt
{
static int resize_host(vkt_bcn_ctx *ctx, const uint8_t *, uint32_t, uint32_t, uint32_t, uint8_t *, uint32_t, uint32_t)
{
    return fail(ctx, -2, "vkt_bcn_cuda_resize_u8: not implemented yet");
}
static int compress_chain(vkt_bcn_ctx *ctx, uint32_t, const uint8_t *, uint32_t, uint32_t, uint32_t, int, const vkt_bc7_params *, void *const *)
{
    return fail(ctx, -2, "vkt_bcn_cuda_compress: not implemented yet");
}
}// namespace vkt
