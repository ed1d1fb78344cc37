// resize_core.cuh -- crocore::Image_<uint8_t>::resize on the GPU, bit-exact, and the whole vierkant::bcn::compress() chain.
//
// Reference: Image_<uint8_t>::resize (/root/reference/extern/crocore/src/Image.cpp:239-247) == stbir_resize_uint8 with its
// defaults (extern/crocore/src/stb_image_resize.h:2462-2470): Catmull-Rom where a dimension is enlarged, Mitchell otherwise
// (also at 1:1 -- which is NOT the identity), clamp-to-edge, linear colour space.  vierkant::bcn::compress resizes level 0
// and then every level from the previous one (src/texture_block_compression.cpp:99-101,141-146).
//
// Design.  stbir is separable and, per output sample, a fixed-order sum:
//     out(x, y) = SUM_rows( SUM_cols( u8 / 255.0f * ch ) * cv ),  every sum starting from 0.0f, terms in ascending input
//     index, multiply and add rounded separately (the reference build has no FMA), margin samples being clamped
//     copies of the edge sample that still enter as separate terms.
// The host builds the reference's coefficient tables with the same float/double arithmetic (ResizeAxis::build: sample
// ranges :1009-1038, kernels :816-844, per-list normalisation :1040-1199 -- including the quirk that a 5-entry list
// spills into its successor's first slot of the flat table) and turns both the gather (enlarging) and the scatter
// (reducing) formulation into one per-output tap list.  Two kernels then evaluate the sums exactly in that order:
// a horizontal pass into an fp32 band buffer and a vertical pass that also encodes ((int)(sat(f) * 255.0f + 0.5),
// :1737-1763).  Both passes are a few taps per sample and far below the encode kernels' cost; they are HBM-light because
// only a band of rows is kept in fp32.
#pragma once
#include <chrono>
#include <map>
#include <memory>

#include "chain_plan.h"
#include "resize_axis.h"
#include "resize_strip.h"

namespace vkt
{

// device copy of one axis (cached per context slot and (in, out) pair)
struct DeviceAxis
{
    int in_size = 0, out_size = 0;
    int *d_start = nullptr;
    int2 *d_tap = nullptr;// {input sample, coefficient bits}: one 64-bit load per tap
    std::vector<int> first_in, last_in;// per output: smallest / largest input sample (band planning)
    // Regular axes (every power-of-two chain level): output o reads exactly `reg_taps` samples, clamp(reg_step * o + reg_off + t),
    // t ascending -- 3 taps at 1:1 (Mitchell, x-1 .. x+1), 8 taps at 2:1 (2x-3 .. 2x+4).  reg_taps == 0: irregular, general passes.
    int reg_taps = 0, reg_step = 0, reg_off = 0;
    // ... and when every output carries the SAME reg_taps coefficient bits (true for both regular cases: the filter is sampled at
    // the same offsets for every output), they are kept here and handed to the strip kernel by value
    bool reg_uniform = false;
    float reg_coef[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    ~DeviceAxis()
    {
        cudaFree(d_start);
        cudaFree(d_tap);
    }
};

// ------------------------------------------------------------------------------------------------ device passes
constexpr int kResizeTileW = 32, kResizeTileH = 8;// threads of a 256-thread CTA as a tile of output samples
// Horizontal pass: rows [row0, row0 + rows) of the source -> fp32 band, one thread per (row, output column); a CTA is a
// 32-column x 8-row tile, so its eight rows share their tap lists through L1.  (Decoding the tile's source run once into
// shared memory instead of per tap was measured slower: the pass is bound by its loads, not by the decode arithmetic.)
template<int C>
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t *__restrict__ src, int in_w, int row0, int rows, int out_w,
                                                        const int *__restrict__ start, const int2 *__restrict__ tap, float *__restrict__ band)
{
    const int x = blockIdx.x * kResizeTileW + int(threadIdx.x % kResizeTileW), r = blockIdx.y * kResizeTileH + int(threadIdx.x / kResizeTileW);
    if(x >= out_w || r >= rows) { return; }
    const uint8_t *row = src + size_t(row0 + r) * size_t(in_w) * C;
    float acc[C];
#pragma unroll
    for(int c = 0; c < C; ++c) { acc[c] = 0.0f; }
    const int t1 = __ldg(start + x + 1);
    for(int t = __ldg(start + x); t < t1; ++t)
    {
        const int2 tp = __ldg(tap + t);
        const uint8_t *p = row + size_t(tp.x) * C;
        const float w = __int_as_float(tp.y);
        uint32_t v[C];
        if(C == 4)
        {
            const uint32_t q = __ldg(reinterpret_cast<const uint32_t *>(p));
            v[0] = q & 255u, v[1 % C] = (q >> 8) & 255u, v[2 % C] = (q >> 16) & 255u, v[3 % C] = q >> 24;
        }
        else
        {
#pragma unroll
            for(int c = 0; c < C; ++c) { v[c] = p[c]; }
        }
#pragma unroll
        for(int c = 0; c < C; ++c) { acc[c] = __fadd_rn(acc[c], __fmul_rn(resize_decode_u8(v[c]), w)); }
    }
    float *o = band + (size_t(r) * size_t(out_w) + size_t(x)) * C;
    // (the band comes from cudaMalloc and a sample is C floats: 16-byte / 8-byte aligned for C == 4 / 2)
    if(C == 4) { *reinterpret_cast<float4 *>(o) = make_float4(acc[0], acc[1 % C], acc[2 % C], acc[3 % C]); }
    else if(C == 2) { *reinterpret_cast<float2 *>(o) = make_float2(acc[0], acc[1 % C]); }
    else
    {
#pragma unroll
        for(int c = 0; c < C; ++c) { o[c] = acc[c]; }
    }
}

// Vertical pass + encode: output rows [y0, y0 + rows), one thread per (output row, column).
template<int C>
__global__ void __launch_bounds__(256) resize_v_kernel(const float *__restrict__ band, int band_row0, int out_w, int y0, int rows,
                                                        const int *__restrict__ start, const int2 *__restrict__ tap, uint8_t *__restrict__ dst)
{
    // 32 x 8 tile: neighbouring output rows read mostly the same band rows (5 of 5 taps shifted by one at 1:1), which then
    // come from L1 instead of L2
    const int x = blockIdx.x * kResizeTileW + int(threadIdx.x % kResizeTileW), ry = blockIdx.y * kResizeTileH + int(threadIdx.x / kResizeTileW);
    const int y = y0 + ry;
    if(x >= out_w || ry >= rows) { return; }
    float acc[C];
#pragma unroll
    for(int c = 0; c < C; ++c) { acc[c] = 0.0f; }
    const int t1 = __ldg(start + y + 1);
    for(int t = __ldg(start + y); t < t1; ++t)
    {
        const int2 tp = __ldg(tap + t);
        const float *p = band + (size_t(tp.x - band_row0) * size_t(out_w) + size_t(x)) * C;
        const float w = __int_as_float(tp.y);
        // one 128-bit (64-bit) load per tap: a warp's 32 samples are then 4 (2) cache lines in one request instead of
        // four strided 32-bit requests over the same lines
        float v[C];
        if(C == 4)
        {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(p));
            v[0] = q.x, v[1 % C] = q.y, v[2 % C] = q.z, v[3 % C] = q.w;
        }
        else if(C == 2)
        {
            const float2 q = __ldg(reinterpret_cast<const float2 *>(p));
            v[0] = q.x, v[1 % C] = q.y;
        }
        else
        {
#pragma unroll
            for(int c = 0; c < C; ++c) { v[c] = p[c]; }
        }
#pragma unroll
        for(int c = 0; c < C; ++c) { acc[c] = __fadd_rn(acc[c], __fmul_rn(v[c], w)); }
    }
    uint32_t q[C];
#pragma unroll
    for(int c = 0; c < C; ++c)
    {
        float f = acc[c];
        f = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);// stbir__saturate :572-581
        const float s = __fmul_rn(f, 255.0f);
        // (int)((double) s + 0.5): s is in [0, 255]; floor(s + 0.5) without a second rounding
        const int i = __float2int_rz(s);
        q[c] = uint32_t(i) + ((__fsub_rn(s, (float) i) >= 0.5f) ? 1u : 0u);
    }
    uint8_t *o = dst + (size_t(y) * size_t(out_w) + size_t(x)) * C;
    if(C == 4) { *reinterpret_cast<uint32_t *>(o) = q[0] | (q[1 % C] << 8) | (q[2 % C] << 16) | (q[3 % C] << 24); }
    else
    {
#pragma unroll
        for(int c = 0; c < C; ++c) { o[c] = uint8_t(q[c]); }
    }
}

// Fused pass for regular axes (DeviceAxis::reg_*) and RGBA: one thread owns an output column and walks a strip of output rows.
// Its S * T horizontal taps (sample, coefficient) stay in registers for the whole strip; the horizontally filtered samples of
// the last T - S input rows stay in a register window, so every input row is filtered once per strip and the vertical sum
// reads registers.  The arithmetic is that of resize_h_kernel followed by resize_v_kernel, operation for operation (same
// order, same coefficient bits, multiply and add rounded separately), so the bytes are the same -- but there is no fp32 band in
// memory and an output sample costs S * T coalesced 32-bit loads instead of ~30 (1:1) to ~180 (2:1) L1 wavefronts per warp row.
// Rows outside the image are clamped copies of the edge row, recomputed (the tap lists name them as separate terms).
template<int S, int T>
__global__ void __launch_bounds__(128) resize_fused_kernel(const uint8_t *__restrict__ src, int in_w, int in_h, int out_w, int y0, int y1, int strip,
                                                            int off_y, const int *__restrict__ xstart, const int2 *__restrict__ xtap,
                                                            const int *__restrict__ ystart, const int2 *__restrict__ ytap, uint8_t *__restrict__ dst)
{
    const int x = blockIdx.x * 128 + int(threadIdx.x);
    const int ya = y0 + int(blockIdx.y) * strip, yb = min(ya + strip, y1);
    if(x >= out_w || ya >= yb) { return; }
    int hx[T];
    float hc[T];
    {
        const int t0 = __ldg(xstart + x);
#pragma unroll
        for(int t = 0; t < T; ++t)
        {
            const int2 tp = __ldg(xtap + t0 + t);
            hx[t] = tp.x, hc[t] = __int_as_float(tp.y);
        }
    }
    // horizontally filtered sample of (virtual) input row v at this thread's column
    auto hrow = [&](int v) -> float4 {
        const int r = v < 0 ? 0 : (v >= in_h ? in_h - 1 : v);
        const uint32_t *row = reinterpret_cast<const uint32_t *>(src) + size_t(r) * size_t(in_w);
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
        for(int t = 0; t < T; ++t)
        {
            const uint32_t q = __ldg(row + hx[t]);
            a0 = __fadd_rn(a0, __fmul_rn(resize_decode_u8(q & 255u), hc[t]));
            a1 = __fadd_rn(a1, __fmul_rn(resize_decode_u8((q >> 8) & 255u), hc[t]));
            a2 = __fadd_rn(a2, __fmul_rn(resize_decode_u8((q >> 16) & 255u), hc[t]));
            a3 = __fadd_rn(a3, __fmul_rn(resize_decode_u8(q >> 24), hc[t]));
        }
        return make_float4(a0, a1, a2, a3);
    };
    float4 win[T];// win[t] = filtered virtual row S * y + off_y + t of the current output row y
#pragma unroll
    for(int t = 0; t < T - S; ++t) { win[t + S] = hrow(S * ya + off_y + t); }
    uint32_t *out = reinterpret_cast<uint32_t *>(dst) + size_t(ya) * size_t(out_w) + size_t(x);
#pragma unroll 1
    for(int y = ya; y < yb; ++y, out += out_w)
    {
#pragma unroll
        for(int t = 0; t < T - S; ++t) { win[t] = win[t + S]; }
#pragma unroll
        for(int t = T - S; t < T; ++t) { win[t] = hrow(S * y + off_y + t); }
        const int t0 = __ldg(ystart + y);
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
        for(int t = 0; t < T; ++t)
        {
            const float w = __int_as_float(__ldg(ytap + t0 + t).y);
            a0 = __fadd_rn(a0, __fmul_rn(win[t].x, w));
            a1 = __fadd_rn(a1, __fmul_rn(win[t].y, w));
            a2 = __fadd_rn(a2, __fmul_rn(win[t].z, w));
            a3 = __fadd_rn(a3, __fmul_rn(win[t].w, w));
        }
        uint32_t q = 0;
        const float acc[4] = {a0, a1, a2, a3};
#pragma unroll
        for(int c = 0; c < 4; ++c)
        {
            float f = acc[c];
            f = f < 0.0f ? 0.0f : (f > 1.0f ? 1.0f : f);// stbir__saturate :572-581
            const float sc = __fmul_rn(f, 255.0f);
            const int i = __float2int_rz(sc);// (int)((double) s + 0.5), as in resize_v_kernel
            q |= (uint32_t(i) + ((__fsub_rn(sc, (float) i) >= 0.5f) ? 1u : 0u)) << (8 * c);
        }
        *out = q;
    }
}


// ---- strip kernel, second version (round 2): the per-thread body lives in resize_strip.h (host and device).
template<int S, int T, int NC>
__global__ void __launch_bounds__(128) resize_strip_kernel(const uint8_t *__restrict__ src, int in_w, int in_h, int out_w, int y0, int y1, int strip,
                                                            const FusedCoef cx, const FusedCoef cy, uint8_t *__restrict__ dst)
{
    resize_strip_thread<S, T, NC>(src, in_w, in_h, out_w, y0, y1, strip, cx, cy, dst, int(blockIdx.x * 128 + threadIdx.x), int(blockIdx.y));
}

}// namespace vkt

// per-slot cache of axis tables (keyed by in/out size); lives beside the DeviceSlot, guarded by the slot mutex
// Entries are shared: a caller keeps the shared_ptr for as long as it (or a kernel it queued) reads the tables, so an
// eviction can never free tables that are still referenced (DeviceAxis's destructor calls cudaFree, which waits for the
// device).
using vkt_axis_ptr = std::shared_ptr<const vkt::DeviceAxis>;
struct vkt_axis_cache
{
    std::map<std::pair<int, int>, vkt_axis_ptr> axes;
};

namespace vkt
{

constexpr size_t kAxisCacheEntries = 256;

static int get_axis(vkt_bcn_ctx *ctx, DeviceSlot *s, int in, int out, vkt_axis_ptr *res)
{
    if(!s->axis_cache) { s->axis_cache = new vkt_axis_cache; }
    auto &m = s->axis_cache->axes;
    auto it = m.find({in, out});
    if(it != m.end())
    {
        *res = it->second;
        return VKT_BCN_OK;
    }
    if(m.size() >= kAxisCacheEntries)
    {
        // drop what nobody holds; entries the running call still references stay (and stay valid)
        for(auto e = m.begin(); e != m.end();) { e = (e->second.use_count() == 1) ? m.erase(e) : std::next(e); }
    }
    ResizeAxis h;
    h.build(in, out);
    auto d = std::make_shared<DeviceAxis>();
    d->in_size = in, d->out_size = out;
    const size_t n = h.idx.size();
    VKT_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&d->d_start), (size_t(out) + 1) * sizeof(int)));
    VKT_CUDA(ctx, cudaMalloc(reinterpret_cast<void **>(&d->d_tap), std::max<size_t>(n, 1) * sizeof(int2)));
    std::vector<int2> taps(n);
    for(size_t t = 0; t < n; ++t)
    {
        int bits;
        memcpy(&bits, &h.coef[t], sizeof(bits));
        taps[t] = make_int2(h.idx[t], bits);
    }
    VKT_CUDA(ctx, cudaMemcpyAsync(d->d_start, h.start.data(), (size_t(out) + 1) * sizeof(int), cudaMemcpyHostToDevice, s->stream));
    VKT_CUDA(ctx, cudaMemcpyAsync(d->d_tap, taps.data(), n * sizeof(int2), cudaMemcpyHostToDevice, s->stream));
    VKT_CUDA(ctx, cudaStreamSynchronize(s->stream));// the host vectors die with this scope
    d->first_in.resize(size_t(out)), d->last_in.resize(size_t(out));
    for(int o = 0; o < out; ++o)
    {
        int lo = in - 1, hi = 0;
        for(int t = h.start[size_t(o)]; t < h.start[size_t(o) + 1]; ++t) { lo = std::min(lo, h.idx[size_t(t)]), hi = std::max(hi, h.idx[size_t(t)]); }
        if(h.start[size_t(o)] == h.start[size_t(o) + 1]) { lo = hi = 0; }
        d->first_in[size_t(o)] = lo, d->last_in[size_t(o)] = hi;
    }
    // regular structure?  (every output: the same number of taps, samples clamp(step * o + off + t))
    if(out > 0 && (in == out || in == 2 * out))
    {
        const int step = in / out, T = h.start[1] - h.start[0];
        bool regular = (T == 3 && step == 1) || (T == 8 && step == 2);
        const int off = regular ? (step == 1 ? -1 : -3) : 0;
        for(int o = 0; o < out && regular; ++o)
        {
            regular = (h.start[size_t(o) + 1] - h.start[size_t(o)] == T);
            for(int t = 0; t < T && regular; ++t)
            {
                const int v = step * o + off + t;
                regular = h.idx[size_t(h.start[size_t(o)] + t)] == (v < 0 ? 0 : (v >= in ? in - 1 : v));
            }
        }
        if(regular)
        {
            d->reg_taps = T, d->reg_step = step, d->reg_off = off;
            bool uniform = true;
            for(int o = 1; o < out && uniform; ++o)
            {
                uniform = memcmp(&h.coef[size_t(h.start[size_t(o)])], &h.coef[size_t(h.start[0])], size_t(T) * sizeof(float)) == 0;
            }
            d->reg_uniform = uniform;
            for(int t = 0; t < T; ++t) { d->reg_coef[t] = h.coef[size_t(h.start[0] + t)]; }
        }
    }
    *res = d;
    m[{in, out}] = std::move(d);
    return VKT_BCN_OK;
}

constexpr size_t kResizeBandBytes = size_t(256) << 20;// fp32 band buffer budget per slot

// d_src (w x h x comps, tightly packed, device) -> d_dst (ow x oh x comps).  Work is queued on `stream`.
// out_y0 / out_y1 restrict the call to a range of output rows (out_y1 == 0: all rows).
static int resize_device(vkt_bcn_ctx *ctx, DeviceSlot *s, const uint8_t *d_src, uint32_t w, uint32_t h, uint32_t comps, uint8_t *d_dst,
                         uint32_t ow, uint32_t oh, cudaStream_t stream, uint32_t out_y0 = 0, uint32_t out_y1 = 0)
{
    if(out_y1 == 0) { out_y1 = oh; }
    vkt_axis_ptr ax, ay;
    int rc = get_axis(ctx, s, int(w), int(ow), &ax);
    if(rc) { return rc; }
    if((rc = get_axis(ctx, s, int(h), int(oh), &ay))) { return rc; }
    static const bool no_fused = getenv("VKT_BCN_NO_FUSED_RESIZE") != nullptr;// (tuning / A-B tests)
    // (a thread of the strip kernels walks a strip of rows in sequence: calls with very few output samples -- the smallest levels
    // of a chain -- would leave the device to a handful of long-running threads; they keep the sample-parallel general passes)
    const uint64_t outputs = uint64_t(out_y1 > out_y0 ? out_y1 - out_y0 : 0) * ow;
    static const int strip_env = getenv("VKT_BCN_RESIZE_STRIP") ? atoi(getenv("VKT_BCN_RESIZE_STRIP")) : 0;// (tuning)
    const bool regular = comps == 4 && ax->reg_taps && ax->reg_taps == ay->reg_taps && ax->reg_step == ay->reg_step && !no_fused;
    // second version of the strip kernel: uniform coefficients as parameters, NC columns per thread, 128-bit loads.  At a third
    // of the general passes' instructions it also takes the smaller calls (the row bands of levels 0 and 1, levels down to
    // 256^2), with shorter strips while a call has few threads (a strip of n output rows filters n + 2 (1:1) or 2 n + 6 (2:1)
    // input rows)
    static const bool old_strip = getenv("VKT_BCN_OLD_STRIP") != nullptr;// (A-B tests)
    static const uint64_t strip_min = getenv("VKT_BCN_STRIP_MIN") ? strtoull(getenv("VKT_BCN_STRIP_MIN"), nullptr, 10) : (1u << 16);// (tests reach the edge cases with small images)
    const uint32_t nc = (regular && ax->reg_step == 1) ? 4u : 2u;
    const bool strip2 = regular && !old_strip && outputs > 0 && outputs >= strip_min && ax->reg_uniform && ay->reg_uniform &&
                        ax->reg_off == -(ax->reg_taps - ax->reg_step) / 2 && ay->reg_off == ax->reg_off && (ow % nc) == 0 && (w % 4u) == 0 &&
                        (reinterpret_cast<uintptr_t>(d_src) & 15u) == 0 && (reinterpret_cast<uintptr_t>(d_dst) & 15u) == 0;
    if(strip2 || (regular && outputs >= (1u << 20) && (reinterpret_cast<uintptr_t>(d_src) & 3u) == 0 && (reinterpret_cast<uintptr_t>(d_dst) & 3u) == 0))
    {
        // regular axes (every level of a power-of-two chain): the fused pass, no fp32 band
        const uint32_t rows = out_y1 - out_y0;
        int strip = strip_env > 0 ? strip_env : 16;// measured on 4096^2 chains: 8 / 16 / 32 rows per strip 3.10 / 3.08 / 3.15 ms (general passes: 3.23)
        while((rows + uint32_t(strip) - 1u) / uint32_t(strip) > 65535u) { strip *= 2; }// (grid.y limit: more than a million output rows)
        if(strip2)
        {
            const int least = (ax->reg_step == 1) ? 4 : 8;
            while(strip_env <= 0 && strip > least && uint64_t(ow / nc) * ((rows + uint32_t(strip) - 1u) / uint32_t(strip)) < 16384u) { strip /= 2; }
            FusedCoef cx, cy;
            memcpy(cx.c, ax->reg_coef, sizeof(cx.c)), memcpy(cy.c, ay->reg_coef, sizeof(cy.c));
            const dim3 grid2((ow / nc + 127u) / 128u, (rows + uint32_t(strip) - 1u) / uint32_t(strip));
            if(ax->reg_step == 1)
            {
                resize_strip_kernel<1, 3, 4><<<grid2, 128, 0, stream>>>(d_src, int(w), int(h), int(ow), int(out_y0), int(out_y1), strip, cx, cy, d_dst);
            }
            else
            {
                resize_strip_kernel<2, 8, 2><<<grid2, 128, 0, stream>>>(d_src, int(w), int(h), int(ow), int(out_y0), int(out_y1), strip, cx, cy, d_dst);
            }
            VKT_CUDA(ctx, cudaGetLastError());
            count(ctx, 1, 0, 0);
            return VKT_BCN_OK;
        }
        const dim3 grid((ow + 127u) / 128u, (rows + uint32_t(strip) - 1u) / uint32_t(strip));
        if(ax->reg_step == 1)
        {
            resize_fused_kernel<1, 3><<<grid, 128, 0, stream>>>(d_src, int(w), int(h), int(ow), int(out_y0), int(out_y1), strip, ay->reg_off, ax->d_start,
                                                                 ax->d_tap, ay->d_start, ay->d_tap, d_dst);
        }
        else
        {
            resize_fused_kernel<2, 8><<<grid, 128, 0, stream>>>(d_src, int(w), int(h), int(ow), int(out_y0), int(out_y1), strip, ay->reg_off, ax->d_start,
                                                                 ax->d_tap, ay->d_start, ay->d_tap, d_dst);
        }
        VKT_CUDA(ctx, cudaGetLastError());
        count(ctx, 1, 0, 0);
        return VKT_BCN_OK;
    }
    const size_t row_bytes = size_t(ow) * comps * sizeof(float);
    // output-row bands whose input-row span fits the budget (at least one output row per band)
    uint32_t y0 = out_y0;
    while(y0 < out_y1)
    {
        const int r_lo = ay->first_in[y0];
        uint32_t y1 = y0 + 1;
        int r_hi = ay->last_in[y0];
        while(y1 < out_y1 && (y1 - y0) < 32768u && (std::max(r_hi, ay->last_in[y1]) - r_lo + 1) <= 32768 &&
              size_t(std::max(r_hi, ay->last_in[y1]) - r_lo + 1) * row_bytes <= kResizeBandBytes)
        {
            r_hi = std::max(r_hi, ay->last_in[y1]);
            ++y1;
        }
        const int rows_in = r_hi - r_lo + 1;
        if((rc = ensure(ctx, &s->d_tmp, &s->tmp_cap, size_t(rows_in) * row_bytes))) { return rc; }
        float *band = static_cast<float *>(s->d_tmp);
        const dim3 gh((ow + kResizeTileW - 1) / kResizeTileW, (uint32_t(rows_in) + kResizeTileH - 1) / kResizeTileH),
                gv((ow + kResizeTileW - 1) / kResizeTileW, (y1 - y0 + kResizeTileH - 1) / kResizeTileH);
        switch(comps)
        {
#define VKT_RESIZE_CASE(C)                                                                                                     \
    case C:                                                                                                                    \
        resize_h_kernel<C><<<gh, 256, 0, stream>>>(d_src, int(w), r_lo, rows_in, int(ow), ax->d_start, ax->d_tap, band); \
        resize_v_kernel<C><<<gv, 256, 0, stream>>>(band, r_lo, int(ow), int(y0), int(y1 - y0), ay->d_start, ay->d_tap, d_dst); \
        break;
            VKT_RESIZE_CASE(1)
            VKT_RESIZE_CASE(2)
            VKT_RESIZE_CASE(3)
            VKT_RESIZE_CASE(4)
#undef VKT_RESIZE_CASE
            default: return fail(ctx, VKT_BCN_ERR_INVALID, "comps must be 1..4 (got %u)", comps);
        }
        VKT_CUDA(ctx, cudaGetLastError());
        count(ctx, 2, 0, 0);
        y0 = y1;
    }
    return VKT_BCN_OK;
}

static int resize_host(vkt_bcn_ctx *ctx, const uint8_t *pixels, uint32_t w, uint32_t h, uint32_t comps, uint8_t *out, uint32_t ow, uint32_t oh)
{
    if(!pixels || !out) { return fail(ctx, VKT_BCN_ERR_INVALID, "null buffer"); }
    if(!w || !h || !ow || !oh || comps < 1 || comps > 4) { return fail(ctx, VKT_BCN_ERR_INVALID, "bad resize arguments"); }
    DeviceSlot *s = ctx->slots[0];
    std::lock_guard<std::mutex> g(s->mtx);
    VKT_CUDA(ctx, cudaSetDevice(s->device));
    const size_t in_bytes = size_t(w) * h * comps, out_bytes = size_t(ow) * oh * comps;
    int rc = ensure(ctx, &s->d_in, &s->in_cap, align_up(in_bytes, 256) + align_up(out_bytes, 256));
    if(rc) { return rc; }
    uint8_t *d_src = static_cast<uint8_t *>(s->d_in), *d_dst = d_src + align_up(in_bytes, 256);
    VKT_CUDA(ctx, cudaMemcpyAsync(d_src, pixels, in_bytes, cudaMemcpyHostToDevice, s->stream));
    if((rc = resize_device(ctx, s, d_src, w, h, comps, d_dst, ow, oh, s->stream))) { return rc; }
    VKT_CUDA(ctx, cudaMemcpyAsync(out, d_dst, out_bytes, cudaMemcpyDeviceToHost, s->stream));
    VKT_CUDA(ctx, cudaStreamSynchronize(s->stream));
    count(ctx, 0, in_bytes, out_bytes);
    return VKT_BCN_OK;
}

// The whole of vierkant::bcn::compress() (src/texture_block_compression.cpp:64-154).  Every device of the context
// uploads the source, runs the (cheap) resize chain itself and encodes its share of every level's block rows, so no
// device-to-device traffic is needed ("halo recompute" taken to the whole chain).
// Queue one chain on `slots` (one slot per participating device; the caller holds their mutexes).  Nothing is waited for:
// chain_wait() does that.
// `pixels` and `level_blocks[l]` may be host memory (pinned for full speed) or device memory of any device of the context:
// the copies that read / write them are issued with cudaMemcpyDefault (unified addressing), so a caller that owns a device
// buffer -- e.g. a Vulkan staging or image buffer imported with cudaImportExternalMemory -- gets its
// blocks without a trip through the host.
// One process per GPU (vkt_bcn_cuda_compress_shard_begin / _end): this context is worker `rank` of `world` workers that split
// ONE chain exactly as the devices of a multi-device context do (chain_plan.h).  phase 1 queues the worker's row slices of the
// sliced levels (and, for rank > 0, the copy of its rows of the last sliced level into `handover`, host memory every worker
// sees); phase 2, on rank 0 after the callers have met at a barrier, continues the chain from `handover` through the small
// levels.  No device-to-device traffic and no collective: the hand-over is a few hundred KB per worker.
struct ChainShard
{
    uint32_t rank = 0, world = 1;
    uint8_t *handover = nullptr;
    int phase = 1;
    cudaEvent_t handed_over = nullptr;// out (phase 1, rank > 0): fires when this worker's rows have landed in `handover`
};

// level_blocks == nullptr (vkt_bcn_cuda_compress_alloc): the destinations are not known yet.  Every level is then staged in the
// slot's pinned mirror and the pending hand-overs carry (level, offset); the caller fills in the addresses before chain_wait().
static int chain_enqueue(vkt_bcn_ctx *ctx, const std::vector<DeviceSlot *> &slots, uint32_t mode, const uint8_t *pixels, uint32_t width,
                         uint32_t height, uint32_t comps, int generate_mipmaps, const vkt_bc7_params *params, void *const *level_blocks,
                         std::vector<std::pair<cudaEvent_t, std::string>> *marks_out, ChainShard *shard = nullptr)
{
    const bool deferred = (level_blocks == nullptr) && !shard;
    if(mode != VKT_BCN_MODE_BC7 && mode != VKT_BCN_MODE_BC5) { return fail(ctx, VKT_BCN_ERR_INVALID, "unknown mode %u", mode); }
    if(!pixels || (!level_blocks && !deferred) || !width || !height) { return fail(ctx, VKT_BCN_ERR_INVALID, "null buffer or empty image"); }
    if(comps != 3 && comps != 4) { return fail(ctx, VKT_BCN_ERR_INVALID, "comps must be 3 or 4 (got %u)", comps); }
    vkt_bcn_plan plan;
    if(vkt_bcn_cuda_compress_plan(width, height, generate_mipmaps, &plan)) { return fail(ctx, VKT_BCN_ERR_INVALID, "bad size"); }
    for(uint32_t l = 0; l < plan.num_levels && !deferred; ++l)
    {
        if(!level_blocks[l]) { return fail(ctx, VKT_BCN_ERR_INVALID, "level_blocks[%u] is null", l); }
    }
    if(mode == VKT_BCN_MODE_BC7)
    {
        vkt_bc7_params def;
        vkt_bc7_params_init_inline(&def);
        Bc7KernelParams kp;
        const int rc = bc7_prepare_params(params ? params : &def, &kp);
        if(rc) { return fail(ctx, rc, rc == VKT_BCN_ERR_UNSUPPORTED ? "unsupported bc7 parameters" : "invalid bc7 parameters"); }
    }
    const uint32_t G = shard ? shard->world : uint32_t(slots.size());
    if(shard && (shard->rank >= shard->world || slots.size() != 1)) { return fail(ctx, VKT_BCN_ERR_INVALID, "bad shard (rank %u of %u)", shard->rank, shard->world); }
    const size_t src_bytes = size_t(width) * height * comps;
    // Pageable caller memory (what the C++ drop-in passes: a malloc'ed image in, std::vector blocks out) is staged through
    // pinned buffers of the slot with the copy pool (host_copy.h); pinned and device memory is used in place.
    const bool stage_in = is_pageable_host(pixels);
    bool stage_out[16] = {}, any_stage_out = false;
    for(uint32_t l = 0; l < plan.num_levels; ++l) { stage_out[l] = deferred || is_pageable_host(level_blocks[l]), any_stage_out = any_stage_out || stage_out[l]; }
    size_t lvl_off[16], lvl_total = 0, out_off[16], out_total = 0;
    for(uint32_t l = 0; l < plan.num_levels; ++l)
    {
        lvl_off[l] = lvl_total, out_off[l] = out_total;
        lvl_total += align_up(size_t(plan.level_width[l]) * plan.level_height[l] * comps, 256);
        out_total += align_up(size_t(plan.level_num_blocks[l]) * 16, 256);
    }
    int rc = VKT_BCN_OK;
    // Level 0 is pipelined in row bands over several streams per device -- upload (stream2) -> resize (stream, high
    // priority) -> encode (stream4/5/6 round robin, so the tail wave of one band overlaps the heads of the next ones) ->
    // download (stream3) -- so that PCIe traffic hides behind the encode kernels.  Bands are small at both ends: the first
    // encode starts after 1/32 of the upload, and the drain at the end of the call is that of a small launch.  The
    // remaining levels (a quarter of the work) are queued behind the last level-0 resize on the high-priority stream, so
    // they run in the middle of the band sequence, not after it.
    struct Band
    {
        uint32_t y0, y1;// pixel rows of level 0
        bool encode;    // false: halo rows, resized here only because this device's slices of the next levels read them
    };
    // graded split of the block rows [r0, r1): 1, 2, 3, 4 ... 4, 3, 2, 1 thirty-seconds when there is enough work
    auto graded = [&](uint32_t r0, uint32_t r1, std::vector<Band> &out, bool resident) {
        static const uint8_t kBig[] = {1, 3, 6, 10, 14, 18, 22, 26, 29, 31, 32};// cumulative 32nds
        static const uint8_t kQuarters[] = {8, 16, 24, 32};
        static const uint8_t kMid[] = {16, 32};
        static const uint8_t kOne[] = {32};
        // measured (profiles/r1_ss_sweep.txt): a band needs enough CTAs to be worth its launches and its block latency --
        // 4096^2 wants the graded schedule, 2048^2 four equal bands (1.09 -> 1.00 ms), 1024^2 two, 512^2 one (0.32 -> 0.28 ms)
        const uint64_t blocks = uint64_t(r1 - r0) * (plan.level_width[0] / 4);
        // (a source that already lives on the device has no upload to hide: two bands, so that the encoder starts after half of the
        // level-0 resize, instead of the graded schedule whose small first and last bands cannot fill the device)
        const bool big = !resident && blocks >= (1u << 19), quarters = !resident && blocks >= (1u << 17), mid = blocks >= (1u << 15);
        const uint8_t *f = big ? kBig : (quarters ? kQuarters : (mid ? kMid : kOne));
        size_t nf = big ? sizeof(kBig) : (quarters ? sizeof(kQuarters) : (mid ? sizeof(kMid) : sizeof(kOne)));
        // tuning experiments only: VKT_BCN_BANDS="2,8,16,24,30,32" (cumulative 32nds) replaces the schedule
        uint8_t custom[32];
        if(const char *e = getenv("VKT_BCN_BANDS"))
        {
            size_t n = 0;
            for(const char *q = e; *q && n < sizeof(custom);)
            {
                custom[n++] = uint8_t(strtoul(q, const_cast<char **>(&q), 10));
                if(*q == ',') { ++q; }
            }
            if(n && custom[n - 1] == 32) { f = custom, nf = n; }
        }
        uint32_t prev = r0;
        for(size_t k = 0; k < nf; ++k)
        {
            const uint32_t e = (k + 1 == nf) ? r1 : r0 + uint32_t(uint64_t(r1 - r0) * f[k] / 32);
            if(e > prev) { out.push_back({prev * 4, e * 4, true}), prev = e; }
        }
    };
    // Several devices: levels [0, M) are sliced by block rows, the small levels [M, L) finished by device 0 (chain_plan.h)
    const ChainSplit split = chain_split(plan.level_height, plan.num_levels, G);
    const uint32_t L = split.levels, M = split.sliced, Geff = split.devices;
    const bool tail = M < L;
    DeviceSlot *ev_slot = nullptr;// events come from the current slot's pool (created once per context, reused by every call)
    // VKT_BCN_TRACE=1: events carry timestamps and the call prints its device timeline to stderr (diagnostics only)
    static const bool trace = getenv("VKT_BCN_TRACE") != nullptr;
    std::vector<std::pair<cudaEvent_t, std::string>> marks;
    auto new_event = [&](cudaEvent_t *e, const char *what = nullptr, uint32_t k = 0) -> cudaError_t {
        DeviceSlot *s = ev_slot;
        if(s->events_used == s->event_pool.size())
        {
            cudaEvent_t fresh;
            const cudaError_t r = cudaEventCreateWithFlags(&fresh, trace ? cudaEventDefault : cudaEventDisableTiming);
            if(r != cudaSuccess) { return r; }
            s->event_pool.push_back(fresh);
        }
        *e = s->event_pool[s->events_used++];
        if(trace && what) { marks.emplace_back(*e, std::string(what) + " " + std::to_string(k)); }
        return cudaSuccess;
    };
    auto mark = [&](cudaStream_t st, const char *what, uint32_t k = 0) {
        if(!trace) { return; }
        cudaEvent_t e;
        if(new_event(&e, what, k) == cudaSuccess) { cudaEventRecord(e, st); }
    };
    std::vector<cudaEvent_t> gathered(Geff, nullptr);
    size_t gather_row_bytes = 0;
    uint8_t *h_gather = shard ? shard->handover : nullptr;
    if(tail && shard)
    {
        gather_row_bytes = size_t(plan.level_width[M - 1]) * comps;
        if(!h_gather) { return fail(ctx, VKT_BCN_ERR_INVALID, "this chain needs a hand-over buffer (vkt_bcn_cuda_compress_shard_plan)"); }
    }
    else if(tail)
    {
        gather_row_bytes = size_t(plan.level_width[M - 1]) * comps;
        const size_t need = gather_row_bytes * plan.level_height[M - 1];
        if(ctx->stage_cap < need)
        {
            if(ctx->h_stage) { cudaFreeHost(ctx->h_stage); }
            ctx->h_stage = nullptr, ctx->stage_cap = 0;
            VKT_CUDA(ctx, cudaHostAlloc(&ctx->h_stage, need, cudaHostAllocPortable));
            ctx->stage_cap = need;
        }
        h_gather = static_cast<uint8_t *>(ctx->h_stage);
    }
    // blocks at d_ptr (inside slot sl's d_out) -> bytes [off, off + bytes) of the caller's level l, queued on st
    auto fetch = [&](DeviceSlot *sl, uint32_t l, size_t off, const void *d_ptr, size_t bytes, cudaStream_t st) -> int {
        uint8_t *user = deferred ? nullptr : static_cast<uint8_t *>(level_blocks[l]) + off;
        if(stage_out[l])
        {
            uint8_t *staged = static_cast<uint8_t *>(sl->h_out) + (static_cast<const uint8_t *>(d_ptr) - static_cast<const uint8_t *>(sl->d_out));
            VKT_CUDA(ctx, cudaMemcpyAsync(staged, d_ptr, bytes, cudaMemcpyDeviceToHost, st));
            cudaEvent_t landed;
            VKT_CUDA(ctx, new_event(&landed));
            VKT_CUDA(ctx, cudaEventRecord(landed, st));
            sl->pending.push_back({landed, staged, user, bytes, l, off});// chain_wait() moves it on once the event has fired
        }
        else { VKT_CUDA(ctx, cudaMemcpyAsync(user, d_ptr, bytes, cudaMemcpyDefault, st)); }
        count(ctx, 0, 0, bytes);
        return VKT_BCN_OK;
    };
    for(uint32_t g = 0; g < Geff && !rc; ++g)
    {
        if(shard && (g != shard->rank || shard->phase != 1)) { continue; }// a process shard: this worker's slices only
        DeviceSlot *s = slots[shard ? 0 : g];
        VKT_CUDA(ctx, cudaSetDevice(s->device));
        ev_slot = s, s->events_used = 0;
        // a source that already lives on this device is read in place (no copy, no room for a copy)
        const bool src_in_place = device_pointer_on(pixels, s->device);
        s->src_in_place = src_in_place;
        if((rc = ensure(ctx, &s->d_in, &s->in_cap, (src_in_place ? 0 : align_up(src_bytes, 256)) + lvl_total))) { break; }
        if((rc = ensure(ctx, &s->d_out, &s->out_cap, out_total))) { break; }
        uint8_t *d_src = src_in_place ? const_cast<uint8_t *>(pixels) : static_cast<uint8_t *>(s->d_in);
        uint8_t *d_lvl = static_cast<uint8_t *>(s->d_in) + (src_in_place ? 0 : align_up(src_bytes, 256));
        s->pending.clear();
        if(any_stage_out && (rc = ensure_pinned(ctx, &s->h_out, &s->h_out_cap, out_total))) { break; }
        // vertical tap ranges of every sliced level: ay[l] maps rows of level l-1 (l == 0: the source) to rows of level l
        std::vector<vkt_axis_ptr> ay(M);
        for(uint32_t l = 0; l < M && !rc; ++l) { rc = get_axis(ctx, s, int(l ? plan.level_height[l - 1] : height), int(plan.level_height[l]), &ay[l]); }
        if(rc) { break; }
        // own[l]: this device's block rows of level l;  need[l]: the pixel rows of level l it has to produce (own rows
        // plus whatever its rows of level l+1 read)
        std::vector<RowTaps> taps(M);
        for(uint32_t l = 0; l < M; ++l) { taps[l] = {ay[l]->first_in.data(), ay[l]->last_in.data()}; }
        const DeviceRows dr = device_rows(split, plan.level_height, taps.data(), g);
        const std::vector<std::pair<uint32_t, uint32_t>> &own = dr.own, &need = dr.need;
        // This device's bands of level 0, in processing order: its own block rows first (graded, encoded), then the halo
        // rows above and below (resized only).  Source rows are uploaded in the same order, each band fetching just the
        // rows its resize taps reach that are not on the device yet.
        std::vector<Band> bands;
        // (two bands only when nothing crosses the link in either direction: host destinations still want their downloads hidden)
        bool dst_on_device = !any_stage_out;
        for(uint32_t l = 0; l < plan.num_levels && dst_on_device && !deferred; ++l) { dst_on_device = device_pointer_on(level_blocks[l], s->device); }
        graded(own[0].first, own[0].second, bands, src_in_place && dst_on_device && !deferred);
        if(need[0].second > own[0].second * 4) { bands.push_back({own[0].second * 4, need[0].second, false}); }
        if(need[0].first < own[0].first * 4) { bands.push_back({need[0].first, own[0].first * 4, false}); }
        const uint32_t K = uint32_t(bands.size());
        const size_t src_row = size_t(width) * comps;
        std::vector<std::pair<uint32_t, uint32_t>> have;// disjoint source-row intervals already queued for upload
        std::vector<cudaEvent_t> ready(K, nullptr);     // recorded on the upload stream once band k's source rows are queued
        std::vector<char> queued(K, 0);
        uint32_t in_base = 0;// first source row held by the pinned input staging buffer
        if(stage_in)
        {
            const std::pair<uint32_t, uint32_t> all = source_rows(taps[0], need[0].first, need[0].second, height);
            in_base = all.first;
            if((rc = ensure_pinned(ctx, &s->h_in, &s->h_in_cap, size_t(all.second - all.first) * src_row))) { break; }
        }
        mark(s->stream2, "t0");
        auto upload = [&](uint32_t k) -> int {
            if(queued[k]) { return VKT_BCN_OK; }
            queued[k] = 1;
            const std::pair<uint32_t, uint32_t> rows = source_rows(taps[0], bands[k].y0, bands[k].y1, height);
            const uint32_t lo = rows.first, hi = rows.second;
            // subtract what is already there
            std::vector<std::pair<uint32_t, uint32_t>> todo;
            if(lo < hi && !src_in_place) { todo.push_back({lo, hi}); }
            for(const auto &iv: have)
            {
                std::vector<std::pair<uint32_t, uint32_t>> next;
                for(const auto &t: todo)
                {
                    if(iv.second <= t.first || iv.first >= t.second)
                    {
                        next.push_back(t);
                        continue;
                    }
                    if(t.first < iv.first) { next.push_back({t.first, iv.first}); }
                    if(iv.second < t.second) { next.push_back({iv.second, t.second}); }
                }
                todo.swap(next);
            }
            for(const auto &t: todo)
            {
                const uint8_t *from = pixels + size_t(t.first) * src_row;
                if(stage_in)
                {
                    uint8_t *staged = static_cast<uint8_t *>(s->h_in) + size_t(t.first - in_base) * src_row;
                    copy_pool(ctx).copy(staged, from, size_t(t.second - t.first) * src_row);
                    from = staged;
                }
                VKT_CUDA(ctx, cudaMemcpyAsync(d_src + size_t(t.first) * src_row, from,
                                              size_t(t.second - t.first) * src_row, cudaMemcpyDefault, s->stream2));
                count(ctx, 0, size_t(t.second - t.first) * src_row, 0);
                have.push_back(t);
            }
            VKT_CUDA(ctx, new_event(&ready[k], "upload done", k));
            VKT_CUDA(ctx, cudaEventRecord(ready[k], s->stream2));
            return VKT_BCN_OK;
        };
        const uint32_t w0 = plan.level_width[0], h0 = plan.level_height[0];
        uint8_t *lvl0 = d_lvl + lvl_off[0];
        const size_t row_px0 = size_t(w0) * comps * 4, row_blk0 = size_t(w0 / 4) * 16;
        uint32_t lane = 0;
        // Level 1 (three quarters of the remaining levels' work) follows level 0 band by band on the same stream: once the
        // graded bands have produced level-0 rows [bands[0].y0, y), every level-1 row whose taps stay inside that range is
        // resized.  The level is then complete when level 0 is (one device) or right after the halo bands (several), and
        // its encode starts about a millisecond earlier than behind the whole chain of ever smaller resize launches.
        // l1_a .. l1_b: the rows of level 1 done this way (tap ranges ascend with the output row).
        uint32_t l1_a = 0, l1_b = 0;
        if(M > 1)
        {
            l1_a = need[1].first;
            while(l1_a < need[1].second && uint32_t(ay[1]->first_in[l1_a]) < bands[0].y0) { ++l1_a; }
            l1_b = l1_a;
        }
        for(uint32_t k = 0; k < K && !rc; ++k)
        {
            if((rc = upload(k))) { break; }
            // keep the link busy: queue the next upload before this band's kernels -- unless the source is pageable, where
            // an upload starts with host work (the copy into the pinned staging buffer): that goes behind the launches
            if(!stage_in && k + 1 < K && (rc = upload(k + 1))) { break; }
            VKT_CUDA(ctx, cudaStreamWaitEvent(s->stream, ready[k], 0));
            if((rc = resize_device(ctx, s, d_src, width, height, comps, lvl0, w0, h0, s->stream, bands[k].y0, bands[k].y1))) { break; }
            if(!bands[k].encode) { continue; }
            cudaEvent_t resized;
            VKT_CUDA(ctx, new_event(&resized, "resize done", k));
            VKT_CUDA(ctx, cudaEventRecord(resized, s->stream));
            if(M > 1)
            {
                uint32_t b = l1_b;
                while(b < need[1].second && uint32_t(ay[1]->last_in[b]) < bands[k].y1) { ++b; }
                if(b > l1_b)
                {
                    if((rc = resize_device(ctx, s, lvl0, w0, h0, comps, d_lvl + lvl_off[1], plan.level_width[1], plan.level_height[1], s->stream, l1_b, b)))
                    {
                        break;
                    }
                    l1_b = b;
                }
            }
            const uint32_t e0 = bands[k].y0 / 4, e1 = bands[k].y1 / 4;
            cudaStream_t enc = (lane % 3u == 0) ? s->stream4 : ((lane % 3u == 1) ? s->stream5 : s->stream6);
            ++lane;
            VKT_CUDA(ctx, cudaStreamWaitEvent(enc, resized, 0));
            mark(enc, "encode start", k);
            uint8_t *d_blk = static_cast<uint8_t *>(s->d_out) + out_off[0] + size_t(e0) * row_blk0;
            if(mode == VKT_BCN_MODE_BC7) { rc = launch_bc7(ctx, s, lvl0 + size_t(e0) * row_px0, w0, (e1 - e0) * 4, comps, w0 * comps, params, d_blk, enc); }
            else { rc = launch_bc5(ctx, s, lvl0 + size_t(e0) * row_px0, w0, (e1 - e0) * 4, comps, w0 * comps, d_blk, enc); }
            if(rc) { break; }
            cudaEvent_t enc_done;
            VKT_CUDA(ctx, new_event(&enc_done, "encode done", k));
            VKT_CUDA(ctx, cudaEventRecord(enc_done, enc));
            VKT_CUDA(ctx, cudaStreamWaitEvent(s->stream3, enc_done, 0));
            if((rc = fetch(s, 0, size_t(e0) * row_blk0, d_blk, size_t(e1 - e0) * row_blk0, s->stream3))) { break; }
            mark(s->stream3, "download done", k);
            if(stage_in && k + 1 < K && (rc = upload(k + 1))) { break; }
        }
        if(rc) { break; }
        // sliced mip levels: resize the needed rows, then this device's block rows of every level in one set of launches
        struct Slice
        {
            uint32_t level, r0, r1;
        };
        std::vector<Slice> slices;
        std::vector<DevImage> dev;
        for(uint32_t l = 1; l < M && !rc; ++l)
        {
            const uint32_t w = plan.level_width[l], h = plan.level_height[l];
            uint8_t *cur = d_lvl + lvl_off[l];
            const uint8_t *prev = d_lvl + lvl_off[l - 1];
            const uint32_t pw = plan.level_width[l - 1], ph = plan.level_height[l - 1];
            if(l == 1 && l1_b > l1_a)
            {
                // what the band loop left: rows that read the halo bands
                if(need[1].first < l1_a && (rc = resize_device(ctx, s, prev, pw, ph, comps, cur, w, h, s->stream, need[1].first, l1_a))) { break; }
                if(l1_b < need[1].second && (rc = resize_device(ctx, s, prev, pw, ph, comps, cur, w, h, s->stream, l1_b, need[1].second))) { break; }
            }
            else if((rc = resize_device(ctx, s, prev, pw, ph, comps, cur, w, h, s->stream, need[l].first, need[l].second))) { break; }
            const uint32_t r0 = own[l].first, r1 = own[l].second;
            const size_t row_px = size_t(w) * comps * 4, row_blk = size_t(w / 4) * 16;
            const DevImage img{cur + size_t(r0) * row_px, w, (r1 - r0) * 4, comps, w * comps, static_cast<uint8_t *>(s->d_out) + out_off[l] + size_t(r0) * row_blk};
            // Level 1 of a 2048^2 .. 4096^2 texture is encoded on its own, on an encode lane, while the chain of small levels is
            // still being resized (measured, profiles/r1_tt_sweep.txt: +3 % at both sizes).  A larger one stays with the other
            // levels on the prioritised stream: queued behind every level-0 band it would end the call with a long download
            // (8192^2: -2 %; a prioritised lane of its own measured no better, profiles/r1_uu_sweep.txt).
            static const uint64_t l1_lane_min = getenv("VKT_BCN_L1_LANE_MIN") ? strtoull(getenv("VKT_BCN_L1_LANE_MIN"), nullptr, 10) : (1u << 16);// (tuning)
            const uint64_t l1_blocks = uint64_t(r1 - r0) * (w / 4);
            static const uint64_t l1_lane_max = getenv("VKT_BCN_L1_LANE_MAX") ? strtoull(getenv("VKT_BCN_L1_LANE_MAX"), nullptr, 10) : (1u << 19);// (tuning)
            if(l == 1 && M > 2 && l1_blocks >= l1_lane_min && l1_blocks < l1_lane_max)
            {
                cudaEvent_t l1_ready, l1_done;
                VKT_CUDA(ctx, new_event(&l1_ready, "level 1 resized"));
                VKT_CUDA(ctx, cudaEventRecord(l1_ready, s->stream));
                cudaStream_t enc = (lane % 3u == 0) ? s->stream4 : ((lane % 3u == 1) ? s->stream5 : s->stream6);
                ++lane;
                VKT_CUDA(ctx, cudaStreamWaitEvent(enc, l1_ready, 0));
                if(mode == VKT_BCN_MODE_BC7) { rc = launch_bc7_batch(ctx, s, &img, 1, params, enc); }
                else { rc = launch_bc5(ctx, s, img.d_px, img.w, img.h, img.comps, img.stride, img.d_out, enc); }
                if(rc) { break; }
                VKT_CUDA(ctx, new_event(&l1_done, "level 1 encode done"));
                VKT_CUDA(ctx, cudaEventRecord(l1_done, enc));
                VKT_CUDA(ctx, cudaStreamWaitEvent(s->stream3, l1_done, 0));
                if((rc = fetch(s, 1, size_t(r0) * row_blk, img.d_out, size_t(r1 - r0) * row_blk, s->stream3))) { break; }
                mark(s->stream3, "level 1 download done");
                continue;
            }
            slices.push_back({l, r0, r1});
            dev.push_back(img);
        }
        if(rc) { break; }
        mark(s->stream, "mip resizes done");
        if(tail && g > 0)
        {
            // this device's rows of level M-1 -> the pinned gather buffer (device 0 continues the chain from there)
            const size_t off = size_t(own[M - 1].first) * 4 * gather_row_bytes, bytes = size_t(own[M - 1].second - own[M - 1].first) * 4 * gather_row_bytes;
            VKT_CUDA(ctx, cudaMemcpyAsync(h_gather + off, d_lvl + lvl_off[M - 1] + off, bytes, cudaMemcpyDeviceToHost, s->stream));
            count(ctx, 0, 0, bytes);
            VKT_CUDA(ctx, new_event(&gathered[g]));
            VKT_CUDA(ctx, cudaEventRecord(gathered[g], s->stream));
            if(shard) { shard->handed_over = gathered[g]; }
        }
        auto encode_and_fetch = [&]() -> int {
            if(!dev.empty())
            {
                if(mode == VKT_BCN_MODE_BC7)
                {
                    const int r = launch_bc7_batch(ctx, s, dev.data(), uint32_t(dev.size()), params, s->stream);
                    if(r) { return r; }
                }
                else
                {
                    for(const DevImage &d: dev)
                    {
                        const int r = launch_bc5(ctx, s, d.d_px, d.w, d.h, d.comps, d.stride, d.d_out, s->stream);
                        if(r) { return r; }
                    }
                }
            }
            mark(s->stream, "mip encode done");
            for(size_t k = 0; k < slices.size(); ++k)
            {
                const Slice &sl = slices[k];
                const size_t row_blk = size_t(plan.level_width[sl.level] / 4) * 16, bytes = size_t(sl.r1 - sl.r0) * row_blk;
                const int r = fetch(s, sl.level, size_t(sl.r0) * row_blk, dev[k].d_out, bytes, s->stream);
                if(r) { return r; }
            }
            mark(s->stream, "mip download done");
            return VKT_BCN_OK;
        };
        if((rc = encode_and_fetch())) { break; }
    }
    if(tail && !rc && (!shard || (shard->phase == 2 && shard->rank == 0)))
    {
        // small levels [M, L) on device 0: wait for every device's rows of level M-1, fetch them, continue the chain
        // (a process shard: the workers' barrier between phase 1 and phase 2 was that wait)
        DeviceSlot *s = slots[0];
        VKT_CUDA(ctx, cudaSetDevice(s->device));
        ev_slot = s;
        uint8_t *d_lvl = static_cast<uint8_t *>(s->d_in) + (s->src_in_place ? 0 : align_up(src_bytes, 256));
        for(uint32_t g = 1; g < Geff && !shard; ++g) { VKT_CUDA(ctx, cudaStreamWaitEvent(s->stream, gathered[g], 0)); }
        {
            // rows of the other devices only (device 0's own rows are in place and may still be read by its encode kernels)
            const size_t off = size_t(plan.level_height[M - 1] / 4 * 1 / Geff) * 4 * gather_row_bytes;
            const size_t total = gather_row_bytes * plan.level_height[M - 1];
            VKT_CUDA(ctx, cudaMemcpyAsync(d_lvl + lvl_off[M - 1] + off, h_gather + off, total - off, cudaMemcpyHostToDevice, s->stream));
            count(ctx, 0, total - off, 0);
        }
        std::vector<DevImage> dev;
        for(uint32_t l = M; l < L && !rc; ++l)
        {
            const uint32_t w = plan.level_width[l], h = plan.level_height[l];
            uint8_t *cur = d_lvl + lvl_off[l];
            if((rc = resize_device(ctx, s, d_lvl + lvl_off[l - 1], plan.level_width[l - 1], plan.level_height[l - 1], comps, cur, w, h, s->stream))) { break; }
            dev.push_back({cur, w, h, comps, w * comps, static_cast<uint8_t *>(s->d_out) + out_off[l]});
        }
        if(!rc && mode == VKT_BCN_MODE_BC7) { rc = launch_bc7_batch(ctx, s, dev.data(), uint32_t(dev.size()), params, s->stream); }
        else if(!rc)
        {
            for(const DevImage &d: dev)
            {
                if((rc = launch_bc5(ctx, s, d.d_px, d.w, d.h, d.comps, d.stride, d.d_out, s->stream))) { break; }
            }
        }
        for(uint32_t l = M; l < L && !rc; ++l)
        {
            const size_t bytes = size_t(plan.level_num_blocks[l]) * 16;
            rc = fetch(s, l, 0, static_cast<uint8_t *>(s->d_out) + out_off[l], bytes, s->stream);
        }
    }
    if(marks_out) { marks_out->swap(marks); }
    return rc;
}

static int chain_wait(vkt_bcn_ctx *ctx, const std::vector<DeviceSlot *> &slots)
{
    int rc = VKT_BCN_OK;
    for(DeviceSlot *s: slots)
    {
        if(cudaSetDevice(s->device) != cudaSuccess) { continue; }
        // blocks staged for a pageable destination: hand each piece over as soon as it has landed, while the GPU still works
        cudaError_t e = cudaSuccess;
        for(const DeviceSlot::PendingCopy &pc: s->pending)
        {
            const cudaError_t e1 = cudaEventSynchronize(pc.ready);
            if(e1 == cudaSuccess && pc.dst) { copy_pool(ctx).copy(pc.dst, pc.src, pc.bytes); }
            else if(e == cudaSuccess) { e = e1; }
        }
        s->pending.clear();
        {
            const cudaError_t e0 = cudaStreamSynchronize(s->stream);
            if(e == cudaSuccess) { e = e0; }
        }
        for(cudaStream_t st: {s->stream2, s->stream3, s->stream4, s->stream5, s->stream6})
        {
            const cudaError_t e2 = cudaStreamSynchronize(st);
            if(e == cudaSuccess) { e = e2; }
        }
        if(e != cudaSuccess && !rc) { rc = fail(ctx, VKT_BCN_ERR_CUDA, "stream synchronize failed: %s", cudaGetErrorString(e)); }
    }
    return rc;
}

// Lanes.  A slot (streams, buffers, tables, axis cache) carries one chain at a time.  A single-device context answers concurrent
// compress() calls from several host threads -- a loader that fans its textures out -- on different lanes of the device, up to
// kMaxLanes, so their uploads, kernels and downloads overlap like the textures of a compress_batch call; only when every lane is
// busy does a call wait.  Lock order everywhere: primary slots in index order, then slots2 in index order.
constexpr size_t kMaxLanes = 4;

// make sure slots2 holds at least `want` entries (entry k: device k % G)
static int grow_lanes(vkt_bcn_ctx *ctx, size_t want)
{
    std::lock_guard<std::mutex> g(ctx->lanes_mtx);
    const size_t G = ctx->slots.size();
    if(ctx->slots2.capacity() < 64) { ctx->slots2.reserve(64); }// entries are never moved once published
    while(ctx->slots2.size() < want && ctx->slots2.size() < 64)
    {
        auto *s = new DeviceSlot;
        s->device = ctx->slots[ctx->slots2.size() % G]->device;
        const cudaError_t e = init_slot(s, ctx->host_tables);
        if(e != cudaSuccess)
        {
            // a half-made lane must not be found by the next call
            const int dev = s->device;
            destroy_slot(s);
            return fail(ctx, VKT_BCN_ERR_CUDA, "extra lane of device %d: %s", dev, cudaGetErrorString(e));
        }
        ctx->slots2.push_back(s);
    }
    return VKT_BCN_OK;
}

static size_t lane_count(vkt_bcn_ctx *ctx)
{
    std::lock_guard<std::mutex> g(ctx->lanes_mtx);
    return ctx->slots2.size();
}

// a free lane of a single-device context, locked; *slot receives it
static int acquire_lane(vkt_bcn_ctx *ctx, DeviceSlot **slot, std::unique_lock<std::mutex> *lock)
{
    for(;;)
    {
        const size_t extra = lane_count(ctx);
        for(size_t k = 0; k <= extra; ++k)
        {
            DeviceSlot *s = k ? ctx->slots2[k - 1] : ctx->slots[0];
            std::unique_lock<std::mutex> l(s->mtx, std::try_to_lock);
            if(l.owns_lock())
            {
                *slot = s, *lock = std::move(l);
                return VKT_BCN_OK;
            }
        }
        if(extra + 1 >= kMaxLanes) { break; }
        const int rc = grow_lanes(ctx, extra + 1);
        if(rc) { return rc; }
    }
    *slot = ctx->slots[0];// every lane is busy: queue up behind the primary one
    *lock = std::unique_lock<std::mutex>(ctx->slots[0]->mtx);
    return VKT_BCN_OK;
}

// The whole of vierkant::bcn::compress() for one image on every device of the context.
static int compress_chain(vkt_bcn_ctx *ctx, uint32_t mode, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t comps,
                          int generate_mipmaps, const vkt_bc7_params *params, void *const *level_blocks)
{
    std::vector<std::unique_lock<std::mutex>> locks;
    std::vector<DeviceSlot *> use(ctx->slots);
    if(ctx->slots.size() == 1)
    {
        locks.emplace_back();
        const int rl = acquire_lane(ctx, &use[0], &locks[0]);
        if(rl) { return rl; }
    }
    else
    {
        for(DeviceSlot *s: ctx->slots) { locks.emplace_back(s->mtx); }
    }
    std::vector<std::pair<cudaEvent_t, std::string>> marks;
    const auto h0 = std::chrono::steady_clock::now();
    int rc = chain_enqueue(ctx, use, mode, pixels, width, height, comps, generate_mipmaps, params, level_blocks, &marks);
    const auto h1 = std::chrono::steady_clock::now();
    const int rw = chain_wait(ctx, use);
    if(!rc) { rc = rw; }
    if(!marks.empty() && ctx->slots.size() == 1)// VKT_BCN_TRACE=1
    {
        const auto h2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[vkt trace] host: everything queued after %.3f ms, waited until %.3f ms\n", std::chrono::duration<double, std::milli>(h1 - h0).count(),
                std::chrono::duration<double, std::milli>(h2 - h0).count());
        for(const auto &m: marks)
        {
            float ms = 0.0f;
            if(cudaEventElapsedTime(&ms, marks[0].first, m.first) == cudaSuccess) { fprintf(stderr, "[vkt trace] %8.3f ms  %s\n", ms, m.second.c_str()); }
        }
    }
    return rc;
}

// vkt_bcn_cuda_compress_alloc: the chain is queued first; the caller's allocator is asked for every level's memory while the GPU
// works (vierkant's compress_result_t::levels[l].resize() -- 22 MB of freshly mapped, zero-filled pages for a 4096^2 chain --
// otherwise sits in front of the first upload), then the pieces are handed over as their downloads land.
static int compress_chain_alloc(vkt_bcn_ctx *ctx, uint32_t mode, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t comps,
                                int generate_mipmaps, const vkt_bc7_params *params, vkt_bcn_alloc_fn alloc_level, void *user)
{
    if(!alloc_level) { return fail(ctx, VKT_BCN_ERR_INVALID, "null allocator"); }
    vkt_bcn_plan plan;
    if(vkt_bcn_cuda_compress_plan(width, height, generate_mipmaps, &plan)) { return fail(ctx, VKT_BCN_ERR_INVALID, "bad size"); }
    std::vector<std::unique_lock<std::mutex>> locks;
    std::vector<DeviceSlot *> use(ctx->slots);
    if(ctx->slots.size() == 1)
    {
        locks.emplace_back();
        const int rl = acquire_lane(ctx, &use[0], &locks[0]);
        if(rl) { return rl; }
    }
    else
    {
        for(DeviceSlot *s: ctx->slots) { locks.emplace_back(s->mtx); }
    }
    // The allocator runs on a helper thread while this one queues the chain (with a pageable source the queueing itself is
    // host work: the source rows are staged band by band): level 0 first, one call at a time.
    uint8_t *base[16] = {};
    uint32_t failed_level = ~0u;
    static const bool trace = getenv("VKT_BCN_TRACE") != nullptr;
    const auto h0 = std::chrono::steady_clock::now();
    double alloc_ms = 0.0;
    auto allocate_all = [&] {
        const auto a0 = std::chrono::steady_clock::now();
        struct Done
        {
            double *ms;
            std::chrono::steady_clock::time_point t0;
            ~Done() { *ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
        } done{&alloc_ms, a0};
        for(uint32_t l = 0; l < plan.num_levels; ++l)
        {
            base[l] = static_cast<uint8_t *>(alloc_level(user, l, size_t(plan.level_num_blocks[l]) * 16));
            if(!base[l])
            {
                failed_level = l;
                return;
            }
        }
    };
    std::thread allocator;
    bool threaded = true;
    try
    {
        allocator = std::thread(allocate_all);
    } catch(...)// (no thread to be had: allocate after queueing, on this one)
    {
        threaded = false;
    }
    int rc = chain_enqueue(ctx, use, mode, pixels, width, height, comps, generate_mipmaps, params, nullptr, nullptr);
    const auto h1 = std::chrono::steady_clock::now();
    if(threaded) { allocator.join(); }
    else { allocate_all(); }
    const auto h2 = std::chrono::steady_clock::now();
    if(!rc && failed_level != ~0u) { rc = fail(ctx, VKT_BCN_ERR_OOM, "the caller's allocator returned null for level %u", failed_level); }
    for(DeviceSlot *s: use)
    {
        for(DeviceSlot::PendingCopy &pc: s->pending) { pc.dst = (!rc && base[pc.level]) ? base[pc.level] + pc.offset : nullptr; }
    }
    const int rw = chain_wait(ctx, use);
    if(trace)
    {
        const auto h3 = std::chrono::steady_clock::now();
        auto ms = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(t - h0).count(); };
        fprintf(stderr, "[vkt trace] compress_alloc host: queued after %.3f ms, allocator took %.3f ms (joined at %.3f), done at %.3f ms\n", ms(h1), alloc_ms,
                ms(h2), ms(h3));
    }
    return rc ? rc : rw;
}

// Several textures (vkt_bcn_cuda_compress_batch): two lanes per device -- the primary slot and a second one with its own
// streams and buffers -- each taking whole chains in turn.  The host only waits for a lane when it wants to reuse it, and by
// then the other lane of that device has a complete chain queued, so the device never idles between textures.
static int compress_many(vkt_bcn_ctx *ctx, const vkt_bcn_source *sources, uint32_t n, int generate_mipmaps, const vkt_bc7_params *params)
{
    std::vector<std::unique_lock<std::mutex>> locks;
    for(DeviceSlot *s: ctx->slots) { locks.emplace_back(s->mtx); }
    const size_t G = ctx->slots.size();
    // Lanes per device: two for large textures (the next texture's upload and resizes run under the current one's kernels); a
    // 1024^2 chain is one wave of blocks and spends most of its 0.4 ms waiting for block latencies, so small textures get four
    // lanes -- four chains in flight per device keep it full (measured: profiles/r2_s_batch_lanes.txt).
    uint64_t max_px = 0;
    for(uint32_t t = 0; t < n; ++t) { max_px = std::max<uint64_t>(max_px, uint64_t(sources[t].width) * sources[t].height); }
    size_t per_device = (max_px <= (uint64_t(1) << 22)) ? 4 : 2;
    if(const char *e = getenv("VKT_BCN_BATCH_LANES")) { per_device = size_t(std::max(1, std::min(8, atoi(e)))); }// (tuning)
    {
        const int rg = grow_lanes(ctx, (per_device - 1) * G);
        if(rg) { return rg; }
    }
    for(size_t k = 0; k < (per_device - 1) * G; ++k) { locks.emplace_back(ctx->slots2[k]->mtx); }// (entries are stable once published)
    std::vector<std::vector<DeviceSlot *>> lanes;// lane k: device k % G, set k / G
    for(size_t k = 0; k < per_device * G; ++k) { lanes.push_back({(k < G) ? ctx->slots[k] : ctx->slots2[k - G]}); }
    std::vector<char> busy(lanes.size(), 0);
    int rc = VKT_BCN_OK;
    for(uint32_t t = 0; t < n && !rc; ++t)
    {
        const size_t k = t % lanes.size();
        if(busy[k]) { rc = chain_wait(ctx, lanes[k]); }
        busy[k] = 0;
        if(rc) { break; }
        const vkt_bcn_source &src = sources[t];
        rc = chain_enqueue(ctx, lanes[k], src.mode, src.pixels, src.width, src.height, src.comps, generate_mipmaps, params, src.level_blocks, nullptr);
        busy[k] = 1;
    }
    for(size_t k = 0; k < lanes.size(); ++k)
    {
        if(busy[k])
        {
            const int rw = chain_wait(ctx, lanes[k]);
            if(!rc) { rc = rw; }
        }
    }
    return rc;
}

}// namespace vkt
