// resize_strip.h -- the per-thread body of the resize strip kernels (resize_core.cuh), written so that it also compiles as plain
// host C++: tests/host_emul runs it thread by thread against the oracle.  On the device every helper below is the intrinsic it names.
#pragma once
#include <cstdint>
#include <cstring>

#include "resize_axis.h"

namespace vkt
{

// ---- strip kernel, second version (round 2): the same sums with a third of the instructions.
// What the first version spends per output sample at 2:1 is 643 instructions, 256 of them decoding (every source sample is decoded
// by the four threads whose taps reach it, each time I2F.U8 on the conversion unit + 3 FMA-pipe instructions) and 29 + 12 encoding
// four channels (F2I, I2F, compare, two adds; compare, min, select to saturate).  Here
//   * a thread owns NC adjacent output columns (4 at 1:1, 2 at 2:1): the source samples its taps share are loaded (128-bit) and
//     decoded once -- 1.5 (1:1) / 5 (2:1) samples per output column and input row instead of 3 / 8;
//   * float(v) is the bit pattern 0x4B000000 | v (= 2^23 + v, one PRMT straight from the packed pixel) minus 2^23, and
//     v / 255.0f two more FMA-pipe instructions (resize_decode_split): no conversion-unit instruction;
//   * the coefficients of a regular axis are the same for every output (DeviceAxis::reg_uniform): kernel parameters, i.e. constant
//     bank operands -- no tap tables, no registers;
//   * the register window of filtered rows is indexed modulo T at compile time (the row loop is unrolled over one period), so
//     nothing is moved between registers;
//   * saturation rides on the last vertical addition (add.sat), and the encode is resize_encode_u8: two round-toward-zero
//     additions and a byte permute, no conversion.
// Sums, operand order and roundings are those of resize_h_kernel / resize_v_kernel; the only liberty is that a sum starts with
// its first product instead of 0.0f + product, which can turn a +0 into a -0 and nothing else -- the encoded byte is 0 either way.
struct FusedCoef
{
    float c[8];
};

#if defined(__CUDA_ARCH__)
typedef float4 StripF4;
typedef uint4 StripU4;
__device__ __forceinline__ StripU4 strip_ld4(const uint32_t *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ uint32_t strip_ld1(const uint32_t *p) { return __ldg(p); }
__device__ __forceinline__ void strip_st4(uint32_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) { *reinterpret_cast<uint4 *>(p) = make_uint4(a, b, c, d); }
__device__ __forceinline__ void strip_st2(uint32_t *p, uint32_t a, uint32_t b) { *reinterpret_cast<uint2 *>(p) = make_uint2(a, b); }
__device__ __forceinline__ float strip_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float strip_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float strip_add_sat(float a, float b)
{
    float d;
    asm("add.rn.sat.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}
// resize_encode_u8 without its mask: the byte sits in the low mantissa bits
__device__ __forceinline__ uint32_t strip_encode_bits(float s) { return __float_as_uint(__fadd_rz(__fadd_rz(s, 0.5f), 8388608.0f)); }
__device__ __forceinline__ uint32_t strip_pack4(uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3)
{
    return __byte_perm(__byte_perm(e0, e1, 0x0040u), __byte_perm(e2, e3, 0x0040u), 0x5410u);
}
template<int CH>
__device__ __forceinline__ float resize_byte_as_float(uint32_t q)
{
    return __fadd_rn(__uint_as_float(__byte_perm(q, 0x4B000000u, 0x7440u + CH)), -8388608.0f);
}
#else
struct StripF4
{
    float x, y, z, w;
};
struct StripU4
{
    uint32_t x, y, z, w;
};
inline StripU4 strip_ld4(const uint32_t *p)
{
    StripU4 v;
    std::memcpy(&v, p, sizeof(v));
    return v;
}
inline uint32_t strip_ld1(const uint32_t *p) { return *p; }
inline void strip_st4(uint32_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) { p[0] = a, p[1] = b, p[2] = c, p[3] = d; }
inline void strip_st2(uint32_t *p, uint32_t a, uint32_t b) { p[0] = a, p[1] = b; }
// (host builds use -ffp-contract=off: a product and a sum round separately, as __fmul_rn / __fadd_rn do)
inline float strip_mul(float a, float b) { return a * b; }
inline float strip_add(float a, float b) { return a + b; }
inline float strip_add_sat(float a, float b)// add.rn.sat.f32: clamped to [0, 1], NaN -> +0
{
    const float d = a + b;
    return (d > 0.0f) ? (d < 1.0f ? d : 1.0f) : 0.0f;
}
inline uint32_t strip_encode_bits(float s) { return resize_encode_u8(s); }
inline uint32_t strip_pack4(uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3) { return (e0 & 255u) | ((e1 & 255u) << 8) | ((e2 & 255u) << 16) | ((e3 & 255u) << 24); }
template<int CH>
inline float resize_byte_as_float(uint32_t q)
{
    const uint32_t bits = 0x4B000000u | ((q >> (8 * CH)) & 255u);// 2^23 + v
    float m;
    std::memcpy(&m, &bits, sizeof(m));
    return m - 8388608.0f;
}
#endif

// One thread of the strip kernel: column group k (output columns NC * k ..), strip number `by` of the call's rows [y0, y1).
// Host and device: the CUDA kernel (resize_core.cuh) calls it with k / by from its thread and block indices, the host emulation
// (tests/host_emul) loops over them -- so the column-group edge cases and the window indexing are checked without a GPU.
template<int S, int T, int NC>
VKT_RESIZE_HD inline void resize_strip_thread(const uint8_t *__restrict__ src, int in_w, int in_h, int out_w, int y0, int y1, int strip,
                                              const FusedCoef &cx, const FusedCoef &cy, uint8_t *__restrict__ dst, int k, int by)
{
    static_assert((S == 1 && T == 3 && NC == 4) || (S == 2 && T == 8 && NC == 2), "1:1 Mitchell (3 taps) or 2:1 Mitchell (8 taps)");
    constexpr int OFF = -(T - S) / 2;           // first tap of output o is sample S * o + OFF: -1 / -3
    constexpr int NP = S * (NC - 1) + T;        // source samples per input row and thread: 6 / 10
    constexpr int P = (S == 1) ? T : T / S;     // output rows after which the window slots repeat: 3 / 4
    const int x = k * NC;
    const int ya = y0 + by * strip, yb = (ya + strip < y1) ? ya + strip : y1;
    if(x >= out_w || ya >= yb) { return; }
    const bool first = (k == 0), last = (x + NC >= out_w);// (out_w is a multiple of NC: checked by the caller)

    // horizontally filtered samples of (virtual) input row v at this thread's NC columns
    auto hrow = [&](int v, StripF4 (&h)[NC]) {
        const int r = v < 0 ? 0 : (v >= in_h ? in_h - 1 : v);
        const uint32_t *row = reinterpret_cast<const uint32_t *>(src) + size_t(r) * size_t(in_w);
        uint32_t px[NP];// samples S * x + OFF + i, clamped to the row
        if(S == 1)
        {
            // 4k-1 | 4k .. 4k+3 | 4k+4
            const StripU4 m = strip_ld4(row + x);
            px[0] = strip_ld1(row + (first ? 0 : x - 1));
            px[1] = m.x, px[2] = m.y, px[3] = m.z, px[4] = m.w;
            px[NP - 1] = strip_ld1(row + (last ? in_w - 1 : x + 4));
        }
        else
        {
            // 4k-3 .. 4k+6 out of the three aligned groups 4k-4.., 4k.., 4k+4.. (edge threads repeat the edge sample)
            const int g = 2 * x;// 4k
            const StripU4 b = strip_ld4(row + g);
            StripU4 a = b, c = b;
            if(!first) { a = strip_ld4(row + g - 4); }
            else { a.y = a.z = a.w = b.x; }
            if(!last) { c = strip_ld4(row + g + 4); }
            else { c.x = c.y = c.z = b.w; }
            px[0] = a.y, px[1] = a.z, px[2] = a.w, px[3] = b.x, px[4] = b.y, px[5] = b.z, px[6] = b.w;
            px[7] = c.x, px[8] = c.y, px[NP - 1] = c.z;
        }
#pragma unroll
        for(int i = 0; i < NP; ++i)
        {
            const float d0 = resize_decode_split(resize_byte_as_float<0>(px[i])), d1 = resize_decode_split(resize_byte_as_float<1>(px[i]));
            const float d2 = resize_decode_split(resize_byte_as_float<2>(px[i])), d3 = resize_decode_split(resize_byte_as_float<3>(px[i]));
#pragma unroll
            for(int j = 0; j < NC; ++j)
            {
                const int t = i - S * j;// sample i is tap t of column j (taps ascend with i: the reference's order)
                if(t == 0)
                {
                    h[j].x = strip_mul(d0, cx.c[0]), h[j].y = strip_mul(d1, cx.c[0]);
                    h[j].z = strip_mul(d2, cx.c[0]), h[j].w = strip_mul(d3, cx.c[0]);
                }
                else if(t > 0 && t < T)
                {
                    h[j].x = strip_add(h[j].x, strip_mul(d0, cx.c[t])), h[j].y = strip_add(h[j].y, strip_mul(d1, cx.c[t]));
                    h[j].z = strip_add(h[j].z, strip_mul(d2, cx.c[t])), h[j].w = strip_add(h[j].w, strip_mul(d3, cx.c[t]));
                }
            }
        }
    };

    // win[(S * m + t) % T] = filtered virtual row S * (ya + m) + OFF + t, the t-th tap of output row ya + m
    StripF4 win[T][NC];
    const int base = S * ya + OFF;
#pragma unroll
    for(int t = 0; t < T - S; ++t) { hrow(base + t, win[t]); }
    uint32_t *out = reinterpret_cast<uint32_t *>(dst) + size_t(ya) * size_t(out_w) + size_t(x);
#pragma unroll 1
    for(int m0 = 0; ya + m0 < yb; m0 += P)
    {
#pragma unroll
        for(int u = 0; u < P; ++u)
        {
            if(ya + m0 + u >= yb) { break; }
#pragma unroll
            for(int t = T - S; t < T; ++t) { hrow(base + S * (m0 + u) + t, win[(S * u + t) % T]); }
            uint32_t q[NC];
#pragma unroll
            for(int j = 0; j < NC; ++j)
            {
                const StripF4 w0 = win[(S * u) % T][j];
                float a0 = strip_mul(w0.x, cy.c[0]), a1 = strip_mul(w0.y, cy.c[0]), a2 = strip_mul(w0.z, cy.c[0]), a3 = strip_mul(w0.w, cy.c[0]);
#pragma unroll
                for(int t = 1; t < T - 1; ++t)
                {
                    const StripF4 w = win[(S * u + t) % T][j];
                    a0 = strip_add(a0, strip_mul(w.x, cy.c[t])), a1 = strip_add(a1, strip_mul(w.y, cy.c[t]));
                    a2 = strip_add(a2, strip_mul(w.z, cy.c[t])), a3 = strip_add(a3, strip_mul(w.w, cy.c[t]));
                }
                const StripF4 wl = win[(S * u + T - 1) % T][j];
                // stbir__saturate (:572-581) on the last addition, then (int)(f * 255.0f + 0.5)
                a0 = strip_add_sat(a0, strip_mul(wl.x, cy.c[T - 1])), a1 = strip_add_sat(a1, strip_mul(wl.y, cy.c[T - 1]));
                a2 = strip_add_sat(a2, strip_mul(wl.z, cy.c[T - 1])), a3 = strip_add_sat(a3, strip_mul(wl.w, cy.c[T - 1]));
                const uint32_t e0 = strip_encode_bits(strip_mul(a0, 255.0f));
                const uint32_t e1 = strip_encode_bits(strip_mul(a1, 255.0f));
                const uint32_t e2 = strip_encode_bits(strip_mul(a2, 255.0f));
                const uint32_t e3 = strip_encode_bits(strip_mul(a3, 255.0f));
                // (resize_encode_u8 four times; the bytes sit in the low mantissa bits)
                q[j] = strip_pack4(e0, e1, e2, e3);
            }
            if(NC == 4) { strip_st4(out, q[0], q[1 % NC], q[2 % NC], q[3 % NC]); }
            else { strip_st2(out, q[0], q[1 % NC]); }
            out += out_w;
        }
    }
}


}// namespace vkt
