// bc7_core.cuh -- the per-lane BC7 block search (device code; one lane owns one 4x4 block, one warp a batch of 32).
//
// B200-native re-design of bc7enc_rdo's bc7enc_compress_block (the function vierkant::bcn::compress calls for every
// block, /root/reference/src/texture_block_compression.cpp:132).  "bc7enc.cpp:N" = /root/reference/extern/bc7enc_rdo/bc7enc.cpp.
//
// Design:
//   * lane == block.  All float work (PCA, least squares, endpoint quantisation) is inherently sequential per subset
//     (SURVEY.md F12), so it runs once per lane with zero redundancy.
//   * per-lane state lives in a shared-memory column (texels + their YCbCr, hoisted once per block), conflict-free.
//   * partition estimator: the candidate list is walked warp-uniformly (every lane scores the same partition in the
//     same iteration; per subset a rolled loop over that subset's texel PAIRS, listed per partition in __constant__
//     memory) and a warp ballot skips candidates no lane of the batch still needs; for the filterbank phase the CTA
//     regroups its blocks by key partition so that the lanes of a warp need the same candidates.  Bounding boxes use
//     16x2 SIMD min/max (VIMNMX3.U16x2), palettes are built two channels per IMAD, projections are IDP.4A dot products,
//     the selected palette entry is a packed RGBA word.  Opaque blocks run the estimator before anything else.
//   * two-subset modes fit the larger subset of every lane first (warp loop trips: max(8..15) + max(1..8)).
//   * selector search: palette in registers in YCbCr, unrolled over the N entries; error and selector are fused in
//     one key (err * 16 + j) so the argmin is a single VIMNMX per candidate (first minimum wins, as the reference).
//   * every colour-cell search has ONE call site for least-squares + quantise + evaluate (a small stage machine walks
//     PCA -> least squares -> uber perturbations), which keeps the kernel inside the instruction cache.
//   * early-outs of the reference that only save work (estimator partial sums, subset-1 skip) are replaced by
//     completed sums; the argmin is identical because errors are non-negative and every comparison is strict.
//   * bit-exactness: every float op that must match goes through explicit round-to-nearest intrinsics (never
//     contracted to FMA), IEEE division / sqrt, truncating conversions, reference operation order.
//
// The same source also compiles as plain host C++ (tests/host_emul) so that the search logic can be checked against
// the reference on a machine without a GPU.  That build is test infrastructure; the product has no CPU path.
#pragma once
#include <stdint.h>

#include "bc7_tables.h"

#if defined(__CUDACC__)
#define VKT_FN __host__ __device__ __forceinline__
#define VKT_NOINLINE __host__ __device__ __noinline__
#else
#define VKT_FN inline
#define VKT_NOINLINE
#include <cmath>
#endif

namespace vkt
{

// ---------------------------------------------------------------------------------------------------- uniform tables
// Indexed only by warp-uniform values (the estimator's loop counter): __constant__ memory on the device.
#define VKT_PART2_INIT                                                                                                         \
    {0xCCCC, 0x8888, 0xEEEE, 0xECC8, 0xC880, 0xFEEC, 0xFEC8, 0xEC80, 0xC800, 0xFFEC, 0xFE80, 0xE800, 0xFFE8, 0xFF00, 0xFFF0, 0xF000, \
     0xF710, 0x008E, 0x7100, 0x08CE, 0x008C, 0x7310, 0x3100, 0x8CCE, 0x088C, 0x3110, 0x6666, 0x366C, 0x17E8, 0x0FF0, 0x718E, 0x399C, \
     0xAAAA, 0xF0F0, 0x5A5A, 0x33CC, 0x3C3C, 0x55AA, 0x9696, 0xA55A, 0x73CE, 0x13C8, 0x324C, 0x3BDC, 0x6996, 0xC33C, 0x9966, 0x0660, \
     0x0272, 0x04E4, 0x4E40, 0x2720, 0xC936, 0x936C, 0x39C6, 0x639C, 0x9336, 0x9CC6, 0x817E, 0xE718, 0xCCF0, 0x0FCC, 0x7744, 0xEE22}
#define VKT_ORDER_INIT                                                                                                         \
    {0,  13, 1,  2,  15, 14, 10, 16, 3,  23, 26, 6,  7,  21, 19, 29, 8,  4,  9,  20, 5,  31, 22, 17, 18, 11, 12, 30, 24, 25, 28, 27, \
     32, 33, 34, 45, 46, 51, 49, 50, 48, 38, 39, 37, 53, 52, 54, 36, 57, 58, 55, 41, 40, 42, 43, 59, 44, 56, 47, 35, 60, 63, 62, 61}
// Estimator work lists (the same data as Bc7Tables::est_idx / est_n0), generated at compile time: texel indices of
// subset 0 (ascending) followed by those of subset 1, and |subset 0|.
struct EstLists
{
    uint8_t idx[64][16];
    uint32_t n0[64];
};
constexpr EstLists make_est_lists()
{
    constexpr uint32_t part2[64] = VKT_PART2_INIT;
    EstLists t{};
    for(int p = 0; p < 64; ++p)
    {
        int n = 0;
        for(int sub = 0; sub < 2; ++sub)
        {
            for(int i = 0; i < 16; ++i)
            {
                if((int) ((part2[p] >> i) & 1u) == sub) { t.idx[p][n++] = (uint8_t) i; }
            }
            if(sub == 0) { t.n0[p] = (uint32_t) n; }
        }
    }
    return t;
}
// The same lists as texel PAIRS, the unit of the estimator's loops: pairs [0, n[p] & 255) walk subset 0, pairs
// [n[p] & 255, n[p] >> 8) subset 1; an odd subset ends with a pair that names its last texel twice (i0 == i1).
struct alignas(16) EstPair
{
    uint32_t i0, i1;
    uint32_t second;// all ones if i1 is a texel of its own, 0 if it repeats i0 (its error must not be counted again)
    uint32_t pad;
};
struct EstTrips
{
    EstPair t[64][10];// at most (n0 + 1) / 2 + (n1 + 1) / 2 = 9 pairs
    uint32_t n[64];
};
constexpr EstTrips make_est_trips()
{
    const EstLists l = make_est_lists();
    EstTrips t{};
    for(int p = 0; p < 64; ++p)
    {
        int m = 0;
        const int n0 = (int) l.n0[p];
        for(int k = 0; k < n0; k += 2)
        {
            t.t[p][m].i0 = l.idx[p][k], t.t[p][m].i1 = l.idx[p][(k + 1 < n0) ? k + 1 : k], t.t[p][m].second = (k + 1 < n0) ? 0xFFFFFFFFu : 0u;
            ++m;
        }
        const int m0 = m;
        for(int k = n0; k < 16; k += 2)
        {
            t.t[p][m].i0 = l.idx[p][k], t.t[p][m].i1 = l.idx[p][(k + 1 < 16) ? k + 1 : k], t.t[p][m].second = (k + 1 < 16) ? 0xFFFFFFFFu : 0u;
            ++m;
        }
        t.n[p] = (uint32_t) m0 | ((uint32_t) m << 8);
    }
    return t;
}
// Uber-level selector rescaling (bc7enc.cpp:1381-1410): nibble s of m[k][ly + 2][hy - (max - 1)] is
// clamp(floor(max * (s - ly) / (hy - ly) + .5f), 0, max) for max = 3, 7, 15 (k = 0, 1, 2), ly in [-2, 1], hy in
// [max - 1, max + 2].  Generated with integer arithmetic -- floor((2 * num + den) / (2 * den)): the float expression
// cannot land on the other side of an integer, its operands being integers below 2^8 -- and checked against the
// reference's float expression when a context is created (bc7_tables.cpp).
struct UberMaps
{
    uint64_t m[3][4][4];
};
constexpr UberMaps make_uber_maps()
{
    UberMaps t{};
    for(int k = 0; k < 3; ++k)
    {
        const int max_sel = (k == 0) ? 3 : (k == 1) ? 7 : 15;
        for(int ly = -2; ly <= 1; ++ly)
        {
            for(int hy = max_sel - 1; hy <= max_sel + 2; ++hy)
            {
                uint64_t map = 0;
                for(int sel = 0; sel <= max_sel; ++sel)
                {
                    const int num = 2 * max_sel * (sel - ly) + (hy - ly), den = 2 * (hy - ly);// den > 0
                    int v = (num >= 0) ? num / den : -((-num + den - 1) / den);          // floor
                    v = v < 0 ? 0 : (v > max_sel ? max_sel : v);
                    map |= (uint64_t) v << (4 * sel);
                }
                t.m[k][ly + 2][hy - (max_sel - 1)] = map;
            }
        }
    }
    return t;
}
#if defined(__CUDACC__)
__constant__ UberMaps c_uber = make_uber_maps();
__constant__ uint32_t c_part2[64] = VKT_PART2_INIT;// bc7enc.cpp:60-70 packed: bit i = subset of texel i
__constant__ uint32_t c_order[64] = VKT_ORDER_INIT;// bc7enc.cpp:1765-1775
__constant__ EstTrips c_trips = make_est_trips();
#endif
static const uint32_t h_part2[64] = VKT_PART2_INIT;
static const uint32_t h_order[64] = VKT_ORDER_INIT;
static const EstTrips h_trips = make_est_trips();
static const UberMaps h_uber = make_uber_maps();
#if defined(__CUDA_ARCH__)
#define VKT_UTAB(name) c_##name
#else
#define VKT_UTAB(name) h_##name
#endif

// ---------------------------------------------------------------------------------------------------- numerics
#if defined(__CUDA_ARCH__)
VKT_FN float fmul(float a, float b) { return __fmul_rn(a, b); }
VKT_FN float fadd(float a, float b) { return __fadd_rn(a, b); }
VKT_FN float fsub(float a, float b) { return __fsub_rn(a, b); }
VKT_FN float fdiv(float a, float b) { return __fdiv_rn(a, b); }
VKT_FN float fsqrt(float a) { return __fsqrt_rn(a); }
VKT_FN int f2i(float a) { return __float2int_rz(a); }// operands are always saturated first (bc7enc.cpp:871)
VKT_FN float u64_to_f(uint64_t a) { return __ull2float_rn(a); }
VKT_FN uint64_t f_to_u64(float a) { return __float2ull_rz(a); }
VKT_FN double u64_to_d(uint64_t a) { return __ull2double_rn(a); }
VKT_FN uint64_t d_to_u64(double a) { return __double2ull_rz(a); }
VKT_FN double dmul(double a, double b) { return __dmul_rn(a, b); }
VKT_FN double dadd(double a, double b) { return __dadd_rn(a, b); }
VKT_FN int popc32(uint32_t m) { return __popc(m); }
VKT_FN int ctz32(uint32_t m) { return __ffs(m) - 1; }
// true if any lane of the currently converged group of the warp holds `p`
VKT_FN bool warp_any(bool p) { return __ballot_sync(__activemask(), p) != 0u; }
// Warp-cooperative helpers: only called where the whole warp is converged (see estimate_partition).
constexpr uint32_t kWarpLanes = 32;
VKT_FN uint32_t warp_lane() { return threadIdx.x & 31u; }
VKT_FN uint32_t warp_ballot(bool p) { return __ballot_sync(0xFFFFFFFFu, p); }
VKT_FN void warp_sync() { __syncwarp(); }
VKT_FN uint64_t warp_min_u64(uint64_t v)
{
#pragma unroll
    for(int d = 16; d > 0; d >>= 1)
    {
        const uint64_t o = __shfl_xor_sync(0xFFFFFFFFu, v, d);
        v = o < v ? o : v;
    }
    return v;
}
VKT_FN uint32_t dp4a_u8(uint32_t a, uint32_t b, uint32_t c) { return __dp4a(a, b, c); }
VKT_FN uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
VKT_FN uint32_t vmin_u16x2(uint32_t a, uint32_t b) { return __vminu2(a, b); }
VKT_FN uint32_t vmax_u16x2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
#else
VKT_FN float fmul(float a, float b) { return a * b; }
VKT_FN float fadd(float a, float b) { return a + b; }
VKT_FN float fsub(float a, float b) { return a - b; }
VKT_FN float fdiv(float a, float b) { return a / b; }
VKT_FN float fsqrt(float a) { return sqrtf(a); }
VKT_FN int f2i(float a) { return (int) a; }
VKT_FN float u64_to_f(uint64_t a) { return (float) a; }
VKT_FN uint64_t f_to_u64(float a) { return (uint64_t) a; }
VKT_FN double u64_to_d(uint64_t a) { return (double) a; }
VKT_FN uint64_t d_to_u64(double a) { return (uint64_t) a; }
VKT_FN double dmul(double a, double b) { return a * b; }
VKT_FN double dadd(double a, double b) { return a + b; }
VKT_FN int popc32(uint32_t m) { return __builtin_popcount(m); }
VKT_FN int ctz32(uint32_t m) { return __builtin_ctz(m); }
VKT_FN bool warp_any(bool p) { return p; }
constexpr uint32_t kWarpLanes = 1;// host emulation: a "warp" is one lane
VKT_FN uint32_t warp_lane() { return 0; }
VKT_FN uint32_t warp_ballot(bool p) { return p ? 1u : 0u; }
VKT_FN void warp_sync() {}
VKT_FN uint64_t warp_min_u64(uint64_t v) { return v; }
VKT_FN uint32_t dp4a_u8(uint32_t a, uint32_t b, uint32_t c)
{
    return c + (a & 255) * (b & 255) + ((a >> 8) & 255) * ((b >> 8) & 255) + ((a >> 16) & 255) * ((b >> 16) & 255) + (a >> 24) * (b >> 24);
}
VKT_FN uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
    const uint64_t pool = (uint64_t) a | ((uint64_t) b << 32);
    uint32_t r = 0;
    for(int i = 0; i < 4; ++i)
    {
        const uint32_t n = (s >> (4 * i)) & 15u;
        uint32_t byte = (uint32_t) (pool >> (8 * (n & 7u))) & 255u;
        if(n & 8u) { byte = (byte & 128u) ? 255u : 0u; }
        r |= byte << (8 * i);
    }
    return r;
}
VKT_FN uint32_t vmin_u16x2(uint32_t a, uint32_t b)
{
    const uint32_t l = (a & 0xFFFFu) < (b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu), h = (a >> 16) < (b >> 16) ? (a >> 16) : (b >> 16);
    return l | (h << 16);
}
VKT_FN uint32_t vmax_u16x2(uint32_t a, uint32_t b)
{
    const uint32_t l = (a & 0xFFFFu) > (b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu), h = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16);
    return l | (h << 16);
}
#endif

// Pin a value in a register: stops ptxas from re-deriving it inside a hot loop (it otherwise trades one register for
// two extra FMA-pipe instructions per use in the selector search).
#if defined(__CUDA_ARCH__)
VKT_FN void keep(int &v) { asm volatile("" : "+r"(v)); }
// a - b computed on the ALU pipe (VIADDMNMX; operands are far below the clamp) given -b: ptxas puts plain subtractions
// of the selector search on the FMA pipe (IMAD.IADD), which is already the busier one there.
VKT_FN int sub_alu(int a, int neg_b) { return __viaddmin_s32(a, neg_b, 0x7FFFFFF0); }
#else
VKT_FN void keep(int &) {}
VKT_FN int sub_alu(int a, int neg_b) { return a + neg_b; }
#endif

VKT_FN float satf(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }// bc7enc.cpp:12-13 (NaN passes through)
VKT_FN int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
VKT_FN float sqf(float v) { return fmul(v, v); }
VKT_FN uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
VKT_FN uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }
VKT_FN int iabs(int a) { return a < 0 ? -a : a; }

// the (uint64)(err * weight + .5f) round trip, bc7enc.cpp:2167,2184,2234,2239,2326,2377,2382
VKT_FN uint64_t weigh(uint64_t err, float w) { return f_to_u64(fadd(fmul(u64_to_f(err), w), .5f)); }
// (uint64_t)((double)err * w + .5f), bc7enc.cpp:1819: the float weight and the .5f are promoted, product and sum in double
VKT_FN uint64_t weigh_d(uint64_t err, float w) { return d_to_u64(dadd(dmul(u64_to_d(err), (double) w), 0.5)); }

// byte c of a packed RGBA8 word (c is a compile-time constant at every call site): one PRMT
VKT_FN uint32_t byte_of(uint32_t v, int c) { return prmt(v, 0u, 0x4440u | (uint32_t) c); }
VKT_FN uint32_t pack4(uint32_t r, uint32_t g, uint32_t b, uint32_t a) { return r | (g << 8) | (b << 16) | (a << 24); }

// BC7 interpolation weight j of an N-entry palette: {0,21,43,64}, {0,9,...,64}, {0,4,9,...,64} (bc7enc.cpp:48-50)
VKT_FN constexpr int selw(int N, int j) { return (j * 64 + (N - 1) / 2) / (N - 1); }

constexpr uint64_t kNoErr = ~0ull;
constexpr uint64_t kIdentityPerm = 0xFEDCBA9876543210ull;

// ---------------------------------------------------------------------------------------------------- parameters
// Encoder knobs after host-side preprocessing (vkt_bc7_params -> Bc7KernelParams, bc7_params.h).
struct Bc7KernelParams
{
    uint32_t mode_mask;
    uint32_t max_partitions;
    uint32_t w[4];// final integer error weights (perceptual: {w0*4, int(w1*4*pr), int(w2*4*pb), w3*4}, bc7enc.cpp:2409-2420)
    uint32_t uber_level;
    uint32_t try_least_squares;
    uint32_t filterbank;
    uint32_t force_alpha;
    uint32_t bias_mode1_pbits;
    uint32_t key28;// 1 if every per-texel error is provably < 2^28 for these weights (packed argmin keys are then exact)
    uint32_t w16[4];// w * 16, for those keys: prepared on the host so the selector search multiplies by them directly (the
                    // compiler otherwise multiplies by w and shifts every key)
    float pbit1_weight;
    float mode1_w, mode5_w, mode6_w, mode7_w;
    // The three knobs bc7enc_rdo's RDO post-processor drives; only read by the extended kernel variant (KV == kKvExt),
    // which the host launches when one of them departs from its default.
    uint32_t ext;              // 1 if any of them is set: launch the extended variant
    uint32_t force_selectors;  // bc7enc.cpp:697
    uint32_t quant_mode6;      // bc7enc.cpp:838,885
    uint64_t forced_sel;       // m_selectors[16], one nibble per cell element
    float low_freq_weight;     // bc7enc.cpp:1819
    const uint8_t *m6_reduced; // g_mode6_reduced_quant[2048][2] (bc7enc.cpp:188-211) in device memory, [value][p]
    const uint32_t *opt7;      // Bc7Tables::opt7 where the kernel can read it (global memory on the device: the table stays out of shared memory)
};

// Kernel variant (template parameter KV of everything below):
//   kKvWide  per-texel errors may need more than 28 bits: 64-bit (error, selector) bookkeeping
//   kKvKey28 every per-texel error is provably < 2^28 (host-checked from the weights): error and selector share one
//            32-bit argmin key -- the variant every sane weight set runs
//   kKvExt   kKvWide plus forced selectors, reduced mode-6 endpoint quantisation and the low-frequency partition weight;
//            its estimator keeps the reference's work-saving early-outs, which that weight makes observable
// (bool arguments still select the first two: false -> kKvWide, true -> kKvKey28)
constexpr int kKvWide = 0, kKvKey28 = 1, kKvExt = 2;

// One texel of the block with its hoisted YCbCr (bc7enc.cpp:511-516): 16 bytes, fetched with a single 128-bit shared load.
struct alignas(16) Texel
{
    uint32_t px;// packed RGBA
    int l, cr, cb;
};

// Per-lane column in shared memory: 16 texel records, texel i (= x + 4y) at p[i * STRIDE] (STRIDE = threads per CTA;
// 1 on the host).  Consecutive lanes are 16 bytes apart, so a warp's 128-bit loads are conflict-free whatever texel
// index each lane asks for.
template<int STRIDE>
struct Lane
{
    Texel *p;
    VKT_FN uint32_t px(int i) const { return p[i * STRIDE].px; }
    VKT_FN int yl(int i) const { return p[i * STRIDE].l; }
    VKT_FN int ycr(int i) const { return p[i * STRIDE].cr; }
    VKT_FN int ycb(int i) const { return p[i * STRIDE].cb; }
    VKT_FN Texel at(int i) const { return p[i * STRIDE]; }
};

// Per-CTA exchange area behind the lane columns (device only): estimate_partition regroups the CTA's blocks by their
// filterbank key through it.
template<int STRIDE>
struct CtaScratch
{
    uint32_t err[STRIDE];  // best estimator error so far (fits 32 bits when KEY28)
    uint16_t info[STRIDE]; // best partition | running << 8
    uint16_t perm[STRIDE]; // slot -> thread whose block the slot's thread adopts
    uint32_t cnt[8][16];   // blocks per (thread index mod 8, key bin); zeroed by the kernel before its first barrier
    uint32_t cnt2[8][8];   // blocks per (thread index mod 8, subset-size class): the second regrouping (encode_block); zeroed likewise
};

// A colour cell = n texels of the block, listed by the nibbles of `perm` (texel index of cell element k at bits [4k,4k+4)).
struct CellRef
{
    uint64_t perm;
    int n;
    VKT_FN int at(int k) const { return (int) ((uint32_t) (perm >> (4 * k)) & 15u); }
};

// result of one colour-cell search (color_cell_compressor_results, bc7enc.cpp:477-485)
struct Cell
{
    uint64_t err;
    uint32_t lo, hi;// quantised endpoints, packed RGBA (without p-bits)
    uint32_t pbits; // bit 0 = pbits[0], bit 1 = pbits[1]
    uint64_t sel;   // 4 bits per cell element
};

template<int MODE>
struct ModeTraits;
template<>
struct ModeTraits<1>
{
    static constexpr int N = 8, comp_bits = 6;
    static constexpr bool pbits = true, shared = true;
};
template<>
struct ModeTraits<5>
{
    static constexpr int N = 4, comp_bits = 7;
    static constexpr bool pbits = false, shared = false;
};
template<>
struct ModeTraits<6>
{
    static constexpr int N = 16, comp_bits = 7;
    static constexpr bool pbits = true, shared = false;
};
template<>
struct ModeTraits<7>
{
    static constexpr int N = 4, comp_bits = 5;
    static constexpr bool pbits = true, shared = false;
};

// ---------------------------------------------------------------------------------------------------- colour metric
struct Ycc
{
    int l, cr, cb;
};
// bc7enc.cpp:511-516
VKT_FN Ycc to_ycc(int r, int g, int b)
{
    Ycc o;
    o.l = r * 109 + g * 366 + b * 37;
    o.cr = (r << 9) - o.l;
    o.cb = (b << 9) - o.l;
    return o;
}
// same, from a packed word: luma as two IDP.4A (366 = 183 + 183)
VKT_FN Ycc to_ycc_packed(uint32_t c)
{
    Ycc o;
    o.l = (int) dp4a_u8(c, 0x0025B76Du, dp4a_u8(c, 0x0000B700u, 0u));// (109,183,37,0) + (0,183,0,0)
    o.cr = (int) (byte_of(c, 0) << 9) - o.l;
    o.cb = (int) (byte_of(c, 2) << 9) - o.l;
    return o;
}
// perceptual distance (candidate e1, source e2), bc7enc.cpp:517-519,528: arithmetic shift, uint32 wrap-around products
VKT_FN uint32_t dist_ycc(const Ycc &e1, const Ycc &e2, const uint32_t w[4])
{
    const int dl = (e1.l - e2.l) >> 8, dcr = (e1.cr - e2.cr) >> 8, dcb = (e1.cb - e2.cb) >> 8;
    return w[0] * (uint32_t) (dl * dl) + w[1] * (uint32_t) (dcr * dcr) + w[2] * (uint32_t) (dcb * dcb);
}
// metric between two packed colours; PERC selects bc7enc.cpp:509-520 vs 521-526; ALPHA adds bc7enc.cpp:533-534.
template<bool PERC, bool ALPHA>
VKT_FN uint64_t dist_px(uint32_t cand, uint32_t src, const uint32_t w[4])
{
    uint32_t e;
    if(PERC) { e = dist_ycc(to_ycc_packed(cand), to_ycc_packed(src), w); }
    else
    {
        const int dr = (int) byte_of(cand, 0) - (int) byte_of(src, 0), dg = (int) byte_of(cand, 1) - (int) byte_of(src, 1),
                  db = (int) byte_of(cand, 2) - (int) byte_of(src, 2);
        e = w[0] * (uint32_t) (dr * dr) + w[1] * (uint32_t) (dg * dg) + w[2] * (uint32_t) (db * db);
    }
    uint64_t t = e;
    if(ALPHA)
    {
        const int da = (int) byte_of(cand, 3) - (int) byte_of(src, 3);
        t += (uint64_t) (w[3] * (uint32_t) (da * da));
    }
    return t;
}

// endpoint replication to 8 bits (scale_color, bc7enc.cpp:487-503) on a packed word, all four bytes at once.
// Every byte of q is < 2^NBITS, so the shifted copies never cross byte boundaries.
template<int NBITS>
VKT_FN uint32_t expand_packed(uint32_t q)
{
    if(NBITS == 8) { return q; }
    // per byte b: (b << (8-n)) | (b >> (2n-8)); the right-shifted copy is masked so neighbours cannot leak in
    return (q << (8 - NBITS)) | ((q >> (2 * NBITS - 8)) & (0x01010101u * ((1u << (8 - NBITS)) - 1u)));
}

// ---------------------------------------------------------------------------------------------------- single colour
// pack_mode1_to_one_color (bc7enc.cpp:537-585) / pack_mode7_to_one_color (bc7enc.cpp:587-643)
template<int MODE, bool PERC, int STRIDE>
VKT_FN uint64_t solid_cell(const Bc7Tables &T, const Bc7KernelParams &P, Lane<STRIDE> L, CellRef cell, uint32_t color, Cell &out)
{
    const uint32_t r = byte_of(color, 0), g = byte_of(color, 1), b = byte_of(color, 2), a = byte_of(color, 3);
    uint32_t best_err = 0xFFFFFFFFu, best_p = 0;
    uint32_t c = 0;
    if(MODE == 1)
    {
#pragma unroll
        for(uint32_t p = 0; p < 2; ++p)
        {
            const uint32_t err = (T.opt1[r][p] & 0xFFFF) + (T.opt1[g][p] & 0xFFFF) + (T.opt1[b][p] & 0xFFFF);
            if(err < best_err) { best_err = err, best_p = p; }// the reference's break on err == 0 cannot change the argmin
        }
        const uint32_t er = T.opt1[r][best_p], eg = T.opt1[g][best_p], eb = T.opt1[b][best_p];
        out.lo = pack4((er >> 16) & 255, (eg >> 16) & 255, (eb >> 16) & 255, 0);
        out.hi = pack4(er >> 24, eg >> 24, eb >> 24, 0);
        out.pbits = best_p;// pbits[1] = 0
        const uint32_t low = expand_packed<7>(((out.lo << 1) & 0x00FEFEFEu) | (best_p * 0x00010101u));
        const uint32_t high = expand_packed<7>(((out.hi << 1) & 0x00FEFEFEu) | (best_p * 0x00010101u));
#pragma unroll
        for(int i = 0; i < 3; ++i) { c |= (((byte_of(low, i) * (64 - 18) + byte_of(high, i) * 18 + 32) >> 6) & 255u) << (8 * i); }
        c |= 255u << 24;
    }
    else
    {
#pragma unroll
        for(uint32_t p = 0; p < 4; ++p)
        {
            const uint32_t err = (P.opt7[r * 4 + p] & 0xFFFF) + (P.opt7[g * 4 + p] & 0xFFFF) + (P.opt7[b * 4 + p] & 0xFFFF) + (P.opt7[a * 4 + p] & 0xFFFF);
            if(err < best_err) { best_err = err, best_p = p; }
        }
        const uint32_t hp = best_p >> 1, lp = best_p & 1;
        const uint32_t er = P.opt7[r * 4 + best_p], eg = P.opt7[g * 4 + best_p], eb = P.opt7[b * 4 + best_p], ea = P.opt7[a * 4 + best_p];
        out.lo = pack4((er >> 16) & 255, (eg >> 16) & 255, (eb >> 16) & 255, (ea >> 16) & 255);
        out.hi = pack4(er >> 24, eg >> 24, eb >> 24, ea >> 24);
        out.pbits = lp | (hp << 1);
        // NB bc7enc.cpp:627-631 shifts the 6-bit value left by 2 and ORs in (value >> 6) == 0: no bit replication here,
        // unlike the table construction (:256-262).  Reproduced as is.
        const uint32_t low = (((out.lo << 1) & 0xFEFEFEFEu) | (lp * 0x01010101u)) << 2;
        const uint32_t high = (((out.hi << 1) & 0xFEFEFEFEu) | (hp * 0x01010101u)) << 2;
#pragma unroll
        for(int i = 0; i < 4; ++i) { c |= (((byte_of(low, i) * (64 - 21) + byte_of(high, i) * 21 + 32) >> 6) & 255u) << (8 * i); }
    }
    // selectors: 2 (mode 1, BC7ENC_MODE_1_OPTIMAL_INDEX) or 1 (mode 7) for every element of the cell
    const uint64_t nib = (MODE == 1) ? 0x2222222222222222ull : 0x1111111111111111ull;
    out.sel = (cell.n >= 16) ? nib : (nib & ((1ull << (4 * cell.n)) - 1ull));
    uint64_t total = 0;
    for(int k = 0; k < cell.n; ++k) { total += dist_px<PERC, MODE == 7>(c, L.px(cell.at(k)), P.w); }
    out.err = total;
    return total;
}

// ---------------------------------------------------------------------------------------------------- evaluate_solution
// bc7enc.cpp:645-831.  lo/hi are quantised endpoints (no p-bits), pbits bit0/bit1.  Updates `best` on strict improvement.
// ROOMY: the calling kernel runs at two CTAs per SM (alpha blocks, uber levels) and has registers to spare.
template<int MODE, bool ALPHA, bool PERC, int KV, int STRIDE, bool ROOMY = false>
VKT_FN void evaluate(const Bc7KernelParams &P, Lane<STRIDE> L, CellRef cell, uint32_t lo, uint32_t hi, uint32_t pbits, Cell &best)
{
    typedef ModeTraits<MODE> M;
    constexpr int N = M::N;
    constexpr int NBITS = M::comp_bits + (M::pbits ? 1 : 0);
    constexpr bool KEY28 = (KV == kKvKey28);
    uint32_t qlo = lo, qhi = hi;
    if(M::pbits)
    {
        const uint32_t pl = pbits & 1u, ph = M::shared ? (pbits & 1u) : ((pbits >> 1) & 1u);
        qlo = ((lo << 1) & 0xFEFEFEFEu) | (pl * 0x01010101u);
        qhi = ((hi << 1) & 0xFEFEFEFEu) | (ph * 0x01010101u);
    }
    const uint32_t c0 = expand_packed<NBITS>(qlo), c1 = expand_packed<NBITS>(qhi);
    // interpolated colours, two channels per multiply: v = c0*(64-w) + c1*w + 32 in 16-bit lanes (<= 16352, no carries)
    const uint32_t c0_rb = c0 & 0x00FF00FFu, c1_rb = c1 & 0x00FF00FFu, c0_ga = (c0 >> 8) & 0x00FF00FFu, c1_ga = (c1 >> 8) & 0x00FF00FFu;
    uint32_t pal[N];// packed RGBA palette entries (alpha lane is garbage-free: computed like the others)
#pragma unroll
    for(int j = 0; j < N; ++j)
    {
        if(j == 0) { pal[j] = c0; }
        else if(j == N - 1) { pal[j] = c1; }
        else
        {
            const uint32_t w = (uint32_t) selw(N, j);
            const uint32_t rb = ((c0_rb * (64u - w) + c1_rb * w + 0x00200020u) >> 6) & 0x00FF00FFu;
            const uint32_t ga = ((c0_ga * (64u - w) + c1_ga * w + 0x00200020u) >> 6) & 0x00FF00FFu;
            pal[j] = rb | (ga << 8);
        }
    }

    uint64_t total = 0, sel = 0;
    if((KV == kKvExt) && P.force_selectors)
    {
        // bc7enc.cpp:697-712: element k of the cell takes m_selectors[k] (the cell's own numbering, also in the two-subset modes)
        for(int k = 0; k < cell.n; ++k)
        {
            const uint32_t s = (uint32_t) (P.forced_sel >> (4 * k)) & 15u;// < N (host-checked against the enabled modes)
            uint32_t c = pal[0];
#pragma unroll
            for(int j = 1; j < N; ++j) { c = (s == (uint32_t) j) ? pal[j] : c; }
            total += dist_px<PERC, ALPHA>(c, L.px(cell.at(k)), P.w);
            sel |= (uint64_t) s << (4 * k);
        }
    }
    else if(PERC)
    {
        int pl[N], pcr[N], pcb[N], pa[N];
#pragma unroll
        for(int j = 0; j < N; ++j)
        {
            const Ycc y = to_ycc_packed(pal[j]);
            pl[j] = y.l, pcr[j] = y.cr, pcb[j] = y.cb;
            keep(pcr[j]), keep(pcb[j]);
            pa[j] = ALPHA ? (int) byte_of(pal[j], 3) : 0;
        }
        if(KEY28)
        {
            // error < 2^28 for every texel (checked on the host from the weights): key = err * 16 + j, one min per candidate.
            const uint32_t w0 = P.w16[0], w1 = P.w16[1], w2 = P.w16[2], w3 = P.w16[3];
#if defined(VKT_EVAL_PAIR)
            // two texels per trip (independent key chains) where the register budget allows it: the kernels that run at two
            // CTAs per SM (alpha blocks, uber levels) have few warps to hide the dependent-issue latency of one chain
            constexpr bool PAIR = ROOMY && (N <= VKT_EVAL_PAIR);
#else
            constexpr bool PAIR = false;
#endif
            if(PAIR)
            {
                for(int k = 0; k < cell.n; k += 2)
                {
                    const bool two = (k + 1 < cell.n);
                    const Texel t0 = L.at(cell.at(k)), t1 = L.at(cell.at(two ? k + 1 : k));
                    const int l20 = t0.l, ncr20 = -t0.cr, ncb20 = -t0.cb, l21 = t1.l, ncr21 = -t1.cr, ncb21 = -t1.cb;
                    const int a20 = ALPHA ? (int) (t0.px >> 24) : 0, a21 = ALPHA ? (int) (t1.px >> 24) : 0;
                    uint32_t key0 = 0xFFFFFFFFu, key1 = 0xFFFFFFFFu;
#pragma unroll
                    for(int j = 0; j < N; ++j)
                    {
                        const int dl0 = (pl[j] - l20) >> 8, dcr0 = sub_alu(pcr[j], ncr20) >> 8, dcb0 = sub_alu(pcb[j], ncb20) >> 8;
                        const int dl1 = (pl[j] - l21) >> 8, dcr1 = sub_alu(pcr[j], ncr21) >> 8, dcb1 = sub_alu(pcb[j], ncb21) >> 8;
                        uint32_t e0 = w0 * (uint32_t) (dl0 * dl0) + (uint32_t) j, e1 = w0 * (uint32_t) (dl1 * dl1) + (uint32_t) j;
                        e0 += w1 * (uint32_t) (dcr0 * dcr0), e1 += w1 * (uint32_t) (dcr1 * dcr1);
                        e0 += w2 * (uint32_t) (dcb0 * dcb0), e1 += w2 * (uint32_t) (dcb1 * dcb1);
                        if(ALPHA)
                        {
                            const int da0 = pa[j] - a20, da1 = pa[j] - a21;
                            e0 += w3 * (uint32_t) (da0 * da0), e1 += w3 * (uint32_t) (da1 * da1);
                        }
                        key0 = umin(key0, e0), key1 = umin(key1, e1);
                    }
                    total += (uint64_t) (key0 >> 4);
                    sel |= (uint64_t) (key0 & 15u) << (4 * k);
                    if(two)
                    {
                        total += (uint64_t) (key1 >> 4);
                        sel |= (uint64_t) (key1 & 15u) << (4 * (k + 1));
                    }
                }
            }
            else
            {
            // The opaque uber kernels (two CTAs per SM, registers to spare) request the next texel's record before the current one
            // is searched: uber 4 2.08 -> 2.05 ms per 2048^2 level.  The default opaque kernel gains nothing once its blocks are
            // regrouped in candidate order (80-register budget), the alpha kernels lose 0.5 % (profiles/r2_ee_binperm_exp.txt).
#if defined(VKT_EVAL_NO_PREFETCH)
            constexpr bool PREFETCH = false;
#else
            constexpr bool PREFETCH = ROOMY && !ALPHA && (MODE != 5);
#endif
            Texel tn = L.at(cell.at(0));
            for(int k = 0; k < cell.n; ++k)
            {
                const Texel t = PREFETCH ? tn : L.at(cell.at(k));
                if(PREFETCH) { tn = L.at(cell.at((k + 1 < cell.n) ? k + 1 : k)); }
                const int l2 = t.l, ncr2 = -t.cr, ncb2 = -t.cb;
                const int a2 = ALPHA ? (int) (t.px >> 24) : 0;
                uint32_t key = 0xFFFFFFFFu;
#pragma unroll
                for(int j = 0; j < N; ++j)
                {
                    const int dl = (pl[j] - l2) >> 8, dcr = sub_alu(pcr[j], ncr2) >> 8, dcb = sub_alu(pcb[j], ncb2) >> 8;
                    uint32_t e = w0 * (uint32_t) (dl * dl) + (uint32_t) j;
                    e += w1 * (uint32_t) (dcr * dcr);
                    e += w2 * (uint32_t) (dcb * dcb);
                    if(ALPHA)
                    {
                        const int da = pa[j] - a2;
                        e += w3 * (uint32_t) (da * da);
                    }
                    key = umin(key, e);
                }
                total += (uint64_t) (key >> 4);
                sel |= (uint64_t) (key & 15u) << (4 * k);
            }
            }
        }
        else
        {
            for(int k = 0; k < cell.n; ++k)
            {
                const int i = cell.at(k);
                Ycc y2;
                y2.l = L.yl(i), y2.cr = L.ycr(i), y2.cb = L.ycb(i);
                const int a2 = ALPHA ? (int) (L.px(i) >> 24) : 0;
                uint64_t be = kNoErr;
                uint32_t bs = 0;
#pragma unroll
                for(int j = 0; j < N; ++j)
                {
                    Ycc y1;
                    y1.l = pl[j], y1.cr = pcr[j], y1.cb = pcb[j];
                    uint64_t e = dist_ycc(y1, y2, P.w);
                    if(ALPHA)
                    {
                        const int da = pa[j] - a2;
                        e += (uint64_t) (P.w[3] * (uint32_t) (da * da));
                    }
                    if(e < be) { be = e, bs = (uint32_t) j; }
                }
                total += be;
                sel |= (uint64_t) bs << (4 * k);
            }
        }
    }
    else
    {
        // linear metric: project onto the endpoint axis, test the two nearest palette entries (bc7enc.cpp:714-777)
        const int lr = byte_of(c0, 0), lg = byte_of(c0, 1), lb = byte_of(c0, 2), la = ALPHA ? (int) byte_of(c0, 3) : 0;
        const int dr = (int) byte_of(c1, 0) - lr, dg = (int) byte_of(c1, 1) - lg, db = (int) byte_of(c1, 2) - lb;
        const int da = ALPHA ? (int) byte_of(c1, 3) - la : 0;
        const int sq = ALPHA ? (dr * dr + dg * dg + db * db + da * da) : (dr * dr + dg * dg + db * db);
        const float f = fdiv((float) N, fadd((float) sq, .00000125f));
        for(int k = 0; k < cell.n; ++k)
        {
            const uint32_t s = L.px(cell.at(k));
            const int r = byte_of(s, 0), g = byte_of(s, 1), b = byte_of(s, 2), a = byte_of(s, 3);
            int dot = (r - lr) * dr + (g - lg) * dg + (b - lb) * db;
            if(ALPHA) { dot += (a - la) * da; }
            int si = f2i(fadd(fmul((float) dot, f), .5f));
            si = clampi(si, 1, N - 1);
            uint32_t p0 = pal[0], p1 = pal[1];
#pragma unroll
            for(int j = 1; j < N; ++j)
            {
                if(si == j) { p0 = pal[j - 1], p1 = pal[j]; }
            }
            const uint64_t e0 = dist_px<false, ALPHA>(p0, s, P.w), e1 = dist_px<false, ALPHA>(p1, s, P.w);
            uint64_t be = e1;// ties keep the upper entry in both variants (bc7enc.cpp:737 and :766)
            if(ALPHA)
            {
                if(e1 > e0) { be = e0, --si; }
            }
            else
            {
                if(e0 < be) { be = e0, --si; }
            }
            total += be;
            sel |= (uint64_t) (uint32_t) si << (4 * k);
        }
    }

    if(total < best.err)
    {
        best.err = total;
        best.lo = lo;
        best.hi = hi;
        best.pbits = pbits;
        best.sel = sel;
    }
}

// ---------------------------------------------------------------------------------------------------- find_optimal_solution
// bc7enc.cpp:868-1099 (+ fixDegenerateEndpoints :833-866).  xl/xh are float endpoints in [0,1] (saturated here).
template<int MODE, bool ALPHA, bool PERC, int KV, int STRIDE, bool ROOMY = false>
VKT_FN uint64_t fit(const Bc7Tables &T, const Bc7KernelParams &P, Lane<STRIDE> L, CellRef cell, const float xl_in[4], const float xh_in[4],
                    Cell &best)
{
    typedef ModeTraits<MODE> M;
    float xl[4], xh[4];
#pragma unroll
    for(int c = 0; c < 4; ++c) { xl[c] = satf(xl_in[c]), xh[c] = satf(xh_in[c]); }
    constexpr int NCOMP = ALPHA ? 4 : 3;
    uint32_t blo = 0, bhi = 0, bpb = 0;

    if(M::pbits)
    {
        constexpr int iscalep = (1 << (M::comp_bits + 1)) - 1;
        constexpr float scalep = (float) iscalep;
        if(!M::shared)
        {
            // independent p-bits (modes 6, 7), bc7enc.cpp:901-973
            float best0 = 1e+9f, best1 = 1e+9f;
            const bool reduced = (KV == kKvExt) && (M::comp_bits == 7) && P.quant_mode6;
            if(reduced)
            {
                // bc7enc.cpp:885-899: 64 endpoint levels from a table, low endpoint with p = 0, high endpoint with p = 1
                bpb = 2u;
#pragma unroll
                for(int c = 0; c < 4; ++c)
                {
                    blo |= (uint32_t) P.m6_reduced[f2i(fadd(fmul(xl[c], 2047.0f), .5f)) * 2 + 0] << (8 * c);
                    bhi |= (uint32_t) P.m6_reduced[f2i(fadd(fmul(xh[c], 2047.0f), .5f)) * 2 + 1] << (8 * c);
                }
            }
#pragma unroll
            for(int p = 0; p < 2; ++p)
            {
                if(reduced) { break; }
                uint32_t qlo = 0, qhi = 0;
                float e0 = 0.0f, e1 = 0.0f;
#pragma unroll
                for(int c = 0; c < 4; ++c)
                {
                    int ql, qh;
                    if(M::comp_bits == 5)
                    {
                        int vl = f2i(fmul(xl[c], 31.0f));
                        vl += (xl[c] > T.mid7[vl][p]) ? 1 : 0;
                        ql = clampi(vl * 2 + p, p, 63 - 1 + p);
                        int vh = f2i(fmul(xh[c], 31.0f));
                        vh += (xh[c] > T.mid7[vh][p]) ? 1 : 0;
                        qh = clampi(vh * 2 + p, p, 63 - 1 + p);
                    }
                    else
                    {
                        // (x * scalep - p) / 2.0f: halving is exact, so the multiply by .5f is the same correctly rounded value
                        ql = clampi(f2i(fadd(fmul(fsub(fmul(xl[c], scalep), (float) p), .5f), .5f)) * 2 + p, p, iscalep - 1 + p);
                        qh = clampi(f2i(fadd(fmul(fsub(fmul(xh[c], scalep), (float) p), .5f), .5f)) * 2 + p, p, iscalep - 1 + p);
                    }
                    qlo |= (uint32_t) ql << (8 * c);
                    qhi |= (uint32_t) qh << (8 * c);
                }
                const uint32_t slo = expand_packed<M::comp_bits + 1>(qlo), shi = expand_packed<M::comp_bits + 1>(qhi);
#pragma unroll
                for(int c = 0; c < NCOMP; ++c)
                {
                    e0 = fadd(e0, sqf(fsub((float) byte_of(slo, c), fmul(xl[c], 255.0f))));
                    e1 = fadd(e1, sqf(fsub((float) byte_of(shi, c), fmul(xh[c], 255.0f))));
                }
                if(p == 1)
                {
                    e0 = fmul(e0, P.pbit1_weight);
                    e1 = fmul(e1, P.pbit1_weight);
                }
                if(e0 < best0)
                {
                    best0 = e0;
                    bpb = (bpb & ~1u) | (uint32_t) p;
                    blo = (qlo >> 1) & 0x7F7F7F7Fu;
                }
                if(e1 < best1)
                {
                    best1 = e1;
                    bpb = (bpb & ~2u) | ((uint32_t) p << 1);
                    bhi = (qhi >> 1) & 0x7F7F7F7Fu;
                }
            }
        }
        else if(P.bias_mode1_pbits)
        {
            // bc7enc.cpp:977-1006
            float x = 0.0f;
#pragma unroll
            for(int c = 0; c < 3; ++c)
            {
                const float t = x < xl[c] ? xl[c] : x;// std::max(a, b) = (a < b) ? b : a
                x = t < xh[c] ? xh[c] : t;
            }
            const int p = (x > fdiv(253.0f, 255.0f)) ? 1 : 0;
            uint32_t qlo = 0, qhi = 0;
#pragma unroll
            for(int c = 0; c < 4; ++c)
            {
                int vl = f2i(fmul(xl[c], 63.0f));
                vl += (xl[c] > T.mid1[vl][p]) ? 1 : 0;
                int vh = f2i(fmul(xh[c], 63.0f));
                vh += (xh[c] > T.mid1[vh][p]) ? 1 : 0;
                qlo |= (uint32_t) clampi(vl * 2 + p, p, 127 - 1 + p) << (8 * c);
                qhi |= (uint32_t) clampi(vh * 2 + p, p, 127 - 1 + p) << (8 * c);
            }
            bpb = (uint32_t) p * 3u;
            blo = (qlo >> 1) & 0x7F7F7F7Fu;
            bhi = (qhi >> 1) & 0x7F7F7F7Fu;
        }
        else
        {
            // shared p-bit (mode 1), bc7enc.cpp:1009-1058
            float beste = 1e+9f;
#pragma unroll
            for(int p = 0; p < 2; ++p)
            {
                uint32_t qlo = 0, qhi = 0;
#pragma unroll
                for(int c = 0; c < 4; ++c)
                {
                    int vl = f2i(fmul(xl[c], 63.0f));
                    vl += (xl[c] > T.mid1[vl][p]) ? 1 : 0;
                    int vh = f2i(fmul(xh[c], 63.0f));
                    vh += (xh[c] > T.mid1[vh][p]) ? 1 : 0;
                    qlo |= (uint32_t) clampi(vl * 2 + p, p, 127 - 1 + p) << (8 * c);
                    qhi |= (uint32_t) clampi(vh * 2 + p, p, 127 - 1 + p) << (8 * c);
                }
                const uint32_t slo = expand_packed<7>(qlo), shi = expand_packed<7>(qhi);
                float e = 0.0f;
#pragma unroll
                for(int c = 0; c < NCOMP; ++c)
                {
                    e = fadd(e, fadd(sqf(fsub(T.unit8[byte_of(slo, c)], xl[c])), sqf(fsub(T.unit8[byte_of(shi, c)], xh[c]))));// s / 255.0f
                }
                if(p == 1) { e = fmul(e, P.pbit1_weight); }
                if(e < beste)
                {
                    beste = e;
                    bpb = (uint32_t) p * 3u;
                    blo = (qlo >> 1) & 0x7F7F7F7Fu;
                    bhi = (qhi >> 1) & 0x7F7F7F7Fu;
                }
            }
        }

        if((MODE == 1) || ((KV == kKvExt) && (MODE == 6) && P.quant_mode6))
        {
            // fixDegenerateEndpoints, bc7enc.cpp:833-866, iscale = iscalep >> 1 = 63 (mode 1) / 127 (mode 6, reduced quantisation)
            constexpr uint32_t iscale = (uint32_t) (iscalep >> 1);
#pragma unroll
            for(int c = 0; c < 3; ++c)
            {
                uint32_t l = byte_of(blo, c), h = byte_of(bhi, c);
                if(l == h && (fabsf(fsub(xl[c], xh[c])) > 0.0f))
                {
                    if(l > (iscale >> 1))
                    {
                        if(l > 0) { l--; }
                        else if(h < iscale) { h++; }
                    }
                    else
                    {
                        if(h < iscale) { h++; }
                        else if(l > 0) { l--; }
                    }
                    blo = (blo & ~(255u << (8 * c))) | (l << (8 * c));
                    bhi = (bhi & ~(255u << (8 * c))) | (h << (8 * c));
                }
            }
        }
        if((best.err == kNoErr) || (blo != best.lo) || (bhi != best.hi) || ((bpb & 3u) != (best.pbits & 3u)))
        {
            evaluate<MODE, ALPHA, PERC, KV, STRIDE, ROOMY>(P, L, cell, blo, bhi, bpb, best);
        }
    }
    else
    {
        // no p-bits (mode 5 colour, 7 bits), bc7enc.cpp:1067-1096
#pragma unroll
        for(int c = 0; c < 4; ++c)
        {
            int vl = f2i(fmul(xl[c], 127.0f));
            vl += (xl[c] > T.mid5[vl]) ? 1 : 0;
            int vh = f2i(fmul(xh[c], 127.0f));
            vh += (xh[c] > T.mid5[vh]) ? 1 : 0;
            blo |= (uint32_t) clampi(vl, 0, 127) << (8 * c);
            bhi |= (uint32_t) clampi(vh, 0, 127) << (8 * c);
        }
        if((best.err == kNoErr) || (blo != best.lo) || (bhi != best.hi))
        {
            evaluate<MODE, ALPHA, PERC, KV, STRIDE, ROOMY>(P, L, cell, blo, bhi, best.pbits, best);
        }
    }
    return best.err;
}

// ---------------------------------------------------------------------------------------------------- least squares
// compute_least_squares_endpoints_rgb/rgba, bc7enc.cpp:287-408, followed by the 1/255 scaling of bc7enc.cpp:1293-1294.
// Selector k of the cell is nibble k of `sel`.  Accumulation is sequential over the cell's texels.
template<int MODE, bool ALPHA, int STRIDE>
VKT_FN void least_squares(const Bc7Tables &T, Lane<STRIDE> L, CellRef cell, uint64_t sel, float xl[4], float xh[4])
{
    constexpr int N = ModeTraits<MODE>::N;
    constexpr int NC = ALPHA ? 4 : 3;
    const float(*wx)[4] = (N == 4) ? T.w2x : (N == 8) ? T.w3x : T.w4x;
    float z00 = 0.0f, z10 = 0.0f, z11 = 0.0f;
    float q00[NC], t[NC];
#pragma unroll
    for(int c = 0; c < NC; ++c) { q00[c] = 0.0f, t[c] = 0.0f; }
    for(int k = 0; k < cell.n; ++k)
    {
        const uint32_t s = (uint32_t) (sel >> (4 * k)) & 15u;
        const uint32_t v = L.px(cell.at(k));
        z00 = fadd(z00, wx[s][0]);
        z10 = fadd(z10, wx[s][1]);
        z11 = fadd(z11, wx[s][2]);
        const float w = wx[s][3];
#pragma unroll
        for(int c = 0; c < NC; ++c)
        {
            const float pc = (float) byte_of(v, c);
            q00[c] = fadd(q00[c], fmul(w, pc));
            t[c] = fadd(t[c], pc);
        }
    }
    const float z01 = z10;
    float det = fsub(fmul(z00, z11), fmul(z01, z10));
    if(det != 0.0f) { det = fdiv(1.0f, det); }
    const float iz00 = fmul(z11, det), iz01 = fmul(-z01, det), iz10 = fmul(-z10, det), iz11 = fmul(z00, det);
    const float inv255 = fdiv(1.0f, 255.0f);
#pragma unroll
    for(int c = 0; c < NC; ++c)
    {
        const float q10 = fsub(t[c], q00[c]);
        float l = fadd(fmul(iz00, q00[c]), fmul(iz01, q10));
        float h = fadd(fmul(iz10, q00[c]), fmul(iz11, q10));
        if((l < 0.0f) || (h > 255.0f))
        {
            // rare: re-scan the channel (bc7enc.cpp:333-347)
            uint32_t lo_v = 0xFFFFFFFFu, hi_v = 0;
            for(int k = 0; k < cell.n; ++k)
            {
                const uint32_t pc = (L.px(cell.at(k)) >> (8 * c)) & 255u;
                lo_v = umin(lo_v, pc), hi_v = umax(hi_v, pc);
            }
            if(lo_v == hi_v) { l = (float) lo_v, h = (float) hi_v; }
        }
        xl[c] = fmul(l, inv255);
        xh[c] = fmul(h, inv255);
    }
    if(!ALPHA) { xl[3] = xh[3] = fmul(255.0f, inv255); }
}

// ---------------------------------------------------------------------------------------------------- color_cell_compression
// bc7enc.cpp:1101-1441
// UBER == false compiles the search without the uber-level stages (P.uber_level must be 0): the default-parameter
// kernels then carry only the two stages they execute, which keeps their working set of instructions small.
template<int MODE, bool ALPHA, bool PERC, int KV, bool UBER, int STRIDE>
VKT_FN uint64_t compress_cell(const Bc7Tables &T, const Bc7KernelParams &P, Lane<STRIDE> L, CellRef cell, Cell &out)
{
    typedef ModeTraits<MODE> M;
    const int n = cell.n;
    out.err = kNoErr;
    out.lo = out.hi = 0;
    out.pbits = 0;
    out.sel = 0;

    const uint32_t first = L.px(cell.at(0));
    float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    bool same = true;
    for(int k = 0; k < n; ++k)
    {
        const uint32_t v = L.px(cell.at(k));
        if(MODE == 7) { same = same && (v == first); }
        else { same = same && (((v ^ first) & 0x00FFFFFFu) == 0u); }
#pragma unroll
        for(int c = 0; c < 4; ++c) { sum[c] = fadd(sum[c], (float) byte_of(v, c)); }
    }
    if((MODE == 1 || MODE == 7) && same) { return solid_cell<(MODE == 7) ? 7 : 1, PERC, STRIDE>(T, P, L, cell, first, out); }

    // mean (bc7enc.cpp:1144-1156): sums of small integers are exact in float in any order
    float mean_s[4], mean[4];
    {
        const float inv_n = fdiv(1.0f, (float) n);
        const float inv_n255 = fdiv(1.0f, fmul((float) n, 255.0f));
#pragma unroll
        for(int c = 0; c < 4; ++c)
        {
            mean_s[c] = fmul(sum[c], inv_n);
            mean[c] = satf(fmul(sum[c], inv_n255));
        }
    }

    float axis[4];
    if(ALPHA)
    {
        // incremental RGBA PCA, bc7enc.cpp:1160-1177
        axis[0] = axis[1] = axis[2] = axis[3] = 0.0f;
        for(int k = 0; k < n; ++k)
        {
            const uint32_t v = L.px(cell.at(k));
            float col[4], nrm[4];
#pragma unroll
            for(int c = 0; c < 4; ++c) { col[c] = fsub((float) byte_of(v, c), mean_s[c]); }
#pragma unroll
            for(int c = 0; c < 4; ++c) { nrm[c] = k ? axis[c] : col[c]; }
            {
                float s = fadd(fadd(fadd(fmul(nrm[0], nrm[0]), fmul(nrm[1], nrm[1])), fmul(nrm[2], nrm[2])), fmul(nrm[3], nrm[3]));
                if(s != 0.0f)
                {
                    s = fdiv(1.0f, fsqrt(s));
#pragma unroll
                    for(int c = 0; c < 4; ++c) { nrm[c] = fmul(nrm[c], s); }
                }
            }
#pragma unroll
            for(int j = 0; j < 4; ++j)
            {
                // dot(color * color[j], n): ((a0*n0 + a1*n1) + a2*n2) + a3*n3 with a_c = color[c] * color[j]
                const float d = fadd(fadd(fadd(fmul(fmul(col[0], col[j]), nrm[0]), fmul(fmul(col[1], col[j]), nrm[1])),
                                          fmul(fmul(col[2], col[j]), nrm[2])),
                                     fmul(fmul(col[3], col[j]), nrm[3]));
                axis[j] = fadd(axis[j], d);
            }
        }
        float s = fadd(fadd(fadd(fmul(axis[0], axis[0]), fmul(axis[1], axis[1])), fmul(axis[2], axis[2])), fmul(axis[3], axis[3]));
        if(s != 0.0f)
        {
            s = fdiv(1.0f, fsqrt(s));
#pragma unroll
            for(int c = 0; c < 4; ++c) { axis[c] = fmul(axis[c], s); }
        }
    }
    else
    {
        // covariance + 3 power iterations, bc7enc.cpp:1181-1218
        float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f, c3 = 0.0f, c4 = 0.0f, c5 = 0.0f;
        for(int k = 0; k < n; ++k)
        {
            const uint32_t v = L.px(cell.at(k));
            const float r = fsub((float) byte_of(v, 0), mean_s[0]), g = fsub((float) byte_of(v, 1), mean_s[1]),
                        b = fsub((float) byte_of(v, 2), mean_s[2]);
            c0 = fadd(c0, fmul(r, r));
            c1 = fadd(c1, fmul(r, g));
            c2 = fadd(c2, fmul(r, b));
            c3 = fadd(c3, fmul(g, g));
            c4 = fadd(c4, fmul(g, b));
            c5 = fadd(c5, fmul(b, b));
        }
        float vr = .9f, vg = 1.0f, vb = .7f;
#pragma unroll
        for(int it = 0; it < 3; ++it)
        {
            float r = fadd(fadd(fmul(vr, c0), fmul(vg, c1)), fmul(vb, c2));
            float g = fadd(fadd(fmul(vr, c1), fmul(vg, c3)), fmul(vb, c4));
            float b = fadd(fadd(fmul(vr, c2), fmul(vg, c4)), fmul(vb, c5));
            float m = fabsf(r) > fabsf(g) ? fabsf(r) : fabsf(g);
            m = m > fabsf(b) ? m : fabsf(b);
            if(m > 1e-10f)
            {
                m = fdiv(1.0f, m);
                r = fmul(r, m), g = fmul(g, m), b = fmul(b, m);
            }
            vr = r, vg = g, vb = b;
        }
        float len = fadd(fadd(fmul(vr, vr), fmul(vg, vg)), fmul(vb, vb));
        if(len < 1e-10f) { axis[0] = axis[1] = axis[2] = axis[3] = 0.0f; }
        else
        {
            len = fdiv(1.0f, fsqrt(len));
            axis[0] = fmul(vr, len), axis[1] = fmul(vg, len), axis[2] = fmul(vb, len), axis[3] = 0.0f;
        }
    }

    // fallback axis, bc7enc.cpp:1223-1230
    if(fadd(fadd(fadd(fmul(axis[0], axis[0]), fmul(axis[1], axis[1])), fmul(axis[2], axis[2])), fmul(axis[3], axis[3])) < .5f)
    {
        if(PERC) { axis[0] = .213f, axis[1] = .715f, axis[2] = .072f, axis[3] = ALPHA ? .715f : 0.0f; }
        else { axis[0] = 1.0f, axis[1] = 1.0f, axis[2] = 1.0f, axis[3] = ALPHA ? 1.0f : 0.0f; }
        float s = fadd(fadd(fadd(fmul(axis[0], axis[0]), fmul(axis[1], axis[1])), fmul(axis[2], axis[2])), fmul(axis[3], axis[3]));
        if(s != 0.0f)
        {
            s = fdiv(1.0f, fsqrt(s));
#pragma unroll
            for(int c = 0; c < 4; ++c) { axis[c] = fmul(axis[c], s); }
        }
    }

    // projection extrema, bc7enc.cpp:1232-1246
    float l = 1e+9f, h = -1e+9f;
    for(int k = 0; k < n; ++k)
    {
        const uint32_t v = L.px(cell.at(k));
        float q[4];
#pragma unroll
        for(int c = 0; c < 4; ++c) { q[c] = fsub((float) byte_of(v, c), mean_s[c]); }
        const float d = fadd(fadd(fadd(fmul(q[0], axis[0]), fmul(q[1], axis[1])), fmul(q[2], axis[2])), fmul(q[3], axis[3]));
        l = (l < d) ? l : d;// minimumf / maximumf, bc7enc.cpp:17,21
        h = (h > d) ? h : d;
    }
    l = fmul(l, fdiv(1.0f, 255.0f));
    h = fmul(h, fdiv(1.0f, 255.0f));

    float xl[4], xh[4];
#pragma unroll
    for(int c = 0; c < 4; ++c)
    {
        xl[c] = satf(fadd(mean[c], fmul(axis[c], l)));
        xh[c] = satf(fadd(mean[c], fmul(axis[c], h)));
    }
    {
        // bc7enc.cpp:1258: dot with (1,1,1,1) = ((x*1 + y*1) + z*1) + w*1
        const float dl = fadd(fadd(fadd(xl[0], xl[1]), xl[2]), xl[3]);
        const float dh = fadd(fadd(fadd(xh[0], xh[1]), xh[2]), xh[3]);
        if(dl > dh)
        {
#pragma unroll
            for(int c = 0; c < 4; ++c)
            {
                const float t = xl[c];
                xl[c] = xh[c], xh[c] = t;
            }
        }
    }

    // Stage machine: one call site for (least squares ->) quantise -> evaluate.
    //   stage 0        PCA endpoints                                        bc7enc.cpp:1279
    //   stage 1        least squares from the current selectors            bc7enc.cpp:1282-1298
    //   stage 2,3,4    uber >= 1: min+1 / max-1 / both                      bc7enc.cpp:1300-1378
    //   stage 5..      uber >= 2: selector rescaling (ly, hy)               bc7enc.cpp:1380-1410
    constexpr int max_sel_v = M::N - 1;
    uint64_t base = 0;
    uint32_t min_sel = 16, max_sel = 0;
    int stage = 0, ly = 0, hy = 0, Q = 1;
    for(;;)
    {
        if(stage > 0)
        {
            uint64_t trial = 0;
            if(stage == 1) { trial = out.sel; }
            else if(stage <= 4)
            {
                if(stage == 2)
                {
                    base = out.sel;
                    for(int k = 0; k < n; ++k)
                    {
                        const uint32_t s = (uint32_t) (base >> (4 * k)) & 15u;
                        min_sel = umin(min_sel, s);
                        max_sel = umax(max_sel, s);
                    }
                }
                for(int k = 0; k < n; ++k)
                {
                    uint32_t s = (uint32_t) (base >> (4 * k)) & 15u;
                    if(stage != 3 && (s == min_sel) && (s < (uint32_t) max_sel_v)) { s++; }
                    else if(stage != 2 && (s == max_sel) && (s > 0)) { s--; }
                    trial |= (uint64_t) s << (4 * k);
                }
            }
            else
            {
                // clampf(floorf(max * (s - ly) / (hy - ly) + .5f), 0, max) per selector, bc7enc.cpp:1399, from the table
                const uint64_t map = VKT_UTAB(uber).m[(M::N == 4) ? 0 : (M::N == 8) ? 1 : 2][ly + 2][hy - (max_sel_v - 1)];
                for(int k = 0; k < n; ++k)
                {
                    const uint32_t s4 = (uint32_t) (base >> (4 * k)) & 15u;
                    trial |= (uint64_t) ((uint32_t) (map >> (4 * s4)) & 15u) << (4 * k);
                }
            }
            least_squares<MODE, ALPHA, STRIDE>(T, L, cell, trial, xl, xh);
        }
        if(!fit<MODE, ALPHA, PERC, KV, STRIDE, (UBER || ALPHA || MODE == 5)>(T, P, L, cell, xl, xh, out)) { return 0; }

        // advance
        if(stage == 0)
        {
            if(P.try_least_squares) { stage = 1; }
            else if(UBER && P.uber_level > 0) { stage = 2; }
            else { break; }
        }
        else if(stage == 1)
        {
            if(UBER && P.uber_level > 0) { stage = 2; }
            else { break; }
        }
        else if(stage < 4) { ++stage; }
        else
        {
            if(stage == 4)
            {
                const uint32_t thresh = ((uint32_t) n * 56u) >> 4;
                if(!((P.uber_level >= 2) && (out.err > thresh))) { break; }
                Q = (P.uber_level >= 4) ? ((int) P.uber_level - 2) : 1;
                ly = -Q, hy = max_sel_v - 1;
                stage = 5;
            }
            else
            {
                ++hy;
                if(hy > max_sel_v + Q) { hy = max_sel_v - 1, ++ly; }
            }
            if((ly == 0) && (hy == max_sel_v)) { ++hy; }// skip the identity mapping (hy range always continues past it)
            if(ly > 1) { break; }
        }
    }

    if(MODE == 1 || MODE == 7)
    {
        // mean as a single colour, bc7enc.cpp:1413-1438
        Cell avg;
        const uint32_t r = (uint32_t) f2i(fadd(.5f, fmul(mean[0], 255.0f))), g = (uint32_t) f2i(fadd(.5f, fmul(mean[1], 255.0f))),
                       b = (uint32_t) f2i(fadd(.5f, fmul(mean[2], 255.0f))), a = (uint32_t) f2i(fadd(.5f, fmul(mean[3], 255.0f)));
        const uint64_t e = solid_cell<(MODE == 7) ? 7 : 1, PERC, STRIDE>(T, P, L, cell, pack4(r, g, b, a), avg);
        if(e < out.err) { out = avg; }
    }
    return out.err;
}

// ---------------------------------------------------------------------------------------------------- partition estimate
// color_cell_compression_est_mode1 / _mode7 (bc7enc.cpp:1443-1709) for BOTH subsets of one partition, sums completed.
// `part` is warp-uniform.  T.est_idx[part] lists the texels of subset 0 (ascending) then those of subset 1, so each
// subset is a compact rolled loop over its own texels (no per-texel membership branches; the hot loop stays inside the
// L0 instruction cache) that handles two texels per trip for instruction-level parallelism.  All four channels are
// always processed: for mode 1 the block is opaque (alpha == 255 everywhere), so the alpha lane has zero extent and
// contributes exactly nothing.
//
// One texel against the subset's palette: project, pick the entry (thresholds are non-decreasing because the palette
// is monotone along a non-negative axis, so "highest satisfied threshold" is a balanced select tree), metric.
template<bool M7, bool PERC, int N>
VKT_FN uint32_t estimate_texel(const Bc7KernelParams &P, const Texel t, uint32_t axb, const uint32_t (&pal)[N], const int (&thr)[N - 1])
{
    const int d = (int) dp4a_u8(t.px, axb, 0u);
    uint32_t c;
    if(N == 8)
    {
        const uint32_t c01 = (d >= thr[0]) ? pal[1] : pal[0], c23 = (d >= thr[2]) ? pal[3] : pal[2];
        const uint32_t c45 = (d >= thr[4]) ? pal[5] : pal[4], c67 = (d >= thr[6]) ? pal[7] : pal[6];
        const uint32_t c03 = (d >= thr[1]) ? c23 : c01, c47 = (d >= thr[5]) ? c67 : c45;
        c = (d >= thr[3]) ? c47 : c03;
    }
    else
    {
        const uint32_t c01 = (d >= thr[0]) ? pal[1] : pal[0], c23 = (d >= thr[2]) ? pal[3] : pal[2];
        c = (d >= thr[1]) ? c23 : c01;
    }
    uint32_t e;
    if(PERC)
    {
        // mode 1 palettes carry g again in the (unused) alpha byte: luma is then a single dot product, 183 g + 183 g = 366 g
        Ycc e1;
        e1.l = M7 ? (int) dp4a_u8(c, 0x0025B76Du, dp4a_u8(c, 0x0000B700u, 0u)) : (int) dp4a_u8(c, 0xB725B76Du, 0u);
        e1.cr = (int) (byte_of(c, 0) << 9) - e1.l;
        e1.cb = (int) (byte_of(c, 2) << 9) - e1.l;
        const int dl = (e1.l - t.l) >> 8, dcr = (e1.cr - t.cr) >> 8, dcb = (e1.cb - t.cb) >> 8;
        // uint32 products, bc7enc.cpp:1533,1670
        e = (P.w[0] * (uint32_t) dl * (uint32_t) dl) + (P.w[1] * (uint32_t) dcr * (uint32_t) dcr) + (P.w[2] * (uint32_t) dcb * (uint32_t) dcb);
        if(M7)
        {
            const int dca = (int) (t.px >> 24) - (int) (c >> 24);
            e += P.w[3] * (uint32_t) dca * (uint32_t) dca;
        }
    }
    else
    {
        e = 0;
#pragma unroll
        for(int ch = 0; ch < (M7 ? 4 : 3); ++ch)
        {
            const int dd = (int) byte_of(c, ch) - (int) byte_of(t.px, ch);
            e += P.w[ch] * (uint32_t) (dd * dd);
        }
    }
    return e;
}

// (tuning: unroll factors of the two warp-uniform texel-pair loops; 1 = rolled, the shipped setting -- see DESIGN.md 5.0)
#ifndef VKT_EST_UNROLL_BBOX
#define VKT_EST_UNROLL_BBOX 1
#endif
#ifndef VKT_EST_UNROLL_ERR
#define VKT_EST_UNROLL_ERR 1
#endif
constexpr int kEstUnrollBbox = VKT_EST_UNROLL_BBOX, kEstUnrollErr = VKT_EST_UNROLL_ERR;
// UNI: `part` is warp-uniform (lists from __constant__ memory, uniform loop control); otherwise lane-varying (shared memory).
template<bool M7, bool PERC, int KV, bool UNI, int STRIDE>
VKT_FN uint64_t estimate_pair(const Bc7Tables &T, const Bc7KernelParams &P, Lane<STRIDE> L, uint32_t part, uint64_t best_so_far)
{
    constexpr int N = M7 ? 4 : 8;
    constexpr bool KEY28 = (KV == kKvKey28);
    // kKvExt keeps the reference's early-outs (a subset's running sum above the best error so far ends that subset,
    // bc7enc.cpp:1536-1537 / :1573-1574 / :1673-1674 / :1705-1706; the second subset is skipped unless the first stayed
    // below it, :1810): a low-frequency weight below 1 can make such a truncated sum the new minimum.  The other variants
    // complete every sum -- with a weight of exactly 1 the argmin is the same.
    constexpr bool EARLY = (KV == kKvExt);
    // UNI: texel pairs from __constant__ memory (one 64-bit broadcast load per trip, no index arithmetic)
    const EstPair *trips = VKT_UTAB(trips).t[UNI ? part : 0];
    const uint32_t tn = VKT_UTAB(trips).n[UNI ? part : 0];
    const uint8_t *order = T.est_idx[part];// !UNI (lane-varying partition) and the wide-error path
    const int n0 = (int) T.est_n0[part];
    uint64_t total = 0;
#pragma unroll 1
    for(int s = 0; s < 2; ++s)
    {
        if(EARLY && s && !(total < best_so_far)) { break; }
        const int k0 = s ? n0 : 0, k1 = s ? 16 : n0;                                           // !UNI: texel positions
        const uint32_t q0 = s ? (tn & 255u) : 0u, q1 = s ? (tn >> 8) : (tn & 255u);// UNI: pair positions
        // pass 1: bounding box, 16x2 SIMD lanes (r | g << 16) and (b | a << 16); an odd tail repeats its last texel
        uint32_t l_rg = 0x00FF00FFu, l_ba = 0x00FF00FFu, h_rg = 0u, h_ba = 0u;
        auto grow = [&](uint32_t v0, uint32_t v1) {
            const uint32_t rg0 = prmt(v0, 0u, 0x4140u), ba0 = prmt(v0, 0u, 0x4342u);
            const uint32_t rg1 = prmt(v1, 0u, 0x4140u), ba1 = prmt(v1, 0u, 0x4342u);
            l_rg = vmin_u16x2(l_rg, vmin_u16x2(rg0, rg1)), h_rg = vmax_u16x2(h_rg, vmax_u16x2(rg0, rg1));
            l_ba = vmin_u16x2(l_ba, vmin_u16x2(ba0, ba1)), h_ba = vmax_u16x2(h_ba, vmax_u16x2(ba0, ba1));
        };
        if(UNI)
        {
#if defined(VKT_EST_BBOX_PREFETCH)
            // software-pipelined: the next pair's texels are requested before the current pair is folded in, so the shared
            // load latency hides behind the min / max work (the last trip re-requests its own pair: no branch, no overrun)
            EstPair w = trips[q0];
            uint32_t v0 = L.px((int) w.i0), v1 = L.px((int) w.i1);
#pragma unroll 1
            for(uint32_t q = q0; q < q1; ++q)
            {
                w = trips[umin(q + 1u, q1 - 1u)];
                const uint32_t n0v = L.px((int) w.i0), n1v = L.px((int) w.i1);
                grow(v0, v1);
                v0 = n0v, v1 = n1v;
            }
#else
#pragma unroll kEstUnrollBbox
            for(uint32_t q = q0; q < q1; ++q)
            {
                const EstPair w = trips[q];
                grow(L.px((int) w.i0), L.px((int) w.i1));
            }
#endif
        }
        else
        {
#pragma unroll 1
            for(int k = k0; k < k1; k += 2) { grow(L.px(order[k]), L.px(order[(k + 1 < k1) ? k + 1 : k])); }
        }
        // palette: lo*(64-w) + hi*w + 32 = 64*lo + (hi-lo)*w + 32 per 16-bit lane (<= 16352: no carries between lanes)
        // -- computed four-fold, see below
        const uint32_t ax_rg = h_rg - l_rg, ax_ba = h_ba - l_ba;// hi >= lo per lane (subsets are never empty)
        const uint32_t axb = prmt(ax_rg, ax_ba, M7 ? 0x6420u : 0x1420u);// (ar, ag, ab, aa) as bytes; mode 1: (ar, ag, ab, 0)
        const uint32_t c4_rg = l_rg * 256u + 0x00800080u, c4_ba = l_ba * 256u + 0x00800080u;
        uint32_t pal[N];
        int thr[N - 1];
        {
            int dots[N];
#pragma unroll
            for(int j = 0; j < N; ++j)
            {
                // packed (r, g, b, a) for mode 7, (r, g, b, g) for mode 1 (estimate_texel's single-dot luma)
                if(j == 0) { pal[j] = prmt(l_rg, l_ba, M7 ? 0x6420u : 0x2420u); }
                else if(j == N - 1) { pal[j] = prmt(h_rg, h_ba, M7 ? 0x6420u : 0x2420u); }
                else
                {
                    const uint32_t w = (uint32_t) selw(N, j);
                    // 4 * (lo*64 + (hi-lo)*w + 32) <= 65408 still fits the 16-bit lane: the interpolated channel is then byte 1 / 3
                    // of each word and PRMT gathers it without a shift
                    pal[j] = prmt(ax_rg * (4u * w) + c4_rg, ax_ba * (4u * w) + c4_ba, M7 ? 0x7531u : 0x3531u);
                }
                dots[j] = (int) dp4a_u8(pal[j], axb, 0u);
            }
#pragma unroll
            for(int j = 0; j < N - 1; ++j) { thr[j] = (dots[j] + dots[j + 1] + 1) >> 1; }
        }
        // pass 2: two texels per trip; an odd tail evaluates its last texel twice and drops the copy
        if(KEY28)
        {
            uint32_t sum = 0;// every term < 2^28 (host-checked), at most 16 terms
            if(UNI)
            {
#pragma unroll kEstUnrollErr
                for(uint32_t q = q0; q < q1; ++q)
                {
                    const EstPair w = trips[q];
                    const uint32_t e0 = estimate_texel<M7, PERC, N>(P, L.at((int) w.i0), axb, pal, thr);
                    const uint32_t e1 = estimate_texel<M7, PERC, N>(P, L.at((int) w.i1), axb, pal, thr);
                    sum += e0 + (e1 & w.second);
                }
            }
            else
            {
#pragma unroll 1
                for(int k = k0; k < k1; k += 2)
                {
                    const bool two = (k + 1 < k1);
                    const Texel t0 = L.at(order[k]), t1 = L.at(order[two ? k + 1 : k]);
                    const uint32_t e0 = estimate_texel<M7, PERC, N>(P, t0, axb, pal, thr);
                    const uint32_t e1 = estimate_texel<M7, PERC, N>(P, t1, axb, pal, thr);
                    sum += e0 + (two ? e1 : 0u);
                }
            }
            total += sum;
        }
        else
        {
            uint64_t sub = 0;
#pragma unroll 1
            for(int k = k0; k < k1; ++k)
            {
                const uint32_t e = estimate_texel<M7, PERC, N>(P, L.at(order[k]), axb, pal, thr);
                // perceptual: `int ie; total_err += ie` (bc7enc.cpp:1533-1535, sign-extended); linear: a uint32_t sum (:1572)
                sub += PERC ? (uint64_t) (int64_t) (int32_t) e : (uint64_t) e;
                if(EARLY && (sub > best_so_far)) { break; }
            }
            total += sub;
        }
    }
    return total;
}

// estimate_partition, bc7enc.cpp:1754-1838.  Must be called by the whole (converged) warp; `active` lanes want a result,
// the others only lend their issue slots.
//   iterations 0..34   warp-uniform scan: every lane scores the same partition for its own block; a ballot skips
//                      candidates no lane of the batch needs (filterbank, bc7enc.cpp:1786-1799).
//   iterations 35..63  only blocks whose best candidate after iteration 34 is the checkerboard (partition 34) go on
//                      (bc7enc.cpp:1829-1830; ~2 % of blocks).  Scanning them warp-uniformly would make the whole warp pay
//                      29 more candidates for one lane, so the warp turns around instead: the 29 candidates of ONE such
//                      block are spread over the lanes (lane-varying partition, the block's column read by all lanes as
//                      a shared-memory broadcast) and the winner is a (error, iteration) warp minimum -- the same
//                      "first strictly smaller wins" the sequential scan implements.
template<bool M7, bool PERC, int KV, int STRIDE>
VKT_FN uint32_t estimate_partition(const Bc7Tables &T, const Bc7KernelParams &P, Lane<STRIDE> L, bool active)
{
    const uint32_t total_partitions = umin(P.max_partitions, 64u);
    if(total_partitions <= 1) { return 0; }
    constexpr bool KEY28 = (KV == kKvKey28);
    // (kKvExt: a candidate's early-outs depend on the best error of all candidates before it, so the whole scan stays in
    // sequence per block)
    constexpr uint32_t kUniformIters = (KV == kKvExt) ? 64 : 35, kKeyIters = 14;
    const uint32_t uniform_end = umin(total_partitions, kUniformIters);
    uint64_t best_err = kNoErr;
    uint32_t best_partition = 0, best_it = 0;
    uint32_t key = 0;
    bool running = active;
    Lane<STRIDE> Lc = L;// the block this lane is scoring (its own, until the regrouping below)
#if defined(__CUDA_ARCH__)
    // Filterbank phase (iterations 14..34): which candidates a block still needs depends only on its key partition -- the
    // best of the first 14 (bc7enc.cpp:1786-1799, :1832-1833) -- and a warp pays for the union over its 32 blocks (21 of 21
    // on average, where one block needs 14.2).  So at iteration 14 the CTA's blocks are regrouped by key: a counting sort
    // through shared memory hands every lane the column pointer and the running state of another block, the warps run the
    // phase on key-sorted batches (union 16.3 on the same data), and the results travel back the same way.  Which lane
    // scores a block has no influence on the block's result.  All warps of the CTA reach this point (the early exit of
    // the scan is held back until then), and the whole CTA calls estimate_partition exactly once.  Measured: -1.9 % kernel
    // time (the three barriers eat most of the 13 % fewer candidates); keeping the adopted blocks for the rest of the
    // kernel instead of sending results back was slower.
    // (opaque kernels only: on the alpha kernels, 2 CTAs per SM, the barriers cost more than the smaller union saves)
#if defined(VKT_EST_REGROUP_ALPHA)
    constexpr bool kRegroupM7 = true;
#else
    constexpr bool kRegroupM7 = false;
#endif
    const bool regroup = (!M7 || kRegroupM7) && KEY28 && (STRIDE >= 64) && (STRIDE % 32 == 0) && P.filterbank && (uniform_end > kKeyIters);// CTA-uniform
    CtaScratch<STRIDE> *S = reinterpret_cast<CtaScratch<STRIDE> *>(L.p - threadIdx.x + 16 * STRIDE);
    uint32_t src = threadIdx.x;
    bool adopted = false;
#else
    const bool regroup = false;
    (void) best_it;
#endif
#pragma unroll 1
    for(uint32_t it = 0; it < uniform_end; ++it)
    {
#if defined(__CUDA_ARCH__)
        if(regroup && it == kKeyIters)
        {
            running = running && (best_err > 0);
            // One counting sort over the CTA's blocks: warp w adopts blocks [32 w, 32 w + 32) of the key order.  (Round 1 sorted
            // within each class of thread index mod 8 and handed a block to a thread of its class, so that the 128-bit loads of
            // the adopted columns kept their conflict-free bank pattern -- at the price of warps that mix eight key windows.
            // Measured in round 2: the purer warps are worth more than the extra shared-memory wavefronts, 2.000 -> 1.984 ms;
            // VKT_REGROUP_BY_CLASS restores the old placement.)
            // The sort key is not the key iteration itself but its place in an order that puts keys with similar candidate sets
            // next to each other: a warp takes 32 consecutive blocks of the sorted CTA, i.e. two or three neighbouring bins, and
            // pays for the union of their sets.  The order was annealed on the candidate sets of T.pred against the key histograms
            // of four textures and a uniform one (all within 0.1 of each other): mean union per warp 17.3 -> 15.7 candidates of
            // the 14.1 a block needs, the slowest warp of a CTA -- which the write-back barrier waits for -- 21.0 -> 18.0.
#if !defined(VKT_REGROUP_PLAIN_BINS)
            constexpr uint64_t kBinOfKey = 0xC2178590AD63B4ull;// nibble k: bin of key iteration k = {4,11,3,6,13,10,0,9,5,8,7,1,2,12}
            const uint32_t bin = running ? ((uint32_t) (kBinOfKey >> (4u * best_it)) & 15u) : 15u;// best_it < 14
#else
            const uint32_t bin = running ? best_it : 15u;// best_it < 14
#endif
#if !defined(VKT_REGROUP_BY_CLASS)
            const uint32_t cls = 0u;// one sort over the whole CTA (see above)
#else
            const uint32_t cls = threadIdx.x & 7u;
#endif
            const uint32_t pos = atomicAdd(&S->cnt[cls][bin], 1u);
            S->err[threadIdx.x] = (uint32_t) best_err;// kNoErr -> 0xFFFFFFFF, above every real error (<= 2^32 - 16)
            S->info[threadIdx.x] = (uint16_t) (best_partition | (running ? 256u : 0u));
            __syncthreads();
            uint32_t q = pos;
#pragma unroll
            for(uint32_t b = 0; b < 15; ++b) { q += (b < bin) ? S->cnt[cls][b] : 0u; }
            constexpr uint32_t kPerWarp = 4;// lanes of one class in a warp
#if !defined(VKT_REGROUP_BY_CLASS)
            S->perm[q] = (uint16_t) threadIdx.x;
#else
            S->perm[(q / kPerWarp) * 32u + (q % kPerWarp) * 8u + cls] = (uint16_t) threadIdx.x;
#endif
            __syncthreads();
            src = S->perm[threadIdx.x];
            const uint32_t e = S->err[src], info = S->info[src];
            best_err = (e == 0xFFFFFFFFu) ? kNoErr : (uint64_t) e;
            best_partition = info & 255u, key = best_partition, running = (info & 256u) != 0u;
            Lc.p = L.p + ((int) src - (int) threadIdx.x);
            adopted = true;
        }
#endif
        const uint32_t part = VKT_UTAB(order)[it];
        running = running && (best_err > 0);// loop condition of the reference
        bool need = running;
        if(need && P.filterbank && (it >= 14) && (it <= 34))
        {
            if((T.pred[part] & (1u << (key + 1))) == 0)
            {
                if(it == 34) { running = false; }
                need = false;
            }
        }
        if(!warp_any(need))
        {
            if(!warp_any(running) && !(regroup && it < kKeyIters)) { break; }// (a warp must not leave before the regrouping)
            continue;
        }
        uint64_t err = estimate_pair<M7, PERC, KV, true, STRIDE>(T, P, Lc, part, best_err);
        // bc7enc.cpp:1817-1820; with m_low_frequency_partition_weight == 1.0f (every variant but kKvExt) it is the identity
        if((KV == kKvExt) && (part < 16)) { err = weigh_d(err, P.low_freq_weight); }
        if(need)
        {
            if(err < best_err) { best_err = err, best_partition = part, best_it = it; }
            if((part == 34) && (best_partition != 34)) { running = false; }
            if(it == 13) { key = best_partition; }
        }
    }
#if defined(__CUDA_ARCH__)
    if(regroup)// (CTA-uniform; `adopted` is true in every warp: the scan cannot end before iteration 14)
    {
        if(adopted)
        {
            S->err[src] = (best_err == kNoErr) ? 0xFFFFFFFFu : (uint32_t) best_err;
            S->info[src] = (uint16_t) (best_partition | (running ? 256u : 0u));
        }
        __syncthreads();
        const uint32_t e = S->err[threadIdx.x], info = S->info[threadIdx.x];
        best_err = (e == 0xFFFFFFFFu) ? kNoErr : (uint64_t) e;
        best_partition = info & 255u, running = (info & 256u) != 0u;
    }
#endif
    if(total_partitions > kUniformIters)
    {
        warp_sync();// lanes are about to read each other's columns (written by their owners in prepare_lane)
        running = running && (best_err > 0);
        uint32_t pending = warp_ballot(running);
        const uint32_t lane = warp_lane();
#pragma unroll 1
        while(pending)
        {
            const uint32_t src = (uint32_t) ctz32(pending);
            pending &= pending - 1u;
            const Lane<STRIDE> Ls{L.p + ((int) src - (int) lane)};// lane columns are consecutive records
            uint64_t best_key = kNoErr;
#pragma unroll 1
            for(uint32_t base = kUniformIters; base < total_partitions; base += kWarpLanes)
            {
                const uint32_t it = base + lane;
                if(it < total_partitions)
                {
                    const uint64_t err = estimate_pair<M7, PERC, KV, false, STRIDE>(T, P, Ls, VKT_UTAB(order)[it], kNoErr);
                    const uint64_t k = (err << 6) | (uint64_t) it;// err < 2^36: no overflow
                    best_key = k < best_key ? k : best_key;
                }
            }
            best_key = warp_min_u64(best_key);
            if(lane == src)
            {
                const uint64_t err = best_key >> 6;
                if(err < best_err) { best_err = err, best_partition = VKT_UTAB(order)[(uint32_t) best_key & 63u]; }
            }
        }
    }
    return best_partition;
}

// ---------------------------------------------------------------------------------------------------- bit packing
struct Bits128
{
    uint64_t lo, hi;
    uint32_t ofs;
    VKT_FN void put(uint32_t v, uint32_t n)// LSB-first, set_block_bits bc7enc.cpp:1840-1852
    {
        if(ofs < 64)
        {
            lo |= (uint64_t) v << ofs;
            if(ofs + n > 64) { hi |= (uint64_t) v >> (64 - ofs); }
        }
        else { hi |= (uint64_t) v << (ofs - 64); }
        ofs += n;
    }
};

struct BlockSolution
{
    uint32_t mode, partition;
    uint64_t sel, asel;// 4 bits per texel, block texel order
    uint32_t lo[2], hi[2];
    uint32_t pbits[2];
};

// encode_bc7_block restricted to the emitted modes 1, 5, 6, 7 (bc7enc.cpp:1867-2037, layouts SURVEY.md App. B)
VKT_FN void pack_block(const Bc7Tables &T, const BlockSolution &s, uint32_t out[4])
{
    const uint32_t mode = s.mode;
    const bool two = (mode == 1) || (mode == 7);
    const uint32_t mask = two ? T.part2[s.partition] : 0u;
    const uint32_t ibits = (mode == 6) ? 4u : (mode == 1) ? 3u : 2u;
    const uint32_t cbits = (mode == 1) ? 6u : (mode == 7) ? 5u : 7u;
    const uint32_t abits = (mode == 5) ? 8u : (mode == 6) ? 7u : (mode == 7) ? 5u : 0u;
    const uint32_t top = 1u << (ibits - 1), full = (1u << ibits) - 1;
    uint64_t sel = s.sel, asel = s.asel;
    uint32_t lo[2] = {s.lo[0], s.lo[1]}, hi[2] = {s.hi[0], s.hi[1]}, pb[2] = {s.pbits[0], s.pbits[1]};
    const uint32_t anchor1 = two ? T.anchor2[s.partition] : 16u;
    // 4-bit-per-texel masks of the two subsets
    uint64_t sub1 = 0;
    for(int i = 0; i < 16; ++i) { sub1 |= (uint64_t) ((mask >> i) & 1u) << (4 * i); }
    sub1 *= 15ull;
    const uint64_t fullmask = 0x1111111111111111ull * full;
#pragma unroll
    for(uint32_t k = 0; k < 2; ++k)
    {
        if(k == 1 && !two) { break; }
        const uint32_t a = k ? anchor1 : 0u;
        const uint64_t members = k ? sub1 : ~sub1;
        if((uint32_t) (sel >> (4 * a)) & top)
        {
            sel = (sel & ~members) | ((fullmask - sel) & members & fullmask);
            if(mode == 5)
            {
                // separate alpha selectors: only RGB trades places here
                const uint32_t l0 = lo[k], h0 = hi[k];
                lo[k] = (h0 & 0x00FFFFFFu) | (l0 & 0xFF000000u);
                hi[k] = (l0 & 0x00FFFFFFu) | (h0 & 0xFF000000u);
            }
            else
            {
                const uint32_t t = lo[k];
                lo[k] = hi[k], hi[k] = t;
            }
            if(mode != 1) { pb[k] = ((pb[k] & 1u) << 1) | ((pb[k] >> 1) & 1u); }
        }
        if(mode == 5)
        {
            if((uint32_t) asel & 2u)
            {
                asel = 0x3333333333333333ull - asel;
                const uint32_t t = lo[0];
                lo[0] = (lo[0] & 0x00FFFFFFu) | (hi[0] & 0xFF000000u);
                hi[0] = (hi[0] & 0x00FFFFFFu) | (t & 0xFF000000u);
            }
        }
    }

    Bits128 w = {0, 0, 0};
    w.put(1u << mode, mode + 1);
    if(mode == 5) { w.put(0, 2); }
    if(two) { w.put(s.partition, 6); }
    const uint32_t ncomp = (mode >= 4) ? 4u : 3u, nsub = two ? 2u : 1u;
    for(uint32_t c = 0; c < ncomp; ++c)
    {
        for(uint32_t k = 0; k < nsub; ++k)
        {
            const uint32_t nb = (c == 3) ? abits : cbits;
            w.put((lo[k] >> (8 * c)) & 255u, nb);
            w.put((hi[k] >> (8 * c)) & 255u, nb);
        }
    }
    if(mode != 5)
    {
        for(uint32_t k = 0; k < nsub; ++k)
        {
            w.put(pb[k] & 1u, 1);
            if(mode != 1) { w.put((pb[k] >> 1) & 1u, 1); }
        }
    }
    for(uint32_t i = 0; i < 16; ++i)
    {
        const uint32_t nb = (i == 0 || i == anchor1) ? ibits - 1 : ibits;
        w.put((uint32_t) (sel >> (4 * i)) & 15u, nb);
    }
    if(mode == 5)
    {
        for(uint32_t i = 0; i < 16; ++i) { w.put((uint32_t) (asel >> (4 * i)) & 15u, i == 0 ? 1u : 2u); }
    }
    out[0] = (uint32_t) w.lo, out[1] = (uint32_t) (w.lo >> 32), out[2] = (uint32_t) w.hi, out[3] = (uint32_t) (w.hi >> 32);
}

// ---------------------------------------------------------------------------------------------------- two-subset modes
// mode 1 (bc7enc.cpp:2336-2397) / mode 7 (bc7enc.cpp:2193-2259): estimate, fit both subsets, arbitrate.
// Called by the whole converged warp; lanes with want == false only help the estimator.
// Returns the weighted error, or kNoErr when not better than best_err.
template<int MODE, bool PERC, int KV, bool UBER, int STRIDE>
VKT_FN uint64_t two_subset_cells(const Bc7Tables &T, const Bc7KernelParams &P, Lane<STRIDE> L, uint32_t part, bool want, uint64_t best_err,
                                 BlockSolution &sol)
{
    constexpr bool ALPHA = (MODE == 7);
    if(!want) { return kNoErr; }
    const uint32_t mask = T.part2[part];
    // element lists of the two subsets (ascending texel order, as the reference gathers them): nibble-packed work list
    const int n0 = (int) T.est_n0[part], n1 = 16 - n0;
    const uint64_t perm = T.est_perm[part];
    const uint64_t perm0 = perm & ((1ull << (4 * n0)) - 1ull), perm1 = perm >> (4 * n0);// 1 <= n0 <= 15
    Cell c[2];
    const float mw = (MODE == 1) ? P.mode1_w : P.mode7_w;
    uint64_t trial = 0;
    // Every lane fits its LARGER subset first: the warp's texel loops then run max(8..15) + max(1..8) trips instead of twice
    // max(1..15).  The two fits are independent and their errors only add, so the order changes neither the result nor
    // the outcome of the early exit below (weigh() is monotone: a partial sum above best_err means the total is too).
    const int second = (n1 > n0) ? 0 : 1;// subset fitted in the second pass
#pragma unroll 1
    for(int pass = 0; pass < 2; ++pass)
    {
        const int s = pass ? second : 1 - second;
        const CellRef cell = {s ? perm1 : perm0, s ? n1 : n0};
        Cell r;
        trial += compress_cell<MODE, ALPHA, PERC, KV, UBER, STRIDE>(T, P, L, cell, r);
        if(s) { c[1] = r; }
        else { c[0] = r; }
        if(weigh(trial, mw) > best_err) { return kNoErr; }// bc7enc.cpp:2377/2234: cannot be adopted any more
    }
    const uint64_t werr = weigh(trial, mw);
    if(!(werr < best_err)) { return kNoErr; }
    sol.mode = MODE;
    sol.partition = part;
    uint64_t sel = 0;
    {
        int k0 = 0, k1 = 0;
        for(int i = 0; i < 16; ++i)
        {
            uint32_t s;
            if((mask >> i) & 1u) { s = (uint32_t) (c[1].sel >> (4 * k1++)) & 15u; }
            else { s = (uint32_t) (c[0].sel >> (4 * k0++)) & 15u; }
            sel |= (uint64_t) s << (4 * i);
        }
    }
    sol.sel = sel;
    sol.asel = 0;
    sol.lo[0] = c[0].lo, sol.hi[0] = c[0].hi, sol.pbits[0] = c[0].pbits;
    sol.lo[1] = c[1].lo, sol.hi[1] = c[1].hi, sol.pbits[1] = c[1].pbits;
    return werr;
}

// ---------------------------------------------------------------------------------------------------- mode 5 alpha
// scalar alpha search of handle_alpha_block_mode5, bc7enc.cpp:2067-2136.  Returns the alpha error.
template<int STRIDE>
VKT_FN uint64_t mode5_alpha(const Bc7Tables &T, const Bc7KernelParams &P, Lane<STRIDE> L, uint32_t lo_a, uint32_t hi_a, uint32_t &out_lo,
                            uint32_t &out_hi, uint64_t &out_sel)
{
    if(lo_a == hi_a)
    {
        out_lo = lo_a, out_hi = hi_a, out_sel = 0;
        return 0;
    }
    uint64_t best = kNoErr;
    const uint32_t passes = (P.uber_level >= 1) ? 3u : 2u;
    for(uint32_t pass = 0; pass < passes; ++pass)
    {
        int v[4];
        v[0] = (int) lo_a, v[3] = (int) hi_a;
        v[1] = (v[0] * (64 - 21) + v[3] * 21 + 32) >> 6;
        v[2] = (v[0] * (64 - 43) + v[3] * 43 + 32) >> 6;
        uint64_t tsel = 0, terr = 0;
        float z00 = 0.0f, z10 = 0.0f, z11 = 0.0f, q00 = 0.0f, t = 0.0f;
        for(int i = 0; i < 16; ++i)
        {
            const int a = (int) (L.px(i) >> 24);
            int s = 0;
            int be = iabs(a - v[0]);
            int e = iabs(a - v[1]);
            if(e < be) { be = e, s = 1; }
            e = iabs(a - v[2]);
            if(e < be) { be = e, s = 2; }
            e = iabs(a - v[3]);
            if(e < be) { be = e, s = 3; }
            tsel |= (uint64_t) (uint32_t) s << (4 * i);
            terr += (uint64_t) ((uint32_t) (be * be) * P.w[3]);
            // compute_least_squares_endpoints_a, bc7enc.cpp:410-427 (accumulated in the same texel order)
            z00 = fadd(z00, T.w2x[s][0]);
            z10 = fadd(z10, T.w2x[s][1]);
            z11 = fadd(z11, T.w2x[s][2]);
            q00 = fadd(q00, fmul(T.w2x[s][3], (float) a));
            t = fadd(t, (float) a);
        }
        if(terr < best)
        {
            best = terr;
            out_lo = lo_a, out_hi = hi_a, out_sel = tsel;
        }
        if(pass != passes - 1u)
        {
            const float q10 = fsub(t, q00);
            const float z01 = z10;
            float det = fsub(fmul(z00, z11), fmul(z01, z10));
            if(det != 0.0f) { det = fdiv(1.0f, det); }
            const float iz00 = fmul(z11, det), iz01 = fmul(-z01, det), iz10 = fmul(-z10, det), iz11 = fmul(z00, det);
            const float xl = fadd(fmul(iz00, q00), fmul(iz01, q10));
            const float xh = fadd(fmul(iz10, q00), fmul(iz11, q10));
            // bc7enc.cpp:445-459 only acts when every alpha is equal, which cannot happen here (min_a != max_a on entry)
            // (int)floor(x + .5f) with x86 cvttss2si semantics for out-of-range values (-> INT_MIN), then clamp
            const float fl = floorf(fadd(xl, .5f)), fh = floorf(fadd(xh, .5f));
            const int il = (fl >= -2147483648.0f && fl < 2147483648.0f) ? f2i(fl) : (int) 0x80000000;
            const int ih = (fh >= -2147483648.0f && fh < 2147483648.0f) ? f2i(fh) : (int) 0x80000000;
            uint32_t nlo = (uint32_t) clampi(il, 0, 255), nhi = (uint32_t) clampi(ih, 0, 255);
            if(nlo > nhi)
            {
                const uint32_t tt = nlo;
                nlo = nhi, nhi = tt;
            }
            if((nlo == lo_a) && (nhi == hi_a)) { break; }
            lo_a = nlo, hi_a = nhi;
        }
    }
    return best;
}

// ---------------------------------------------------------------------------------------------------- block entry
// Fill the YCbCr part of the lane column from its 16 texels (hoisted once per block).
template<int STRIDE>
VKT_FN void prepare_lane(Lane<STRIDE> L)
{
#pragma unroll
    for(int i = 0; i < 16; ++i)
    {
        Texel t;
        t.px = L.p[i * STRIDE].px;
        const Ycc y = to_ycc_packed(t.px);
        t.l = y.l, t.cr = y.cr, t.cb = y.cb;
        L.p[i * STRIDE] = t;
    }
}

// The dispatch of bc7enc_compress_block (bc7enc.cpp:2422-2437): does the block take the alpha path?
VKT_FN bool block_has_alpha(const Bc7KernelParams &P, const uint32_t px[16])
{
    uint32_t and_all = 0xFFFFFFFFu;
#pragma unroll
    for(int i = 0; i < 16; ++i) { and_all &= px[i]; }
    return P.force_alpha || ((and_all >> 24) != 255u);
}

// bc7enc_compress_block (bc7enc.cpp:2402-2438) after its dispatch: ALPHA == false is handle_opaque_block (:2293-2400),
// ALPHA == true is handle_alpha_block (:2139-2291).  The caller classifies blocks first so that a warp only ever holds
// blocks of one kind; the whole warp must call this converged (estimate_partition is warp-cooperative).
// L: the lane column with texels [0,16) filled; the YCbCr part is filled here.
// gid: an identifier of the lane's block that the caller can map back to its output slot.  Returns the identifier of the
// block this lane actually encoded -- on the device the CTA's opaque blocks change hands once (see below).
template<bool PERC, int KV, bool ALPHA, bool UBER, int STRIDE>
VKT_FN uint32_t encode_block(const Bc7Tables &T, const Bc7KernelParams &P, Lane<STRIDE> L, uint32_t out[4], uint32_t gid = 0)
{
    if(PERC) { prepare_lane<STRIDE>(L); }
    const CellRef whole = {kIdentityPerm, 16};

    BlockSolution sol;
    sol.mode = 6, sol.partition = 0, sol.sel = 0, sol.asel = 0;
    sol.lo[0] = sol.lo[1] = sol.hi[0] = sol.hi[1] = 0;
    sol.pbits[0] = sol.pbits[1] = 0;
    uint64_t best_err = kNoErr;

    // The partition estimate (for mode 1 / mode 7) depends on the block alone.  Opaque blocks run it FIRST: nothing of the
    // other modes is live during its loops (7 registers more for them at the 80-register budget; -2.7 % kernel time).  The
    // reference runs it after mode 6 and skips it for blocks mode 6 already encodes exactly; here those few blocks estimate
    // in vain.  Alpha blocks keep the reference's order (modes 6 and 5 often reach zero error on transparent blocks, and
    // the skipped estimates are worth more than the registers: measured).
    const bool do17 = ALPHA ? ((P.mode_mask & (1u << 7)) != 0) : ((P.max_partitions > 0) && (P.mode_mask & (1u << 1)));// warp-uniform
    uint32_t part17 = 0;
    if(!ALPHA && do17) { part17 = estimate_partition<false, PERC, KV, STRIDE>(T, P, L, true); }
#if defined(__CUDA_ARCH__) && !defined(VKT_NO_SIZE_REGROUP)
    // Second regrouping, by subset size.  The colour-cell searches of mode 1 loop over the texels of a subset, and the lanes
    // of a warp hold partitions with different subset sizes: a raster-order warp runs max(8..15) + max(1..8) = 22 trips where
    // one block needs 16 (20.4 of 32 lanes active in those loops).  Now that every block has its partition, the CTA's blocks
    // are sorted by the size of their larger subset (8 classes) with the same counting sort as in estimate_partition -- over the
    // whole CTA: pure warps beat the bank pattern here as well, 1.984 -> 1.960 ms, uber 4 2.135 -> 2.08 ms -- and every lane KEEPS
    // the block it adopts to the end of the kernel: what travels is the column pointer, the partition and the block's identifier (the
    // caller stores the result in that block's slot).  Which lane encodes a block has no influence on the block's result.
    if(!ALPHA && do17 && (STRIDE >= 64) && (STRIDE % 32 == 0))// CTA-uniform
    {
        CtaScratch<STRIDE> *S = reinterpret_cast<CtaScratch<STRIDE> *>(L.p - threadIdx.x + 16 * STRIDE);
        const uint32_t n0 = T.est_n0[part17];
        const uint32_t bin = umax(n0, 16u - n0) - 8u;// 0..7
#if !defined(VKT_REGROUP_BY_CLASS)
        const uint32_t cls = 0u;
#else
        const uint32_t cls = threadIdx.x & 7u;
#endif
        const uint32_t pos = atomicAdd(&S->cnt2[cls][bin], 1u);
        S->err[threadIdx.x] = gid;
        S->info[threadIdx.x] = (uint16_t) part17;
        __syncthreads();
        uint32_t q = pos;
#pragma unroll
        for(uint32_t b = 0; b < 7; ++b) { q += (b < bin) ? S->cnt2[cls][b] : 0u; }
#if !defined(VKT_REGROUP_BY_CLASS)
        S->perm[q] = (uint16_t) threadIdx.x;
#else
        S->perm[(q / 4u) * 32u + (q % 4u) * 8u + cls] = (uint16_t) threadIdx.x;
#endif
        __syncthreads();
        const uint32_t src = S->perm[threadIdx.x];
        gid = S->err[src];
        part17 = S->info[src];
        L.p = L.p + ((int) src - (int) threadIdx.x);
    }
#endif

    if(P.mode_mask & (1u << 6))
    {
        Cell c6;
        best_err = weigh(compress_cell<6, ALPHA, PERC, KV, UBER, STRIDE>(T, P, L, whole, c6), P.mode6_w);
        sol.sel = c6.sel, sol.lo[0] = c6.lo, sol.hi[0] = c6.hi, sol.pbits[0] = c6.pbits;
    }
    if(!ALPHA)
    {
        if(do17)
        {
            BlockSolution s1 = sol;
            if(two_subset_cells<1, PERC, KV, UBER, STRIDE>(T, P, L, part17, best_err > 0, best_err, s1) != kNoErr) { sol = s1; }
        }
    }
    else
    {
        if((best_err > 0) && (P.mode_mask & (1u << 5)))
        {
            uint32_t min_a = 255, max_a = 0;
            for(int i = 0; i < 16; ++i)
            {
                const uint32_t a = L.px(i) >> 24;
                min_a = umin(min_a, a), max_a = umax(max_a, a);
            }
            Cell c5;
            uint64_t e5 = compress_cell<5, false, PERC, KV, UBER, STRIDE>(T, P, L, whole, c5);
            uint32_t alo = 0, ahi = 0;
            uint64_t asel = 0;
            e5 += mode5_alpha<STRIDE>(T, P, L, min_a, max_a, alo, ahi, asel);
            e5 = weigh(e5, P.mode5_w);
            if(e5 < best_err)
            {
                best_err = e5;
                sol.mode = 5, sol.partition = 0;
                sol.sel = c5.sel, sol.asel = asel;
                sol.lo[0] = (c5.lo & 0x00FFFFFFu) | (alo << 24);
                sol.hi[0] = (c5.hi & 0x00FFFFFFu) | (ahi << 24);
                sol.pbits[0] = 0;
            }
        }
        if(do17)
        {
            part17 = estimate_partition<true, PERC, KV, STRIDE>(T, P, L, best_err > 0);
            BlockSolution s7 = sol;
            if(two_subset_cells<7, PERC, KV, UBER, STRIDE>(T, P, L, part17, best_err > 0, best_err, s7) != kNoErr) { sol = s7; }
        }
    }
    pack_block(T, sol, out);
    return gid;
}

}// namespace vkt
