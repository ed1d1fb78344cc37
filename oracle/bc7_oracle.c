/* oracle/bc7_oracle.c -- TEST INFRASTRUCTURE ONLY (see bc7_oracle.h).
 *
 * Scalar C restatement of bc7enc_rdo's bc7enc_compress_block (the function vierkant::bcn::compress calls per block,
 * /root/reference/src/texture_block_compression.cpp:132).  "bc7enc.cpp:N" below = /root/reference/extern/bc7enc_rdo/bc7enc.cpp.
 *
 * The arithmetic (operation order, float32 rounding points, truncating conversions, strict '<' tie-breaks, early-outs)
 * follows the reference exactly; the code is organised differently (packed partition masks, one least-squares routine
 * for RGB/RGBA, one quantiser helper).  Must be compiled with -ffp-contract=off (oracle/Makefile) so no FMA is formed.
 */
#include "bc7_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint8_t c[4]; } rgba8;
typedef struct { float c[4]; } vec4;

/* ------------------------------------------------------------------------------------------------ tables */
/* interpolation weights, bc7enc.cpp:48-50 (BC7 specification) */
static const uint32_t W2[4] = {0, 21, 43, 64};
static const uint32_t W3[8] = {0, 9, 18, 27, 37, 46, 55, 64};
static const uint32_t W4[16] = {0, 4, 9, 13, 17, 21, 26, 30, 34, 38, 43, 47, 51, 55, 60, 64};
/* least-squares weight tuples {w*w, (1-w)*w, (1-w)*(1-w), w} as the 6-digit decimal literals of bc7enc.cpp:52-57.
 * They are data, not formulas: (21/64)^2 evaluated in float is a different number. */
static const float W2X[4][4] = {{0.000000f, 0.000000f, 1.000000f, 0.000000f},
                                {0.107666f, 0.220459f, 0.451416f, 0.328125f},
                                {0.451416f, 0.220459f, 0.107666f, 0.671875f},
                                {1.000000f, 0.000000f, 0.000000f, 1.000000f}};
static const float W3X[8][4] = {{0.000000f, 0.000000f, 1.000000f, 0.000000f}, {0.019775f, 0.120850f, 0.738525f, 0.140625f},
                                {0.079102f, 0.202148f, 0.516602f, 0.281250f}, {0.177979f, 0.243896f, 0.334229f, 0.421875f},
                                {0.334229f, 0.243896f, 0.177979f, 0.578125f}, {0.516602f, 0.202148f, 0.079102f, 0.718750f},
                                {0.738525f, 0.120850f, 0.019775f, 0.859375f}, {1.000000f, 0.000000f, 0.000000f, 1.000000f}};
static const float W4X[16][4] = {
        {0.000000f, 0.000000f, 1.000000f, 0.000000f}, {0.003906f, 0.058594f, 0.878906f, 0.062500f},
        {0.019775f, 0.120850f, 0.738525f, 0.140625f}, {0.041260f, 0.161865f, 0.635010f, 0.203125f},
        {0.070557f, 0.195068f, 0.539307f, 0.265625f}, {0.107666f, 0.220459f, 0.451416f, 0.328125f},
        {0.165039f, 0.241211f, 0.352539f, 0.406250f}, {0.219727f, 0.249023f, 0.282227f, 0.468750f},
        {0.282227f, 0.249023f, 0.219727f, 0.531250f}, {0.352539f, 0.241211f, 0.165039f, 0.593750f},
        {0.451416f, 0.220459f, 0.107666f, 0.671875f}, {0.539307f, 0.195068f, 0.070557f, 0.734375f},
        {0.635010f, 0.161865f, 0.041260f, 0.796875f}, {0.738525f, 0.120850f, 0.019775f, 0.859375f},
        {0.878906f, 0.058594f, 0.003906f, 0.937500f}, {1.000000f, 0.000000f, 0.000000f, 1.000000f}};

/* BC7 two-subset partitions as bit masks: bit i = subset of texel i (packed form of bc7enc.cpp:60-70). */
static const uint16_t PART2[64] = {
        0xCCCC, 0x8888, 0xEEEE, 0xECC8, 0xC880, 0xFEEC, 0xFEC8, 0xEC80, 0xC800, 0xFFEC, 0xFE80, 0xE800, 0xFFE8,
        0xFF00, 0xFFF0, 0xF000, 0xF710, 0x008E, 0x7100, 0x08CE, 0x008C, 0x7310, 0x3100, 0x8CCE, 0x088C, 0x3110,
        0x6666, 0x366C, 0x17E8, 0x0FF0, 0x718E, 0x399C, 0xAAAA, 0xF0F0, 0x5A5A, 0x33CC, 0x3C3C, 0x55AA, 0x9696,
        0xA55A, 0x73CE, 0x13C8, 0x324C, 0x3BDC, 0x6996, 0xC33C, 0x9966, 0x0660, 0x0272, 0x04E4, 0x4E40, 0x2720,
        0xC936, 0x936C, 0x39C6, 0x639C, 0x9336, 0x9CC6, 0x817E, 0xE718, 0xCCF0, 0x0FCC, 0x7744, 0xEE22};
/* anchor texel of the second subset, bc7enc.cpp:94 (BC7 specification) */
static const uint8_t ANCHOR2[64] = {15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 15, 2,  8,  2,  2, 8,
                                    8,  15, 2,  8,  2,  2,  8,  8,  2,  2,  15, 15, 6,  8,  2,  8,  15, 15, 2,  8,  2, 2,
                                    2,  15, 15, 6,  6,  2,  6,  8,  15, 15, 2,  2,  15, 15, 15, 15, 15, 2,  2,  15};
/* partition scan order of estimate_partition, bc7enc.cpp:1765-1775 (1-based in the reference, 0-based here) */
static const uint8_t PART_ORDER[64] = {0,  13, 1,  2,  15, 14, 10, 16, 3,  23, 26, 6,  7,  21, 19, 29,
                                       8,  4,  9,  20, 5,  31, 22, 17, 18, 11, 12, 30, 24, 25, 28, 27,
                                       32, 33, 34, 45, 46, 51, 49, 50, 48, 38, 39, 37, 53, 52, 54, 36,
                                       57, 58, 55, 41, 40, 42, 43, 59, 44, 56, 47, 35, 60, 63, 62, 61};
/* filterbank predictors, bc7enc.cpp:1714-1751: bit (k+1) set = evaluate this partition when key partition k won */
#define B(x) (1u << (x))
static const uint32_t PART_PRED[35] = {
        ~0u, ~0u, ~0u, ~0u, ~0u,
        B(1) | B(2) | B(8), B(1) | B(3) | B(7), ~0u, ~0u, B(2) | B(8) | B(16), B(7) | B(3) | B(15), ~0u,
        B(8) | B(14) | B(16), B(7) | B(14) | B(15), ~0u, ~0u, ~0u, ~0u,
        B(14) | B(15), B(16) | B(22) | B(14), B(17) | B(24) | B(14), B(2) | B(14) | B(15) | B(1), ~0u,
        B(1) | B(3) | B(14) | B(16) | B(22), ~0u, B(1) | B(2) | B(15) | B(17) | B(24), B(1) | B(3) | B(22), ~0u, ~0u, ~0u,
        B(14) | B(15) | B(16) | B(17), ~0u, ~0u,
        B(1) | B(2) | B(3) | B(27) | B(4) | B(24), B(14) | B(15) | B(16) | B(11) | B(17) | B(27)};
#undef B

typedef struct { uint16_t err; uint8_t lo, hi; } opt_ep;
static opt_ep g_opt1[256][2];    /* [colour][pbit]         bc7enc.cpp:213-240 */
static opt_ep g_opt7[256][2][2]; /* [colour][hi_p][lo_p]   bc7enc.cpp:242-282 */
static float g_mid1[64][2];      /* bc7enc.cpp:150-169 */
static float g_mid5[128];        /* bc7enc.cpp:171-186 */
static float g_mid7[32][2];      /* bc7enc.cpp:129-148 */
static uint8_t g_m6_reduced[2048][2]; /* bc7enc.cpp:188-211 (only with quant_mode6_endpoints) */
static pthread_once_t g_once = PTHREAD_ONCE_INIT;

static int imin(int a, int b) { return a < b ? a : b; }

/* bc7enc.cpp:124-285 */
static void build_tables(void)
{
    for(uint32_t p = 0; p < 2; p++)
    {
        for(uint32_t i = 0; i < 32; i++)
        {
            uint32_t vl = ((i << 1) | p) << 2;
            vl |= vl >> 6;
            uint32_t vh = (((uint32_t) imin(31, (int) i + 1) << 1) | p) << 2;
            vh |= vh >> 6;
            float lo = vl / 255.0f, hi = vh / 255.0f;
            g_mid7[i][p] = (i == 31) ? 1.0f : (lo + hi) / 2.0f;
        }
        for(uint32_t i = 0; i < 64; i++)
        {
            uint32_t vl = ((i << 1) | p) << 1;
            vl |= vl >> 7;
            uint32_t vh = (((uint32_t) imin(63, (int) i + 1) << 1) | p) << 1;
            vh |= vh >> 7;
            float lo = vl / 255.0f, hi = vh / 255.0f;
            g_mid1[i][p] = (i == 63) ? 1.0f : (lo + hi) / 2.0f;
        }
    }
    for(uint32_t i = 0; i < 128; i++)
    {
        uint32_t vl = i << 1;
        vl |= vl >> 7;
        uint32_t vh = (uint32_t) imin(127, (int) i + 1) << 1;
        vh |= vh >> 7;
        float lo = vl / 255.0f, hi = vh / 255.0f;
        g_mid5[i] = (i == 127) ? 1.0f : (lo + hi) / 2.0f;
    }
    for(uint32_t p = 0; p < 2; p++)
    {
        for(uint32_t i = 0; i < 2048; i++)
        {
            float f = i / 2047.0f, best = 1e+9f;
            int best_index = 0;
            for(int j = 0; j < 64; j++)
            {
                int ik = (j * 127 + 31) / 63;
                float k = ((ik << 1) + p) / 255.0f;
                float e = fabsf(k - f);
                if(e < best) { best = e; best_index = ik; }
            }
            g_m6_reduced[i][p] = (uint8_t) best_index;
        }
    }
    for(int c = 0; c < 256; c++)
    {
        for(uint32_t lp = 0; lp < 2; lp++)
        {
            opt_ep best = {0xFFFF, 0, 0}; /* lo/hi are always overwritten: err 65025 max < 65535 */
            for(uint32_t l = 0; l < 64; l++)
            {
                uint32_t low = ((l << 1) | lp) << 1;
                low |= low >> 7;
                for(uint32_t h = 0; h < 64; h++)
                {
                    uint32_t high = ((h << 1) | lp) << 1;
                    high |= high >> 7;
                    int k = (int) ((low * (64 - W3[2]) + high * W3[2] + 32) >> 6);
                    int err = (k - c) * (k - c);
                    if(err < best.err) { best.err = (uint16_t) err; best.lo = (uint8_t) l; best.hi = (uint8_t) h; }
                }
            }
            g_opt1[c][lp] = best;
        }
        for(uint32_t hp = 0; hp < 2; hp++)
        {
            for(uint32_t lp = 0; lp < 2; lp++)
            {
                opt_ep best = {0xFFFF, 0, 0};
                for(uint32_t l = 0; l < 32; l++)
                {
                    uint32_t low = ((l << 1) | lp) << 2;
                    low |= low >> 6;
                    for(uint32_t h = 0; h < 32; h++)
                    {
                        uint32_t high = ((h << 1) | hp) << 2;
                        high |= high >> 6;
                        int k = (int) ((low * (64 - W2[1]) + high * W2[1] + 32) >> 6);
                        int err = (k - c) * (k - c);
                        if(err < best.err) { best.err = (uint16_t) err; best.lo = (uint8_t) l; best.hi = (uint8_t) h; }
                    }
                }
                g_opt7[c][hp][lp] = best;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------ helpers */
static float satf(float v) { if(v < 0.0f) v = 0.0f; else if(v > 1.0f) v = 1.0f; return v; }        /* bc7enc.cpp:12-13 */
static int clampi(int v, int lo, int hi) { if(v < lo) v = lo; else if(v > hi) v = hi; return v; }   /* bc7enc.cpp:11 */
static float sqf(float v) { return v * v; }
static float dot4(const vec4 *a, const vec4 *b)                                                     /* bc7enc.cpp:43 */
{
    return a->c[0] * b->c[0] + a->c[1] * b->c[1] + a->c[2] * b->c[2] + a->c[3] * b->c[3];
}
static void normalize4(vec4 *v)                                                                     /* bc7enc.cpp:45 */
{
    float s = v->c[0] * v->c[0] + v->c[1] * v->c[1] + v->c[2] * v->c[2] + v->c[3] * v->c[3];
    if(s != 0.0f)
    {
        s = 1.0f / sqrtf(s);
        for(int i = 0; i < 4; i++) v->c[i] *= s;
    }
}

/* the (uint64)(err * weight + .5f) round trip of bc7enc.cpp:2167,2184,2234,2239,2326,2377,2382 */
static uint64_t weigh(uint64_t err, float w) { return (uint64_t) ((float) err * w + .5f); }

/* per-cell configuration / state, bc7enc.cpp:462-485 */
typedef struct
{
    uint32_t n;
    const rgba8 *px;
    uint32_t nsel;
    const uint32_t *w;
    const float (*wx)[4];
    uint32_t comp_bits;
    uint32_t weights[4];
    int has_alpha, has_pbits, share_pbit, perceptual;
} cell_cfg;

typedef struct
{
    uint64_t best_err;
    rgba8 lo, hi;
    uint32_t pbits[2];
    uint8_t *sel;
    uint8_t *sel_tmp;
} cell_out;

/* bc7enc.cpp:487-503 */
static rgba8 expand_endpoint(const rgba8 *q, const cell_cfg *cfg)
{
    const uint32_t n = cfg->comp_bits + (cfg->has_pbits ? 1 : 0);
    rgba8 r;
    for(int i = 0; i < 4; i++)
    {
        uint32_t v = (uint32_t) q->c[i] << (8 - n);
        v |= v >> n;
        r.c[i] = (uint8_t) v;
    }
    return r;
}

/* bc7enc.cpp:505-535.  Argument order matters: (candidate, source); >> is arithmetic on negative ints. */
static uint64_t dist_rgb(const rgba8 *e1, const rgba8 *e2, int perceptual, const uint32_t w[4])
{
    int dr, dg, db;
    if(perceptual)
    {
        const int l1 = e1->c[0] * 109 + e1->c[1] * 366 + e1->c[2] * 37;
        const int cr1 = ((int) e1->c[0] << 9) - l1, cb1 = ((int) e1->c[2] << 9) - l1;
        const int l2 = e2->c[0] * 109 + e2->c[1] * 366 + e2->c[2] * 37;
        const int cr2 = ((int) e2->c[0] << 9) - l2, cb2 = ((int) e2->c[2] << 9) - l2;
        dr = (l1 - l2) >> 8;
        dg = (cr1 - cr2) >> 8;
        db = (cb1 - cb2) >> 8;
    }
    else
    {
        dr = (int) e1->c[0] - (int) e2->c[0];
        dg = (int) e1->c[1] - (int) e2->c[1];
        db = (int) e1->c[2] - (int) e2->c[2];
    }
    return w[0] * (uint32_t) (dr * dr) + w[1] * (uint32_t) (dg * dg) + w[2] * (uint32_t) (db * db); /* u32 wrap */
}
static uint64_t dist_rgba(const rgba8 *e1, const rgba8 *e2, int perceptual, const uint32_t w[4])
{
    int da = (int) e1->c[3] - (int) e2->c[3];
    return dist_rgb(e1, e2, perceptual, w) + (w[3] * (uint32_t) (da * da));
}

/* bc7enc.cpp:537-585 */
static uint64_t solid_mode1(const cell_cfg *cfg, cell_out *out, uint32_t r, uint32_t g, uint32_t b, uint8_t *sel)
{
    uint32_t best_err = UINT32_MAX, best_p = 0;
    for(uint32_t p = 0; p < 2; p++)
    {
        uint32_t err = g_opt1[r][p].err + g_opt1[g][p].err + g_opt1[b][p].err;
        if(err < best_err)
        {
            best_err = err;
            best_p = p;
            if(!best_err) break;
        }
    }
    const opt_ep *er = &g_opt1[r][best_p], *eg = &g_opt1[g][best_p], *eb = &g_opt1[b][best_p];
    out->lo = (rgba8){{er->lo, eg->lo, eb->lo, 0}};
    out->hi = (rgba8){{er->hi, eg->hi, eb->hi, 0}};
    out->pbits[0] = best_p;
    out->pbits[1] = 0;
    memset(sel, 2, cfg->n);
    rgba8 p;
    for(int i = 0; i < 3; i++)
    {
        uint32_t low = (((uint32_t) out->lo.c[i] << 1) | best_p) << 1;
        low |= low >> 7;
        uint32_t high = (((uint32_t) out->hi.c[i] << 1) | best_p) << 1;
        high |= high >> 7;
        p.c[i] = (uint8_t) ((low * (64 - W3[2]) + high * W3[2] + 32) >> 6);
    }
    p.c[3] = 255;
    uint64_t total = 0;
    for(uint32_t i = 0; i < cfg->n; i++) total += dist_rgb(&p, &cfg->px[i], cfg->perceptual, cfg->weights);
    out->best_err = total;
    return total;
}

/* bc7enc.cpp:587-643 */
static uint64_t solid_mode7(const cell_cfg *cfg, cell_out *out, uint32_t r, uint32_t g, uint32_t b, uint32_t a,
                            uint8_t *sel)
{
    uint32_t best_err = UINT32_MAX, best_p = 0;
    for(uint32_t p = 0; p < 4; p++)
    {
        uint32_t hp = p >> 1, lp = p & 1;
        uint32_t err = g_opt7[r][hp][lp].err + g_opt7[g][hp][lp].err + g_opt7[b][hp][lp].err + g_opt7[a][hp][lp].err;
        if(err < best_err)
        {
            best_err = err;
            best_p = p;
            if(!best_err) break;
        }
    }
    uint32_t hp = best_p >> 1, lp = best_p & 1;
    const opt_ep *e[4] = {&g_opt7[r][hp][lp], &g_opt7[g][hp][lp], &g_opt7[b][hp][lp], &g_opt7[a][hp][lp]};
    for(int i = 0; i < 4; i++)
    {
        out->lo.c[i] = e[i]->lo;
        out->hi.c[i] = e[i]->hi;
    }
    out->pbits[0] = lp;
    out->pbits[1] = hp;
    for(uint32_t i = 0; i < cfg->n; i++) sel[i] = 1;
    rgba8 p;
    for(int i = 0; i < 4; i++)
    {
        uint32_t low = ((uint32_t) out->lo.c[i] << 1) | lp;
        uint32_t high = ((uint32_t) out->hi.c[i] << 1) | hp;
        low = (low << 2) | (low >> 6);
        high = (high << 2) | (high >> 6);
        p.c[i] = (uint8_t) ((low * (64 - W2[1]) + high * W2[1] + 32) >> 6);
    }
    uint64_t total = 0;
    for(uint32_t i = 0; i < cfg->n; i++) total += dist_rgba(&p, &cfg->px[i], cfg->perceptual, cfg->weights);
    out->best_err = total;
    return total;
}

/* bc7enc.cpp:645-831 */
static uint64_t try_endpoints(const rgba8 *lo, const rgba8 *hi, const uint32_t pbits[2], const cell_cfg *cfg,
                              cell_out *out, const port_bc7_params *cp)
{
    rgba8 qlo = *lo, qhi = *hi;
    if(cfg->has_pbits)
    {
        uint32_t pl = pbits[0], ph = cfg->share_pbit ? pbits[0] : pbits[1];
        for(int i = 0; i < 4; i++)
        {
            qlo.c[i] = (uint8_t) ((lo->c[i] << 1) | pl);
            qhi.c[i] = (uint8_t) ((hi->c[i] << 1) | ph);
        }
    }
    const rgba8 c0 = expand_endpoint(&qlo, cfg), c1 = expand_endpoint(&qhi, cfg);
    const uint32_t N = cfg->nsel;
    rgba8 pal[16];
    pal[0] = c0;
    pal[N - 1] = c1;
    const uint32_t nc = cfg->has_alpha ? 4 : 3;
    for(uint32_t i = 1; i < N - 1; i++)
        for(uint32_t j = 0; j < nc; j++)
            pal[i].c[j] = (uint8_t) ((c0.c[j] * (64 - cfg->w[i]) + c1.c[j] * cfg->w[i] + 32) >> 6);

    const int lr = c0.c[0], lg = c0.c[1], lb = c0.c[2];
    const int dr = c1.c[0] - lr, dg = c1.c[1] - lg, db = c1.c[2] - lb;
    uint64_t total = 0;

    if(cp->force_selectors)
    {
        for(uint32_t i = 0; i < cfg->n; i++)
        {
            const uint32_t s = cp->selectors[i];
            total += cfg->has_alpha ? dist_rgba(&pal[s], &cfg->px[i], cfg->perceptual, cfg->weights)
                                    : dist_rgb(&pal[s], &cfg->px[i], cfg->perceptual, cfg->weights);
            out->sel_tmp[i] = (uint8_t) s;
        }
    }
    else if(!cfg->perceptual)
    {
        if(cfg->has_alpha)
        {
            const int la = c0.c[3], da = c1.c[3] - la;
            const float f = N / (float) (dr * dr + dg * dg + db * db + da * da + .00000125f);
            for(uint32_t i = 0; i < cfg->n; i++)
            {
                const rgba8 *c = &cfg->px[i];
                int r = c->c[0], g = c->c[1], b = c->c[2], a = c->c[3];
                int s = (int) ((float) ((r - lr) * dr + (g - lg) * dg + (b - lb) * db + (a - la) * da) * f + .5f);
                s = clampi(s, 1, (int) N - 1);
                uint64_t e0 = dist_rgba(&pal[s - 1], c, 0, cfg->weights);
                uint64_t e1 = dist_rgba(&pal[s], c, 0, cfg->weights);
                if(e1 > e0) { e1 = e0; --s; }
                total += e1;
                out->sel_tmp[i] = (uint8_t) s;
            }
        }
        else
        {
            const float f = N / (float) (dr * dr + dg * dg + db * db + .00000125f);
            for(uint32_t i = 0; i < cfg->n; i++)
            {
                const rgba8 *c = &cfg->px[i];
                int r = c->c[0], g = c->c[1], b = c->c[2];
                int s = (int) ((float) ((r - lr) * dr + (g - lg) * dg + (b - lb) * db) * f + .5f);
                s = clampi(s, 1, (int) N - 1);
                uint64_t e0 = dist_rgb(&pal[s - 1], c, 0, cfg->weights);
                uint64_t e1 = dist_rgb(&pal[s], c, 0, cfg->weights);
                int bs = s;
                uint64_t be = e1;
                if(e0 < be) { be = e0; bs = s - 1; }
                total += be;
                out->sel_tmp[i] = (uint8_t) bs;
            }
        }
    }
    else
    {
        for(uint32_t i = 0; i < cfg->n; i++)
        {
            uint64_t be = UINT64_MAX;
            uint32_t bs = 0;
            for(uint32_t j = 0; j < N; j++)
            {
                uint64_t e = cfg->has_alpha ? dist_rgba(&pal[j], &cfg->px[i], 1, cfg->weights)
                                            : dist_rgb(&pal[j], &cfg->px[i], 1, cfg->weights);
                if(e < be) { be = e; bs = j; }
            }
            total += be;
            out->sel_tmp[i] = (uint8_t) bs;
        }
    }

    if(total < out->best_err)
    {
        out->best_err = total;
        out->lo = *lo;
        out->hi = *hi;
        out->pbits[0] = pbits[0];
        out->pbits[1] = pbits[1];
        memcpy(out->sel, out->sel_tmp, cfg->n);
    }
    return total;
}

/* bc7enc.cpp:833-866 */
static void nudge_degenerate(uint32_t mode, rgba8 *lo, rgba8 *hi, const vec4 *xl, const vec4 *xh, uint32_t iscale,
                             const port_bc7_params *cp)
{
    if(!((mode == 1) || ((mode == 6) && cp->quant_mode6_endpoints))) return;
    for(int i = 0; i < 3; i++)
    {
        if(lo->c[i] != hi->c[i]) continue;
        if(!(fabs(xl->c[i] - xh->c[i]) > 0.0f)) continue;
        if(lo->c[i] > (iscale >> 1))
        {
            if(lo->c[i] > 0) lo->c[i]--;
            else if(hi->c[i] < iscale) hi->c[i]++;
        }
        else
        {
            if(hi->c[i] < iscale) hi->c[i]++;
            else if(lo->c[i] > 0) lo->c[i]--;
        }
    }
}

static int same_rgba(const rgba8 *a, const rgba8 *b) { return memcmp(a->c, b->c, 4) == 0; }

/* bc7enc.cpp:868-1099 */
static uint64_t fit_endpoints(uint32_t mode, vec4 xl, vec4 xh, const cell_cfg *cfg, cell_out *out,
                              const port_bc7_params *cp)
{
    for(int i = 0; i < 4; i++)
    {
        xl.c[i] = satf(xl.c[i]);
        xh.c[i] = satf(xh.c[i]);
    }
    if(cfg->has_pbits)
    {
        const int iscalep = (1 << (cfg->comp_bits + 1)) - 1;
        const float scalep = (float) iscalep;
        const int ncomp = cfg->has_alpha ? 4 : 3;
        uint32_t best_pbits[2];
        rgba8 best_lo, best_hi;

        if(!cfg->share_pbit)
        {
            if((cfg->comp_bits == 7) && cp->quant_mode6_endpoints)
            {
                best_pbits[0] = 0;
                best_pbits[1] = 1;
                for(int c = 0; c < 4; c++)
                {
                    best_lo.c[c] = g_m6_reduced[(int) ((xl.c[c] * 2047.0f) + .5f)][0];
                    best_hi.c[c] = g_m6_reduced[(int) ((xh.c[c] * 2047.0f) + .5f)][1];
                }
            }
            else
            {
                float best_e0 = 1e+9f, best_e1 = 1e+9f;
                for(int p = 0; p < 2; p++)
                {
                    rgba8 qlo, qhi;
                    if(cfg->comp_bits == 5)
                    {
                        for(int c = 0; c < 4; c++)
                        {
                            int vl = (int) (xl.c[c] * 31.0f);
                            vl += (xl.c[c] > g_mid7[vl][p]);
                            qlo.c[c] = (uint8_t) clampi(vl * 2 + p, p, 63 - 1 + p);
                            int vh = (int) (xh.c[c] * 31.0f);
                            vh += (xh.c[c] > g_mid7[vh][p]);
                            qhi.c[c] = (uint8_t) clampi(vh * 2 + p, p, 63 - 1 + p);
                        }
                    }
                    else
                    {
                        for(int c = 0; c < 4; c++)
                        {
                            qlo.c[c] = (uint8_t) clampi(((int) ((xl.c[c] * scalep - p) / 2.0f + .5f)) * 2 + p, p, iscalep - 1 + p);
                            qhi.c[c] = (uint8_t) clampi(((int) ((xh.c[c] * scalep - p) / 2.0f + .5f)) * 2 + p, p, iscalep - 1 + p);
                        }
                    }
                    rgba8 slo = expand_endpoint(&qlo, cfg), shi = expand_endpoint(&qhi, cfg);
                    float e0 = 0, e1 = 0;
                    for(int i = 0; i < ncomp; i++)
                    {
                        e0 += sqf(slo.c[i] - xl.c[i] * 255.0f);
                        e1 += sqf(shi.c[i] - xh.c[i] * 255.0f);
                    }
                    if(p == 1)
                    {
                        e0 *= cp->pbit1_weight;
                        e1 *= cp->pbit1_weight;
                    }
                    if(e0 < best_e0)
                    {
                        best_e0 = e0;
                        best_pbits[0] = (uint32_t) p;
                        for(int c = 0; c < 4; c++) best_lo.c[c] = qlo.c[c] >> 1;
                    }
                    if(e1 < best_e1)
                    {
                        best_e1 = e1;
                        best_pbits[1] = (uint32_t) p;
                        for(int c = 0; c < 4; c++) best_hi.c[c] = qhi.c[c] >> 1;
                    }
                }
            }
        }
        else if((mode == 1) && cp->bias_mode1_pbits)
        {
            float x = 0.0f;
            for(int c = 0; c < 3; c++)
            {
                float t = x > xl.c[c] ? x : xl.c[c]; /* std::max(x, xl) then std::max(.., xh), bc7enc.cpp:981 */
                x = t > xh.c[c] ? t : xh.c[c];
            }
            int p = (x > (253.0f / 255.0f)) ? 1 : 0;
            rgba8 qlo, qhi;
            for(int c = 0; c < 4; c++)
            {
                int vl = (int) (xl.c[c] * 63.0f);
                vl += (xl.c[c] > g_mid1[vl][p]);
                qlo.c[c] = (uint8_t) clampi(vl * 2 + p, p, 127 - 1 + p);
                int vh = (int) (xh.c[c] * 63.0f);
                vh += (xh.c[c] > g_mid1[vh][p]);
                qhi.c[c] = (uint8_t) clampi(vh * 2 + p, p, 127 - 1 + p);
            }
            best_pbits[0] = best_pbits[1] = (uint32_t) p;
            for(int c = 0; c < 4; c++)
            {
                best_lo.c[c] = qlo.c[c] >> 1;
                best_hi.c[c] = qhi.c[c] >> 1;
            }
        }
        else
        {
            float best_e = 1e+9f;
            for(int p = 0; p < 2; p++)
            {
                rgba8 qlo, qhi;
                if(cfg->comp_bits == 6)
                {
                    for(int c = 0; c < 4; c++)
                    {
                        int vl = (int) (xl.c[c] * 63.0f);
                        vl += (xl.c[c] > g_mid1[vl][p]);
                        qlo.c[c] = (uint8_t) clampi(vl * 2 + p, p, 127 - 1 + p);
                        int vh = (int) (xh.c[c] * 63.0f);
                        vh += (xh.c[c] > g_mid1[vh][p]);
                        qhi.c[c] = (uint8_t) clampi(vh * 2 + p, p, 127 - 1 + p);
                    }
                }
                else
                {
                    for(int c = 0; c < 4; c++)
                    {
                        qlo.c[c] = (uint8_t) clampi(((int) ((xl.c[c] * scalep - p) / 2.0f + .5f)) * 2 + p, p, iscalep - 1 + p);
                        qhi.c[c] = (uint8_t) clampi(((int) ((xh.c[c] * scalep - p) / 2.0f + .5f)) * 2 + p, p, iscalep - 1 + p);
                    }
                }
                rgba8 slo = expand_endpoint(&qlo, cfg), shi = expand_endpoint(&qhi, cfg);
                float e = 0;
                for(int i = 0; i < ncomp; i++) e += sqf((slo.c[i] / 255.0f) - xl.c[i]) + sqf((shi.c[i] / 255.0f) - xh.c[i]);
                if(p == 1) e *= cp->pbit1_weight;
                if(e < best_e)
                {
                    best_e = e;
                    best_pbits[0] = best_pbits[1] = (uint32_t) p;
                    for(int c = 0; c < 4; c++)
                    {
                        best_lo.c[c] = qlo.c[c] >> 1;
                        best_hi.c[c] = qhi.c[c] >> 1;
                    }
                }
            }
        }

        nudge_degenerate(mode, &best_lo, &best_hi, &xl, &xh, (uint32_t) (iscalep >> 1), cp);

        if((out->best_err == UINT64_MAX) || !same_rgba(&best_lo, &out->lo) || !same_rgba(&best_hi, &out->hi) ||
           (best_pbits[0] != out->pbits[0]) || (best_pbits[1] != out->pbits[1]))
            try_endpoints(&best_lo, &best_hi, best_pbits, cfg, out, cp);
    }
    else
    {
        const int iscale = (1 << cfg->comp_bits) - 1;
        const float scale = (float) iscale;
        rgba8 tlo, thi;
        if(cfg->comp_bits == 7)
        {
            for(int c = 0; c < 4; c++)
            {
                int vl = (int) (xl.c[c] * 127.0f);
                vl += (xl.c[c] > g_mid5[vl]);
                tlo.c[c] = (uint8_t) clampi(vl, 0, 127);
                int vh = (int) (xh.c[c] * 127.0f);
                vh += (xh.c[c] > g_mid5[vh]);
                thi.c[c] = (uint8_t) clampi(vh, 0, 127);
            }
        }
        else
        {
            for(int c = 0; c < 4; c++)
            {
                tlo.c[c] = (uint8_t) clampi((int) (xl.c[c] * scale + .5f), 0, 255);
                thi.c[c] = (uint8_t) clampi((int) (xh.c[c] * scale + .5f), 0, 255);
            }
        }
        nudge_degenerate(mode, &tlo, &thi, &xl, &xh, (uint32_t) iscale, cp);
        if((out->best_err == UINT64_MAX) || !same_rgba(&tlo, &out->lo) || !same_rgba(&thi, &out->hi))
            try_endpoints(&tlo, &thi, out->pbits, cfg, out, cp);
    }
    return out->best_err;
}

/* bc7enc.cpp:287-408: one routine for RGB (nc = 3, alpha endpoints := 255) and RGBA (nc = 4); channels are independent */
static void lsq_endpoints(uint32_t n, const uint8_t *sel, const float (*wx)[4], vec4 *xl, vec4 *xh, const rgba8 *px,
                          int nc)
{
    float z00 = 0.0f, z01, z10 = 0.0f, z11 = 0.0f;
    float q00[4] = {0, 0, 0, 0}, t[4] = {0, 0, 0, 0};
    for(uint32_t i = 0; i < n; i++)
    {
        const uint32_t s = sel[i];
        z00 += wx[s][0];
        z10 += wx[s][1];
        z11 += wx[s][2];
        float w = wx[s][3];
        for(int c = 0; c < nc; c++)
        {
            q00[c] += w * px[i].c[c];
            t[c] += px[i].c[c];
        }
    }
    float q10[4];
    for(int c = 0; c < nc; c++) q10[c] = t[c] - q00[c];
    z01 = z10;
    float det = z00 * z11 - z01 * z10;
    if(det != 0.0f) det = 1.0f / det;
    float iz00 = z11 * det, iz01 = -z01 * det, iz10 = -z10 * det, iz11 = z00 * det;
    for(int c = 0; c < nc; c++)
    {
        xl->c[c] = iz00 * q00[c] + iz01 * q10[c];
        xh->c[c] = iz10 * q00[c] + iz11 * q10[c];
    }
    if(nc == 3) xl->c[3] = xh->c[3] = 255.0f;
    for(int c = 0; c < nc; c++)
    {
        if((xl->c[c] < 0.0f) || (xh->c[c] > 255.0f))
        {
            uint32_t lo = UINT32_MAX, hi = 0;
            for(uint32_t i = 0; i < n; i++)
            {
                if(px[i].c[c] < lo) lo = px[i].c[c];
                if(px[i].c[c] > hi) hi = px[i].c[c];
            }
            if(lo == hi)
            {
                xl->c[c] = (float) lo;
                xh->c[c] = (float) hi;
            }
        }
    }
}

/* bc7enc.cpp:410-460 */
static void lsq_endpoints_alpha(const uint8_t *sel, float *xl, float *xh, const rgba8 *px)
{
    float z00 = 0.0f, z01, z10 = 0.0f, z11 = 0.0f, q00 = 0.0f, q10, t = 0.0f;
    for(uint32_t i = 0; i < 16; i++)
    {
        const uint32_t s = sel[i];
        z00 += W2X[s][0];
        z10 += W2X[s][1];
        z11 += W2X[s][2];
        float w = W2X[s][3];
        q00 += w * px[i].c[3];
        t += px[i].c[3];
    }
    q10 = t - q00;
    z01 = z10;
    float det = z00 * z11 - z01 * z10;
    if(det != 0.0f) det = 1.0f / det;
    float iz00 = z11 * det, iz01 = -z01 * det, iz10 = -z10 * det, iz11 = z00 * det;
    *xl = iz00 * q00 + iz01 * q10;
    *xh = iz10 * q00 + iz11 * q10;
    if((*xl < 0.0f) || (*xh > 255.0f))
    {
        uint32_t lo = UINT32_MAX, hi = 0;
        for(uint32_t i = 0; i < 16; i++)
        {
            if(px[i].c[3] < lo) lo = px[i].c[3];
            if(px[i].c[3] > hi) hi = px[i].c[3];
        }
        if(lo == hi)
        {
            *xl = (float) lo;
            *xh = (float) hi;
        }
    }
}

/* least squares from a selector vector, scale to [0,1], fit.  bc7enc.cpp:1285-1297 and its repeats */
static uint64_t refit(uint32_t mode, const uint8_t *sel, const cell_cfg *cfg, cell_out *out, const port_bc7_params *cp)
{
    vec4 xl = {{0, 0, 0, 0}}, xh = {{0, 0, 0, 0}};
    lsq_endpoints(cfg->n, sel, cfg->wx, &xl, &xh, cfg->px, cfg->has_alpha ? 4 : 3);
    for(int c = 0; c < 4; c++)
    {
        xl.c[c] = xl.c[c] * (1.0f / 255.0f);
        xh.c[c] = xh.c[c] * (1.0f / 255.0f);
    }
    return fit_endpoints(mode, xl, xh, cfg, out, cp);
}

/* bc7enc.cpp:1101-1441 */
static uint64_t compress_cell(uint32_t mode, const cell_cfg *cfg, cell_out *out, const port_bc7_params *cp)
{
    out->best_err = UINT64_MAX;
    const uint32_t n = cfg->n;
    const rgba8 *px = cfg->px;

    if(mode == 1 || mode == 7)
    {
        int same = 1;
        const int nc = (mode == 7) ? 4 : 3;
        for(uint32_t i = 1; i < n && same; i++)
            for(int c = 0; c < nc; c++)
                if(px[i].c[c] != px[0].c[c]) { same = 0; break; }
        if(same)
            return (mode == 1) ? solid_mode1(cfg, out, px[0].c[0], px[0].c[1], px[0].c[2], out->sel)
                               : solid_mode7(cfg, out, px[0].c[0], px[0].c[1], px[0].c[2], px[0].c[3], out->sel);
    }

    vec4 mean = {{0, 0, 0, 0}}, axis;
    for(uint32_t i = 0; i < n; i++)
        for(int c = 0; c < 4; c++) mean.c[c] = mean.c[c] + (float) px[i].c[c];
    vec4 mean_s;
    {
        const float inv_n = 1.0f / (float) n, inv_n255 = 1.0f / (float) (n * 255.0f);
        for(int c = 0; c < 4; c++)
        {
            mean_s.c[c] = mean.c[c] * inv_n;
            mean.c[c] = satf(mean.c[c] * inv_n255);
        }
    }

    if(cfg->has_alpha)
    {
        /* incremental PCA, bc7enc.cpp:1160-1177 */
        for(int c = 0; c < 4; c++) axis.c[c] = 0.0f;
        for(uint32_t i = 0; i < n; i++)
        {
            vec4 col, a, b, c, d, nrm;
            for(int k = 0; k < 4; k++) col.c[k] = (float) px[i].c[k] - mean_s.c[k];
            for(int k = 0; k < 4; k++)
            {
                a.c[k] = col.c[k] * col.c[0];
                b.c[k] = col.c[k] * col.c[1];
                c.c[k] = col.c[k] * col.c[2];
                d.c[k] = col.c[k] * col.c[3];
            }
            nrm = i ? axis : col;
            normalize4(&nrm);
            axis.c[0] += dot4(&a, &nrm);
            axis.c[1] += dot4(&b, &nrm);
            axis.c[2] += dot4(&c, &nrm);
            axis.c[3] += dot4(&d, &nrm);
        }
        normalize4(&axis);
    }
    else
    {
        /* covariance + 3 power iterations, bc7enc.cpp:1181-1218 */
        float cov[6] = {0, 0, 0, 0, 0, 0};
        for(uint32_t i = 0; i < n; i++)
        {
            float r = px[i].c[0] - mean_s.c[0], g = px[i].c[1] - mean_s.c[1], b = px[i].c[2] - mean_s.c[2];
            cov[0] += r * r;
            cov[1] += r * g;
            cov[2] += r * b;
            cov[3] += g * g;
            cov[4] += g * b;
            cov[5] += b * b;
        }
        float vr = .9f, vg = 1.0f, vb = .7f;
        for(int it = 0; it < 3; it++)
        {
            float r = vr * cov[0] + vg * cov[1] + vb * cov[2];
            float g = vr * cov[1] + vg * cov[3] + vb * cov[4];
            float b = vr * cov[2] + vg * cov[4] + vb * cov[5];
            float m = fabsf(r) > fabsf(g) ? fabsf(r) : fabsf(g);
            m = m > fabsf(b) ? m : fabsf(b);
            if(m > 1e-10f)
            {
                m = 1.0f / m;
                r *= m;
                g *= m;
                b *= m;
            }
            vr = r;
            vg = g;
            vb = b;
        }
        float len = vr * vr + vg * vg + vb * vb;
        if(len < 1e-10f) { axis = (vec4){{0, 0, 0, 0}}; }
        else
        {
            len = 1.0f / sqrtf(len);
            axis = (vec4){{vr * len, vg * len, vb * len, 0}};
        }
    }

    if(dot4(&axis, &axis) < .5f)
    {
        if(cfg->perceptual) axis = (vec4){{.213f, .715f, .072f, cfg->has_alpha ? .715f : 0}};
        else axis = (vec4){{1.0f, 1.0f, 1.0f, cfg->has_alpha ? 1.0f : 0}};
        normalize4(&axis);
    }

    float l = 1e+9f, h = -1e+9f;
    for(uint32_t i = 0; i < n; i++)
    {
        vec4 q;
        for(int k = 0; k < 4; k++) q.c[k] = (float) px[i].c[k] - mean_s.c[k];
        float d = dot4(&q, &axis);
        if(d < l) l = d;
        if(d > h) h = d;
    }
    l *= (1.0f / 255.0f);
    h *= (1.0f / 255.0f);

    vec4 c_lo, c_hi;
    for(int k = 0; k < 4; k++)
    {
        c_lo.c[k] = satf(mean.c[k] + axis.c[k] * l);
        c_hi.c[k] = satf(mean.c[k] + axis.c[k] * h);
    }
    {
        /* dot with (1,1,1,1): ((x*1 + y*1) + z*1) + w*1, bc7enc.cpp:1258 */
        vec4 ones = {{1.0f, 1.0f, 1.0f, 1.0f}};
        if(dot4(&c_lo, &ones) > dot4(&c_hi, &ones))
        {
            vec4 t = c_lo;
            c_lo = c_hi;
            c_hi = t;
        }
    }

    if(!fit_endpoints(mode, c_lo, c_hi, cfg, out, cp)) return 0;

    if(cp->try_least_squares)
        if(!refit(mode, out->sel, cfg, out, cp)) return 0;

    if(cp->uber_level > 0)
    {
        /* bc7enc.cpp:1300-1411 */
        uint8_t base[16], trial[16];
        memcpy(base, out->sel, n);
        const int max_sel_v = (int) cfg->nsel - 1;
        uint32_t min_sel = 16, max_sel = 0;
        for(uint32_t i = 0; i < n; i++)
        {
            if(base[i] < min_sel) min_sel = base[i];
            if(base[i] > max_sel) max_sel = base[i];
        }
        for(int variant = 0; variant < 3; variant++)
        {
            for(uint32_t i = 0; i < n; i++)
            {
                uint32_t s = base[i];
                if(variant != 1 && (s == min_sel) && (s < cfg->nsel - 1)) s++;
                else if(variant != 0 && (s == max_sel) && (s > 0)) s--;
                trial[i] = (uint8_t) s;
            }
            if(!refit(mode, trial, cfg, out, cp)) return 0;
        }
        const uint32_t thresh = (n * 56) >> 4;
        if((cp->uber_level >= 2) && (out->best_err > thresh))
        {
            const int Q = (cp->uber_level >= 4) ? ((int) cp->uber_level - 2) : 1;
            for(int ly = -Q; ly <= 1; ly++)
            {
                for(int hy = max_sel_v - 1; hy <= (max_sel_v + Q); hy++)
                {
                    if((ly == 0) && (hy == max_sel_v)) continue;
                    for(uint32_t i = 0; i < n; i++)
                    {
                        float v = floorf((float) max_sel_v * ((float) base[i] - (float) ly) / ((float) hy - (float) ly) + .5f);
                        if(v < 0) v = 0; else if(v > (float) max_sel_v) v = (float) max_sel_v;
                        trial[i] = (uint8_t) v;
                    }
                    if(!refit(mode, trial, cfg, out, cp)) return 0;
                }
            }
        }
    }

    if(mode == 1 || mode == 7)
    {
        /* try the mean as a single colour, bc7enc.cpp:1413-1438 */
        cell_out avg = *out;
        const uint32_t r = (uint32_t) (int) (.5f + mean.c[0] * 255.0f), g = (uint32_t) (int) (.5f + mean.c[1] * 255.0f),
                       b = (uint32_t) (int) (.5f + mean.c[2] * 255.0f), a = (uint32_t) (int) (.5f + mean.c[3] * 255.0f);
        uint64_t e = (mode == 1) ? solid_mode1(cfg, &avg, r, g, b, out->sel_tmp) : solid_mode7(cfg, &avg, r, g, b, a, out->sel_tmp);
        if(e < out->best_err)
        {
            *out = avg;
            memcpy(out->sel, out->sel_tmp, n);
            out->best_err = e;
        }
    }
    return out->best_err;
}

/* bc7enc.cpp:1443-1709: bbox-diagonal estimate for one subset; nch = 3 (mode 1, 8 levels) or 4 (mode 7, 4 levels) */
static uint64_t estimate_cell(uint32_t n, const rgba8 *px, int nch, int perceptual, const uint32_t w[4], uint64_t best_so_far)
{
    uint32_t lo[4] = {255, 255, 255, 255}, hi[4] = {0, 0, 0, 0};
    for(uint32_t i = 0; i < n; i++)
        for(int c = 0; c < nch; c++)
        {
            if(px[i].c[c] < lo[c]) lo[c] = px[i].c[c];
            if(px[i].c[c] > hi[c]) hi[c] = px[i].c[c];
        }
    const int N = (nch == 3) ? 8 : 4;
    const uint32_t *iw = (nch == 3) ? W3 : W2;
    int pal[8][4];
    for(int i = 0; i < N; i++)
        for(int c = 0; c < nch; c++)
            pal[i][c] = (i == 0) ? (int) lo[c] : (i == N - 1) ? (int) hi[c] : (int) (uint8_t) ((lo[c] * (64 - iw[i]) + hi[c] * iw[i] + 32) >> 6);
    int ax[4] = {0, 0, 0, 0};
    for(int c = 0; c < nch; c++) ax[c] = (int) hi[c] - (int) lo[c];
    int dots[8], thr[7];
    for(int i = 0; i < N; i++)
    {
        int d = 0;
        for(int c = 0; c < nch; c++) d += pal[i][c] * ax[c];
        dots[i] = d;
    }
    for(int i = 0; i < N - 1; i++) thr[i] = (dots[i] + dots[i + 1] + 1) >> 1;

    int l1[8], cr1[8], cb1[8];
    if(perceptual)
        for(int j = 0; j < N; j++)
        {
            l1[j] = pal[j][0] * 109 + pal[j][1] * 366 + pal[j][2] * 37;
            cr1[j] = (pal[j][0] << 9) - l1[j];
            cb1[j] = (pal[j][2] << 9) - l1[j];
        }

    uint64_t total = 0;
    for(uint32_t i = 0; i < n; i++)
    {
        const rgba8 *c = &px[i];
        int d = 0;
        for(int k = 0; k < nch; k++) d += ax[k] * c->c[k];
        int s = 0;
        for(int k = N - 2; k >= 0; k--)
            if(d >= thr[k]) { s = k + 1; break; }
        if(perceptual)
        {
            const int l2 = c->c[0] * 109 + c->c[1] * 366 + c->c[2] * 37;
            const int cr2 = ((int) c->c[0] << 9) - l2, cb2 = ((int) c->c[2] << 9) - l2;
            const int dl = (l1[s] - l2) >> 8, dcr = (cr1[s] - cr2) >> 8, dcb = (cb1[s] - cb2) >> 8;
            /* uint32 products, then (int): bc7enc.cpp:1533,1670 */
            uint32_t e = (w[0] * (uint32_t) dl * (uint32_t) dl) + (w[1] * (uint32_t) dcr * (uint32_t) dcr) + (w[2] * (uint32_t) dcb * (uint32_t) dcb);
            if(nch == 4)
            {
                const int dca = (int) c->c[3] - pal[s][3];
                e += w[3] * (uint32_t) dca * (uint32_t) dca;
            }
            int ie = (int) e;
            total += (uint64_t) (int64_t) ie;
        }
        else
        {
            int dr = pal[s][0] - (int) c->c[0], dg = pal[s][1] - (int) c->c[1], db = pal[s][2] - (int) c->c[2];
            uint32_t e = w[0] * (uint32_t) (dr * dr) + w[1] * (uint32_t) (dg * dg) + w[2] * (uint32_t) (db * db);
            if(nch == 4)
            {
                int da = pal[s][3] - (int) c->c[3];
                e += w[3] * (uint32_t) (da * da);
            }
            total += e;
        }
        if(total > best_so_far) break;
    }
    return total;
}

/* bc7enc.cpp:1754-1838 */
static uint32_t estimate_partition(const rgba8 *px, const port_bc7_params *cp, const uint32_t w[4], uint32_t mode)
{
    const uint32_t total_partitions = cp->max_partitions < 64 ? cp->max_partitions : 64;
    if(total_partitions <= 1) return 0;
    uint64_t best_err = UINT64_MAX;
    uint32_t best_partition = 0;
    int key = 0;
    for(uint32_t it = 0; (it < total_partitions) && (best_err > 0); it++)
    {
        const uint32_t part = PART_ORDER[it];
        if(cp->mode17_partition_estimation_filterbank && (it >= 14) && (it <= 34))
        {
            if((PART_PRED[part] & (1u << (key + 1))) == 0)
            {
                if(it == 34) break;
                continue;
            }
        }
        rgba8 sub[2][16];
        uint32_t cnt[2] = {0, 0};
        for(uint32_t i = 0; i < 16; i++)
        {
            uint32_t s = (PART2[part] >> i) & 1;
            sub[s][cnt[s]++] = px[i];
        }
        uint64_t err = 0;
        for(uint32_t s = 0; (s < 2) && (err < best_err); s++)
            err += estimate_cell(cnt[s], sub[s], (mode == 7) ? 4 : 3, (int) cp->perceptual, w, best_err);
        if(part < 16) err = (uint64_t) ((double) err * cp->low_frequency_partition_weight + .5f);
        if(err < best_err)
        {
            best_err = err;
            best_partition = part;
        }
        if((part == 34) && (best_partition != 34)) break;
        if(it == 13) key = (int) best_partition;
    }
    return best_partition;
}

/* ------------------------------------------------------------------------------------------------ packing */
typedef struct
{
    uint32_t mode, partition;
    uint8_t sel[16], asel[16];
    rgba8 lo[2], hi[2];
    uint32_t pbits[2][2];
} block_solution;

static void put_bits(uint8_t *bytes, uint32_t val, uint32_t nbits, uint32_t *ofs) /* bc7enc.cpp:1840-1852, LSB first */
{
    for(uint32_t i = 0; i < nbits; i++, (*ofs)++)
        if((val >> i) & 1) bytes[*ofs >> 3] |= (uint8_t) (1u << (*ofs & 7));
}

/* bc7enc.cpp:1867-2037 restricted to the modes the encoder emits (1, 5, 6, 7); layouts in SURVEY.md App. B */
static void pack_block(uint8_t *out, const block_solution *sol)
{
    static const uint8_t color_bits[8] = {4, 6, 5, 7, 5, 7, 7, 5}, alpha_bits[8] = {0, 0, 0, 0, 6, 8, 7, 5};
    static const uint8_t index_bits[8] = {3, 3, 2, 2, 2, 2, 4, 2};
    const uint32_t mode = sol->mode;
    const uint32_t subsets = (mode == 1 || mode == 7) ? 2 : 1;
    const uint32_t mask = (subsets == 2) ? PART2[sol->partition] : 0;
    const int separate_alpha = (mode == 5), shared_p = (mode == 1), has_p = (mode != 5);
    uint8_t sel[16], asel[16];
    rgba8 lo[2], hi[2];
    uint32_t pb[2][2];
    memcpy(sel, sol->sel, 16);
    memcpy(asel, sol->asel, 16);
    memcpy(lo, sol->lo, sizeof lo);
    memcpy(hi, sol->hi, sizeof hi);
    memcpy(pb, sol->pbits, sizeof pb);
    int anchor[2] = {-1, -1};
    const uint32_t nidx = 1u << index_bits[mode];
    for(uint32_t k = 0; k < subsets; k++)
    {
        const uint32_t a = k ? ANCHOR2[sol->partition] : 0;
        anchor[k] = (int) a;
        if(sel[a] & (nidx >> 1))
        {
            for(uint32_t i = 0; i < 16; i++)
                if(((mask >> i) & 1) == k) sel[i] = (uint8_t) ((nidx - 1) - sel[i]);
            const int nswap = separate_alpha ? 3 : 4;
            for(int q = 0; q < nswap; q++)
            {
                uint8_t t = lo[k].c[q];
                lo[k].c[q] = hi[k].c[q];
                hi[k].c[q] = t;
            }
            if(!shared_p)
            {
                uint32_t t = pb[k][0];
                pb[k][0] = pb[k][1];
                pb[k][1] = t;
            }
        }
        if(separate_alpha)
        {
            const uint32_t na = 4; /* mode 5: 2-bit alpha indices */
            if(asel[a] & (na >> 1))
            {
                for(uint32_t i = 0; i < 16; i++) asel[i] = (uint8_t) ((na - 1) - asel[i]);
                uint8_t t = lo[k].c[3];
                lo[k].c[3] = hi[k].c[3];
                hi[k].c[3] = t;
            }
        }
    }
    memset(out, 0, 16);
    uint32_t ofs = 0;
    put_bits(out, 1u << mode, mode + 1, &ofs);
    if(mode == 5) put_bits(out, 0, 2, &ofs); /* rotation */
    if(subsets == 2) put_bits(out, sol->partition, 6, &ofs);
    const uint32_t ncomp = (mode >= 4) ? 4 : 3;
    for(uint32_t c = 0; c < ncomp; c++)
        for(uint32_t s = 0; s < subsets; s++)
        {
            const uint32_t nb = (c == 3) ? alpha_bits[mode] : color_bits[mode];
            put_bits(out, lo[s].c[c], nb, &ofs);
            put_bits(out, hi[s].c[c], nb, &ofs);
        }
    if(has_p)
        for(uint32_t s = 0; s < subsets; s++)
        {
            put_bits(out, pb[s][0], 1, &ofs);
            if(!shared_p) put_bits(out, pb[s][1], 1, &ofs);
        }
    for(int i = 0; i < 16; i++)
    {
        uint32_t nb = index_bits[mode];
        if(i == anchor[0] || i == anchor[1]) nb--;
        put_bits(out, sel[i], nb, &ofs);
    }
    if(separate_alpha)
        for(int i = 0; i < 16; i++)
        {
            uint32_t nb = 2;
            if(i == anchor[0]) nb--;
            put_bits(out, asel[i], nb, &ofs);
        }
}

/* ------------------------------------------------------------------------------------------------ block paths */
static void split_subsets(const rgba8 *px, uint32_t part, rgba8 sub[2][16], uint8_t idx[2][16], uint32_t cnt[2])
{
    cnt[0] = cnt[1] = 0;
    for(uint32_t i = 0; i < 16; i++)
    {
        const uint32_t s = (PART2[part] >> i) & 1;
        sub[s][cnt[s]] = px[i];
        idx[s][cnt[s]] = (uint8_t) i;
        cnt[s]++;
    }
}

/* two-subset trial shared by mode 1 (bc7enc.cpp:2336-2397) and mode 7 (bc7enc.cpp:2193-2259) */
static uint64_t two_subset_trial(uint32_t mode, const rgba8 *px, cell_cfg *cfg, const port_bc7_params *cp, float mode_weight,
                                 uint64_t best_err, block_solution *sol, uint8_t *sel_tmp)
{
    const uint32_t part = estimate_partition(px, cp, cfg->weights, mode);
    if(mode == 1)
    {
        cfg->w = W3; cfg->wx = W3X; cfg->nsel = 8; cfg->comp_bits = 6; cfg->has_pbits = 1; cfg->share_pbit = 1;
    }
    else
    {
        cfg->w = W2; cfg->wx = W2X; cfg->nsel = 4; cfg->comp_bits = 5; cfg->has_pbits = 1; cfg->share_pbit = 0; cfg->has_alpha = 1;
    }
    rgba8 sub[2][16];
    uint8_t idx[2][16], ssel[2][16];
    uint32_t cnt[2];
    cell_out res[2];
    memset(res, 0, sizeof res);
    split_subsets(px, part, sub, idx, cnt);
    uint64_t trial = 0;
    for(uint32_t s = 0; s < 2; s++)
    {
        cfg->n = cnt[s];
        cfg->px = sub[s];
        res[s].sel = ssel[s];
        res[s].sel_tmp = sel_tmp;
        trial += compress_cell(mode, cfg, &res[s], cp);
        if(weigh(trial, mode_weight) > best_err) break;
    }
    const uint64_t werr = weigh(trial, mode_weight);
    if(werr < best_err)
    {
        sol->mode = mode;
        sol->partition = part;
        for(uint32_t s = 0; s < 2; s++)
        {
            for(uint32_t i = 0; i < cnt[s]; i++) sol->sel[idx[s][i]] = ssel[s][i];
            sol->lo[s] = res[s].lo;
            sol->hi[s] = res[s].hi;
            sol->pbits[s][0] = res[s].pbits[0];
            sol->pbits[s][1] = res[s].pbits[1];
        }
        return werr;
    }
    return UINT64_MAX;
}

/* bc7enc.cpp:2293-2400 */
static void encode_opaque(uint8_t *out, const rgba8 *px, const port_bc7_params *cp, cell_cfg *cfg)
{
    uint8_t sel_tmp[16];
    block_solution sol;
    memset(&sol, 0, sizeof sol);
    uint64_t best_err = UINT64_MAX;
    cfg->perceptual = (int) cp->perceptual;
    cfg->n = 16;
    cfg->px = px;
    cfg->has_alpha = 0;
    if(cp->mode_mask & (1u << 6))
    {
        cfg->w = W4; cfg->wx = W4X; cfg->nsel = 16; cfg->comp_bits = 7; cfg->has_pbits = 1; cfg->share_pbit = 0;
        cell_out r6;
        memset(&r6, 0, sizeof r6);
        r6.sel = sol.sel;
        r6.sel_tmp = sel_tmp;
        best_err = weigh(compress_cell(6, cfg, &r6, cp), cp->mode6_error_weight);
        sol.mode = 6;
        sol.lo[0] = r6.lo;
        sol.hi[0] = r6.hi;
        sol.pbits[0][0] = r6.pbits[0];
        sol.pbits[0][1] = r6.pbits[1];
    }
    if((best_err > 0) && (cp->max_partitions > 0) && (cp->mode_mask & (1u << 1)))
    {
        block_solution s1;
        memset(&s1, 0, sizeof s1);
        uint64_t e = two_subset_trial(1, px, cfg, cp, cp->mode1_error_weight, best_err, &s1, sel_tmp);
        if(e != UINT64_MAX) sol = s1;
    }
    pack_block(out, &sol);
}

/* bc7enc.cpp:2039-2137 */
static void mode5_trial(const rgba8 *px, const port_bc7_params *cp, cell_cfg *cfg, uint32_t lo_a, uint32_t hi_a,
                        block_solution *sol, uint64_t *err5)
{
    cfg->w = W2; cfg->wx = W2X; cfg->nsel = 4; cfg->comp_bits = 7; cfg->has_pbits = 0; cfg->share_pbit = 0; cfg->has_alpha = 0;
    cfg->perceptual = (int) cp->perceptual;
    cfg->n = 16;
    cfg->px = px;
    cell_out r5;
    memset(&r5, 0, sizeof r5);
    uint8_t sel_tmp[16];
    r5.sel = sol->sel;
    r5.sel_tmp = sel_tmp;
    *err5 = compress_cell(5, cfg, &r5, cp);
    sol->lo[0] = r5.lo;
    sol->hi[0] = r5.hi;
    if(lo_a == hi_a)
    {
        sol->lo[0].c[3] = (uint8_t) lo_a;
        sol->hi[0].c[3] = (uint8_t) hi_a;
        memset(sol->asel, 0, 16);
        return;
    }
    uint64_t alpha_err = UINT64_MAX;
    const uint32_t passes = (cp->uber_level >= 1) ? 3 : 2;
    for(uint32_t pass = 0; pass < passes; pass++)
    {
        int32_t v[4];
        v[0] = (int32_t) lo_a;
        v[3] = (int32_t) hi_a;
        v[1] = (v[0] * (64 - 21) + v[3] * 21 + 32) >> 6;
        v[2] = (v[0] * (64 - 43) + v[3] * 43 + 32) >> 6;
        uint8_t tsel[16];
        uint64_t terr = 0;
        for(uint32_t i = 0; i < 16; i++)
        {
            const int32_t a = px[i].c[3];
            int s = 0;
            int32_t be = abs(a - v[0]);
            int e = abs(a - v[1]); if(e < be) { be = e; s = 1; }
            e = abs(a - v[2]); if(e < be) { be = e; s = 2; }
            e = abs(a - v[3]); if(e < be) { be = e; s = 3; }
            tsel[i] = (uint8_t) s;
            uint32_t a_err = (uint32_t) (be * be) * cfg->weights[3];
            terr += a_err;
        }
        if(terr < alpha_err)
        {
            alpha_err = terr;
            sol->lo[0].c[3] = (uint8_t) lo_a;
            sol->hi[0].c[3] = (uint8_t) hi_a;
            memcpy(sol->asel, tsel, 16);
        }
        if(pass != passes - 1)
        {
            float xl, xh;
            lsq_endpoints_alpha(tsel, &xl, &xh, px);
            uint32_t nlo = (uint32_t) clampi((int) floor(xl + .5f), 0, 255);
            uint32_t nhi = (uint32_t) clampi((int) floor(xh + .5f), 0, 255);
            if(nlo > nhi) { uint32_t t = nlo; nlo = nhi; nhi = t; }
            if((nlo == lo_a) && (nhi == hi_a)) break;
            lo_a = nlo;
            hi_a = nhi;
        }
    }
    *err5 += alpha_err;
}

/* bc7enc.cpp:2139-2291 */
static void encode_alpha(uint8_t *out, const rgba8 *px, const port_bc7_params *cp, cell_cfg *cfg)
{
    cfg->w = W4; cfg->wx = W4X; cfg->nsel = 16; cfg->comp_bits = 7; cfg->has_pbits = 1; cfg->share_pbit = 0; cfg->has_alpha = 1;
    cfg->perceptual = (int) cp->perceptual;
    cfg->n = 16;
    cfg->px = px;
    block_solution s6, s5, s7;
    memset(&s6, 0, sizeof s6);
    memset(&s5, 0, sizeof s5);
    memset(&s7, 0, sizeof s7);
    uint64_t best_err = UINT64_MAX;
    uint32_t best_mode = 0;
    uint8_t sel_tmp[16];
    if(cp->mode_mask & (1u << 6))
    {
        cell_out r6;
        memset(&r6, 0, sizeof r6);
        r6.sel = s6.sel;
        r6.sel_tmp = sel_tmp;
        best_err = weigh(compress_cell(6, cfg, &r6, cp), cp->mode6_error_weight);
        best_mode = 6;
        s6.mode = 6;
        s6.lo[0] = r6.lo;
        s6.hi[0] = r6.hi;
        s6.pbits[0][0] = r6.pbits[0];
        s6.pbits[0][1] = r6.pbits[1];
    }
    if((best_err > 0) && (cp->mode_mask & (1u << 5)))
    {
        uint32_t lo_a = 255, hi_a = 0;
        for(uint32_t i = 0; i < 16; i++)
        {
            if(px[i].c[3] < lo_a) lo_a = px[i].c[3];
            if(px[i].c[3] > hi_a) hi_a = px[i].c[3];
        }
        uint64_t e5;
        mode5_trial(px, cp, cfg, lo_a, hi_a, &s5, &e5);
        e5 = weigh(e5, cp->mode5_error_weight);
        if(e5 < best_err)
        {
            best_err = e5;
            best_mode = 5;
            s5.mode = 5;
        }
    }
    if((best_err > 0) && (cp->mode_mask & (1u << 7)))
    {
        uint64_t e = two_subset_trial(7, px, cfg, cp, cp->mode7_error_weight, best_err, &s7, sel_tmp);
        if(e != UINT64_MAX)
        {
            best_err = e;
            best_mode = 7;
        }
    }
    pack_block(out, best_mode == 7 ? &s7 : best_mode == 5 ? &s5 : &s6);
}

/* ------------------------------------------------------------------------------------------------ public */
void port_bc7_params_init(port_bc7_params *p) /* bc7enc.h:95-113 */
{
    memset(p, 0, sizeof *p);
    p->mode_mask = UINT32_MAX;
    p->max_partitions = 64;
    p->weights[0] = 128; p->weights[1] = 64; p->weights[2] = 16; p->weights[3] = 32;
    p->perceptual = 1;
    p->try_least_squares = 1;
    p->mode17_partition_estimation_filterbank = 1;
    p->pbit1_weight = p->mode1_error_weight = p->mode5_error_weight = p->mode6_error_weight = p->mode7_error_weight = 1.0f;
    p->low_frequency_partition_weight = 1.0f;
}

/* bc7enc.cpp:2402-2438 */
int port_bc7_encode_block(const uint8_t *rgba64, const port_bc7_params *params, uint8_t *out16)
{
    pthread_once(&g_once, build_tables);
    port_bc7_params def;
    if(!params)
    {
        port_bc7_params_init(&def);
        params = &def;
    }
    const rgba8 *px = (const rgba8 *) rgba64;
    cell_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    if(params->perceptual)
    {
        const float pr = (.5f / (1.0f - .2126f)) * (.5f / (1.0f - .2126f));
        const float pb = (.5f / (1.0f - .0722f)) * (.5f / (1.0f - .0722f));
        cfg.weights[0] = (uint32_t) (int) (params->weights[0] * 4.0f);
        cfg.weights[1] = (uint32_t) (int) (params->weights[1] * 4.0f * pr);
        cfg.weights[2] = (uint32_t) (int) (params->weights[2] * 4.0f * pb);
        cfg.weights[3] = params->weights[3] * 4;
    }
    else memcpy(cfg.weights, params->weights, sizeof cfg.weights);

    int alpha = params->force_alpha != 0;
    for(int i = 0; i < 16 && !alpha; i++) alpha = px[i].c[3] < 255;
    if(alpha) encode_alpha(out16, px, params, &cfg);
    else encode_opaque(out16, px, params, &cfg);
    return alpha;
}

typedef struct
{
    const uint8_t *px;
    uint8_t *out;
    uint64_t n;
    const port_bc7_params *params;
    uint64_t *next;
} job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *) arg;
    for(;;)
    {
        uint64_t b0 = __atomic_fetch_add(j->next, 256, __ATOMIC_RELAXED);
        if(b0 >= j->n) break;
        uint64_t b1 = b0 + 256 < j->n ? b0 + 256 : j->n;
        for(uint64_t b = b0; b < b1; b++) port_bc7_encode_block(j->px + 64 * b, j->params, j->out + 16 * b);
    }
    return NULL;
}

void port_bc7_encode_blocks(const uint8_t *px, uint64_t num_blocks, const port_bc7_params *params, uint8_t *out, int threads)
{
    pthread_once(&g_once, build_tables);
    uint64_t next = 0;
    job_t job = {px, out, num_blocks, params, &next};
    if(threads <= 1) { worker(&job); return; }
    if(threads > 256) threads = 256;
    pthread_t th[256];
    for(int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &job);
    for(int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}

/* ------------------------------------------------------------------------------------------------ decoder */
/* BC7 block decoder for the modes bc7enc emits (1, 5, 6, 7), written from the BC7 format definition (same results
 * as bc7decomp::unpack_bc7, bc7decomp.cpp:594, on those modes; pinned in tests/test_oracle_pinning.py).  Blocks of any
 * other mode decode to transparent black.  Used only for the PSNR fallback metric. */
static uint32_t get_bits(const uint8_t *b, uint32_t *ofs, uint32_t n)
{
    uint32_t v = 0;
    for(uint32_t i = 0; i < n; i++, (*ofs)++) v |= (uint32_t) ((b[*ofs >> 3] >> (*ofs & 7)) & 1) << i;
    return v;
}

static void unpack_block(const uint8_t *b, uint8_t *px)
{
    memset(px, 0, 64);
    uint32_t mode = 0;
    while(mode < 8 && !((b[0] >> mode) & 1)) mode++;
    if(!(mode == 1 || mode == 5 || mode == 6 || mode == 7)) return;
    static const uint8_t cbits[8] = {4, 6, 5, 7, 5, 7, 7, 5}, abits[8] = {0, 0, 0, 0, 6, 8, 7, 5}, ibits[8] = {3, 3, 2, 2, 2, 2, 4, 2};
    const uint32_t subsets = (mode == 1 || mode == 7) ? 2 : 1;
    uint32_t ofs = mode + 1, rot = 0, part = 0;
    if(mode == 5) rot = get_bits(b, &ofs, 2);
    if(subsets == 2) part = get_bits(b, &ofs, 6);
    uint32_t ep[4][4]; /* [subset*2 + lo/hi][comp] */
    const uint32_t ncomp = (mode >= 4) ? 4 : 3;
    for(uint32_t c = 0; c < ncomp; c++)
        for(uint32_t e = 0; e < subsets * 2; e++) ep[e][c] = get_bits(b, &ofs, c == 3 ? abits[mode] : cbits[mode]);
    if(mode != 5)
    {
        uint32_t pb[4];
        for(uint32_t e = 0; e < subsets * 2; e++)
        {
            if(mode == 1) { if(!(e & 1)) pb[e] = get_bits(b, &ofs, 1); else pb[e] = pb[e - 1]; }
            else pb[e] = get_bits(b, &ofs, 1);
        }
        for(uint32_t e = 0; e < subsets * 2; e++)
            for(uint32_t c = 0; c < ncomp; c++) ep[e][c] = (ep[e][c] << 1) | pb[e];
    }
    for(uint32_t e = 0; e < subsets * 2; e++)
    {
        for(uint32_t c = 0; c < ncomp; c++)
        {
            const uint32_t n = (c == 3 ? abits[mode] : cbits[mode]) + (mode != 5 ? 1 : 0);
            uint32_t v = ep[e][c] << (8 - n);
            ep[e][c] = v | (v >> n);
        }
        if(ncomp == 3) ep[e][3] = 255;
    }
    const uint32_t mask = subsets == 2 ? PART2[part] : 0;
    const uint32_t *w = ibits[mode] == 2 ? W2 : ibits[mode] == 3 ? W3 : W4;
    uint32_t csel[16], asel[16];
    for(uint32_t i = 0; i < 16; i++)
    {
        uint32_t nb = ibits[mode];
        if(i == 0 || (subsets == 2 && i == ANCHOR2[part])) nb--;
        csel[i] = get_bits(b, &ofs, nb);
    }
    if(mode == 5)
        for(uint32_t i = 0; i < 16; i++) asel[i] = get_bits(b, &ofs, i == 0 ? 1 : 2);
    for(uint32_t i = 0; i < 16; i++)
    {
        const uint32_t s = (mask >> i) & 1;
        uint32_t out[4];
        for(uint32_t c = 0; c < 4; c++)
        {
            const uint32_t wt = (mode == 5 && c == 3) ? W2[asel[i]] : w[csel[i]];
            out[c] = (ep[2 * s][c] * (64 - wt) + ep[2 * s + 1][c] * wt + 32) >> 6;
        }
        if(rot)
        {
            uint32_t t = out[3];
            out[3] = out[rot - 1];
            out[rot - 1] = t;
        }
        for(uint32_t c = 0; c < 4; c++) px[4 * i + c] = (uint8_t) out[c];
    }
}

void port_bc7_unpack_blocks(const uint8_t *blocks, uint64_t num_blocks, uint8_t *px)
{
    for(uint64_t b = 0; b < num_blocks; b++) unpack_block(blocks + 16 * b, px + 64 * b);
}
