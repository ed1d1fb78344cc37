/* oracle/stbir_oracle.c -- TEST INFRASTRUCTURE ONLY (see bc7_oracle.h for the rules).
 *
 * CPU restatement of crocore::Image_<uint8_t>::resize (/root/reference/extern/crocore/src/Image.cpp:239-247), i.e.
 * stbir_resize_uint8 with its defaults (extern/crocore/src/stb_image_resize.h:2462-2470): Catmull-Rom when a dimension
 * is enlarged, Mitchell otherwise (also at 1:1), clamp-to-edge, linear colour space, float accumulation.
 * vierkant::bcn::compress runs every mip level -- including level 0 -- through it (src/texture_block_compression.cpp:101).
 *
 * Structure follows the reference (filter tables first, then horizontal and vertical passes with the reference's
 * accumulation order), but without its ring buffer: all horizontally filtered rows are kept.
 *   stbir__calculate_sample_range_*     stb_image_resize.h:1009-1038
 *   stbir__calculate_coefficients_*     :1040-1124
 *   stbir__normalize_downsample_...     :1126-1199
 *   stbir__calculate_filters            :1203-1241
 *   decode  u8 / 255.0f                 :1304-1312
 *   horizontal passes                   :1450-1531 (gather, enlarging), :1533-1670 (scatter, reducing)
 *   vertical passes                     :1870-1985 (gather), :1987-2060 (scatter); row loops :2067-2204
 *   encode  (int)(saturate(f) * 255.0f + 0.5)   :1737-1746,1753-1763
 * Parity status: PINNED against the unmodified reference (oracle/_ref, ref_resize_u8) in tests/test_oracle_pinning.py.
 * Compile with -ffp-contract=off (the reference's x86-64 build has no FMA).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "bc7_oracle.h"

typedef struct
{
    int n0, n1;
} span_t;

/* one resampled dimension */
typedef struct
{
    int in_size, out_size;
    int enlarge;     /* scale > 1: per-output gather lists; else per-input scatter lists */
    float scale;
    int margin;      /* filter_pixel_margin */
    int width;       /* coefficient_width: slots per list in the flat table */
    int count;       /* number of lists */
    span_t *span;
    float *coef;     /* flat, `width` slots per list; the reference lets a list spill into its successor's first slot */
} axis_t;

/* stbir__filter_catmullrom / stbir__filter_mitchell, :816-844; float arithmetic throughout */
static float kernel_catmullrom(float x)
{
    x = (float) fabs(x);
    if(x < 1.0f) { return 1 - x * x * (2.5f - 1.5f * x); }
    else if(x < 2.0f) { return 2 - x * (4 + x * (0.5f * x - 2.5f)); }
    return 0.0f;
}
static float kernel_mitchell(float x)
{
    x = (float) fabs(x);
    if(x < 1.0f) { return (16 + x * x * (21 * x - 36)) / 18; }
    else if(x < 2.0f) { return (32 + x * (-60 + x * (36 - 7 * x))) / 18; }
    return 0.0f;
}

static void axis_build(axis_t *a, int in_size, int out_size)
{
    const float support = 2.0f; /* both default filters, :852-856 */
    memset(a, 0, sizeof(*a));
    a->in_size = in_size, a->out_size = out_size;
    a->scale = ((float) out_size / in_size) / (1.0f - 0.0f); /* stbir__calculate_transform :2233-2234, s0 = 0, s1 = 1 */
    const float shift = 0.0f * out_size / (1.0f - 0.0f);     /* :2236 */
    a->enlarge = a->scale > 1;
    a->width = (int) ceil(support * 2);                      /* stbir__get_coefficient_width :915-921 */
    const int pixel_width = a->enlarge ? (int) ceil(support * 2) : (int) ceil(support * 2 / a->scale); /* :892-901 */
    a->margin = pixel_width / 2;
    a->count = a->enlarge ? out_size : in_size + a->margin * 2; /* stbir__get_contributors :923-929 */
    a->span = (span_t *) calloc((size_t) a->count + 1, sizeof(span_t));
    a->coef = (float *) calloc((size_t) (a->count + 3) * (size_t) a->width, sizeof(float));

    if(a->enlarge)
    {
        const float radius = support * a->scale; /* out_pixels_radius, :1210 */
        for(int n = 0; n < a->count; ++n)
        {
            /* stbir__calculate_sample_range_upsample :1009-1022 (the +-0.5 are double constants) */
            const float centre = (float) n + 0.5f;
            const float lo = centre - radius, hi = centre + radius;
            const float in_lo = (lo + shift) / a->scale, in_hi = (hi + shift) / a->scale;
            const float in_centre = (centre + shift) / a->scale;
            int first = (int) (floor(in_lo + 0.5));
            const int last = (int) (floor(in_hi - 0.5));
            /* stbir__calculate_coefficients_upsample :1040-1093 */
            float *g = a->coef + (size_t) a->width * n;
            span_t *s = &a->span[n];
            float total = 0;
            s->n0 = first, s->n1 = last;
            for(int i = 0; i <= last - first; i++)
            {
                const float tap_centre = (float) (i + first) + 0.5f;
                g[i] = kernel_catmullrom(in_centre - tap_centre);
                if(i == 0 && !g[i])
                {
                    s->n0 = ++first;
                    i--;
                    continue;
                }
                total += g[i];
            }
            const float norm = 1 / total;
            for(int i = 0; i <= last - first; i++) { g[i] *= norm; }
            for(int i = last - first; i >= 0; i--)
            {
                if(g[i]) { break; }
                s->n1 = s->n0 + i - 1;
            }
        }
    }
    else
    {
        const float radius = support / a->scale; /* in_pixels_radius, :1225 */
        for(int n = 0; n < a->count; ++n)
        {
            /* stbir__calculate_sample_range_downsample :1025-1038 */
            const float centre = (float) (n - a->margin) + 0.5f;
            const float lo = centre - radius, hi = centre + radius;
            const float out_lo = lo * a->scale - shift, out_hi = hi * a->scale - shift;
            const float out_centre = centre * a->scale - shift;
            const int first = (int) (floor(out_lo + 0.5)), last = (int) (floor(out_hi - 0.5));
            /* stbir__calculate_coefficients_downsample :1095-1124 */
            float *g = a->coef + (size_t) a->width * n;
            span_t *s = &a->span[n];
            s->n0 = first, s->n1 = last;
            for(int i = 0; i <= last - first; i++)
            {
                const float x = ((float) (i + first) + 0.5f) - out_centre;
                g[i] = kernel_mitchell(x) * a->scale;
            }
            for(int i = last - first; i >= 0; i--)
            {
                if(g[i]) { break; }
                s->n1 = s->n0 + i - 1;
            }
        }
        /* stbir__normalize_downsample_coefficients :1126-1199 */
        for(int i = 0; i < out_size; i++)
        {
            float total = 0;
            for(int j = 0; j < a->count; j++)
            {
                if(i >= a->span[j].n0 && i <= a->span[j].n1) { total += a->coef[(size_t) a->width * j + (i - a->span[j].n0)]; }
                else if(i < a->span[j].n0) { break; }
            }
            const float norm = 1 / total;
            for(int j = 0; j < a->count; j++)
            {
                if(i >= a->span[j].n0 && i <= a->span[j].n1) { a->coef[(size_t) a->width * j + (i - a->span[j].n0)] *= norm; }
                else if(i < a->span[j].n0) { break; }
            }
        }
        for(int j = 0; j < a->count; j++)
        {
            float *g = a->coef + (size_t) a->width * j;
            int skip = 0;
            while(g[skip] == 0) { skip++; }
            a->span[j].n0 += skip;
            while(a->span[j].n0 < 0)
            {
                a->span[j].n0++;
                skip++;
            }
            const int range = a->span[j].n1 - a->span[j].n0 + 1;
            const int max = a->width < range ? a->width : range;
            for(int i = 0; i < max; i++)
            {
                if(i + skip >= a->width) { break; }
                g[i] = g[i + skip];
            }
        }
        for(int j = 0; j < a->count; j++)
        {
            if(a->span[j].n1 > out_size - 1) { a->span[j].n1 = out_size - 1; }
        }
    }
}

static void axis_free(axis_t *a)
{
    free(a->span);
    free(a->coef);
}

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void port_resize_u8(const uint8_t *in, uint32_t w, uint32_t h, uint32_t comps, uint8_t *out, uint32_t ow, uint32_t oh)
{
    axis_t ax, ay;
    axis_build(&ax, (int) w, (int) ow);
    axis_build(&ay, (int) h, (int) oh);
    const int C = (int) comps;
    const size_t row_f = (size_t) ow * C;
    /* horizontal pass of every input row (rows outside the image are clamped copies of these) */
    float *hbuf = (float *) calloc((size_t) h * row_f, sizeof(float));
    float *dec = (float *) malloc(((size_t) w + 2 * (size_t) ax.margin) * C * sizeof(float));
    for(uint32_t y = 0; y < h; ++y)
    {
        const uint8_t *src = in + (size_t) y * w * C;
        float *d0 = dec + (size_t) ax.margin * C; /* index 0 of the decode buffer, :1246-1251 */
        for(int x = -ax.margin; x < (int) w + ax.margin; ++x)
        {
            const int sx = clampi(x, 0, (int) w - 1);
            for(int c = 0; c < C; ++c) { d0[x * C + c] = ((float) src[sx * C + c]) / 255.0f; }
        }
        float *dst = hbuf + (size_t) y * row_f;
        if(ax.enlarge)
        {
            for(int x = 0; x < (int) ow; ++x)
            {
                const float *g = ax.coef + (size_t) ax.width * x;
                int k = 0;
                for(int t = ax.span[x].n0; t <= ax.span[x].n1; ++t, ++k)
                {
                    for(int c = 0; c < C; ++c) { dst[x * C + c] += d0[t * C + c] * g[k]; }
                }
            }
        }
        else
        {
            for(int x = 0; x < ax.count; ++x)
            {
                const float *g = ax.coef + (size_t) ax.width * x;
                const int in_x = x - ax.margin;
                for(int k = ax.span[x].n0; k <= ax.span[x].n1; ++k)
                {
                    for(int c = 0; c < C; ++c) { dst[k * C + c] += d0[in_x * C + c] * g[k - ax.span[x].n0]; }
                }
            }
        }
    }
    /* vertical pass */
    float *acc = (float *) calloc((size_t) oh * row_f, sizeof(float));
    if(ay.enlarge)
    {
        for(int y = 0; y < (int) oh; ++y)
        {
            const float *g = ay.coef + (size_t) ay.width * y;
            float *dst = acc + (size_t) y * row_f;
            int k = 0;
            for(int t = ay.span[y].n0; t <= ay.span[y].n1; ++t, ++k)
            {
                const float *src = hbuf + (size_t) clampi(t, 0, (int) h - 1) * row_f;
                for(size_t i = 0; i < row_f; ++i) { dst[i] += src[i] * g[k]; }
            }
        }
    }
    else
    {
        for(int j = 0; j < ay.count; ++j)
        {
            const float *g = ay.coef + (size_t) ay.width * j;
            const float *src = hbuf + (size_t) clampi(j - ay.margin, 0, (int) h - 1) * row_f;
            for(int k = ay.span[j].n0; k <= ay.span[j].n1; ++k)
            {
                float *dst = acc + (size_t) k * row_f;
                const float cf = g[k - ay.span[j].n0];
                for(size_t i = 0; i < row_f; ++i) { dst[i] += src[i] * cf; }
            }
        }
    }
    /* encode: (unsigned char)(int)(saturate(f) * 255.0f + 0.5), the addition in double */
    for(size_t i = 0; i < (size_t) oh * row_f; ++i)
    {
        float f = acc[i];
        if(f < 0) { f = 0; }
        else if(f > 1) { f = 1; }
        out[i] = (uint8_t) (int) ((f * 255.0f) + 0.5);
    }
    free(acc);
    free(dec);
    free(hbuf);
    axis_free(&ax);
    axis_free(&ay);
}
