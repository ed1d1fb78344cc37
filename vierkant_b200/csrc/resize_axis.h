// resize_axis.h -- host-side tap lists of the stbir-exact resize (see resize_core.cuh for the design).  Plain C++, no
// CUDA: shared by the C-ABI shim and by the GPU-less host emulation harness (tests/host_emul).
// Build with -ffp-contract=off: the coefficient arithmetic must round exactly like the reference's x86-64 build.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace vkt
{

// ------------------------------------------------------------------------------------------------ host: tap lists
#if defined(__CUDACC__)
#define VKT_RESIZE_HD __host__ __device__
#else
#define VKT_RESIZE_HD
#endif
// stbir decodes a sample as u8 / 255.0f (:1252-1291).  The correctly rounded quotient without a division: one Newton step on
// v * (1/255) -- q1 = fma(fma(-q0, 255, v), 1/255, q0) -- equals v / 255.0f for all 256 values (checked exhaustively in
// tests/test_host_emul.py); four FMA-pipe instructions instead of a ~10-instruction IEEE division or a shared-memory
// lookup (the pass is bound by its load/store instructions).
VKT_RESIZE_HD inline float resize_decode_u8(uint32_t v)
{
    const float fv = (float) v, r = 1.0f / 255.0f;
#if defined(__CUDA_ARCH__)
    const float q0 = __fmul_rn(fv, r);
    return __fmaf_rn(__fmaf_rn(-q0, 255.0f, fv), r, q0);
#else
    const float q0 = fv * r;
    return fmaf(fmaf(-q0, 255.0f, fv), r, q0);
#endif
}

// The same quotient from a sample that already is a float (the fused pass builds float(v) with a byte permute and one add):
// v / 255 = v * hi + v * lo with hi = RN(1/255), lo = RN(1/255 - hi).  fma(v, hi, RN(v * lo)) rounds once, from a value that is
// exact to ~2^-50 -- correctly rounded for all 256 values (exhaustive: tests/test_host_emul.py); two FMA-pipe instructions.
VKT_RESIZE_HD inline float resize_decode_split(float fv)
{
    const float hi = 1.0f / 255.0f;                                          // 0x3B808081
    const float lo = (float) (1.0 / 255.0 - (double) (1.0f / 255.0f));// 0xAF7EFEFF
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(fv, hi, __fmul_rn(fv, lo));
#else
    return fmaf(fv, hi, fv * lo);
#endif
}

// stbir's encode of a saturated sample s = sat(f) * 255.0f in [0, 255] (:1737-1763): (int)((double) s + 0.5), i.e. floor(s + 0.5)
// of the EXACT sum.  Two additions that round toward zero do the same without a conversion: RZ(s + 0.5) never crosses an
// integer upwards (integers are representable), so its floor is that of the exact sum, and RZ(. + 2^23) leaves that floor in the
// low mantissa bits.  Exhaustively equal for every float in [0, 255] (tests/test_host_emul.py).
VKT_RESIZE_HD inline uint32_t resize_encode_u8(float s)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(__fadd_rz(__fadd_rz(s, 0.5f), 8388608.0f)) & 255u;
#else
    auto rz = [](double x) {// x rounded to float toward zero
        float f = (float) x;
        if(std::fabs((double) f) > std::fabs(x)) { f = std::nextafterf(f, 0.0f); }
        return f;
    };
    const float u = rz((double) rz((double) s + 0.5) + 8388608.0);
    uint32_t bits;
    std::memcpy(&bits, &u, sizeof(bits));
    return bits & 255u;
#endif
}

struct ResizeAxis
{
    int in_size = 0, out_size = 0;
    std::vector<int> start;  // out_size + 1 offsets into idx / coef
    std::vector<int> idx;    // input sample, already clamped to [0, in_size)
    std::vector<float> coef;

    static float catmullrom(float x)// stbir__filter_catmullrom :816-829
    {
        x = (float) fabs(x);
        if(x < 1.0f) { return 1 - x * x * (2.5f - 1.5f * x); }
        else if(x < 2.0f) { return 2 - x * (4 + x * (0.5f * x - 2.5f)); }
        return 0.0f;
    }
    static float mitchell(float x)// stbir__filter_mitchell :831-844
    {
        x = (float) fabs(x);
        if(x < 1.0f) { return (16 + x * x * (21 * x - 36)) / 18; }
        else if(x < 2.0f) { return (32 + x * (-60 + x * (36 - 7 * x))) / 18; }
        return 0.0f;
    }

    void build(int in, int out)
    {
        in_size = in, out_size = out;
        const float support = 2.0f;// stbir__support_two for both default filters
        const float scale = ((float) out / in) / (1.0f - 0.0f);// stbir__calculate_transform :2233
        const float shift = 0.0f * out / (1.0f - 0.0f);
        const bool enlarge = scale > 1;
        const int width = (int) ceil(support * 2);// slots per list in the reference's flat coefficient table
        const int margin = (enlarge ? (int) ceil(support * 2) : (int) ceil(support * 2 / scale)) / 2;
        const int lists = enlarge ? out : in + 2 * margin;
        std::vector<int> n0(size_t(lists) + 1), n1(size_t(lists) + 1), last0(size_t(lists) + 1);
        std::vector<float> flat((size_t(lists) + 3) * size_t(width), 0.0f);
        auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
        start.assign(size_t(out) + 1, 0);
        idx.clear(), coef.clear();

        if(enlarge)
        {
            const float radius = support * scale;
            for(int n = 0; n < lists; ++n)
            {
                const float centre = (float) n + 0.5f;
                const float lo = centre - radius, hi = centre + radius;
                const float in_lo = (lo + shift) / scale, in_hi = (hi + shift) / scale;
                const float in_centre = (centre + shift) / scale;
                int first = (int) (floor(in_lo + 0.5));// double arithmetic, as the reference
                const int last = (int) (floor(in_hi - 0.5));
                float *g = flat.data() + size_t(width) * n;
                float total = 0;
                n0[n] = first, n1[n] = last;
                for(int i = 0; i <= last - first; i++)
                {
                    const float tap_centre = (float) (i + first) + 0.5f;
                    g[i] = catmullrom(in_centre - tap_centre);
                    if(i == 0 && !g[i])
                    {
                        n0[n] = ++first;
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
                    n1[n] = n0[n] + i - 1;
                }
            }
            // the passes read the table only after all lists were written (a spilled 5th entry has been overwritten)
            for(int n = 0; n < out; ++n)
            {
                start[n] = (int) idx.size();
                const float *g = flat.data() + size_t(width) * n;
                for(int t = n0[n], k = 0; t <= n1[n]; ++t, ++k)
                {
                    idx.push_back(clampi(t, 0, in - 1));
                    coef.push_back(g[k]);
                }
            }
            start[out] = (int) idx.size();
            return;
        }

        const float radius = support / scale;
        for(int n = 0; n < lists; ++n)
        {
            const float centre = (float) (n - margin) + 0.5f;
            const float lo = centre - radius, hi = centre + radius;
            const float out_lo = lo * scale - shift, out_hi = hi * scale - shift;
            const float out_centre = centre * scale - shift;
            const int first = (int) (floor(out_lo + 0.5)), last = (int) (floor(out_hi - 0.5));
            float *g = flat.data() + size_t(width) * n;
            n0[n] = first, n1[n] = last, last0[n] = last;
            for(int i = 0; i <= last - first; i++)
            {
                const float x = ((float) (i + first) + 0.5f) - out_centre;
                g[i] = mitchell(x) * scale;
            }
            for(int i = last - first; i >= 0; i--)
            {
                if(g[i]) { break; }
                n1[n] = n0[n] + i - 1;
            }
        }
        // per-output normalisation, :1126-1160.  The reference scans every list from 0 and stops at the first one that
        // starts after i; lists whose (untrimmed, monotone) range ended before i cannot qualify, so the scan may start
        // at the first list that can still reach i -- the visited qualifying lists and their order are the same.
        {
            int lo_list = 0;
            for(int i = 0; i < out; i++)
            {
                while(lo_list < lists && last0[lo_list] < i) { ++lo_list; }
                float total = 0;
                for(int j = lo_list; j < lists; j++)
                {
                    if(i >= n0[j] && i <= n1[j]) { total += flat[size_t(width) * j + size_t(i - n0[j])]; }
                    else if(i < n0[j]) { break; }
                }
                const float norm = 1 / total;
                for(int j = lo_list; j < lists; j++)
                {
                    if(i >= n0[j] && i <= n1[j]) { flat[size_t(width) * j + size_t(i - n0[j])] *= norm; }
                    else if(i < n0[j]) { break; }
                }
            }
        }
        // drop leading zeros and outputs left of the image, :1162-1199
        for(int j = 0; j < lists; j++)
        {
            float *g = flat.data() + size_t(width) * j;
            int skip = 0;
            while(g[skip] == 0 && (g + skip) < flat.data() + flat.size() - 1) { skip++; }
            n0[j] += skip;
            while(n0[j] < 0)
            {
                n0[j]++;
                skip++;
            }
            const int range = n1[j] - n0[j] + 1;
            const int max = width < range ? width : range;
            for(int i = 0; i < max; i++)
            {
                if(i + skip >= width) { break; }
                g[i] = g[i + skip];
            }
        }
        for(int j = 0; j < lists; j++) { n1[j] = n1[j] < out - 1 ? n1[j] : out - 1; }
        // scatter -> gather: output k receives list j's entry (k - n0[j]); lists are visited in ascending j, which is the
        // order in which the reference's scatter loops add into output k
        std::vector<int> count(size_t(out) + 1, 0);
        for(int j = 0; j < lists; j++)
        {
            for(int k = n0[j]; k <= n1[j]; ++k) { count[size_t(k)]++; }
        }
        for(int k = 0; k < out; ++k) { start[size_t(k) + 1] = start[size_t(k)] + count[size_t(k)]; }
        idx.assign(size_t(start[out]), 0), coef.assign(size_t(start[out]), 0.0f);
        std::vector<int> fill(start.begin(), start.end() - 1);
        for(int j = 0; j < lists; j++)
        {
            const float *g = flat.data() + size_t(width) * j;
            for(int k = n0[j]; k <= n1[j]; ++k)
            {
                const int o = fill[size_t(k)]++;
                idx[size_t(o)] = clampi(j - margin, 0, in - 1);
                coef[size_t(o)] = g[k - n0[j]];
            }
        }
    }
};

}// namespace vkt
