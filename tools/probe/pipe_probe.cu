// pipe_probe.cu -- measured issue rates of the integer instructions the BC7 kernels are made of, on this GPU.
// Eight independent chains per thread, 256-thread CTAs; reported as warp instructions per clock per SM sub-partition
// (1.0 = the scheduler's issue limit).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ROUNDS 8
template<int OP>
__device__ __forceinline__ void step(uint32_t &a, uint32_t &b, uint32_t k)
{
    if(OP == 0) { a = a * b + k; }                                            // IMAD r*r+r
    if(OP == 1) { a = a * 0x9E3779B1u + b; }                                  // IMAD r*imm+r
    if(OP == 2) { a = a ^ (b | k); }                                          // LOP3
    if(OP == 3) { a = __funnelshift_r(a, b, 8); }                             // SHF
    if(OP == 4) { a = __viaddmin_s32((int) a, (int) b, 0x7FFFFFF0); }         // VIADDMNMX
    if(OP == 5) { a = __dp4a(a, b, k); }                                      // IDP.4A
    if(OP == 6) { a = min(min(a, b), k) + 1; }                                // VIMNMX3 (+ IADD)
    if(OP == 7) { a = __byte_perm(a, b, 0x6420); }                            // PRMT
    if(OP == 8) { a = (a >= k) ? b : a; }                                     // ISETP + SEL
    if(OP == 9) { a = a * b + k; b = b ^ a; }                                 // IMAD + LOP3 (balanced)
    if(OP == 10) { a = (uint32_t) ((int) (a - b) >> 8); }                     // IADD + SHF
    if(OP == 11) { a = __mulhi((int) a, 1 << 24) + b; }                       // IMAD.HI
    if(OP == 12) { int d = (int) (a - k) >> 8; b += (uint32_t) (d * d) * 103u; a += b; }// metric term: sub, shf, imad, imad, (add)
    if(OP == 13) { a = a * b + k; b = (uint32_t) ((int) b >> 3) + a; }        // IMAD + SHF + IADD
    if(OP == 15) { asm("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(a) : "r"(k), "r"(b)); }                     // IDP.2A.LO.S16.U8
    if(OP == 16) { asm("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(a) : "r"(k), "r"(b)); b = b ^ a; }         // IDP.2A + LOP3
    if(OP == 17) { a = __dp4a(a, b, k); b = b ^ a; }                                                          // IDP.4A + LOP3
    if(OP == 14) { float f = __uint_as_float(a); f = __fmaf_rn(f, 1.0001f, 0.5f); a = __float_as_uint(f); }// FFMA
}
template<int OP>
__global__ void __launch_bounds__(256) probe(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t a[CHAINS], b[CHAINS];
#pragma unroll
    for(int k = 0; k < CHAINS; ++k) { a[k] = seed + threadIdx.x * 8u + k, b[k] = seed * 3u + k + blockIdx.x; }
#pragma unroll 1
    for(int i = 0; i < iters; ++i)
    {
#pragma unroll
        for(int r = 0; r < ROUNDS; ++r)
        {
#pragma unroll
            for(int k = 0; k < CHAINS; ++k) { step<OP>(a[k], b[k], seed + r); }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for(int k = 0; k < CHAINS; ++k) { acc += a[k] ^ b[k]; }
    if(acc == 0x12345u) { out[blockIdx.x * 256 + threadIdx.x] = acc; }
}
template<int OP>
static void run(const char *name, int ops_per_step, int ctas_per_sm, int sms, double mhz, uint32_t *d)
{
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    probe<OP><<<sms * ctas_per_sm, 256>>>(d, 100, 1);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for(int rep = 0; rep < 3; ++rep)
    {
        cudaEventRecord(e0);
        probe<OP><<<sms * ctas_per_sm, 256>>>(d, iters, 17 + rep);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double warp_inst = double(sms) * ctas_per_sm * 8 /*warps*/ * iters * ROUNDS * CHAINS * ops_per_step;
    const double per_clk_smsp = warp_inst / (best * 1e-3) / (mhz * 1e6) / (sms * 4.0);
    printf("%-34s %d CTAs/SM: %6.3f ms  %.3f warp-inst/clk/SMSP (counting %d inst per step)\n", name, ctas_per_sm, best, per_clk_smsp, ops_per_step);
}
// dependent-issue latency: ONE chain per thread, one warp per scheduler
template<int OP>
__global__ void __launch_bounds__(128) latency(uint32_t *out, int iters, uint32_t seed, uint32_t w)
{
    uint32_t a = seed + threadIdx.x, b = seed * 3u + blockIdx.x;
#pragma unroll 1
    for(int i = 0; i < iters; ++i)
    {
#pragma unroll
        for(int r = 0; r < 64; ++r)
        {
            if(OP == 0) { a = a * b + seed; }                                   // IMAD r*r+r
            if(OP == 1) { a = a * w + b; }                                      // IMAD r*UR+r (w is a kernel parameter)
            if(OP == 2) { a = __funnelshift_r(a, b, 8); }                       // SHF
            if(OP == 3) { a = __viaddmin_s32((int) a, (int) b, 0x7FFFFFF0); }   // VIADDMNMX
            if(OP == 4) { int d = __viaddmin_s32((int) a, (int) b, 0x7FFFFFF0) >> 8; a = (uint32_t) d * ((uint32_t) d * w) + a; }// the metric chain: VIADDMNMX, SHF, IMAD(UR), IMAD
            if(OP == 5) { a = __dp4a(a, b, a); }                                // IDP.4A
            if(OP == 6) { a = (a >= b) ? seed : a + 1; }                        // ISETP + SEL (+ IADD)
        }
    }
    if(a == 0x12345u) { out[blockIdx.x * 128 + threadIdx.x] = a; }
}
template<int OP>
static void lat(const char *name, int inst_per_step, int sms, double mhz, uint32_t *d)
{
    const int iters = 2000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    latency<OP><<<sms, 128>>>(d, 10, 1, 3);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for(int rep = 0; rep < 3; ++rep)
    {
        cudaEventRecord(e0);
        latency<OP><<<sms, 128>>>(d, iters, 17 + rep, 103);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    const double cycles = best * 1e-3 * mhz * 1e6 / (double(iters) * 64);
    printf("latency %-40s %.2f cycles per step (%d dependent instructions: %.2f each)\n", name, cycles, inst_per_step, cycles / inst_per_step);
}
int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double mhz = khz / 1000.0;
    printf("%s, %d SMs, %.0f MHz (max)\n", p.name, p.multiProcessorCount, mhz);
    uint32_t *d;
    cudaMalloc(&d, 1 << 24);
    lat<0>("IMAD r*r+r", 1, p.multiProcessorCount, mhz, d);
    lat<1>("IMAD r*UR+r", 1, p.multiProcessorCount, mhz, d);
    lat<2>("SHF", 1, p.multiProcessorCount, mhz, d);
    lat<3>("VIADDMNMX", 1, p.multiProcessorCount, mhz, d);
    lat<4>("VIADDMNMX > SHF > IMAD(UR) > IMAD", 4, p.multiProcessorCount, mhz, d);
    lat<5>("IDP.4A", 1, p.multiProcessorCount, mhz, d);
    lat<6>("ISETP > SEL (> IADD)", 3, p.multiProcessorCount, mhz, d);
    for(int c: {3, 8})
    {
        run<0>("IMAD r*r+r", 1, c, p.multiProcessorCount, mhz, d);
        run<1>("IMAD r*imm+r", 1, c, p.multiProcessorCount, mhz, d);
        run<2>("LOP3", 1, c, p.multiProcessorCount, mhz, d);
        run<3>("SHF", 1, c, p.multiProcessorCount, mhz, d);
        run<4>("VIADDMNMX", 1, c, p.multiProcessorCount, mhz, d);
        run<5>("IDP.4A", 1, c, p.multiProcessorCount, mhz, d);
        run<6>("VIMNMX3 + IADD", 2, c, p.multiProcessorCount, mhz, d);
        run<7>("PRMT", 1, c, p.multiProcessorCount, mhz, d);
        run<8>("ISETP + SEL", 2, c, p.multiProcessorCount, mhz, d);
        run<9>("IMAD + LOP3", 2, c, p.multiProcessorCount, mhz, d);
        run<10>("IADD + SHF (fused: LEA.HI)", 1, c, p.multiProcessorCount, mhz, d);
        run<11>("mulhi 2^24 + add (fused: LEA.HI)", 1, c, p.multiProcessorCount, mhz, d);
        run<12>("metric term (sub shf imad imad add)", 5, c, p.multiProcessorCount, mhz, d);
        run<13>("IMAD + LEA", 2, c, p.multiProcessorCount, mhz, d);
        run<14>("FFMA", 1, c, p.multiProcessorCount, mhz, d);
        run<15>("IDP.2A", 1, c, p.multiProcessorCount, mhz, d);
        run<16>("IDP.2A + LOP3", 2, c, p.multiProcessorCount, mhz, d);
        run<17>("IDP.4A + LOP3", 2, c, p.multiProcessorCount, mhz, d);
    }
    return 0;
}
