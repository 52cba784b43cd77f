// pipes.cu — instruction-throughput microbenchmark for the op mix an LDPC min-sum row update can be built from.
// Measures warp-instructions per clock per SM for candidate s16x2 / f16x2 / scalar ops on sm_100a, so the
// decode kernel's arithmetic representation is chosen from numbers.  Not product code.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define ITERS 1024
#define REP 8
#define NCHAIN 8   // independent dependency chains per thread

#define DEF_KERNEL(NAME, INIT, BODY)                                                     \
__global__ void __launch_bounds__(1024) k_##NAME(unsigned* out, unsigned seed, long long* clk) { \
    unsigned r[NCHAIN];                                                                  \
    _Pragma("unroll") for (int c = 0; c < NCHAIN; c++) r[c] = seed * (threadIdx.x + 1) + c * 0x01010101u; \
    unsigned k1 = seed ^ 0x00070007u, k2 = seed | 0x001f001fu; asm volatile("" : "+r"(k1), "+r"(k2)); INIT;                    \
    long long t0 = clock64();                                                            \
    _Pragma("unroll 1") for (int i = 0; i < ITERS; i++) {                                \
        _Pragma("unroll") for (int rep = 0; rep < REP; rep++) {                          \
        _Pragma("unroll") for (int c = 0; c < NCHAIN; c++) { unsigned& x = r[c]; BODY; } } \
    }                                                                                    \
    long long t1 = clock64();                                                            \
    unsigned acc = 0;                                                                    \
    _Pragma("unroll") for (int c = 0; c < NCHAIN; c++) acc ^= r[c];                      \
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;                                    \
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;                                     \
}

static __device__ __forceinline__ unsigned h2u(__half2 h) { return *reinterpret_cast<unsigned*>(&h); }
static __device__ __forceinline__ __half2 u2h(unsigned u) { return *reinterpret_cast<__half2*>(&u); }

DEF_KERNEL(lop3,      , x = (x ^ k1) & r[(c + 1) % NCHAIN])
DEF_KERNEL(iadd,      , x = x + r[(c + 1) % NCHAIN] + k1)
DEF_KERNEL(imad,      , x = x * k1 + k2)
DEF_KERNEL(prmt,      , x = __byte_perm(x, k1, 0x5140))
DEF_KERNEL(vimnmx16,  , x = (rep & 1) ? __vmins2(x, k1) : __vmaxs2(x, k2))
DEF_KERNEL(vimnmx3_16,, x = __vimin3_s16x2(x, k1, k2))
DEF_KERNEL(viaddmnmx16,, x = __viaddmin_s16x2(x, k1, k2))
DEF_KERNEL(viaddmnmx16relu,, x = __viaddmin_s16x2_relu(x, k1, k2))
DEF_KERNEL(viadd16,   , x = __vadd2(x, k1))
DEF_KERNEL(vimnmx32,  , x = (rep & 1) ? (unsigned)min((int)x, (int)k1) : (unsigned)max((int)x, (int)k2))
DEF_KERNEL(viaddmnmx32,, x = (unsigned)__viaddmin_s32((int)x, (int)k1, (int)k2))
DEF_KERNEL(vabsdiff4, , x = __vabsdiffu4(x, k1))
DEF_KERNEL(hadd2,     , x = h2u(__hadd2(u2h(x), u2h(k1))))
DEF_KERNEL(hfma2,     , x = h2u(__hfma2(u2h(x), u2h(k1), u2h(k2))))
DEF_KERNEL(hmnmx2,    , x = (rep & 1) ? h2u(__hmin2(u2h(x), u2h(k1))) : h2u(__hmax2(u2h(x), u2h(k2))))
DEF_KERNEL(hmnmx2abs, , x = h2u(__hmin2(__habs2(u2h(x)), u2h(k1))))
DEF_KERNEL(hset2eq,   , x = __heq2_mask(u2h(x), u2h(k1)))
DEF_KERNEL(hset2gt,   , x = __hgt2_mask(u2h(x), u2h(k1)))
DEF_KERNEL(fadd,      , x = __float_as_uint(__uint_as_float(x) + __uint_as_float(k1)))
DEF_KERNEL(fmnmx,     , x = (rep & 1) ? __float_as_uint(fminf(__uint_as_float(x), __uint_as_float(k1))) : __float_as_uint(fmaxf(__uint_as_float(x), __uint_as_float(k2))))
DEF_KERNEL(shf,       , x = __funnelshift_l(x, k1, 3))
DEF_KERNEL(sel,       , x = (x > k1) ? r[(c + 1) % NCHAIN] : x)   // ISETP + SEL
// mixes: one ALU-pipe op + one FMA-pipe op per step
DEF_KERNEL(mix_lop3_imad,   , x = ((x ^ k1) & r[(c + 1) % NCHAIN]); x = x * k1 + k2)
DEF_KERNEL(mix_hmnmx_hfma,  , x = h2u(__hmin2(u2h(x), u2h(k1))); asm volatile("" : "+r"(x)); x = h2u(__hfma2(u2h(x), u2h(k1), u2h(k2))))
DEF_KERNEL(mix_vimnmx_imad, , x = __vmins2(x, k1); asm volatile("" : "+r"(x)); x = x * k1 + k2)
DEF_KERNEL(mix_lop3_hfma,   , x = (x ^ k1) & r[(c + 1) % NCHAIN]; x = h2u(__hfma2(u2h(x), u2h(k1), u2h(k2))))

// shared-memory pipe: one LDS.32 + one STS.32 per step, conflict-free
__global__ void __launch_bounds__(1024) k_lds_sts(unsigned* out, unsigned seed, long long* clk) {
    extern __shared__ unsigned sm[];
    const int T = blockDim.x;
    for (int i = threadIdx.x; i < 16 * T; i += T) sm[i] = seed + i;
    __syncthreads();
    unsigned acc = seed;
    long long t0 = clock64();
    for (int i = 0; i < ITERS * REP; i++) {
        #pragma unroll
        for (int c = 0; c < NCHAIN; c++) {
            unsigned v = sm[((c + i) & 15) * T + threadIdx.x];
            sm[((c + i + 8) & 15) * T + threadIdx.x] = v + acc;
            acc ^= v;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void __launch_bounds__(1024) k_lds_only(unsigned* out, unsigned seed, long long* clk) {
    extern __shared__ unsigned sm[];
    const int T = blockDim.x;
    for (int i = threadIdx.x; i < 16 * T; i += T) sm[i] = seed + i;
    __syncthreads();
    unsigned acc = seed;
    long long t0 = clock64();
    for (int i = 0; i < ITERS * REP; i++) {
        #pragma unroll
        for (int c = 0; c < NCHAIN; c++) acc ^= sm[((c + i) & 15) * T + threadIdx.x];
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <typename K>
static void run(const char* name, K kern, int threads, int ops_per_step, size_t smem, unsigned* d_out, long long* d_clk, int sms) {
    const int blocks = sms;   // one CTA per SM
    if (smem) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<blocks, threads, smem>>>(d_out, 0x00010003u, d_clk);   // warm-up
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<blocks, threads, smem>>>(d_out, 0x00010003u, d_clk);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%-22s ERROR %s\n", name, cudaGetErrorString(err)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long clk[1024]; cudaMemcpy(clk, d_clk, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += (double)clk[i]; avg /= blocks;
    double warp_instr = (double)ITERS * REP * NCHAIN * ops_per_step * (threads / 32);
    printf("{\"op\": \"%s\", \"threads\": %d, \"warp_instr_per_clk_per_sm\": %.3f, \"lanes_per_clk_per_sm\": %.1f, \"ms\": %.3f, \"gops_chip\": %.1f, \"cycles\": %.0f}\n",
           name, threads, warp_instr / avg, 32.0 * warp_instr / avg, ms, 32.0 * warp_instr * blocks / (ms * 1e6), avg);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
    unsigned* d_out; long long* d_clk;
    cudaMalloc(&d_out, sizeof(unsigned) * 1024 * 1024); cudaMalloc(&d_clk, sizeof(long long) * 1024);
    const int sms = p.multiProcessorCount;
    for (int threads : {256, 1024}) {
#define RUN(NAME, OPS) run(#NAME, k_##NAME, threads, OPS, 0, d_out, d_clk, sms)
        RUN(lop3, 1); RUN(iadd, 1); RUN(imad, 1); RUN(prmt, 1); RUN(shf, 1); RUN(sel, 2);
        RUN(vimnmx16, 1); RUN(vimnmx3_16, 1); RUN(viaddmnmx16, 1); RUN(viaddmnmx16relu, 1); RUN(viadd16, 1);
        RUN(vimnmx32, 1); RUN(viaddmnmx32, 1); RUN(vabsdiff4, 1);
        RUN(hadd2, 1); RUN(hfma2, 1); RUN(hmnmx2, 1); RUN(hmnmx2abs, 1); RUN(hset2eq, 1); RUN(hset2gt, 1);
        RUN(fadd, 1); RUN(fmnmx, 1);
        RUN(mix_lop3_imad, 2); RUN(mix_hmnmx_hfma, 2); RUN(mix_vimnmx_imad, 2); RUN(mix_lop3_hfma, 2);
        run("lds_sts", k_lds_sts, threads, 2, sizeof(unsigned) * 16 * threads, d_out, d_clk, sms);
        run("lds_only", k_lds_only, threads, 1, sizeof(unsigned) * 16 * threads, d_out, d_clk, sms);
    }
    return 0;
}
