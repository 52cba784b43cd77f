// channel.cuh — LLR quantiser, synthetic BPSK/AWGN channel and BER/FER counters on the device.
//
// quantize_kernel      : q = clamp((int)(scale*y), -sat, sat)      (ref: code/x86/CFixPointConversion/CFastFixConversion.cpp:55-65;
//                        GPU twin code/gpu_fixed/decoder_template/GPU_Scheduled_functions.cu:54-64)
// awgn_kernel          : all-zero codeword, BPSK 0 -> -1, y = -1 + sigma*n, Box-Muller, then the same quantiser, fused
//                        (ref: GenerateNoiseAndTransform code/gpu_fixed/awgn_channel/CChanel_AWGN_SIMD.cu:7-30).  The reference
//                        draws from cuRAND XORWOW seed 1234 / MKL MT2203, neither reproducible here; this generator
//                        is counter-based (Philox4x32-10 keyed by seed, counter = frame and position) so any frame can be
//                        regenerated independently on any GPU — what frame sharding across ranks needs.
// count_errors_kernel  : bit/frame errors over the first n - n_checks positions (ref: code/gpu_fixed/ber_analyzer/CErrorAnalyzer.cpp:142-150)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ldpcb200 {

// (int)x as the reference's x86 build computes it: truncation toward zero, and the "integer indefinite" INT_MIN for NaN or
// |x| >= 2^31 (cvttss2si), so that even absurd inputs quantise like the reference (to -sat).
__device__ __forceinline__ int quantize_one(float p, int sat)
{
    int v = (fabsf(p) < 2147483648.0f) ? __float2int_rz(p) : (int)0x80000000;
    v = max(v, -sat);
    return min(v, sat);
}

static __global__ void quantize_kernel(const float* __restrict__ y, int8_t* __restrict__ q, size_t count, float scale, int sat)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        q[i] = (int8_t)quantize_one(__fmul_rn(scale, y[i]), sat);
    }
}

__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// channel value -> the decoder's input type: quantised int8 / int16, or the raw float y (norm_channel = false: no 2/sigma^2
// scaling, ref: code/x86/main_p.cpp:124)
template <class S> __device__ __forceinline__ S awgn_out(float y, float scale, int sat) { return (S)quantize_one(__fmul_rn(scale, y), sat); }
template <> __device__ __forceinline__ float awgn_out<float>(float y, float, int) { return y; }

// one thread = 4 consecutive positions of one frame
template <class S>
__global__ void awgn_kernel(S* __restrict__ q, size_t frames, int n, float sigma, uint64_t seed, uint64_t first_frame, float scale, int sat)
{
    const int quads = (n + 3) / 4;
    const size_t total = frames * (size_t)quads;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f = i / quads;
        const int p = (int)(i % quads) * 4;
        const uint64_t gf = first_frame + f;
        uint32_t c[4] = { (uint32_t)gf, (uint32_t)(gf >> 32), (uint32_t)p, 0x4C445043u };
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        float g[4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const float u1 = ((float)(c[2 * h] >> 8) + 0.5f) * (1.0f / 16777216.0f);        // (0,1)
            const float u2 = ((float)(c[2 * h + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
            const float r = sqrtf(-2.0f * __logf(u1));
            float sn, cs; __sincosf(6.283185307179586f * u2, &sn, &cs);
            g[2 * h] = r * cs; g[2 * h + 1] = r * sn;
        }
#pragma unroll
        for (int b = 0; b < 4; b++) {
            if (p + b < n) {
                const float y = -1.0f + sigma * g[b];
                q[f * (size_t)n + p + b] = awgn_out<S>(y, scale, sat);
            }
        }
    }
}

// one warp per frame; out[0] += bit errors, out[1] += frame errors
static __global__ void count_errors_kernel(const uint8_t* __restrict__ hard, size_t frames, int n, int k_info, int packed, unsigned long long* out)
{
    const int lane = threadIdx.x & 31;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long be = 0, fe = 0;
    for (size_t f = warp; f < frames; f += nwarps) {
        unsigned int cnt = 0;
        if (!packed) { for (int i = lane; i < k_info; i += 32) cnt += hard[f * (size_t)n + i] & 1u; }
        else {
            const int nb = (n + 7) / 8;
            for (int b = lane; b * 8 < k_info; b += 32) {
                unsigned int v = hard[f * (size_t)nb + b];
                const int valid = min(8, k_info - b * 8);
                v &= (1u << valid) - 1u;
                cnt += __popc(v);
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
        be += cnt; fe += (cnt != 0);
    }
    if (lane == 0 && (be | fe)) { atomicAdd(out, be); atomicAdd(out + 1, fe); }
}

}  // namespace ldpcb200
