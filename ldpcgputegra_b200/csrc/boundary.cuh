// boundary.cuh — layout conversion between the reference's frame-major buffers and the frame-parallel kernel's interleaved
// state, with the hard decision fused into the way out.  Included by ldpc_b200.cu only (non-template kernels).
#pragma once
#include "rowops.cuh"

namespace ldpcb200 {

// ---- layout conversion at the boundary ------------------------------------------------------------------------------
// frame-major int8 [F][N]  ->  V[n][T] biased bytes.  Replaces Interleaver_uint8 (ref: code/gpu_fixed/transpose/
// GPU_Transpose_uint8.cu:80-130) with no T%32 / N%128 restriction.  Tile = 128 frames x 128 variables; 4x4 byte
// transposes in registers (PRMT) and an XOR-swizzled 32-bit shared tile, conflict-free on both sides.
__device__ __forceinline__ void transpose4x4(const uint32_t g[4], uint32_t o[4])
{
    const uint32_t t0 = __byte_perm(g[0], g[1], 0x5140), t1 = __byte_perm(g[2], g[3], 0x5140);
    const uint32_t t2 = __byte_perm(g[0], g[1], 0x7362), t3 = __byte_perm(g[2], g[3], 0x7362);
    o[0] = __byte_perm(t0, t1, 0x5410); o[1] = __byte_perm(t0, t1, 0x7632);
    o[2] = __byte_perm(t2, t3, 0x5410); o[3] = __byte_perm(t2, t3, 0x7632);
}

__global__ void __launch_bounds__(256) interleave_kernel(const int8_t* __restrict__ llr, uint32_t* __restrict__ V, size_t frames, int n, int T, int lo, int hi)
{
    __shared__ uint32_t tile[128 * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t f0 = (size_t)blockIdx.x * 128;
    const int n0 = blockIdx.y * 128;
    const bool fast = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(llr) & 3) == 0);
    for (int tw = warp; tw < 32; tw += 8) {
        uint32_t g[4], o[4];
        const int nn = n0 + 4 * lane;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const size_t f = f0 + 4 * tw + k;
            uint32_t w = 0u;
            if (f < frames && nn < n) {
                if (fast) w = __ldg(reinterpret_cast<const uint32_t*>(llr + f * n + nn));
                else { for (int b = 0; b < 4; b++) if (nn + b < n) w |= (uint32_t)(uint8_t)llr[f * n + nn + b] << (8 * b); }
            }
            g[k] = bias_bytes(w, lo, hi);
        }
        transpose4x4(g, o);
#pragma unroll
        for (int i = 0; i < 4; i++) tile[(4 * lane + i) * 32 + (tw ^ lane)] = o[i];
    }
    __syncthreads();
    const int t0 = blockIdx.x * 32;
    for (int r = warp; r < 128; r += 8) {
        const int nn = n0 + r;
        if (nn < n && t0 + lane < T) V[(size_t)nn * T + t0 + lane] = tile[r * 32 + (lane ^ ((r >> 2) & 31))];
    }
}

// V[n][T] -> hard decisions, frame-major.  Replaces InvInterleaver_uint8 and its fused vsetgts4 (ref: GPU_Transpose_uint8.cu:9-78).
// PACKED=false: one byte per bit in {0,1} (the reference's output).  PACKED=true: LSB-first bits, ceil(n/8) bytes per frame.
template <bool PACKED>
__global__ void __launch_bounds__(256) deinterleave_hard_kernel(const uint32_t* __restrict__ V, uint8_t* __restrict__ hard, size_t frames, int n, int T, int lo)
{
    __shared__ uint32_t tile[128 * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t f0 = (size_t)blockIdx.x * 128;
    const int n0 = blockIdx.y * 128;
    const int t0 = blockIdx.x * 32;
    const uint32_t thr = (uint32_t)((-lo) & 0xFF) * 0x01010101u;   // v > 0  <=>  biased byte > -lo
    for (int r = warp; r < 128; r += 8) {
        const int nn = n0 + r;
        uint32_t w = 0u;
        if (nn < n && t0 + lane < T) w = __vcmpgtu4(V[(size_t)nn * T + t0 + lane], thr) & 0x01010101u;
        tile[r * 32 + (lane ^ ((r >> 2) & 31))] = w;
    }
    __syncthreads();
    if (!PACKED) {
        const bool fast = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(hard) & 3) == 0);
        for (int tw = warp; tw < 32; tw += 8) {
            uint32_t g[4], o[4];
#pragma unroll
            for (int i = 0; i < 4; i++) g[i] = tile[(4 * lane + i) * 32 + (tw ^ lane)];
            transpose4x4(g, o);   // o[k] = bytes (n..n+3) of frame 4tw+k
            const int nn = n0 + 4 * lane;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const size_t f = f0 + 4 * tw + k;
                if (f < frames && nn < n) {
                    if (fast) *reinterpret_cast<uint32_t*>(hard + f * n + nn) = o[k];
                    else { for (int b = 0; b < 4; b++) if (nn + b < n) hard[f * n + nn + b] = (uint8_t)(o[k] >> (8 * b)); }
                }
            }
        }
    } else {
        const int nb = (n + 7) / 8;
        for (int tw = warp; tw < 32; tw += 8) {
#pragma unroll
            for (int q = 0; q < 4; q++) {          // 4 groups of 32 variables
                const int r = 32 * q + lane;
                const uint32_t w = tile[r * 32 + (tw ^ ((r >> 2) & 31))];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t bits = __ballot_sync(0xFFFFFFFFu, (w >> (8 * k)) & 1u);
                    const size_t f = f0 + 4 * tw + k;
                    const int byte0 = (n0 + 32 * q) / 8;
                    if (lane < 4 && f < frames && byte0 + lane < nb) hard[f * nb + byte0 + lane] = (uint8_t)(bits >> (8 * lane));
                }
            }
        }
    }
}

// parity-check access: V/MSG -> frame-major int8 posteriors [F][n] and messages [F][m]
__global__ void fp_debug_state_kernel(const uint32_t* __restrict__ V, const uint32_t* __restrict__ MSG, int8_t* post, int8_t* msgs,
                                      size_t frames, int n, int m, int T, int lo, int have_msgs)
{
    const size_t total = frames * (size_t)(n + m);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f = i / (size_t)(n + m);
        const int k = (int)(i % (size_t)(n + m));
        const int t = (int)(f >> 2), b = (int)(f & 3);
        if (k < n) { if (post) post[f * n + k] = (int8_t)((int)((V[(size_t)k * T + t] >> (8 * b)) & 0xFF) + lo); }
        else if (msgs) { const int e = k - n; msgs[f * (size_t)m + e] = have_msgs ? (int8_t)((int)((MSG[(size_t)e * T + t] >> (8 * b)) & 0xFF) - 128) : (int8_t)0; }
    }
}

// the same for the staged kernel's COMPRESSED messages (kernel_fp.cuh: fp_row_math_c; four words per row and thread): every edge's
// message is re-expanded, so ldpc_b200_debug_state returns the bytes the uncompressed kernels hold.  edge_row[e] = row << 4 | slot.
__global__ void fc_debug_state_kernel(const uint32_t* __restrict__ V, const uint32_t* __restrict__ MSGC, const uint32_t* __restrict__ edge_row,
                                      int8_t* post, int8_t* msgs, size_t frames, int n, int m, int T, int lo, int have_msgs)
{
    const size_t total = frames * (size_t)(n + m);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f = i / (size_t)(n + m);
        const int k = (int)(i % (size_t)(n + m));
        const int t = (int)(f >> 2), b = (int)(f & 3);
        if (k < n) { if (post) post[f * n + k] = (int8_t)((int)((V[(size_t)k * T + t] >> (8 * b)) & 0xFF) + lo); }
        else if (msgs) {
            const int e = k - n;
            int v = 0;
            if (have_msgs) {
                const uint32_t er = edge_row[e], row = er >> 4, j = er & 15u, g = (uint32_t)b >> 1, sh = 16u * ((uint32_t)b & 1u);
                const uint32_t cw = (MSGC[(size_t)(4 * row + g) * T + t] >> sh) & 0xFFFFu, es = (MSGC[(size_t)(4 * row + 2 + g) * T + t] >> sh) & 0xFFFFu;
                const int mag = (int)(((es >> j) & 1u) ? (cw >> 8) : (cw & 0xFFu));
                v = ((es >> (8 + j)) & 1u) ? -mag : mag;
            }
            msgs[f * (size_t)m + e] = (int8_t)v;
        }
    }
}

}  // namespace ldpcb200
