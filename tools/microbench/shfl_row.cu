// shfl_row.cu — the check-node update of one layered min-sum row, two mappings, same arithmetic (rowops.cuh), same on-chip state:
//
//   in-lane  : one lane owns a whole row of one frame pair; the two-smallest search and the parity run over registers
//              (track2 / xor3) — the mapping kernel_rp.cuh uses;
//   shuffle  : the row's edges are spread over an 8-lane segment (degree 6-7 rows: 1-2 lanes idle), the (min1, min2, parity)
//              reduction is a 3-round xor-butterfly of warp shuffles, every lane then finishes the row constants itself —
//              the mapping BASELINE.json's north star names ("min1/min2/sign/index reduction done with warp shuffles").
//
// Workload: a synthetic 576x288-like code (6 levels of 48 rows, degree 7, all variables of a level distinct, so a level's rows are
// independent exactly as in the level schedule), P frame pairs per CTA resident in shared memory as binary16x2 (posterior
// U[n], messages MS[M]), one CTA per SM, a barrier between levels, ITERS iterations.  Both kernels read and write the same arrays;
// the host checks that they end bit-identical and prints edge updates per second.  Measurement tool, not product code:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o shfl_row shfl_row.cu && ./shfl_row
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
#include "../../ldpcgputegra_b200/csrc/rowops.cuh"

using namespace ldpcb200;

constexpr int N = 576, LEVELS = 6, ROWS_PER_LEVEL = 48, D = 7, ROWS = LEVELS * ROWS_PER_LEVEL, M = ROWS * D;
constexpr int P = 20;                                   // frame pairs per CTA: 20 * (576 + 2016) * 4 B = 207 KB of shared memory
constexpr int SEM = LDPC_SEM_X86_SSE, ALGO = LDPC_ALGO_OMS;

struct Args { const uint16_t* idx; uint32_t* state; ldpc_params_t prm; int iters; };   // state: [CTA][P][N + M] words, in and out

__device__ __forceinline__ h2 extrinsic(uint32_t u, uint32_t m, const RowConsts& K) { return __hmin2(__hsub2_sat(bits_h2(u), bits_h2(m)), K.top); }

// ---------------------------------------------------------------------------------------------------------------- in-lane
__global__ void __launch_bounds__(ROWS_PER_LEVEL * P) k_inlane(const __grid_constant__ Args A)
{
    extern __shared__ uint32_t sm[];
    uint16_t* s_idx = reinterpret_cast<uint16_t*>(sm + P * (N + M));
    uint32_t* g = A.state + (size_t)blockIdx.x * P * (N + M);
    for (int i = threadIdx.x; i < P * (N + M); i += blockDim.x) sm[i] = g[i];
    for (int i = threadIdx.x; i < M; i += blockDim.x) s_idx[i] = A.idx[i];
    __syncthreads();
    RowConsts K; make_consts<SEM>(K, A.prm);
    const int p = threadIdx.x / ROWS_PER_LEVEL, r = threadIdx.x % ROWS_PER_LEVEL;       // lane -> (pair, row of the level)
    uint32_t* U = sm + p * (N + M); uint32_t* MS = U + N;
    for (int it = 0; it < A.iters; it++)
        for (int L = 0; L < LEVELS; L++) {
            const int e0 = (L * ROWS_PER_LEVEL + r) * D;
            h2 xu[D], a[D]; uint32_t f[D]; int v[D];
#pragma unroll
            for (int j = 0; j < D; j++) { v[j] = s_idx[e0 + j]; xu[j] = extrinsic(U[v[j]], MS[e0 + j], K); }
            RowState s; row_pass1<SEM, ALGO, false, D>(xu, a, f, s, K);
            RowOut o; row_finish<SEM, ALGO>(s, D, K, K.msg_c, o);
            RowOutS q; fold_sign(o, q);
#pragma unroll
            for (int j = 0; j < D; j++) { h2 msg, un; pass2_edge_s(xu[j], a[j], f[j], q, K, msg, un); MS[e0 + j] = h2_bits(msg); U[v[j]] = h2_bits(un); }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < P * (N + M); i += blockDim.x) g[i] = sm[i];
}

// ---------------------------------------------------------------------------------------------------------------- shuffle
// 8-lane segment = one (pair, row) task; lane j < D holds edge j, lane 7 holds the neutral element.  1024 threads = 128 segments;
// a level has ROWS_PER_LEVEL * P = 960 tasks -> 7.5 rounds.
__global__ void __launch_bounds__(1024) k_shuffle(const __grid_constant__ Args A)
{
    extern __shared__ uint32_t sm[];
    uint16_t* s_idx = reinterpret_cast<uint16_t*>(sm + P * (N + M));
    uint32_t* g = A.state + (size_t)blockIdx.x * P * (N + M);
    for (int i = threadIdx.x; i < P * (N + M); i += blockDim.x) sm[i] = g[i];
    for (int i = threadIdx.x; i < M; i += blockDim.x) s_idx[i] = A.idx[i];
    __syncthreads();
    RowConsts K; make_consts<SEM>(K, A.prm);
    const int seg = threadIdx.x >> 3, j = threadIdx.x & 7, nseg = blockDim.x >> 3;
    const bool edge = j < D;
    for (int it = 0; it < A.iters; it++)
        for (int L = 0; L < LEVELS; L++) {
            for (int base = 0; base < ROWS_PER_LEVEL * P; base += nseg) {        // uniform trip count: full-mask shuffles
                const int task = base + seg;
                const bool live = task < ROWS_PER_LEVEL * P;
                const int p = live ? task / ROWS_PER_LEVEL : 0, r = live ? task % ROWS_PER_LEVEL : 0;
                uint32_t* U = sm + p * (N + M); uint32_t* MS = U + N;
                const int e = (L * ROWS_PER_LEVEL + r) * D + (edge ? j : 0);
                const int v = s_idx[e];
                const h2 xu = extrinsic(U[v], MS[e], K);
                const h2 t = signed_contrib(xu, K);
                const h2 a = magnitude<SEM, ALGO, false>(t, K);
                h2 min1 = edge ? a : K.min_init, min2 = K.min_init;
                uint32_t par = edge ? h2_bits(t) : 0u;
#pragma unroll
                for (int d = 1; d < 8; d <<= 1) {
                    const h2 o1 = bits_h2(__shfl_xor_sync(0xFFFFFFFFu, h2_bits(min1), d));
                    const h2 o2 = bits_h2(__shfl_xor_sync(0xFFFFFFFFu, h2_bits(min2), d));
                    par ^= __shfl_xor_sync(0xFFFFFFFFu, par, d);
                    const h2 n1 = __hmin2(min1, o1);
                    min2 = __hmin2(__hmax2(min1, o1), __hmin2(min2, o2));
                    min1 = n1;
                }
                RowState s; s.min1 = min1; s.min2 = min2; s.par = par;
                RowOut o; row_finish<SEM, ALGO>(s, D, K, K.msg_c, o);
                RowOutS q; fold_sign(o, q);
                h2 msg, un; pass2_edge_s(xu, a, h2_bits(t), q, K, msg, un);
                if (live && edge) { MS[e] = h2_bits(msg); U[v] = h2_bits(un); }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < P * (N + M); i += blockDim.x) g[i] = sm[i];
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main(int argc, char** argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 200;
    int dev = 0, sms = 0; CK(cudaSetDevice(dev)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    std::mt19937 rng(7);
    std::vector<uint16_t> idx(M);
    for (int L = 0; L < LEVELS; L++) {                       // a level's 48 x 7 variables are distinct: its rows are independent
        std::vector<int> perm(N); for (int i = 0; i < N; i++) perm[i] = i;
        std::shuffle(perm.begin(), perm.end(), rng);
        for (int k = 0; k < ROWS_PER_LEVEL * D; k++) idx[L * ROWS_PER_LEVEL * D + k] = (uint16_t)perm[k];
    }
    const size_t words = (size_t)sms * P * (N + M);
    std::vector<uint32_t> init(words);
    for (size_t c = 0; c < (size_t)sms * P; c++) {           // posteriors: biased LLRs in [-31, 31] -> (v + 127)/256 as binary16x2; messages 0
        uint32_t* U = init.data() + c * (N + M);
        for (int i = 0; i < N; i++) {
            const int v0 = (int)(rng() % 63) - 31, v1 = (int)(rng() % 63) - 31;
            const __half2 h = __floats2half2_rn((v0 + 127) / 256.0f, (v1 + 127) / 256.0f);
            memcpy(&U[i], &h, 4);
        }
        for (int i = 0; i < M; i++) U[N + i] = 0u;
    }
    uint16_t* d_idx; uint32_t *d_a, *d_b;
    CK(cudaMalloc(&d_idx, M * 2)); CK(cudaMemcpy(d_idx, idx.data(), M * 2, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_a, words * 4)); CK(cudaMalloc(&d_b, words * 4));
    Args A{}; A.idx = d_idx; A.iters = iters;
    A.prm.algo = ALGO; A.prm.semantics = SEM; A.prm.offset = 1; A.prm.factor_q5 = 29; A.prm.sat_var = 127; A.prm.sat_msg = 31;
    const size_t smem = (size_t)P * (N + M) * 4 + M * 2;
    CK(cudaFuncSetAttribute(k_inlane, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_shuffle, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms[2] = { 0, 0 };
    for (int which = 0; which < 2; which++) {
        uint32_t* d = which ? d_b : d_a; A.state = d;
        for (int rep = 0; rep < 3; rep++) {                  // the last repetition is the one timed and compared
            CK(cudaMemcpy(d, init.data(), words * 4, cudaMemcpyHostToDevice));
            CK(cudaEventRecord(e0));
            if (which) k_shuffle<<<sms, 1024, smem>>>(A); else k_inlane<<<sms, ROWS_PER_LEVEL * P, smem>>>(A);
            CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
            CK(cudaEventElapsedTime(&ms[which], e0, e1));
        }
    }
    std::vector<uint32_t> ra(words), rb(words);
    CK(cudaMemcpy(ra.data(), d_a, words * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(rb.data(), d_b, words * 4, cudaMemcpyDeviceToHost));
    size_t diff = 0; for (size_t i = 0; i < words; i++) diff += ra[i] != rb[i];
    const double edges = (double)sms * P * 2.0 * M * iters;                // frame-edge updates (two frames per word)
    printf("{\"sms\": %d, \"pairs_per_cta\": %d, \"iters\": %d, \"inlane_ms\": %.3f, \"shuffle_ms\": %.3f, \"inlane_Tedge_s\": %.3f, \"shuffle_Tedge_s\": %.3f, "
           "\"shuffle_over_inlane\": %.2f, \"words_differing\": %zu}\n",
           sms, P, iters, ms[0], ms[1], edges / ms[0] / 1e9, edges / ms[1] / 1e9, ms[1] / ms[0], diff);
    return diff ? 2 : 0;
}
