// diverged_wait.cu — what a warp pays for its collectives after ONE lane has waited on an mbarrier by itself.
//
// The staged kernel's producer warps (kernel_fs.cuh) used to do `if (lane == 0) mbar_wait(slot free)` at the top of every row and
// then a handful of __shfl_sync / __ballot_sync / __syncwarp.  profiles/r02_ncu_fs_small_v4.txt shows every one of those collectives
// on the compiler's DIVERGENT path (BRA.DIV taken, WARPSYNC.COLLECTIVE ... ENDCOLLECTIVE), which is what held every small DVB-S2 batch
// at 578 ns per row.  This is the same loop in isolation: one warp, an mbarrier whose phase 0 is already complete (the wait returns at
// once), NSHFL shuffles and one __syncwarp per iteration, timed with clock64 —
//   variant 0: lane 0 alone executes the wait loop (the old producer);
//   variant 1: every lane executes it (one warp instruction either way; the producer since round 2);
//   variant 2: no wait at all (the floor).
// Measurement tool, not product code:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o diverged_wait diverged_wait.cu && ./diverged_wait
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)     // the product's wait loop, verbatim
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}

constexpr int NSHFL = 6;

template <int VARIANT>
__global__ void k(int iters, uint32_t* out, long long* cycles)
{
    __shared__ __align__(8) unsigned long long bar_mem;
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bar_mem);
    const int lane = threadIdx.x & 31;
    if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);            // phase 0 completes: a wait on parity 0 passes at once from here on
    __syncwarp();
    uint32_t v = (uint32_t)lane * 2654435761u;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (VARIANT == 0) { if (lane == 0) mbar_wait(bar, 0u); }
        else if (VARIANT == 1) mbar_wait(bar, 0u);
        __syncwarp();
#pragma unroll
        for (int s = 0; s < NSHFL; s++) v = v * 3u + __shfl_sync(0xFFFFFFFFu, v, (lane + s + 1) & 31);
    }
    const long long t1 = clock64();
    out[threadIdx.x] = v;
    if (lane == 0) *cycles = t1 - t0;
}

int main()
{
    const int iters = 200000;
    uint32_t* d_out; long long* d_cyc;
    cudaMalloc(&d_out, 32 * sizeof(uint32_t)); cudaMalloc(&d_cyc, sizeof(long long));
    long long cyc[3]; uint32_t chk[3];
    for (int rep = 0; rep < 2; rep++) {         // the second round is the measurement
        k<0><<<1, 32>>>(iters, d_out, d_cyc); cudaMemcpy(&cyc[0], d_cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&chk[0], d_out, 4, cudaMemcpyDeviceToHost);
        k<1><<<1, 32>>>(iters, d_out, d_cyc); cudaMemcpy(&cyc[1], d_cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&chk[1], d_out, 4, cudaMemcpyDeviceToHost);
        k<2><<<1, 32>>>(iters, d_out, d_cyc); cudaMemcpy(&cyc[2], d_cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&chk[2], d_out, 4, cudaMemcpyDeviceToHost);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    printf("{\"iters\": %d, \"collectives_per_iter\": %d, \"cycles_per_iter\": {\"lane0_waits_alone\": %.1f, \"every_lane_waits\": %.1f, \"no_wait\": %.1f}, "
           "\"same_result\": %s}\n", iters, NSHFL + 1, (double)cyc[0] / iters, (double)cyc[1] / iters, (double)cyc[2] / iters,
           (chk[0] == chk[1] && chk[1] == chk[2]) ? "true" : "false");
    return 0;
}
