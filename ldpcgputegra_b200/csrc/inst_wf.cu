// inst_wf.cu — instantiations of the warp-per-frame on-chip engine (kernel_wf.cuh, kernel 6): storage type x schedule
#include <cuda_runtime.h>
#include "kernel_wf.cuh"

namespace ldpcb200 {

template <class S, bool FLOOD, bool PAIR>
static int do_wf(const WfArgs<S>& a, int blocks, int threads, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(wf_decode_kernel<S, FLOOD, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    wf_decode_kernel<S, FLOOD, PAIR><<<blocks, threads, smem, st>>>(a);
    return (int)cudaGetLastError();
}

int launch_wf(const WfArgs<float>& a, int blocks, int threads, size_t smem, cudaStream_t st)
{
    if (a.pair) return a.flooding ? do_wf<float, true, true>(a, blocks, threads, smem, st) : do_wf<float, false, true>(a, blocks, threads, smem, st);
    return a.flooding ? do_wf<float, true, false>(a, blocks, threads, smem, st) : do_wf<float, false, false>(a, blocks, threads, smem, st);
}
int launch_wf(const WfArgs<int16_t>& a, int blocks, int threads, size_t smem, cudaStream_t st)
{
    if (a.pair) return a.flooding ? do_wf<int16_t, true, true>(a, blocks, threads, smem, st) : do_wf<int16_t, false, true>(a, blocks, threads, smem, st);
    return a.flooding ? do_wf<int16_t, true, false>(a, blocks, threads, smem, st) : do_wf<int16_t, false, false>(a, blocks, threads, smem, st);
}
int launch_wf(const WfArgs<int8_t>& a, int blocks, int threads, size_t smem, cudaStream_t st)
{
    if (a.pair) return a.flooding ? do_wf<int8_t, true, true>(a, blocks, threads, smem, st) : do_wf<int8_t, false, true>(a, blocks, threads, smem, st);
    return a.flooding ? do_wf<int8_t, true, false>(a, blocks, threads, smem, st) : do_wf<int8_t, false, false>(a, blocks, threads, smem, st);
}

}  // namespace ldpcb200
