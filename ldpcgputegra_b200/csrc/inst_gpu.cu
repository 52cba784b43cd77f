// inst_gpu.cu — kernel instantiations for semantics mode LDPC_SEM_GPU_FIXED (see launch.cuh)
#define LDPC_INST_SEM LDPC_SEM_GPU_FIXED
#include "launch.cuh"

namespace ldpcb200 {

int launch_fp_gpu(int algo, int et, const FpArgs& args, int blocks, cudaStream_t st)
{
    switch (algo) {
    case LDPC_ALGO_MS: LDPC_CASE(do_fp, LDPC_SEM_GPU_FIXED, LDPC_ALGO_MS, args, blocks, st);
    case LDPC_ALGO_OMS: LDPC_CASE(do_fp, LDPC_SEM_GPU_FIXED, LDPC_ALGO_OMS, args, blocks, st);
    case LDPC_ALGO_NMS:
    case LDPC_ALGO_2NMS: LDPC_CASE(do_fp, LDPC_SEM_GPU_FIXED, LDPC_ALGO_NMS, args, blocks, st);
    }
    return (int)cudaErrorInvalidValue;
}

int launch_rp_gpu(int algo, int et, const RpArgs& args, int blocks, int threads, size_t smem, cudaStream_t st)
{
    switch (algo) {
    case LDPC_ALGO_MS: LDPC_CASE(do_rp, LDPC_SEM_GPU_FIXED, LDPC_ALGO_MS, args, blocks, threads, smem, st);
    case LDPC_ALGO_OMS: LDPC_CASE(do_rp, LDPC_SEM_GPU_FIXED, LDPC_ALGO_OMS, args, blocks, threads, smem, st);
    case LDPC_ALGO_NMS:
    case LDPC_ALGO_2NMS: LDPC_CASE(do_rp, LDPC_SEM_GPU_FIXED, LDPC_ALGO_NMS, args, blocks, threads, smem, st);
    }
    return (int)cudaErrorInvalidValue;
}


}  // namespace ldpcb200
