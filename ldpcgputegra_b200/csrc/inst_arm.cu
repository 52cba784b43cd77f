// inst_arm.cu — kernel instantiations for semantics mode LDPC_SEM_ARM_SCALAR (see launch.cuh)
#define LDPC_INST_SEM LDPC_SEM_ARM_SCALAR
#include "launch.cuh"

namespace ldpcb200 {

int launch_fp_arm(int algo, int et, const FpArgs& args, int blocks, cudaStream_t st)
{
    switch (algo) {
    case LDPC_ALGO_OMS: LDPC_CASE(do_fp, LDPC_SEM_ARM_SCALAR, LDPC_ALGO_OMS, args, blocks, st);
    }
    return (int)cudaErrorInvalidValue;
}

int launch_rp_arm(int algo, int et, const RpArgs& args, int blocks, int threads, size_t smem, cudaStream_t st)
{
    switch (algo) {
    case LDPC_ALGO_OMS: LDPC_CASE(do_rp, LDPC_SEM_ARM_SCALAR, LDPC_ALGO_OMS, args, blocks, threads, smem, st);
    }
    return (int)cudaErrorInvalidValue;
}


}  // namespace ldpcb200
