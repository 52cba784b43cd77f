// inst_fc_arm.cu — instantiations of the bulk-copy-staged frame-parallel kernel on COMPRESSED messages (kernel_fs.cuh, CMP = true) for semantics mode LDPC_SEM_ARM_SCALAR: its own
// translation unit so that the staged kernel's 6 variants per (semantics, algorithm) compile beside the on-chip kernel's (see launch.cuh)
#define LDPC_INST_SEM LDPC_SEM_ARM_SCALAR
#include "launch.cuh"

namespace ldpcb200 {

int launch_fc_arm(int algo, const FsArgs& args, int blocks, size_t smem, cudaStream_t st)
{
    switch (algo) {
    case LDPC_ALGO_OMS: return do_fs<LDPC_SEM_ARM_SCALAR, LDPC_ALGO_OMS, true>(args, blocks, smem, st);
    }
    return (int)cudaErrorInvalidValue;
}

}  // namespace ldpcb200
