// inst_fc_uniform.cu — instantiations of the bulk-copy-staged frame-parallel kernel on COMPRESSED messages (kernel_fs.cuh, CMP = true) for semantics mode LDPC_SEM_UNIFORM: its own
// translation unit so that the staged kernel's 6 variants per (semantics, algorithm) compile beside the on-chip kernel's (see launch.cuh)
#define LDPC_INST_SEM LDPC_SEM_UNIFORM
#include "launch.cuh"

namespace ldpcb200 {

int launch_fc_uniform(int algo, const FsArgs& args, int blocks, size_t smem, cudaStream_t st)
{
    switch (algo) {
    case LDPC_ALGO_OMS: return do_fs<LDPC_SEM_UNIFORM, LDPC_ALGO_OMS, true>(args, blocks, smem, st);
    case LDPC_ALGO_NMS:
    case LDPC_ALGO_2NMS: return do_fs<LDPC_SEM_UNIFORM, LDPC_ALGO_NMS, true>(args, blocks, smem, st);
    }
    return (int)cudaErrorInvalidValue;
}

}  // namespace ldpcb200
