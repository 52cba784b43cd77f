// inst_fs_x86.cu — instantiations of the bulk-copy-staged frame-parallel kernel (kernel_fs.cuh) for semantics mode LDPC_SEM_X86_SSE: its own
// translation unit so that the staged kernel's 6 variants per (semantics, algorithm) compile beside the on-chip kernel's (see launch.cuh)
#define LDPC_INST_SEM LDPC_SEM_X86_SSE
#include "launch.cuh"

namespace ldpcb200 {

int launch_fs_x86_nms(const FsArgs& args, int blocks, size_t smem, cudaStream_t st);

int launch_fs_x86(int algo, const FsArgs& args, int blocks, size_t smem, cudaStream_t st)
{
    switch (algo) {
    case LDPC_ALGO_OMS: return do_fs<LDPC_SEM_X86_SSE, LDPC_ALGO_OMS>(args, blocks, smem, st);
    case LDPC_ALGO_NMS:
    case LDPC_ALGO_2NMS: return launch_fs_x86_nms(args, blocks, smem, st);        // inst_fs_x86_nms.cu (build parallelism)
    }
    return (int)cudaErrorInvalidValue;
}

}  // namespace ldpcb200
