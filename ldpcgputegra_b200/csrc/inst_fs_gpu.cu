// inst_fs_gpu.cu — instantiations of the bulk-copy-staged frame-parallel kernel (kernel_fs.cuh) for semantics mode LDPC_SEM_GPU_FIXED: its own
// translation unit so that the staged kernel's 6 variants per (semantics, algorithm) compile beside the on-chip kernel's (see launch.cuh)
#define LDPC_INST_SEM LDPC_SEM_GPU_FIXED
#include "launch.cuh"

namespace ldpcb200 {

int launch_fs_gpu_ms(const FsArgs& args, int blocks, size_t smem, cudaStream_t st);
int launch_fs_gpu_nms(const FsArgs& args, int blocks, size_t smem, cudaStream_t st);

int launch_fs_gpu(int algo, const FsArgs& args, int blocks, size_t smem, cudaStream_t st)
{
    switch (algo) {
    case LDPC_ALGO_MS: return launch_fs_gpu_ms(args, blocks, smem, st);            // inst_fs_gpu_ms.cu (build parallelism)
    case LDPC_ALGO_OMS: return do_fs<LDPC_SEM_GPU_FIXED, LDPC_ALGO_OMS>(args, blocks, smem, st);
    case LDPC_ALGO_NMS:
    case LDPC_ALGO_2NMS: return launch_fs_gpu_nms(args, blocks, smem, st);         // inst_fs_gpu_nms.cu
    }
    return (int)cudaErrorInvalidValue;
}

}  // namespace ldpcb200
