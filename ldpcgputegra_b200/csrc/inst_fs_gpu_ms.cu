// inst_fs_gpu_ms.cu — the MS part of inst_fs_gpu.cu (staged frame-parallel kernel, semantics mode LDPC_SEM_GPU_FIXED): its own
// translation unit so that the three algorithms compile side by side
#define LDPC_INST_SEM LDPC_SEM_GPU_FIXED
#include "launch.cuh"

namespace ldpcb200 {

int launch_fs_gpu_ms(const FsArgs& args, int blocks, size_t smem, cudaStream_t st) { return do_fs<LDPC_SEM_GPU_FIXED, LDPC_ALGO_MS>(args, blocks, smem, st); }

}  // namespace ldpcb200
