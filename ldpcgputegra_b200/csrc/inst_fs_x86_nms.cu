// inst_fs_x86_nms.cu — the NMS half of inst_fs_x86.cu (staged frame-parallel kernel, semantics mode LDPC_SEM_X86_SSE): its own translation
// unit so that the two halves compile side by side (the OMS half carries the quirk-row bodies and is the longest unit of the build)
#define LDPC_INST_SEM LDPC_SEM_X86_SSE
#include "launch.cuh"

namespace ldpcb200 {

int launch_fs_x86_nms(const FsArgs& args, int blocks, size_t smem, cudaStream_t st) { return do_fs<LDPC_SEM_X86_SSE, LDPC_ALGO_NMS>(args, blocks, smem, st); }

}  // namespace ldpcb200
