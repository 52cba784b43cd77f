// launch.cuh — per-semantics kernel instantiation units.  Each inst_<sem>.cu includes this with LDPC_INST_SEM defined, so
// the template instantiations compile in parallel and a semantics mode costs nothing in the others' kernels.
#pragma once
#include <cuda_runtime.h>
#include "kernel_fp.cuh"
#include "kernel_rp.cuh"
#include "kernel_fs.cuh"

namespace ldpcb200 {

// algo: ldpc_algo_t (2NMS shares the NMS instantiation: only the rescale constants differ). Returns cudaError_t as int.
typedef int (*fp_launch_fn)(int algo, int et, const FpArgs& args, int blocks, cudaStream_t st);
typedef int (*fs_launch_fn)(int algo, const FsArgs& args, int blocks, size_t smem, cudaStream_t st);
typedef int (*rp_launch_fn)(int algo, int et, const RpArgs& args, int blocks, int threads, size_t smem, cudaStream_t st);

int launch_fp_x86(int, int, const FpArgs&, int, cudaStream_t);
int launch_fp_uniform(int, int, const FpArgs&, int, cudaStream_t);
int launch_fp_arm(int, int, const FpArgs&, int, cudaStream_t);
int launch_fp_gpu(int, int, const FpArgs&, int, cudaStream_t);
int launch_fs_x86(int, const FsArgs&, int, size_t, cudaStream_t);
int launch_fs_uniform(int, const FsArgs&, int, size_t, cudaStream_t);
int launch_fs_arm(int, const FsArgs&, int, size_t, cudaStream_t);
int launch_fs_gpu(int, const FsArgs&, int, size_t, cudaStream_t);
int launch_fc_x86(int, const FsArgs&, int, size_t, cudaStream_t);        // the same kernel on compressed messages (inst_fc_*.cu)
int launch_fc_uniform(int, const FsArgs&, int, size_t, cudaStream_t);
int launch_fc_arm(int, const FsArgs&, int, size_t, cudaStream_t);
int launch_fc_gpu(int, const FsArgs&, int, size_t, cudaStream_t);
int launch_rp_x86(int, int, const RpArgs&, int, int, size_t, cudaStream_t);
int launch_rp_uniform(int, int, const RpArgs&, int, int, size_t, cudaStream_t);
int launch_rp_arm(int, int, const RpArgs&, int, int, size_t, cudaStream_t);
int launch_rp_gpu(int, int, const RpArgs&, int, int, size_t, cudaStream_t);

#ifdef LDPC_INST_SEM
template <int SEM, int ALGO, bool ET>
static int do_fp(const FpArgs& a, int blocks, cudaStream_t st)
{
    fp_decode_kernel<SEM, ALGO, ET><<<blocks, FP_BLOCK, 0, st>>>(a);
    return (int)cudaGetLastError();
}
template <int SEM, int ALGO, bool ET>
static int do_rp(const RpArgs& a, int blocks, int threads, size_t smem, cudaStream_t st)
{
    // per-device attribute, cheap and idempotent: set on every launch so any device of the process is covered
    if (a.static_nrows > 0) {
        cudaError_t e = cudaFuncSetAttribute(rp_decode_kernel<SEM, ALGO, ET, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        rp_decode_kernel<SEM, ALGO, ET, true><<<blocks, threads, smem, st>>>(a);
    } else {
        cudaError_t e = cudaFuncSetAttribute(rp_decode_kernel<SEM, ALGO, ET, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        rp_decode_kernel<SEM, ALGO, ET, false><<<blocks, threads, smem, st>>>(a);
    }
    return (int)cudaGetLastError();
}
template <int SEM, int ALGO, bool CMP = false>
static int do_fs(const FsArgs& a, int blocks, size_t smem, cudaStream_t st)
{
#define FS_LAUNCH(NCV, MD, ETV)                                                                                                              \
    { cudaError_t e = cudaFuncSetAttribute(fs_decode_kernel<SEM, ALGO, NCV, MD, ETV, CMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return (int)e;                                                                                                      \
      fs_decode_kernel<SEM, ALGO, NCV, MD, ETV, CMP><<<blocks, NCV + FS_PRODUCER_THREADS, smem, st>>>(a); }
    if (CMP && a.max_deg > 8) return (int)cudaErrorInvalidValue;        // compressed rows carry 8 edge bits per frame
    // one CTA per SM and staircase runs in the code: the paired-row instantiation (kernel_fs.cuh: fs_row_stair2), no register cap
    if (!CMP && a.pipe2 && a.nc == 128 && a.max_deg <= 8) {
#define FS_LAUNCH2(ETV)                                                                                                                       \
    { cudaError_t e = cudaFuncSetAttribute(fs_decode_kernel<SEM, ALGO, 128, 8, ETV, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return (int)e;                                                                                                      \
      fs_decode_kernel<SEM, ALGO, 128, 8, ETV, false, true><<<blocks, 128 + FS_PRODUCER_THREADS, smem, st>>>(a); }
        if constexpr (!CMP) { if (a.et) FS_LAUNCH2(true) else FS_LAUNCH2(false) }
#undef FS_LAUNCH2
        return (int)cudaGetLastError();
    }
    // CTA width x widest row body x early termination.  The kernel's registers are those of its widest body, and the two producer
    // warps get the same allocation as the consumers: 256 consumers per CTA keep 16 consumer warps on an SM where 128 keep 12
    // (DESIGN.md 3.2b).  320 consumers capped at 80 registers (20 consumer warps) were measured 3 % SLOWER at a balanced batch and
    // the degree-10 body needs 120 registers (one wide CTA per SM, never a win): neither is instantiated.
    if (a.et) {
        if (a.nc == 256 && a.max_deg <= 8) FS_LAUNCH(256, 8, true)
        else { if (a.max_deg <= 8) FS_LAUNCH(128, 8, true) else if constexpr (!CMP) { if (a.max_deg <= FS_MAXDEG) FS_LAUNCH(128, FS_MAXDEG, true) else FS_LAUNCH(128, FS_GEN_MAXDEG, true) } }
    } else {
        if (a.nc == 256 && a.max_deg <= 8) FS_LAUNCH(256, 8, false)
        else { if (a.max_deg <= 8) FS_LAUNCH(128, 8, false) else if constexpr (!CMP) { if (a.max_deg <= FS_MAXDEG) FS_LAUNCH(128, FS_MAXDEG, false) else FS_LAUNCH(128, FS_GEN_MAXDEG, false) } }
    }
#undef FS_LAUNCH
    return (int)cudaGetLastError();
}
#define LDPC_CASE(FN, SEM, ALGO, ...) return et ? FN<SEM, ALGO, true>(__VA_ARGS__) : FN<SEM, ALGO, false>(__VA_ARGS__)
#endif

}  // namespace ldpcb200
