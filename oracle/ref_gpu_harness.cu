/*
 * ref_gpu_harness.cu — C-callable shell around the REFERENCE's own gpu_fixed kernels, built for sm_100a.
 * TEST INFRASTRUCTURE ONLY; contains none of the reference's code.  oracle/Makefile compiles it together with the reference
 * kernel sources where they lie under /root/reference (decoder_{ms,oms,nms,2nms}/cuda/CUDA_x_SIMD.cu, transpose/
 * GPU_Transpose_uint8.cu, matrix/constantes_decoder.cpp) into oracle/_ref/libref_gpu_<code>.so.  The code is selected without
 * touching the read-only tree: -DCODE=<n> -DCONSTANTES_MANAGEMENT skips the reference's dispatcher header and
 * -include <code>/constantes_gpu.h supplies _N/_K/_M/DEG_x.
 *
 * It runs exactly the launch sequence of CGPU_Decoder_OMS_SIMD::decode (ref: code/gpu_fixed/decoder_oms/CGPU_Decoder_OMS_SIMD.cu:97-149)
 * — H2D into the message buffer, Interleaver_uint8, the decode kernel <<<T/128,128>>>, InvInterleaver_uint8, D2H — and
 * additionally copies out the interleaved posteriors/messages before the de-interleave overwrites the message buffer.
 * It pins the GPU_FIXED mode of the oracle on the B200 box (tests/test_parity_gpu.py::test_reference_gpu_kernels) and is
 * "the kernel to beat" in bench.py --ref-gpu.
 */
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

extern __global__ void LDPC_Sched_Stage_1_MS_SIMD(unsigned int*, unsigned int*, unsigned int*, unsigned int);
extern __global__ void LDPC_Sched_Stage_1_OMS_SIMD(unsigned int*, unsigned int*, unsigned int*, unsigned int);
extern __global__ void LDPC_Sched_Stage_1_NMS_SIMD(unsigned int*, unsigned int*, unsigned int*, unsigned int);
extern __global__ void LDPC_Sched_Stage_1_2NMS_SIMD(unsigned int*, unsigned int*, unsigned int*, unsigned int);
extern __global__ void Interleaver_uint8(int* in, int* out, int taille_frame, int nb_frames);
extern __global__ void InvInterleaver_uint8(int* in, int* out, int taille_frame, int nb_frames);
extern const unsigned int PosNoeudsVariable[_M];

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "ref_gpu: %s -> %s\n", #x, cudaGetErrorString(e)); return -2; } } while (0)

extern "C" {

void ref_gpu_info(int* out4) { out4[0] = _N; out4[1] = _K; out4[2] = _M; out4[3] = NB_DEGRES; }
void ref_gpu_table(uint32_t* pos) { for (int i = 0; i < _M; i++) pos[i] = PosNoeudsVariable[i]; }

/* frames % 512 == 0 (the reference needs nb_frames(threads) % 128 == 0, 4 frames per thread). post/msgs nullable.
 * ms_out (nullable) receives the kernel time in ms (cudaEvent around the decode kernel only) and the whole
 * H2D..D2H time in ms_out[1]. */
int ref_gpu_decode(int algo, const int8_t* llr, uint8_t* hard, int8_t* post, int8_t* msgs, size_t frames, int iters, float* ms_out)
{
    if (frames % 512 || iters < 1) return -1;
    const size_t T = frames / 4;
    unsigned int *d_pos = nullptr, *d_msg = nullptr, *d_v = nullptr;
    CK(cudaMalloc((void**)&d_pos, sizeof(unsigned int) * _M));
    CK(cudaMemcpy(d_pos, PosNoeudsVariable, sizeof(unsigned int) * _M, cudaMemcpyHostToDevice));
    CK(cudaMalloc((void**)&d_msg, sizeof(unsigned int) * (size_t)_M * T));      /* d_MSG_C_2_V (ref: CGPUDecoder.cpp:36) */
    /* device_V (ref: CGPUDecoder.cpp:37), padded by 128 rows: InvInterleaver_uint8 reads whole 128-row chunks with its bounds check
     * commented out (GPU_Transpose_uint8.cu:18,27-35) and discards what lies past _N only when writing (:48), so for _N % 128 != 0
     * (576, 1200) it reads up to 64 rows beyond the array — harmless next to other allocations at the reference's small batch sizes,
     * an illegal address at 64 Ki frames (seen here). */
    CK(cudaMalloc((void**)&d_v, sizeof(unsigned int) * (size_t)(_N + 128) * T));
    cudaEvent_t e0, e1, e2, e3;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    cudaEventRecord(e0);
    CK(cudaMemcpy(d_msg, llr, (size_t)_N * T * 4, cudaMemcpyHostToDevice));
    dim3 grid(1, (unsigned)(T / 32)), threads(32, 32);
    Interleaver_uint8<<<grid, threads>>>((int*)d_msg, (int*)d_v, _N, (int)T);
    cudaEventRecord(e1);
    switch (algo) {
    case 0: LDPC_Sched_Stage_1_MS_SIMD<<<(unsigned)(T / 128), 128>>>(d_v, d_msg, d_pos, iters); break;
    case 1: LDPC_Sched_Stage_1_OMS_SIMD<<<(unsigned)(T / 128), 128>>>(d_v, d_msg, d_pos, iters); break;
    case 2: LDPC_Sched_Stage_1_NMS_SIMD<<<(unsigned)(T / 128), 128>>>(d_v, d_msg, d_pos, iters); break;
    case 3: LDPC_Sched_Stage_1_2NMS_SIMD<<<(unsigned)(T / 128), 128>>>(d_v, d_msg, d_pos, iters); break;
    default: return -1;
    }
    cudaEventRecord(e2);
    CK(cudaGetLastError());
    if (post || msgs) {
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> w((size_t)(post && msgs ? (_M > _N ? _M : _N) : (msgs ? _M : _N)) * T);
        if (post) {
            CK(cudaMemcpy(w.data(), d_v, (size_t)_N * T * 4, cudaMemcpyDeviceToHost));
            for (size_t f = 0; f < frames; f++) for (int n = 0; n < _N; n++) post[f * _N + n] = (int8_t)(w[(size_t)n * T + f / 4] >> (8 * (f % 4)));
        }
        if (msgs) {
            CK(cudaMemcpy(w.data(), d_msg, (size_t)_M * T * 4, cudaMemcpyDeviceToHost));
            for (size_t f = 0; f < frames; f++) for (int e = 0; e < _M; e++) msgs[f * (size_t)_M + e] = (int8_t)(w[(size_t)e * T + f / 4] >> (8 * (f % 4)));
        }
    }
    InvInterleaver_uint8<<<grid, threads>>>((int*)d_v, (int*)d_msg, _N, (int)T);
    CK(cudaMemcpy(hard, d_msg, (size_t)_N * T * 4, cudaMemcpyDeviceToHost));
    cudaEventRecord(e3);
    CK(cudaDeviceSynchronize());
    if (ms_out) { cudaEventElapsedTime(&ms_out[0], e1, e2); cudaEventElapsedTime(&ms_out[1], e0, e3); }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
    cudaFree(d_pos); cudaFree(d_msg); cudaFree(d_v);
    return 0;
}

}  /* extern "C" */
