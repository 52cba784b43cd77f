/*
 * ldpc_b200.h — C ABI of the B200-native LDPC decoder (drop-in boundary).
 *
 * Every entry point replaces one piece of the reference's decoder boundary;
 * the reference interface it stands in for is cited as (ref: path:line), paths
 * relative to the reference tree (boiseHPSim/ldpcGpuTegra).
 *
 * Conventions (differences from the reference are deliberate and listed):
 *   - plain C, pointers + sizes, no C++/torch types; int return = 0 on success,
 *     negative ldpc_status_t on failure.  The reference prints and exit(0)s
 *     (ref: code/gpu_fixed/custom_api/custom_cuda.cu:5-17); this library never
 *     exits — the C++ adapters in ldpcgputegra_b200/adapters restore that behaviour.
 *   - a handle is single-threaded: one per (GPU, host thread), like one
 *     CGPUDecoder object per OpenMP section (ref: code/gpu_fixed/test.cpp:241-281).
 *   - I/O buffers are caller-owned.  LLR input is frame-major int8 [F][N]
 *     (ref: code/x86/CDecoder/OMS/CDecoder_OMS_fixed_SSE.cpp:140-149,
 *      code/gpu_fixed/transpose/GPU_Transpose_uint8.cu:80-130); hard-decision
 *     output is frame-major, one byte per bit in {0,1} (ref: code/x86/CTools/CTools.cpp:370,
 *     code/gpu_fixed/transpose/GPU_Transpose_uint8.cu:29) or bit-packed (new; LSB-first,
 *     bit n of frame f is (out[f*ceil(N/8) + n/8] >> (n%8)) & 1).
 *   - there is NO CPU fallback: every compute entry point needs a CUDA device.
 */
#ifndef LDPC_B200_H
#define LDPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDPC_B200_ABI_VERSION 1
#define LDPC_MAX_DEG_CLASSES 8

typedef enum {
    LDPC_OK = 0,
    LDPC_ERR_INVALID = -1,     /* bad argument / unsupported combination          */
    LDPC_ERR_CUDA = -2,        /* CUDA runtime failure (see ldpc_b200_last_error) */
    LDPC_ERR_NO_DEVICE = -3,   /* no usable GPU: the library has no CPU path      */
    LDPC_ERR_IO = -4,          /* code-table file unreadable / malformed          */
    LDPC_ERR_NOMEM = -5,
    LDPC_ERR_UNSUPPORTED = -6
} ldpc_status_t;

/* check-node update rule (ref: decoder_{ms,oms,nms,2nms}/cuda/CUDA_x_SIMD.cu:25, x86 CDecoder_{OMS,NMS}_fixed_SSE) */
typedef enum { LDPC_ALGO_MS = 0, LDPC_ALGO_OMS = 1, LDPC_ALGO_NMS = 2, LDPC_ALGO_2NMS = 3 } ldpc_algo_t;

/* message-passing schedule.  Every reference decoder is horizontal-layered; flooding is new. */
typedef enum { LDPC_SCHED_LAYERED = 0, LDPC_SCHED_FLOODING = 1 } ldpc_schedule_t;

/* arithmetic type of posteriors / messages */
typedef enum { LDPC_DTYPE_I8 = 0, LDPC_DTYPE_I16 = 1, LDPC_DTYPE_F32 = 2 } ldpc_dtype_t;

/*
 * Which of the reference's (mutually non-identical) fixed-point semantics to reproduce bit-exactly.
 *   X86_SSE    : the binary the reference builds (ref: code/x86/CDecoder/OMS/CDecoder_OMS_fixed_SSE.cpp:122-574,
 *                NMS/CDecoder_NMS_fixed_SSE.cpp:125-368): -127 rail, zero counts positive, and for OMS rows of degree
 *                class >= 1 the magnitude is |min(x, sat_msg)| (negative side unclamped).
 *   UNIFORM    : the AVX2 twin (ref: code/x86/CDecoder/OMS/CDecoder_OMS_fixed_AVX.cpp:240-346): min(|x|, sat_msg) everywhere.
 *   ARM_SCALAR : the scalar decoder of the ARM tree (ref: code/ldpc_decoder_arm/CDecoder/OMS/CDecoder_OMS_fixed_x86.cpp:61-200):
 *                run-time rails, zero counts negative, the only reference decoder with a syndrome stop criterion.
 *   GPU_FIXED  : the gpu_fixed kernels (ref: code/gpu_fixed/decoder_oms/cuda/CUDA_OMS_SIMD.cu:25-262 and siblings):
 *                -128 rail, unclamped |x| (up to 128), zero counts negative, literal offset 1 / clamp 31 / 0.75 / 0.875.
 */
typedef enum { LDPC_SEM_X86_SSE = 0, LDPC_SEM_UNIFORM = 1, LDPC_SEM_ARM_SCALAR = 2, LDPC_SEM_GPU_FIXED = 3 } ldpc_semantics_t;

typedef enum { LDPC_ET_NONE = 0,
               LDPC_ET_SYNDROME = 1   /* per frame, after each full iteration: stop when every check's extrinsic-sign parity
                                         is satisfied (ref: ldpc_decoder_arm/.../CDecoder_OMS_fixed_x86.cpp:150-192) */
} ldpc_early_term_t;

typedef enum { LDPC_OUT_BYTES = 0, LDPC_OUT_PACKED = 1 } ldpc_out_format_t;

/*
 * Code table = the reference's compile-time header, as data.
 * (ref: code/x86/Constantes/576x288/constantes_sse.h:26-60, code/gpu_fixed/matrix/576x288/constantes_gpu.h:6-39 +
 *  constantes_decoder.h:3).  n = _N, n_checks = _K (number of CHECK nodes, not info bits), m = _M,
 * deg[i]/rows[i] = DEG_(i+1) / DEG_(i+1)_COMPUTATIONS, pos = PosNoeudsVariable[_M] in reference row order.
 */
typedef struct {
    int32_t n;
    int32_t n_checks;
    int32_t m;
    int32_t nb_deg;
    int32_t deg[LDPC_MAX_DEG_CLASSES];
    int32_t rows[LDPC_MAX_DEG_CLASSES];
    uint32_t* pos;        /* [m]; owned by whoever filled the struct (ldpc_b200_free_code for loader-filled ones) */
} ldpc_code_t;

/*
 * Decoder parameters = the reference's setters + the literals its GPU kernels hard-code.
 * (ref: setOffset CDecoder_OMS_fixed_SSE.cpp:101-109, setFactor CDecoder_NMS_fixed_SSE.cpp:107-111,
 *  setVarRange/setMsgRange CDecoder_fixed.cpp:32-43, defaults code/x86/main_p.cpp:90-104,133-139)
 */
typedef struct {
    int32_t algo;         /* ldpc_algo_t                                   */
    int32_t schedule;     /* ldpc_schedule_t                               */
    int32_t dtype;        /* ldpc_dtype_t                                  */
    int32_t semantics;    /* ldpc_semantics_t                              */
    int32_t offset;       /* OMS offset, default 1                         */
    int32_t factor_q5;    /* x86 NMS: (min*factor)>>5, default 29          */
    float   factor1;      /* GPU NMS/2NMS + float NMS: min1 scale, 0.75    */
    float   factor2;      /* GPU 2NMS: min2 scale, 0.875 (NMS: = factor1)  */
    int32_t sat_var;      /* posterior rail, 127 (int8) / up to 32767 (i16)*/
    int32_t sat_msg;      /* message clamp, 31                             */
    int32_t llr_scale;    /* FACTEUR_BETA = 8                              */
    int32_t sat_llr;      /* quantiser clamp, 31                           */
    int32_t early_term;   /* ldpc_early_term_t                             */
    int32_t out_format;   /* ldpc_out_format_t                             */
    int32_t kernel;       /* 0 = auto; 1 = frame-parallel (HBM-resident state); 2 = row-parallel on-chip (short codes);
                             3 = generic engine (fp32 arithmetic: int16 / float / flooding, and int8 layered as a cross-check);
                             4 = frame-parallel with the state staged through shared memory by cp.async.bulk / tensor-map copies (long codes; row degrees 3..32);
                             5 = generic engine with the state on chip, a CTA owns F frames (int16 / float / flooding on short codes);
                             6 = generic engine with the state on chip, ONE WARP PER FRAME and a global work queue (the same modes; chosen
                                 when >= 8 frames fit per SM and the schedule fills the lanes: flooding, or layered with wide levels) */
    int32_t reserved[5];  /* 0 for production use.  Experiment knobs of this implementation, used by tools/ and the A/B tests only:
                             [0],[1] = (warps, pairs) of an on-chip group; [2] = kernel waves per pipeline chunk of decode();
                             [3] = 1: descriptor-driven on-chip plan, 3: 32-row steps, 5: pair-slowest lane mapping;
                             [4] = stage-ring depth of the staged kernel (bits 0..7) | its CTA width (bits 8..11: 1 = 128, 2 = 256 consumer threads)
                                   | message lines through a TMA tensor map (bits 12..13: 1 = never, 2 = always)
                                   | posterior lines through tile::gather4 (bits 14..15, likewise)
                                   | compressed messages, four words per row instead of one per edge (bits 16..17: 2 = on, rows of degree <= 8;
                                     bit-exact, 2/3 of the state, measured slightly SLOWER: off unless asked for — DESIGN.md 3.2b)
                                   | bit 18: no register-carried staircase runs | bits 19..20: paired staircase rows (1 = never, 2 = always) */
} ldpc_params_t;

typedef struct ldpc_b200_handle_s* ldpc_handle;

/* library-level ------------------------------------------------------------------------------------------------------ */
int         ldpc_b200_abi_version(void);
int         ldpc_b200_device_count(void);                       /* 0 when no CUDA device is visible (never negative)   */
const char* ldpc_b200_status_string(int status);
void        ldpc_b200_default_params(ldpc_params_t* p);         /* OMS, layered, int8, X86_SSE, offset 1, 127/31, bytes  */

/* H-matrix load (ref: the `#include "./576x288/constantes_sse.h"` of code/x86/Constantes/constantes_sse.h:1 and the CODE
 * switch of code/gpu_fixed/matrix/constantes_gpu.h:18-76).  Host-only, no GPU needed. */
int  ldpc_b200_load_code_header(ldpc_code_t* out, const char* header_path, const char* table_path /* nullable: GPU flavour
                                keeps the index array in a second file, constantes_decoder.h */);
int  ldpc_b200_load_code_table(ldpc_code_t* out, const char* path);      /* this library's own compact binary table */
int  ldpc_b200_save_code_table(const ldpc_code_t* code, const char* path);
int  ldpc_b200_check_code(const ldpc_code_t* code);                      /* structural validation                    */
void ldpc_b200_free_code(ldpc_code_t* code);
/* depth of the dependency DAG of the layered schedule in reference row order; fills level_of_row[n_checks]
 * when non-null.  Returns the number of levels or a negative status. */
int  ldpc_b200_level_schedule(const ldpc_code_t* code, int32_t* level_of_row);

/* decoder object (ref: CGPUDecoder::CGPUDecoder(nb_frames,n,k,m) code/gpu_fixed/decoder_template/CGPUDecoder.cpp:14-38;
 * CreateDecoder(...) code/x86/CDecoder/DecoderLibrary.h:44-134).  max_frames is a capacity in FRAMES (the reference's
 * nb_frames counts GPU threads = 4 frames each). */
int  ldpc_b200_create(ldpc_handle* h, const ldpc_code_t* code, const ldpc_params_t* params, int device, size_t max_frames);
void ldpc_b200_destroy(ldpc_handle h);                          /* ref: CGPUDecoder::~CGPUDecoder CGPUDecoder.cpp:41-54 */
const char* ldpc_b200_last_error(ldpc_handle h);                /* nullable handle: last create() error of this thread  */
int  ldpc_b200_get_info(ldpc_handle h, int what, int64_t* value);
enum { LDPC_INFO_KERNEL = 0,            /* which decode kernel the handle selected (1 .. 6)                        */
       LDPC_INFO_LEVELS = 1,            /* level-schedule depth                                                    */
       LDPC_INFO_SMEM_BYTES = 2,        /* dynamic shared memory per CTA                                            */
       LDPC_INFO_FRAMES_PER_CTA = 3,
       LDPC_INFO_LAUNCHES = 4,          /* kernels launched by this handle so far (for bench.py's gpu_launches)     */
       LDPC_INFO_STREAM_SLOTS = 5,
       LDPC_INFO_DEVICE = 6,
       LDPC_INFO_FS_STAIR_ROWS = 7,     /* staged kernel: rows of H inside register-carried staircase runs (0: none)  */
       LDPC_INFO_FS_VARIANT = 8 };      /* staged kernel, last launch: rows per consumer group (1 | 2) | 16 if the two-pass
                                           wide-row body ran | 32 if the messages were compressed | 256 x consumers per CTA */

/* LLR quantisation (ref: CFastFixConversion::generate code/x86/CFixPointConversion/CFastFixConversion.cpp:55-65;
 * GPU twin LDPC_Convert_Float_LLR_to_8b_Fixed_Point code/gpu_fixed/decoder_template/GPU_Scheduled_functions.cu:54-64):
 * q = clamp((int)(llr_scale*y), -sat_llr, sat_llr).  Host buffers; runs on the GPU. */
int  ldpc_b200_quantize(ldpc_handle h, const float* y, int8_t* q, size_t count);

/* decode(frames, iterations) with HOST buffers, blocking (ref: CGPU_Decoder_OMS_SIMD::decode
 * code/gpu_fixed/decoder_oms/CGPU_Decoder_OMS_SIMD.cu:97-149; CDecoder_OMS_fixed_SSE::decode CDecoder_OMS_fixed_SSE.cpp:114-120).
 * llr: int8 [frames][n] (int16 [frames][n] for LDPC_DTYPE_I16, float for F32).  hard: [frames][n] bytes or [frames][ceil(n/8)].
 * iters_done: nullable, [frames] — iterations actually executed per frame (new: no reference API returns it).
 * Large batches are pipelined internally over the handle's stream slots (H2D / decode / D2H overlap). */
int  ldpc_b200_decode(ldpc_handle h, const void* llr, uint8_t* hard, size_t frames, int iters, uint8_t* iters_done);

/* the multi-stream pipeline, exposed (ref: CGPU_Decoder_MS_SIMD::decode_stream code/gpu_fixed/decoder_ms/CGPU_Decoder_MS_SIMD.cu:219-275;
 * intent of code/ldpc_multiStream/queue/handler.cpp:51-138).  Buffers should be pinned (ldpc_b200_host_alloc) to overlap. */
int  ldpc_b200_decode_async(ldpc_handle h, int slot, const void* llr, uint8_t* hard, size_t frames, int iters, uint8_t* iters_done);
int  ldpc_b200_sync(ldpc_handle h, int slot /* -1 = all */);
int  ldpc_b200_host_alloc(void** p, size_t bytes);              /* pinned; ref: CUDA_MALLOC_HOST custom_cuda.cu:30-60, CTrame.cpp:38-41 */
/* the same for buffers the host only WRITES and the GPU reads (LLR input): write-combined pinned memory — no cache snooping on the
 * way to the device; reading it back on the CPU is slow.  Freed with ldpc_b200_host_free. */
int  ldpc_b200_host_alloc_input(void** p, size_t bytes);
int  ldpc_b200_host_free(void* p);
/* device buffers for the device-resident entry points below, for C hosts that do not link the CUDA runtime themselves
 * (ref: CUDA_MALLOC_DEVICE code/gpu_fixed/custom_api/custom_cuda.cu:62-142).  Allocated on the handle's device. */
int  ldpc_b200_device_alloc(ldpc_handle h, void** p, size_t bytes);
int  ldpc_b200_device_free(ldpc_handle h, void* p);

/* CUDA streams across this ABI are cudaStream_t passed as void*.  For every entry point that takes a HANDLE, NULL means the
 * handle's slot-0 stream (a non-blocking stream: it is NOT ordered against the legacy default stream).  ldpc_b200_stream returns
 * a slot's stream so that a C host can put calls that take no handle (ldpc_b200_encode_device) on the same stream. */
void* ldpc_b200_stream(ldpc_handle h, int slot);

/* decode with DEVICE buffers on a caller-supplied CUDA stream (NULL = handle's slot-0 stream).
 * This is what `value` in bench.py times.  d_iters_done nullable.  The call works in the scratch state of slot 0 (frame-parallel
 * and generic kernels: posteriors, messages, iteration counts): consecutive calls on DIFFERENT streams, or a call followed by
 * decode()/decode_async(slot 0), are ordered one behind the other by an event (they never overlap); the caller's own buffers
 * are the caller's to order. */
int  ldpc_b200_decode_device(ldpc_handle h, const void* d_llr, uint8_t* d_hard, size_t frames, int iters,
                             uint8_t* d_iters_done, void* cuda_stream);

/* parity-check access to the decoder state after the last blocking decode of <= capacity frames: final posteriors
 * [frames][n] and check-to-variable messages [frames][m], frame-major, widened to the dtype's storage (int8 / int16 / float).
 * Must be enabled with ldpc_b200_set_debug(h, 1) before decoding.  (ref: protected var_nodes/var_mesgs of
 * CDecoder_fixed_SSE.h:37-38, device_V/d_MSG_C_2_V of CGPUDecoder.h:22-23) */
int  ldpc_b200_set_debug(ldpc_handle h, int enable);
int  ldpc_b200_debug_state(ldpc_handle h, void* posteriors, void* msgs, size_t frames);

/* synthetic channel + counters on the device (ref: GenerateNoiseAndTransform code/gpu_fixed/awgn_channel/CChanel_AWGN_SIMD.cu:7-30;
 * CErrorAnalyzer::generate code/gpu_fixed/ber_analyzer/CErrorAnalyzer.cpp:119-159).  All-zero codeword, BPSK 0 -> -1,
 * y = -1 + sigma*n, quantised with the handle's llr_scale/sat_llr.  Counter-based RNG: frame f of (seed) is reproducible. */
int  ldpc_b200_awgn_device(ldpc_handle h, void* d_llr, size_t frames, float sigma, uint64_t seed, uint64_t first_frame, void* cuda_stream);
int  ldpc_b200_awgn(ldpc_handle h, void* llr_host, size_t frames, float sigma, uint64_t seed, uint64_t first_frame);
/* counts over the first (n - n_checks) positions of every frame, all-zero codeword assumed; out[0]=bit errors, out[1]=frame errors */
int  ldpc_b200_count_errors_device(ldpc_handle h, const uint8_t* d_hard, size_t frames, uint64_t* out2_host, void* cuda_stream);

/* systematic encoder derived from the parity-check table (ref: the `-encoder` option, code/x86/main_p.cpp:232-233,
 * EncoderLibrary code/x86/CEncoder/EncoderLibrary.h, GenericEncoder::encode code/x86/CEncoder/GenericEncoder.cpp:38-78 — a DVB-S2
 * IRA encoder driven by a second table; this one needs only H).  Information bits = the first n - n_checks positions, parity
 * positions solved from H c = 0 (peeling, then a dense GF(2) inverse for what peeling cannot reach).  LDPC_ERR_UNSUPPORTED when
 * the last n_checks columns of H are singular.  Bits are bytes in {0,1}, frame-major.  Runs on the GPU. */
typedef struct ldpc_b200_encoder_s* ldpc_encoder;
int  ldpc_b200_encoder_create(ldpc_encoder* e, const ldpc_code_t* code, int device);
void ldpc_b200_encoder_destroy(ldpc_encoder e);
const char* ldpc_b200_encoder_last_error(ldpc_encoder e);      /* nullable: last create() error of this thread */
int  ldpc_b200_encoder_info(ldpc_encoder e, int* n_phases, int* dense_unknowns);
int  ldpc_b200_encode(ldpc_encoder e, const uint8_t* info /*[frames][n - n_checks]*/, uint8_t* codeword /*[frames][n]*/, size_t frames);   /* host buffers */
/* device buffers; d_info nullable = counter-based random information bits of (seed, first_frame), first_frame % 32 == 0
 * (ref: rand()%2 in GenericEncoder.cpp:47-51).  The encoder has no handle, so here cuda_stream == NULL is the legacy default
 * stream AND the call returns only when the codeword is complete (a decoder handle's streams are non-blocking streams that the
 * default stream does not order); to pipeline, pass an explicit stream — ldpc_b200_stream(h, 0) puts it on the decoder's. */
int  ldpc_b200_encode_device(ldpc_encoder e, const uint8_t* d_info, uint8_t* d_codeword, size_t frames, uint64_t seed, uint64_t first_frame, void* cuda_stream);
/* channel and counters for a transmitted codeword: BPSK 0 -> -1, 1 -> +1 (ref: CChanelAWGN_MKL.cpp:129-139), same noise stream as
 * ldpc_b200_awgn_device for the same (seed, frame); errors counted against the codeword over the information part
 * (ref: CErrorAnalyzer::generate, buf_en_bits, code/x86/CErrorAnalyzer/CErrorAnalyzer.cpp:123-137) */
int  ldpc_b200_awgn_codeword_device(ldpc_handle h, void* d_llr, const uint8_t* d_codeword, size_t frames, float sigma, uint64_t seed, uint64_t first_frame, void* cuda_stream);
int  ldpc_b200_count_errors_ref_device(ldpc_handle h, const uint8_t* d_hard, const uint8_t* d_codeword, size_t frames, uint64_t* out2_host, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_H */
