// kernel_gp.cuh — the generic engine: one thread owns ONE frame, arithmetic in fp32, state in HBM, any code table.
//
// What it covers (everything the binary16x2 int8-layered kernels of kernel_fp/kernel_rp do not):
//   * int16 storage (LDPC_DTYPE_I16)        — the reference has no vectorised int16 decoder; the model is its scalar decoder
//                                              with run-time rails (ref: code/ldpc_decoder_arm/CDecoder/OMS/CDecoder_OMS_fixed_x86.cpp:30-32,61-148)
//   * float min-sum (LDPC_DTYPE_F32)         — declared but never defined in the reference
//                                              (ref: code/gpu_fixed/decoder_template/GPU_Scheduled_functions.h:31-34,54-61)
//   * the flooding schedule, every dtype     — named only in a banner (ref: code/gpu_fixed/main.cpp:95)
//   * int8 layered as well (params.kernel = 3), which pins this engine to the reference-checked oracle in all four semantics.
// Integer modes run in fp32 on integer-valued numbers: |values| <= 32767, products with factor_q5 <= 2^23, column sums
// < 2^24 — every operation is exact, so the results are bit-identical to the integer oracle.  The float mode issues exactly
// the oracle's operations in the oracle's order (single-rounded __fmul_rn/__fadd_rn, column sums in ascending edge order).
//
// Layout (frame-interleaved, every warp access is one contiguous line):  V[n][T], MSG[e][T], LLR[n][T] (flooding only) of S.
// Roofline: HBM.  Algorithmic bytes per frame-iteration: layered 4*M*sizeof(S); flooding (4*M + 2*N)*sizeof(S).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ldpc_b200.h"

namespace ldpcb200 {

#define GP_BLOCK 128
#define GP_MAXDEG 4096   // run-time-degree rows keep nothing per edge in registers

struct GpMode {
    int sem;            // ldpc_semantics_t (ignored when is_float)
    int algo;           // ldpc_algo_t
    int is_float;
    int wide;           // int16 storage
    float lo, hi;       // rails of contributions and posteriors (fixed point)
    float sat_msg;      // magnitude clamp
    float off;          // OMS offset (float mode: offset / llr_scale)
    float f1, f2;       // float / GPU_FIXED normalisation of min1 (-> c2) and min2 (-> c1)
    float factor;       // x86 NMS: (min * factor) >> 5
    float pack_sat;     // x86 NMS: saturation of the repack (127, or 32767 for int16)
    float min_init;
    int x86;            // zero counts positive + degree-parity term (X86_SSE, UNIFORM)
    int quirk;          // X86_SSE OMS: rows of class >= 1 use |min(x, sat_msg)|
};

template <class S>
struct GpArgs {
    S* V;                    // [n][T]
    S* MSG;                  // [m][T]
    const S* LLR;            // [n][T], flooding only
    const uint32_t* pos;     // [m] reference edge table
    const int32_t* cptr;     // [n+1] column pointers (flooding)
    const int32_t* cedge;    // [m] edges of each column, ascending
    uint8_t* iters_done;     // [T], nullable
    int T, n, m, nb_deg;
    int deg[LDPC_MAX_DEG_CLASSES];
    int rows[LDPC_MAX_DEG_CLASSES];
    int iters, flooding, et;
    GpMode md;
};

template <class S> struct GpIO;
template <> struct GpIO<float> {
    static __device__ __forceinline__ float ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct GpIO<int16_t> {
    static __device__ __forceinline__ float ld(const int16_t* p) { return (float)*p; }
    static __device__ __forceinline__ void st(int16_t* p, float v) { *p = (int16_t)__float2int_rn(v); }
};
template <> struct GpIO<int8_t> {
    static __device__ __forceinline__ float ld(const int8_t* p) { return (float)*p; }
    static __device__ __forceinline__ void st(int8_t* p, float v) { *p = (int8_t)__float2int_rn(v); }
};

template <class S> struct GpIsFloat { static const bool value = false; };
template <> struct GpIsFloat<float> { static const bool value = true; };

__device__ __forceinline__ float gp_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// magnitude entering the min search (oracle/ldpc_oracle.c: magnitude())
template <bool FLT>
__device__ __forceinline__ float gp_magnitude(const GpMode& md, float x, int cls)
{
    if (FLT || md.sem == LDPC_SEM_GPU_FIXED) return fabsf(x);
    if (md.quirk && cls >= 1) return fabsf(fminf(x, md.sat_msg));
    return fminf(fabsf(x), md.sat_msg);
}

// the two magnitudes of a row (oracle/ldpc_oracle.c: row_constants())
template <bool FLT>
__device__ __forceinline__ void gp_constants(const GpMode& md, float min1, float min2, int cls, bool first, float& c1, float& c2)
{
    if (FLT) {
        if (md.algo == LDPC_ALGO_OMS) { c1 = fmaxf(__fsub_rn(min2, md.off), 0.0f); c2 = fmaxf(__fsub_rn(min1, md.off), 0.0f); }
        else { c1 = __fmul_rn(min2, md.f2); c2 = __fmul_rn(min1, md.f1); }
        return;
    }
    if (md.sem == LDPC_SEM_GPU_FIXED) {
        switch (md.algo) {
        case LDPC_ALGO_MS: c1 = fminf(min2, 31.0f); c2 = fminf(min1, 31.0f); break;
        case LDPC_ALGO_OMS:
            c1 = fmaxf(min2 - 1.0f, 0.0f); c2 = fmaxf(min1 - 1.0f, 0.0f);
            if (!(first && cls >= 1)) { c1 = fminf(c1, 31.0f); c2 = fminf(c2, 31.0f); }
            break;
        default: c1 = truncf(__fmul_rn(min2, md.f2)); c2 = truncf(__fmul_rn(min1, md.f1)); break;   // NMS: f2 = f1 = 0.75 ; 2NMS: 0.875 / 0.75
        }
        return;
    }
    if (md.algo == LDPC_ALGO_NMS) {   // x86: (min * factor) >> 5, repacked with saturation
        c1 = fminf(floorf(min2 * md.factor * 0.03125f), md.pack_sat);
        c2 = fminf(floorf(min1 * md.factor * 0.03125f), md.pack_sat);
        return;
    }
    c1 = fminf(fmaxf(min2 - md.off, 0.0f), md.sat_msg);
    c2 = fminf(fmaxf(min1 - md.off, 0.0f), md.sat_msg);
}

// One check row for one frame, degree known at compile time: contributions stay in registers.  WRITE_V = layered (posteriors
// updated in place).
// (state pointers, stride and mode are passed one by one, not as a struct: the on-chip engine calls this with shared-memory
// pointers held in registers, and the mode must stay a kernel parameter in the constant bank)
// FLT = float mode (compile-time: float boundary type <=> float arithmetic without rails); IDX32 = the state fits 32-bit element
// offsets (shared memory): the first profile of the on-chip engine spent 10 of 52 instructions per edge on 64-bit address math.
template <bool IDX32> struct GpIdx { typedef size_t type; };
template <> struct GpIdx<true> { typedef uint32_t type; };

template <class S, int D, bool WRITE_V, bool FLT, bool IDX32>
__device__ __forceinline__ void gp_row(const GpMode& md, S* V, S* MSG, const uint32_t* pos, int T, int t, size_t e, int cls, bool first)
{
    typedef typename GpIdx<IDX32>::type ix_t;
    const bool x86 = !FLT && md.x86;
    float x[D], a[D];
    uint32_t idx[D];
    float min1 = md.min_init, min2 = md.min_init;
    int par = 0;
#pragma unroll
    for (int j = 0; j < D; j++) idx[j] = __ldg(pos + e + j);
#pragma unroll
    for (int j = 0; j < D; j++) {
        const float v = GpIO<S>::ld(V + (ix_t)idx[j] * (ix_t)T + (ix_t)t);
        const float m = first ? 0.0f : GpIO<S>::ld(MSG + (ix_t)(e + j) * (ix_t)T + (ix_t)t);
        float xx = __fsub_rn(v, m);
        if (!FLT) xx = gp_clamp(xx, md.lo, md.hi);
        x[j] = xx;
        const float aa = gp_magnitude<FLT>(md, xx, cls);
        a[j] = aa;
        const float old = min1;
        min1 = fminf(min1, aa);
        min2 = fminf(min2, fmaxf(aa, old));
        par ^= x86 ? (xx < 0.0f) : (xx > 0.0f);
    }
    float c1, c2;
    gp_constants<FLT>(md, min1, min2, cls, first, c1, c2);
    const int k = x86 ? (D & 1) : 1;       // negate = par ^ flag_j ^ k  (x86: degree parity; others: keep = par ^ pos_j)
#pragma unroll
    for (int j = 0; j < D; j++) {
        const float mag = (a[j] == min1) ? c1 : c2;
        const int flag = x86 ? (x[j] < 0.0f) : (x[j] > 0.0f);
        const float msg = (par ^ flag ^ k) ? -mag : mag;
        GpIO<S>::st(MSG + (ix_t)(e + j) * (ix_t)T + (ix_t)t, msg);
        if (WRITE_V) {
            float vn = __fadd_rn(x[j], msg);
            if (!FLT) vn = gp_clamp(vn, md.lo, md.hi);
            GpIO<S>::st(V + (ix_t)idx[j] * (ix_t)T + (ix_t)t, vn);
        }
    }
}

// run-time degree (rows wider than 8: 2048x384 has degree 32, DVB-S2 rate 1/9 degree 27): two passes, the contributions are
// recomputed from memory in the second one (L1/L2 hits) instead of being parked in local memory
template <class S, bool WRITE_V, bool FLT, bool IDX32>
__device__ __noinline__ void gp_row_rt(const GpMode& md, S* V, S* MSG, const uint32_t* pos, int T, int t, size_t e, int D, int cls, bool first)
{
    typedef typename GpIdx<IDX32>::type ix_t;
    const bool x86 = !FLT && md.x86;
    float min1 = md.min_init, min2 = md.min_init;
    int par = 0;
#pragma unroll 1
    for (int j = 0; j < D; j++) {
        const float v = GpIO<S>::ld(V + (ix_t)__ldg(pos + e + j) * (ix_t)T + (ix_t)t);
        const float m = first ? 0.0f : GpIO<S>::ld(MSG + (ix_t)(e + j) * (ix_t)T + (ix_t)t);
        float xx = __fsub_rn(v, m);
        if (!FLT) xx = gp_clamp(xx, md.lo, md.hi);
        const float aa = gp_magnitude<FLT>(md, xx, cls);
        const float old = min1;
        min1 = fminf(min1, aa);
        min2 = fminf(min2, fmaxf(aa, old));
        par ^= x86 ? (xx < 0.0f) : (xx > 0.0f);
    }
    float c1, c2;
    gp_constants<FLT>(md, min1, min2, cls, first, c1, c2);
    const int k = x86 ? (D & 1) : 1;
#pragma unroll 1
    for (int j = 0; j < D; j++) {
        const size_t vi = (ix_t)__ldg(pos + e + j) * (ix_t)T + (ix_t)t;
        const float v = GpIO<S>::ld(V + vi);
        const float m = first ? 0.0f : GpIO<S>::ld(MSG + (ix_t)(e + j) * (ix_t)T + (ix_t)t);
        float xx = __fsub_rn(v, m);
        if (!FLT) xx = gp_clamp(xx, md.lo, md.hi);
        const float mag = (gp_magnitude<FLT>(md, xx, cls) == min1) ? c1 : c2;
        const int flag = x86 ? (xx < 0.0f) : (xx > 0.0f);
        const float msg = (par ^ flag ^ k) ? -mag : mag;
        GpIO<S>::st(MSG + (ix_t)(e + j) * (ix_t)T + (ix_t)t, msg);
        if (WRITE_V) {
            float vn = __fadd_rn(xx, msg);
            if (!FLT) vn = gp_clamp(vn, md.lo, md.hi);
            GpIO<S>::st(V + vi, vn);
        }
    }
}

template <class S, bool WRITE_V>
__device__ __forceinline__ void gp_all_rows(const GpArgs<S>& A, int t, bool first)
{
    size_t e = 0;
    for (int c = 0; c < A.nb_deg; c++) {
        const int D = A.deg[c], R = A.rows[c];
#define GP_CASE(DD) case DD: for (int r = 0; r < R; r++, e += DD) gp_row<S, DD, WRITE_V, GpIsFloat<S>::value, false>(A.md, A.V, A.MSG, A.pos, A.T, t, e, c, first); break;
        switch (D) {
            GP_CASE(3) GP_CASE(4) GP_CASE(5) GP_CASE(6) GP_CASE(7) GP_CASE(8)
        default:
            for (int r = 0; r < R; r++, e += D) gp_row_rt<S, WRITE_V, GpIsFloat<S>::value, false>(A.md, A.V, A.MSG, A.pos, A.T, t, e, D, c, first);
        }
#undef GP_CASE
    }
}

// flooding, variable-node half: posterior = clamp(llr + sum of the column's new messages), ascending edge order
template <class S>
__device__ __forceinline__ void gp_vn_pass(const GpArgs<S>& A, int t)
{
    const GpMode& md = A.md;
    for (int n = 0; n < A.n; n++) {
        float s = GpIO<S>::ld(A.LLR + (size_t)n * A.T + t);
        const int k1 = __ldg(A.cptr + n + 1);
        for (int k = __ldg(A.cptr + n); k < k1; k++) s = __fadd_rn(s, GpIO<S>::ld(A.MSG + (size_t)__ldg(A.cedge + k) * A.T + t));
        if (!md.is_float) s = gp_clamp(s, md.lo, md.hi);
        GpIO<S>::st(A.V + (size_t)n * A.T + t, s);
    }
}

// stop criterion.  Fixed-point layered: parity of (sat(v - m) > 0) with the updated messages
// (ref: code/ldpc_decoder_arm/CDecoder/OMS/CDecoder_OMS_fixed_x86.cpp:150-178); float and flooding: parity of the hard decisions.
template <class S>
__device__ __forceinline__ bool gp_syndrome_ok(const GpArgs<S>& A, int t)
{
    const GpMode& md = A.md;
    const bool posterior = md.is_float || A.flooding;
    size_t e = 0;
    for (int c = 0; c < A.nb_deg; c++) {
        const int D = A.deg[c];
        for (int r = 0; r < A.rows[c]; r++) {
            int par = 0;
            for (int j = 0; j < D; j++, e++) {
                float xx = GpIO<S>::ld(A.V + (size_t)__ldg(A.pos + e) * A.T + t);
                if (!posterior) xx = gp_clamp(xx - GpIO<S>::ld(A.MSG + e * A.T + t), md.lo, md.hi);
                par ^= (xx > 0.0f);
            }
            if (par) return false;
        }
    }
    return true;
}

template <class S>
__global__ void __launch_bounds__(GP_BLOCK, 8) gp_decode_kernel(const __grid_constant__ GpArgs<S> A)
{
    const int t = blockIdx.x * GP_BLOCK + threadIdx.x;
    if (t >= A.T) return;
    int it = 0;
    while (it < A.iters) {
        if (A.flooding) { gp_all_rows<S, false>(A, t, it == 0); gp_vn_pass<S>(A, t); }
        else gp_all_rows<S, true>(A, t, it == 0);
        it++;
        if (A.et && it < A.iters && gp_syndrome_ok<S>(A, t)) break;
    }
    if (A.iters_done) A.iters_done[t] = (uint8_t)it;
}

// ---- boundary: frame-major <-> frame-interleaved, any element type ---------------------------------------------------------
// src [F][X] -> dst [X][T]; columns t >= F are zero-filled; fixed-point inputs are clamped to the rails (the reference clamps
// at first use, which is the same for every variable that participates in a check).  block (32, 8), tile 32 x 32.
template <class S>
__global__ void gp_interleave_kernel(const S* __restrict__ src, S* __restrict__ dst, size_t F, int X, int T, int do_clamp, float lo, float hi)
{
    __shared__ S tile[32][33];
    const size_t f0 = (size_t)blockIdx.x * 32;
    const int x0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const size_t f = f0 + i; const int x = x0 + threadIdx.x;
        S v = (S)0;
        if (f < F && x < X) {
            v = src[f * (size_t)X + x];
            if (do_clamp) v = (S)fminf(fmaxf((float)v, lo), hi);
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int x = x0 + i; const size_t f = f0 + threadIdx.x;
        if (x < X && f < (size_t)T) dst[(size_t)x * T + f] = tile[threadIdx.x][i];
    }
}

// src [X][T] -> dst [F][X]  (debug state: posteriors, messages)
template <class S>
__global__ void gp_deinterleave_kernel(const S* __restrict__ src, S* __restrict__ dst, size_t F, int X, int T)
{
    __shared__ S tile[32][33];
    const size_t f0 = (size_t)blockIdx.x * 32;
    const int x0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int x = x0 + i; const size_t f = f0 + threadIdx.x;
        tile[i][threadIdx.x] = (x < X && f < (size_t)T) ? src[(size_t)x * T + f] : (S)0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const size_t f = f0 + i; const int x = x0 + threadIdx.x;
        if (f < F && x < X) dst[f * (size_t)X + x] = tile[threadIdx.x][i];
    }
}

// hard decisions: V [N][T] -> bytes [F][N] in {0,1} (ref: code/x86/CTools/CTools.cpp:370) or LSB-first packed [F][ceil(N/8)]
template <class S, bool PACKED>
__global__ void gp_hard_kernel(const S* __restrict__ V, uint8_t* __restrict__ hard, size_t F, int N, int T)
{
    __shared__ uint8_t tile[32][33];
    const size_t f0 = (size_t)blockIdx.x * 32;
    const int x0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int x = x0 + i; const size_t f = f0 + threadIdx.x;
        tile[i][threadIdx.x] = (x < N && f < (size_t)T) ? (uint8_t)((float)V[(size_t)x * T + f] > 0.0f) : (uint8_t)0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const size_t f = f0 + i; const int x = x0 + threadIdx.x;
        const uint8_t b = tile[threadIdx.x][i];
        if (!PACKED) {
            if (f < F && x < N) hard[f * (size_t)N + x] = b;
        } else {
            const uint32_t w = __ballot_sync(0xFFFFFFFFu, b != 0);     // bits x0 .. x0+31 of frame f
            const int nb = (N + 7) / 8, byte = x0 / 8 + (int)threadIdx.x;
            if (threadIdx.x < 4 && f < F && byte < nb) hard[f * (size_t)nb + byte] = (uint8_t)(w >> (8 * threadIdx.x));
        }
    }
}

}  // namespace ldpcb200
