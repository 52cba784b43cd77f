// rowops.cuh — the check-node / variable-node arithmetic of one layered min-sum row, two frames per 32-bit register.
//
// Representation.  The reference computes in saturating int8 (ref: code/x86/CDecoder/OMS/
// CDecoder_OMS_fixed_SSE.cpp:28-81; code/gpu_fixed/decoder_oms/cuda/CUDA_OMS_SIMD.cu:160-187 on SIMD-in-a-word PTX).
// sm_100a has no byte SIMD (the __v*4 intrinsics expand to 5-10 LOP3/PRMT/IADD each) but issues HFMA2-class ops on the
// FMA pipe and HMNMX2/LOP3/PRMT on the ALU pipe concurrently, 64 lanes/clk/SM each (profiles/r01_pipe_microbench.jsonl).
// So every int8 quantity q is carried as the fp16 number q/256 in one half of a __half2:
//   - all values are multiples of 2^-8 with magnitude < 2, so every add, multiply-by-+-1 and min/max below is EXACT;
//   - HFMA2.SAT clamps to [0,1]: with posteriors stored biased, U = (v - lo)/256, one fused op performs
//     "subtract message, saturate at the lower rail" and a single HMNMX2 finishes the upper rail;
//   - the outgoing sign is a multiplication by +-1.0 (FMA pipe) instead of compare/select/negate chains;
//   - the "is this edge the minimum" select is d = SAT((a - min1)*256) in {0,1} followed by one FMA.
// Bit-exactness against the integer oracle is a theorem about exact small-integer arithmetic in binary16, and is
// checked exhaustively by tests/test_parity_gpu.py (posteriors and messages, saturating inputs included).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/ldpc_b200.h"

namespace ldpcb200 {

typedef __half2 h2;

__device__ __forceinline__ uint32_t h2_bits(h2 v) { return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ h2 bits_h2(uint32_t u) { return *reinterpret_cast<h2*>(&u); }
__device__ __forceinline__ h2 h2_const(float f) { return __float2half2_rn(f); }

// Per-launch constants (all exact in binary16).  Built once per thread; they live in (uniform) registers.
struct RowConsts {
    h2 top;        // (hi - lo)/256: upper rail of the biased posterior
    h2 lo;         // lo/256 (negative): T = Xu + lo is the signed contribution / 256
    h2 lo_flag;    // X86/UNIFORM: lo/256 ; GPU/ARM: (lo - 1)/256 — the value whose SIGN BIT is the parity flag
    h2 msg;        // sat_msg/256
    h2 msg_c;      // clamp applied to the row constants c1/c2 (sat_msg/256, or 1.0 where the reference forgets it)
    h2 off;        // -offset/256
    h2 min_init;   // running-min initial value / 256
    h2 k256;       // 256.0
    h2 one;        // 1.0
    uint32_t num1, sh1, num2, sh2;   // NMS rescale of min2 (-> c1) and min1 (-> c2): (k*num)>>sh
    uint32_t c64;  // 0x64646464, from a kernel argument (see bytes01_to_w)
};

template <int SEM>
__device__ __forceinline__ void make_consts(RowConsts& K, const ldpc_params_t& p)
{
    const int lo = (SEM == LDPC_SEM_GPU_FIXED) ? -128 : -p.sat_var;
    const int hi = (SEM == LDPC_SEM_ARM_SCALAR) ? p.sat_var : 127;
    const int sat_msg = (SEM == LDPC_SEM_GPU_FIXED) ? 31 : p.sat_msg;
    K.top = h2_const((float)(hi - lo) / 256.0f);
    K.lo = h2_const((float)lo / 256.0f);
    K.lo_flag = h2_const((float)((SEM == LDPC_SEM_GPU_FIXED || SEM == LDPC_SEM_ARM_SCALAR) ? lo - 1 : lo) / 256.0f);
    K.msg = h2_const((float)sat_msg / 256.0f);
    K.msg_c = K.msg;
    K.off = h2_const(-(float)((SEM == LDPC_SEM_GPU_FIXED) ? 1 : p.offset) / 256.0f);
    const int mi = (SEM == LDPC_SEM_ARM_SCALAR) ? p.sat_var + 1 : ((SEM == LDPC_SEM_GPU_FIXED) ? 127 : p.sat_var);
    K.min_init = h2_const((float)mi / 256.0f);
    K.k256 = h2_const(256.0f);
    K.c64 = 0x64646464u;       // the frame-parallel kernels overwrite this with their exp_word argument
    K.one = h2_const(1.0f);
    if (SEM == LDPC_SEM_GPU_FIXED) {
        // trunc(min*0.75f) = (3*min)>>2 ; trunc(min*0.875f) = (7*min)>>3  (ref: CUDA_NMS_SIMD.cu:76-83, CUDA_2NMS_SIMD.cu:76-83)
        K.num2 = 3; K.sh2 = 2;
        if (p.algo == LDPC_ALGO_2NMS) { K.num1 = 7; K.sh1 = 3; } else { K.num1 = 3; K.sh1 = 2; }
    } else {
        K.num1 = K.num2 = (uint32_t)p.factor_q5; K.sh1 = K.sh2 = 5;   // (ref: CDecoder_NMS_fixed_SSE.cpp:196-208)
    }
}

// running state of one row for one register (= two frames)
struct RowState {
    h2 min1, min2;
    uint32_t par;     // XOR of the flag words: bit 15 of each half is the parity
};

__device__ __forceinline__ void row_begin(RowState& s, const RowConsts& K) { s.min1 = K.min_init; s.min2 = K.min_init; s.par = 0u; }

// signed contribution / 256
__device__ __forceinline__ h2 signed_contrib(h2 xu, const RowConsts& K) { return __hadd2(xu, K.lo); }

// magnitude entering the min search (see oracle/ldpc_oracle.c: magnitude())
// Q = "quirk row": X86_SSE OMS rows of degree class >= 1 clamp before the abs (ref: CDecoder_OMS_fixed_SSE.cpp:211 vs :293)
template <int SEM, int ALGO, bool Q>
__device__ __forceinline__ h2 magnitude(h2 t, const RowConsts& K)
{
    if (SEM == LDPC_SEM_GPU_FIXED) return __habs2(t);
    if (SEM == LDPC_SEM_X86_SSE && ALGO == LDPC_ALGO_OMS && Q) return __habs2(__hmin2(t, K.msg));
    return __hmin2(__habs2(t), K.msg);
}

// pass 1, one edge: xu = biased saturated contribution (already clamped to [0, top])
template <int SEM, int ALGO, bool Q>
__device__ __forceinline__ h2 pass1_edge(RowState& s, h2 xu, const RowConsts& K)
{
    h2 t = signed_contrib(xu, K);
    h2 a = magnitude<SEM, ALGO, Q>(t, K);
    h2 old = s.min1;
    s.min1 = __hmin2(s.min1, a);
    s.min2 = __hmin2(s.min2, __hmax2(a, old));
    if (SEM == LDPC_SEM_GPU_FIXED || SEM == LDPC_SEM_ARM_SCALAR) s.par ^= h2_bits(__hadd2(xu, K.lo_flag));   // sign bit <=> x <= 0
    else s.par ^= h2_bits(t);                                                                                 // sign bit <=> x < 0
    return a;
}

// (k*num)>>sh on the integer value of q (q = k/256, 0 <= k <= 128), clamped to 127, back to /256.  Integer route because
// k*num exceeds binary16's exact range.
__device__ __forceinline__ h2 rescale(h2 q, uint32_t num, uint32_t sh, const RowConsts& K)
{
    uint32_t w = h2_bits(__hfma2(q, K.k256, h2_const(1024.0f))) & 0x03FF03FFu;   // 0x6400|k -> k
    uint32_t p = (w * num) >> sh;
    p &= (0xFFFFu >> sh) * 0x00010001u;
    p = __vminu2(p, 0x007F007Fu);
    return __hfma2(bits_h2(p | 0x64006400u), h2_const(1.0f / 256.0f), h2_const(-4.0f));
}

struct RowOut {
    h2 c1;       // magnitude for the edge(s) holding min1
    h2 dc;       // c2 - c1
    h2 nmin1;    // -min1*256
    uint32_t sgn;  // bit 15 of each half = row parity ^ degree parity ; other bits = 1.0 (0x3C00)
};

// (a & b) ^ c in one LOP3; written in PTX so the compiler cannot re-associate the parity xors into every edge
__device__ __forceinline__ uint32_t and_xor(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x6A;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int SEM, int ALGO>
__device__ __forceinline__ void row_finish(const RowState& s, int deg, const RowConsts& K, h2 msg_c, RowOut& o)
{
    h2 c1, c2;
    if (ALGO == LDPC_ALGO_OMS) {
        c1 = __hmin2(__hadd2_sat(s.min2, K.off), msg_c);       // min(max(min-offset,0), sat_msg) (ref: CDecoder_OMS_fixed_SSE.cpp:229-230)
        c2 = __hmin2(__hadd2_sat(s.min1, K.off), msg_c);
    } else if (ALGO == LDPC_ALGO_MS) {
        c1 = __hmin2(s.min2, K.msg); c2 = __hmin2(s.min1, K.msg);   // (ref: CUDA_MS_SIMD.cu:173-174)
    } else {
        c1 = rescale(s.min2, K.num1, K.sh1, K); c2 = rescale(s.min1, K.num2, K.sh2, K);
    }
    o.c1 = c1;
    o.dc = __hsub2(c2, c1);
    o.nmin1 = __hneg2(__hmul2(s.min1, K.k256));
    o.sgn = and_xor(s.par, 0x80008000u, ((deg & 1) ? 0x80008000u : 0u) | 0x3C003C00u);
}

// ---- register-resident row update used by the on-chip kernel: all D contributions of a row are in registers -------------
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// merge two new magnitudes into the running (min1, min2): 5 min/max ops per 2 edges instead of 6
__device__ __forceinline__ void track2(RowState& s, h2 a, h2 b)
{
    const h2 p = __hmin2(a, b), q = __hmax2(a, b);
    const h2 n1 = __hmin2(s.min1, p);
    s.min2 = __hmin2(__hmin2(__hmax2(s.min1, p), s.min2), q);
    s.min1 = n1;
}
__device__ __forceinline__ void track1(RowState& s, h2 a)
{
    const h2 old = s.min1;
    s.min1 = __hmin2(s.min1, a);
    s.min2 = __hmin2(s.min2, __hmax2(a, old));
}

// pass 1 over a whole row: xu[] in, a[] (magnitudes) and f[] (words whose sign bits are the parity flags) out
template <int SEM, int ALGO, bool Q, int D>
__device__ __forceinline__ void row_pass1(const h2 (&xu)[D], h2 (&a)[D], uint32_t (&f)[D], RowState& s, const RowConsts& K)
{
    row_begin(s, K);
#pragma unroll
    for (int j = 0; j < D; j++) {
        const h2 t = signed_contrib(xu[j], K);
        a[j] = magnitude<SEM, ALGO, Q>(t, K);
        f[j] = (SEM == LDPC_SEM_GPU_FIXED || SEM == LDPC_SEM_ARM_SCALAR) ? h2_bits(__hadd2(xu[j], K.lo_flag)) : h2_bits(t);
    }
#pragma unroll
    for (int j = 0; j + 1 < D; j += 2) { track2(s, a[j], a[j + 1]); s.par = xor3(s.par, f[j], f[j + 1]); }
    if (D & 1) { track1(s, a[D - 1]); s.par ^= f[D - 1]; }
}

// pass 2, one edge, flag word from pass 1
__device__ __forceinline__ void pass2_edge_f(h2 xu, h2 a, uint32_t f, const RowOut& o, const RowConsts& K, h2& msg, h2& unew)
{
    const h2 d = __hfma2_sat(a, K.k256, o.nmin1);
    const h2 mag = __hfma2(d, o.dc, o.c1);
    const h2 sigma = bits_h2(and_xor(f, 0x80008000u, o.sgn));
    msg = __hmul2(sigma, mag);
    unew = __hmin2(__hfma2_sat(sigma, mag, xu), K.top);
}

// The same with the row's sign folded into the two magnitudes once per row (sc1 = +-c1, sdc = +-(c2 - c1)): the edge's own
// sign is then a single LOP3 on the bit pattern, (f & 0x8000) ^ mag — 5 instructions per edge instead of 6, one FMA-pipe op
// fewer.  A zero magnitude may come out as -0: it adds and compares like +0, and no flag is ever taken from a message.
struct RowOutS { h2 sc1, sdc, nmin1; };
__device__ __forceinline__ void fold_sign(const RowOut& o, RowOutS& q)
{
    const h2 rs = bits_h2(o.sgn);             // +-1.0 per half: row parity ^ degree parity
    q.sc1 = __hmul2(rs, o.c1); q.sdc = __hmul2(rs, o.dc); q.nmin1 = o.nmin1;
}
__device__ __forceinline__ void pass2_edge_s(h2 xu, h2 a, uint32_t f, const RowOutS& q, const RowConsts& K, h2& msg, h2& unew)
{
    const h2 d = __hfma2_sat(a, K.k256, q.nmin1);
    const h2 smag = __hfma2(d, q.sdc, q.sc1);
    msg = bits_h2(and_xor(f, 0x80008000u, h2_bits(smag)));
    unew = __hmin2(__hadd2_sat(xu, msg), K.top);
}

// pass 2, one edge: returns the new message (signed, /256) and the new biased posterior
template <int SEM>
__device__ __forceinline__ void pass2_edge(h2 xu, h2 a, const RowOut& o, const RowConsts& K, h2& msg, h2& unew)
{
    h2 flag = (SEM == LDPC_SEM_GPU_FIXED || SEM == LDPC_SEM_ARM_SCALAR) ? __hadd2(xu, K.lo_flag) : signed_contrib(xu, K);
    h2 d = __hfma2_sat(a, K.k256, o.nmin1);                    // 0 where a == min1, 1 elsewhere
    h2 mag = __hfma2(d, o.dc, o.c1);                           // c1 or c2
    h2 sigma = bits_h2(and_xor(h2_bits(flag), 0x80008000u, o.sgn)); // +-1.0
    msg = __hmul2(sigma, mag);
    unew = __hmin2(__hfma2_sat(sigma, mag, xu), K.top);        // saturating add at both rails
}

// ---- storage conversions -------------------------------------------------------------------------------------------
// A byte b in [0,255] becomes the binary16 number 1024+b by byte-permuting it under the exponent byte 0x64.
// PRMT encodes ONE immediate: with the literal 0x64646464 as the second source ptxas keeps the SELECTOR in a register and
// re-materialises it before every use (first SASS of kernel_fs: 25 extra moves per degree-7 row; an asm "mov" is folded just the
// same).  The exponent word therefore arrives as a kernel argument (RowConsts::c64, host-set to 0x64646464), which ptxas cannot
// fold, and the selector becomes the immediate.
__device__ __forceinline__ h2 bytes01_to_w(uint32_t word, uint32_t c64) { return bits_h2(__byte_perm(word, c64, 0x4140)); }
__device__ __forceinline__ h2 bytes23_to_w(uint32_t word, uint32_t c64) { return bits_h2(__byte_perm(word, c64, 0x4342)); }
// w = 1024 + b  ->  b/256
__device__ __forceinline__ h2 w_to_q(h2 w) { return __hfma2(w, h2_const(1.0f / 256.0f), h2_const(-4.0f)); }
// q (multiple of 1/256, integer part q*256 + bias in [0,255]) -> binary16 bits 0x6400|byte
__device__ __forceinline__ uint32_t q_to_w(h2 q, float bias) { return h2_bits(__hfma2(q, h2_const(256.0f), h2_const(1024.0f + bias))); }
__device__ __forceinline__ uint32_t pack_bytes(uint32_t w_lo, uint32_t w_hi) { return __byte_perm(w_lo, w_hi, 0x6420); }

// clamp int8 x4 to [lo, hi] and bias by -lo (result bytes in [0, hi-lo])
__device__ __forceinline__ uint32_t bias_bytes(uint32_t w, int lo, int hi)
{
    if (lo == -128 && hi == 127) return w ^ 0x80808080u;
    const uint32_t L = (uint32_t)(lo & 0xFF) * 0x01010101u, H = (uint32_t)(hi & 0xFF) * 0x01010101u;
    w = __vmaxs4(__vmins4(w, H), L);
    return __vsub4(w, L);
}

// base + idx * pitch_bytes, written so that it compiles to ONE instruction (IMAD.WIDE.U32 with the 64-bit base as addend).
// V[(size_t)idx * T + t] with an int T costs five (wide multiply, high-part multiply for the sign extension, shift pair, add pair).
__device__ __forceinline__ uint32_t* word_at(uint32_t* base, uint32_t idx, uint32_t pitch_bytes)
{
    return reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(base) + (uint64_t)idx * (uint64_t)pitch_bytes);
}

// Shared memory is addressed through 32-bit shared-window offsets and explicit ld/st.shared in the hot loop: generic 64-bit
// pointers cost 3-4 extra integer instructions per access (first profile: profiles/r01_ncu_rp_v2_g11.txt, 36.7 warp
// instructions per edge against 21 in the row body itself).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds_u128(uint32_t a) { uint4 v; asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

}  // namespace ldpcb200
