// kernel_fp.cuh — "frame-parallel" decoder: one thread owns 4 frames and walks every row of H in reference order.
//
// This is the universal path: it is bit-exact for ANY code table because rows are visited strictly in table order
// (DVB-S2 64800x32400 in reference order is a 32 399-deep dependency chain — SURVEY App. C — so it has no intra-frame
// parallelism to exploit).  Decoder state lives in HBM in a frame-interleaved layout so that every access of a warp is one
// 128-byte line:
//     V  [n][T] u32 : byte k of word (n,t) = biased posterior (v - lo) of frame 4t+k          (ref layout it replaces:
//     MSG[e][T] u32 : byte k of word (e,t) = message + 128 of frame 4t+k                       GPU_Transpose_uint8.cu:120-126,
//                                                                                              CUDA_OMS_SIMD.cu:41,84-85)
// Differences from the reference kernel (ref: code/gpu_fixed/decoder_oms/cuda/CUDA_OMS_SIMD.cu:25-262): no block barriers
// (the reference refills a shared iTable behind two __syncthreads per row); edge indices are warp-uniform read-only loads;
// arithmetic is the exact binary16x2 formulation of rowops.cuh instead of emulated byte SIMD; first iteration skips the
// message loads (as the reference's peeled iteration does); optional per-frame syndrome early termination.
// Roofline: HBM.  Algorithmic bytes per frame-iteration = 4*M (posterior r/w + message r/w, 1 B each per edge).
#pragma once
#include "rowops.cuh"

namespace ldpcb200 {

struct FpArgs {
    uint32_t* V;
    uint32_t* MSG;
    const uint32_t* pos;
    uint8_t* iters_done;     // [4*T], nullable
    int T;                   // words per variable (= threads), multiple of 32
    int n, m, nb_deg;
    int deg[LDPC_MAX_DEG_CLASSES];
    int rows[LDPC_MAX_DEG_CLASSES];
    int iters;
    uint32_t exp_word;       // 0x64646464 (rowops.cuh: bytes01_to_w)
    ldpc_params_t prm;
};

#define FP_BLOCK 128

// ---- the arithmetic of one row on packed words: wv[j] = biased posteriors of 4 frames, wm[j] = messages + 128 -----------
// G = frame pairs per word: 2 (four frames in 32 bits) or 1 (two frames in the low 16 bits; the staged kernel's small-batch variant)
template <int SEM, int ALGO, int D, bool FIRST, bool ET, bool Q, int G = 2>
__device__ __forceinline__ void fp_row_math(const uint32_t (&wv)[D], const uint32_t (&wm)[D], const RowConsts& K, uint32_t keep_lo, uint32_t keep_hi,
                                            uint32_t (&nv)[D], uint32_t (&nm)[D])
{
    uint32_t ov[2][D], om[2][D];
    const h2 inv256 = h2_const(1.0f / 256.0f), half = h2_const(0.5f), m4 = h2_const(-4.0f);
#pragma unroll
    for (int g = 0; g < G; g++) {
        h2 xu[D], a[D];
        uint32_t f[D];
#pragma unroll
        for (int j = 0; j < D; j++) {
            h2 wU = g ? bytes23_to_w(wv[j], K.c64) : bytes01_to_w(wv[j], K.c64);
            h2 nM = m4;                                                        // -(0) - 4
            if (!FIRST) { h2 wM = g ? bytes23_to_w(wm[j], K.c64) : bytes01_to_w(wm[j], K.c64); nM = __hfma2(wM, __hneg2(inv256), half); }
            xu[j] = __hmin2(__hfma2_sat(wU, inv256, nM), K.top);               // clamp(v - m) in the biased domain
        }
        RowState s;
        row_pass1<SEM, ALGO, Q, D>(xu, a, f, s, K);                            // pairwise (min1, min2) merge, 3-input parity xors
        RowOut o; row_finish<SEM, ALGO>(s, D, K, K.msg_c, o);
        RowOutS q; fold_sign(o, q);
#pragma unroll
        for (int j = 0; j < D; j++) {
            h2 msg, unew;
            pass2_edge_s(xu[j], a[j], f[j], q, K, msg, unew);
            ov[g][j] = q_to_w(unew, 0.0f);
            om[g][j] = q_to_w(msg, 128.0f);
        }
    }
#pragma unroll
    for (int j = 0; j < D; j++) {
        if (G == 2) { nv[j] = pack_bytes(ov[0][j], ov[1][j]); nm[j] = pack_bytes(om[0][j], om[1][j]); }
        else { nv[j] = __byte_perm(ov[0][j], 0u, 0x4420); nm[j] = __byte_perm(om[0][j], 0u, 0x4420); }
        if (ET) {   // frozen frames keep their state
            const uint32_t keep = __byte_perm(keep_lo, keep_hi, 0x6420);
            nv[j] = (wv[j] & keep) | (nv[j] & ~keep);
            nm[j] = FIRST ? nm[j] : ((wm[j] & keep) | (nm[j] & ~keep));
        }
    }
}

// ---- the same row on COMPRESSED messages (staged kernel, rows of degree <= 8) ------------------------------------------
// A min-sum row sends only two magnitudes: c1 to the edge(s) that held the minimum, c2 to the others, each with its own sign.  The row's
// D message words are therefore stored as FOUR words per row and thread (4 frames), whatever D:
//     cm[g]     (g = frame pair 0 | 1)  bytes { c1 frame 2g, c2 frame 2g, c1 frame 2g+1, c2 frame 2g+1 }
//     cm[2 + g]                         halves: bits 0..7 = "edge j gets c2" (the d of pass 2), bits 8..15 = "message j is negative"
// Re-expansion of edge j is a shift that brings both of its bits to bit 7 / bit 15 of the half, a sign-replicating PRMT (the c2 mask),
// a bit select and the sign xor — all exact, the expanded message is the word the uncompressed kernel would have loaded (a zero
// magnitude may carry a sign: -0 adds like +0).  Packing costs three FMA-pipe operations per edge: the bits are ACCUMULATED as
// subnormal binary16 numbers (bit j of a half = 2^(j-24); d and the sign indicator are 0.0 / 1.0, the sums stay below 2^-16: exact).
template <int SEM, int ALGO, int D, bool FIRST, bool Q>
__device__ __forceinline__ void fp_expand_pair(uint32_t cw, uint32_t es, const RowConsts& K, h2 (&nM)[D])
{
    const h2 inv256 = h2_const(1.0f / 256.0f), m4 = h2_const(-4.0f);
    const uint32_t c1q = h2_bits(__hfma2(bits_h2(__byte_perm(cw, K.c64, 0x4240)), inv256, m4));       // c1/256 per half
    const uint32_t c2q = h2_bits(__hfma2(bits_h2(__byte_perm(cw, K.c64, 0x4341)), inv256, m4));
#pragma unroll
    for (int j = 0; j < D; j++) {
        const uint32_t x = es << (7 - j);
        uint32_t mask;                                                       // 0xFFFF per half where bit 7 of its low byte is set
        asm("prmt.b32 %0, %1, %1, 0xAA88;" : "=r"(mask) : "r"(x));           // selector msb = replicate the byte's sign
        const uint32_t mag = (c2q & mask) | (c1q & ~mask);
        nM[j] = __hsub2(m4, bits_h2(and_xor(x, 0x80008000u, mag)));          // -(message)/256 - 4
    }
}

template <int SEM, int ALGO, int D, bool FIRST, bool ET, bool Q>
__device__ __forceinline__ void fp_row_math_c(const uint32_t (&wv)[D], const uint32_t (&cm)[4], const RowConsts& K, uint32_t keep_lo, uint32_t keep_hi,
                                              uint32_t (&nv)[D], uint32_t (&ncm)[4])
{
    static_assert(D <= 8, "compressed rows carry 8 edge bits per frame");
    uint32_t ov[2][D];
    const h2 inv256 = h2_const(1.0f / 256.0f), m4 = h2_const(-4.0f);
#pragma unroll
    for (int g = 0; g < 2; g++) {
        h2 xu[D], a[D], nM[D];
        uint32_t f[D];
        if (!FIRST) fp_expand_pair<SEM, ALGO, D, FIRST, Q>(cm[g], cm[2 + g], K, nM);
#pragma unroll
        for (int j = 0; j < D; j++) {
            const h2 wU = g ? bytes23_to_w(wv[j], K.c64) : bytes01_to_w(wv[j], K.c64);
            xu[j] = __hmin2(__hfma2_sat(wU, inv256, FIRST ? m4 : nM[j]), K.top);
        }
        RowState s;
        row_pass1<SEM, ALGO, Q, D>(xu, a, f, s, K);
        RowOut o; row_finish<SEM, ALGO>(s, D, K, K.msg_c, o);
        RowOutS q; fold_sign(o, q);
        h2 accd = bits_h2(0u), accs = bits_h2(0u);
#pragma unroll
        for (int j = 0; j < D; j++) {
            const h2 d = __hfma2_sat(a[j], K.k256, q.nmin1);                 // 0 where a == min1, 1 elsewhere
            const h2 smag = __hfma2(d, q.sdc, q.sc1);
            const h2 msg = bits_h2(and_xor(f[j], 0x80008000u, h2_bits(smag)));
            const h2 unew = __hmin2(__hadd2_sat(xu[j], msg), K.top);
            ov[g][j] = q_to_w(unew, 0.0f);
            const h2 bit = bits_h2(0x00010001u << j);                        // 2^(j-24) per half
            accd = __hfma2(d, bit, accd);
            accs = __hfma2(__hmul2_sat(msg, h2_const(-256.0f)), bit, accs);   // |message| >= 1/256 or zero: 1 where negative, else 0
        }
        uint32_t cw = __byte_perm(q_to_w(o.c1, 0.0f), q_to_w(__hadd2(o.c1, o.dc), 0.0f), 0x6240);
        uint32_t es = __byte_perm(h2_bits(accd), h2_bits(accs), 0x6240);
        if (ET && !FIRST) {   // frozen frames keep their state
            const uint32_t keep = g ? keep_hi : keep_lo;
            cw = (cm[g] & keep) | (cw & ~keep);
            es = (cm[2 + g] & keep) | (es & ~keep);
        }
        ncm[g] = cw; ncm[2 + g] = es;
    }
#pragma unroll
    for (int j = 0; j < D; j++) {
        nv[j] = pack_bytes(ov[0][j], ov[1][j]);
        if (ET) {
            const uint32_t keep = __byte_perm(keep_lo, keep_hi, 0x6420);
            nv[j] = (wv[j] & keep) | (nv[j] & ~keep);
        }
    }
}

// ---- one row, degree known at compile time, x kept in registers --------------------------------------------------
template <int SEM, int ALGO, int D, bool FIRST, bool ET, bool Q>
__device__ __forceinline__ void fp_row(const FpArgs& A, int t, size_t e, const RowConsts& K, uint32_t keep_lo, uint32_t keep_hi)
{
    uint32_t idx[D], wv[D], wm[D], nv[D], nm[D];
    const uint32_t T4 = 4u * (uint32_t)A.T;
    uint32_t* const vt = A.V + t;
    uint32_t* const mp = A.MSG + (e * (size_t)A.T + (size_t)t);
#pragma unroll
    for (int j = 0; j < D; j++) idx[j] = __ldg(A.pos + e + j);
#pragma unroll
    for (int j = 0; j < D; j++) wv[j] = *word_at(vt, idx[j], T4);
#pragma unroll
    for (int j = 0; j < D; j++) wm[j] = FIRST ? 0x80808080u : *word_at(mp, (uint32_t)j, T4);
    fp_row_math<SEM, ALGO, D, FIRST, ET, Q>(wv, wm, K, keep_lo, keep_hi, nv, nm);
#pragma unroll
    for (int j = 0; j < D; j++) {
        *word_at(vt, idx[j], T4) = nv[j];
        *word_at(mp, (uint32_t)j, T4) = nm[j];
    }
}

// ---- one row, run-time degree: two passes over the edges, contributions recomputed from memory in pass 2 ----------
template <int SEM, int ALGO, bool FIRST, bool ET, bool Q>
__device__ __noinline__ void fp_row_generic(const FpArgs& A, int t, size_t e, int D, const RowConsts& K, uint32_t keep_lo, uint32_t keep_hi)
{
    const h2 inv256 = h2_const(1.0f / 256.0f), half = h2_const(0.5f);
    RowState s[2]; row_begin(s[0], K); row_begin(s[1], K);
    for (int j = 0; j < D; j++) {
        const uint32_t wv = A.V[(size_t)__ldg(A.pos + e + j) * A.T + t];
        const uint32_t wm = FIRST ? 0x80808080u : A.MSG[(e + j) * A.T + t];
#pragma unroll
        for (int g = 0; g < 2; g++) {
            h2 wU = g ? bytes23_to_w(wv, K.c64) : bytes01_to_w(wv, K.c64);
            h2 wM = g ? bytes23_to_w(wm, K.c64) : bytes01_to_w(wm, K.c64);
            h2 xu = __hmin2(__hfma2_sat(wU, inv256, __hfma2(wM, __hneg2(inv256), half)), K.top);
            pass1_edge<SEM, ALGO, Q>(s[g], xu, K);
        }
    }
    RowOut o[2]; row_finish<SEM, ALGO>(s[0], D, K, K.msg_c, o[0]); row_finish<SEM, ALGO>(s[1], D, K, K.msg_c, o[1]);
    for (int j = 0; j < D; j++) {
        const size_t vi = (size_t)__ldg(A.pos + e + j) * A.T + t;
        const uint32_t wv = A.V[vi];
        const uint32_t wm = FIRST ? 0x80808080u : A.MSG[(e + j) * A.T + t];
        uint32_t ov[2], om[2];
#pragma unroll
        for (int g = 0; g < 2; g++) {
            h2 wU = g ? bytes23_to_w(wv, K.c64) : bytes01_to_w(wv, K.c64);
            h2 wM = g ? bytes23_to_w(wm, K.c64) : bytes01_to_w(wm, K.c64);
            h2 xu = __hmin2(__hfma2_sat(wU, inv256, __hfma2(wM, __hneg2(inv256), half)), K.top);
            h2 a = magnitude<SEM, ALGO, Q>(signed_contrib(xu, K), K);
            h2 msg, unew;
            pass2_edge<SEM>(xu, a, o[g], K, msg, unew);
            ov[g] = q_to_w(unew, 0.0f); om[g] = q_to_w(msg, 128.0f);
        }
        uint32_t nv = pack_bytes(ov[0], ov[1]), nm = pack_bytes(om[0], om[1]);
        if (ET) {
            const uint32_t keep = __byte_perm(keep_lo, keep_hi, 0x6420);
            nv = (wv & keep) | (nv & ~keep);
            nm = FIRST ? nm : ((wm & keep) | (nm & ~keep));
        }
        A.V[vi] = nv;
        A.MSG[(e + j) * A.T + t] = nm;
    }
}

template <int SEM, int ALGO, bool FIRST, bool ET, bool Q>
__device__ __forceinline__ void fp_class(const FpArgs& A, int t, size_t e, int D, int R, const RowConsts& K, uint32_t keep_lo, uint32_t keep_hi)
{
#define FP_CASE(DD) case DD: for (int r = 0; r < R; r++, e += DD) fp_row<SEM, ALGO, DD, FIRST, ET, Q>(A, t, e, K, keep_lo, keep_hi); break;
    switch (D) {
        FP_CASE(3) FP_CASE(4) FP_CASE(5) FP_CASE(6) FP_CASE(7) FP_CASE(8)
    default:
        for (int r = 0; r < R; r++, e += D) fp_row_generic<SEM, ALGO, FIRST, ET, Q>(A, t, e, D, K, keep_lo, keep_hi);
    }
#undef FP_CASE
}

template <int SEM, int ALGO, bool FIRST, bool ET>
__device__ __forceinline__ void fp_iteration(const FpArgs& A, int t, RowConsts& K, uint32_t keep_lo, uint32_t keep_hi)
{
    size_t e = 0;
    for (int c = 0; c < A.nb_deg; c++) {
        const int D = A.deg[c], R = A.rows[c];
        // the reference's OMS kernel forgets the 31-clamp for the second degree class in its peeled first iteration
        // (ref: CUDA_OMS_SIMD.cu:113-114 vs :73-74)
        K.msg_c = (SEM == LDPC_SEM_GPU_FIXED && ALGO == LDPC_ALGO_OMS && FIRST && c >= 1) ? K.one : K.msg;
        if (SEM == LDPC_SEM_X86_SSE && ALGO == LDPC_ALGO_OMS && c >= 1) fp_class<SEM, ALGO, FIRST, ET, true>(A, t, e, D, R, K, keep_lo, keep_hi);
        else fp_class<SEM, ALGO, FIRST, ET, false>(A, t, e, D, R, K, keep_lo, keep_hi);
        e += (size_t)D * R;
    }
}

// syndrome stop criterion: for every row the parity of (x > 0) over its edges, x = sat(v - m) with the UPDATED messages
// (ref: code/ldpc_decoder_arm/CDecoder/OMS/CDecoder_OMS_fixed_x86.cpp:150-178).  Returns bit15-of-each-half words: 1 = a check failed.
__device__ __forceinline__ void fp_syndrome(const FpArgs& A, int t, const RowConsts& K, int lo, uint32_t& bad_lo, uint32_t& bad_hi)
{
    const h2 inv256 = h2_const(1.0f / 256.0f), half = h2_const(0.5f);
    const h2 lo_np = h2_const((float)(lo - 1) / 256.0f);     // sign bit of (xu + lo_np) <=> x <= 0
    bad_lo = bad_hi = 0u;
    size_t e = 0;
    for (int c = 0; c < A.nb_deg; c++) {
        const int D = A.deg[c];
        const uint32_t dpar = (D & 1) ? 0x80008000u : 0u;    // XOR of pos flags = (D&1) ^ XOR of (x<=0) flags
        for (int r = 0; r < A.rows[c]; r++) {
            uint32_t p0 = dpar, p1 = dpar;
            for (int j = 0; j < D; j++, e++) {
                const uint32_t wv = A.V[(size_t)__ldg(A.pos + e) * A.T + t];
                const uint32_t wm = A.MSG[e * A.T + t];
                h2 x0 = __hmin2(__hfma2_sat(bytes01_to_w(wv, K.c64), inv256, __hfma2(bytes01_to_w(wm, K.c64), __hneg2(inv256), half)), K.top);
                h2 x1 = __hmin2(__hfma2_sat(bytes23_to_w(wv, K.c64), inv256, __hfma2(bytes23_to_w(wm, K.c64), __hneg2(inv256), half)), K.top);
                p0 ^= h2_bits(__hadd2(x0, lo_np));
                p1 ^= h2_bits(__hadd2(x1, lo_np));
            }
            bad_lo |= p0; bad_hi |= p1;
        }
    }
    bad_lo &= 0x80008000u; bad_hi &= 0x80008000u;
}

template <int SEM, int ALGO, bool ET>
__global__ void __launch_bounds__(FP_BLOCK) fp_decode_kernel(const __grid_constant__ FpArgs A)
{
    const int t = blockIdx.x * FP_BLOCK + threadIdx.x;
    if (t >= A.T) return;
    RowConsts K; make_consts<SEM>(K, A.prm); K.c64 = A.exp_word;
    const int lo = (SEM == LDPC_SEM_GPU_FIXED) ? -128 : -A.prm.sat_var;
    uint32_t keep_lo = 0u, keep_hi = 0u;           // 0xFFFF per half = frame frozen (early-terminated)
    uint32_t done[4] = { 0u, 0u, 0u, 0u };
    int it = 0;
    if (A.iters > 0) {
        fp_iteration<SEM, ALGO, true, ET>(A, t, K, 0u, 0u);
        it = 1;
        for (;;) {
            if (it >= A.iters) break;
            if (ET) {
                uint32_t b0, b1;
                fp_syndrome(A, t, K, lo, b0, b1);
                // frames that pass now and were not frozen before stop at iteration `it`
                const uint32_t pass_lo = ~b0 & 0x80008000u, pass_hi = ~b1 & 0x80008000u;
                if ((pass_lo & 0x00008000u) && !done[0]) done[0] = it;
                if ((pass_lo & 0x80000000u) && !done[1]) done[1] = it;
                if ((pass_hi & 0x00008000u) && !done[2]) done[2] = it;
                if ((pass_hi & 0x80000000u) && !done[3]) done[3] = it;
                keep_lo = (done[0] ? 0x0000FFFFu : 0u) | (done[1] ? 0xFFFF0000u : 0u);
                keep_hi = (done[2] ? 0x0000FFFFu : 0u) | (done[3] ? 0xFFFF0000u : 0u);
                if (done[0] && done[1] && done[2] && done[3]) break;
            }
            fp_iteration<SEM, ALGO, false, ET>(A, t, K, keep_lo, keep_hi);
            it++;
        }
    }
    if (A.iters_done) {
#pragma unroll
        for (int k = 0; k < 4; k++) A.iters_done[4 * (size_t)t + k] = (uint8_t)((ET && done[k]) ? done[k] : it);
    }
}

}  // namespace ldpcb200
