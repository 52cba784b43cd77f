// kernel_rp.cuh — "row-parallel" on-chip decoder for codes whose whole state fits in shared memory.
//
// A GROUP of G warps owns P frame PAIRS (the two halves of every __half2) for all iterations; nothing but LLR-in and
// bits-out touches HBM (the reference keeps posteriors AND messages in global memory and re-reads them every iteration:
// code/gpu_fixed/decoder_oms/cuda/CUDA_OMS_SIMD.cu:160-187).  Lanes are (pair, row) tasks: the layered schedule in reference
// row order is cut into LEVELS of mutually independent rows (ldpc_b200_level_schedule), so running a level's
// rows concurrently and the levels in order is bit-identical to the sequential reference loop
// (ref: code/x86/CDecoder/OMS/CDecoder_OMS_fixed_SSE.cpp:156-554).  A level is executed in "steps" of <= 32 rows of one
// degree; the P*nrows tasks of a step are spread over the G*32 lanes of the group (576x288: 4 pairs x 24 rows = 3 full
// warps), and only a named barrier over the group's G warps (a __syncwarp when G = 1) separates levels.
//
// Shared memory per CTA:  steps | idx[] u16 (step-transposed byte offsets into U) | flags | per pair: U[n] h2 (biased
// posteriors), MS[] h2 (messages, step-transposed: edge j of row z of a step lives at msg_off + j*stride + z, so the index
// loads and the message accesses of a warp are conflict-free and, for the specialised (degree, stride) variants, every
// per-edge address is an immediate offset).
// Roofline: SM issue slots; HBM traffic = N bytes in + N (or N/8) bytes out per frame, once.
#pragma once
#include "rowops.cuh"

namespace ldpcb200 {

#define RP_MAX_THREADS 768
#define RP_STATIC_THREADS 576     // static plan: at most 18 warps per CTA -> 112 registers per thread
#define RP_MAX_GROUPS 15          // named barriers 1..15

struct RpStep {         // read by the kernel as one 128-bit load (first four words) — keep the field order
    uint16_t deg;       // degree of every row of this step
    uint16_t nrows;     // rows in this step (<= 32)
    uint16_t stride;    // element stride between consecutive edges of a row in idx[] / MS[] (>= nrows)
    uint8_t cls;        // degree class (0 = first): selects the X86_SSE quirk / the GPU first-iteration clamp
    uint8_t sync;       // 1 = a new level starts here: group barrier before
    uint32_t msg_off;   // element offset of this step's block in idx[] / MS[]
    uint32_t magic;     // ceil(65536 / nrows): task -> (pair, row) without a division
    uint32_t variant;   // 0 = generic (run-time degree/stride) ; 1.. = specialised (degree, stride) instantiation
    uint32_t pad[3];    // 32 bytes: keeps every later shared-memory region 16-byte aligned
};
static_assert(sizeof(RpStep) == 32, "RpStep must stay 32 bytes");

struct RpRun {
    int32_t first, count, variant, quirk;     // consecutive steps sharing one instantiation; quirk = degree class >= 1
    uint32_t msg_off0;                        // msg_off of the first step; step s of the run starts at msg_off0 + s * stride * deg
    uint32_t syncmask;                        // bit s = a new level starts at step s of the run (count <= 32)
    uint32_t pad[2];
};
static_assert(sizeof(RpRun) == 32, "RpRun must stay 32 bytes");

struct RpArgs {
    const int8_t* llr;        // [frames][n] frame-major
    uint8_t* hard;            // [frames][n] or [frames][ceil(n/8)]
    uint8_t* iters_done;      // nullable [frames]
    int8_t* dbg_post;         // nullable [frames][n]
    int8_t* dbg_msgs;         // nullable [frames][m]   (reference edge order)
    const uint16_t* idx_t;    // [m_elems]  step-transposed byte offsets (4*variable)
    const uint32_t* edge_of;  // [m_elems]  step-transposed -> reference edge number, 0xFFFFFFFF for padding (debug only)
    const RpStep* steps;
    const RpRun* runs;       // [nruns] consecutive steps sharing one instantiation
    size_t frames;
    int n, m, nsteps, nruns;
    int pair_words;           // words per pair in shared memory (n_pad + m_elems + bank-spreading pad)
    int n_pad;                // U length in words (multiple of 4)
    int m_elems;              // padded message elements per pair (multiple of 4)
    int G, P;                 // warps per group, pairs per full group
    int static_nrows;         // > 0: every step has this many rows and P * nrows <= 32 * G — lane t owns (pair t / nrows, row t % nrows) for the whole decode
    int pair_fastest;         // static plan, P = 4, pair pitch = 8 or 24 (mod 32 words): lane t owns (pair t % 4, row t / 4) instead — a warp is then
                              // 8 rows x 4 pairs whose four 8-bank windows tile the 32 banks: posterior gathers and message accesses conflict-free
    int groups;               // groups per CTA
    int slots;                // pair slots per CTA (the last group may own fewer than P)
    int iters;
    int packed;
    ldpc_params_t prm;
};

__device__ __forceinline__ void group_sync(int G, int bar_id)
{
    if (G == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(G * 32) : "memory");
}

// ---- one row, degree and stride known at compile time: every per-edge address is base + immediate --------------------
// ub = shared address of the pair's U, msa = shared address of MS[msg_off + z], ixa = shared address of idx[msg_off + z]
template <int SEM, int ALGO, int D, int NR, bool FIRST, bool ET, bool Q>
__device__ __forceinline__ void rp_row(uint32_t ub, uint32_t msa, uint32_t ixa, const RowConsts& K, h2 msg_c, uint32_t keep)
{
    uint32_t ua[D], f[D];
    h2 xu[D], a[D], mo[D], uo[D];
#pragma unroll
    for (int j = 0; j < D; j++) ua[j] = ub + lds_u16(ixa + 2 * NR * j);
#pragma unroll
    for (int j = 0; j < D; j++) uo[j] = bits_h2(lds_u32(ua[j]));
    if (!FIRST) {
#pragma unroll
        for (int j = 0; j < D; j++) mo[j] = bits_h2(lds_u32(msa + 4 * NR * j));
    }
#pragma unroll
    for (int j = 0; j < D; j++) xu[j] = FIRST ? uo[j] : __hmin2(__hadd2_sat(uo[j], __hneg2(mo[j])), K.top);   // FIRST: m = 0, U already inside the rails
    RowState s;
    row_pass1<SEM, ALGO, Q, D>(xu, a, f, s, K);
    RowOut o; row_finish<SEM, ALGO>(s, D, K, msg_c, o);
    RowOutS q; fold_sign(o, q);
#pragma unroll
    for (int j = 0; j < D; j++) {
        h2 msg, unew;
        pass2_edge_s(xu[j], a[j], f[j], q, K, msg, unew);
        if (ET) {   // frozen frame (keep = 0xFFFF in its half) retains posterior and message
            unew = bits_h2((h2_bits(uo[j]) & keep) | (h2_bits(unew) & ~keep));
            if (!FIRST) msg = bits_h2((h2_bits(mo[j]) & keep) | (h2_bits(msg) & ~keep));
        }
        sts_u32(ua[j], h2_bits(unew));
        sts_u32(msa + 4 * NR * j, h2_bits(msg));
    }
}

// ---- run-time degree / stride: two passes, contributions recomputed in pass 2 ------------------------------------------
template <int SEM, int ALGO, bool FIRST, bool ET, bool Q>
__device__ __noinline__ void rp_row_generic(uint32_t ub, uint32_t msa, uint32_t ixa, int D, int stride, const RowConsts& K, h2 msg_c, uint32_t keep)
{
    RowState s; row_begin(s, K);
    for (int j = 0; j < D; j++) {
        const h2 u = bits_h2(lds_u32(ub + lds_u16(ixa + 2 * stride * j)));
        const h2 xu = FIRST ? u : __hmin2(__hadd2_sat(u, __hneg2(bits_h2(lds_u32(msa + 4 * stride * j)))), K.top);
        pass1_edge<SEM, ALGO, Q>(s, xu, K);
    }
    RowOut o; row_finish<SEM, ALGO>(s, D, K, msg_c, o);
    for (int j = 0; j < D; j++) {
        const uint32_t ua = ub + lds_u16(ixa + 2 * stride * j);
        const h2 u = bits_h2(lds_u32(ua)), mold = FIRST ? h2_const(0.0f) : bits_h2(lds_u32(msa + 4 * stride * j));
        const h2 xu = FIRST ? u : __hmin2(__hadd2_sat(u, __hneg2(mold)), K.top);
        const h2 a = magnitude<SEM, ALGO, Q>(signed_contrib(xu, K), K);
        h2 msg, unew;
        pass2_edge<SEM>(xu, a, o, K, msg, unew);
        if (ET) {
            unew = bits_h2((h2_bits(u) & keep) | (h2_bits(unew) & ~keep));
            if (!FIRST) msg = bits_h2((h2_bits(mold) & keep) | (h2_bits(msg) & ~keep));
        }
        sts_u32(ua, h2_bits(unew));
        sts_u32(msa + 4 * stride * j, h2_bits(msg));
    }
}

// specialised variants: id = 1 + (deg - 6) * 5 + stride index, for deg in {6,7,8}, stride in {24,32,64,96,128}
__host__ __device__ __forceinline__ int rp_variant_id(int deg, int stride)
{
    const int si = stride == 24 ? 0 : stride == 32 ? 1 : stride == 64 ? 2 : stride == 96 ? 3 : stride == 128 ? 4 : -1;
    if (deg < 6 || deg > 8 || si < 0) return 0;
    return 1 + (deg - 6) * 5 + si;
}
#define RP_ALL_CASES                                                                                  \
    RP_CASE(1, 6, 24) RP_CASE(2, 6, 32) RP_CASE(3, 6, 64) RP_CASE(4, 6, 96) RP_CASE(5, 6, 128)        \
    RP_CASE(6, 7, 24) RP_CASE(7, 7, 32) RP_CASE(8, 7, 64) RP_CASE(9, 7, 96) RP_CASE(10, 7, 128)       \
    RP_CASE(11, 8, 24) RP_CASE(12, 8, 32) RP_CASE(13, 8, 64) RP_CASE(14, 8, 96) RP_CASE(15, 8, 128)

// everything the step loop needs, in registers: shared-window addresses and group geometry
struct RpCtx {
    uint32_t steps_s, idx_s, flags_s, state_s;   // shared addresses: step table, index table, this group's flags, this group's pairs
    uint32_t pair_bytes, ms_off;                 // bytes per pair; byte offset of MS inside a pair (= 4 * n_pad)
    int gl, GT, G, bar;                          // lane index in the group, lanes in the group, warps in the group, barrier id
};

// one run = consecutive steps that share the instantiation (degree, stride, quirk class)
template <int SEM, int ALGO, int D, int NR, bool FIRST, bool ET, bool Q>
__device__ __forceinline__ void rp_run(const RpCtx& c, int first, int count, int valid_pairs, const RowConsts& K)
{
#pragma unroll 1
    for (int s = first; s < first + count; s++) {
        const uint4 sd = lds_u128(c.steps_s + 32u * s);         // {deg | nrows<<16, stride | cls<<16 | sync<<24, msg_off, magic}
        const int nrows = (int)(sd.x >> 16);
        if (sd.y >> 24) group_sync(c.G, c.bar);
        const h2 msg_c = (SEM == LDPC_SEM_GPU_FIXED && ALGO == LDPC_ALGO_OMS && FIRST && ((sd.y >> 16) & 0xFF) >= 1) ? K.one : K.msg;
        const int tasks = valid_pairs * nrows;
        for (int t = c.gl; t < tasks; t += c.GT) {
            const uint32_t p = ((uint32_t)t * sd.w) >> 16;
            const uint32_t e = sd.z + ((uint32_t)t - p * nrows);
            const uint32_t ub = c.state_s + p * c.pair_bytes;
            uint32_t keep = 0u;
            if (ET) keep = lds_u32(c.flags_s + 8u * p + 4u);
            if (D > 0) rp_row<SEM, ALGO, (D > 0 ? D : 1), (D > 0 ? NR : 1), FIRST, ET, Q>(ub, ub + c.ms_off + 4u * e, c.idx_s + 2u * e, K, msg_c, keep);
            else rp_row_generic<SEM, ALGO, FIRST, ET, Q>(ub, ub + c.ms_off + 4u * e, c.idx_s + 2u * e, (int)(sd.x & 0xFFFFu), (int)(sd.y & 0xFFFFu), K, msg_c, keep);
        }
    }
}

// Static plan (all steps have the same number of rows and a full group carries at most one task per lane, e.g. 576x288:
// 4 pairs x 24 rows on 3 warps): a lane keeps its (pair, row slot) for the whole decode, so a step costs nothing but the
// row itself — no descriptor load, no task -> (pair, row) arithmetic, addresses advance by a compile-time constant.
// (profiles/r01_ncu_rp_v3_g34.txt: 32.9 warp instructions per warp-edge-update against ~22 in the row body.)
struct RpLane { uint32_t ub, ms_lane, ix_lane, keep, pair; bool active; };

// row body with the addresses and the old messages already in registers (software-pipelined caller)
template <int SEM, int ALGO, int D, int NR, bool FIRST, bool ET, bool Q>
__device__ __forceinline__ void rp_row_pre(const uint32_t (&ua)[D], const h2 (&uo)[D], const h2 (&mo)[D], uint32_t msa, const RowConsts& K, h2 msg_c, uint32_t keep)
{
    uint32_t f[D];
    h2 xu[D], a[D];
#pragma unroll
    for (int j = 0; j < D; j++) xu[j] = FIRST ? uo[j] : __hmin2(__hadd2_sat(uo[j], __hneg2(mo[j])), K.top);
    RowState s;
    row_pass1<SEM, ALGO, Q, D>(xu, a, f, s, K);
    RowOut o; row_finish<SEM, ALGO>(s, D, K, msg_c, o);
    RowOutS q; fold_sign(o, q);
#pragma unroll
    for (int j = 0; j < D; j++) {
        h2 msg, unew;
        pass2_edge_s(xu[j], a[j], f[j], q, K, msg, unew);
        if (ET) {   // frozen frame (keep = 0xFFFF in its half) retains posterior and message
            unew = bits_h2((h2_bits(uo[j]) & keep) | (h2_bits(unew) & ~keep));
            if (!FIRST) msg = bits_h2((h2_bits(mo[j]) & keep) | (h2_bits(msg) & ~keep));
        }
        sts_u32(ua[j], h2_bits(unew));
        sts_u32(msa + 4 * NR * j, h2_bits(msg));
    }
}

// Software pipeline over the steps of a run: the index loads and the old messages of step s+1 belong to this lane alone
// (no other lane writes them), so they are issued before step s's arithmetic and cross the level barrier in flight; only the
// posterior gathers have to wait for the barrier.
template <int SEM, int ALGO, int D, int NR, bool FIRST, bool ET, bool Q>
__device__ __forceinline__ void rp_run_static(const RpCtx& c, const RpRun& r, const RpLane& L, const RowConsts& K)
{
    const h2 msg_c = (SEM == LDPC_SEM_GPU_FIXED && ALGO == LDPC_ALGO_OMS && FIRST && r.quirk) ? K.one : K.msg;
    uint32_t mask = r.syncmask, msa = L.ms_lane + 4u * r.msg_off0, ixa = L.ix_lane + 2u * r.msg_off0;
    uint32_t ix[D];
    h2 mo[D];
#pragma unroll
    for (int j = 0; j < D; j++) { ix[j] = 0u; mo[j] = bits_h2(0u); }
    if (L.active) {
#pragma unroll
        for (int j = 0; j < D; j++) ix[j] = lds_u16(ixa + 2 * NR * j);
        if (!FIRST) {
#pragma unroll
            for (int j = 0; j < D; j++) mo[j] = bits_h2(lds_u32(msa + 4 * NR * j));
        }
    }
#pragma unroll 1
    for (int s = 0; s < r.count; s++, mask >>= 1, msa += 4u * NR * D, ixa += 2u * NR * D) {
        if (mask & 1u) group_sync(c.G, c.bar);
        if (L.active) {
            uint32_t ua[D];
            h2 uo[D];
#pragma unroll
            for (int j = 0; j < D; j++) { ua[j] = L.ub + ix[j]; uo[j] = bits_h2(lds_u32(ua[j])); }
            rp_row_pre<SEM, ALGO, D, NR, FIRST, ET, Q>(ua, uo, mo, msa, K, msg_c, L.keep);
            if (s + 1 < r.count) {      // the row's registers are dead here: the next step's loads cross the barrier in flight
#pragma unroll
                for (int j = 0; j < D; j++) ix[j] = lds_u16(ixa + 2 * NR * (D + j));
                if (!FIRST) {
#pragma unroll
                    for (int j = 0; j < D; j++) mo[j] = bits_h2(lds_u32(msa + 4 * NR * (D + j)));
                }
            }
        }
    }
}

template <int SEM, int ALGO, bool FIRST, bool ET, bool Q>
__device__ __forceinline__ void rp_dispatch_static(const RpCtx& c, const RpRun& r, const RpLane& L, const RowConsts& K)
{
#define RP_CASE(ID, DD, NN) case ID: rp_run_static<SEM, ALGO, DD, NN, FIRST, ET, Q>(c, r, L, K); break;
    switch (r.variant) {
        RP_ALL_CASES
    }
#undef RP_CASE
}

template <int SEM, int ALGO, bool FIRST, bool ET>
__device__ __forceinline__ void rp_iteration_static(const RpCtx& c, uint32_t runs_s, int nruns, const RpLane& L, const RowConsts& K)
{
    for (int i = 0; i < nruns; i++) {
        const uint4 a = lds_u128(runs_s + 32u * i), b = lds_u128(runs_s + 32u * i + 16u);
        RpRun r; r.first = (int)a.x; r.count = (int)a.y; r.variant = (int)a.z; r.quirk = (int)a.w; r.msg_off0 = b.x; r.syncmask = b.y;
        if (SEM == LDPC_SEM_X86_SSE && ALGO == LDPC_ALGO_OMS && r.quirk) rp_dispatch_static<SEM, ALGO, FIRST, ET, true>(c, r, L, K);
        else rp_dispatch_static<SEM, ALGO, FIRST, ET, false>(c, r, L, K);
    }
    group_sync(c.G, c.bar);
}

template <int SEM, int ALGO, bool FIRST, bool ET, bool Q>
__device__ __forceinline__ void rp_dispatch(const RpCtx& c, const RpRun& r, int valid_pairs, const RowConsts& K)
{
#define RP_CASE(ID, DD, NN) case ID: rp_run<SEM, ALGO, DD, NN, FIRST, ET, Q>(c, r.first, r.count, valid_pairs, K); break;
    switch (r.variant) {
        RP_ALL_CASES
    default: rp_run<SEM, ALGO, 0, 0, FIRST, ET, Q>(c, r.first, r.count, valid_pairs, K);
    }
#undef RP_CASE
}

template <int SEM, int ALGO, bool FIRST, bool ET>
__device__ __forceinline__ void rp_iteration(const RpCtx& c, const RpRun* runs, int nruns, int valid_pairs, const RowConsts& K)
{
    for (int i = 0; i < nruns; i++) {
        const RpRun r = runs[i];
        // x86-SSE OMS clamps before the abs for rows of degree class >= 1 (ref: CDecoder_OMS_fixed_SSE.cpp:211 vs :293)
        if (SEM == LDPC_SEM_X86_SSE && ALGO == LDPC_ALGO_OMS && r.quirk) rp_dispatch<SEM, ALGO, FIRST, ET, true>(c, r, valid_pairs, K);
        else rp_dispatch<SEM, ALGO, FIRST, ET, false>(c, r, valid_pairs, K);
    }
    group_sync(c.G, c.bar);
}

// per-frame syndrome criterion: parity of (x > 0) over every row, x = sat(v - m) with the UPDATED messages
// (ref: code/ldpc_decoder_arm/CDecoder/OMS/CDecoder_OMS_fixed_x86.cpp:150-178).  ORs "some check failed" bits into flags[2p].
__device__ __forceinline__ void rp_syndrome(const RpCtx& c, uint32_t* flags, int nsteps, int valid_pairs, const RowConsts& K, int lo)
{
    const h2 lo_np = h2_const((float)(lo - 1) / 256.0f);
    for (int s = 0; s < nsteps; s++) {
        const uint4 sd = lds_u128(c.steps_s + 32u * s);
        const int nrows = (int)(sd.x >> 16), deg = (int)(sd.x & 0xFFFFu), stride = (int)(sd.y & 0xFFFFu);
        const int tasks = valid_pairs * nrows;
        for (int t = c.gl; t < tasks; t += c.GT) {
            const uint32_t p = ((uint32_t)t * sd.w) >> 16;
            const uint32_t e = sd.z + ((uint32_t)t - p * nrows);
            const uint32_t ub = c.state_s + p * c.pair_bytes, msa = ub + c.ms_off + 4u * e, ixa = c.idx_s + 2u * e;
            uint32_t par = (deg & 1) ? 0x80008000u : 0u;
            for (int j = 0; j < deg; j++) {
                const h2 u = bits_h2(lds_u32(ub + lds_u16(ixa + 2 * stride * j)));
                const h2 xu = __hmin2(__hadd2_sat(u, __hneg2(bits_h2(lds_u32(msa + 4 * stride * j)))), K.top);
                par ^= h2_bits(__hadd2(xu, lo_np));
            }
            par &= 0x80008000u;
            if (par) atomicOr(&flags[2 * p], par);
        }
    }
}

// LLR int8 (two frames) -> biased binary16 posteriors of the pair
__device__ __forceinline__ void rp_load_pair(const RpArgs& A, size_t f0, h2* U, int lane, int lo, int hi)
{
    const int n = A.n;
    const bool have1 = f0 + 1 < A.frames;
    const int8_t* p0 = A.llr + f0 * (size_t)n;
    const int8_t* p1 = A.llr + (f0 + 1) * (size_t)n;
    if ((n % 16 == 0) && ((reinterpret_cast<uintptr_t>(A.llr) & 15) == 0)) {
        for (int c = lane; c < n / 16; c += 32) {
            const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(p0) + c);
            uint4 q1 = make_uint4(0u, 0u, 0u, 0u);
            if (have1) q1 = __ldg(reinterpret_cast<const uint4*>(p1) + c);
            const uint32_t w0[4] = { q0.x, q0.y, q0.z, q0.w }, w1[4] = { q1.x, q1.y, q1.z, q1.w };
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t b0 = bias_bytes(w0[k], lo, hi), b1 = bias_bytes(w1[k], lo, hi);
                // half 0 = frame f0, half 1 = frame f0+1: (0x6400|byte) pairs, then /256 - 4
                uint4 out;
                out.x = h2_bits(w_to_q(bits_h2((__byte_perm(b0, b1, 0x4400) & 0x00FF00FFu) | 0x64006400u)));
                out.y = h2_bits(w_to_q(bits_h2((__byte_perm(b0, b1, 0x5511) & 0x00FF00FFu) | 0x64006400u)));
                out.z = h2_bits(w_to_q(bits_h2((__byte_perm(b0, b1, 0x6622) & 0x00FF00FFu) | 0x64006400u)));
                out.w = h2_bits(w_to_q(bits_h2((__byte_perm(b0, b1, 0x7733) & 0x00FF00FFu) | 0x64006400u)));
                *reinterpret_cast<uint4*>(U + 16 * c + 4 * k) = out;
            }
        }
    } else {
        for (int i = lane; i < n; i += 32) {
            int v0 = p0[i], v1 = have1 ? p1[i] : 0;
            v0 = min(max(v0, lo), hi) - lo; v1 = min(max(v1, lo), hi) - lo;
            U[i] = w_to_q(bits_h2((uint32_t)v0 | ((uint32_t)v1 << 16) | 0x64006400u));
        }
    }
}

// hard decisions of the pair: bit = posterior > 0  <=>  biased byte > -lo
__device__ __forceinline__ void rp_store_pair(const RpArgs& A, size_t f0, const h2* U, int lane, int lo)
{
    const int n = A.n;
    const bool have1 = f0 + 1 < A.frames;
    const h2 thr = h2_const((float)(-lo) / 256.0f);
    if (!A.packed) {
        uint8_t* o0 = A.hard + f0 * (size_t)n;
        uint8_t* o1 = A.hard + (f0 + 1) * (size_t)n;
        if ((n % 16 == 0) && ((reinterpret_cast<uintptr_t>(A.hard) & 15) == 0)) {
            for (int c = lane; c < n / 16; c += 32) {
                uint32_t r0[4], r1[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint4 q = *reinterpret_cast<const uint4*>(U + 16 * c + 4 * k);
                    const uint32_t g0 = __hgt2_mask(bits_h2(q.x), thr) & 0x00010001u, g1 = __hgt2_mask(bits_h2(q.y), thr) & 0x00010001u;
                    const uint32_t g2 = __hgt2_mask(bits_h2(q.z), thr) & 0x00010001u, g3 = __hgt2_mask(bits_h2(q.w), thr) & 0x00010001u;
                    // g*: byte0 = frame0 bit, byte2 = frame1 bit
                    const uint32_t lo01 = __byte_perm(g0, g1, 0x6240), lo23 = __byte_perm(g2, g3, 0x6240);   // [g0.b0,g1.b0,g0.b2,g1.b2]
                    r0[k] = __byte_perm(lo01, lo23, 0x5410);
                    r1[k] = __byte_perm(lo01, lo23, 0x7632);
                }
                reinterpret_cast<uint4*>(o0)[c] = make_uint4(r0[0], r0[1], r0[2], r0[3]);
                if (have1) reinterpret_cast<uint4*>(o1)[c] = make_uint4(r1[0], r1[1], r1[2], r1[3]);
            }
        } else {
            for (int i = lane; i < n; i += 32) {
                const uint32_t g = __hgt2_mask(U[i], thr);
                o0[i] = (uint8_t)(g & 1u);
                if (have1) o1[i] = (uint8_t)((g >> 16) & 1u);
            }
        }
    } else {
        const int nb = (n + 7) / 8;
        uint8_t* o0 = A.hard + f0 * (size_t)nb;
        uint8_t* o1 = A.hard + (f0 + 1) * (size_t)nb;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const uint32_t g = (i < n) ? __hgt2_mask(U[i], thr) : 0u;
            const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, g & 1u), b1 = __ballot_sync(0xFFFFFFFFu, (g >> 16) & 1u);
            if (lane < 4 && base / 8 + lane < nb) {
                o0[base / 8 + lane] = (uint8_t)(b0 >> (8 * lane));
                if (have1) o1[base / 8 + lane] = (uint8_t)(b1 >> (8 * lane));
            }
        }
    }
}

// STAT = static plan (see rp_run_static): its own instantiation so that it gets the register budget of an 18-warp CTA
// (the software pipeline keeps two steps of addresses and messages live) and carries none of the descriptor-driven code.
template <int SEM, int ALGO, bool ET, bool STAT>
__global__ void __launch_bounds__(STAT ? RP_STATIC_THREADS : RP_MAX_THREADS, 1) rp_decode_kernel(const __grid_constant__ RpArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: steps | runs | idx | flags (2 words per pair slot) | pair states
    RpStep* steps = reinterpret_cast<RpStep*>(smem_raw);
    RpRun* runs = reinterpret_cast<RpRun*>(steps + A.nsteps);
    uint16_t* idx = reinterpret_cast<uint16_t*>(runs + A.nruns);
    const size_t idx_bytes = (((size_t)A.m_elems * 2 + 15) / 16) * 16;
    uint32_t* flags_all = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(idx) + idx_bytes);
    const size_t flag_bytes = (((size_t)A.slots * 8 + 15) / 16) * 16;
    unsigned char* state_all = reinterpret_cast<unsigned char*>(flags_all) + flag_bytes;
    const int pair_bytes = A.pair_words * 4;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = warp / A.G, wig = warp - grp * A.G;
    unsigned char* gstate = state_all + (size_t)grp * A.P * pair_bytes;
    uint32_t* gflags = flags_all + 2 * grp * A.P;
    const int gpairs = min(A.P, A.slots - grp * A.P);
    RpCtx c;
    c.steps_s = smem_u32(steps); c.idx_s = smem_u32(idx); c.flags_s = smem_u32(gflags); c.state_s = smem_u32(gstate);
    c.pair_bytes = (uint32_t)pair_bytes; c.ms_off = 4u * A.n_pad;
    c.gl = wig * 32 + lane; c.GT = A.G * 32; c.G = A.G; c.bar = 1 + grp;

    for (int i = threadIdx.x; i < A.nsteps * (int)(sizeof(RpStep) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(steps)[i] = reinterpret_cast<const uint32_t*>(A.steps)[i];
    for (int i = threadIdx.x; i < A.nruns * (int)(sizeof(RpRun) / 4); i += blockDim.x)
        reinterpret_cast<int32_t*>(runs)[i] = reinterpret_cast<const int32_t*>(A.runs)[i];
    for (int i = threadIdx.x; i < A.m_elems; i += blockDim.x) idx[i] = A.idx_t[i];
    __syncthreads();

    RowConsts K; make_consts<SEM>(K, A.prm);
    const int lo = (SEM == LDPC_SEM_GPU_FIXED) ? -128 : -A.prm.sat_var;
    const int hi = (SEM == LDPC_SEM_ARM_SCALAR) ? A.prm.sat_var : 127;
    // 32-bit pair arithmetic in the loop (a launch never carries 2^31 pairs; the host chunks): 64-bit bounds made the compiler
    // re-derive min(gpairs, pairs - first) with wide ops in every step
    const uint32_t pairs = (uint32_t)((A.frames + 1) / 2);
    const uint32_t slot0 = (uint32_t)(grp * A.P);
    for (uint32_t base = blockIdx.x * (uint32_t)A.slots; base < pairs; base += gridDim.x * (uint32_t)A.slots) {
        const uint32_t first = base + slot0;               // global pair index of this group's pair 0
        if (first >= pairs) break;                         // uniform over the group
        const int valid = (int)min((uint32_t)gpairs, pairs - first);
        {   // pull the NEXT set's LLRs into L2 while this set is decoded (the load below would otherwise expose HBM latency
            // to the whole group once per set)
            const size_t next = (size_t)first + (size_t)gridDim.x * A.slots;
            if (next < pairs) {
                const size_t nvalid = min((size_t)gpairs, (size_t)pairs - next);
                const size_t bytes = min(nvalid * 2 * (size_t)A.n, (A.frames - 2 * next) * (size_t)A.n);
                const char* src = reinterpret_cast<const char*>(A.llr) + 2 * next * (size_t)A.n;
                for (size_t off = (size_t)c.gl * 128; off < bytes; off += (size_t)c.GT * 128)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(src + off));
            }
        }
        for (int p = wig; p < valid; p += A.G) {
            rp_load_pair(A, 2 * ((size_t)first + p), reinterpret_cast<h2*>(gstate + (size_t)p * pair_bytes), lane, lo, hi);
            if (ET && lane == 0) { gflags[2 * p] = 0u; gflags[2 * p + 1] = 0u; }
        }
        group_sync(c.G, c.bar);
        RpLane L;                                          // static plan: this lane's (pair, row slot) for the whole decode
        const bool stat = STAT;
        {
            uint32_t lp = stat ? (uint32_t)c.gl / (uint32_t)A.static_nrows : 0u, lz = (uint32_t)c.gl - lp * (uint32_t)A.static_nrows;
            if (stat && A.pair_fastest) { lp = (uint32_t)c.gl & 3u; lz = (uint32_t)c.gl >> 2; }
            L.active = stat && (int)lp < valid && lz < (uint32_t)A.static_nrows;
            L.pair = lp;
            L.ub = c.state_s + lp * c.pair_bytes; L.ms_lane = L.ub + c.ms_off + 4u * lz; L.ix_lane = c.idx_s + 2u * lz; L.keep = 0u;
        }
        const uint32_t runs_s = smem_u32(runs);
        int it = 0;
        uint32_t done_lo = 0u, done_hi = 0u;               // ET bookkeeping of pair gl lives in lane gl of the group
        if (A.iters > 0) {
            if (stat) rp_iteration_static<SEM, ALGO, true, ET>(c, runs_s, A.nruns, L, K);
            else rp_iteration<SEM, ALGO, true, ET>(c, runs, A.nruns, valid, K);
            it = 1;
            while (it < A.iters) {
                if (ET) {
                    rp_syndrome(c, gflags, A.nsteps, valid, K, lo);
                    group_sync(c.G, c.bar);
                    if (c.gl < valid) {
                        const uint32_t bad = gflags[2 * c.gl];
                        if (!(bad & 0x00008000u) && !done_lo) done_lo = it;
                        if (!(bad & 0x80000000u) && !done_hi) done_hi = it;
                        gflags[2 * c.gl] = (done_lo && done_hi) ? 0u : 1u;                                   // "still running"
                        gflags[2 * c.gl + 1] = (done_lo ? 0x0000FFFFu : 0u) | (done_hi ? 0xFFFF0000u : 0u);  // freeze mask
                    }
                    group_sync(c.G, c.bar);
                    uint32_t running = 0u;
                    for (int p = 0; p < valid; p++) running |= gflags[2 * p];
                    group_sync(c.G, c.bar);
                    if (c.gl < valid) gflags[2 * c.gl] = 0u;       // cleared for the next syndrome pass (ordered by the barriers of the iteration)
                    if (!running) break;
                    if (stat && L.active) L.keep = lds_u32(c.flags_s + 8u * L.pair + 4u);
                }
                if (stat) rp_iteration_static<SEM, ALGO, false, ET>(c, runs_s, A.nruns, L, K);
                else rp_iteration<SEM, ALGO, false, ET>(c, runs, A.nruns, valid, K);
                it++;
            }
        }
        group_sync(c.G, c.bar);
        if (A.iters_done && c.gl < valid) {
            const size_t f0 = 2 * ((size_t)first + c.gl);
            A.iters_done[f0] = (uint8_t)((ET && done_lo) ? done_lo : it);
            if (f0 + 1 < A.frames) A.iters_done[f0 + 1] = (uint8_t)((ET && done_hi) ? done_hi : it);
        }
        for (int p = wig; p < valid; p += A.G) {
            const size_t f0 = 2 * ((size_t)first + p);
            const h2* U = reinterpret_cast<const h2*>(gstate + (size_t)p * pair_bytes);
            const h2* MS = U + A.n_pad;
            rp_store_pair(A, f0, U, lane, lo);
            if (A.dbg_post) {
                for (int i = lane; i < A.n; i += 32) {
                    const uint32_t w = q_to_w(U[i], 0.0f);
                    A.dbg_post[f0 * (size_t)A.n + i] = (int8_t)((int)(w & 0xFFu) + lo);
                    if (f0 + 1 < A.frames) A.dbg_post[(f0 + 1) * (size_t)A.n + i] = (int8_t)((int)((w >> 16) & 0xFFu) + lo);
                }
            }
            if (A.dbg_msgs) {
                for (int i = lane; i < A.m_elems; i += 32) {
                    const uint32_t e = A.edge_of[i];
                    if (e == 0xFFFFFFFFu) continue;
                    const uint32_t w = (A.iters > 0) ? q_to_w(MS[i], 128.0f) : 0x00800080u;
                    A.dbg_msgs[f0 * (size_t)A.m + e] = (int8_t)((int)(w & 0xFFu) - 128);
                    if (f0 + 1 < A.frames) A.dbg_msgs[(f0 + 1) * (size_t)A.m + e] = (int8_t)((int)((w >> 16) & 0xFFu) - 128);
                }
            }
        }
        group_sync(c.G, c.bar);
    }
}

}  // namespace ldpcb200
