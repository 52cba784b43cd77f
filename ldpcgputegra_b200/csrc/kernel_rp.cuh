// kernel_rp.cuh — "row-parallel" on-chip decoder for codes whose whole state fits in shared memory.
//
// One WARP owns one frame PAIR (the two halves of every __half2) for all iterations; nothing but LLR-in and bits-out
// touches HBM (the reference keeps posteriors AND messages in global memory and re-reads them every iteration:
// code/gpu_fixed/decoder_oms/cuda/CUDA_OMS_SIMD.cu:160-187).  Lanes are rows: the layered schedule in reference row order
// is cut into LEVELS of mutually independent rows (ldpc_b200_level_schedule), so running a level's rows
// concurrently and the levels in order gives bit-identical results to the sequential reference loop
// (ref: code/x86/CDecoder/OMS/CDecoder_OMS_fixed_SSE.cpp:156-554).  A level is executed in "steps" of <= 32 rows of one
// degree; only a __syncwarp separates levels.
//
// Shared memory per CTA:  idx[M] u16 (step-transposed edge table) | steps | per warp: U[n] h2 (biased posteriors),
// MS[M] h2 (messages, step-transposed: edge j of the row on lane z of a step lives at msg_off + j*nrows + z so that both
// the u16 index loads and the message accesses of a warp are conflict-free).
// Roofline: SM issue / shared-memory pipe; HBM traffic = N + N (or N/8) bytes per frame, once.
#pragma once
#include "rowops.cuh"

namespace ldpcb200 {

#define RP_MAX_THREADS 768

struct RpStep {
    int32_t deg;        // degree of every row of this step
    int32_t cls;        // degree class (0 = first) — selects the X86_SSE quirk / GPU first-iteration clamp
    int32_t nrows;      // rows (= active lanes) in this step, <= 32
    int32_t msg_off;    // offset of this step's block in idx[] / MS[]
    int32_t sync;       // 1 = a new level starts here: __syncwarp before
    int32_t pad[3];
};

struct RpArgs {
    const int8_t* llr;        // [frames][n] frame-major
    uint8_t* hard;            // [frames][n] or [frames][ceil(n/8)]
    uint8_t* iters_done;      // nullable [frames]
    int8_t* dbg_post;         // nullable [frames][n]
    int8_t* dbg_msgs;         // nullable [frames][m]   (reference edge order)
    const uint16_t* idx_t;    // [m]  step-transposed variable indices
    const uint32_t* edge_of;  // [m]  step-transposed -> reference edge number (debug only)
    const RpStep* steps;
    size_t frames;
    int n, m, nsteps, n_pad;  // n_pad: U row length in words
    int iters;
    int packed;
    ldpc_params_t prm;
};

template <int SEM, int ALGO, int D, bool FIRST, bool ET, bool Q>
__device__ __forceinline__ void rp_row(h2* __restrict__ U, h2* __restrict__ MS, const uint16_t* __restrict__ idx, int a0, int stride,
                                       const RowConsts& K, uint32_t keep)
{
    int vi[D];
    h2 xu[D], a[D], mo[D], uo[D];
    RowState s; row_begin(s, K);
#pragma unroll
    for (int j = 0; j < D; j++) vi[j] = idx[a0 + j * stride];
#pragma unroll
    for (int j = 0; j < D; j++) uo[j] = U[vi[j]];
    if (!FIRST) {
#pragma unroll
        for (int j = 0; j < D; j++) mo[j] = MS[a0 + j * stride];
    }
#pragma unroll
    for (int j = 0; j < D; j++) {
        xu[j] = FIRST ? uo[j] : __hmin2(__hadd2_sat(uo[j], __hneg2(mo[j])), K.top);   // FIRST: m = 0 and U is already inside the rails
        a[j] = pass1_edge<SEM, ALGO, Q>(s, xu[j], K);
    }
    RowOut o; row_finish<SEM, ALGO>(s, D, K, o);
#pragma unroll
    for (int j = 0; j < D; j++) {
        h2 msg, unew;
        pass2_edge<SEM>(xu[j], a[j], o, K, msg, unew);
        if (ET) {   // frozen frame (keep = 0xFFFF in its half) retains posterior and message
            unew = bits_h2((h2_bits(uo[j]) & keep) | (h2_bits(unew) & ~keep));
            if (!FIRST) msg = bits_h2((h2_bits(mo[j]) & keep) | (h2_bits(msg) & ~keep));
        }
        U[vi[j]] = unew;
        MS[a0 + j * stride] = msg;
    }
}

// run-time degree (> 8): two passes, contributions recomputed in pass 2
template <int SEM, int ALGO, bool FIRST, bool ET, bool Q>
__device__ __noinline__ void rp_row_generic(h2* __restrict__ U, h2* __restrict__ MS, const uint16_t* __restrict__ idx, int a0, int stride, int D,
                                            const RowConsts& K, uint32_t keep)
{
    RowState s; row_begin(s, K);
    for (int j = 0; j < D; j++) {
        h2 u = U[idx[a0 + j * stride]];
        h2 xu = FIRST ? u : __hmin2(__hadd2_sat(u, __hneg2(MS[a0 + j * stride])), K.top);
        pass1_edge<SEM, ALGO, Q>(s, xu, K);
    }
    RowOut o; row_finish<SEM, ALGO>(s, D, K, o);
    for (int j = 0; j < D; j++) {
        const int vi = idx[a0 + j * stride];
        h2 u = U[vi], mold = FIRST ? h2_const(0.0f) : MS[a0 + j * stride];
        h2 xu = FIRST ? u : __hmin2(__hadd2_sat(u, __hneg2(mold)), K.top);
        h2 a = magnitude<SEM, ALGO, Q>(signed_contrib(xu, K), K);
        h2 msg, unew;
        pass2_edge<SEM>(xu, a, o, K, msg, unew);
        if (ET) {
            unew = bits_h2((h2_bits(u) & keep) | (h2_bits(unew) & ~keep));
            if (!FIRST) msg = bits_h2((h2_bits(mold) & keep) | (h2_bits(msg) & ~keep));
        }
        U[vi] = unew;
        MS[a0 + j * stride] = msg;
    }
}

template <int SEM, int ALGO, bool FIRST, bool ET, bool Q>
__device__ __forceinline__ void rp_step(const RpStep& st, h2* U, h2* MS, const uint16_t* idx, int lane, const RowConsts& K, uint32_t keep)
{
    if (lane >= st.nrows) return;
    const int a0 = st.msg_off + lane, stride = st.nrows;
#define RP_CASE(DD) case DD: rp_row<SEM, ALGO, DD, FIRST, ET, Q>(U, MS, idx, a0, stride, K, keep); break;
    switch (st.deg) {
        RP_CASE(3) RP_CASE(4) RP_CASE(5) RP_CASE(6) RP_CASE(7) RP_CASE(8)
    default: rp_row_generic<SEM, ALGO, FIRST, ET, Q>(U, MS, idx, a0, stride, st.deg, K, keep);
    }
#undef RP_CASE
}

template <int SEM, int ALGO, bool FIRST, bool ET>
__device__ __forceinline__ void rp_iteration(const RpStep* steps, int nsteps, h2* U, h2* MS, const uint16_t* idx, int lane, RowConsts& K, uint32_t keep)
{
    for (int s = 0; s < nsteps; s++) {
        const RpStep st = steps[s];
        if (st.sync) __syncwarp();
        K.msg_c = (SEM == LDPC_SEM_GPU_FIXED && ALGO == LDPC_ALGO_OMS && FIRST && st.cls >= 1) ? K.one : K.msg;
        if (SEM == LDPC_SEM_X86_SSE && ALGO == LDPC_ALGO_OMS && st.cls >= 1) rp_step<SEM, ALGO, FIRST, ET, true>(st, U, MS, idx, lane, K, keep);
        else rp_step<SEM, ALGO, FIRST, ET, false>(st, U, MS, idx, lane, K, keep);
    }
    __syncwarp();
}

// per-frame syndrome criterion (see fp_syndrome): returns a word whose bit 15 / bit 31 says "some check of frame 0 / 1 failed"
__device__ __forceinline__ uint32_t rp_syndrome(const RpStep* steps, int nsteps, const h2* U, const h2* MS, const uint16_t* idx, int lane,
                                                const RowConsts& K, int lo)
{
    const h2 lo_np = h2_const((float)(lo - 1) / 256.0f);
    uint32_t bad = 0u;
    for (int s = 0; s < nsteps; s++) {
        const RpStep st = steps[s];
        if (lane < st.nrows) {
            uint32_t p = (st.deg & 1) ? 0x80008000u : 0u;
            for (int j = 0; j < st.deg; j++) {
                const int a = st.msg_off + j * st.nrows + lane;
                h2 xu = __hmin2(__hadd2_sat(U[idx[a]], __hneg2(MS[a])), K.top);
                p ^= h2_bits(__hadd2(xu, lo_np));
            }
            bad |= p;
        }
    }
    bad &= 0x80008000u;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) bad |= __shfl_xor_sync(0xFFFFFFFFu, bad, d);
    return bad;
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

template <int SEM, int ALGO, bool ET>
__global__ void __launch_bounds__(RP_MAX_THREADS, 1) rp_decode_kernel(const __grid_constant__ RpArgs A)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: steps | idx_t | per-warp { U[n_pad] , MS[m] }
    RpStep* steps = reinterpret_cast<RpStep*>(smem_raw);
    uint16_t* idx = reinterpret_cast<uint16_t*>(steps + A.nsteps);
    const size_t idx_bytes = (((size_t)A.m * 2 + 15) / 16) * 16;
    h2* state = reinterpret_cast<h2*>(reinterpret_cast<unsigned char*>(idx) + idx_bytes);
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_pad = ((A.m + 3) / 4) * 4;
    h2* U = state + (size_t)warp * (A.n_pad + m_pad);
    h2* MS = U + A.n_pad;

    for (int i = threadIdx.x; i < A.nsteps * (int)(sizeof(RpStep) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(steps)[i] = reinterpret_cast<const uint32_t*>(A.steps)[i];
    for (int i = threadIdx.x; i < A.m; i += blockDim.x) idx[i] = A.idx_t[i];
    __syncthreads();

    RowConsts K; make_consts<SEM>(K, A.prm);
    const int lo = (SEM == LDPC_SEM_GPU_FIXED) ? -128 : -A.prm.sat_var;
    const int hi = (SEM == LDPC_SEM_ARM_SCALAR) ? A.prm.sat_var : 127;
    const size_t pairs = (A.frames + 1) / 2;
    for (size_t pair = (size_t)blockIdx.x * warps + warp; pair < pairs; pair += (size_t)gridDim.x * warps) {
        const size_t f0 = 2 * pair;
        rp_load_pair(A, f0, U, lane, lo, hi);
        __syncwarp();
        int it = 0;
        uint32_t done0 = 0u, done1 = 0u, keep = 0u;
        if (A.iters > 0) {
            rp_iteration<SEM, ALGO, true, ET>(steps, A.nsteps, U, MS, idx, lane, K, 0u);
            it = 1;
            while (it < A.iters) {
                if (ET) {
                    const uint32_t bad = rp_syndrome(steps, A.nsteps, U, MS, idx, lane, K, lo);
                    if (!(bad & 0x00008000u) && !done0) done0 = it;
                    if (!(bad & 0x80000000u) && !done1) done1 = it;
                    keep = (done0 ? 0x0000FFFFu : 0u) | (done1 ? 0xFFFF0000u : 0u);
                    if (done0 && done1) break;
                }
                rp_iteration<SEM, ALGO, false, ET>(steps, A.nsteps, U, MS, idx, lane, K, keep);
                it++;
            }
        }
        rp_store_pair(A, f0, U, lane, lo);
        if (A.iters_done && lane == 0) {
            A.iters_done[f0] = (uint8_t)((ET && done0) ? done0 : it);
            if (f0 + 1 < A.frames) A.iters_done[f0 + 1] = (uint8_t)((ET && done1) ? done1 : it);
        }
        if (A.dbg_post) {
            for (int i = lane; i < A.n; i += 32) {
                const uint32_t w = q_to_w(U[i], 0.0f);
                A.dbg_post[f0 * (size_t)A.n + i] = (int8_t)((int)(w & 0xFFu) + lo);
                if (f0 + 1 < A.frames) A.dbg_post[(f0 + 1) * (size_t)A.n + i] = (int8_t)((int)((w >> 16) & 0xFFu) + lo);
            }
        }
        if (A.dbg_msgs) {
            for (int i = lane; i < A.m; i += 32) {
                const uint32_t w = (A.iters > 0) ? q_to_w(MS[i], 128.0f) : 0x00800080u;
                const size_t e = A.edge_of[i];
                A.dbg_msgs[f0 * (size_t)A.m + e] = (int8_t)((int)(w & 0xFFu) - 128);
                if (f0 + 1 < A.frames) A.dbg_msgs[(f0 + 1) * (size_t)A.m + e] = (int8_t)((int)((w >> 16) & 0xFFu) - 128);
            }
        }
        __syncwarp();
    }
}

}  // namespace ldpcb200
