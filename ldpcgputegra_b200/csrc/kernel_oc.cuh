// kernel_oc.cuh — the generic engine with the whole decoder state ON CHIP (kernel 5): int16 / float / flooding for codes
// short enough that a handful of frames fit in shared memory (576x288 float flooding: 11.9 KB per frame, 19 frames per SM).
//
// Same arithmetic as kernel_gp.cuh — it calls the very same gp_row / gp_row_rt functions, through generic pointers that
// happen to point into shared memory — but nothing except LLR-in and bits-out touches HBM, and the intra-frame parallelism of
// the schedule is used: a CTA owns F frames, a task is (row, frame) with the frame index fastest (consecutive lanes hit
// consecutive words: V[n][F], MSG[e][F], LLR[n][F] as fp32), the layered schedule runs level by level
// (ldpc_b200_level_schedule: rows of one level share no variable, so running them concurrently and the levels in order is
// bit-identical to the sequential reference loop), the flooding schedule is one level of all rows followed by a
// variable-node phase over (variable, frame) tasks.  A __syncthreads separates levels / phases.
// What it replaces in the reference: nothing — the reference has no int16, float or flooding decoder; the
// thing it is measured against is kernel_gp.cuh's HBM-resident version of the same modes.
// Roofline: SM issue slots; HBM sees sizeof(S)*N bytes in and N (or N/8) bytes out per frame, once.
#pragma once
#include "kernel_gp.cuh"

namespace ldpcb200 {

#define OC_MAX_THREADS 1024         // the host picks the CTA size that needs the fewest rounds per iteration (multiple of 32); 64 registers x 1024 fit one SM
#define OC_MAXF 64

struct OcRow { uint32_t e0; uint16_t deg; uint16_t cls; };

template <class S>
struct OcArgs {
    const S* llr;            // [frames][n] frame-major, boundary type
    uint8_t* hard;           // [frames][n] or [frames][ceil(n/8)]
    uint8_t* iters_done;     // nullable [frames]
    S* dbg_post;             // nullable [frames][n]
    S* dbg_msgs;             // nullable [frames][m]
    const uint32_t* pos;
    const int32_t* cptr; const int32_t* cedge;
    const OcRow* rows;       // rows in level order
    const int32_t* level_ptr;   // [nlevels + 1]
    size_t frames;
    int n, m, n_checks, nlevels, F, iters, flooding, et, packed, threads;
    int packed_syn;          // stop criterion on packed hard-decision words (float or flooding, F <= 32): n words of shared memory behind the state
    GpMode md;
};

// flooding, variable-node half for one (variable, frame): clamp(llr + sum of the column's new messages), ascending edge order
template <bool FLT>
__device__ __forceinline__ void oc_vn_one(const GpMode& md, float* V, const float* MSG, const float* LLR, const int32_t* cptr, const int32_t* cedge, int F, int f, int n)
{
    float s = LLR[n * F + f];
    const int k1 = __ldg(cptr + n + 1);
    for (int k = __ldg(cptr + n); k < k1; k++) s = __fadd_rn(s, MSG[__ldg(cedge + k) * F + f]);
    if (!FLT) s = gp_clamp(s, md.lo, md.hi);
    V[n * F + f] = s;
}

// Hard decisions of F <= 32 frames of one variable as ONE word (bit f = frame f): lanes tid = slot * F + f of a slot are consecutive,
// so a ballot hands every run of lanes that shares a slot its bits at once; the run's first lane ORs them into hb[nn] (a slot's F
// frames straddle at most two warps, and two slots never share a word).  Every thread of the CTA must call this (full-mask ballot).
__device__ __forceinline__ void oc_pack_bits(uint32_t* hb, int nn, bool b, bool mine, int F, int f)
{
    const uint32_t bits = __ballot_sync(0xFFFFFFFFu, b);
    const int lane = (int)(threadIdx.x & 31u);
    if (mine && (lane == 0 || f == 0)) {
        const int cnt = min(F - f, 32 - lane);
        const uint32_t w = ((bits >> lane) & (cnt >= 32 ? 0xFFFFFFFFu : ((1u << cnt) - 1u))) << f;
        if (w) atomicOr(hb + nn, w);
    }
}

// A thread keeps ONE frame f = tid % F and one task slot = tid / F for the whole decode and walks rows (variables) slot, slot + R,
// ...: the first version derived (row, frame) from a flat task number with a run-time division per task, and ncu showed the two
// light phases — variable nodes and stop criterion, 3 to 7 loads per task — spending most of their instructions on that
// (profiles/r01_ncu_oc_v2.txt: 38 % + 18 % of the kernel's instructions against 36 % for the check-node phase).
template <bool WRITE_V, bool FLT>
__device__ __forceinline__ void oc_level(const GpMode& md, float* V, float* MSG, const uint32_t* pos, const OcRow* rows, int r0, int nr, int F, int f, int slot, int R, bool first)
{
    for (int ri = slot; ri < nr; ri += R) {
        const OcRow row = rows[r0 + ri];
#define OC_CASE(DD) case DD: gp_row<float, DD, WRITE_V, FLT, true>(md, V, MSG, pos, F, f, row.e0, row.cls, first); break;
        switch (row.deg) {
            OC_CASE(3) OC_CASE(4) OC_CASE(5) OC_CASE(6) OC_CASE(7) OC_CASE(8)
        default: gp_row_rt<float, WRITE_V, FLT, true>(md, V, MSG, pos, F, f, row.e0, row.deg, row.cls, first);
        }
#undef OC_CASE
    }
}

template <class S>
__global__ void __launch_bounds__(OC_MAX_THREADS, 1) oc_decode_kernel(const __grid_constant__ OcArgs<S> A)
{
    extern __shared__ __align__(16) float oc_smem[];
    __shared__ int s_done[OC_MAXF], s_bad[OC_MAXF];
    const int F = A.F, n = A.n, m = A.m, tid = threadIdx.x;
    const int f = tid % F, slot = tid / F, R = (int)blockDim.x / F;        // this thread's frame and task slot (slot >= R: idle in the task phases)
    float* Vs = oc_smem; float* Ms = Vs + (size_t)n * F; float* Ls = Ms + (size_t)m * F;
    const bool posterior_syndrome = GpIsFloat<S>::value || A.flooding;
    uint32_t* const hb = reinterpret_cast<uint32_t*>(Ls + (A.flooding ? (size_t)n * F : 0));     // [n] packed hard decisions (packed_syn only)
    __shared__ uint32_t s_badw;

    for (size_t base = (size_t)blockIdx.x * F; base < A.frames; base += (size_t)gridDim.x * F) {
        const int valid = (int)min((size_t)F, A.frames - base);
        for (int i = tid; i < valid * n; i += (int)blockDim.x) {          // frame-major LLRs -> [n][F] fp32, clamped to the rails
            const int f = i / n, nn = i - f * n;
            float v = (float)A.llr[(base + f) * (size_t)n + nn];
            if (!GpIsFloat<S>::value) v = gp_clamp(v, A.md.lo, A.md.hi);
            Vs[(size_t)nn * F + f] = v;
            if (A.flooding) Ls[(size_t)nn * F + f] = v;
        }
        if (tid < OC_MAXF) { s_done[tid] = tid < valid ? 0 : 255; s_bad[tid] = 0; }
        if (A.packed_syn) { for (int i = tid; i < n; i += (int)blockDim.x) hb[i] = 0u; if (tid == 0) s_badw = 0u; }
        if (A.iters == 0 && A.dbg_msgs) for (int i = tid; i < m * F; i += (int)blockDim.x) Ms[i] = 0.0f;
        __syncthreads();
        int it = 0;
        while (it < A.iters) {
            const bool first = it == 0;
            const bool live = slot < R && !s_done[f];          // s_done only changes between the barriers at the end of an iteration
            for (int L = 0; L < A.nlevels; L++) {
                const int r0 = __ldg(A.level_ptr + L), nr = __ldg(A.level_ptr + L + 1) - r0;
                if (live) {
                    if (A.flooding) oc_level<false, GpIsFloat<S>::value>(A.md, Vs, Ms, A.pos, A.rows, r0, nr, F, f, slot, R, first);
                    else oc_level<true, GpIsFloat<S>::value>(A.md, Vs, Ms, A.pos, A.rows, r0, nr, F, f, slot, R, first);
                }
                __syncthreads();
            }
            const bool check = A.et && it + 1 < A.iters;          // a stop test follows this iteration
            if (A.flooding) {
                if (check && A.packed_syn) {                      // the same pass also packs the new hard decisions (uniform trip count: ballots)
                    for (int base = 0; base < n; base += R) {
                        const int nn = base + slot;
                        const bool mine = slot < R && nn < n;
                        bool b = false;
                        if (live && mine) { oc_vn_one<GpIsFloat<S>::value>(A.md, Vs, Ms, Ls, A.cptr, A.cedge, F, f, nn); b = Vs[(size_t)nn * F + f] > 0.0f; }
                        oc_pack_bits(hb, nn, b, mine, F, f);
                    }
                } else if (live) {
                    for (int nn = slot; nn < n; nn += R) oc_vn_one<GpIsFloat<S>::value>(A.md, Vs, Ms, Ls, A.cptr, A.cedge, F, f, nn);
                }
                __syncthreads();
            } else if (check && A.packed_syn) {
                for (int base = 0; base < n; base += R) {
                    const int nn = base + slot;
                    const bool mine = slot < R && nn < n;
                    oc_pack_bits(hb, nn, live && mine && Vs[(size_t)nn * F + f] > 0.0f, mine, F, f);
                }
                __syncthreads();
            }
            it++;
            if (A.et && it < A.iters) {
                if (A.packed_syn) {
                    // one task per ROW for all frames at once: XOR of the row's packed words; a set bit = that frame fails the check
                    uint32_t w = 0u;
                    for (int ri = tid; ri < A.n_checks; ri += (int)blockDim.x) {
                        const OcRow row = A.rows[ri];
                        uint32_t x = 0u;
                        for (int j = 0; j < row.deg; j++) x ^= hb[__ldg(A.pos + row.e0 + j)];
                        w |= x;
                    }
                    w = __reduce_or_sync(0xFFFFFFFFu, w);
                    if ((tid & 31) == 0 && w) atomicOr(&s_badw, w);
                    __syncthreads();
                    if (tid < F && ((s_badw >> tid) & 1u)) s_bad[tid] = 1;
                    for (int i = tid; i < n; i += (int)blockDim.x) hb[i] = 0u;       // ready for the next iteration's bits
                } else if (live) {
                    // stop criterion per frame (see gp_syndrome_ok): rows slot, slot + R, ... of frame f; failures OR-ed into s_bad
                    int bad = 0;
                    for (int ri = slot; ri < A.n_checks && !bad; ri += R) {
                        const OcRow row = A.rows[ri];
                        int par = 0;
                        for (int j = 0; j < row.deg; j++) {
                            float xx = Vs[__ldg(A.pos + row.e0 + j) * (uint32_t)F + (uint32_t)f];
                            if (!posterior_syndrome) xx = gp_clamp(xx - Ms[(row.e0 + j) * (uint32_t)F + (uint32_t)f], A.md.lo, A.md.hi);
                            par ^= (xx > 0.0f);
                        }
                        bad = par;
                    }
                    if (bad) s_bad[f] = 1;
                }
                __syncthreads();
                int running = 0;
                if (tid < F) {
                    if (!s_done[tid] && !s_bad[tid]) s_done[tid] = it;
                    running = s_done[tid] == 0;
                    s_bad[tid] = 0;
                }
                if (tid == 0) s_badw = 0u;
                if (!__syncthreads_or(running)) break;
            }
        }
        if (A.iters_done && tid < valid) A.iters_done[base + tid] = (uint8_t)((s_done[tid] > 0 && s_done[tid] < 255) ? s_done[tid] : it);
        // hard decisions (ref: bit = posterior > 0, code/x86/CTools/CTools.cpp:370)
        if (!A.packed) {
            for (int i = tid; i < valid * n; i += (int)blockDim.x) {
                const int f = i / n, nn = i - f * n;
                A.hard[(base + f) * (size_t)n + nn] = (uint8_t)(Vs[(size_t)nn * F + f] > 0.0f);
            }
        } else {
            const int nb = (n + 7) / 8, chunks = (n + 31) / 32, warp = tid >> 5, lane = tid & 31;
            for (int w = warp; w < valid * chunks; w += (int)blockDim.x / 32) {
                const int f = w / chunks, c = w - f * chunks, nn = c * 32 + lane;
                const uint32_t bits = __ballot_sync(0xFFFFFFFFu, nn < n && Vs[(size_t)nn * F + f] > 0.0f);
                if (lane < 4 && c * 4 + lane < nb) A.hard[(base + f) * (size_t)nb + c * 4 + lane] = (uint8_t)(bits >> (8 * lane));
            }
        }
        if (A.dbg_post) for (int i = tid; i < valid * n; i += (int)blockDim.x) { const int f = i / n, nn = i - f * n; GpIO<S>::st(A.dbg_post + (base + f) * (size_t)n + nn, Vs[(size_t)nn * F + f]); }
        if (A.dbg_msgs) for (int i = tid; i < valid * m; i += (int)blockDim.x) { const int f = i / m, ee = i - f * m; GpIO<S>::st(A.dbg_msgs + (base + f) * (size_t)m + ee, Ms[(size_t)ee * F + f]); }
        __syncthreads();
    }
}

}  // namespace ldpcb200
