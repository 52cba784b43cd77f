// kernel_wf.cuh — the generic engine with the state on chip and ONE WARP PER FRAME (kernel 6): float min-sum, int16 storage and the
// flooding schedule for short codes (BASELINE.json configs[2]: float normalised min-sum, flooding, per-frame syndrome stop).
//
// kernel_oc.cuh (kernel 5) gives a CTA F frames and spreads (row, frame) tasks over its threads, with __syncthreads between phases:
// every frame of the CTA runs until the CTA's slowest frame has stopped, lanes that share a warp but not a row gather from unrelated
// shared-memory words (profiles/r01_ncu_oc_v3.txt: 35 % of the wavefronts were bank conflicts), and the code tables are read from
// global memory per task.  Here a warp owns one frame for its whole decode and takes the next one from a global work queue the
// moment its frame stops:
//   * lanes are ROWS (check-node pass, stop criterion) or VARIABLES (variable-node pass of the flooding schedule) of that frame;
//     messages are stored step-transposed (edge j of row z of a 32-row step at off + 32 j + z), so message accesses are one
//     wavefront and, for quasi-cyclic tables, posterior gathers of consecutive rows fall on consecutive words;
//   * nothing but __syncwarp separates phases and levels — no CTA barrier, no atomics, and early termination is per frame for real:
//     the work is the MEAN number of iterations, not the maximum over a CTA's frames;
//   * step descriptors, edge table and column table live in shared memory as 16-bit byte offsets.
// Arithmetic = kernel_gp.cuh's (fp32; integer modes on integer-valued floats, float mode in the oracle's operation order with
// single-rounded operations, column sums in ascending edge order), so results are compared bit for bit with oracle/ldpc_oracle.c.
// What it replaces in the reference: nothing — the reference has no float, int16 or flooding decoder; parity unpinned
// for float and flooding, pinned for int16 (K7) and int8 layered.
// Roofline: SM issue slots; HBM sees sizeof(S)*N bytes in and N (or N/8) bytes out per frame, once.
#pragma once
#include "kernel_gp.cuh"
#include "rowops.cuh"

namespace ldpcb200 {

#define WF_MAX_WARPS 24            // warps (= frames in flight) per CTA: 768 threads leave 85 registers per thread

struct WfRun {             // 16 bytes, read as one 128-bit shared load: consecutive 32-row steps of ONE degree class (and, layered, one level)
    uint16_t deg, cls;     // row degree, degree class
    uint16_t nsteps, last; // steps in the run; rows in its last step (every other step has 32)
    uint32_t off;          // element offset of the run's first step in idx_t[] / MSG[]: edge j of row z of step s at off + 32 (s deg + j) + z
    uint32_t sync;         // 1 = a new level starts here (layered schedule): the warp synchronises first
};
static_assert(sizeof(WfRun) == 16, "WfRun must stay 16 bytes");

struct WfVRun {            // 16 bytes: consecutive 32-variable steps of ONE column degree (flooding: variable-node pass; variables sorted by degree)
    uint16_t dv, pad;
    uint16_t nsteps, last;
    uint32_t off;          // element offset in cm_t[]: k-th message of variable z of step s at off + 32 (s dv + k) + z
    uint32_t voff;         // element offset in var_t[]: variable z of step s at voff + 32 s + z
};
static_assert(sizeof(WfVRun) == 16, "WfVRun must stay 16 bytes");

template <class S>
struct WfArgs {
    const S* llr;              // [frames][n] frame-major, boundary type
    uint8_t* hard;             // [frames][n] or [frames][ceil(n/8)]
    uint8_t* iters_done;       // nullable [frames]
    S* dbg_post;               // nullable [frames][n]
    S* dbg_msgs;               // nullable [frames][m]
    const WfRun* runs;         // [nruns]
    const uint16_t* idx_t;     // [m_elems] byte offset of the edge's posterior inside V (4 * variable)
    const WfVRun* vruns;       // [nvruns]
    const uint16_t* cm_t;      // [cm_elems] byte offset of the column's k-th message inside MSG (ascending reference edge order)
    const uint16_t* var_t;     // [var_elems] byte offset (4 * variable) of the variables in variable-node-pass order
    const uint32_t* edge_of;   // [m_elems] step-transposed -> reference edge, 0xFFFFFFFF = padding (debug only)
    unsigned int* counter;     // work queue: next frame to decode (zeroed before the launch)
    size_t frames;
    int n, m, n_pad, m_elems, nruns, nvruns, cm_elems, var_elems, iters, flooding, et, packed;
    int pair;                  // two steps of a run per pass (the PAIR instantiation): chosen by the host when an SM holds fewer than 12 frames
    uint32_t off_vruns, off_idx, off_cm, off_var, off_state;      // byte offsets of the shared-memory regions (each 16-byte aligned)
    GpMode md;
};

// defined in inst_wf.cu (its own translation unit so that the six instantiations compile beside the others); return cudaError_t as int
int launch_wf(const WfArgs<float>& a, int blocks, int threads, size_t smem, cudaStream_t st);
int launch_wf(const WfArgs<int16_t>& a, int blocks, int threads, size_t smem, cudaStream_t st);
int launch_wf(const WfArgs<int8_t>& a, int blocks, int threads, size_t smem, cudaStream_t st);

__device__ __forceinline__ float lds_f32(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }

// sign bit of the result <=> x > 0 (x = +-0 gives +0): the oracle's "x > 0" flag as a bit pattern that can be XOR-ed
__device__ __forceinline__ uint32_t wf_pos_flag(float x) { return __float_as_uint(__fsub_rn(0.0f, x)); }

// R check rows of R consecutive steps of a run at once (degree known at compile time): all loads first, then the arithmetic of the R
// independent rows (the compiler interleaves them: instruction-level parallelism is what a warp-per-frame mapping is short of —
// profiles/r02_ncu_wf_v2.txt: 18 warps per SM, 30 % of the stall samples on fixed-latency dependencies, 20 % on shared-memory loads),
// then the stores.  vb = shared address of the frame's V, msa = shared address of MSG[off + lane] of the first row, ixa = shared
// address of idx_t[off + lane]; the next step's row is 32 D elements further.  WRITE_V = layered schedule (posteriors updated in
// place; rows of one run belong to one level, so they share no variable).
template <bool FLT, int D, bool WRITE_V, int R>
__device__ __forceinline__ void wf_rows(const GpMode& md, uint32_t vb, uint32_t msa, uint32_t ixa, int cls, bool first)
{
    uint32_t ua[R][D];
    float x[R][D], a[R][D], msg[R][D];
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int j = 0; j < D; j++) ua[r][j] = vb + lds_u16(ixa + 64 * (D * r + j));
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int j = 0; j < D; j++) {
            float xx = __fsub_rn(lds_f32(ua[r][j]), lds_f32(msa + 128 * (D * r + j)));
            if (!FLT) xx = gp_clamp(xx, md.lo, md.hi);
            x[r][j] = xx;
            a[r][j] = gp_magnitude<FLT>(md, xx, cls);
        }
#pragma unroll
    for (int r = 0; r < R; r++) {
        float min1 = md.min_init, min2 = md.min_init;
#pragma unroll
        for (int j = 0; j + 1 < D; j += 2) {      // pairwise merge: the two smallest values (with multiplicity) whatever the order
            const float p = fminf(a[r][j], a[r][j + 1]), q = fmaxf(a[r][j], a[r][j + 1]);
            min2 = fminf(fminf(fmaxf(min1, p), min2), q);
            min1 = fminf(min1, p);
        }
        if (D & 1) { const float old = min1; min1 = fminf(min1, a[r][D - 1]); min2 = fminf(min2, fmaxf(a[r][D - 1], old)); }
        float c1, c2;
        gp_constants<FLT>(md, min1, min2, cls, first, c1, c2);
        if (FLT || !md.x86) {
            // keep the sign iff par ^ (x > 0), par = XOR of all (x > 0) flags — float, ARM_SCALAR and GPU_FIXED (oracle: update_row)
            uint32_t par = 0x80000000u;                                  // the complement, so that `neg` below comes out directly
#pragma unroll
            for (int j = 0; j < D; j++) par ^= wf_pos_flag(x[r][j]);
#pragma unroll
            for (int j = 0; j < D; j++) {
                const float mag = (a[r][j] == min1) ? c1 : c2;
                const uint32_t neg = (par ^ wf_pos_flag(x[r][j])) & 0x80000000u;      // set <=> !(par ^ flag): negate
                msg[r][j] = __uint_as_float(__float_as_uint(mag) ^ neg);             // ARM_SCALAR's clamp of the signed message is folded into c1/c2 (gp_constants)
            }
        } else {
            // x86 semantics: sign bit, zero counts positive, degree-parity term (ref: CDecoder_OMS_fixed_SSE.cpp:180-190,232-244)
            int par = D & 1;
#pragma unroll
            for (int j = 0; j < D; j++) par ^= (x[r][j] < 0.0f);
#pragma unroll
            for (int j = 0; j < D; j++) {
                const float mag = (a[r][j] == min1) ? c1 : c2;
                msg[r][j] = (par ^ (int)(x[r][j] < 0.0f)) ? -mag : mag;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int j = 0; j < D; j++) {
            sts_f32(msa + 128 * (D * r + j), msg[r][j]);
            if (WRITE_V) {
                float vn = __fadd_rn(x[r][j], msg[r][j]);
                if (!FLT) vn = gp_clamp(vn, md.lo, md.hi);
                sts_f32(ua[r][j], vn);
            }
        }
}

// run-time degree (rows wider than 8): two passes, contributions recomputed from shared memory in the second
template <bool FLT, bool WRITE_V>
__device__ __noinline__ void wf_row_rt(const GpMode& md, uint32_t vb, uint32_t msa, uint32_t ixa, int D, int cls, bool first)
{
    const bool x86 = !FLT && md.x86;
    float min1 = md.min_init, min2 = md.min_init;
    int par = 0;
#pragma unroll 1
    for (int j = 0; j < D; j++) {
        float xx = __fsub_rn(lds_f32(vb + lds_u16(ixa + 64 * j)), lds_f32(msa + 128 * j));
        if (!FLT) xx = gp_clamp(xx, md.lo, md.hi);
        const float aa = gp_magnitude<FLT>(md, xx, cls), old = min1;
        min1 = fminf(min1, aa);
        min2 = fminf(min2, fmaxf(aa, old));
        par ^= x86 ? (xx < 0.0f) : (xx > 0.0f);
    }
    float c1, c2;
    gp_constants<FLT>(md, min1, min2, cls, first, c1, c2);
    const int k = x86 ? (D & 1) : 1;
#pragma unroll 1
    for (int j = 0; j < D; j++) {
        const uint32_t ua = vb + lds_u16(ixa + 64 * j);
        float xx = __fsub_rn(lds_f32(ua), lds_f32(msa + 128 * j));
        if (!FLT) xx = gp_clamp(xx, md.lo, md.hi);
        const float mag = (gp_magnitude<FLT>(md, xx, cls) == min1) ? c1 : c2;
        const int flag = x86 ? (xx < 0.0f) : (xx > 0.0f);
        const float msg = (par ^ flag ^ k) ? -mag : mag;
        sts_f32(msa + 128 * j, msg);
        if (WRITE_V) {
            float vn = __fadd_rn(xx, msg);
            if (!FLT) vn = gp_clamp(vn, md.lo, md.hi);
            sts_f32(ua, vn);
        }
    }
}

// stop criterion of one row: float and flooding — parity of the hard decisions (posterior > 0); fixed-point layered — parity of
// (sat(v - m) > 0) with the updated messages (ref: code/ldpc_decoder_arm/CDecoder/OMS/CDecoder_OMS_fixed_x86.cpp:150-178)
template <bool POSTERIOR>
__device__ __forceinline__ uint32_t wf_row_syndrome(const GpMode& md, uint32_t vb, uint32_t msa, uint32_t ixa, int D)
{
    uint32_t par = 0u;
#pragma unroll 2
    for (int j = 0; j < D; j++) {
        float xx = lds_f32(vb + lds_u16(ixa + 64 * j));
        if (!POSTERIOR) xx = gp_clamp(xx - lds_f32(msa + 128 * j), md.lo, md.hi);
        par ^= wf_pos_flag(xx);
    }
    return par & 0x80000000u;
}

// stop criterion of R rows of consecutive steps, degree known at compile time (see wf_row_syndrome)
template <bool POSTERIOR, int D, int R>
__device__ __forceinline__ uint32_t wf_rows_syndrome(const GpMode& md, uint32_t vb, uint32_t msa, uint32_t ixa)
{
    uint32_t ua[R][D];
    float xx[R][D];
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int j = 0; j < D; j++) ua[r][j] = vb + lds_u16(ixa + 64 * (D * r + j));
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int j = 0; j < D; j++) {
            xx[r][j] = lds_f32(ua[r][j]);
            if (!POSTERIOR) xx[r][j] = gp_clamp(xx[r][j] - lds_f32(msa + 128 * (D * r + j)), md.lo, md.hi);
        }
    uint32_t bad = 0u;
#pragma unroll
    for (int r = 0; r < R; r++) {
        uint32_t par = 0u;
#pragma unroll
        for (int j = 0; j + 1 < D; j += 2) par = xor3(par, wf_pos_flag(xx[r][j]), wf_pos_flag(xx[r][j + 1]));
        if (D & 1) par ^= wf_pos_flag(xx[r][D - 1]);
        bad |= par;
    }
    return bad & 0x80000000u;
}

// variable-node update of R variables (of R consecutive steps) of column degree DV (flooding): clamp(llr + sum of the column's
// messages), ascending edge order.  ca = shared address of cm_t[off + lane], ta = shared address of var_t[voff + lane].
template <bool FLT, int DV, int R>
__device__ __forceinline__ void wf_vars(const GpMode& md, uint32_t vb, uint32_t lb, uint32_t mb, uint32_t ca, uint32_t ta)
{
    uint32_t va[R], off[R][DV > 0 ? DV : 1];
    float s[R], mk[R][DV > 0 ? DV : 1];
#pragma unroll
    for (int r = 0; r < R; r++) {
        va[r] = lds_u16(ta + 64 * r);
#pragma unroll
        for (int k = 0; k < DV; k++) off[r][k] = lds_u16(ca + 64 * (DV * r + k));
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        s[r] = lds_f32(lb + va[r]);
#pragma unroll
        for (int k = 0; k < DV; k++) mk[r][k] = lds_f32(mb + off[r][k]);
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
#pragma unroll
        for (int k = 0; k < DV; k++) s[r] = __fadd_rn(s[r], mk[r][k]);
        if (!FLT) s[r] = gp_clamp(s[r], md.lo, md.hi);
    }
#pragma unroll
    for (int r = 0; r < R; r++) sts_f32(vb + va[r], s[r]);
}
template <bool FLT>
__device__ __noinline__ void wf_var_rt(const GpMode& md, uint32_t vb, uint32_t lb, uint32_t mb, uint32_t ca, uint32_t ta, int dv)
{
    const uint32_t va = lds_u16(ta);
    float s = lds_f32(lb + va);
#pragma unroll 1
    for (int k = 0; k < dv; k++) s = __fadd_rn(s, lds_f32(mb + lds_u16(ca + 64 * k)));
    if (!FLT) s = gp_clamp(s, md.lo, md.hi);
    sts_f32(vb + va, s);
}

// PAIR = two steps of a run per pass through the row code (for codes whose state leaves an SM only a few warps: the second row is the
// latency hiding the missing warps would have been — 2304x1152 float flooding, 4 warps per SM: 4.8 -> 6.4 M frames/s; with 18 warps
// per SM the doubled code only costs instruction-cache misses: 576x288 48 -> 41 M frames/s, profiles/r02_sweep_wf_v3.jsonl)
template <class S, bool FLOOD, bool PAIR>
__global__ void __launch_bounds__(WF_MAX_WARPS * 32, 1) wf_decode_kernel(const __grid_constant__ WfArgs<S> A)
{
    constexpr bool FLT = GpIsFloat<S>::value;
    extern __shared__ __align__(16) unsigned char wf_smem[];
    // layout: runs | vruns | idx_t (u16) | cm_t (u16) | var_t (u16) | per warp: V[n_pad] | LLR[n_pad] (flooding) | MSG[m_elems]
    WfRun* runs = reinterpret_cast<WfRun*>(wf_smem);
    WfVRun* vruns = reinterpret_cast<WfVRun*>(wf_smem + A.off_vruns);
    uint16_t* idx = reinterpret_cast<uint16_t*>(wf_smem + A.off_idx);
    uint16_t* cm = reinterpret_cast<uint16_t*>(wf_smem + A.off_cm);
    uint16_t* vt = reinterpret_cast<uint16_t*>(wf_smem + A.off_var);
    float* state_all = reinterpret_cast<float*>(wf_smem + A.off_state);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int frame_words = A.n_pad * (FLOOD ? 2 : 1) + A.m_elems;
    float* V = state_all + (size_t)warp * frame_words;
    float* L = V + A.n_pad;                          // flooding only
    float* M = V + A.n_pad * (FLOOD ? 2 : 1);

    for (int i = threadIdx.x; i < A.nruns * 4; i += blockDim.x) reinterpret_cast<uint32_t*>(runs)[i] = reinterpret_cast<const uint32_t*>(A.runs)[i];
    for (int i = threadIdx.x; i < A.nvruns * 4; i += blockDim.x) reinterpret_cast<uint32_t*>(vruns)[i] = reinterpret_cast<const uint32_t*>(A.vruns)[i];
    for (int i = threadIdx.x; i < A.m_elems; i += blockDim.x) idx[i] = A.idx_t[i];
    for (int i = threadIdx.x; i < A.cm_elems; i += blockDim.x) cm[i] = A.cm_t[i];
    for (int i = threadIdx.x; i < A.var_elems; i += blockDim.x) vt[i] = A.var_t[i];
    __syncthreads();

    const uint32_t runs_s = smem_u32(runs), vruns_s = smem_u32(vruns), idx_s = smem_u32(idx), cm_s = smem_u32(cm), vt_s = smem_u32(vt);
    const uint32_t vb = smem_u32(V), lb = smem_u32(L), mb = smem_u32(M);
    const GpMode& md = A.md;
    const bool posterior_syndrome = FLT || FLOOD;
    const int n = A.n;

    for (;;) {
        unsigned int fq = 0;
        if (lane == 0) fq = atomicAdd(A.counter, 1u);
        fq = __shfl_sync(0xFFFFFFFFu, fq, 0);
        if ((size_t)fq >= A.frames) break;
        const size_t f = fq;
        // frame-major LLRs -> V (and the channel copy the flooding variable-node pass starts from), clamped to the rails; MSG = 0
        const S* src = A.llr + f * (size_t)n;
        for (int i = lane; i < n; i += 32) {
            float v = (float)src[i];
            if (!FLT) v = gp_clamp(v, md.lo, md.hi);
            V[i] = v;
            if (FLOOD) L[i] = v;
        }
        for (int i = lane; i < A.m_elems; i += 32) M[i] = 0.0f;
        __syncwarp();

        int it = 0;
        while (it < A.iters) {
            const bool first = it == 0;
            // ---- check-node pass: runs of 32-row steps of one degree, lanes = rows ----
#pragma unroll 1
            for (int r = 0; r < A.nruns; r++) {
                const uint4 rd = lds_u128(runs_s + 16u * r);
                const int deg = (int)(rd.x & 0xFFFFu), cls = (int)(rd.x >> 16), nsteps = (int)(rd.y & 0xFFFFu), last = (int)(rd.y >> 16);
                if (!FLOOD && rd.w) __syncwarp();                  // a new level reads what the previous one wrote
                uint32_t msa = mb + 4u * (rd.z + (uint32_t)lane), ixa = idx_s + 2u * (rd.z + (uint32_t)lane);
#define WF_CASE(DD)                                                                                                             \
    case DD: {                                                                                                                  \
        int s = 0;                                                                                                              \
        for (; PAIR && s + 2 < nsteps; s += 2, msa += 256u * DD, ixa += 128u * DD) wf_rows<FLT, DD, !FLOOD, PAIR ? 2 : 1>(md, vb, msa, ixa, cls, first); /* full steps, two at a time */ \
        for (; s < nsteps; s++, msa += 128u * DD, ixa += 64u * DD)                                                              \
            if (s + 1 < nsteps || lane < last) wf_rows<FLT, DD, !FLOOD, 1>(md, vb, msa, ixa, cls, first);                      \
    } break;
                switch (deg) {
                    WF_CASE(3) WF_CASE(4) WF_CASE(5) WF_CASE(6) WF_CASE(7) WF_CASE(8)
                default:
                    for (int s = 0; s < nsteps; s++, msa += 128u * deg, ixa += 64u * deg)
                        if (s + 1 < nsteps || lane < last) wf_row_rt<FLT, !FLOOD>(md, vb, msa, ixa, deg, cls, first);
                }
#undef WF_CASE
            }
            __syncwarp();
            // ---- flooding: variable-node pass, lanes = variables (sorted by column degree) ----
            if (FLOOD) {
#pragma unroll 1
                for (int r = 0; r < A.nvruns; r++) {
                    const uint4 rd = lds_u128(vruns_s + 16u * r);
                    const int dv = (int)(rd.x & 0xFFFFu), nsteps = (int)(rd.y & 0xFFFFu), last = (int)(rd.y >> 16);
                    uint32_t ca = cm_s + 2u * (rd.z + (uint32_t)lane), ta = vt_s + 2u * (rd.w + (uint32_t)lane);
#define WF_VCASE(DD)                                                                                                            \
    case DD: {                                                                                                                  \
        int s = 0;                                                                                                              \
        for (; PAIR && s + 2 < nsteps; s += 2, ca += 128u * DD, ta += 128u) wf_vars<FLT, DD, PAIR ? 2 : 1>(md, vb, lb, mb, ca, ta);                \
        for (; s < nsteps; s++, ca += 64u * DD, ta += 64u)                                                                      \
            if (s + 1 < nsteps || lane < last) wf_vars<FLT, DD, 1>(md, vb, lb, mb, ca, ta);                                    \
    } break;
                    switch (dv) {
                        WF_VCASE(1) WF_VCASE(2) WF_VCASE(3) WF_VCASE(4) WF_VCASE(5) WF_VCASE(6)
                    default:
                        for (int s = 0; s < nsteps; s++, ca += 64u * dv, ta += 64u)
                            if (s + 1 < nsteps || lane < last) wf_var_rt<FLT>(md, vb, lb, mb, ca, ta, dv);
                    }
#undef WF_VCASE
                }
                __syncwarp();
            }
            it++;
            // ---- per-frame stop criterion: leaves at the first step that holds an unsatisfied check (early iterations fail at once) ----
            if (A.et && it < A.iters) {
                bool bad_any = false;
#pragma unroll 1
                for (int r = 0; r < A.nruns && !bad_any; r++) {
                    const uint4 rd = lds_u128(runs_s + 16u * r);
                    const int deg = (int)(rd.x & 0xFFFFu), nsteps = (int)(rd.y & 0xFFFFu), last = (int)(rd.y >> 16);
                    uint32_t msa = mb + 4u * (rd.z + (uint32_t)lane), ixa = idx_s + 2u * (rd.z + (uint32_t)lane);
#define WF_SCASE(DD)                                                                                                            \
    case DD:                                                                                                                    \
        for (int s = 0; s < nsteps && !bad_any; s++, msa += 128u * DD, ixa += 64u * DD) {                                       \
            uint32_t bad = 0u;                                                                                                  \
            if (s + 1 < nsteps || lane < last)                                                                                  \
                bad = posterior_syndrome ? wf_rows_syndrome<true, DD, 1>(md, vb, msa, ixa) : wf_rows_syndrome<false, DD, 1>(md, vb, msa, ixa); \
            bad_any = __any_sync(0xFFFFFFFFu, bad);                                                                             \
        }                                                                                                                       \
        break;
                    switch (deg) {
                        WF_SCASE(3) WF_SCASE(4) WF_SCASE(5) WF_SCASE(6) WF_SCASE(7) WF_SCASE(8)
                    default:
                        for (int s = 0; s < nsteps && !bad_any; s++, msa += 128u * deg, ixa += 64u * deg) {
                            uint32_t bad = 0u;
                            if (s + 1 < nsteps || lane < last)
                                bad = posterior_syndrome ? wf_row_syndrome<true>(md, vb, msa, ixa, deg) : wf_row_syndrome<false>(md, vb, msa, ixa, deg);
                            bad_any = __any_sync(0xFFFFFFFFu, bad);
                        }
                    }
#undef WF_SCASE
                }
                if (!bad_any) break;
            }
        }
        // ---- outputs ----
        if (A.iters_done && lane == 0) A.iters_done[f] = (uint8_t)it;
        if (!A.packed) {
            uint8_t* o = A.hard + f * (size_t)n;
            if ((n % 4 == 0) && ((reinterpret_cast<uintptr_t>(A.hard) & 3) == 0)) {
                for (int i = lane; i < n / 4; i += 32) {
                    const float4 q = *reinterpret_cast<const float4*>(V + 4 * i);
                    reinterpret_cast<uint32_t*>(o)[i] = (uint32_t)(q.x > 0.0f) | ((uint32_t)(q.y > 0.0f) << 8) | ((uint32_t)(q.z > 0.0f) << 16) | ((uint32_t)(q.w > 0.0f) << 24);
                }
            } else {
                for (int i = lane; i < n; i += 32) o[i] = (uint8_t)(V[i] > 0.0f);
            }
        } else {
            const int nb = (n + 7) / 8;
            uint8_t* o = A.hard + f * (size_t)nb;
            for (int base = 0; base < n; base += 32) {
                const int i = base + lane;
                const uint32_t bits = __ballot_sync(0xFFFFFFFFu, i < n && V[i] > 0.0f);
                if (lane < 4 && base / 8 + lane < nb) o[base / 8 + lane] = (uint8_t)(bits >> (8 * lane));
            }
        }
        if (A.dbg_post) for (int i = lane; i < n; i += 32) GpIO<S>::st(A.dbg_post + f * (size_t)n + i, V[i]);
        if (A.dbg_msgs) {
            for (int i = lane; i < A.m_elems; i += 32) {
                const uint32_t e = A.edge_of[i];
                if (e != 0xFFFFFFFFu) GpIO<S>::st(A.dbg_msgs + f * (size_t)A.m + e, M[i]);
            }
        }
        __syncwarp();
    }
}

}  // namespace ldpcb200
