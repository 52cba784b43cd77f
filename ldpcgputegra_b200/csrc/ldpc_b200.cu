// ldpc_b200.cu — the C ABI (include/ldpc_b200.h): handle, kernel selection, stream-slot pipeline, boundary conversions.
//
// Replaces the reference's decoder objects: CGPUDecoder + CGPU_Decoder_{MS,OMS,NMS,2NMS}_SIMD::decode
// (ref: code/gpu_fixed/decoder_template/CGPUDecoder.cpp:14-60, code/gpu_fixed/decoder_oms/CGPU_Decoder_OMS_SIMD.cu:97-149)
// and the x86 CreateDecoder() products (ref: code/x86/CDecoder/DecoderLibrary.h:44-134).  Where the reference does
// blocking cudaMemcpy H2D -> Interleaver -> kernel -> InvInterleaver -> blocking D2H on the default stream, decode() here
// cuts the batch into chunks and runs H2D / decode / D2H of consecutive chunks on rotating stream slots so the three
// overlap (the pipeline ldpc_multiStream intended: code/ldpc_multiStream/queue/handler.cpp:51-138).
// There is NO CPU fallback: without a CUDA device every compute entry point returns LDPC_ERR_NO_DEVICE.
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ldpc_b200.h"
#include "boundary.cuh"
#include "channel.cuh"
#include "launch.cuh"
#include "kernel_gp.cuh"
#include "kernel_oc.cuh"
#include "kernel_wf.cuh"

using namespace ldpcb200;

namespace {

constexpr int kSlots = 4;
thread_local std::string g_create_error;

struct Slot {
    cudaStream_t stream = nullptr;
    int8_t* d_llr = nullptr;     size_t llr_bytes = 0;
    uint8_t* d_hard = nullptr;   size_t hard_bytes = 0;
    uint8_t* d_iters = nullptr;  size_t iters_bytes = 0;
    uint32_t* d_V = nullptr;     size_t v_bytes = 0;       // frame-parallel kernel state
    uint32_t* d_MSG = nullptr;   size_t msg_bytes = 0;
    uint8_t* d_LLR0 = nullptr;   size_t llr0_bytes = 0;    // generic engine, flooding: interleaved channel values
    int T = 0;                                              // words per variable the V/MSG buffers were laid out for
    // The slot's scratch (d_V, d_MSG, d_LLR0, d_iters) may be used from a caller's stream (decode_device) as well as from the slot's
    // own: `busy` is recorded behind every use and a use from ANOTHER stream waits on it first.
    cudaEvent_t busy = nullptr; cudaStream_t last = nullptr; bool used = false;
    unsigned int* d_queue = nullptr;                        // warp-per-frame engine (kernel 6): the launch's work-queue counter
};

}  // namespace

struct ldpc_b200_handle_s {
    int device = 0;
    ldpc_code_t code{};
    ldpc_params_t prm{};
    size_t max_frames = 0, chunk_frames = 0;
    int kernel = 0;                 // 1 = frame-parallel, 2 = row-parallel on-chip, 3 = generic engine (fp32 arithmetic)
    int elem = 1;                   // bytes per LLR / posterior / message element at the boundary (1, 2 or 4)
    GpMode gp_mode{};
    int32_t* d_cptr = nullptr; int32_t* d_cedge = nullptr;
    OcRow* d_oc_rows = nullptr; int32_t* d_oc_levels = nullptr; int oc_nlevels = 0, oc_F = 0, oc_threads = 0; size_t oc_smem = 0; bool oc_packed_syn = false;   // on-chip generic engine (kernel 5)
    // warp-per-frame on-chip generic engine (kernel 6)
    WfRun* d_wf_runs = nullptr; WfVRun* d_wf_vruns = nullptr; uint16_t* d_wf_idx = nullptr; uint16_t* d_wf_cm = nullptr; uint16_t* d_wf_var = nullptr; uint32_t* d_wf_edge_of = nullptr;
    int wf_nruns = 0, wf_nvruns = 0, wf_melems = 0, wf_cmelems = 0, wf_varelems = 0, wf_npad = 0, wf_warps = 0, wf_ctas_per_sm = 1; size_t wf_smem = 0;
    uint32_t wf_off_vruns = 0, wf_off_idx = 0, wf_off_cm = 0, wf_off_var = 0, wf_off_state = 0;
    int levels = 0, sms = 0;
    // row-parallel plan
    int rp_G = 0, rp_P = 0, rp_groups = 0, rp_slots = 0, rp_npad = 0, rp_melems = 0, rp_nsteps = 0, rp_nruns = 0, rp_pair_words = 0, rp_static = 0, rp_pair_fastest = 0; size_t rp_smem = 0;
    RpStep* d_steps = nullptr; RpRun* d_runs = nullptr; uint16_t* d_idx_t = nullptr; uint32_t* d_edge_of = nullptr;
    uint32_t* d_pos = nullptr;
    uint32_t* d_pos2 = nullptr; int fs_max_deg = 0;    // staged frame-parallel kernel: edge table with hazard flags
    int fs_nseg = 0, fs_seg_deg[FS_MAXSEG] = {}, fs_seg_rows[FS_MAXSEG] = {}, fs_seg_cls[FS_MAXSEG] = {}, fs_seg_stair[FS_MAXSEG] = {};   // consumer-side segments (kernel_fs.cuh: FsArgs::seg_*)
    int64_t fs_last_variant = 0;                       // LDPC_INFO_FS_VARIANT
    uint32_t* d_edge_row = nullptr;                    // ... and (row << 4 | slot) of every edge, for re-expanding compressed messages (debug_state)
    Slot slot[kSlots];
    bool debug = false;
    int8_t* d_dbg_post = nullptr; int8_t* d_dbg_msgs = nullptr; size_t dbg_post_bytes = 0, dbg_msgs_bytes = 0, dbg_frames = 0; int dbg_iters = 0;
    unsigned long long* d_counters = nullptr;
    float* d_qy = nullptr; int8_t* d_qq = nullptr; size_t qy_bytes = 0, qq_bytes = 0;     // scratch of ldpc_b200_quantize, grown on demand, kept
    int64_t launches = 0;
    std::string err;
};

namespace {

int fail(ldpc_handle h, int status, const std::string& msg)
{
    if (h) h->err = msg; else g_create_error = msg;
    return status;
}

#define CU_TRY(h, call)                                                                                         \
    do { cudaError_t e__ = (call);                                                                              \
         if (e__ != cudaSuccess) return fail(h, LDPC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } while (0)

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup: no link-time dependency on libcuda
typedef CUresult (*tm_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
tm_encode_fn tm_encoder()
{
    static tm_encode_fn fn = [] {
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { cudaGetLastError(); p = nullptr; }
        return (tm_encode_fn)p;
    }();
    return fn;
}
// 2-D map over a state array X[rows][T] of 32-bit words, box = nc words x box_rows rows
bool make_state_map(CUtensorMap* tm, uint32_t* base, size_t rows, int T, int nc, int box_rows)
{
    tm_encode_fn enc = tm_encoder();
    if (!enc) return false;
    const cuuint64_t dims[2] = { (cuuint64_t)T, (cuuint64_t)rows }, strides[1] = { (cuuint64_t)T * 4 };
    const cuuint32_t box[2] = { (cuuint32_t)nc, (cuuint32_t)box_rows }, estr[2] = { 1, 1 };
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int lo_rail(const ldpc_params_t& p) { return p.semantics == LDPC_SEM_GPU_FIXED ? -128 : -p.sat_var; }
int hi_rail(const ldpc_params_t& p) { return p.semantics == LDPC_SEM_ARM_SCALAR ? p.sat_var : 127; }

template <typename T>
int ensure(ldpc_handle h, T** p, size_t* have, size_t need)
{
    if (*have >= need) return LDPC_OK;
    if (*p) { CU_TRY(h, cudaFree(*p)); *p = nullptr; *have = 0; }
    CU_TRY(h, cudaMalloc((void**)p, need));
    *have = need;
    return LDPC_OK;
}

int validate_params(const ldpc_code_t* c, const ldpc_params_t* p, std::string& why)
{
    if (p->dtype != LDPC_DTYPE_I8 && p->dtype != LDPC_DTYPE_I16 && p->dtype != LDPC_DTYPE_F32) { why = "unknown dtype"; return LDPC_ERR_INVALID; }
    if (p->schedule != LDPC_SCHED_LAYERED && p->schedule != LDPC_SCHED_FLOODING) { why = "unknown schedule"; return LDPC_ERR_INVALID; }
    if (p->algo < LDPC_ALGO_MS || p->algo > LDPC_ALGO_2NMS) { why = "unknown algo"; return LDPC_ERR_INVALID; }
    if (p->early_term != LDPC_ET_NONE && p->early_term != LDPC_ET_SYNDROME) { why = "unknown early_term"; return LDPC_ERR_INVALID; }
    if (p->out_format != LDPC_OUT_BYTES && p->out_format != LDPC_OUT_PACKED) { why = "unknown out_format"; return LDPC_ERR_INVALID; }
    if (p->kernel < 0 || p->kernel > 6) { why = "unknown kernel id"; return LDPC_ERR_INVALID; }
    const bool generic = p->dtype != LDPC_DTYPE_I8 || p->schedule != LDPC_SCHED_LAYERED;
    if (generic && (p->kernel == 1 || p->kernel == 2 || p->kernel == 4)) { why = "kernels 1, 2 and 4 are int8 layered only: int16, float and flooding run on the generic engine (kernel 0 or 3)"; return LDPC_ERR_UNSUPPORTED; }
    for (int i = 0; i < c->nb_deg; i++) if (c->deg[i] > GP_MAXDEG && (generic || p->kernel == 3 || p->kernel == 5 || p->kernel == 6)) { why = "generic engine: row degree > 4096"; return LDPC_ERR_UNSUPPORTED; }
    if (p->dtype == LDPC_DTYPE_F32) {
        if (p->algo == LDPC_ALGO_OMS && (p->offset < 0 || p->llr_scale <= 0)) { why = "float OMS: offset >= 0 and llr_scale > 0 (offset is offset/llr_scale in channel units)"; return LDPC_ERR_INVALID; }
        if ((p->algo == LDPC_ALGO_NMS || p->algo == LDPC_ALGO_2NMS) && !(p->factor1 > 0.0f && p->factor2 > 0.0f)) { why = "float NMS: factors must be positive"; return LDPC_ERR_INVALID; }
        return LDPC_OK;
    }
    const bool wide = p->dtype == LDPC_DTYPE_I16;
    const int rail = wide ? 32767 : 127;
    switch (p->semantics) {
    case LDPC_SEM_X86_SSE:
        if (wide) { why = "int16: UNIFORM or ARM_SCALAR semantics (the SSE decoder is int8 only, CDecoder_OMS_fixed_SSE.cpp:114-120)"; return LDPC_ERR_UNSUPPORTED; }
        /* fall through */
    case LDPC_SEM_UNIFORM:
        if (p->algo != LDPC_ALGO_OMS && p->algo != LDPC_ALGO_NMS) { why = "x86 semantics: OMS or NMS only"; return LDPC_ERR_UNSUPPORTED; }
        if (!wide && p->sat_var != 127) { why = "x86 semantics: sat_var must be 127 (the reference exits otherwise, CDecoder_OMS_fixed_SSE.cpp:114-120)"; return LDPC_ERR_INVALID; }
        if (wide && (p->sat_var < 1 || p->sat_var > rail)) { why = "int16 path: 1 <= sat_var <= 32767"; return LDPC_ERR_INVALID; }
        break;
    case LDPC_SEM_ARM_SCALAR:
        if (p->algo != LDPC_ALGO_OMS) { why = "ARM scalar semantics: OMS only"; return LDPC_ERR_UNSUPPORTED; }
        if (p->sat_var < 1 || p->sat_var > rail) { why = wide ? "int16 path: 1 <= sat_var <= 32767" : "int8 path: 1 <= sat_var <= 127"; return LDPC_ERR_INVALID; }
        break;
    case LDPC_SEM_GPU_FIXED:
        if (wide) { why = "int16: UNIFORM or ARM_SCALAR semantics (the gpu_fixed kernels are int8x4 only)"; return LDPC_ERR_UNSUPPORTED; }
        break;
    default: why = "unknown semantics"; return LDPC_ERR_INVALID;
    }
    if (p->sat_msg < 1 || p->sat_msg > rail || p->offset < 0 || p->offset > rail) { why = "sat_msg/offset out of range"; return LDPC_ERR_INVALID; }
    if (p->factor_q5 < 0 || p->factor_q5 > 255) { why = "factor_q5 must be in [0,255]"; return LDPC_ERR_INVALID; }
    if (p->algo == LDPC_ALGO_NMS || p->algo == LDPC_ALGO_2NMS)
        for (int i = 0; i < c->nb_deg; i++) if (c->deg[i] < 2) { why = "NMS needs row degree >= 2"; return LDPC_ERR_UNSUPPORTED; }
    return LDPC_OK;
}

GpMode make_gp_mode(const ldpc_params_t& p)
{
    GpMode md{};
    md.is_float = p.dtype == LDPC_DTYPE_F32; md.wide = p.dtype == LDPC_DTYPE_I16;
    md.sem = p.semantics; md.algo = p.algo;
    const bool gpu = !md.is_float && p.semantics == LDPC_SEM_GPU_FIXED, arm = !md.is_float && p.semantics == LDPC_SEM_ARM_SCALAR;
    md.lo = gpu ? -128.0f : -(float)p.sat_var;
    md.hi = (arm || md.wide) ? (float)p.sat_var : 127.0f;
    md.sat_msg = gpu ? 31.0f : (float)p.sat_msg;
    md.off = md.is_float ? (float)p.offset / (float)p.llr_scale : (gpu ? 1.0f : (float)p.offset);
    if (md.is_float) {
        md.f1 = p.algo == LDPC_ALGO_MS ? 1.0f : p.factor1;
        md.f2 = p.algo == LDPC_ALGO_MS ? 1.0f : (p.algo == LDPC_ALGO_2NMS ? p.factor2 : p.factor1);
    } else { md.f1 = 0.75f; md.f2 = p.algo == LDPC_ALGO_2NMS ? 0.875f : 0.75f; }   // GPU_FIXED literals (CUDA_NMS_SIMD.cu:76-83, CUDA_2NMS_SIMD.cu:76-83)
    md.factor = (float)p.factor_q5; md.pack_sat = md.wide ? 32767.0f : 127.0f;
    md.min_init = md.is_float ? INFINITY : (arm ? (float)(p.sat_var + 1) : (gpu ? 127.0f : (float)p.sat_var));
    md.x86 = !md.is_float && (p.semantics == LDPC_SEM_X86_SSE || p.semantics == LDPC_SEM_UNIFORM);
    md.quirk = !md.is_float && p.semantics == LDPC_SEM_X86_SSE && p.algo == LDPC_ALGO_OMS;
    return md;
}

// Build the row-parallel plan: levels -> steps of <= 32 same-degree rows, step-transposed index table, (G, P) grouping.
struct RpPlan {
    std::vector<RpStep> steps; std::vector<RpRun> runs; std::vector<uint16_t> idx_t; std::vector<uint32_t> edge_of;
    int G = 1, P = 1, m_elems = 0, pair_pad = 0, slots = 0, static_nrows = 0, pair_fastest = 0; bool pair_pad_ok = false;
};

int build_rp_plan(ldpc_handle h, RpPlan& plan, size_t smem_budget)
{
    const ldpc_code_t& c = h->code;
    std::vector<int32_t> level(c.n_checks);
    const int levels = ldpc_b200_level_schedule(&c, level.data());
    if (levels < 0) return levels;
    h->levels = levels;
    std::vector<int> row_cls(c.n_checks); std::vector<uint32_t> row_e0(c.n_checks);
    { int r = 0; uint32_t e = 0;
      for (int k = 0; k < c.nb_deg; k++) for (int q = 0; q < c.rows[k]; q++, r++) { row_cls[r] = k; row_e0[r] = e; e += c.deg[k]; } }
    std::vector<std::vector<int>> by_level(levels);
    for (int r = 0; r < c.n_checks; r++) by_level[level[r]].push_back(r);
    struct Tmp { RpStep st; std::vector<int> rows; };
    std::vector<Tmp> tmp;
    int off = 0;
    // Step size.  Normally a (level, degree class) group of rows is cut into steps of <= 32 rows.  When every group of the code
    // is a multiple of one size S in (32, 128] (QC codes with a large circulant: 2304x1152 -> 96, 1944x972 -> 81, 1248x624 -> 52),
    // steps of S rows let ONE frame pair occupy 2-4 warps, which is where such codes get their parallelism from: only a handful of
    // their pairs fit in shared memory (2304x1152: 5 per SM = 5 warps with 32-row steps, 15 with 96-row steps).
    int big_step = 0;
    if (h->prm.reserved[3] != 3) {       // reserved[3] == 3: force the 32-row steps (A/B experiments)
        int g = 0;
        for (int L = 0; L < levels; L++)
            for (int k = 0; k < c.nb_deg; k++) {
                int cnt = 0;
                for (int r : by_level[L]) cnt += row_cls[r] == k;
                if (cnt) { int a = g, b = cnt; while (b) { const int t_ = a % b; a = b; b = t_; } g = a; }
            }
        int S = g;
        for (int d = 2; S > 128 && d <= g; d++) if (g % d == 0 && g / d <= 128) S = g / d;
        bool degs_ok = true;
        for (int k = 0; k < c.nb_deg; k++) degs_ok = degs_ok && c.deg[k] >= 6 && c.deg[k] <= 8;     // the strides 64/96/128 exist as specialised variants only
        if (S > 32 && S <= 128 && degs_ok) big_step = S;
    }
    for (int L = 0; L < levels; L++) {
        bool first = true;
        for (int k = 0; k < c.nb_deg; k++) {
            std::vector<int> rows;
            for (int r : by_level[L]) if (row_cls[r] == k) rows.push_back(r);
            if (rows.empty()) continue;
            const int rounds = big_step ? (int)rows.size() / big_step : ((int)rows.size() + 31) / 32;
            const int per = ((int)rows.size() + rounds - 1) / rounds;
            for (int q = 0; q < rounds; q++) {
                const int b = q * per, e = std::min((int)rows.size(), b + per);
                if (e <= b) continue;
                Tmp t; RpStep& st = t.st; memset(&st, 0, sizeof(st));
                st.deg = (uint16_t)c.deg[k]; st.cls = (uint8_t)k; st.nrows = (uint16_t)(e - b); st.sync = first ? 1 : 0;
                first = false;
                // specialised instantiations exist for degree 6..8 with an element stride of 24, 32, 64, 96 or 128 (immediate addressing)
                int stride = st.nrows;
                if (st.deg >= 6 && st.deg <= 8 && st.nrows > 16) stride = st.nrows <= 24 ? 24 : (st.nrows + 31) / 32 * 32;
                st.stride = (uint16_t)stride;
                st.variant = (uint32_t)rp_variant_id(st.deg, stride);
                st.msg_off = (uint32_t)off;
                st.magic = (65536u + st.nrows - 1) / st.nrows;
                t.rows.assign(rows.begin() + b, rows.begin() + e);
                off += stride * st.deg;
                tmp.push_back(t);
            }
        }
    }
    plan.m_elems = (off + 3) / 4 * 4;
    plan.idx_t.assign(plan.m_elems, 0); plan.edge_of.assign(plan.m_elems, 0xFFFFFFFFu);
    for (auto& t : tmp) {
        const RpStep& st = t.st;
        for (int z = 0; z < st.nrows; z++)
            for (int j = 0; j < st.deg; j++) {
                const uint32_t ref_e = row_e0[t.rows[z]] + j;
                plan.idx_t[st.msg_off + j * st.stride + z] = (uint16_t)(4 * c.pos[ref_e]);
                plan.edge_of[st.msg_off + j * st.stride + z] = ref_e;
            }
        plan.steps.push_back(st);
    }
    // every way into an iteration ends with a group barrier (load, previous iteration, syndrome bookkeeping): the first level needs none
    if (!plan.steps.empty()) plan.steps[0].sync = 0;
    // runs: consecutive steps that share one kernel instantiation (degree, stride, x86 quirk class)
    for (int i = 0; i < (int)plan.steps.size(); i++) {
        const RpStep& st = plan.steps[i];
        const int quirk = st.cls >= 1 ? 1 : 0;
        if (!plan.runs.empty() && plan.runs.back().variant == (int)st.variant && plan.runs.back().quirk == quirk && plan.runs.back().count < 32
            && st.msg_off == plan.runs.back().msg_off0 + (uint32_t)plan.runs.back().count * st.stride * st.deg) {
            plan.runs.back().syncmask |= (uint32_t)st.sync << plan.runs.back().count;
            plan.runs.back().count++;
        } else plan.runs.push_back(RpRun{ i, 1, (int)st.variant, quirk, st.msg_off, (uint32_t)st.sync, { 0u, 0u } });
    }
    // grouping: G warps share P pairs so that the P*nrows tasks of a step fill whole warps.  Score = lane efficiency x
    // occupancy (measured on 576x288, profiles/r01_sweep_groupings.jsonl: (3,4) with 18 warps 0.78 ms, (1,1) with 23 warps at
    // 75 % lanes 0.86 ms, (1,4) with 100 % lanes but 6 warps 1.01 ms).
    bool uniform = true;
    for (auto& st : plan.steps) uniform = uniform && st.nrows == plan.steps[0].nrows;
    const int nr0 = plan.steps[0].nrows;
    const int base_words = (c.n + 3) / 4 * 4 + plan.m_elems;
    auto pad_for = [&](int P) {
        // a warp that straddles two pairs must not hit the same banks in both: pair pitch = nrows (mod 32 words) puts task t
        // of a step in bank t % 32 whatever pair it belongs to
        if (P > 1 && uniform && nr0 % 4 == 0) return ((nr0 % 32) - base_words % 32 + 32) % 32;
        return 0;
    };
    auto slots_for = [&](int G, int P, int pad) {
        const size_t pair_bytes = (size_t)(base_words + pad) * 4;
        int slots = 0;
        for (int s_try = 1; s_try <= 256; s_try++) {
            const size_t fixed = plan.steps.size() * sizeof(RpStep) + plan.runs.size() * sizeof(RpRun) + (((size_t)plan.m_elems * 2 + 15) / 16) * 16 + (((size_t)s_try * 8 + 15) / 16) * 16;
            const int groups = (s_try + P - 1) / P;
            if (fixed + pair_bytes * s_try > smem_budget || groups * G > RP_MAX_THREADS / 32 || (G > 1 && groups > RP_MAX_GROUPS)) break;
            slots = s_try;
        }
        return slots;
    };
    double best = -1.0;
    for (int G = 1; G <= 4; G++)
        for (int P = 1; P <= 8; P++) {
            const int pad = pad_for(P), slots = slots_for(G, P, pad);
            if (slots < 1) continue;
            const int full = slots / P, rest = slots % P;
            double useful = 0, issued = 0;
            for (auto& st : plan.steps) {
                const int lanes = 32 * G;
                useful += (double)slots * st.nrows * st.deg;
                issued += (double)full * ((P * st.nrows + lanes - 1) / lanes) * lanes * st.deg;
                if (rest) issued += (double)((rest * st.nrows + lanes - 1) / lanes) * lanes * st.deg;
            }
            const int warps = (full + (rest ? 1 : 0)) * G;
            const double score = useful / issued * std::min(1.0, warps / 16.0) - 0.002 * G * P;
            if (score > best) { best = score; plan.G = G; plan.P = P; }
        }
    // experiment knob (not part of the reference's parameter set): reserved[0]/[1] force the grouping
    if (h->prm.reserved[0] > 0 && h->prm.reserved[0] <= 4 && h->prm.reserved[1] > 0 && h->prm.reserved[1] <= 8) { plan.G = h->prm.reserved[0]; plan.P = h->prm.reserved[1]; }
    plan.pair_pad = pad_for(plan.P);
    plan.pair_pad_ok = plan.P > 1 && uniform && nr0 % 4 == 0;       // pad_for() really produced pitch = nrows (mod 32)
    plan.slots = slots_for(plan.G, plan.P, plan.pair_pad);
    // experiment knob (tools/occupancy_sweep.py): fewer frame pairs per SM than shared memory allows, to measure what resident warps are
    // worth to this kernel — the question behind a compressed message format ("compressed messages")
    if (const char* cap = getenv("LDPC_B200_RP_MAX_SLOTS")) { const int c_ = atoi(cap); if (c_ >= 1) plan.slots = std::min(plan.slots, c_); }
    // static plan: uniform steps, every run specialised, at most one task per lane of a full group
    bool all_special = true;
    for (auto& r : plan.runs) all_special = all_special && r.variant != 0;
    plan.static_nrows = (uniform && all_special && plan.P * nr0 <= 32 * plan.G && h->prm.reserved[3] != 1) ? nr0 : 0;
    // lane -> (pair, row) with the pair index fastest when that makes a warp's four 8-row windows tile the 32 banks (pair pitch
    // = nrows (mod 32) by pad_for(): 24 -> offsets 0/24/16/8, 8 -> 0/8/16/24); profiles/r01_ncu_rp_v6.txt had 23 % extra wavefronts
    if (plan.static_nrows && plan.P == 4 && plan.pair_pad_ok && (nr0 % 32 == 24 || nr0 % 32 == 8) && 4 * nr0 == 32 * plan.G && h->prm.reserved[3] != 5)
        plan.pair_fastest = 1;
    if (plan.static_nrows) {   // the static kernel is compiled for at most RP_STATIC_THREADS threads
        const int max_groups = RP_STATIC_THREADS / (32 * plan.G);
        if (max_groups < 1) plan.static_nrows = 0;
        else plan.slots = std::min(plan.slots, max_groups * plan.P);
    }
    return LDPC_OK;
}

void destroy_impl(ldpc_handle h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    for (auto& s : h->slot) {
        if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
        if (s.busy) cudaEventDestroy(s.busy);
        cudaFree(s.d_queue);
        cudaFree(s.d_llr); cudaFree(s.d_hard); cudaFree(s.d_iters); cudaFree(s.d_V); cudaFree(s.d_MSG); cudaFree(s.d_LLR0);
    }
    cudaFree(h->d_steps); cudaFree(h->d_runs); cudaFree(h->d_idx_t); cudaFree(h->d_edge_of); cudaFree(h->d_pos); cudaFree(h->d_pos2); cudaFree(h->d_edge_row);
    cudaFree(h->d_qy); cudaFree(h->d_qq);
    cudaFree(h->d_wf_runs); cudaFree(h->d_wf_vruns); cudaFree(h->d_wf_idx); cudaFree(h->d_wf_cm); cudaFree(h->d_wf_var); cudaFree(h->d_wf_edge_of);
    cudaFree(h->d_dbg_post); cudaFree(h->d_dbg_msgs); cudaFree(h->d_counters); cudaFree(h->d_cptr); cudaFree(h->d_cedge); cudaFree(h->d_oc_rows); cudaFree(h->d_oc_levels);
    free(h->code.pos);
    delete h;
}

int ensure_debug(ldpc_handle h, size_t frames, int iters)
{
    int rc;
    if ((rc = ensure(h, &h->d_dbg_post, &h->dbg_post_bytes, frames * (size_t)h->code.n * h->elem))) return rc;
    if ((rc = ensure(h, &h->d_dbg_msgs, &h->dbg_msgs_bytes, frames * (size_t)h->code.m * h->elem))) return rc;
    h->dbg_frames = frames; h->dbg_iters = iters;
    return LDPC_OK;
}

size_t hard_row_bytes(ldpc_handle h) { return h->prm.out_format == LDPC_OUT_PACKED ? (size_t)(h->code.n + 7) / 8 : (size_t)h->code.n; }

// A slot's scratch is about to be used on stream st: order it behind the slot's previous use when that ran on another stream.
int slot_enter(ldpc_handle h, Slot& s, cudaStream_t st)
{
    if (s.used && s.last != st) CU_TRY(h, cudaStreamWaitEvent(st, s.busy, 0));
    return LDPC_OK;
}
int slot_leave(ldpc_handle h, Slot& s, cudaStream_t st)
{
    CU_TRY(h, cudaEventRecord(s.busy, st));
    s.used = true; s.last = st;
    return LDPC_OK;
}

// device buffers of one slot for `frames` frames through decode_async: allocated (or grown) BEFORE a pipeline starts, so that no
// cudaFree / cudaMalloc — both device-wide synchronisation points — lands between the asynchronous copies and kernels of other slots
int reserve_slot(ldpc_handle h, Slot& s, size_t frames, bool want_iters)
{
    const ldpc_code_t& c = h->code;
    const size_t el = (size_t)h->elem;
    int rc;
    if ((rc = ensure(h, &s.d_llr, &s.llr_bytes, frames * (size_t)c.n * el))) return rc;
    if ((rc = ensure(h, &s.d_hard, &s.hard_bytes, frames * hard_row_bytes(h)))) return rc;
    if (h->kernel == 3) {
        const size_t T = (frames + 31) / 32 * 32;
        if ((rc = ensure(h, &s.d_V, &s.v_bytes, (size_t)c.n * T * el))) return rc;
        if ((rc = ensure(h, &s.d_MSG, &s.msg_bytes, (size_t)c.m * T * el))) return rc;
        if (h->prm.schedule == LDPC_SCHED_FLOODING && (rc = ensure(h, &s.d_LLR0, &s.llr0_bytes, (size_t)c.n * T * el))) return rc;
        if (want_iters && (rc = ensure(h, &s.d_iters, &s.iters_bytes, T))) return rc;
    } else if (h->kernel == 1 || h->kernel == 4) {
        const size_t tq = h->kernel == 4 ? FS_MAX_CONSUMERS : 32, T = ((frames + 3) / 4 + tq - 1) / tq * tq;       // the widest row pitch launch_decode may pick
        if ((rc = ensure(h, &s.d_V, &s.v_bytes, (size_t)c.n * T * 4))) return rc;
        if ((rc = ensure(h, &s.d_MSG, &s.msg_bytes, (size_t)c.m * T * 4))) return rc;
        if (want_iters && (rc = ensure(h, &s.d_iters, &s.iters_bytes, 4 * T))) return rc;
    } else if (want_iters && (rc = ensure(h, &s.d_iters, &s.iters_bytes, frames))) return rc;
    return LDPC_OK;
}

// generic engine: interleave (+ clamp) -> decode -> hard decisions; S = storage type of the boundary and of the HBM state
template <class S>
int launch_decode_gp(ldpc_handle h, Slot& s, const void* d_llr, uint8_t* d_hard, size_t frames, int iters, uint8_t* d_iters, cudaStream_t st, bool want_debug)
{
    const ldpc_code_t& c = h->code;
    const int T = (int)((frames + 31) / 32 * 32);
    const bool flooding = h->prm.schedule == LDPC_SCHED_FLOODING;
    int rc;
    if ((rc = ensure(h, &s.d_V, &s.v_bytes, (size_t)c.n * T * sizeof(S)))) return rc;
    if ((rc = ensure(h, &s.d_MSG, &s.msg_bytes, (size_t)c.m * T * sizeof(S)))) return rc;
    if (flooding && (rc = ensure(h, &s.d_LLR0, &s.llr0_bytes, (size_t)c.n * T * sizeof(S)))) return rc;
    s.T = T;
    const dim3 blk(32, 8), gn((unsigned)(T / 32), (unsigned)((c.n + 31) / 32)), gm((unsigned)(T / 32), (unsigned)((c.m + 31) / 32));
    S* V = reinterpret_cast<S*>(s.d_V); S* MSG = reinterpret_cast<S*>(s.d_MSG); S* LLR0 = reinterpret_cast<S*>(s.d_LLR0);
    gp_interleave_kernel<S><<<gn, blk, 0, st>>>(reinterpret_cast<const S*>(d_llr), V, frames, c.n, T, h->gp_mode.is_float ? 0 : 1, h->gp_mode.lo, h->gp_mode.hi);
    CU_TRY(h, cudaGetLastError());
    if (flooding) CU_TRY(h, cudaMemcpyAsync(LLR0, V, (size_t)c.n * T * sizeof(S), cudaMemcpyDeviceToDevice, st));
    if (iters == 0 && want_debug) CU_TRY(h, cudaMemsetAsync(MSG, 0, (size_t)c.m * T * sizeof(S), st));
    GpArgs<S> a{};
    a.V = V; a.MSG = MSG; a.LLR = LLR0; a.pos = h->d_pos; a.cptr = h->d_cptr; a.cedge = h->d_cedge; a.iters_done = nullptr;
    a.T = T; a.n = c.n; a.m = c.m; a.nb_deg = c.nb_deg;
    for (int i = 0; i < LDPC_MAX_DEG_CLASSES; i++) { a.deg[i] = c.deg[i]; a.rows[i] = c.rows[i]; }
    a.iters = iters; a.flooding = flooding; a.et = h->prm.early_term == LDPC_ET_SYNDROME; a.md = h->gp_mode;
    if (d_iters) {
        if ((rc = ensure(h, &s.d_iters, &s.iters_bytes, (size_t)T))) return rc;
        a.iters_done = s.d_iters;
    }
    gp_decode_kernel<S><<<T / GP_BLOCK + (T % GP_BLOCK ? 1 : 0), GP_BLOCK, 0, st>>>(a);
    CU_TRY(h, cudaGetLastError());
    if (h->prm.out_format == LDPC_OUT_PACKED) gp_hard_kernel<S, true><<<gn, blk, 0, st>>>(V, d_hard, frames, c.n, T);
    else gp_hard_kernel<S, false><<<gn, blk, 0, st>>>(V, d_hard, frames, c.n, T);
    CU_TRY(h, cudaGetLastError());
    h->launches += 3;
    if (d_iters && d_iters != s.d_iters) CU_TRY(h, cudaMemcpyAsync(d_iters, s.d_iters, frames, cudaMemcpyDeviceToDevice, st));
    if (want_debug) {
        gp_deinterleave_kernel<S><<<gn, blk, 0, st>>>(V, reinterpret_cast<S*>(h->d_dbg_post), frames, c.n, T);
        gp_deinterleave_kernel<S><<<gm, blk, 0, st>>>(MSG, reinterpret_cast<S*>(h->d_dbg_msgs), frames, c.m, T);
        CU_TRY(h, cudaGetLastError());
        h->launches += 2;
    }
    return LDPC_OK;
}

// generic engine, on-chip state: frame-major in, hard decisions out, nothing else touches HBM
template <class S>
int launch_decode_oc(ldpc_handle h, const void* d_llr, uint8_t* d_hard, size_t frames, int iters, uint8_t* d_iters, cudaStream_t st, bool want_debug)
{
    const ldpc_code_t& c = h->code;
    OcArgs<S> a{};
    a.llr = reinterpret_cast<const S*>(d_llr); a.hard = d_hard; a.iters_done = d_iters;
    a.dbg_post = want_debug ? reinterpret_cast<S*>(h->d_dbg_post) : nullptr; a.dbg_msgs = want_debug ? reinterpret_cast<S*>(h->d_dbg_msgs) : nullptr;
    a.pos = h->d_pos; a.cptr = h->d_cptr; a.cedge = h->d_cedge; a.rows = h->d_oc_rows; a.level_ptr = h->d_oc_levels;
    a.frames = frames; a.n = c.n; a.m = c.m; a.n_checks = c.n_checks; a.nlevels = h->oc_nlevels; a.F = h->oc_F; a.iters = iters;
    a.flooding = h->prm.schedule == LDPC_SCHED_FLOODING; a.et = h->prm.early_term == LDPC_ET_SYNDROME;
    a.packed = h->prm.out_format == LDPC_OUT_PACKED; a.md = h->gp_mode; a.packed_syn = h->oc_packed_syn ? 1 : 0;
    const int blocks = (int)std::min<size_t>((size_t)h->sms, (frames + h->oc_F - 1) / h->oc_F);
    CU_TRY(h, cudaFuncSetAttribute(oc_decode_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->oc_smem));
    a.threads = h->oc_threads;
    oc_decode_kernel<S><<<blocks, h->oc_threads, h->oc_smem, st>>>(a);
    CU_TRY(h, cudaGetLastError());
    h->launches += 1;
    return LDPC_OK;
}

// ---- warp-per-frame on-chip generic engine (kernel 6): plan = steps of <= 32 same-degree rows (per level when layered), the edge
// table step-transposed, and for the flooding schedule the column table in steps of 32 consecutive variables --------------------------
struct WfPlan {
    std::vector<WfRun> runs; std::vector<uint16_t> idx_t; std::vector<uint32_t> edge_of; std::vector<WfVRun> vruns; std::vector<uint16_t> cm_t, var_t;
    int m_elems = 0, n_pad = 0, levels = 1; size_t lanes_used = 0, lanes_total = 0;
};

int build_wf_plan(const ldpc_code_t& c, bool flooding, WfPlan& plan)
{
    std::vector<int32_t> level((size_t)c.n_checks, 0);
    if (!flooding) { plan.levels = ldpc_b200_level_schedule(&c, level.data()); if (plan.levels < 0) return plan.levels; }
    std::vector<int> row_cls((size_t)c.n_checks); std::vector<uint32_t> row_e0((size_t)c.n_checks);
    { int r = 0; uint32_t e = 0;
      for (int k = 0; k < c.nb_deg; k++) for (int q = 0; q < c.rows[k]; q++, r++) { row_cls[r] = k; row_e0[r] = e; e += c.deg[k]; } }
    std::vector<std::vector<int>> by_level((size_t)plan.levels);
    for (int r = 0; r < c.n_checks; r++) by_level[level[r]].push_back(r);
    std::vector<uint32_t> elem_of_edge((size_t)c.m, 0u);
    uint32_t off = 0;
    // check-node side: a run = the rows of one (level, degree class) in reference order, cut into steps of 32
    for (int L = 0; L < plan.levels; L++) {
        bool first = true;
        for (int k = 0; k < c.nb_deg; k++) {
            std::vector<int> rows;
            for (int r : by_level[L]) if (row_cls[r] == k) rows.push_back(r);
            if (rows.empty()) continue;
            const uint32_t d = (uint32_t)c.deg[k], nsteps = (uint32_t)(rows.size() + 31) / 32;
            WfRun run{}; run.deg = (uint16_t)d; run.cls = (uint16_t)k; run.nsteps = (uint16_t)nsteps; run.last = (uint16_t)(rows.size() - 32 * (nsteps - 1));
            run.off = off; run.sync = (first && L > 0) ? 1u : 0u;
            first = false;
            plan.idx_t.resize(off + 32u * d * nsteps, 0); plan.edge_of.resize(off + 32u * d * nsteps, 0xFFFFFFFFu);
            for (size_t q = 0; q < rows.size(); q++)
                for (uint32_t j = 0; j < d; j++) {
                    const uint32_t ref_e = row_e0[rows[q]] + j, el = off + 32u * ((uint32_t)(q / 32) * d + j) + (uint32_t)(q % 32);
                    plan.idx_t[el] = (uint16_t)(4 * c.pos[ref_e]); plan.edge_of[el] = ref_e; elem_of_edge[ref_e] = el;
                }
            off += 32u * d * nsteps;
            plan.lanes_used += rows.size(); plan.lanes_total += 32u * nsteps;
            plan.runs.push_back(run);
        }
    }
    plan.m_elems = (int)off;
    plan.n_pad = (c.n + 3) / 4 * 4;
    // variable-node side (flooding): variables sorted by column degree (stable: index order inside a degree), a run = one degree
    if (flooding) {
        std::vector<std::vector<uint32_t>> col((size_t)c.n);
        for (int e = 0; e < c.m; e++) col[c.pos[e]].push_back((uint32_t)e);                 // ascending edge order inside a column
        size_t maxdv = 0;
        for (auto& v : col) maxdv = std::max(maxdv, v.size());
        uint32_t coff = 0, voff = 0;
        for (size_t dv = 0; dv <= maxdv; dv++) {
            std::vector<int> vars;
            for (int v = 0; v < c.n; v++) if (col[v].size() == dv) vars.push_back(v);
            if (vars.empty()) continue;
            const uint32_t nsteps = (uint32_t)(vars.size() + 31) / 32;
            WfVRun run{}; run.dv = (uint16_t)dv; run.nsteps = (uint16_t)nsteps; run.last = (uint16_t)(vars.size() - 32 * (nsteps - 1)); run.off = coff; run.voff = voff;
            plan.cm_t.resize(coff + 32u * (uint32_t)dv * nsteps, 0); plan.var_t.resize(voff + 32u * nsteps, 0);
            for (size_t q = 0; q < vars.size(); q++) {
                plan.var_t[voff + q] = (uint16_t)(4 * vars[q]);
                for (size_t k = 0; k < dv; k++)
                    plan.cm_t[coff + 32u * ((uint32_t)(q / 32) * (uint32_t)dv + (uint32_t)k) + (uint32_t)(q % 32)] = (uint16_t)(4 * elem_of_edge[col[vars[q]][k]]);
            }
            coff += 32u * (uint32_t)dv * nsteps; voff += 32u * nsteps;
            plan.vruns.push_back(run);
        }
    }
    return LDPC_OK;
}

template <class S>
int launch_decode_wf(ldpc_handle h, Slot& s, const void* d_llr, uint8_t* d_hard, size_t frames, int iters, uint8_t* d_iters, cudaStream_t st, bool want_debug)
{
    const ldpc_code_t& c = h->code;
    if (!s.d_queue) CU_TRY(h, cudaMalloc((void**)&s.d_queue, sizeof(unsigned int)));
    CU_TRY(h, cudaMemsetAsync(s.d_queue, 0, sizeof(unsigned int), st));
    WfArgs<S> a{};
    a.llr = reinterpret_cast<const S*>(d_llr); a.hard = d_hard; a.iters_done = d_iters;
    a.dbg_post = want_debug ? reinterpret_cast<S*>(h->d_dbg_post) : nullptr; a.dbg_msgs = want_debug ? reinterpret_cast<S*>(h->d_dbg_msgs) : nullptr;
    a.runs = h->d_wf_runs; a.idx_t = h->d_wf_idx; a.vruns = h->d_wf_vruns; a.cm_t = h->d_wf_cm; a.var_t = h->d_wf_var; a.edge_of = h->d_wf_edge_of; a.counter = s.d_queue;
    a.frames = frames; a.n = c.n; a.m = c.m; a.n_pad = h->wf_npad; a.m_elems = h->wf_melems; a.nruns = h->wf_nruns; a.nvruns = h->wf_nvruns; a.cm_elems = h->wf_cmelems; a.var_elems = h->wf_varelems;
    a.iters = iters; a.flooding = h->prm.schedule == LDPC_SCHED_FLOODING; a.et = h->prm.early_term == LDPC_ET_SYNDROME; a.packed = h->prm.out_format == LDPC_OUT_PACKED;
    a.off_vruns = h->wf_off_vruns; a.off_idx = h->wf_off_idx; a.off_cm = h->wf_off_cm; a.off_var = h->wf_off_var; a.off_state = h->wf_off_state; a.md = h->gp_mode;
    a.pair = h->wf_warps * h->wf_ctas_per_sm < 12;
    const size_t per_cta = (size_t)h->wf_warps;
    const int blocks = (int)std::min<size_t>((size_t)h->sms * h->wf_ctas_per_sm, (frames + per_cta - 1) / per_cta);
    CU_TRY(h, (cudaError_t)launch_wf(a, blocks, h->wf_warps * 32, h->wf_smem, st));
    h->launches += 1;
    return LDPC_OK;
}

// decode `frames` frames that are already in device memory, on stream st.  For the frame-parallel kernel the slot's V/MSG
// state is used, `frames` must fit it.
int launch_decode(ldpc_handle h, Slot& s, const int8_t* d_llr, uint8_t* d_hard, size_t frames, int iters, uint8_t* d_iters, cudaStream_t st, bool want_debug)
{
    const ldpc_code_t& c = h->code;
    if (h->kernel == 6) {
        if (h->elem == 4) return launch_decode_wf<float>(h, s, d_llr, d_hard, frames, iters, d_iters, st, want_debug);
        if (h->elem == 2) return launch_decode_wf<int16_t>(h, s, d_llr, d_hard, frames, iters, d_iters, st, want_debug);
        return launch_decode_wf<int8_t>(h, s, d_llr, d_hard, frames, iters, d_iters, st, want_debug);
    }
    if (h->kernel == 5) {
        if (h->elem == 4) return launch_decode_oc<float>(h, d_llr, d_hard, frames, iters, d_iters, st, want_debug);
        if (h->elem == 2) return launch_decode_oc<int16_t>(h, d_llr, d_hard, frames, iters, d_iters, st, want_debug);
        return launch_decode_oc<int8_t>(h, d_llr, d_hard, frames, iters, d_iters, st, want_debug);
    }
    if (h->kernel == 3) {
        if (h->elem == 4) return launch_decode_gp<float>(h, s, d_llr, d_hard, frames, iters, d_iters, st, want_debug);
        if (h->elem == 2) return launch_decode_gp<int16_t>(h, s, d_llr, d_hard, frames, iters, d_iters, st, want_debug);
        return launch_decode_gp<int8_t>(h, s, d_llr, d_hard, frames, iters, d_iters, st, want_debug);
    }
    const int et = h->prm.early_term == LDPC_ET_SYNDROME;
    const int lo = lo_rail(h->prm), hi = hi_rail(h->prm);
    if (h->kernel == 2) {
        RpArgs a{};
        a.llr = d_llr; a.hard = d_hard; a.iters_done = d_iters;
        a.dbg_post = want_debug ? h->d_dbg_post : nullptr; a.dbg_msgs = want_debug ? h->d_dbg_msgs : nullptr;
        a.idx_t = h->d_idx_t; a.edge_of = h->d_edge_of; a.steps = h->d_steps;
        a.runs = h->d_runs; a.nruns = h->rp_nruns; a.pair_words = h->rp_pair_words;
        a.frames = frames; a.n = c.n; a.m = c.m; a.nsteps = h->rp_nsteps; a.n_pad = h->rp_npad; a.m_elems = h->rp_melems;
        a.static_nrows = h->rp_static; a.pair_fastest = h->rp_pair_fastest;
        a.G = h->rp_G; a.P = h->rp_P; a.groups = h->rp_groups; a.slots = h->rp_slots; a.iters = iters;
        a.packed = h->prm.out_format == LDPC_OUT_PACKED; a.prm = h->prm;
        const size_t pairs = (frames + 1) / 2;
        const int blocks = (int)std::min<size_t>((size_t)h->sms, (pairs + h->rp_slots - 1) / h->rp_slots);
        rp_launch_fn fn = h->prm.semantics == LDPC_SEM_X86_SSE ? launch_rp_x86 : h->prm.semantics == LDPC_SEM_UNIFORM ? launch_rp_uniform
                        : h->prm.semantics == LDPC_SEM_ARM_SCALAR ? launch_rp_arm : launch_rp_gpu;
        CU_TRY(h, (cudaError_t)fn(h->prm.algo, et, a, blocks, h->rp_groups * h->rp_G * 32, h->rp_smem, st));
        h->launches += 1;
        return LDPC_OK;
    }
    // frame-parallel: interleave -> decode -> de-interleave + hard decision
    const size_t t4 = (frames + 3) / 4;
    // consumer threads per CTA of the staged kernel.  The kernel's registers (96 for rows up to degree 8) keep 2 CTAs of 256 + 64
    // threads on an SM (16 consumer warps) or, capped at 80 registers with a few spills, 4 CTAs of 128 + 64 (16 consumer warps, twice
    // the producers).  A/B on one box, DVB-S2: 256 Ki frames 464 vs 493 ms, 128 Ki 276 vs 260 ms, 64 Ki 274 vs 209 ms — the wide
    // CTA wins once there are two of them for every SM.  reserved[4] bits 8.. force 128 (1) or 256 (2).
    const size_t fs_words = (t4 + 255) / 256 * 256;
    const int fs_knob = (h->prm.reserved[4] >> 8) & 15;
    const int fs_nc = h->fs_max_deg > 8 ? FS_CONSUMERS : fs_knob == 2 ? 256 : fs_knob == 1 ? FS_CONSUMERS
                    : (fs_words / 256 >= (size_t)(2 * h->sms * 17 / 20) ? 256 : FS_CONSUMERS);
    const int tq = h->kernel != 4 ? 32 : fs_nc;
    const int T = (int)((t4 + tq - 1) / tq * tq);
    int rc;
    if ((rc = ensure(h, &s.d_V, &s.v_bytes, (size_t)c.n * T * 4))) return rc;
    // staged kernel, compressed messages (kernel_fp.cuh: fp_row_math_c): four words per row and thread instead of one per edge.
    // reserved[4] bits 16..17: 1 = never, 2 = always (rows of degree <= 8 only); default: see DESIGN.md 3.2b
    const int cmp_knob = (h->prm.reserved[4] >> 16) & 3;
    if (cmp_knob == 2 && (h->kernel != 4 || h->fs_max_deg > 8)) return fail(h, LDPC_ERR_UNSUPPORTED, "compressed messages: staged kernel with row degrees <= 8 only");
    const bool fs_cmp = h->kernel == 4 && h->fs_max_deg <= 8 && cmp_knob == 2;
    const size_t msg_lines_total = fs_cmp ? (size_t)4 * c.n_checks : (size_t)c.m;
    if ((rc = ensure(h, &s.d_MSG, &s.msg_bytes, msg_lines_total * T * 4))) return rc;
    s.T = T;
    dim3 tg((unsigned)((frames + 127) / 128), (unsigned)((c.n + 127) / 128));
    // the way in covers every word of the row pitch T (the staged kernel rounds T up to its CTA width and runs all T threads): padding
    // frames get zero LLRs — an all-zero, syndrome-passing frame — instead of whatever the allocation held
    dim3 tgi((unsigned)((T + 31) / 32), tg.y);
    interleave_kernel<<<tgi, 256, 0, st>>>(d_llr, s.d_V, frames, c.n, T, lo, hi);
    CU_TRY(h, cudaGetLastError());
    FpArgs a{};
    a.V = s.d_V; a.MSG = s.d_MSG; a.pos = h->d_pos; a.iters_done = nullptr; a.T = T; a.n = c.n; a.m = c.m; a.nb_deg = c.nb_deg;
    for (int i = 0; i < LDPC_MAX_DEG_CLASSES; i++) { a.deg[i] = c.deg[i]; a.rows[i] = c.rows[i]; }
    a.iters = iters; a.exp_word = 0x64646464u; a.prm = h->prm;
    uint8_t* d_it4 = nullptr;
    if (d_iters) {   // kernel writes 4*T entries (padding frames included) into the slot's scratch, then the valid part is copied
        if ((rc = ensure(h, &s.d_iters, &s.iters_bytes, (size_t)4 * T))) return rc;
        d_it4 = s.d_iters; a.iters_done = d_it4;
    }
    if (h->kernel == 4) {
        FsArgs f{};
        f.V = s.d_V; f.MSG = s.d_MSG; f.pos2 = h->d_pos2; f.T = T; f.n = c.n; f.m = c.m; f.nb_deg = c.nb_deg;
        for (int i = 0; i < LDPC_MAX_DEG_CLASSES; i++) { f.deg[i] = c.deg[i]; f.rows[i] = c.rows[i]; }
        f.iters = iters; f.max_deg = h->fs_max_deg; f.exp_word = 0x64646464u; f.prm = h->prm;
        // consumers per CTA: 128.  Wider CTAs (256 / 512 consumers = 1 KB / 2 KB lines, a quarter of the bulk-copy requests) were built to
        // test whether the copy engine's request rate bounds the kernel: it does not (DVB-S2, 256 Ki frames: 570 / 576 / 583 ms for
        // 128 / 256 / 512) — the consumers' issue slots do (profiles/r01_ncu_fs_v2.txt: 66 % issue-active, 636 warp instructions per row).
        const int nc = fs_nc;
        f.nc = nc;
        // ring depth: as deep as shared memory allows for the CTAs that will share an SM, at most the hazard window
        const int ctas = T / nc;
        int per_sm = std::min(nc == 128 ? 4 : 2, std::max(1, (ctas + h->sms - 1) / h->sms));
        // message lines as one 2-D tensor copy per row (TMA tensor map) instead of D one-dimensional bulk copies, posterior lines four
        // at a time (tile::gather4): the copy engine serves requests one after the other (~46 cycles each), which is what bounds a
        // batch too small to put 16 consumer warps on an SM.  reserved[4] bits 12..13 (message map) and 14..15 (gather4): 1 = never,
        // 2 = always; default: both whenever the encoder is available.  A/B on one box (profiles/r02_sweep_fs_g4.jsonl), DVB-S2, bulk ->
        // message map -> + gather4: 16 Ki frames 243 -> 225 -> 195 ms, 64 Ki 246 -> 230 -> 201 ms, 128 Ki 306 -> 291 -> 266 ms, 303 104
        // frames 521 -> 514 -> 514 ms (gather4 fetches the hazard lines too; the balanced batch does not notice).
        const int tm_knob = (h->prm.reserved[4] >> 12) & 3, g4_knob = (h->prm.reserved[4] >> 14) & 3;
        f.use_tm = f.use_g4 = 0;
        if (tm_knob != 1 && iters > 0) {
            bool ok = true;
            if (fs_cmp) ok = make_state_map(&f.tm_msg[0], s.d_MSG, msg_lines_total, T, nc, 4);
            else for (int i = 0; i < c.nb_deg && ok; i++) ok = make_state_map(&f.tm_msg[i], s.d_MSG, (size_t)c.m, T, nc, c.deg[i]);
            if (!ok && tm_knob == 2) return fail(h, LDPC_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available or refused the message map");
            f.use_tm = ok ? 1 : 0;
        }
        if (iters > 0 && g4_knob != 1) {
            const bool ok = make_state_map(&f.tm_v, s.d_V, (size_t)c.n, T, nc, 1);
            if (!ok && g4_knob == 2) return fail(h, LDPC_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available or refused the posterior map");
            f.use_g4 = ok ? 1 : 0;
        }
        f.msg_line0 = f.use_g4 ? (f.max_deg + 3) / 4 * 4 : f.max_deg;
        f.nseg = h->fs_nseg;
        for (int i = 0; i < h->fs_nseg; i++) { f.seg_deg[i] = h->fs_seg_deg[i]; f.seg_rows[i] = h->fs_seg_rows[i]; f.seg_cls[i] = h->fs_seg_cls[i]; f.seg_stair[i] = h->fs_seg_stair[i]; }
        f.cmp = fs_cmp ? 1 : 0; f.msg_lines = fs_cmp ? 4 : f.max_deg;
        const size_t line = (size_t)nc * 4, stage_bytes = (size_t)(f.msg_line0 + f.msg_lines) * line + FS_P2_BYTES, fwd_bytes = (size_t)FS_FWD * f.max_deg * line;
        auto stages_for = [&](int ps) { return (int)(((long)(220 * 1024) / ps - (long)fwd_bytes - 256) / (long)stage_bytes); };
        while (per_sm > 1 && stages_for(per_sm) < 4) per_sm--;          // wide rows make wide stages: fewer CTAs per SM rather than a ring too shallow to hide anything
        int stages = stages_for(per_sm);
        stages = std::max(2, std::min(stages, FS_HAZARD - 1));   // a stage is handed back one row late (fs_row)
        if ((h->prm.reserved[4] & 255) >= 2 && (h->prm.reserved[4] & 255) < FS_HAZARD) stages = h->prm.reserved[4] & 255;     // experiment knob
        f.stages = stages;
        {   // paired staircase rows: only where at most two CTAs share an SM (124-146 registers), reserved[4] bits 19..20: 1 = never, 2 = always
            const int p2_knob = (h->prm.reserved[4] >> 19) & 3;
            bool any_stair = false;
            for (int i = 0; i < h->fs_nseg; i++) any_stair = any_stair || h->fs_seg_stair[i];
            f.pipe2 = (any_stair && !fs_cmp && nc == 128 && f.max_deg <= 8 && p2_knob != 1 && (p2_knob == 2 || ctas <= 2 * h->sms) && stages >= 4) ? 1 : 0;     // the pair holds three stages

        }
        const size_t smem = (size_t)((16 * stages + 127) / 128 * 128) + fwd_bytes + stages * stage_bytes;
        fs_launch_fn fn = h->prm.semantics == LDPC_SEM_X86_SSE ? (fs_cmp ? launch_fc_x86 : launch_fs_x86) : h->prm.semantics == LDPC_SEM_UNIFORM ? (fs_cmp ? launch_fc_uniform : launch_fs_uniform)
                        : h->prm.semantics == LDPC_SEM_ARM_SCALAR ? (fs_cmp ? launch_fc_arm : launch_fs_arm) : (fs_cmp ? launch_fc_gpu : launch_fs_gpu);
        f.et = et; f.iters_done = d_it4;
        h->fs_last_variant = (f.pipe2 ? 2 : 1) | (f.max_deg > FS_MAXDEG ? 16 : 0) | (fs_cmp ? 32 : 0) | 256 * (int64_t)nc;
        if (iters > 0) CU_TRY(h, (cudaError_t)fn(h->prm.algo, f, ctas, smem, st));
        else if (d_it4) CU_TRY(h, cudaMemsetAsync(d_it4, 0, (size_t)4 * T, st));
    } else {
    fp_launch_fn fn = h->prm.semantics == LDPC_SEM_X86_SSE ? launch_fp_x86 : h->prm.semantics == LDPC_SEM_UNIFORM ? launch_fp_uniform
                    : h->prm.semantics == LDPC_SEM_ARM_SCALAR ? launch_fp_arm : launch_fp_gpu;
    CU_TRY(h, (cudaError_t)fn(h->prm.algo, et, a, (T + FP_BLOCK - 1) / FP_BLOCK, st));
    }
    if (h->prm.out_format == LDPC_OUT_PACKED) deinterleave_hard_kernel<true><<<tg, 256, 0, st>>>(s.d_V, d_hard, frames, c.n, T, lo);
    else deinterleave_hard_kernel<false><<<tg, 256, 0, st>>>(s.d_V, d_hard, frames, c.n, T, lo);
    CU_TRY(h, cudaGetLastError());
    h->launches += 3;
    if (d_iters && d_iters != d_it4) CU_TRY(h, cudaMemcpyAsync(d_iters, d_it4, frames, cudaMemcpyDeviceToDevice, st));
    if (want_debug) {
        if (fs_cmp) fc_debug_state_kernel<<<1024, 256, 0, st>>>(s.d_V, s.d_MSG, h->d_edge_row, h->d_dbg_post, h->d_dbg_msgs, frames, c.n, c.m, T, lo, iters > 0);
        else fp_debug_state_kernel<<<1024, 256, 0, st>>>(s.d_V, s.d_MSG, h->d_dbg_post, h->d_dbg_msgs, frames, c.n, c.m, T, lo, iters > 0);
        CU_TRY(h, cudaGetLastError());
        h->launches += 1;
    }
    return LDPC_OK;
}

}  // namespace

extern "C" {

int ldpc_b200_abi_version(void) { return LDPC_B200_ABI_VERSION; }

int ldpc_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* ldpc_b200_status_string(int s)
{
    switch (s) {
    case LDPC_OK: return "ok";
    case LDPC_ERR_INVALID: return "invalid argument";
    case LDPC_ERR_CUDA: return "CUDA error";
    case LDPC_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU path)";
    case LDPC_ERR_IO: return "code table unreadable or malformed";
    case LDPC_ERR_NOMEM: return "out of memory";
    case LDPC_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
    }
}

void ldpc_b200_default_params(ldpc_params_t* p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->algo = LDPC_ALGO_OMS; p->schedule = LDPC_SCHED_LAYERED; p->dtype = LDPC_DTYPE_I8; p->semantics = LDPC_SEM_X86_SSE;
    p->offset = 1; p->factor_q5 = 29; p->factor1 = 0.75f; p->factor2 = 0.875f;       // (ref: code/x86/main_p.cpp:133-139)
    p->sat_var = 127; p->sat_msg = 31; p->llr_scale = 8; p->sat_llr = 31;              // (ref: code/x86/main_p.cpp:90-104)
    p->early_term = LDPC_ET_NONE; p->out_format = LDPC_OUT_BYTES; p->kernel = 0;
}

const char* ldpc_b200_last_error(ldpc_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ldpc_b200_create(ldpc_handle* out, const ldpc_code_t* code, const ldpc_params_t* params, int device, size_t max_frames)
{
    if (!out || !code || !params) return fail(nullptr, LDPC_ERR_INVALID, "null argument");
    *out = nullptr;
    int rc = ldpc_b200_check_code(code);
    if (rc) return fail(nullptr, rc, "malformed code table");
    std::string why;
    if ((rc = validate_params(code, params, why))) return fail(nullptr, rc, why);
    int ndev = ldpc_b200_device_count();
    if (ndev <= 0) return fail(nullptr, LDPC_ERR_NO_DEVICE, "no CUDA device visible: the decoder has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, LDPC_ERR_INVALID, "device index out of range");
    ldpc_handle h = new ldpc_b200_handle_s();
    h->device = device; h->prm = *params; h->code = *code;
    h->code.pos = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)code->m);
    if (!h->code.pos) { delete h; return fail(nullptr, LDPC_ERR_NOMEM, "host allocation failed"); }
    memcpy(h->code.pos, code->pos, sizeof(uint32_t) * (size_t)code->m);
    h->max_frames = max_frames ? max_frames : 65536;
#define CREATE_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { std::string m__ = std::string(#call) + ": " + cudaGetErrorString(e__); destroy_impl(h); return fail(nullptr, LDPC_ERR_CUDA, m__); } } while (0)
    CREATE_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CREATE_TRY(cudaGetDeviceProperties(&prop, device));
    h->sms = prop.multiProcessorCount;
    for (auto& s : h->slot) { CREATE_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)); CREATE_TRY(cudaEventCreateWithFlags(&s.busy, cudaEventDisableTiming)); }
    CREATE_TRY(cudaMalloc((void**)&h->d_pos, sizeof(uint32_t) * (size_t)code->m));
    CREATE_TRY(cudaMemcpy(h->d_pos, code->pos, sizeof(uint32_t) * (size_t)code->m, cudaMemcpyHostToDevice));
    CREATE_TRY(cudaMalloc((void**)&h->d_counters, 2 * sizeof(unsigned long long)));

    // kernel selection.  int16 / float / flooding -> the generic engine; int8 layered -> the on-chip row-parallel kernel when the
    // whole state of >= 8 frame pairs per SM fits in shared memory, else the frame-parallel kernel
    const bool generic = params->dtype != LDPC_DTYPE_I8 || params->schedule != LDPC_SCHED_LAYERED || params->kernel == 3 || params->kernel == 5 || params->kernel == 6;
    h->elem = params->dtype == LDPC_DTYPE_F32 ? 4 : (params->dtype == LDPC_DTYPE_I16 ? 2 : 1);
    h->kernel = generic ? 3 : 1;
    if (generic) {
        h->gp_mode = make_gp_mode(*params);
        std::vector<int32_t> cptr((size_t)code->n + 1, 0), cedge((size_t)code->m);
        for (int e = 0; e < code->m; e++) cptr[code->pos[e] + 1]++;
        for (int i = 0; i < code->n; i++) cptr[i + 1] += cptr[i];
        { std::vector<int32_t> fill(cptr.begin(), cptr.end() - 1);
          for (int e = 0; e < code->m; e++) cedge[fill[code->pos[e]]++] = e; }              // ascending edge order inside a column
        CREATE_TRY(cudaMalloc((void**)&h->d_cptr, cptr.size() * sizeof(int32_t)));
        CREATE_TRY(cudaMemcpy(h->d_cptr, cptr.data(), cptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        CREATE_TRY(cudaMalloc((void**)&h->d_cedge, cedge.size() * sizeof(int32_t)));
        CREATE_TRY(cudaMemcpy(h->d_cedge, cedge.data(), cedge.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        const bool flooding = params->schedule == LDPC_SCHED_FLOODING;
        // warp-per-frame on-chip variant (kernel 6): one warp owns one frame; chosen when at least 8 frames fit per SM and the schedule
        // fills the lanes (flooding always does; a layered schedule with narrow levels is better served by kernel 5's (row, frame) tasks)
        if ((params->kernel == 6 || params->kernel == 0) && code->n <= 16383) {
            WfPlan plan;
            if ((rc = build_wf_plan(*code, flooding, plan))) { destroy_impl(h); return fail(nullptr, rc, "level schedule failed"); }
            auto up16 = [](size_t x) { return (x + 15) / 16 * 16; };
            const size_t off_vruns = up16(plan.runs.size() * sizeof(WfRun)), off_idx = up16(off_vruns + plan.vruns.size() * sizeof(WfVRun));
            const size_t off_cm = up16(off_idx + (size_t)plan.m_elems * 2), off_var = up16(off_cm + plan.cm_t.size() * 2), off_state = up16(off_var + plan.var_t.size() * 2);
            const size_t frame_bytes = ((size_t)plan.n_pad * (flooding ? 2 : 1) + plan.m_elems) * sizeof(float);
            const size_t budget = (size_t)prop.sharedMemPerBlockOptin;
            const int fit = off_state < budget ? (int)((budget - off_state) / frame_bytes) : 0;
            const bool lanes_ok = plan.lanes_used * 2 >= plan.lanes_total;
            const bool table_ok = (size_t)plan.m_elems * 4 < 65535;
            if (params->kernel == 6 && (fit < 1 || !table_ok)) { destroy_impl(h); return fail(nullptr, LDPC_ERR_UNSUPPORTED, "code state does not fit in shared memory for the warp-per-frame engine"); }
            if (table_ok && (params->kernel == 6 || (fit >= 8 && lanes_ok))) {
                // several CTAs per SM when the frames are small: at most WF_MAX_WARPS warps per CTA, all of an SM's shared memory used
                int ctas = 1, warps = std::min(fit, WF_MAX_WARPS);
                if (fit > WF_MAX_WARPS) {
                    const size_t sm_total = (size_t)prop.sharedMemPerMultiprocessor;
                    ctas = (int)std::min<size_t>(64 / WF_MAX_WARPS, sm_total / (off_state + (size_t)WF_MAX_WARPS * frame_bytes + 1024));
                    ctas = std::max(ctas, 1);
                }
                h->kernel = 6; h->levels = plan.levels; h->wf_warps = warps; h->wf_ctas_per_sm = ctas;
                h->wf_nruns = (int)plan.runs.size(); h->wf_nvruns = (int)plan.vruns.size(); h->wf_melems = plan.m_elems; h->wf_cmelems = (int)plan.cm_t.size();
                h->wf_varelems = (int)plan.var_t.size(); h->wf_npad = plan.n_pad;
                h->wf_off_vruns = (uint32_t)off_vruns; h->wf_off_idx = (uint32_t)off_idx; h->wf_off_cm = (uint32_t)off_cm; h->wf_off_var = (uint32_t)off_var; h->wf_off_state = (uint32_t)off_state;
                h->wf_smem = off_state + (size_t)warps * frame_bytes;
                auto up = [&](auto** dptr, const void* src, size_t bytes) { if (cudaMalloc((void**)dptr, std::max<size_t>(bytes, 16)) != cudaSuccess) return false; return bytes == 0 || cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess; };
                if (!up(&h->d_wf_runs, plan.runs.data(), plan.runs.size() * sizeof(WfRun)) || !up(&h->d_wf_vruns, plan.vruns.data(), plan.vruns.size() * sizeof(WfVRun)) ||
                    !up(&h->d_wf_idx, plan.idx_t.data(), plan.idx_t.size() * 2) || !up(&h->d_wf_cm, plan.cm_t.data(), plan.cm_t.size() * 2) ||
                    !up(&h->d_wf_var, plan.var_t.data(), plan.var_t.size() * 2) ||
                    !up(&h->d_wf_edge_of, plan.edge_of.data(), plan.edge_of.size() * 4)) { destroy_impl(h); return fail(nullptr, LDPC_ERR_CUDA, "uploading the warp-per-frame plan failed"); }
            }
        }
        // on-chip variant (kernel 5): fp32 state of F frames in shared memory, rows in level order (flooding: one level)
        const size_t per_frame = (size_t)(code->n + code->m + (flooding ? code->n : 0)) * sizeof(float);
        // stop criterion on packed hard-decision words (kernel_oc.cuh: oc_pack_bits) where the criterion reads posteriors (float or
        // flooding): n extra words behind the state, frames-per-CTA <= 32
        const bool want_syn_words = params->early_term != LDPC_ET_NONE && (params->dtype == LDPC_DTYPE_F32 || flooding);
        const size_t syn_bytes = want_syn_words ? (size_t)code->n * sizeof(uint32_t) : 0;
        int F = (int)std::min<size_t>(OC_MAXF, ((size_t)prop.sharedMemPerBlockOptin - 1024 - syn_bytes) / per_frame);
        if (want_syn_words && F > 32) F = (int)std::min<size_t>(OC_MAXF, ((size_t)prop.sharedMemPerBlockOptin - 1024) / per_frame);   // short codes: the per-frame test
        h->oc_packed_syn = want_syn_words && F <= 32;
        if (params->kernel == 5 && F < 1) { destroy_impl(h); return fail(nullptr, LDPC_ERR_UNSUPPORTED, "code state does not fit in shared memory for the on-chip generic engine"); }
        if (h->kernel != 6 && ((params->kernel == 5 && F >= 1) || (params->kernel == 0 && F >= 8))) {
            std::vector<int32_t> level(code->n_checks, 0);
            int levels = 1;
            if (!flooding) { levels = ldpc_b200_level_schedule(code, level.data()); if (levels < 0) { destroy_impl(h); return fail(nullptr, levels, "level schedule failed"); } }
            h->levels = levels;
            std::vector<OcRow> rows_ref((size_t)code->n_checks), rows_sorted;
            { int r = 0; uint32_t e = 0;
              for (int k = 0; k < code->nb_deg; k++) for (int q = 0; q < code->rows[k]; q++, r++) { rows_ref[r] = OcRow{ e, (uint16_t)code->deg[k], (uint16_t)k }; e += code->deg[k]; } }
            std::vector<int32_t> level_ptr((size_t)levels + 1, 0);
            for (int r = 0; r < code->n_checks; r++) level_ptr[level[r] + 1]++;
            for (int L = 0; L < levels; L++) level_ptr[L + 1] += level_ptr[L];
            rows_sorted.resize(code->n_checks);
            { std::vector<int32_t> fill(level_ptr.begin(), level_ptr.end() - 1);
              for (int r = 0; r < code->n_checks; r++) rows_sorted[fill[level[r]]++] = rows_ref[r]; }     // stable: reference order inside a level, same-degree rows adjacent
            CREATE_TRY(cudaMalloc((void**)&h->d_oc_rows, rows_sorted.size() * sizeof(OcRow)));
            CREATE_TRY(cudaMemcpy(h->d_oc_rows, rows_sorted.data(), rows_sorted.size() * sizeof(OcRow), cudaMemcpyHostToDevice));
            CREATE_TRY(cudaMalloc((void**)&h->d_oc_levels, level_ptr.size() * sizeof(int32_t)));
            CREATE_TRY(cudaMemcpy(h->d_oc_levels, level_ptr.data(), level_ptr.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            h->kernel = 5; h->oc_nlevels = levels; h->oc_F = F; h->oc_smem = per_frame * F + (h->oc_packed_syn ? syn_bytes : 0);
            // CTA size.  A thread owns (frame, slot): a level of nr rows takes ceil(nr / R) rounds, R = T / F slots, and a round is
            // latency-bound (one task is a ~600-cycle dependency chain, the CTA's warps overlap theirs).  So: the fewest rounds per
            // iteration, then the fewest threads that achieve them.  (Two earlier models — fewest idle lanes, and a latency/issue sum
            // that is flat in T — picked 128 and 320 threads and were 3x and 1.5x slower than the widest CTA.)
            long best_rounds = -1;
            for (int T = 128; T <= OC_MAX_THREADS; T += 32) {
                const long R = T / F;
                if (R < 1) continue;
                long rounds = 0;
                for (int L = 0; L < levels; L++) rounds += ((long)(level_ptr[L + 1] - level_ptr[L]) + R - 1) / R;
                if (flooding) rounds += ((long)code->n + R - 1) / R;
                if (best_rounds < 0 || rounds < best_rounds) { best_rounds = rounds; h->oc_threads = T; }
            }
        }
    }
    if (!generic && code->n <= 16383 && params->kernel != 1 && params->kernel != 4) {
        RpPlan plan;
        if ((rc = build_rp_plan(h, plan, (size_t)prop.sharedMemPerBlockOptin))) { destroy_impl(h); return fail(nullptr, rc, "level schedule failed"); }
        h->rp_npad = (code->n + 3) / 4 * 4;
        h->rp_melems = plan.m_elems;
        h->rp_pair_words = h->rp_npad + plan.m_elems + plan.pair_pad;
        const size_t pair_bytes = (size_t)h->rp_pair_words * 4;
        auto fixed_bytes = [&](int slots_) { return plan.steps.size() * sizeof(RpStep) + plan.runs.size() * sizeof(RpRun) + (((size_t)plan.m_elems * 2 + 15) / 16) * 16 + (((size_t)slots_ * 8 + 15) / 16) * 16; };
        const int slots = plan.slots;
        // On-chip or frame-parallel?  The on-chip kernel lives on (pairs per SM) x (rows per level) concurrent row tasks: 576x288
        // 23 x 29, 2304x1152 5 x 115, 200x100 71 x 6 all run 25-100 M frames/s; 1200x600 (464 levels: 8 x 1.3) or 816x408 (11 x 6.4)
        // leave most lanes idle and the frame-parallel kernels win by 10-25x once the batch gives them threads
        // (profiles/r01_sweep_v5.jsonl, r01_paper_codes.jsonl).  Small batches starve the frame-parallel kernels instead
        // (4000x2000, 16 Ki frames: 2.3 M frames/s on chip, 0.6 M frame-parallel), so there the on-chip kernel is kept whenever it fits.
        const double tasks_per_sm = (double)plan.slots * code->n_checks / std::max(1, h->levels);
        const bool enough_tasks = tasks_per_sm >= 256.0 || h->max_frames < 32768;
        const int min_slots = params->kernel == 2 ? 1 : (enough_tasks ? 2 : (1 << 30));
        if (slots >= min_slots) {
            h->kernel = 2; h->rp_G = plan.G; h->rp_P = plan.P; h->rp_slots = slots; h->rp_groups = (slots + plan.P - 1) / plan.P;
            h->rp_nsteps = (int)plan.steps.size(); h->rp_nruns = (int)plan.runs.size(); h->rp_static = plan.static_nrows; h->rp_pair_fastest = plan.pair_fastest;
            h->rp_smem = fixed_bytes(slots) + pair_bytes * slots;
            CREATE_TRY(cudaMalloc((void**)&h->d_runs, plan.runs.size() * sizeof(RpRun)));
            CREATE_TRY(cudaMemcpy(h->d_runs, plan.runs.data(), plan.runs.size() * sizeof(RpRun), cudaMemcpyHostToDevice));
            CREATE_TRY(cudaMalloc((void**)&h->d_steps, plan.steps.size() * sizeof(RpStep)));
            CREATE_TRY(cudaMemcpy(h->d_steps, plan.steps.data(), plan.steps.size() * sizeof(RpStep), cudaMemcpyHostToDevice));
            CREATE_TRY(cudaMalloc((void**)&h->d_idx_t, plan.idx_t.size() * sizeof(uint16_t)));
            CREATE_TRY(cudaMemcpy(h->d_idx_t, plan.idx_t.data(), plan.idx_t.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
            CREATE_TRY(cudaMalloc((void**)&h->d_edge_of, plan.edge_of.size() * sizeof(uint32_t)));
            CREATE_TRY(cudaMemcpy(h->d_edge_of, plan.edge_of.data(), plan.edge_of.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        } else if (params->kernel == 2) { destroy_impl(h); return fail(nullptr, LDPC_ERR_UNSUPPORTED, "code state does not fit in shared memory for the row-parallel kernel"); }
    } else if (params->kernel == 2) { destroy_impl(h); return fail(nullptr, LDPC_ERR_UNSUPPORTED, "row-parallel kernel needs n <= 16383"); }
    // frame-parallel family: the staged variant (kernel 4) when the code allows it — degrees 3..10, enough rows for the hazard
    // window to be a small part of an iteration
    if (h->kernel == 1) {
        bool ok = code->n_checks >= 8 * FS_HAZARD;
        int dmax = 0;
        for (int i = 0; i < code->nb_deg; i++) { ok = ok && code->deg[i] >= 3 && code->deg[i] <= FS_GEN_MAXDEG && code->n <= (int)FS_IDX_MASK; dmax = std::max(dmax, code->deg[i]); }
        if (params->kernel == 4 && !ok) { destroy_impl(h); return fail(nullptr, LDPC_ERR_UNSUPPORTED, "staged frame-parallel kernel: needs row degrees 3..32 and >= 128 rows"); }
        // rows wider than FS_MAXDEG run through the two-pass row body (fs_row_generic): no row summary, 5-bit writer slots
        const bool wide = dmax > FS_MAXDEG;
        // wide rows: the staged kernel runs one CTA per SM on them (wide stages) and a two-pass row body — 2.6x the plain kernel at 64 Ki
        // DVB-S2 rate-8/9 frames, level with it at 256 Ki (profiles/r02_sweep_fs_wide.jsonl): chosen below 256 Ki frames of capacity, or on request
        if (ok && params->kernel != 1 && (!wide || params->kernel == 4 || h->max_frames < 262144)) {
            // hazard flags: an edge whose variable was touched by one of the FS_HAZARD previous rows (cyclic over the iteration boundary)
            std::vector<uint32_t> pos2((size_t)code->m + FS_P2_PAD, 0u);      // padded: the consumers fetch one row ahead
            std::vector<long> last((size_t)code->n, -(long)(1 << 30));
            std::vector<int> last_slot((size_t)code->n, 0);
            std::vector<uint8_t> stair_row((size_t)code->n_checks, 0);      // exactly one hazard edge: slot D-2, written by slot D-1 of the row before
            for (int lap = 0; lap < 2; lap++) {
                long q = (long)lap * code->n_checks; size_t e = 0;
                for (int k = 0; k < code->nb_deg; k++)
                    for (int r = 0; r < code->rows[k]; r++, q++) {
                        const size_t e0 = e;
                        int nhaz = 0, hslot = 0, wslot = 0; long hback = 0; bool fwd = false;
                        for (int j = 0; j < code->deg[k]; j++, e++) {
                            const uint32_t v = code->pos[e];
                            const long back = q - last[v];
                            if (lap == 1) {
                                uint32_t w = v;
                                if (back <= FS_HAZARD) { w |= FS_F_HAZARD; nhaz++; hslot = j; wslot = last_slot[v]; hback = back; fwd = back <= FS_FWD; }
                                if (back <= FS_FWD) w |= FS_F_FWD | ((uint32_t)(back - 1) << 28) | ((uint32_t)last_slot[v] << (wide ? 23 : 24));
                                pos2[e] = w;
                            }
                            last[v] = q; last_slot[v] = j;
                        }
                        if (lap == 1 && !wide) {     // row summary (kernel_fs.cuh: FS_ROW_SHIFT)
                            const uint32_t hs = nhaz == 0 ? FS_ROW_NONE : (nhaz == 1 && fwd) ? (uint32_t)hslot : FS_ROW_GENERIC;
                            pos2[e0] |= hs << FS_ROW_SHIFT;
                            pos2[e0 + 1] |= (uint32_t)wslot << FS_ROW_SHIFT;
                            pos2[e0 + 2] |= (uint32_t)(hback > 0 ? (hback - 1) & 3 : 0) << FS_ROW_SHIFT;
                            const int D = code->deg[k];
                            stair_row[(size_t)(q - code->n_checks)] = (nhaz == 1 && hslot == D - 2 && wslot == D - 1 && hback == 1 && D >= 6 && D <= 8) ? 1 : 0;
                        }
                    }
            }
            CREATE_TRY(cudaMalloc((void**)&h->d_pos2, pos2.size() * sizeof(uint32_t)));
            CREATE_TRY(cudaMemcpy(h->d_pos2, pos2.data(), pos2.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            h->kernel = 4; h->fs_max_deg = dmax;
            // consumer-side segments: every degree class, cut around its runs of >= FS_STAIR_MIN staircase rows (reserved[4] bit 18: none)
            {
                struct Seg { int deg, rows, cls, stair; };
                std::vector<Seg> segs;
                const bool no_stair = (params->reserved[4] >> 18) & 1;
                size_t row = 0;
                for (int k = 0; k < code->nb_deg; k++) {
                    const size_t r0 = row, r1 = row + (size_t)code->rows[k];
                    size_t a = r0;
                    while (a < r1) {
                        size_t b = a;
                        const bool st = !no_stair && stair_row[a];
                        while (b < r1 && (bool)(!no_stair && stair_row[b]) == st) b++;
                        const bool keep = st && b - a >= FS_STAIR_MIN;
                        if (!segs.empty() && segs.back().cls == k && !segs.back().stair && !keep) segs.back().rows += (int)(b - a);
                        else segs.push_back({ code->deg[k], (int)(b - a), k, keep ? 1 : 0 });
                        a = b;
                    }
                    row = r1;
                }
                if (segs.size() > FS_MAXSEG) { segs.clear(); for (int k = 0; k < code->nb_deg; k++) segs.push_back({ code->deg[k], code->rows[k], k, 0 }); }
                h->fs_nseg = (int)segs.size();
                for (size_t i = 0; i < segs.size(); i++) { h->fs_seg_deg[i] = segs[i].deg; h->fs_seg_rows[i] = segs[i].rows; h->fs_seg_cls[i] = segs[i].cls; h->fs_seg_stair[i] = segs[i].stair; }
            }
            std::vector<uint32_t> edge_row((size_t)code->m);
            { size_t e = 0; uint32_t row = 0;
              for (int k = 0; k < code->nb_deg; k++) for (int r = 0; r < code->rows[k]; r++, row++) for (int j = 0; j < code->deg[k]; j++) edge_row[e++] = (row << 4) | (uint32_t)j; }
            CREATE_TRY(cudaMalloc((void**)&h->d_edge_row, edge_row.size() * sizeof(uint32_t)));
            CREATE_TRY(cudaMemcpy(h->d_edge_row, edge_row.data(), edge_row.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
    }
    if (h->kernel != 2 && !h->levels) h->levels = ldpc_b200_level_schedule(code, nullptr);
#undef CREATE_TRY
    // pipeline granularity of decode(): a quarter of the declared capacity, at least one full wave of the chosen kernel
    // pipeline granularity of decode(): whole waves of the chosen kernel.  H2D, kernel and D2H take about the same time per
    // frame for 576x288 over PCIe Gen5, so the fill/drain of the 3-stage pipeline costs 2 chunks: many small chunks win
    // (measured with tools/e2e_sweep.py: 5 chunks 1.17 ms, 10 chunks of one wave each 1.11 ms).  reserved[2] overrides (waves per chunk).
    const size_t wave = h->kernel == 2 ? (size_t)h->sms * h->rp_slots * 2 : h->kernel == 6 ? (size_t)h->sms * h->wf_warps * h->wf_ctas_per_sm : h->kernel == 5 ? (size_t)h->sms * h->oc_F : (h->kernel == 3 ? (size_t)h->sms * 1024 : (size_t)h->sms * 512 * 4);
    size_t k = h->kernel == 2 ? 1 : (h->kernel == 5 || h->kernel == 6) ? std::max<size_t>(1, h->max_frames / 8 / wave) : std::max<size_t>(1, (h->max_frames / 4 + wave / 2) / wave);
    if (h->prm.reserved[2] > 0) k = (size_t)h->prm.reserved[2];
    h->chunk_frames = std::max<size_t>(std::min<size_t>(h->max_frames, k * wave), 1);
    if (const char* cf = getenv("LDPC_B200_CHUNK_FRAMES")) { const long v = atol(cf); if (v > 0) h->chunk_frames = (size_t)v; }     // experiment knob (tools/e2e_chunks.py)
    // frame-parallel kernels: a wave is 300 Ki frames, so the rule above never split a batch and decode() ran H2D, kernel and D2H
    // back to back.  Quarter the batch over the four stream slots instead (the chunks' kernels share the SMs: a 32 Ki-frame chunk
    // is 64 CTAs), unless reserved[2] says otherwise.
    if ((h->kernel == 1 || h->kernel == 4) && h->prm.reserved[2] <= 0 && h->max_frames >= 8192)
        h->chunk_frames = std::min<size_t>(h->max_frames, (h->max_frames / kSlots + 511) / 512 * 512);
    *out = h;
    return LDPC_OK;
}

void ldpc_b200_destroy(ldpc_handle h) { destroy_impl(h); }

int ldpc_b200_get_info(ldpc_handle h, int what, int64_t* value)
{
    if (!h || !value) return LDPC_ERR_INVALID;
    switch (what) {
    case LDPC_INFO_KERNEL: *value = h->kernel; break;
    case LDPC_INFO_LEVELS: *value = h->levels; break;
    case LDPC_INFO_SMEM_BYTES: *value = h->kernel == 2 ? (int64_t)h->rp_smem : h->kernel == 6 ? (int64_t)h->wf_smem : h->kernel == 5 ? (int64_t)h->oc_smem : (h->kernel == 1 ? 16384 : 0); break;
    case LDPC_INFO_FRAMES_PER_CTA: *value = h->kernel == 2 ? h->rp_slots * 2 : (h->kernel == 3 ? GP_BLOCK : h->kernel == 6 ? h->wf_warps : h->kernel == 5 ? h->oc_F : (h->kernel == 4 ? FS_CONSUMERS * 4 : FP_BLOCK * 4)); break;
    case LDPC_INFO_LAUNCHES: *value = h->launches; break;
    case LDPC_INFO_STREAM_SLOTS: *value = kSlots; break;
    case LDPC_INFO_DEVICE: *value = h->device; break;
    case LDPC_INFO_FS_STAIR_ROWS: { int64_t n = 0; for (int i = 0; i < h->fs_nseg; i++) if (h->fs_seg_stair[i]) n += h->fs_seg_rows[i]; *value = n; break; }
    case LDPC_INFO_FS_VARIANT: *value = h->fs_last_variant; break;
    default: return fail(h, LDPC_ERR_INVALID, "unknown info key");
    }
    return LDPC_OK;
}

void* ldpc_b200_stream(ldpc_handle h, int slot) { return (h && slot >= 0 && slot < kSlots) ? (void*)h->slot[slot].stream : nullptr; }

int ldpc_b200_set_debug(ldpc_handle h, int enable) { if (!h) return LDPC_ERR_INVALID; h->debug = enable != 0; return LDPC_OK; }

int ldpc_b200_host_alloc(void** p, size_t bytes)
{
    if (!p) return LDPC_ERR_INVALID;
    if (ldpc_b200_device_count() <= 0) return LDPC_ERR_NO_DEVICE;
    return cudaMallocHost(p, bytes) == cudaSuccess ? LDPC_OK : LDPC_ERR_NOMEM;
}
int ldpc_b200_host_alloc_input(void** p, size_t bytes)
{
    if (!p) return LDPC_ERR_INVALID;
    if (ldpc_b200_device_count() <= 0) return LDPC_ERR_NO_DEVICE;
    return cudaHostAlloc(p, bytes, cudaHostAllocWriteCombined | cudaHostAllocPortable) == cudaSuccess ? LDPC_OK : LDPC_ERR_NOMEM;
}
int ldpc_b200_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? LDPC_OK : LDPC_ERR_CUDA; }

int ldpc_b200_device_alloc(ldpc_handle h, void** p, size_t bytes)
{
    if (!h || !p) return fail(h, LDPC_ERR_INVALID, "device_alloc: bad argument");
    CU_TRY(h, cudaSetDevice(h->device));
    if (cudaMalloc(p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return fail(h, LDPC_ERR_NOMEM, "device_alloc: cudaMalloc failed"); }
    return LDPC_OK;
}
int ldpc_b200_device_free(ldpc_handle h, void* p)
{
    if (!h) return LDPC_ERR_INVALID;
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaFree(p));
    return LDPC_OK;
}

int ldpc_b200_decode_device(ldpc_handle h, const void* d_llr, uint8_t* d_hard, size_t frames, int iters, uint8_t* d_iters_done, void* cuda_stream)
{
    if (!h || !d_llr || !d_hard || iters < 0) return fail(h, LDPC_ERR_INVALID, "decode_device: bad argument");
    if (d_iters_done && iters > 255) return fail(h, LDPC_ERR_INVALID, "iteration counts are returned as bytes: iters <= 255 when iters_done is requested");
    if (frames == 0) return LDPC_OK;
    CU_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->slot[0].stream;
    const bool dbg = h->debug;
    int rc;
    if (dbg && (rc = ensure_debug(h, frames, iters))) return rc;
    if ((rc = slot_enter(h, h->slot[0], st))) return rc;
    if ((rc = launch_decode(h, h->slot[0], (const int8_t*)d_llr, d_hard, frames, iters, d_iters_done, st, dbg))) return rc;
    return slot_leave(h, h->slot[0], st);
}

int ldpc_b200_decode_async(ldpc_handle h, int slot, const void* llr, uint8_t* hard, size_t frames, int iters, uint8_t* iters_done)
{
    if (!h || slot < 0 || slot >= kSlots || !llr || !hard || iters < 0) return fail(h, LDPC_ERR_INVALID, "decode_async: bad argument");
    if (iters_done && iters > 255) return fail(h, LDPC_ERR_INVALID, "iteration counts are returned as bytes: iters <= 255 when iters_done is requested");
    if (frames == 0) return LDPC_OK;
    CU_TRY(h, cudaSetDevice(h->device));
    Slot& s = h->slot[slot];
    const size_t n = h->code.n, hb = hard_row_bytes(h), el = (size_t)h->elem;
    int rc;
    // a no-op when decode() reserved the slots up front; a direct caller with growing batches pays the (device-synchronising)
    // reallocation here
    if ((rc = reserve_slot(h, s, frames, iters_done != nullptr))) return rc;
    uint8_t* const d_it = iters_done ? s.d_iters : nullptr;
    if ((rc = slot_enter(h, s, s.stream))) return rc;
    CU_TRY(h, cudaMemcpyAsync(s.d_llr, llr, frames * n * el, cudaMemcpyHostToDevice, s.stream));
    const bool dbg = h->debug && slot == 0;
    if (dbg && (rc = ensure_debug(h, frames, iters))) return rc;
    if ((rc = launch_decode(h, s, s.d_llr, s.d_hard, frames, iters, d_it, s.stream, dbg))) return rc;
    CU_TRY(h, cudaMemcpyAsync(hard, s.d_hard, frames * hb, cudaMemcpyDeviceToHost, s.stream));
    if (iters_done) CU_TRY(h, cudaMemcpyAsync(iters_done, d_it, frames, cudaMemcpyDeviceToHost, s.stream));
    return slot_leave(h, s, s.stream);
}

int ldpc_b200_sync(ldpc_handle h, int slot)
{
    if (!h || slot < -1 || slot >= kSlots) return fail(h, LDPC_ERR_INVALID, "sync: bad slot");
    CU_TRY(h, cudaSetDevice(h->device));
    for (int i = 0; i < kSlots; i++) if (slot < 0 || slot == i) CU_TRY(h, cudaStreamSynchronize(h->slot[i].stream));
    return LDPC_OK;
}

int ldpc_b200_decode(ldpc_handle h, const void* llr, uint8_t* hard, size_t frames, int iters, uint8_t* iters_done)
{
    if (!h || !llr || !hard || iters < 0) return fail(h, LDPC_ERR_INVALID, "decode: bad argument");
    const size_t n = h->code.n, hb = hard_row_bytes(h);
    const size_t chunk = (h->debug) ? std::max<size_t>(frames, 1) : h->chunk_frames;   // debug: one chunk so the state is whole
    int rc = LDPC_OK, k = 0;
    if (frames == 0) return LDPC_OK;
    CU_TRY(h, cudaSetDevice(h->device));
    {   // size the slots this call will use BEFORE anything asynchronous is queued: the first chunk is the largest
        const size_t nchunks = (frames + chunk - 1) / chunk, first = std::min(chunk, frames);
        for (int s = 0; s < kSlots && (size_t)s < nchunks && !h->debug; s++)
            if ((rc = reserve_slot(h, h->slot[s], first, iters_done != nullptr))) return rc;
    }
    // Uniform chunks of one kernel wave.  Measured alternatives (profiles/r02_e2e_chunks.jsonl, r02_e2e_ramp.jsonl): smaller chunks lose to
    // per-chunk host work (half a wave 16.8 against 18.4 Gb/s), larger ones to the pipeline's fill and drain, and a ramped schedule
    // (quarter and half chunks at both ends) changes nothing — what separates the call from the host link's rate is a fixed ~0.17 ms.
    for (size_t f = 0; f < frames && rc == LDPC_OK; f += chunk, k++) {
        const size_t cnt = std::min(chunk, frames - f);
        const int slot = h->debug ? 0 : (k % kSlots);
        // A slot's device buffers are reused by chunk k + kSlots on the SAME stream, i.e. in order behind chunk k's copy out: no host
        // synchronisation is needed for that, and not blocking here lets the host queue the whole batch ahead of the copy engines.
        rc = ldpc_b200_decode_async(h, slot, (const int8_t*)llr + f * n * (size_t)h->elem, hard + f * hb, cnt, iters, iters_done ? iters_done + f : nullptr);
    }
    int rc2 = ldpc_b200_sync(h, -1);
    return rc ? rc : rc2;
}

int ldpc_b200_debug_state(ldpc_handle h, void* posteriors, void* msgs, size_t frames)
{
    if (!h) return LDPC_ERR_INVALID;
    if (!h->debug || !h->d_dbg_post || frames > h->dbg_frames) return fail(h, LDPC_ERR_INVALID, "debug_state: enable ldpc_b200_set_debug and decode first (frames <= last decode)");
    CU_TRY(h, cudaSetDevice(h->device));
    CU_TRY(h, cudaDeviceSynchronize());
    if (posteriors) CU_TRY(h, cudaMemcpy(posteriors, h->d_dbg_post, frames * (size_t)h->code.n * h->elem, cudaMemcpyDeviceToHost));
    if (msgs) CU_TRY(h, cudaMemcpy(msgs, h->d_dbg_msgs, frames * (size_t)h->code.m * h->elem, cudaMemcpyDeviceToHost));
    return LDPC_OK;
}

int ldpc_b200_quantize(ldpc_handle h, const float* y, int8_t* q, size_t count)
{
    if (!h || !y || !q) return fail(h, LDPC_ERR_INVALID, "quantize: bad argument");
    if (count == 0) return LDPC_OK;
    CU_TRY(h, cudaSetDevice(h->device));
    int rc;
    if ((rc = ensure(h, &h->d_qy, &h->qy_bytes, count * sizeof(float)))) return rc;       // scratch lives on the handle: no cudaMalloc / cudaFree per call
    if ((rc = ensure(h, &h->d_qq, &h->qq_bytes, count))) return rc;
    cudaStream_t st = h->slot[0].stream;
    CU_TRY(h, cudaMemcpyAsync(h->d_qy, y, count * sizeof(float), cudaMemcpyHostToDevice, st));
    quantize_kernel<<<(unsigned)std::min<size_t>((count + 255) / 256, 65535), 256, 0, st>>>(h->d_qy, h->d_qq, count, (float)h->prm.llr_scale, h->prm.sat_llr);
    CU_TRY(h, cudaGetLastError());
    CU_TRY(h, cudaMemcpyAsync(q, h->d_qq, count, cudaMemcpyDeviceToHost, st));
    const cudaError_t e = cudaStreamSynchronize(st);
    h->launches += 1;
    if (e != cudaSuccess) return fail(h, LDPC_ERR_CUDA, cudaGetErrorString(e));
    return LDPC_OK;
}

int ldpc_b200_awgn_device(ldpc_handle h, void* d_llr, size_t frames, float sigma, uint64_t seed, uint64_t first_frame, void* cuda_stream)
{
    if (!h || !d_llr) return fail(h, LDPC_ERR_INVALID, "awgn_device: bad argument");
    if (frames == 0) return LDPC_OK;
    CU_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->slot[0].stream;
    const size_t total = frames * (size_t)((h->code.n + 3) / 4);
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 1u << 20);
    const float scale = (float)h->prm.llr_scale; const int sat = h->prm.sat_llr;
    if (h->elem == 4) awgn_kernel<float><<<blocks, 256, 0, st>>>((float*)d_llr, frames, h->code.n, sigma, seed, first_frame, scale, sat);
    else if (h->elem == 2) awgn_kernel<int16_t><<<blocks, 256, 0, st>>>((int16_t*)d_llr, frames, h->code.n, sigma, seed, first_frame, scale, sat);
    else awgn_kernel<int8_t><<<blocks, 256, 0, st>>>((int8_t*)d_llr, frames, h->code.n, sigma, seed, first_frame, scale, sat);
    CU_TRY(h, cudaGetLastError());
    h->launches += 1;
    return LDPC_OK;
}

int ldpc_b200_awgn(ldpc_handle h, void* llr_host, size_t frames, float sigma, uint64_t seed, uint64_t first_frame)
{
    if (!h || !llr_host) return fail(h, LDPC_ERR_INVALID, "awgn: bad argument");
    if (frames == 0) return LDPC_OK;
    CU_TRY(h, cudaSetDevice(h->device));
    Slot& s = h->slot[0];
    int rc;
    if ((rc = ensure(h, &s.d_llr, &s.llr_bytes, frames * (size_t)h->code.n * h->elem))) return rc;
    if ((rc = ldpc_b200_awgn_device(h, s.d_llr, frames, sigma, seed, first_frame, s.stream))) return rc;
    CU_TRY(h, cudaMemcpyAsync(llr_host, s.d_llr, frames * (size_t)h->code.n * h->elem, cudaMemcpyDeviceToHost, s.stream));
    CU_TRY(h, cudaStreamSynchronize(s.stream));
    return LDPC_OK;
}

int ldpc_b200_count_errors_device(ldpc_handle h, const uint8_t* d_hard, size_t frames, uint64_t* out2_host, void* cuda_stream)
{
    if (!h || !d_hard || !out2_host) return fail(h, LDPC_ERR_INVALID, "count_errors_device: bad argument");
    CU_TRY(h, cudaSetDevice(h->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->slot[0].stream;
    CU_TRY(h, cudaMemsetAsync(h->d_counters, 0, 2 * sizeof(unsigned long long), st));
    if (frames) {
        count_errors_kernel<<<(unsigned)std::min<size_t>((frames + 7) / 8, 4096), 256, 0, st>>>(d_hard, frames, h->code.n, h->code.n - h->code.n_checks,
                                                                                               h->prm.out_format == LDPC_OUT_PACKED, h->d_counters);
        CU_TRY(h, cudaGetLastError());
        h->launches += 1;
    }
    unsigned long long r[2];
    CU_TRY(h, cudaMemcpyAsync(r, h->d_counters, sizeof(r), cudaMemcpyDeviceToHost, st));
    CU_TRY(h, cudaStreamSynchronize(st));
    out2_host[0] = r[0]; out2_host[1] = r[1];
    return LDPC_OK;
}


/* accessor for encoder.cu's channel / counter entry points (not part of the public header) */
int ldpc_b200_internal_channel_params(ldpc_handle h, int* device, int* n, int* n_checks, int* elem, int* llr_scale, int* sat_llr, int* packed,
                                      unsigned long long** d_counters, void** slot0_stream)
{
    if (!h) return LDPC_ERR_INVALID;
    *device = h->device; *n = h->code.n; *n_checks = h->code.n_checks; *elem = h->elem; *llr_scale = h->prm.llr_scale; *sat_llr = h->prm.sat_llr;
    *packed = h->prm.out_format == LDPC_OUT_PACKED; *d_counters = h->d_counters; *slot0_stream = (void*)h->slot[0].stream;
    return LDPC_OK;
}

}  // extern "C"
