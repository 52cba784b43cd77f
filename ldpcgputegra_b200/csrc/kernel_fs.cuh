// kernel_fs.cuh — frame-parallel decoder with the HBM-resident state STAGED through shared memory by the bulk-copy engine.
//
// Same mapping and arithmetic as kernel_fp.cuh (one thread = 4 frames, rows strictly in reference order: the only legal
// mapping for DVB-S2 64800x32400, whose reference order is a 32 399-deep chain — SURVEY App. C), but the loads no longer sit
// on the threads' critical path: two producer warps (posterior lines / message lines) walk the row list K rows ahead of the
// consumers and pull every row's posterior lines V[idx][t0..t0+NC-1] and message lines MSG[e][t0..t0+NC-1] (NC*4 contiguous bytes
// each, NC = 128 | 256 consumer threads) into a K-stage ring with cp.async.bulk (TMA, non-tensor form) completing on an mbarrier
// per stage; consumers read their word from the stage, do the row, store the results straight to HBM, and hand the stage back.
// Memory-level parallelism is then K rows x 14 lines x NC*4 B per CTA whatever the number of resident threads.  Measured
// (DESIGN.md 3.2b): DVB-S2 21.6 Gb/s = 0.94 of the measured HBM peak on algorithmic bytes at a batch that fills every SM alike.
// (The reference re-reads both arrays through a per-row __syncthreads pair: code/gpu_fixed/decoder_oms/cuda/CUDA_OMS_SIMD.cu:141-187.)
//
// Staleness.  A line prefetched for row q was read up to K rows early, so it is wrong if one of the K rows before q wrote the
// same variable.  The host marks those edges (bit 31 of pos2[e], window FS_HAZARD rows, cyclic over the iteration boundary);
// the producer skips them and the consumer loads that word itself, after its own stores in program order (a V word is only ever
// written by its own thread).  DVB-S2: exactly the staircase parity bit of each row.  Everything else is ordered by
// st.global -> fence.proxy.async -> mbarrier arrive (empty) -> producer wait -> cp.async.bulk.
// A store followed one row later by a load of the same word would still cost an L2 round trip per row (first measurement:
// 910 ns per DVB-S2 row, all of it this load), so the last FS_FWD rows' outputs are also kept in a small per-thread ring in
// shared memory and flagged edges whose writer is that close read the ring instead (pos2 bits 30..26: forward, rows back - 1,
// edge slot of the writer).  Messages are private to their edge and a whole iteration old when they are fetched.
// The common pattern — one forwardable hazard edge per row — is summarised per row (FS_ROW_SHIFT) and served by copying the ring word
// into the stage slot.  Early termination (template parameter ET) adds a second, read-only pass over the ring per iteration for the
// per-frame stop criterion (fs_syndrome_sweep) and a CTA-wide exit.
// Round 2: the message lines of a row arrive as one 2-D tensor-map copy and the posterior lines four at a time through
// tile::gather4; the producers are two separate tight loops whose warps stay CONVERGED (every lane waits for the free slot — a wait by
// lane 0 alone left every later shuffle / __syncwarp on the compiler's divergent path and WAS the time per row of every small batch);
// runs of staircase rows carry their hazard word in a register (fs_row_stair), in pairs sharing one basic block where a CTA has room
// for the registers (fs_row_stair2, the PIPE2 instantiation); rows of degree 11..32 take a two-pass body (fs_row_generic); the messages
// may be stored compressed, four words per row (CMP: bit-exact, fewer bytes, more instructions, not faster).
// Roofline: HBM, 4*M bytes per frame-iteration as for kernel_fp.
#pragma once
#include <cuda.h>            // CUtensorMap (type only: the encoder is looked up at run time, ldpc_b200.cu)
#include "kernel_fp.cuh"

namespace ldpcb200 {

#define FS_CONSUMERS 128                 // the narrow CTA; a CTA has NC = 128 or 256 consumer threads + 2 producer warps, the state arrays' row pitch is a multiple of NC words
#define FS_PRODUCER_THREADS 64           // one warp fetches the posterior lines, one the message lines: a bulk copy costs ~46 issue cycles, and
                                         // the single producer of the first version (1340 cycles per row) bounded small batches
#define FS_MAX_CONSUMERS 512
#define FS_P2_BYTES 128u                   // tail of every stage: the row's edge words (<= FS_MAXDEG of them), written by the posterior-side producer
#define FS_P2_OFFSET(A, NC) ((uint32_t)((A).msg_line0 + (A).msg_lines) * (uint32_t)(NC) * 4u)
#define FS_LINE (FS_CONSUMERS * 4)       // bytes per staged line at NC = 128
#define FS_MAXDEG 10                     // 1200x600, the gpu_fixed tree's default code (matrix/code.h:1), has rows of degree 9
#define FS_GEN_MAXDEG 32                 // rows up to this degree run through the two-pass row body (fs_row_generic): DVB-S2 rates 8/9 and 9/10 have rows of 27 and 30
#define FS_HAZARD 16                     // hazard window in rows = the largest ring depth the host may choose
#define FS_FWD 4                         // rows whose outputs stay in the forwarding ring
#define FS_F_HAZARD 0x80000000u          // pos2 flags: not prefetched (written within the hazard window) ...
#define FS_F_FWD    0x40000000u          // ... and the writer is at most FS_FWD rows back: bits 29..28 = rows back - 1, bits 27..24 = its edge slot
#define FS_IDX_MASK 0x000FFFFFu
// Tables with a row degree above FS_MAXDEG carry no row summary and a 5-bit writer slot in bits 27..23 instead (FS_GEN_SLOT).
#define FS_GEN_SLOT(p2) (((p2) >> 23) & 31u)
// Row summary in bits 23..20 of the row's first three words (every row has >= 3 edges): word 0 = slot of the row's ONLY hazard edge
// when that edge can be forwarded (14 = some other hazard pattern, 15 = no hazard), word 1 = the writer's slot, word 2 = rows back - 1.
#define FS_ROW_SHIFT 20
#define FS_ROW_NONE 15u
#define FS_ROW_GENERIC 14u
#define FS_MAXSEG 16                     // consumer-side segments of the row list (see FsArgs::seg_*)
#define FS_P2_PAD 192                    // words of padding behind pos2: the producers' windows (and the one-row-ahead reads of kernel_fa) run past the last row
#define FS_STAIR_MIN 32                  // shortest run of staircase rows worth a segment of its own

struct FsArgs {
    uint32_t* V;
    uint32_t* MSG;
    const uint32_t* pos2;    // [m + FS_MAXDEG] variable index | hazard flags | row summary (see the FS_* masks)
    int T, n, m, nb_deg;
    int deg[LDPC_MAX_DEG_CLASSES];
    int rows[LDPC_MAX_DEG_CLASSES];
    int iters, stages, max_deg;
    uint32_t exp_word;       // 0x64646464 (rowops.cuh: bytes01_to_w)
    uint8_t* iters_done;     // [4*T], nullable
    int et;                  // per-frame syndrome early termination (the ET instantiation)
    int nc;                  // consumer threads per CTA (128 | 256): a staged line is nc * 4 bytes
    int use_tm;              // the message lines of a row arrive as ONE 2-D tensor copy (box = nc words x row degree over MSG[m][T])
    int use_g4;              // the posterior lines arrive four at a time (tile::gather4 over V[n][T]); hazard lines come along and are ignored
    int msg_line0;           // first message line of a stage: max_deg, or max_deg rounded up to 4 with gather4 (it writes whole groups of four)
    // The consumers walk the row list in SEGMENTS: a degree class, cut where a run of STAIRCASE rows begins or ends.  A staircase row has
    // exactly one hazard edge, at slot D - 2, written by slot D - 1 of the row just before it (the parity chain of DVB-S2 and of every
    // IRA code: row r holds p[r-1] and p[r] as its last two edges).  Inside such a run the word travels from row to row in a REGISTER
    // (no forwarding ring, no patch of the stage slot, no proxy fence for it), and row r's own store of p[r] is dropped: row r + 1
    // overwrites it before anything else can read it.
    int nseg;
    int seg_deg[FS_MAXSEG], seg_rows[FS_MAXSEG], seg_cls[FS_MAXSEG], seg_stair[FS_MAXSEG];
    int pipe2;               // the paired-row instantiation (PIPE2): the host asks for it when a CTA has an SM to itself and the code has staircase runs
    int cmp;                 // compressed messages (the CMP instantiation, rows of degree <= 8): MSG is [4 * rows][T], four words per row and thread
    int msg_lines;           // message lines of a stage: max_deg, or 4 when compressed
    ldpc_params_t prm;
    alignas(64) CUtensorMap tm_msg[LDPC_MAX_DEG_CLASSES];    // one map per degree class: the box height is part of the map
    alignas(64) CUtensorMap tm_v;                            // gather4: box = nc words x 1 row
};

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// tensor-map form: box {NC words, D rows} of MSG[m][T] at (t0, e) -> D consecutive lines of the stage (SASS UTMALDG)
__device__ __forceinline__ void tma_g2s_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// four rows of V[n][T] given by index, nc words each at column t0 -> four consecutive lines of the stage (sm_100 tile::gather4)
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, int c0, int r0, int r1, int r2, int r3, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
// generic-proxy global writes -> visible to later async-proxy (bulk copy) reads.  The .global form is a bare FENCE.VIEW.ASYNC.G;
// the unqualified form costs a MEMBAR.ALL.GPU on top (cuobjdump)
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// generic-proxy shared-memory writes -> ordered before later async-proxy (bulk copy) writes of the same stage slot
__device__ __forceinline__ void fence_proxy_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one row for the consumer: words come from the stage unless the edge is flagged (hazard) — then from global, after this
// thread's own earlier stores
// The stage of the PREVIOUS row is handed back here, between this row's arithmetic and its stores: the fence then only has
// to cover stores that were issued a whole row ago, so it never waits on fresh ones.
template <int SEM, int ALGO, int D, bool FIRST, bool Q, int NC, bool ET, bool CMP>
__device__ __forceinline__ void fs_row(const FsArgs& A, int tid, uint32_t* vt, uint32_t* mp, uint32_t T4, const uint32_t (&p2)[D], uint32_t stage_s, const RowConsts& K,
                                       uint32_t prev_empty, int lane, uint32_t fwd_s, uint32_t q, bool fwd_ok, uint32_t keep_lo, uint32_t keep_hi)
{
    constexpr uint32_t LINE = NC * 4u;
    constexpr int ML = CMP ? 4 : D;                // message words per row and thread (kernel_fp.cuh: fp_row_math_c)
    uint32_t wv[D], wm[ML], nv[D], nm[ML];
    // Hazards (edges the producer did not prefetch because one of the last FS_HAZARD rows wrote the variable).  The common pattern —
    // exactly one such edge per row, its writer at most FS_FWD rows back: the staircase of DVB-S2 and of every IRA code — is
    // served by copying the writer's word from the forwarding ring INTO the stage slot (shared memory takes a run-time slot
    // number, a register array does not), after which the row loads all its words from the stage like a row without hazards.
    // The per-edge flag tests of the first version cost ~40 issued instructions per row for one flagged edge; this costs ~12.
    const uint32_t hs = (p2[0] >> FS_ROW_SHIFT) & 15u;
    bool generic = false;
    if (hs != FS_ROW_NONE) {
        if (hs < FS_ROW_GENERIC && fwd_ok) {
            const uint32_t ws = (p2[1] >> FS_ROW_SHIFT) & 15u, back = ((p2[2] >> FS_ROW_SHIFT) & 3u) + 1u;
            const uint32_t v = lds_u32(fwd_s + ((((q - back) & (FS_FWD - 1)) * (uint32_t)A.max_deg + ws) * NC + tid) * 4u);
            sts_u32(stage_s + hs * LINE + 4 * tid, v);        // read back below by this same thread: program order
            // The slot is overwritten by a bulk copy (async proxy) once the stage has been handed back, one row from now.  PTX orders a
            // generic-proxy access before a later async-proxy access of the same shared location only through a proxy fence of that state
            // space (memory consistency model, "proxy fence": fence.proxy.async[.shared::cta]) followed by the release/acquire chain this
            // thread -> __syncwarp -> lane 0's mbarrier.arrive (release.cta) -> producer's try_wait (acquire.cta) -> cp.async.bulk.
            fence_proxy_async_shared();
        } else generic = true;
    }
#pragma unroll
    for (int j = 0; j < D; j++) wv[j] = lds_u32(stage_s + j * LINE + 4 * tid);
    if (generic) {
#pragma unroll
        for (int j = 0; j < D; j++) {
            if (p2[j] & FS_F_HAZARD) {
                if ((p2[j] & FS_F_FWD) && fwd_ok) {
                    const uint32_t back = ((p2[j] >> 28) & 3u) + 1u, slot = (p2[j] >> 24) & 15u;
                    wv[j] = lds_u32(fwd_s + ((((q - back) & (FS_FWD - 1)) * (uint32_t)A.max_deg + slot) * NC + tid) * 4u);
                } else wv[j] = *word_at(vt, p2[j] & FS_IDX_MASK, T4);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < ML; j++) wm[j] = FIRST ? 0x80808080u : lds_u32(stage_s + (A.msg_line0 + j) * LINE + 4 * tid);
    if constexpr (CMP) fp_row_math_c<SEM, ALGO, D, FIRST, ET, Q>(wv, wm, K, keep_lo, keep_hi, nv, nm);
    else fp_row_math<SEM, ALGO, D, FIRST, ET, Q>(wv, wm, K, keep_lo, keep_hi, nv, nm);   // ET: frozen frames keep their state
    fence_proxy_async_global();                    // the previous rows' stores, before any later bulk copy of the same lines
    __syncwarp();
    if (lane == 0 && prev_empty) mbar_arrive(prev_empty);
#pragma unroll
    for (int j = 0; j < D; j++) {
        *word_at(vt, p2[j] & FS_IDX_MASK, T4) = nv[j];          // vt = V + t, mp = MSG + (first message line of the row) * T + t, T4 = 4 * T
        if (!CMP) *word_at(mp, (uint32_t)j, T4) = nm[j];
        sts_u32(fwd_s + (((q & (FS_FWD - 1)) * (uint32_t)A.max_deg + j) * NC + tid) * 4u, nv[j]);
    }
    if (CMP) {
#pragma unroll
        for (int j = 0; j < ML; j++) *word_at(mp, (uint32_t)j, T4) = nm[j];
    }
}

// a STAIRCASE row (FsArgs::seg_*): the hazard word arrives in `carry`, the word the next row needs leaves in it
template <int SEM, int ALGO, int D, bool FIRST, bool Q, int NC, bool ET, bool CMP>
__device__ __forceinline__ void fs_row_stair(const FsArgs& A, int tid, uint32_t* vt, uint32_t* mp, uint32_t T4, const uint32_t (&p2)[D], uint32_t stage_s, const RowConsts& K,
                                             uint32_t prev_empty, int lane, uint32_t fwd_s, uint32_t q, bool ring_out, bool store_ws, uint32_t& carry,
                                             uint32_t keep_lo, uint32_t keep_hi)
{
    constexpr uint32_t LINE = NC * 4u;
    constexpr int HS = D - 2, WS = D - 1;
    constexpr int ML = CMP ? 4 : D;
    uint32_t wv[D], wm[ML], nv[D], nm[ML];
#pragma unroll
    for (int j = 0; j < D; j++) wv[j] = (j == HS) ? carry : lds_u32(stage_s + j * LINE + 4 * tid);
#pragma unroll
    for (int j = 0; j < ML; j++) wm[j] = FIRST ? 0x80808080u : lds_u32(stage_s + (A.msg_line0 + j) * LINE + 4 * tid);
    if constexpr (CMP) fp_row_math_c<SEM, ALGO, D, FIRST, ET, Q>(wv, wm, K, keep_lo, keep_hi, nv, nm);
    else fp_row_math<SEM, ALGO, D, FIRST, ET, Q>(wv, wm, K, keep_lo, keep_hi, nv, nm);
    carry = nv[WS];
    fence_proxy_async_global();
    __syncwarp();
    if (lane == 0 && prev_empty) mbar_arrive(prev_empty);
#pragma unroll
    for (int j = 0; j < D; j++) {
        if (j != WS || store_ws) *word_at(vt, p2[j] & FS_IDX_MASK, T4) = nv[j];
        if (!CMP) *word_at(mp, (uint32_t)j, T4) = nm[j];
    }
    if (CMP) {
#pragma unroll
        for (int j = 0; j < ML; j++) *word_at(mp, (uint32_t)j, T4) = nm[j];
    }
    if (ring_out) {       // the last FS_FWD rows of a run: the rows after it may forward from them
#pragma unroll
        for (int j = 0; j < D; j++) sts_u32(fwd_s + (((q & (FS_FWD - 1)) * (uint32_t)A.max_deg + j) * NC + tid) * 4u, nv[j]);
    }
}

// TWO consecutive staircase rows in one basic block (the small-batch instantiation, PIPE2): with one consumer warp per scheduler a row
// is a serial chain of ~400 instructions, and everything of row B that does not depend on row A — six of its seven posterior words,
// all its messages, their unpacking, saturation, magnitudes and most of the min search — can issue between A's dependent
// instructions once both rows' words are in registers.  Neither row is one of a run's last FS_FWD (no ring writes, no store of
// slot D - 1).  A's stage is handed back together with the stage of the row before it: its words are in registers by then.
template <int SEM, int ALGO, int D, bool FIRST, bool Q, int NC, bool ET, bool CMP>
__device__ __forceinline__ void fs_row_stair2(const FsArgs& A, int tid, uint32_t* vt, uint32_t* mp, uint32_t T4, const uint32_t (&p2a)[D], const uint32_t (&p2b)[D],
                                              uint32_t st_a, uint32_t st_b, const RowConsts& K, uint32_t prev_empty, uint32_t empty_a, int lane, uint32_t& carry,
                                              uint32_t keep_lo, uint32_t keep_hi)
{
    constexpr uint32_t LINE = NC * 4u;
    constexpr int HS = D - 2, WS = D - 1;
    constexpr int ML = CMP ? 4 : D;
    uint32_t wva[D], wma[ML], nva[D], nma[ML], wvb[D], wmb[ML], nvb[D], nmb[ML];
#pragma unroll
    for (int j = 0; j < D; j++) wva[j] = (j == HS) ? carry : lds_u32(st_a + j * LINE + 4 * tid);
#pragma unroll
    for (int j = 0; j < ML; j++) wma[j] = FIRST ? 0x80808080u : lds_u32(st_a + (A.msg_line0 + j) * LINE + 4 * tid);
#pragma unroll
    for (int j = 0; j < D; j++) wvb[j] = (j == HS) ? 0u : lds_u32(st_b + j * LINE + 4 * tid);
#pragma unroll
    for (int j = 0; j < ML; j++) wmb[j] = FIRST ? 0x80808080u : lds_u32(st_b + (A.msg_line0 + j) * LINE + 4 * tid);
    if constexpr (CMP) fp_row_math_c<SEM, ALGO, D, FIRST, ET, Q>(wva, wma, K, keep_lo, keep_hi, nva, nma);
    else fp_row_math<SEM, ALGO, D, FIRST, ET, Q>(wva, wma, K, keep_lo, keep_hi, nva, nma);
    wvb[HS] = nva[WS];
    if constexpr (CMP) fp_row_math_c<SEM, ALGO, D, FIRST, ET, Q>(wvb, wmb, K, keep_lo, keep_hi, nvb, nmb);
    else fp_row_math<SEM, ALGO, D, FIRST, ET, Q>(wvb, wmb, K, keep_lo, keep_hi, nvb, nmb);
    carry = nvb[WS];
    fence_proxy_async_global();
    __syncwarp();
    if (lane == 0) { if (prev_empty) mbar_arrive(prev_empty); mbar_arrive(empty_a); }
    uint32_t* const mpb = word_at(mp, (uint32_t)ML, T4);
#pragma unroll
    for (int j = 0; j < D; j++) {
        if (j != WS) *word_at(vt, p2a[j] & FS_IDX_MASK, T4) = nva[j];
        if (!CMP) *word_at(mp, (uint32_t)j, T4) = nma[j];
    }
#pragma unroll
    for (int j = 0; j < D; j++) {
        if (j != WS) *word_at(vt, p2b[j] & FS_IDX_MASK, T4) = nvb[j];
        if (!CMP) *word_at(mpb, (uint32_t)j, T4) = nmb[j];
    }
    if (CMP) {
#pragma unroll
        for (int j = 0; j < ML; j++) { *word_at(mp, (uint32_t)j, T4) = nma[j]; *word_at(mpb, (uint32_t)j, T4) = nmb[j]; }
    }
}

// consumer-side ring cursor
struct FsCursor { int stage; uint32_t phase, prev_empty, q; };

// all rows of one degree class.  The row's edge words (variable index + hazard flags) are fetched ONE ROW AHEAD: they come from
// global memory (L1 hit for 3 rows out of 4, L2 otherwise) and used to sit at the head of every row's dependency chain
// (profiles/r01_ncu_fs_v2.txt: 15 % of the stall samples on the long scoreboard).  pos2 is padded by FS_MAXDEG words so that the
// read past the last row is harmless; the words fetched across a class boundary are simply dropped.
template <int SEM, int ALGO, int D, bool FIRST, bool Q, int NC, bool ET, bool CMP, bool STAIR, bool PIPE2>
__device__ __forceinline__ void fs_class(const FsArgs& A, int tid, int t, size_t& e, size_t& rho, int R, const RowConsts& K, int lane, uint32_t bars, uint32_t ring,
                                         uint32_t stage_bytes, uint32_t fwd_s, int Kst, FsCursor& c, uint32_t keep_lo, uint32_t keep_hi)
{
    const uint32_t T4 = 4u * (uint32_t)A.T;
    uint32_t* const vt = A.V + t;
    uint32_t* mp = A.MSG + ((CMP ? 4 * rho : e) * (size_t)A.T + (size_t)t);
    uint32_t p2[D];
    uint32_t carry = 0u;
    rho += (size_t)R;
    for (int r = 0; r < R; r++, e += D) {
        mbar_wait(bars + 8 * c.stage, c.phase);
        const uint32_t st = ring + (uint32_t)c.stage * stage_bytes;
        if constexpr (STAIR && PIPE2) {
            if (r >= 1 && r + 1 + FS_FWD < R) {       // rows r and r + 1 as a pair (row 0 sets up the carry, the last FS_FWD rows feed the ring)
                const int stage_b = c.stage + 1 == Kst ? 0 : c.stage + 1;
                const uint32_t phase_b = c.stage + 1 == Kst ? c.phase ^ 1u : c.phase;
                mbar_wait(bars + 8 * stage_b, phase_b);
                const uint32_t st_b = ring + (uint32_t)stage_b * stage_bytes;
                uint32_t p2b[D];
#pragma unroll
                for (int j4 = 0; j4 < D; j4 += 4) {
                    const uint4 w = lds_u128(st + FS_P2_OFFSET(A, NC) + 4u * j4), wb = lds_u128(st_b + FS_P2_OFFSET(A, NC) + 4u * j4);
                    p2[j4] = w.x; if (j4 + 1 < D) p2[j4 + 1] = w.y; if (j4 + 2 < D) p2[j4 + 2] = w.z; if (j4 + 3 < D) p2[j4 + 3] = w.w;
                    p2b[j4] = wb.x; if (j4 + 1 < D) p2b[j4 + 1] = wb.y; if (j4 + 2 < D) p2b[j4 + 2] = wb.z; if (j4 + 3 < D) p2b[j4 + 3] = wb.w;
                }
                fs_row_stair2<SEM, ALGO, D, FIRST, Q, NC, ET, CMP>(A, tid, vt, mp, T4, p2, p2b, st, st_b, K, c.prev_empty, bars + 8 * (Kst + c.stage), lane, carry, keep_lo, keep_hi);
                mp = word_at(mp, (uint32_t)(2 * (CMP ? 4 : D)), T4);
                c.prev_empty = bars + 8 * (Kst + stage_b); c.q += 2;
                c.stage = stage_b; c.phase = phase_b;
                if (++c.stage == Kst) { c.stage = 0; c.phase ^= 1u; }
                r++; e += D;
                continue;
            }
        }
        // the row's edge words: written behind the stage's lines by the producer that fetched them (round 2: the consumers' own
        // row-ahead __ldg sat on the long scoreboard for 14 % of their samples once the producers had stopped being the bottleneck,
        // profiles/r02_ncu_fs_small_v2.txt) — a broadcast shared-memory load behind the barrier they wait on anyway
#pragma unroll
        for (int j4 = 0; j4 < D; j4 += 4) {              // 16-byte loads: the tail of a stage is 128 bytes, whatever the row's degree
            const uint4 w = lds_u128(st + FS_P2_OFFSET(A, NC) + 4u * j4);
            p2[j4] = w.x; if (j4 + 1 < D) p2[j4 + 1] = w.y; if (j4 + 2 < D) p2[j4 + 2] = w.z; if (j4 + 3 < D) p2[j4 + 3] = w.w;
        }
        if constexpr (STAIR) {
            if (r == 0)     // the run's first row: the word comes from the ring (the row before wrote it there) or, at the very start, from memory
                carry = c.q >= 1 ? lds_u32(fwd_s + ((((c.q - 1u) & (FS_FWD - 1)) * (uint32_t)A.max_deg + (uint32_t)(D - 1)) * NC + tid) * 4u)
                                 : *word_at(vt, p2[D - 2] & FS_IDX_MASK, T4);
            fs_row_stair<SEM, ALGO, D, FIRST, Q, NC, ET, CMP>(A, tid, vt, mp, T4, p2, st, K, c.prev_empty, lane, fwd_s, c.q, r + FS_FWD >= R, r + 1 >= R, carry, keep_lo, keep_hi);
        } else
        fs_row<SEM, ALGO, D, FIRST, Q, NC, ET, CMP>(A, tid, vt, mp, T4, p2, st, K, c.prev_empty, lane, fwd_s, c.q, c.q >= FS_FWD, keep_lo, keep_hi);
        mp = word_at(mp, (uint32_t)(CMP ? 4 : D), T4);
        c.prev_empty = bars + 8 * (Kst + c.stage); c.q++;
        if (++c.stage == Kst) { c.stage = 0; c.phase ^= 1u; }
    }
}

// ---- rows of ANY degree up to FS_GEN_MAXDEG (the MAXD = FS_GEN_MAXDEG instantiation): two passes over the stage, nothing of the row in
// registers — the min search first, then every edge recomputed from the staged words with the row's two constants (the structure of
// fp_row_generic in kernel_fp.cuh, with shared memory where that one reads HBM twice).  Hazard words (forwarding ring or memory) are
// written into their stage slot during pass 1 so that pass 2 finds them there.
template <int SEM, int ALGO, bool FIRST, bool Q, int NC, bool ET>
__device__ __noinline__ void fs_row_generic(const FsArgs& A, int tid, uint32_t* vt, uint32_t* mp, uint32_t T4, int D, uint32_t stage_s, const RowConsts& K,
                                            uint32_t prev_empty, int lane, uint32_t fwd_s, uint32_t q, bool fwd_ok, uint32_t keep_lo, uint32_t keep_hi)
{
    constexpr uint32_t LINE = NC * 4u;
    const h2 inv256 = h2_const(1.0f / 256.0f), half = h2_const(0.5f);
    const uint32_t p2s = stage_s + FS_P2_OFFSET(A, NC), vs = stage_s + 4u * (uint32_t)tid, ms = vs + (uint32_t)A.msg_line0 * LINE;
    RowState s[2]; row_begin(s[0], K); row_begin(s[1], K);
    bool patched = false;
    for (int j = 0; j < D; j++) {
        const uint32_t p2 = lds_u32(p2s + 4u * j);
        uint32_t wv;
        if (p2 & FS_F_HAZARD) {                                    // warp-uniform: the flags belong to the row
            if ((p2 & FS_F_FWD) && fwd_ok) {
                const uint32_t back = ((p2 >> 28) & 3u) + 1u;
                wv = lds_u32(fwd_s + ((((q - back) & (FS_FWD - 1)) * (uint32_t)A.max_deg + FS_GEN_SLOT(p2)) * NC + tid) * 4u);
            } else wv = *word_at(vt, p2 & FS_IDX_MASK, T4);
            sts_u32(vs + j * LINE, wv);
            patched = true;
        } else wv = lds_u32(vs + j * LINE);
        const uint32_t wm = FIRST ? 0x80808080u : lds_u32(ms + j * LINE);
#pragma unroll
        for (int g = 0; g < 2; g++) {
            const h2 wU = g ? bytes23_to_w(wv, K.c64) : bytes01_to_w(wv, K.c64);
            const h2 wM = g ? bytes23_to_w(wm, K.c64) : bytes01_to_w(wm, K.c64);
            const h2 xu = __hmin2(__hfma2_sat(wU, inv256, __hfma2(wM, __hneg2(inv256), half)), K.top);
            pass1_edge<SEM, ALGO, Q>(s[g], xu, K);
        }
    }
    if (patched) fence_proxy_async_shared();                       // generic-proxy writes of stage slots, before the copy engine refills them (see fs_row)
    RowOut o[2]; row_finish<SEM, ALGO>(s[0], D, K, K.msg_c, o[0]); row_finish<SEM, ALGO>(s[1], D, K, K.msg_c, o[1]);
    fence_proxy_async_global();                                    // the previous rows' stores, before any later bulk copy of the same lines
    __syncwarp();
    if (lane == 0 && prev_empty) mbar_arrive(prev_empty);
    const uint32_t keep = __byte_perm(keep_lo, keep_hi, 0x6420);
    const uint32_t fwd_row = fwd_s + (((q & (FS_FWD - 1)) * (uint32_t)A.max_deg) * NC + tid) * 4u;
    for (int j = 0; j < D; j++) {
        const uint32_t p2 = lds_u32(p2s + 4u * j);
        const uint32_t wv = lds_u32(vs + j * LINE);
        const uint32_t wm = FIRST ? 0x80808080u : lds_u32(ms + j * LINE);
        uint32_t ov[2], om[2];
#pragma unroll
        for (int g = 0; g < 2; g++) {
            const h2 wU = g ? bytes23_to_w(wv, K.c64) : bytes01_to_w(wv, K.c64);
            const h2 wM = g ? bytes23_to_w(wm, K.c64) : bytes01_to_w(wm, K.c64);
            const h2 xu = __hmin2(__hfma2_sat(wU, inv256, __hfma2(wM, __hneg2(inv256), half)), K.top);
            const h2 a = magnitude<SEM, ALGO, Q>(signed_contrib(xu, K), K);
            h2 msg, unew;
            pass2_edge<SEM>(xu, a, o[g], K, msg, unew);
            ov[g] = q_to_w(unew, 0.0f); om[g] = q_to_w(msg, 128.0f);
        }
        uint32_t nv = pack_bytes(ov[0], ov[1]), nm = pack_bytes(om[0], om[1]);
        if (ET) {   // frozen frames keep their state
            nv = (wv & keep) | (nv & ~keep);
            nm = FIRST ? nm : ((wm & keep) | (nm & ~keep));
        }
        *word_at(vt, p2 & FS_IDX_MASK, T4) = nv;
        *word_at(mp, (uint32_t)j, T4) = nm;
        sts_u32(fwd_row + (uint32_t)j * LINE, nv);
    }
}

template <int SEM, int ALGO, bool FIRST, bool Q, int NC, bool ET>
__device__ __forceinline__ void fs_class_generic(const FsArgs& A, int tid, int t, size_t& e, int D, int R, const RowConsts& K, int lane, uint32_t bars, uint32_t ring,
                                                 uint32_t stage_bytes, uint32_t fwd_s, int Kst, FsCursor& c, uint32_t keep_lo, uint32_t keep_hi)
{
    const uint32_t T4 = 4u * (uint32_t)A.T;
    uint32_t* const vt = A.V + t;
    uint32_t* mp = A.MSG + (e * (size_t)A.T + (size_t)t);
    for (int r = 0; r < R; r++, e += D) {
        mbar_wait(bars + 8 * c.stage, c.phase);
        fs_row_generic<SEM, ALGO, FIRST, Q, NC, ET>(A, tid, vt, mp, T4, D, ring + (uint32_t)c.stage * stage_bytes, K, c.prev_empty, lane, fwd_s, c.q, c.q >= FS_FWD, keep_lo, keep_hi);
        mp = word_at(mp, (uint32_t)D, T4);
        c.prev_empty = bars + 8 * (Kst + c.stage); c.q++;
        if (++c.stage == Kst) { c.stage = 0; c.phase ^= 1u; }
    }
}

// Stop criterion, consumer side: one more pass over the row list through the same ring — the producers fetch EVERY line of a row
// this time (nothing is written, so nothing can be stale once the update pass has drained) — and per row the parity of (x > 0),
// x = sat(v - m) with the updated messages (ref: code/ldpc_decoder_arm/CDecoder/OMS/CDecoder_OMS_fixed_x86.cpp:150-178; the same
// words as fp_syndrome in kernel_fp.cuh).  Returns bit15-of-each-half words: 1 = a check of that frame failed.
template <int NC, bool CMP>
__device__ __forceinline__ void fs_syndrome_sweep(const FsArgs& A, int tid, int lane, const RowConsts& K, int lo, uint32_t bars, uint32_t ring, uint32_t stage_bytes,
                                                  int Kst, FsCursor& c, uint32_t& bad_lo, uint32_t& bad_hi)
{
    constexpr uint32_t LINE = NC * 4u;
    const h2 inv256 = h2_const(1.0f / 256.0f), half = h2_const(0.5f), lo_np = h2_const((float)(lo - 1) / 256.0f);
    bad_lo = bad_hi = 0u;
    for (int cl = 0; cl < A.nb_deg; cl++) {
        const int D = A.deg[cl];
        const uint32_t dpar = (D & 1) ? 0x80008000u : 0u;
        for (int r = 0; r < A.rows[cl]; r++) {
            mbar_wait(bars + 8 * c.stage, c.phase);
            const uint32_t st = ring + (uint32_t)c.stage * stage_bytes + 4u * (uint32_t)tid;
            uint32_t p0 = dpar, p1 = dpar;
            if constexpr (CMP) {
                // the row's four words re-expanded edge by edge (fp_expand_pair in kernel_fp.cuh, with a run-time degree)
                const h2 m4 = h2_const(-4.0f);
                uint32_t c1q[2], c2q[2], es[2];
#pragma unroll
                for (int g = 0; g < 2; g++) {
                    const uint32_t cw = lds_u32(st + (A.msg_line0 + g) * LINE);
                    es[g] = lds_u32(st + (A.msg_line0 + 2 + g) * LINE) << (8 - D);           // edge 0's bits at bit 7 / bit 15 after one more shift
                    c1q[g] = h2_bits(__hfma2(bits_h2(__byte_perm(cw, K.c64, 0x4240)), inv256, m4));
                    c2q[g] = h2_bits(__hfma2(bits_h2(__byte_perm(cw, K.c64, 0x4341)), inv256, m4));
                }
                for (int j = D - 1; j >= 0; j--) {               // parity is order-free: walk the edges from the last, whose bits are in place first
                    const uint32_t wv = lds_u32(st + j * LINE);
                    h2 x[2];
#pragma unroll
                    for (int g = 0; g < 2; g++) {
                        const uint32_t xs = es[g];
                        es[g] <<= 1;
                        uint32_t mask;
                        asm("prmt.b32 %0, %1, %1, 0xAA88;" : "=r"(mask) : "r"(xs));
                        const uint32_t mag = (c2q[g] & mask) | (c1q[g] & ~mask);
                        const h2 nM = __hsub2(m4, bits_h2(and_xor(xs, 0x80008000u, mag)));
                        x[g] = __hmin2(__hfma2_sat(g ? bytes23_to_w(wv, K.c64) : bytes01_to_w(wv, K.c64), inv256, nM), K.top);
                    }
                    p0 ^= h2_bits(__hadd2(x[0], lo_np));
                    p1 ^= h2_bits(__hadd2(x[1], lo_np));
                }
            } else
            for (int j = 0; j < D; j++) {
                const uint32_t wv = lds_u32(st + j * LINE), wm = lds_u32(st + (A.msg_line0 + j) * LINE);
                const h2 x0 = __hmin2(__hfma2_sat(bytes01_to_w(wv, K.c64), inv256, __hfma2(bytes01_to_w(wm, K.c64), __hneg2(inv256), half)), K.top);
                const h2 x1 = __hmin2(__hfma2_sat(bytes23_to_w(wv, K.c64), inv256, __hfma2(bytes23_to_w(wm, K.c64), __hneg2(inv256), half)), K.top);
                p0 ^= h2_bits(__hadd2(x0, lo_np));
                p1 ^= h2_bits(__hadd2(x1, lo_np));
            }
            bad_lo |= p0; bad_hi |= p1;
            __syncwarp();
            if (lane == 0 && c.prev_empty) mbar_arrive(c.prev_empty);       // same hand-back discipline as the update rows: one row late
            c.prev_empty = bars + 8 * (Kst + c.stage);
            if (++c.stage == Kst) { c.stage = 0; c.phase ^= 1u; }
        }
    }
    bad_lo &= 0x80008000u; bad_hi &= 0x80008000u;
}

// ---- producers ---------------------------------------------------------------------------------------------------------------------
// ET: every iteration but the last is followed by a second pass over the rows for the stop criterion (sweep 1: every line of every row,
// hazard flags ignored), each pass behind a CTA barrier so that nothing it reads can be stale.  Returns true when the CTA is done.
template <bool ET>
__device__ __forceinline__ bool fs_producer_pass_end(int sweeps, int sw)
{
    if (ET && sweeps == 2) {
        if (sw == 0) __syncthreads();                      // the consumers' stores of this iteration are out (they fenced)
        else if (!__syncthreads_or(0)) return true;        // every frame of the CTA passed: the consumers leave too
    }
    return false;
}

// message side: one tensor copy per row (lane 0 alone) or one bulk copy per message line
template <int NC, bool ET, bool CMP, bool TM>
__device__ __forceinline__ void fs_produce_msgs(const FsArgs& A, uint32_t bars, uint32_t ring, uint32_t stage_bytes, int Kst, int t0, int lane)
{
    constexpr uint32_t LINE = NC * 4u;
    int stage = 0; uint32_t phase = 0;
    const uint32_t dst_line0 = ring + (uint32_t)A.msg_line0 * LINE;
    for (int it = 0; it < A.iters; it++) {
        const int sweeps = (ET && it + 1 < A.iters) ? 2 : 1;
        for (int sw = 0; sw < sweeps; sw++) {
            const bool want_msg = it > 0 || sw == 1;
            int line0 = 0;                                              // the row's first line in MSG (an edge number, or 4 x the row number)
            for (int c = 0; c < A.nb_deg; c++) {
                const int R = A.rows[c], ML = CMP ? 4 : A.deg[c];
                const CUtensorMap* const tm = &A.tm_msg[CMP ? 0 : c];
                const uint32_t bytes = want_msg ? (uint32_t)ML * LINE : 0u;
                for (int r = 0; r < R; r++, line0 += ML) {
                    const uint32_t full = bars + 8 * stage, dst = dst_line0 + (uint32_t)stage * stage_bytes;
                    // EVERY lane waits for the free slot (one warp instruction either way): a wait by lane 0 alone leaves the warp diverged
                    // for good, and every __syncwarp / shuffle after it then takes the compiler's WARPSYNC.COLLECTIVE path, 100-250 cycles
                    // apiece (profiles/r02_ncu_fs_small_v4.txt)
                    mbar_wait(bars + 8 * (Kst + stage), phase ^ 1u);             // slot free (passes at once on the first lap)
                    if (lane == 0) mbar_arrive_expect_tx(full, bytes);
                    if (TM) { if (lane == 0 && want_msg) tma_g2s_2d(dst, tm, t0, line0, full); }
                    else {
                        __syncwarp();
                        if (want_msg && lane < ML) bulk_g2s(dst + (uint32_t)lane * LINE, A.MSG + ((size_t)(line0 + lane) * A.T + t0), LINE, full);
                    }
                    if (++stage == Kst) { stage = 0; phase ^= 1u; }
                }
            }
            if (fs_producer_pass_end<ET>(sweeps, sw)) return;
        }
    }
}

// posterior side.  The row's edge words (pos2) are the one thing a producer has to LOAD before it can issue anything, and that load
// queues in the SM's L1TEX path behind every copy already in flight — a whole ring of rows: fetched one row ahead (round 2's first
// version) its latency WAS the producer's time per row, 578 ns, and with it the time per row of every batch too small to put two CTAs
// on an SM, whatever the consumers did.  The words now stream through three 32-word windows held across the warp's lanes (W0 = the
// window the current row starts in, W1, W2 = the next two, loaded 32..64 words = 5..9 rows before their first use); a row's D words
// are picked out of W0 / W1 with two shuffles.  pos2 is padded by FS_P2_PAD words so the windows may run past the last row.
// Everything that depends only on the edge words — the gather indices (four shuffles), the line count of the one-dimensional form —
// is computed BEFORE the wait for a free slot.  The words themselves are left behind the stage's lines for the consumers (with their
// flags), ordered before lane 0's arrive by the __syncwarp.
template <int NC, bool ET, bool CMP, bool G4>
__device__ __forceinline__ void fs_produce_posteriors(const FsArgs& A, uint32_t bars, uint32_t ring, uint32_t stage_bytes, int Kst, int t0, int lane)
{
    constexpr uint32_t LINE = NC * 4u;
    int stage = 0; uint32_t phase = 0;
    const uint32_t p2_off = FS_P2_OFFSET(A, NC) + 4u * (uint32_t)lane;
    for (int it = 0; it < A.iters; it++) {
        const int sweeps = (ET && it + 1 < A.iters) ? 2 : 1;
        for (int sw = 0; sw < sweeps; sw++) {
            const uint32_t* wp = A.pos2 + lane;
            uint32_t W0 = __ldg(wp), W1 = __ldg(wp + 32), W2 = __ldg(wp + 64);
            int off = 0;                                                // the current row's first word, relative to W0's first
            for (int c = 0; c < A.nb_deg; c++) {
                const int D = A.deg[c], R = A.rows[c];
                const int groups = (D + 3) >> 2;                        // gather4: lane g < groups fetches edges 4g .. 4g+3 (the last one repeated to fill the group)
                const int b = 4 * min(lane, groups - 1);
                const int s0 = b, s1 = min(b + 1, D - 1), s2 = min(b + 2, D - 1), s3 = min(b + 3, D - 1);
                for (int r = 0; r < R; r++) {
                    if (off >= 32) { W0 = W1; W1 = W2; wp += 32; W2 = __ldg(wp + 64); off -= 32; }
                    const int pos = off + lane;
                    const uint32_t w_lo = __shfl_sync(0xFFFFFFFFu, W0, pos & 31), w_hi = __shfl_sync(0xFFFFFFFFu, W1, pos & 31);
                    const uint32_t p2_row = lane < D ? (pos < 32 ? w_lo : w_hi) : FS_F_HAZARD;      // with its flags: what the consumers need
                    off += D;
                    const bool fetch = lane < D && (sw == 1 || !(p2_row & FS_F_HAZARD));     // one-dimensional form: this lane's line is wanted
                    const uint32_t idx = p2_row & FS_IDX_MASK;
                    int r0 = 0, r1 = 0, r2 = 0, r3 = 0;
                    uint32_t bytes;
                    if (G4) {
                        r0 = (int)__shfl_sync(0xFFFFFFFFu, idx, s0); r1 = (int)__shfl_sync(0xFFFFFFFFu, idx, s1);
                        r2 = (int)__shfl_sync(0xFFFFFFFFu, idx, s2); r3 = (int)__shfl_sync(0xFFFFFFFFu, idx, s3);
                        bytes = 4u * (uint32_t)groups * LINE;
                    } else bytes = (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, fetch)) * LINE;
                    const uint32_t full = bars + 8 * stage, dst0 = ring + (uint32_t)stage * stage_bytes;
                    mbar_wait(bars + 8 * (Kst + stage), phase ^ 1u);     // slot free (passes at once on the first lap); every lane: see fs_produce_msgs
                    if (lane < D) sts_u32(dst0 + p2_off, p2_row);
                    __syncwarp();                                        // the edge words, before lane 0's arrive (release)
                    if (lane == 0) mbar_arrive_expect_tx(full, bytes);
                    __syncwarp();                                        // the byte count is posted before any copy can complete
                    if (G4) { if (lane < groups) tma_gather4(dst0 + 4u * (uint32_t)lane * LINE, &A.tm_v, t0, r0, r1, r2, r3, full); }
                    else if (fetch) bulk_g2s(dst0 + (uint32_t)lane * LINE, A.V + ((size_t)idx * A.T + t0), LINE, full);
                    if (++stage == Kst) { stage = 0; phase ^= 1u; }
                }
            }
            if (fs_producer_pass_end<ET>(sweeps, sw)) return;
        }
    }
}

// MAXD: the largest row degree this instantiation carries (8 | FS_MAXDEG) — a kernel's register allocation is that of its widest row
// body, and DVB-S2 (degrees 7 and 6) should not pay for the degree-10 body of 1200x600
template <int SEM, int ALGO, int NC, int MAXD, bool ET, bool CMP = false, bool PIPE2 = false>
__global__ void __launch_bounds__(NC + FS_PRODUCER_THREADS, (PIPE2 || MAXD > FS_MAXDEG) ? 1 : (MAXD <= 8 ? 512 : 384) / NC) fs_decode_kernel(const __grid_constant__ FsArgs A)
{
    constexpr uint32_t LINE = NC * 4u;
    extern __shared__ __align__(128) unsigned char fs_smem[];
    // layout: full[K] | empty[K] | pad to 128 | forwarding ring | K stages of 2*max_deg lines
    const int Kst = A.stages;
    const uint32_t bars = smem_u32(fs_smem);
    const uint32_t fwd_s = bars + (uint32_t)((16 * Kst + 127) / 128 * 128);
    const uint32_t ring = fwd_s + FS_FWD * (uint32_t)A.max_deg * LINE;
    const uint32_t stage_bytes = (uint32_t)(A.msg_line0 + A.msg_lines) * LINE + FS_P2_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * NC;
    if (threadIdx.x == 0) {
        for (int k = 0; k < Kst; k++) { mbar_init(bars + 8 * k, 2); mbar_init(bars + 8 * (Kst + k), NC / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp >= NC / 32) {
        // ---------------- two producer warps: warp NC/32 fetches the posterior lines of every row, warp NC/32 + 1 its message lines.
        // Both arrive on the stage's "full" barrier (count 2) with their own byte counts.  (Four producer warps, the edges of a row dealt
        // out alternately to the two warps of a side, changed nothing — profiles/r02_sweep_fs_pw4.jsonl.)
        // Each side is its own tight loop, specialised on the copy form: with one consumer warp per scheduler a batch is only as fast as
        // the SLOWER producer walks the row list, and the first version's common loop (run-time side / form flags, class tables re-read
        // from constant memory every row, a ballot to count lines) was a serial chain of ~130 instructions = 1136 cycles per row —
        // exactly the 578 ns per row every small batch ran at, whatever the ring depth (profiles/r02_ncu_fs_small_v3.txt).
        if (warp > NC / 32) { if (A.use_tm) fs_produce_msgs<NC, ET, CMP, true>(A, bars, ring, stage_bytes, Kst, t0, lane); else fs_produce_msgs<NC, ET, CMP, false>(A, bars, ring, stage_bytes, Kst, t0, lane); }
        else                { if (A.use_g4) fs_produce_posteriors<NC, ET, CMP, true>(A, bars, ring, stage_bytes, Kst, t0, lane); else fs_produce_posteriors<NC, ET, CMP, false>(A, bars, ring, stage_bytes, Kst, t0, lane); }
        return;
    }

    // ---------------- consumers ---------------------------------------------------------------------------------------------
    const int tid = threadIdx.x, t = t0 + tid;
    RowConsts K; make_consts<SEM>(K, A.prm); K.c64 = A.exp_word;
    FsCursor cur{0, 0u, 0u, 0u};
    const int lo = (SEM == LDPC_SEM_GPU_FIXED) ? -128 : -A.prm.sat_var;
    uint32_t keep_lo = 0u, keep_hi = 0u;           // 0xFFFF per half = frame frozen (early-terminated)
    uint32_t done[4] = { 0u, 0u, 0u, 0u };
    int it = 0;
    for (; it < A.iters; it++) {
        size_t e = 0, rho = 0;
        for (int sg = 0; sg < A.nseg; sg++) {
            const int D = A.seg_deg[sg], R = A.seg_rows[sg], c = A.seg_cls[sg];
            const bool stair = A.seg_stair[sg] != 0;
            const bool quirk = SEM == LDPC_SEM_X86_SSE && ALGO == LDPC_ALGO_OMS && c >= 1;
            // the reference's OMS kernel forgets the 31-clamp for the second degree class in its peeled first iteration (CUDA_OMS_SIMD.cu:113-114)
            K.msg_c = (SEM == LDPC_SEM_GPU_FIXED && ALGO == LDPC_ALGO_OMS && it == 0 && c >= 1) ? K.one : K.msg;
#define FS_GO_(DD, FI, QQ, ST) fs_class<SEM, ALGO, DD, FI, QQ, NC, ET, CMP, ST, PIPE2>(A, tid, t, e, rho, R, K, lane, bars, ring, stage_bytes, fwd_s, Kst, cur, keep_lo, keep_hi)
#define FS_GO(DD, FI, QQ) do { if constexpr (DD >= 6 && DD <= 8) { if (stair) FS_GO_(DD, FI, QQ, true); else FS_GO_(DD, FI, QQ, false); } else FS_GO_(DD, FI, QQ, false); } while (0)
#define FS_CASE(DD)                                                                          \
    case DD:                                                                                 \
        if constexpr (DD <= MAXD) {                                                          \
        if constexpr (SEM == LDPC_SEM_X86_SSE && ALGO == LDPC_ALGO_OMS) {                    \
        if (it == 0) { if (quirk) FS_GO(DD, true, true); else FS_GO(DD, true, false); }      \
        else         { if (quirk) FS_GO(DD, false, true); else FS_GO(DD, false, false); }    \
        } else { if (it == 0) FS_GO(DD, true, false); else FS_GO(DD, false, false); }        \
        }                                                                                    \
        break;
            if constexpr (MAXD > FS_MAXDEG) {       // the two-pass row body for every row, whatever its degree
                static_assert(!CMP && !PIPE2, "the generic row body has neither compressed messages nor paired rows");
                if constexpr (SEM == LDPC_SEM_X86_SSE && ALGO == LDPC_ALGO_OMS) {
                    if (it == 0) { if (quirk) fs_class_generic<SEM, ALGO, true, true, NC, ET>(A, tid, t, e, D, R, K, lane, bars, ring, stage_bytes, fwd_s, Kst, cur, keep_lo, keep_hi);
                                   else fs_class_generic<SEM, ALGO, true, false, NC, ET>(A, tid, t, e, D, R, K, lane, bars, ring, stage_bytes, fwd_s, Kst, cur, keep_lo, keep_hi); }
                    else         { if (quirk) fs_class_generic<SEM, ALGO, false, true, NC, ET>(A, tid, t, e, D, R, K, lane, bars, ring, stage_bytes, fwd_s, Kst, cur, keep_lo, keep_hi);
                                   else fs_class_generic<SEM, ALGO, false, false, NC, ET>(A, tid, t, e, D, R, K, lane, bars, ring, stage_bytes, fwd_s, Kst, cur, keep_lo, keep_hi); }
                } else {
                    if (it == 0) fs_class_generic<SEM, ALGO, true, false, NC, ET>(A, tid, t, e, D, R, K, lane, bars, ring, stage_bytes, fwd_s, Kst, cur, keep_lo, keep_hi);
                    else fs_class_generic<SEM, ALGO, false, false, NC, ET>(A, tid, t, e, D, R, K, lane, bars, ring, stage_bytes, fwd_s, Kst, cur, keep_lo, keep_hi);
                }
                (void)stair; (void)rho;
            } else
            switch (D) { FS_CASE(3) FS_CASE(4) FS_CASE(5) FS_CASE(6) FS_CASE(7) FS_CASE(8) FS_CASE(9) FS_CASE(10) }
#undef FS_CASE
#undef FS_GO
#undef FS_GO_
        }
        if (ET && it + 1 < A.iters) {
            __threadfence();                           // once per iteration: this iteration's stores are performed ...
            fence_proxy_async_global();                // ... and visible to the producers' bulk copies of the stop-criterion pass
            __syncthreads();
            uint32_t b0, b1;
            fs_syndrome_sweep<NC, CMP>(A, tid, lane, K, lo, bars, ring, stage_bytes, Kst, cur, b0, b1);
            // frames that pass now and were not frozen before stop at iteration it + 1 (same bookkeeping as fp_decode_kernel)
            const uint32_t pass_lo = ~b0 & 0x80008000u, pass_hi = ~b1 & 0x80008000u;
            if ((pass_lo & 0x00008000u) && !done[0]) done[0] = it + 1;
            if ((pass_lo & 0x80000000u) && !done[1]) done[1] = it + 1;
            if ((pass_hi & 0x00008000u) && !done[2]) done[2] = it + 1;
            if ((pass_hi & 0x80000000u) && !done[3]) done[3] = it + 1;
            keep_lo = (done[0] ? 0x0000FFFFu : 0u) | (done[1] ? 0xFFFF0000u : 0u);
            keep_hi = (done[2] ? 0x0000FFFFu : 0u) | (done[3] ? 0xFFFF0000u : 0u);
            if (!__syncthreads_or(!(done[0] && done[1] && done[2] && done[3]))) { it++; break; }
        }
    }
    if (A.iters_done) {
#pragma unroll
        for (int k = 0; k < 4; k++) A.iters_done[4 * (size_t)t + k] = (uint8_t)((ET && done[k]) ? done[k] : it);
    }
}

}  // namespace ldpcb200
