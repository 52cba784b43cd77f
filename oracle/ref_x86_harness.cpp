/*
 * ref_x86_harness.cpp — C-callable shell around the REFERENCE's own x86 SSE decoders.
 * TEST INFRASTRUCTURE ONLY.  This file contains none of the reference's code: it is compiled together with the
 * reference sources where they lie under /root/reference (see oracle/Makefile), once per code table, into
 * oracle/_ref/libref_x86_<code>.so.  It is what pins the oracle (tests/test_oracle_vs_ref.py, tools/gen_golden.py)
 * and what bench.py times as the CPU baseline (kind "reference").
 *
 * Reference entry points used: CDecoder_OMS_fixed_SSE / CDecoder_NMS_fixed_SSE (setOffset/setFactor, setVarRange,
 * setMsgRange, decode(char*,char*,int)) exactly as CreateDecoder() does (code/x86/CDecoder/DecoderLibrary.h:44-134).
 */
#include <string>
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include <chrono>
#include <thread>
#include <atomic>
#include <vector>
using namespace std;
#include "CDecoder/OMS/CDecoder_OMS_fixed_SSE.h"
#include "CDecoder/NMS/CDecoder_NMS_fixed_SSE.h"

namespace {
struct OmsProbe : public CDecoder_OMS_fixed_SSE { const char* vn() const { return (const char*)var_nodes; } const char* vm() const { return (const char*)var_mesgs; } };
struct NmsProbe : public CDecoder_NMS_fixed_SSE { const char* vn() const { return (const char*)var_nodes; } const char* vm() const { return (const char*)var_mesgs; } };

const int kDeg[] = {
    DEG_1
#if NB_DEGRES >= 2
    , DEG_2
#endif
#if NB_DEGRES >= 3
    , DEG_3
#endif
#if NB_DEGRES >= 4
    , DEG_4
#endif
#if NB_DEGRES >= 5
    , DEG_5
#endif
};
const int kRows[] = {
    DEG_1_COMPUTATIONS
#if NB_DEGRES >= 2
    , DEG_2_COMPUTATIONS
#endif
#if NB_DEGRES >= 3
    , DEG_3_COMPUTATIONS
#endif
#if NB_DEGRES >= 4
    , DEG_4_COMPUTATIONS
#endif
#if NB_DEGRES >= 5
    , DEG_5_COMPUTATIONS
#endif
};

CDecoder* make(int algo, int param, OmsProbe** op, NmsProbe** np)
{
    *op = nullptr; *np = nullptr;
    if (algo == 1) { OmsProbe* d = new OmsProbe(); d->setOffset(param); d->setVarRange(-127, 127); d->setMsgRange(-31, 31); *op = d; return d; }
#if NB_DEGRES <= 2
    if (algo == 2) { NmsProbe* d = new NmsProbe(); d->setFactor(param); d->setVarRange(-127, 127); d->setMsgRange(-31, 31); *np = d; return d; }
#endif
    return nullptr;
}
}  // namespace

extern "C" {

/* out[0..3] = _N, _K(#checks), _M, NB_DEGRES ; deg/rows filled with up to 8 entries */
void ref_x86_info(int* out4, int* deg, int* rows)
{
    out4[0] = _N; out4[1] = _K; out4[2] = _M; out4[3] = NB_DEGRES;
    for (int i = 0; i < NB_DEGRES; i++) { deg[i] = kDeg[i]; rows[i] = kRows[i]; }
}

void ref_x86_table(uint32_t* pos) { for (int i = 0; i < _M; i++) pos[i] = PosNoeudsVariable[i]; }

/* frames must be a multiple of 16 (one reference call = 16 frames). post/msgs nullable, frame-major int8. */
int ref_x86_decode(int algo, int param, const int8_t* llr, uint8_t* hard, int8_t* post, int8_t* msgs, size_t frames, int iters)
{
    if (frames % 16) return -1;
    OmsProbe* op; NmsProbe* np;
    CDecoder* dec = make(algo, param, &op, &np);
    if (!dec) return -6;
    char *in, *out;
    if (posix_memalign((void**)&in, 64, 16 * _N) || posix_memalign((void**)&out, 64, 16 * _N)) return -5;
    for (size_t b = 0; b < frames / 16; b++) {
        memcpy(in, llr + b * 16 * _N, 16 * _N);
        dec->decode(in, out, iters);
        memcpy(hard + b * 16 * _N, out, 16 * _N);
        const char* vn = op ? op->vn() : np->vn();
        const char* vm = op ? op->vm() : np->vm();
        /* internal layout var_nodes[16*n + f] (code/x86/CDecoder/OMS/CDecoder_OMS_fixed_SSE.cpp:143-148) */
        if (post) for (int f = 0; f < 16; f++) for (int n = 0; n < _N; n++) post[(b * 16 + f) * _N + n] = vn[16 * n + f];
        if (msgs) for (int f = 0; f < 16; f++) for (int e = 0; e < _M; e++) msgs[(b * 16 + f) * (size_t)_M + e] = vm[16 * e + f];
    }
    free(in); free(out);
    delete dec;
    return 0;
}

/* CPU baseline: one decoder object per thread (objects own their state), each looping decode() over its slice of 16-frame
 * blocks — the reference's PERF loop (code/x86/main_p.cpp:673-687) with T = threads instead of its cap of 4.
 * Decoders and staging buffers are created before the clock starts.  Returns seconds spent decoding, <0 on error. */
double ref_x86_decode_mt(int algo, int param, const int8_t* llr, uint8_t* hard, size_t frames, int iters, int threads)
{
    if (frames % 16 || threads < 1) return -1.0;
    const size_t blocks = frames / 16;
    const bool direct = (((uintptr_t)llr | (uintptr_t)hard) % 16 == 0) && (_N % 16 == 0);
    struct Slot { CDecoder* dec; char* in; char* out; };
    vector<Slot> slots(threads);
    bool bad = false;
    for (auto& s : slots) {
        OmsProbe* op; NmsProbe* np;
        s.dec = make(algo, param, &op, &np); s.in = s.out = nullptr;
        if (!s.dec || posix_memalign((void**)&s.in, 64, 16 * _N) || posix_memalign((void**)&s.out, 64, 16 * _N)) bad = true;
    }
    double seconds = -1.0;
    if (!bad) {
        atomic<int> ready(0); atomic<bool> go(false);
        vector<thread> pool;
        for (int t = 0; t < threads; t++)
            pool.emplace_back([&, t] {
                Slot& s = slots[t];
                ready.fetch_add(1);
                while (!go.load(memory_order_acquire)) this_thread::yield();
                const size_t b0 = blocks * t / threads, b1 = blocks * (t + 1) / threads;
                for (size_t b = b0; b < b1; b++) {
                    if (direct) s.dec->decode((char*)llr + b * 16 * _N, (char*)hard + b * 16 * _N, iters);
                    else { memcpy(s.in, llr + b * 16 * _N, 16 * _N); s.dec->decode(s.in, s.out, iters); memcpy(hard + b * 16 * _N, s.out, 16 * _N); }
                }
            });
        while (ready.load() < threads) this_thread::yield();
        auto t0 = chrono::steady_clock::now();
        go.store(true, memory_order_release);
        for (auto& th : pool) th.join();
        seconds = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
    }
    for (auto& s : slots) { free(s.in); free(s.out); delete s.dec; }
    return seconds;
}

}  // extern "C"
