/*
 * ref_arm_harness.cpp — C-callable shell around the REFERENCE's scalar decoder of the ARM tree
 * (code/ldpc_decoder_arm/CDecoder/OMS/CDecoder_OMS_fixed_x86.cpp:61-200), the only reference decoder with a
 * syndrome stop criterion and run-time saturation rails.  TEST INFRASTRUCTURE ONLY; contains none of the reference's
 * code.  Built by oracle/Makefile into oracle/_ref/libref_arm_<code>.so with the x86/gpu-order code table forced in
 * (-include), because the ARM tree's own 576x288 table uses another row order.
 *
 * The reference never reports how many iterations it ran.  A decode that stopped at iteration s leaves the same state
 * for every cap >= s, so the harness re-runs with caps 1..iters-1 and takes the smallest cap whose state (posteriors and
 * messages) equals the final one.  The one ambiguous case — a decode that never stopped but sat on a fixed point — is
 * resolved by evaluating the stop criterion on the final state (satisfied() below, the harness's only own arithmetic).
 */
#include <string>
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include <vector>
using namespace std;
#include "CDecoder/OMS/CDecoder_OMS_fixed_x86.h"

namespace {
struct Probe : public CDecoder_OMS_fixed_x86 { const short* vn() const { return var_nodes; } const short* vm() const { return var_mesgs; } };

const int kDeg[]  = { DEG_1
#if NB_DEGRES >= 2
    , DEG_2
#endif
};
const int kRows[] = { DEG_1_COMPUTATIONS
#if NB_DEGRES >= 2
    , DEG_2_COMPUTATIONS
#endif
};

bool satisfied(const short* v, const short* m, int sat_var)
{
    int e = 0;
    for (int c = 0; c < (int)(sizeof(kDeg) / sizeof(kDeg[0])); c++)
        for (int r = 0; r < kRows[c]; r++) {
            int par = 0;
            for (int j = 0; j < kDeg[c]; j++, e++) {
                int x = v[PosNoeudsVariable[e]] - m[e];
                x = x < -sat_var ? -sat_var : (x > sat_var ? sat_var : x);
                par ^= (x > 0);
            }
            if (par) return false;
        }
    return true;
}
}

extern "C" {

void ref_arm_info(int* out4) { out4[0] = _N; out4[1] = _K; out4[2] = _M; out4[3] = NB_DEGRES; }

/* one frame per reference call. post/msgs are int16 frame-major (the reference stores shorts), nullable. */
int ref_arm_decode(int offset, int sat_var, int sat_msg, int early_term, const int8_t* llr, uint8_t* hard,
                   int16_t* post, int16_t* msgs, uint8_t* iters_done, size_t frames, int iters)
{
#if NB_DEGRES > 2
    return -6;
#else
    Probe dec;
    dec.setOffset(offset); dec.setVarRange(-sat_var, sat_var); dec.setMsgRange(-sat_msg, sat_msg);
    vector<signed char> in(_N), out(_N);
    vector<short> fin_v(_N), fin_m(_M);
    for (size_t f = 0; f < frames; f++) {
        memcpy(in.data(), llr + f * _N, _N);
        dec.setEarlyTerm(early_term != 0);
        dec.decode(in.data(), out.data(), iters);
        for (int i = 0; i < _N; i++) hard[f * _N + i] = (uint8_t)out[i];
        memcpy(fin_v.data(), dec.vn(), sizeof(short) * _N);
        memcpy(fin_m.data(), dec.vm(), sizeof(short) * _M);
        if (post) memcpy(post + f * _N, fin_v.data(), sizeof(short) * _N);
        if (msgs) memcpy(msgs + f * (size_t)_M, fin_m.data(), sizeof(short) * _M);
        if (iters_done) {
            int done = iters;
            if (early_term && satisfied(fin_v.data(), fin_m.data(), sat_var)) {
                for (int k = 1; k < iters; k++) {
                    dec.decode(in.data(), out.data(), k);
                    if (!memcmp(fin_v.data(), dec.vn(), sizeof(short) * _N) && !memcmp(fin_m.data(), dec.vm(), sizeof(short) * _M)) { done = k; break; }
                }
            }
            iters_done[f] = (uint8_t)done;
        }
    }
    return 0;
#endif
}

}  // extern "C"
