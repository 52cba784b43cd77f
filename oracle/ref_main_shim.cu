/*
 * ref_main_shim.cu — puts libldpc_b200.so behind the REFERENCE's own simulator, unmodified.
 * TEST INFRASTRUCTURE ONLY; contains none of the reference's code.
 *
 * code/gpu_fixed/main.cpp picks its decoder with `new CGPU_Decoder_{MS,OMS,NMS,2NMS}_SIMD(NB_THREAD_ON_GPU, _N, _K, _M)` /
 * `new CGPU_Decoder_MS_SIMD_v2(...)` (main.cpp:212-229) and calls decoder->decode(...) (main.cpp:268).  Those classes are DECLARED in
 * decoder_<x>/CGPU_Decoder_<x>_SIMD.h and DEFINED in the matching .cu files.  oracle/Makefile compiles main.cpp and every non-decoder
 * source of the reference where they lie (channel, counters, terminal, timers, buffers, code table, CGPUDecoder base class) and
 * links THIS file in place of the five decoder .cu files: it defines the same member functions, each forwarding to the header-only
 * adapter CGPU_Decoder_B200 (ldpcgputegra_b200/adapters) and through it to the C ABI.  Not one line of the reference changes.
 */
#include <map>
#include <cstdio>
#include "decoder_ms/CGPU_Decoder_MS_SIMD.h"
#include "decoder_oms/CGPU_Decoder_OMS_SIMD.h"
#include "decoder_nms/CGPU_Decoder_NMS_SIMD.h"
#include "decoder_2nms/CGPU_Decoder_2NMS_SIMD.h"
#include "decoder_oms_v2/CGPU_Decoder_MS_SIMD_v2.h"
#include "CGPU_Decoder_B200.h"          /* the 4-argument constructor: table from the compiled-in PosNoeudsVariable */

namespace {
std::map<const void*, CGPU_Decoder_B200*>& live() { static std::map<const void*, CGPU_Decoder_B200*> m; return m; }
CGPU_Decoder_B200* make(const void* self, size_t nb, size_t n, size_t k, size_t m, const char* algo)
{
    printf("(II) decoder behind this harness : libldpc_b200 (%s, GPU_FIXED semantics)\n", algo);
    return live()[self] = new CGPU_Decoder_B200(nb, n, k, m, algo);
}
void drop(const void* self) { delete live()[self]; live().erase(self); }
}

#define SHIM(CLS, ALGO)                                                                                              \
    CLS::CLS(size_t nb, size_t n, size_t k, size_t m) : CGPUDecoder(nb, n, k, m) { make(this, nb, n, k, m, ALGO); } \
    CLS::~CLS() { drop(this); }                                                                                      \
    void CLS::initialize() {}                                                                                        \
    void CLS::decode(float var_nodes[_N], int Rprime_fix[_N], int nombre_iterations) { live()[this]->decode(var_nodes, Rprime_fix, nombre_iterations); }

SHIM(CGPU_Decoder_MS_SIMD, "MS")
SHIM(CGPU_Decoder_OMS_SIMD, "OMS")
SHIM(CGPU_Decoder_NMS_SIMD, "NMS")
SHIM(CGPU_Decoder_2NMS_SIMD, "2NMS")
SHIM(CGPU_Decoder_MS_SIMD_v2, "OMS")          /* the experimental oms_v2 kernels share the OMS arithmetic */
void CGPU_Decoder_MS_SIMD::decode_testStream(float a[4000], int b[4000], int it) { live()[this]->decode(a, b, it); }
void CGPU_Decoder_MS_SIMD::decode_stream(float a[4000], int b[4000], int it) { live()[this]->decode(a, b, it); }
