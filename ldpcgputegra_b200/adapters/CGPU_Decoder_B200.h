// CGPU_Decoder_B200.h — header-only C++ adapters that put libldpc_b200.so behind the reference's decoder classes, so the
// reference harnesses link against the new decoder by changing one `new` expression.
//
//   CGPU_Decoder_B200  : same constructor and virtuals as CGPUDecoder and its CGPU_Decoder_{MS,OMS,NMS,2NMS}_SIMD children
//                        (ref: code/gpu_fixed/decoder_template/CGPUDecoder.h:20-37, code/gpu_fixed/decoder_oms/CGPU_Decoder_OMS_SIMD.h)
//   CDecoder_B200      : same setters/decode as CDecoder_{OMS,NMS}_fixed_SSE
//                        (ref: code/x86/CDecoder/template/CDecoder.h:28-40, CDecoder_fixed.h:30-44, OMS/CDecoder_OMS_fixed_SSE.h:26-39)
//
// They compile against the C ABI only (no CUDA headers needed).  Error behaviour mirrors the reference: print and exit(0)
// (ref: code/gpu_fixed/custom_api/custom_cuda.cu:5-17, CGPU_Decoder_OMS_SIMD.cu:101-105).  When included from inside the
// reference tree define LDPC_B200_DERIVE_FROM_REFERENCE before including, after the reference's own class header, and the
// adapters derive from CGPUDecoder / CDecoder_fixed so they fit the reference's factory and pointer types.  Each adapter derives
// only when its own base class has been declared (the two harness trees never include both): the reference headers' include
// guards tell (__CLASS_CGPUDecoder__, code/gpu_fixed/decoder_template/CGPUDecoder.h:14; __CDecoder_fixed__, code/x86/CDecoder/template/CDecoder_fixed.h).
#ifndef CGPU_DECODER_B200_H
#define CGPU_DECODER_B200_H

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../include/ldpc_b200.h"

namespace ldpc_b200_adapters {

inline void die(ldpc_handle h, int rc, const char* what)
{
    printf("(EE) ldpc_b200 %s failed: %s (%s)\n", what, ldpc_b200_status_string(rc), ldpc_b200_last_error(h));
    exit(0);
}

// Fill an ldpc_code_t from the reference's compile-time macros when the including TU has them (constantes_gpu.h / constantes_sse.h).
#if defined(_N) && defined(_K) && defined(_M) && defined(NB_DEGRES) && defined(DEG_1)
template <typename IndexT>
inline ldpc_code_t code_from_reference_macros(const IndexT* table, uint32_t* storage /* [_M] */)
{
    ldpc_code_t c; memset(&c, 0, sizeof(c));
    c.n = _N; c.n_checks = _K; c.m = _M; c.nb_deg = NB_DEGRES;
    c.deg[0] = DEG_1; c.rows[0] = DEG_1_COMPUTATIONS;
#if NB_DEGRES >= 2
    c.deg[1] = DEG_2; c.rows[1] = DEG_2_COMPUTATIONS;
#endif
#if NB_DEGRES >= 3
    c.deg[2] = DEG_3; c.rows[2] = DEG_3_COMPUTATIONS;
#endif
#if NB_DEGRES >= 4
    c.deg[3] = DEG_4; c.rows[3] = DEG_4_COMPUTATIONS;
#endif
#if NB_DEGRES >= 5
    c.deg[4] = DEG_5; c.rows[4] = DEG_5_COMPUTATIONS;
#endif
    for (int i = 0; i < _M; i++) storage[i] = (uint32_t)table[i];
    c.pos = storage;
    return c;
}
#endif

}  // namespace ldpc_b200_adapters

// ---------------------------------------------------------------------------------------------------------------------
// GPU flavour.  nb_frames counts GPU THREADS of the reference = 4 frames each (ref: CGPUDecoder.cpp:14-22), so one decode()
// call moves 4*nb_frames frames: var_nodes really holds int8 [4*nb_frames][n], Rprime_fix receives bytes in {0,1}
// (ref: CGPU_Decoder_OMS_SIMD.cu:111,146).
// ---------------------------------------------------------------------------------------------------------------------
#if defined(LDPC_B200_DERIVE_FROM_REFERENCE) && defined(__CLASS_CGPUDecoder__)
#define LDPC_B200_DERIVE_GPU 1
#endif
#if defined(LDPC_B200_DERIVE_FROM_REFERENCE) && (defined(__CDecoder_fixed__) || defined(LDPC_B200_HAVE_CDECODER_FIXED))
#define LDPC_B200_DERIVE_X86 1
#endif

class CGPU_Decoder_B200
#ifdef LDPC_B200_DERIVE_GPU
    : public CGPUDecoder
#endif
{
public:
    // algo: "MS" | "OMS" | "NMS" | "2NMS" — the reference picks the subclass from argv (code/gpu_fixed/test.cpp:241-276)
    CGPU_Decoder_B200(size_t _nb_frames, size_t n, size_t k, size_t m, const ldpc_code_t& code, const char* algo = "OMS", int device = 0)
#ifdef LDPC_B200_DERIVE_GPU
        : CGPUDecoder(0, n, k, m)
#endif
    { init(_nb_frames, n, k, m, code, algo, device); }

#if defined(_N) && defined(_K) && defined(_M) && defined(NB_DEGRES) && defined(DEG_1)
    // The reference's own constructor shape (ref: CGPU_Decoder_OMS_SIMD(nb_frames, n, k, m), code/gpu_fixed/decoder_oms/CGPU_Decoder_OMS_SIMD.h:7):
    // inside the reference tree the code table is the compiled-in PosNoeudsVariable[_M] (matrix/<code>/constantes_decoder.h, declared by
    // matrix/<code>/constantes_gpu.h:39), so the only line a harness changes is the `new` expression itself.
    CGPU_Decoder_B200(size_t _nb_frames, size_t n, size_t k, size_t m, const char* algo = "OMS", int device = 0)
#ifdef LDPC_B200_DERIVE_GPU
        : CGPUDecoder(0, n, k, m)
#endif
    {
        static uint32_t table[_M];
        const ldpc_code_t code = ldpc_b200_adapters::code_from_reference_macros(PosNoeudsVariable, table);
        init(_nb_frames, n, k, m, code, algo, device);
    }
#endif

private:
    void init(size_t _nb_frames, size_t n, size_t k, size_t m, const ldpc_code_t& code, const char* algo, int device)
    {
        if ((size_t)code.n != n || (size_t)code.n_checks != k || (size_t)code.m != m) { printf("(EE) code table does not match (n,k,m)\n"); exit(0); }
        frames_ = 4 * _nb_frames;
        ldpc_params_t p; ldpc_b200_default_params(&p);
        p.semantics = LDPC_SEM_GPU_FIXED;
        p.algo = !strcmp(algo, "MS") ? LDPC_ALGO_MS : !strcmp(algo, "NMS") ? LDPC_ALGO_NMS : !strcmp(algo, "2NMS") ? LDPC_ALGO_2NMS : LDPC_ALGO_OMS;
        if (p.algo == LDPC_ALGO_2NMS) p.early_term = LDPC_ET_NONE;   // the reference's per-thread break relies on UB; opt in explicitly
        int rc = ldpc_b200_create(&h_, &code, &p, device, frames_);
        if (rc) ldpc_b200_adapters::die(nullptr, rc, "create");
    }

public:
    virtual ~CGPU_Decoder_B200() { ldpc_b200_destroy(h_); }
    virtual void initialize() {}
    virtual void decode(float var_nodes[], int Rprime_fix[], int nombre_iterations)
    {
        int rc = ldpc_b200_decode(h_, var_nodes, (uint8_t*)Rprime_fix, frames_, nombre_iterations, nullptr);
        if (rc) ldpc_b200_adapters::die(h_, rc, "decode");
    }
    // the reference's streamed variant creates a stream per call and still copies synchronously (CGPU_Decoder_MS_SIMD.cu:219-275);
    // here it is the asynchronous slot API: call sync() before reading Rprime_fix.
    virtual void decode_stream(float var_nodes[], int Rprime_fix[], int nombre_iterations)
    {
        int rc = ldpc_b200_decode_async(h_, next_slot_, var_nodes, (uint8_t*)Rprime_fix, frames_, nombre_iterations, nullptr);
        if (rc) ldpc_b200_adapters::die(h_, rc, "decode_stream");
        next_slot_ = (next_slot_ + 1) % 4;
    }
    void sync() { int rc = ldpc_b200_sync(h_, -1); if (rc) ldpc_b200_adapters::die(h_, rc, "sync"); }
    ldpc_handle handle() const { return h_; }

private:
    ldpc_handle h_ = nullptr;
    size_t frames_ = 0;
    int next_slot_ = 0;
};

// ---------------------------------------------------------------------------------------------------------------------
// CPU flavour of the boundary: setters first, then decode(char*, char*, iters) on `frames` frames per call (the SSE
// reference decodes 16: code/x86/CDecoder/OMS/CDecoder_OMS_fixed_SSE.cpp:140-149).  The decoder is created lazily at the
// first decode so that the setters behave like the reference's.
// ---------------------------------------------------------------------------------------------------------------------
class CDecoder_B200
#ifdef LDPC_B200_DERIVE_X86
    : public CDecoder_fixed
#endif
{
public:
    CDecoder_B200(const ldpc_code_t& code, const char* type /* "OMS" | "NMS" */, size_t frames = 16, int device = 0)
        : code_(code), frames_(frames), device_(device)
    {
        ldpc_b200_default_params(&p_);
        p_.semantics = LDPC_SEM_X86_SSE;
        p_.algo = !strcmp(type, "NMS") ? LDPC_ALGO_NMS : LDPC_ALGO_OMS;
        offset_set_ = false;
    }
    virtual ~CDecoder_B200() { ldpc_b200_destroy(h_); }
    virtual void setSigmaChannel(float) {}
    virtual void setNumberOfIterations(int v) { nb_iters_ = v; }
    virtual void setOffset(int o)
    {   // (ref: CDecoder_OMS_fixed_SSE.cpp:101-109) — configuring twice is an error there too
        if (offset_set_) { printf("(EE) Offset value was already configured (%d)\n", p_.offset); exit(0); }
        p_.offset = o; offset_set_ = true;
    }
    virtual void setFactor(int f) { p_.factor_q5 = f; }
    virtual void setVarRange(int vmin, int vmax) { (void)vmin; p_.sat_var = vmax; }
    virtual void setMsgRange(int mmin, int mmax) { (void)mmin; p_.sat_msg = mmax; }
    virtual void decode(char var_nodes[], char Rprime_fix[], int nombre_iterations)
    {
        if (p_.sat_var != 127) exit(0);   // (ref: CDecoder_OMS_fixed_SSE.cpp:114-120)
        if (!h_) { int rc = ldpc_b200_create(&h_, &code_, &p_, device_, frames_); if (rc) ldpc_b200_adapters::die(nullptr, rc, "create"); }
        int rc = ldpc_b200_decode(h_, var_nodes, (uint8_t*)Rprime_fix, frames_, nombre_iterations, nullptr);
        if (rc) ldpc_b200_adapters::die(h_, rc, "decode");
    }
    virtual void decode(float[], char[], int) {}   // compatibility no-op, as in the reference (CDecoder_fixed_SSE.cpp:35-40)

private:
    ldpc_code_t code_;
    ldpc_params_t p_;
    ldpc_handle h_ = nullptr;
    size_t frames_;
    int device_, nb_iters_ = 0;
    bool offset_set_;
};

#endif  // CGPU_DECODER_B200_H
