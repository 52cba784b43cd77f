// adapter_harness.cpp — a miniature of the reference's harness loop (ref: code/gpu_fixed/main.cpp:212-268, code/x86/main_p.cpp:382-385,
// 483-486) written against the header-only adapters: builds a decoder with the reference's constructor shapes, decodes the frames of
// a file and writes the hard decisions.  usage: adapter_harness <code.ldpc> <gpu|x86> <llr.bin> <out.bin> <frames> <iters>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "CGPU_Decoder_B200.h"

int main(int argc, char** argv)
{
    if (argc != 7) { printf("usage\n"); return 2; }
    ldpc_code_t code;
    if (ldpc_b200_load_code_table(&code, argv[1])) { printf("(EE) cannot load %s\n", argv[1]); return 2; }
    const size_t frames = (size_t)atoll(argv[5]); const int iters = atoi(argv[6]);
    std::vector<char> llr(frames * code.n), out(frames * code.n);
    FILE* f = fopen(argv[3], "rb"); if (!f || fread(llr.data(), 1, llr.size(), f) != llr.size()) { printf("(EE) llr file\n"); return 2; } fclose(f);
    if (argv[2][0] == 'g') {
        // reference: new CGPU_Decoder_OMS_SIMD(NB_THREAD_ON_GPU, _N, _K, _M); nb_frames counts threads of 4 frames
        CGPU_Decoder_B200 dec(frames / 4, code.n, code.n_checks, code.m, code, "OMS");
        dec.initialize();
        dec.decode((float*)llr.data(), (int*)out.data(), iters);
        dec.decode_stream((float*)llr.data(), (int*)out.data(), iters); dec.sync();
    } else {
        // reference: CreateDecoder("OMS", "sse", "fixed", ...) -> setOffset / setVarRange / setMsgRange, then decode(char*, char*, iters)
        CDecoder_B200 dec(code, "OMS", frames);
        dec.setOffset(1); dec.setVarRange(-127, 127); dec.setMsgRange(-31, 31);
        dec.decode((float*)nullptr, out.data(), iters);          // the reference's no-op float overload
        dec.decode(llr.data(), out.data(), iters);
    }
    f = fopen(argv[4], "wb"); fwrite(out.data(), 1, out.size(), f); fclose(f);
    printf("decoded %zu frames\n", frames);
    return 0;
}
