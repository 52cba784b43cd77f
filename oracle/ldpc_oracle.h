/*
 * ldpc_oracle.h — CPU restatement of the reference decoders.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product path (ldpcgputegra_b200/, include/) may call, link or import this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker or the timed CPU baseline.
 *
 * Parity pinning: the X86_SSE / UNIFORM / ARM_SCALAR modes are checked bit-exactly (hard decisions,
 * posteriors, messages, iteration counts) against the reference's own decoders compiled from
 * /root/reference into oracle/_ref (tests/test_oracle_vs_ref.py, fixtures in tests/golden/).
 * GPU_FIXED is checked against the reference kernels built for sm_100a (oracle/_ref/libref_gpu.so) on the
 * GPU box.  The float flooding decoder has NO reference implementation: parity unpinned.
 */
#ifndef LDPC_ORACLE_H
#define LDPC_ORACLE_H
#include "../include/ldpc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* fixed-point decoder, all semantics modes; prm->schedule selects layered (every reference decoder) or flooding (own
 * definition, unpinned: no reference decoder floods).  llr/hard/post/msgs are frame-major; elem_bytes = 1 (int8) or 2 (int16)
 * selects the width of llr/post/msgs storage.  post, msgs, iters_done nullable.  Returns 0 or LDPC_ERR_*. */
int oracle_decode_fixed(const ldpc_code_t* code, const ldpc_params_t* prm,
                        const void* llr, uint8_t* hard, void* post, void* msgs, uint8_t* iters_done,
                        size_t frames, int iters, int elem_bytes);

/* float min-sum (MS / offset / normalised / 2-factor normalised), flooding or layered, optional syndrome early termination
 * (own definition; unpinned).  post [frames][n], msgs [frames][m] (check-to-variable), iters_done nullable. */
int oracle_decode_float(const ldpc_code_t* code, const ldpc_params_t* prm,
                        const float* llr, uint8_t* hard, float* post, float* msgs, uint8_t* iters_done,
                        size_t frames, int iters);

/* q = clamp((int)(scale*y), -sat, sat)  (ref: code/x86/CFixPointConversion/CFastFixConversion.cpp:55-65) */
void oracle_quantize(const float* y, int8_t* q, size_t count, int scale, int sat);

/* byte-per-bit -> LSB-first packed */
void oracle_pack_bits(const uint8_t* hard, uint8_t* packed, size_t frames, int n);

/* multi-threaded wrapper used as the timed CPU baseline when oracle/_ref is unavailable: OpenMP over frames */
int oracle_decode_fixed_mt(const ldpc_code_t* code, const ldpc_params_t* prm, const void* llr, uint8_t* hard,
                           size_t frames, int iters, int elem_bytes, int threads);

#ifdef __cplusplus
}
#endif
#endif
