// encoder.cu — systematic encoder derived from the parity-check table itself, on the GPU.
//
// The reference simulates with the all-zero codeword (CFakeEncoder, ref: code/x86/CEncoder/CFakeEncoder.cpp:23-36) or, with
// `-encoder`, with a DVB-S2 IRA encoder driven by a second, hand-made table (GenericEncoder::encode, ref:
// code/x86/CEncoder/GenericEncoder.cpp:38-78: information bits first, parity accumulated through GenericEncoderTable.h, then the
// staircase p[i] ^= p[i-1]).  Here the encoder needs no second table: the information bits are the first n - n_checks positions
// (the convention of the reference's error counters, CErrorAnalyzer.cpp:129-137) and the parity positions are solved from
// H c = 0 —
//   1. peeling: any check with exactly one unknown parity position determines it (an IRA / staircase code is solved entirely
//      this way, in chain order — for DVB-S2 that is the reference encoder's accumulate-then-staircase, by uniqueness of the
//      systematic codeword);
//   2. what peeling cannot reach (the weight-3 column of the 802.16e / 802.11n dual-diagonal part, unstructured codes) is solved
//      by a dense GF(2) inverse of the remaining square system, computed once on the host;
// On the device the codeword is bit-sliced: one 32-bit word carries the same position of 32 frames, so a parity equation is a
// run of XORs; one thread encodes 32 frames.
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ldpc_b200.h"
#include "channel.cuh"

namespace {

struct Phase {            // either a list of peeling steps or one dense solve
    int dense = 0;
    // peeling: step s sets var target[s] = XOR of src[ptr[s] .. ptr[s+1])
    std::vector<int32_t> ptr, src, target;
    // dense: unknowns uvar[r]; equation rows' known-variable lists (eq_ptr/eq_src, r rows) give b; x_i = XOR_{j in inv row i} b_j
    std::vector<int32_t> uvar, eq_ptr, eq_src, inv_ptr, inv_idx;
};

struct DevPhase { int dense, n_steps; int32_t *ptr, *src, *target, *uvar, *eq_ptr, *eq_src, *inv_ptr, *inv_idx; int r; };

}  // namespace

struct ldpc_b200_encoder_s {
    int device = 0, n = 0, k = 0, max_r = 0;
    std::vector<DevPhase> phases;
    std::vector<void*> allocs;
    uint32_t* d_words = nullptr; size_t words_cap = 0;      // bit-sliced codewords [n][W]
    uint32_t* d_tmp = nullptr; size_t tmp_cap = 0;          // dense-solve right-hand sides [max_r][W]
    std::string err;
};

namespace {

thread_local std::string g_enc_error;

// ---- device kernels ---------------------------------------------------------------------------------------------------------
// frame-major bytes [F][k] (or random bits when info == nullptr) -> bit-sliced words [n][W] (positions >= k cleared)
__global__ void enc_pack_kernel(const uint8_t* __restrict__ info, uint32_t* __restrict__ words, size_t frames, int n, int k, int W,
                                uint64_t seed, uint64_t first_frame)
{
    const size_t total = (size_t)n * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int pos = (int)(i / W), w = (int)(i % W);
        uint32_t v = 0u;
        if (pos < k) {
            if (info) {
                for (int b = 0; b < 32; b++) { const size_t f = (size_t)w * 32 + b; if (f < frames && info[f * (size_t)k + pos]) v |= 1u << b; }
            } else {    // counter-based random information bits: word (pos, w) of (seed, first_frame) is reproducible anywhere
                const uint64_t g = first_frame / 32 + (uint64_t)w;
                uint32_t c[4] = { (uint32_t)g, (uint32_t)(g >> 32), (uint32_t)pos, 0x454E4331u };
                ldpcb200::philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
                v = c[0];
            }
        }
        words[i] = v;
    }
}

__global__ void enc_peel_kernel(uint32_t* __restrict__ words, int W, int n_steps, const int32_t* __restrict__ ptr, const int32_t* __restrict__ src,
                                const int32_t* __restrict__ target)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    for (int s = 0; s < n_steps; s++) {
        uint32_t v = 0u;
        for (int q = __ldg(ptr + s); q < __ldg(ptr + s + 1); q++) v ^= words[(size_t)__ldg(src + q) * W + w];
        words[(size_t)__ldg(target + s) * W + w] = v;
    }
}

__global__ void enc_dense_rhs_kernel(const uint32_t* __restrict__ words, uint32_t* __restrict__ tmp, int W, int r, const int32_t* __restrict__ eq_ptr,
                                     const int32_t* __restrict__ eq_src)
{
    const size_t total = (size_t)r * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(i / W), w = (int)(i % W);
        uint32_t v = 0u;
        for (int q = __ldg(eq_ptr + j); q < __ldg(eq_ptr + j + 1); q++) v ^= words[(size_t)__ldg(eq_src + q) * W + w];
        tmp[i] = v;
    }
}

__global__ void enc_dense_solve_kernel(uint32_t* __restrict__ words, const uint32_t* __restrict__ tmp, int W, int r, const int32_t* __restrict__ uvar,
                                       const int32_t* __restrict__ inv_ptr, const int32_t* __restrict__ inv_idx)
{
    const size_t total = (size_t)r * W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int u = (int)(i / W), w = (int)(i % W);
        uint32_t v = 0u;
        for (int q = __ldg(inv_ptr + u); q < __ldg(inv_ptr + u + 1); q++) v ^= tmp[(size_t)__ldg(inv_idx + q) * W + w];
        words[(size_t)__ldg(uvar + u) * W + w] = v;
    }
}

// bit-sliced words -> frame-major bytes [F][n] in {0,1}
__global__ void enc_unpack_kernel(const uint32_t* __restrict__ words, uint8_t* __restrict__ out, size_t frames, int n, int W)
{
    const size_t total = frames * (size_t)n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f = i / n; const int pos = (int)(i % n);
        out[i] = (uint8_t)((words[(size_t)pos * W + f / 32] >> (f % 32)) & 1u);
    }
}

// BPSK bit 0 -> -1, bit 1 -> +1 (ref: code/x86/CChanel/CChanelAWGN_MKL.cpp:129-139), y = s + sigma*n, then the handle's input type
template <class S>
__global__ void awgn_codeword_kernel(S* __restrict__ q, const uint8_t* __restrict__ bits, size_t frames, int n, float sigma, uint64_t seed,
                                     uint64_t first_frame, float scale, int sat)
{
    const int quads = (n + 3) / 4;
    const size_t total = frames * (size_t)quads;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t f = i / quads;
        const int p = (int)(i % quads) * 4;
        const uint64_t gf = first_frame + f;
        uint32_t c[4] = { (uint32_t)gf, (uint32_t)(gf >> 32), (uint32_t)p, 0x4C445043u };      // the same noise as awgn_kernel for (seed, frame, position)
        ldpcb200::philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        float g[4];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const float u1 = ((float)(c[2 * h] >> 8) + 0.5f) * (1.0f / 16777216.0f);
            const float u2 = ((float)(c[2 * h + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
            const float r = sqrtf(-2.0f * __logf(u1));
            float sn, cs; __sincosf(6.283185307179586f * u2, &sn, &cs);
            g[2 * h] = r * cs; g[2 * h + 1] = r * sn;
        }
#pragma unroll
        for (int b = 0; b < 4; b++) {
            if (p + b < n) {
                const float y = (bits[f * (size_t)n + p + b] ? 1.0f : -1.0f) + sigma * g[b];
                q[f * (size_t)n + p + b] = ldpcb200::awgn_out<S>(y, scale, sat);
            }
        }
    }
}

// errors against a reference codeword over the first k_info positions; one warp per frame
__global__ void count_errors_ref_kernel(const uint8_t* __restrict__ hard, const uint8_t* __restrict__ ref, size_t frames, int n, int k_info, int packed,
                                        unsigned long long* out)
{
    const int lane = threadIdx.x & 31;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int nb = (n + 7) / 8;
    unsigned long long be = 0, fe = 0;
    for (size_t f = warp; f < frames; f += nwarps) {
        int e = 0;
        for (int i = lane; i < k_info; i += 32) {
            const int bit = packed ? ((hard[f * (size_t)nb + i / 8] >> (i % 8)) & 1) : (hard[f * (size_t)n + i] != 0);
            e += bit != (ref[f * (size_t)n + i] != 0);
        }
        for (int o = 16; o; o >>= 1) e += __shfl_xor_sync(0xFFFFFFFFu, e, o);
        be += (unsigned long long)e; fe += e != 0;
    }
    if (lane == 0 && (be | fe)) { atomicAdd(out, be); atomicAdd(out + 1, fe); }
}

// ---- host analysis ------------------------------------------------------------------------------------------------------------
int enc_fail(ldpc_encoder e, int status, const std::string& msg) { if (e) e->err = msg; else g_enc_error = msg; return status; }

template <class T>
bool upload(ldpc_encoder e, const std::vector<T>& v, T** d)
{
    *d = nullptr;
    if (v.empty()) return true;
    if (cudaMalloc((void**)d, v.size() * sizeof(T)) != cudaSuccess) return false;
    e->allocs.push_back(*d);
    return cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) == cudaSuccess;
}

// Builds the phase list; returns "" or the reason the code cannot be encoded systematically with the last n_checks positions as parity
std::string analyse(const ldpc_code_t& c, std::vector<Phase>& phases, int& max_r)
{
    const int n = c.n, R = c.n_checks, k = n - R;
    std::vector<int> row_ptr(R + 1, 0);
    { int r = 0, e = 0; for (int q = 0; q < c.nb_deg; q++) for (int z = 0; z < c.rows[q]; z++, r++) { row_ptr[r] = e; e += c.deg[q]; } row_ptr[R] = c.m; }
    std::vector<char> known(n, 0), row_done(R, 0);
    for (int v = 0; v < k; v++) known[v] = 1;
    std::vector<int> unk(R, 0);
    std::vector<std::vector<int>> rows_of(n);
    for (int r = 0; r < R; r++) for (int e = row_ptr[r]; e < row_ptr[r + 1]; e++) { const int v = (int)c.pos[e]; rows_of[v].push_back(r); if (v >= k) unk[r]++; }
    int remaining = R;           // unknown parity positions
    max_r = 0;
    while (remaining > 0) {
        // ---- peel ----
        Phase ph; ph.ptr.push_back(0);
        std::vector<int> stack;
        for (int r = 0; r < R; r++) if (!row_done[r] && unk[r] == 1) stack.push_back(r);
        while (!stack.empty()) {
            const int r = stack.back(); stack.pop_back();
            if (row_done[r] || unk[r] != 1) continue;
            int t = -1;
            for (int e = row_ptr[r]; e < row_ptr[r + 1]; e++) { const int v = (int)c.pos[e]; if (!known[v]) t = v; }
            for (int e = row_ptr[r]; e < row_ptr[r + 1]; e++) { const int v = (int)c.pos[e]; if (v != t) ph.src.push_back(v); }
            ph.ptr.push_back((int)ph.src.size()); ph.target.push_back(t);
            known[t] = 1; row_done[r] = 1; remaining--;
            for (int rr : rows_of[t]) if (!row_done[rr]) { if (--unk[rr] == 1) stack.push_back(rr); }
        }
        if (!ph.target.empty()) phases.push_back(std::move(ph));
        if (remaining == 0) break;
        // ---- dense solve of everything that is left ----
        std::vector<int> uvar, urow;
        for (int v = k; v < n; v++) if (!known[v]) uvar.push_back(v);
        for (int r = 0; r < R; r++) if (!row_done[r] && unk[r] > 0) urow.push_back(r);
        const int r_ = (int)uvar.size(), q_ = (int)urow.size();
        if (r_ > 16384) return "the part of H that peeling cannot solve has " + std::to_string(r_) + " unknowns (> 16384): dense inverse too large";
        if (q_ < r_) return "fewer independent checks than parity positions";
        std::vector<int> col_of(n, -1);
        for (int i = 0; i < r_; i++) col_of[uvar[i]] = i;
        // augmented system [A | I_q] over GF(2), Gauss-Jordan with row selection
        const int wa = (r_ + 63) / 64, wi = (q_ + 63) / 64, ws = wa + wi;
        std::vector<uint64_t> M((size_t)q_ * ws, 0);
        for (int j = 0; j < q_; j++) {
            for (int e = row_ptr[urow[j]]; e < row_ptr[urow[j] + 1]; e++) { const int cidx = col_of[c.pos[e]]; if (cidx >= 0) M[(size_t)j * ws + cidx / 64] ^= 1ull << (cidx % 64); }
            M[(size_t)j * ws + wa + j / 64] |= 1ull << (j % 64);
        }
        std::vector<int> piv_row(r_, -1);
        int next = 0;
        for (int col = 0; col < r_; col++) {
            int p = -1;
            for (int j = next; j < q_; j++) if ((M[(size_t)j * ws + col / 64] >> (col % 64)) & 1) { p = j; break; }
            if (p < 0) return "the parity part of H (last n_checks columns) is singular: no systematic encoder with the information bits first";
            if (p != next) for (int w = 0; w < ws; w++) std::swap(M[(size_t)p * ws + w], M[(size_t)next * ws + w]);
            for (int j = 0; j < q_; j++)
                if (j != next && ((M[(size_t)j * ws + col / 64] >> (col % 64)) & 1))
                    for (int w = 0; w < ws; w++) M[(size_t)j * ws + w] ^= M[(size_t)next * ws + w];
            piv_row[col] = next++;
        }
        Phase dp; dp.dense = 1; dp.uvar.assign(uvar.begin(), uvar.end());
        // b_j = XOR of the KNOWN variables of check urow[j]
        dp.eq_ptr.push_back(0);
        for (int j = 0; j < q_; j++) {
            for (int e = row_ptr[urow[j]]; e < row_ptr[urow[j] + 1]; e++) if (known[c.pos[e]]) dp.eq_src.push_back((int)c.pos[e]);
            dp.eq_ptr.push_back((int)dp.eq_src.size());
        }
        // x_col = XOR_j T[piv_row[col]][j] b_j, T = the transformation accumulated in the identity half
        dp.inv_ptr.push_back(0);
        for (int col = 0; col < r_; col++) {
            const uint64_t* t = &M[(size_t)piv_row[col] * ws + wa];
            for (int j = 0; j < q_; j++) if ((t[j / 64] >> (j % 64)) & 1) dp.inv_idx.push_back(j);
            dp.inv_ptr.push_back((int)dp.inv_idx.size());
        }
        dp.target.assign(1, q_);      // number of right-hand sides
        max_r = std::max(max_r, q_);
        for (int v : uvar) { known[v] = 1; remaining--; for (int rr : rows_of[v]) if (!row_done[rr]) unk[rr]--; }
        for (int r : urow) if (unk[r] == 0) row_done[r] = 1;
        phases.push_back(std::move(dp));
    }
    return "";
}

}  // namespace

extern "C" {

const char* ldpc_b200_encoder_last_error(ldpc_encoder e) { return e ? e->err.c_str() : g_enc_error.c_str(); }

void ldpc_b200_encoder_destroy(ldpc_encoder e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    for (void* p : e->allocs) cudaFree(p);
    cudaFree(e->d_words); cudaFree(e->d_tmp);
    delete e;
}

int ldpc_b200_encoder_create(ldpc_encoder* out, const ldpc_code_t* code, int device)
{
    if (!out || !code) return enc_fail(nullptr, LDPC_ERR_INVALID, "null argument");
    *out = nullptr;
    int rc = ldpc_b200_check_code(code);
    if (rc) return enc_fail(nullptr, rc, "malformed code table");
    std::vector<Phase> phases; int max_r = 0;
    const std::string why = analyse(*code, phases, max_r);
    if (!why.empty()) return enc_fail(nullptr, LDPC_ERR_UNSUPPORTED, why);
    const int ndev = ldpc_b200_device_count();
    if (ndev <= 0) return enc_fail(nullptr, LDPC_ERR_NO_DEVICE, "no CUDA device visible: the encoder runs on the GPU");
    if (device < 0 || device >= ndev) return enc_fail(nullptr, LDPC_ERR_INVALID, "device index out of range");
    if (cudaSetDevice(device) != cudaSuccess) return enc_fail(nullptr, LDPC_ERR_CUDA, "cudaSetDevice failed");
    ldpc_encoder e = new ldpc_b200_encoder_s();
    e->device = device; e->n = code->n; e->k = code->n - code->n_checks; e->max_r = max_r;
    for (auto& ph : phases) {
        DevPhase d{}; d.dense = ph.dense; d.n_steps = ph.dense ? 0 : (int)ph.target.size(); d.r = ph.dense ? (int)ph.uvar.size() : 0;
        bool ok = upload(e, ph.ptr, &d.ptr) && upload(e, ph.src, &d.src) && upload(e, ph.uvar, &d.uvar) && upload(e, ph.eq_ptr, &d.eq_ptr)
               && upload(e, ph.eq_src, &d.eq_src) && upload(e, ph.inv_ptr, &d.inv_ptr) && upload(e, ph.inv_idx, &d.inv_idx);
        if (!ph.dense) ok = ok && upload(e, ph.target, &d.target); else d.n_steps = ph.target[0];     // dense: number of right-hand sides
        if (!ok) { ldpc_b200_encoder_destroy(e); return enc_fail(nullptr, LDPC_ERR_CUDA, "uploading the encoder tables failed"); }
        e->phases.push_back(d);
    }
    *out = e;
    return LDPC_OK;
}

int ldpc_b200_encoder_info(ldpc_encoder e, int* n_phases, int* dense_unknowns)
{
    if (!e) return LDPC_ERR_INVALID;
    if (n_phases) *n_phases = (int)e->phases.size();
    if (dense_unknowns) { int r = 0; for (auto& p : e->phases) r += p.r; *dense_unknowns = r; }
    return LDPC_OK;
}

int ldpc_b200_encode_device(ldpc_encoder e, const uint8_t* d_info, uint8_t* d_codeword, size_t frames, uint64_t seed, uint64_t first_frame, void* cuda_stream)
{
    if (!e || !d_codeword) return enc_fail(e, LDPC_ERR_INVALID, "encode_device: bad argument");
    if (!d_info && first_frame % 32) return enc_fail(e, LDPC_ERR_INVALID, "encode_device: random information bits need first_frame % 32 == 0");
    if (frames == 0) return LDPC_OK;
#define ENC_TRY(call) do { cudaError_t x__ = (call); if (x__ != cudaSuccess) return enc_fail(e, LDPC_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(x__)); } while (0)
    ENC_TRY(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int W = (int)((frames + 31) / 32);
    const size_t need = (size_t)e->n * W * 4, need_tmp = (size_t)std::max(e->max_r, 1) * W * 4;
    if (e->words_cap < need) { ENC_TRY(cudaStreamSynchronize(st)); cudaFree(e->d_words); e->d_words = nullptr; e->words_cap = 0; ENC_TRY(cudaMalloc((void**)&e->d_words, need)); e->words_cap = need; }
    if (e->tmp_cap < need_tmp) { ENC_TRY(cudaStreamSynchronize(st)); cudaFree(e->d_tmp); e->d_tmp = nullptr; e->tmp_cap = 0; ENC_TRY(cudaMalloc((void**)&e->d_tmp, need_tmp)); e->tmp_cap = need_tmp; }
    const unsigned g1 = (unsigned)std::min<size_t>(((size_t)e->n * W + 255) / 256, 1u << 16);
    enc_pack_kernel<<<g1, 256, 0, st>>>(d_info, e->d_words, frames, e->n, e->k, W, seed, first_frame);
    for (auto& p : e->phases) {
        if (!p.dense) enc_peel_kernel<<<(W + 63) / 64, 64, 0, st>>>(e->d_words, W, p.n_steps, p.ptr, p.src, p.target);
        else {
            const unsigned g2 = (unsigned)std::min<size_t>(((size_t)p.n_steps * W + 255) / 256, 1u << 16), g3 = (unsigned)std::min<size_t>(((size_t)p.r * W + 255) / 256, 1u << 16);
            enc_dense_rhs_kernel<<<g2, 256, 0, st>>>(e->d_words, e->d_tmp, W, p.n_steps, p.eq_ptr, p.eq_src);
            enc_dense_solve_kernel<<<g3, 256, 0, st>>>(e->d_words, e->d_tmp, W, p.r, p.uvar, p.inv_ptr, p.inv_idx);
        }
    }
    const unsigned g4 = (unsigned)std::min<size_t>((frames * (size_t)e->n + 255) / 256, 1u << 18);
    enc_unpack_kernel<<<g4, 256, 0, st>>>(e->d_words, d_codeword, frames, e->n, W);
    ENC_TRY(cudaGetLastError());
    // NULL = legacy default stream, which the decoder handles' non-blocking streams are not ordered against: finish before returning,
    // so that a following ldpc_b200_awgn_codeword_device(..., NULL) (slot-0 stream of its handle) reads a complete codeword
    if (!st) ENC_TRY(cudaStreamSynchronize(st));
    return LDPC_OK;
}

int ldpc_b200_encode(ldpc_encoder e, const uint8_t* info, uint8_t* codeword, size_t frames)
{
    if (!e || !info || !codeword) return enc_fail(e, LDPC_ERR_INVALID, "encode: bad argument");
    if (frames == 0) return LDPC_OK;
    ENC_TRY(cudaSetDevice(e->device));
    uint8_t *d_i = nullptr, *d_c = nullptr;
    ENC_TRY(cudaMalloc((void**)&d_i, frames * (size_t)e->k));
    if (cudaMalloc((void**)&d_c, frames * (size_t)e->n) != cudaSuccess) { cudaFree(d_i); return enc_fail(e, LDPC_ERR_NOMEM, "encode: cudaMalloc failed"); }
    cudaMemcpy(d_i, info, frames * (size_t)e->k, cudaMemcpyHostToDevice);
    int rc = ldpc_b200_encode_device(e, d_i, d_c, frames, 0, 0, nullptr);
    if (!rc && cudaMemcpy(codeword, d_c, frames * (size_t)e->n, cudaMemcpyDeviceToHost) != cudaSuccess) rc = enc_fail(e, LDPC_ERR_CUDA, "encode: D2H failed");
    cudaFree(d_i); cudaFree(d_c);
    return rc;
#undef ENC_TRY
}

}  // extern "C"

// ---- channel and counters for non-zero codewords: these need the decoder handle's dtype / quantiser, so they live behind small
// accessors exported by ldpc_b200.cu ------------------------------------------------------------------------------------------
extern "C" int ldpc_b200_internal_channel_params(ldpc_handle h, int* device, int* n, int* n_checks, int* elem, int* llr_scale, int* sat_llr, int* packed,
                                                  unsigned long long** d_counters, void** slot0_stream);

extern "C" int ldpc_b200_awgn_codeword_device(ldpc_handle h, void* d_llr, const uint8_t* d_codeword, size_t frames, float sigma, uint64_t seed,
                                              uint64_t first_frame, void* cuda_stream)
{
    int dev, n, nc, elem, scale, sat, packed; unsigned long long* ctr; void* s0;
    if (!h || !d_llr || !d_codeword || ldpc_b200_internal_channel_params(h, &dev, &n, &nc, &elem, &scale, &sat, &packed, &ctr, &s0)) return LDPC_ERR_INVALID;
    if (frames == 0) return LDPC_OK;
    if (cudaSetDevice(dev) != cudaSuccess) return LDPC_ERR_CUDA;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : (cudaStream_t)s0;
    const size_t total = frames * (size_t)((n + 3) / 4);
    const unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 1u << 20);
    if (elem == 4) awgn_codeword_kernel<float><<<blocks, 256, 0, st>>>((float*)d_llr, d_codeword, frames, n, sigma, seed, first_frame, (float)scale, sat);
    else if (elem == 2) awgn_codeword_kernel<int16_t><<<blocks, 256, 0, st>>>((int16_t*)d_llr, d_codeword, frames, n, sigma, seed, first_frame, (float)scale, sat);
    else awgn_codeword_kernel<int8_t><<<blocks, 256, 0, st>>>((int8_t*)d_llr, d_codeword, frames, n, sigma, seed, first_frame, (float)scale, sat);
    return cudaGetLastError() == cudaSuccess ? LDPC_OK : LDPC_ERR_CUDA;
}

extern "C" int ldpc_b200_count_errors_ref_device(ldpc_handle h, const uint8_t* d_hard, const uint8_t* d_codeword, size_t frames, uint64_t* out2_host, void* cuda_stream)
{
    int dev, n, nc, elem, scale, sat, packed; unsigned long long* ctr; void* s0;
    if (!h || !d_hard || !d_codeword || !out2_host || ldpc_b200_internal_channel_params(h, &dev, &n, &nc, &elem, &scale, &sat, &packed, &ctr, &s0)) return LDPC_ERR_INVALID;
    if (cudaSetDevice(dev) != cudaSuccess) return LDPC_ERR_CUDA;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : (cudaStream_t)s0;
    if (cudaMemsetAsync(ctr, 0, 2 * sizeof(unsigned long long), st) != cudaSuccess) return LDPC_ERR_CUDA;
    if (frames) count_errors_ref_kernel<<<(unsigned)std::min<size_t>((frames + 7) / 8, 4096), 256, 0, st>>>(d_hard, d_codeword, frames, n, n - nc, packed, ctr);
    unsigned long long r[2];
    if (cudaMemcpyAsync(r, ctr, sizeof(r), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return LDPC_ERR_CUDA;
    out2_host[0] = r[0]; out2_host[1] = r[1];
    return LDPC_OK;
}
