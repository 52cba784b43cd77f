/*
 * ldpc_oracle.c — scalar CPU restatement of the reference's layered min-sum decoders.
 * TEST INFRASTRUCTURE ONLY (see ldpc_oracle.h): never linked into or called from the product path.
 *
 * One frame at a time, plain int arithmetic with explicit clamps, so that every saturation the
 * SIMD originals perform implicitly is visible.  Reference lines each piece follows are cited inline
 * (paths relative to the reference tree).
 */
#include "ldpc_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define MAXDEG 64

static inline int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }
static inline int mini(int a, int b) { return a < b ? a : b; }
static inline int maxi(int a, int b) { return a > b ? a : b; }
static inline int absi(int a) { return a < 0 ? -a : a; }

typedef struct {
    int sem, algo, offset, factor_q5, sat_var, sat_msg, et, wide, flooding;
    float f1, f2;
} oparams;

/* contribution x = posterior - old message, with the mode's rails.
 * X86_SSE/UNIFORM: _mm_subs_epi8 then max with min_var   (ref: x86/CDecoder/OMS/CDecoder_OMS_fixed_SSE.cpp:55-56,208)
 * ARM_SCALAR     : SATURATE(a, vSAT_NEG_VAR, vSAT_POS_VAR)  (ref: ldpc_decoder_arm/CDecoder/OMS/CDecoder_OMS_fixed_x86.cpp:30-32,88)
 * GPU_FIXED      : vsubss4 -> [-128,127]                  (ref: gpu_fixed/decoder_oms/cuda/CUDA_OMS_SIMD.cu:166) */
static inline int contrib(const oparams* p, int v, int m, int wide)
{
    int x = v - m;
    switch (p->sem) {
    case LDPC_SEM_X86_SSE:
    case LDPC_SEM_UNIFORM:
        if (!wide) x = clampi(x, -128, 127);
        return clampi(x, -p->sat_var, wide ? p->sat_var : 127);
    case LDPC_SEM_ARM_SCALAR:
        return clampi(x, -p->sat_var, p->sat_var);
    default: /* GPU_FIXED */
        return clampi(x, -128, 127);
    }
}

/* magnitude that enters the min1/min2 search
 * X86_SSE class 0 : min(|x|, max_msg)      (ref: CDecoder_OMS_fixed_SSE.cpp:211)
 * X86_SSE class>=1 (OMS only): |min(x, max_msg)|  (ref: CDecoder_OMS_fixed_SSE.cpp:293,314,364,384)
 * X86 NMS: min(|x|,max_msg) in every class   (ref: x86/CDecoder/NMS/CDecoder_NMS_fixed_SSE.cpp:183,275)
 * UNIFORM : min(|x|, max_msg)               (ref: x86/CDecoder/OMS/CDecoder_OMS_fixed_AVX.cpp:240,325)
 * ARM     : |SATURATE(x, -max_msg, max_msg)| (ref: CDecoder_OMS_fixed_x86.cpp:90)
 * GPU     : vabs4 with wrap-around: |-128| = 128   (ref: simd_functions.h:1104-1136, CUDA_OMS_SIMD.cu:168-169) */
static inline int magnitude(const oparams* p, int x, int cls)
{
    switch (p->sem) {
    case LDPC_SEM_X86_SSE:
        if (p->algo == LDPC_ALGO_OMS && cls >= 1) return absi(mini(x, p->sat_msg));
        return mini(absi(x), p->sat_msg);
    case LDPC_SEM_UNIFORM:
    case LDPC_SEM_ARM_SCALAR:
        return mini(absi(x), p->sat_msg);
    default:
        return absi(x);
    }
}

/* the two magnitudes of a row: c1 goes to the edge(s) holding min1, c2 to every other edge */
static inline void row_constants(const oparams* p, int min1, int min2, int cls, int first_iter, int* c1, int* c2)
{
    switch (p->sem) {
    case LDPC_SEM_X86_SSE:
    case LDPC_SEM_UNIFORM:
        if (p->algo == LDPC_ALGO_NMS && p->wide) {
            /* int16 storage (no reference decoder, own definition): the same (min*factor)>>5 without the 16-bit lane wrap and
             * with the pack saturation at the int16 rail */
            *c1 = mini((min2 * p->factor_q5) >> 5, 32767); *c2 = mini((min1 * p->factor_q5) >> 5, 32767);
        } else if (p->algo == LDPC_ALGO_NMS) {
            /* unpack to u16, mullo_epi16, srli 5, packs_epi16  (ref: CDecoder_NMS_fixed_SSE.cpp:196-208) */
            int t2 = ((min2 * p->factor_q5) & 0xFFFF) >> 5, t1 = ((min1 * p->factor_q5) & 0xFFFF) >> 5;
            *c1 = mini(t2, 127); *c2 = mini(t1, 127);
        } else {
            /* min(subs_epu8(min, offset), max_msg)  (ref: CDecoder_OMS_fixed_SSE.cpp:229-230) */
            *c1 = mini(maxi(min2 - p->offset, 0), p->sat_msg);
            *c2 = mini(maxi(min1 - p->offset, 0), p->sat_msg);
        }
        break;
    case LDPC_SEM_ARM_SCALAR:
        /* (ref: CDecoder_OMS_fixed_x86.cpp:94-95) — message clamp is applied after the sign, in pass 2 */
        *c1 = maxi(min2 - p->offset, 0);
        *c2 = maxi(min1 - p->offset, 0);
        break;
    default: /* GPU_FIXED: literals */
        switch (p->algo) {
        case LDPC_ALGO_MS:   /* (ref: gpu_fixed/decoder_ms/cuda/CUDA_MS_SIMD.cu:73-74,173-174) */
            *c1 = mini(min2, 31); *c2 = mini(min1, 31); break;
        case LDPC_ALGO_OMS:  /* (ref: CUDA_OMS_SIMD.cu:73-74,173-174,220-221); peeled first iteration forgets the clamp for the second degree class (:113-114) */
            *c1 = maxi(min2 - 1, 0); *c2 = maxi(min1 - 1, 0);
            if (!(first_iter && cls >= 1)) { *c1 = mini(*c1, 31); *c2 = mini(*c2, 31); }
            break;
        case LDPC_ALGO_NMS:  /* (ref: gpu_fixed/decoder_nms/cuda/CUDA_NMS_SIMD.cu:73-85) */
            *c1 = (int)(signed char)((float)min2 * 0.750f); *c2 = (int)(signed char)((float)min1 * 0.750f); break;
        default:             /* 2NMS (ref: gpu_fixed/decoder_2nms/cuda/CUDA_2NMS_SIMD.cu:73-85) */
            *c1 = (int)(signed char)((float)min2 * 0.875f); *c2 = (int)(signed char)((float)min1 * 0.750f); break;
        }
    }
}

static int validate(const ldpc_code_t* code, const oparams* p, int wide)
{
    if (!code || code->n <= 0 || code->m <= 0 || code->nb_deg <= 0 || code->nb_deg > LDPC_MAX_DEG_CLASSES) return LDPC_ERR_INVALID;
    long e = 0, r = 0;
    for (int c = 0; c < code->nb_deg; c++) {
        if (code->deg[c] <= 0 || code->deg[c] > MAXDEG || code->rows[c] < 0) return LDPC_ERR_INVALID;
        e += (long)code->deg[c] * code->rows[c]; r += code->rows[c];
    }
    if (e != code->m || r != code->n_checks) return LDPC_ERR_INVALID;
    if (p->sem == LDPC_SEM_X86_SSE || p->sem == LDPC_SEM_UNIFORM) {
        if (p->algo != LDPC_ALGO_OMS && p->algo != LDPC_ALGO_NMS) return LDPC_ERR_UNSUPPORTED;
        if (!wide && p->sat_var > 127) return LDPC_ERR_INVALID;
        if (wide && p->sem == LDPC_SEM_X86_SSE) return LDPC_ERR_UNSUPPORTED;
    } else if (p->sem == LDPC_SEM_ARM_SCALAR) {
        if (p->algo != LDPC_ALGO_OMS) return LDPC_ERR_UNSUPPORTED;
    } else if (p->sem == LDPC_SEM_GPU_FIXED) {
        if (wide) return LDPC_ERR_UNSUPPORTED;
    } else return LDPC_ERR_INVALID;
    return 0;
}

/* rails of the posterior / contribution in the mode (int8: hi is 127 except ARM; wide: +-sat_var) */
static inline int rail_lo(const oparams* p) { return p->sem == LDPC_SEM_GPU_FIXED ? -128 : -p->sat_var; }
static inline int rail_hi(const oparams* p) { return (p->sem == LDPC_SEM_ARM_SCALAR || p->wide) ? p->sat_var : 127; }

/* one check row: contributions from (v, m), new messages into m; posteriors written back only when `write_v` (layered). */
static inline int update_row(const ldpc_code_t* code, const oparams* p, int* v, int* m, int e, int d, int c, int first_iter, int write_v)
{
    const uint32_t* pos = code->pos;
    const int wide = p->wide;
    const int x86 = (p->sem == LDPC_SEM_X86_SSE || p->sem == LDPC_SEM_UNIFORM);
    /* running-min initial value: 127 = vSAT_POS_VAR on x86 (ref: CDecoder_OMS_fixed_SSE.cpp:177-178), 0x7F on GPU
     * (ref: CUDA_OMS_SIMD.cu:51-52), vSAT_POS_VAR+1 in the ARM scalar decoder (ref: CDecoder_OMS_fixed_x86.cpp:78-79) */
    const int min_init = (p->sem == LDPC_SEM_ARM_SCALAR) ? p->sat_var + 1 : (p->sem == LDPC_SEM_GPU_FIXED ? 127 : p->sat_var);
    int x[MAXDEG], a[MAXDEG];
    int min1 = min_init, min2 = min_init, par = 0;
    for (int j = 0; j < d; j++) {
        x[j] = contrib(p, v[pos[e + j]], m[e + j], wide);
        a[j] = magnitude(p, x[j], c);
        /* (ref: CDecoder_OMS_fixed_SSE.cpp:213-215 / CUDA_OMS_SIMD.cu:168-169) */
        int old = min1;
        min1 = mini(min1, a[j]);
        min2 = mini(min2, maxi(a[j], old));
        /* x86: sign bit, zero is positive (ref: :209-210). GPU/ARM: vcmpgts4(x,0) / Signe_de_contrib, zero is negative
         * (ref: CUDA_OMS_SIMD.cu:170, CDecoder_OMS_fixed_x86.cpp:22,91) */
        par ^= x86 ? (x[j] < 0) : (x[j] > 0);
    }
    int c1, c2;
    row_constants(p, min1, min2, c, first_iter, &c1, &c2);
    for (int j = 0; j < d; j++) {
        int mag = (a[j] == min1) ? c1 : c2;
        int msg, vn;
        if (x86) {
            /* sign ^ 0xC0 (odd degree) / 0x40 (even) then _mm_sign_epi8 (ref: CDecoder_OMS_fixed_SSE.cpp:180-190,232-236,243-244) */
            int negate = par ^ (x[j] < 0) ^ (d & 1);
            msg = negate ? -mag : mag;
            int s = x[j] + msg;
            if (!wide) s = clampi(s, -128, 127);
            vn = clampi(s, -p->sat_var, wide ? p->sat_var : 127);   /* adds_epi8 then max(min_var) (ref: :58-59,245) */
        } else if (p->sem == LDPC_SEM_ARM_SCALAR) {
            int keep = par ^ (x[j] > 0);
            msg = keep ? mag : -mag;
            msg = clampi(msg, -p->sat_msg, p->sat_msg);                       /* (ref: CDecoder_OMS_fixed_x86.cpp:101-102) */
            vn = clampi(x[j] + msg, -p->sat_var, p->sat_var);
        } else {
            int keep = par ^ (x[j] > 0);
            msg = keep ? mag : -mag;                                          /* (ref: CUDA_OMS_SIMD.cu:182-183) */
            vn = clampi(x[j] + msg, -128, 127);                               /* vaddss4 (ref: :186) */
        }
        if (write_v) v[pos[e + j]] = vn;
        m[e + j] = msg;
    }
    return par;       /* XOR of the row's flags as seen by this update (GPU_FIXED: of (x > 0)) */
}

/* decode ONE frame held in int arrays v[n] (in: LLR, out: posterior) and m[M] (out: messages). Returns iterations run.
 * Layered = every reference decoder.  Flooding has NO reference implementation — own definition, PARITY UNPINNED:
 * the row update is the mode's own (same rails, magnitudes, sign conventions, constants), applied to every row against the
 * posteriors of the previous iteration; then every posterior is rebuilt as clamp(llr + sum of its new messages) with a wide
 * accumulator; the stop criterion is the syndrome of the hard decisions.  llr0[n] is scratch for the clamped channel values. */
static int decode_frame(const ldpc_code_t* code, const oparams* p, int* v, int* m, int* llr0, int iters)
{
    const uint32_t* pos = code->pos;
    const int wide = p->wide;
    int done = 0;
    memset(m, 0, sizeof(int) * (size_t)code->m);   /* (ref: CDecoder_OMS_fixed_SSE.cpp:129-131); GPU peels iteration 1 instead (CUDA_OMS_SIMD.cu:40-132) */
    if (p->flooding) for (int i = 0; i < code->n; i++) llr0[i] = v[i] = clampi(v[i], rail_lo(p), rail_hi(p));
    for (int it = 0; it < iters; it++) {
        int e = 0;
        for (int c = 0; c < code->nb_deg; c++) {
            const int d = code->deg[c];
            for (int r = 0; r < code->rows[c]; r++, e += d) update_row(code, p, v, m, e, d, c, it == 0, !p->flooding);
        }
        if (p->flooding) {
            for (int i = 0; i < code->n; i++) v[i] = llr0[i];
            for (int q = 0; q < code->m; q++) v[pos[q]] += m[q];
            for (int i = 0; i < code->n; i++) v[i] = clampi(v[i], rail_lo(p), rail_hi(p));
        }
        done = it + 1;
        if (p->et == LDPC_ET_SYNDROME) {
            /* layered: second pass over every row with the updated messages (ref: CDecoder_OMS_fixed_x86.cpp:150-192);
             * flooding: parity of the hard decisions */
            int stop = 1; e = 0;
            for (int c = 0; c < code->nb_deg && stop; c++) {
                const int d = code->deg[c];
                for (int r = 0; r < code->rows[c]; r++, e += d) {
                    int par = 0;
                    for (int j = 0; j < d; j++) par ^= p->flooding ? (v[pos[e + j]] > 0) : (contrib(p, v[pos[e + j]], m[e + j], wide) > 0);
                    if (par) { stop = 0; break; }
                }
            }
            if (stop) break;
        }
    }
    return done;
}

static void load_params(const ldpc_params_t* prm, oparams* p)
{
    p->sem = prm->semantics; p->algo = prm->algo; p->offset = prm->offset; p->factor_q5 = prm->factor_q5;
    p->sat_var = prm->sat_var; p->sat_msg = prm->sat_msg; p->et = prm->early_term; p->f1 = prm->factor1; p->f2 = prm->factor2;
    p->wide = 0; p->flooding = prm->schedule == LDPC_SCHED_FLOODING;
}

static int decode_range(const ldpc_code_t* code, const oparams* p, const void* llr, uint8_t* hard, void* post, void* msgs,
                        uint8_t* iters_done, size_t f0, size_t f1, int iters, int wide)
{
    const int n = code->n, M = code->m;
    int* v = (int*)malloc(sizeof(int) * (size_t)(2 * n + M));
    if (!v) return LDPC_ERR_NOMEM;
    int* m = v + n;
    int* llr0 = m + M;
    for (size_t f = f0; f < f1; f++) {
        if (wide) { const int16_t* q = (const int16_t*)llr + f * n; for (int i = 0; i < n; i++) v[i] = q[i]; }
        else      { const int8_t*  q = (const int8_t*)llr  + f * n; for (int i = 0; i < n; i++) v[i] = q[i]; }
        int done = decode_frame(code, p, v, m, llr0, iters);
        for (int i = 0; i < n; i++) hard[f * n + i] = (uint8_t)(v[i] > 0);   /* (ref: x86/CTools/CTools.cpp:370; GPU_Transpose_uint8.cu:29) */
        if (iters_done) iters_done[f] = (uint8_t)done;
        if (post) {
            if (wide) { int16_t* o = (int16_t*)post + f * n; for (int i = 0; i < n; i++) o[i] = (int16_t)v[i]; }
            else      { int8_t*  o = (int8_t*)post  + f * n; for (int i = 0; i < n; i++) o[i] = (int8_t)v[i]; }
        }
        if (msgs) {
            if (wide) { int16_t* o = (int16_t*)msgs + f * (size_t)M; for (int i = 0; i < M; i++) o[i] = (int16_t)m[i]; }
            else      { int8_t*  o = (int8_t*)msgs  + f * (size_t)M; for (int i = 0; i < M; i++) o[i] = (int8_t)m[i]; }
        }
    }
    free(v);
    return 0;
}

int oracle_decode_fixed(const ldpc_code_t* code, const ldpc_params_t* prm, const void* llr, uint8_t* hard, void* post, void* msgs,
                        uint8_t* iters_done, size_t frames, int iters, int elem_bytes)
{
    if (!prm || !llr || !hard || (elem_bytes != 1 && elem_bytes != 2) || iters < 0) return LDPC_ERR_INVALID;
    oparams p; load_params(prm, &p);
    const int wide = elem_bytes == 2;
    p.wide = wide;
    int rc = validate(code, &p, wide);
    if (rc) return rc;
    return decode_range(code, &p, llr, hard, post, msgs, iters_done, 0, frames, iters, wide);
}

typedef struct { const ldpc_code_t* code; const oparams* p; const void* llr; uint8_t* hard; size_t f0, f1; int iters, wide, rc; } mt_job;
static void* mt_worker(void* arg)
{
    mt_job* j = (mt_job*)arg;
    j->rc = decode_range(j->code, j->p, j->llr, j->hard, NULL, NULL, NULL, j->f0, j->f1, j->iters, j->wide);
    return NULL;
}

int oracle_decode_fixed_mt(const ldpc_code_t* code, const ldpc_params_t* prm, const void* llr, uint8_t* hard,
                           size_t frames, int iters, int elem_bytes, int threads)
{
    if (!prm || !llr || !hard || (elem_bytes != 1 && elem_bytes != 2) || iters < 0) return LDPC_ERR_INVALID;
    oparams p; load_params(prm, &p);
    const int wide = elem_bytes == 2;
    p.wide = wide;
    int rc = validate(code, &p, wide);
    if (rc) return rc;
    if (threads < 1) threads = 1;
    if (threads > 1024) threads = 1024;
    mt_job* jobs = (mt_job*)calloc((size_t)threads, sizeof(mt_job));
    pthread_t* tid = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    if (!jobs || !tid) { free(jobs); free(tid); return LDPC_ERR_NOMEM; }
    int started = 0;
    for (int t = 0; t < threads; t++) {
        mt_job j = { code, &p, llr, hard, frames * (size_t)t / (size_t)threads, frames * (size_t)(t + 1) / (size_t)threads, iters, wide, 0 };
        jobs[t] = j;
        if (pthread_create(&tid[t], NULL, mt_worker, &jobs[t])) { mt_worker(&jobs[t]); tid[t] = 0; } else started |= 1;
    }
    for (int t = 0; t < threads; t++) { if (tid[t]) pthread_join(tid[t], NULL); if (jobs[t].rc) rc = jobs[t].rc; }
    (void)started;
    free(jobs); free(tid);
    return rc;
}

/* ---------------------------------------------------------------------------------------------------------------------
 * Float normalised / offset min-sum.  NO reference implementation exists (float kernels are declarations only:
 * gpu_fixed/decoder_template/GPU_Scheduled_functions.h:31-34,54-61; x86 float decode is an empty stub:
 * x86/CDecoder/template/CDecoder_fixed_SSE.cpp:35-40).  This is the library's own definition — PARITY UNPINNED.
 * Convention kept from the fixed-point decoders: bit 1 <=> LLR > 0, edge keeps a positive sign iff the XOR of the other
 * edges' (x > 0) flags is 1.  Every product/sum is a single rounded fp32 operation, posterior sums run in edge order,
 * so a GPU implementation following the same order is reproducible to the last bit.
 * ------------------------------------------------------------------------------------------------------------------ */
static int decode_frame_float(const ldpc_code_t* code, const ldpc_params_t* prm, const float* llr, float* post, float* c2v, float* v2c, int iters)
{
    const int n = code->n, M = code->m;
    const uint32_t* pos = code->pos;
    const float f1 = prm->factor1, f2 = (prm->algo == LDPC_ALGO_2NMS) ? prm->factor2 : prm->factor1;
    const int flooding = prm->schedule == LDPC_SCHED_FLOODING;
    /* float OMS: the fixed-point offset expressed in channel units (offset / llr_scale, exact for power-of-two scales) */
    const float off = (float)prm->offset / (float)(prm->llr_scale > 0 ? prm->llr_scale : 1);
    int done = 0;
    for (int i = 0; i < n; i++) post[i] = llr[i];
    for (int e = 0; e < M; e++) { c2v[e] = 0.0f; v2c[e] = llr[pos[e]]; }
    for (int it = 0; it < iters; it++) {
        int e = 0;
        for (int c = 0; c < code->nb_deg; c++) {
            const int d = code->deg[c];
            for (int r = 0; r < code->rows[c]; r++, e += d) {
                float x[MAXDEG], a[MAXDEG];
                float min1 = INFINITY, min2 = INFINITY; int par = 0;
                for (int j = 0; j < d; j++) {
                    x[j] = flooding ? v2c[e + j] : (post[pos[e + j]] - c2v[e + j]);
                    a[j] = fabsf(x[j]);
                    float old = min1;
                    min1 = fminf(min1, a[j]);
                    min2 = fminf(min2, fmaxf(a[j], old));
                    par ^= (x[j] > 0.0f);
                }
                volatile float c1, c2;                           /* volatile: one rounding each, no contraction */
                if (prm->algo == LDPC_ALGO_OMS) { c1 = fmaxf(min2 - off, 0.0f); c2 = fmaxf(min1 - off, 0.0f); }
                else { c1 = min2 * f2; c2 = min1 * f1; }
                for (int j = 0; j < d; j++) {
                    float mag = (a[j] == min1) ? c1 : c2;
                    int keep = par ^ (x[j] > 0.0f);
                    float msg = keep ? mag : -mag;
                    c2v[e + j] = msg;
                    if (!flooding) post[pos[e + j]] = x[j] + msg;
                }
            }
        }
        if (flooding) {
            for (int i = 0; i < n; i++) post[i] = llr[i];
            for (int q = 0; q < M; q++) post[pos[q]] = post[pos[q]] + c2v[q];   /* edge order */
            for (int q = 0; q < M; q++) v2c[q] = post[pos[q]] - c2v[q];
        }
        done = it + 1;
        if (prm->early_term == LDPC_ET_SYNDROME) {
            int stop = 1; e = 0;
            for (int c = 0; c < code->nb_deg && stop; c++) {
                const int d = code->deg[c];
                for (int r = 0; r < code->rows[c]; r++, e += d) {
                    int par = 0;
                    for (int j = 0; j < d; j++) par ^= (post[pos[e + j]] > 0.0f);
                    if (par) { stop = 0; break; }
                }
            }
            if (stop) break;
        }
    }
    return done;
}

int oracle_decode_float(const ldpc_code_t* code, const ldpc_params_t* prm, const float* llr, uint8_t* hard, float* post, float* msgs,
                        uint8_t* iters_done, size_t frames, int iters)
{
    if (!code || !prm || !llr || !hard || iters < 0) return LDPC_ERR_INVALID;
    if (prm->algo < LDPC_ALGO_MS || prm->algo > LDPC_ALGO_2NMS) return LDPC_ERR_UNSUPPORTED;
    ldpc_params_t q = *prm;
    if (q.algo == LDPC_ALGO_MS) { q.factor1 = 1.0f; q.factor2 = 1.0f; }
    const int n = code->n, M = code->m;
    float* buf = (float*)malloc(sizeof(float) * (size_t)(n + 2 * (size_t)M));
    if (!buf) return LDPC_ERR_NOMEM;
    float *p = buf, *c2v = buf + n, *v2c = c2v + M;
    for (size_t f = 0; f < frames; f++) {
        int done = decode_frame_float(code, &q, llr + f * n, p, c2v, v2c, iters);
        for (int i = 0; i < n; i++) hard[f * n + i] = (uint8_t)(p[i] > 0.0f);
        if (post) memcpy(post + f * n, p, sizeof(float) * (size_t)n);
        if (msgs) memcpy(msgs + f * (size_t)M, c2v, sizeof(float) * (size_t)M);
        if (iters_done) iters_done[f] = (uint8_t)done;
    }
    free(buf);
    return 0;
}

void oracle_quantize(const float* y, int8_t* q, size_t count, int scale, int sat)
{
    for (size_t i = 0; i < count; i++) {
        int value = (int)((float)scale * y[i]);          /* C truncation toward zero (ref: CFastFixConversion.cpp:59) */
        value = (value > -sat) ? value : -sat;             /* (ref: :60) */
        value = (value < sat) ? value : sat;               /* (ref: :61) */
        q[i] = (int8_t)value;
    }
}

void oracle_pack_bits(const uint8_t* hard, uint8_t* packed, size_t frames, int n)
{
    const int nb = (n + 7) / 8;
    for (size_t f = 0; f < frames; f++)
        for (int b = 0; b < nb; b++) {
            unsigned v = 0;
            for (int k = 0; k < 8 && b * 8 + k < n; k++) v |= (unsigned)(hard[f * n + b * 8 + k] & 1) << k;
            packed[f * nb + b] = (uint8_t)v;
        }
}
