/*
 * ldpc_sim.c — the reference's BER/FER simulator, restated in C over the C ABI of libldpc_b200.so.
 *
 * Same command line, same banner, same "(RT)" progress and "SNR = ... | BER = ... | FER = ..." result lines as the reference's
 * simulators (ref: code/x86/main_p.cpp:112-770 — argv parsing :154-336, banner :339-371, Eb/N0 loop :400-660;
 * code/gpu_fixed/test.cpp:97-640; report lines code/x86/CTerminal/CTerminal.cpp:38-90), so curves can be compared line for line.
 * What differs is where the work runs: channel, decoder and error counters all stay on the GPU (ldpc_b200_awgn_device ->
 * ldpc_b200_decode_device -> ldpc_b200_count_errors_device), one host thread per GPU, frames sharded by counter ranges of the
 * Philox channel generator; only three integers per batch come back to the host.  There is no CPU decode path here.
 *
 *   ldpc_sim -fixed|-float [reference options] [-code 576x288 | -header constantes_sse.h [-table constantes_decoder.h]]
 *            [-gpus n] [-frames n] [-flooding] [-early] [-int16] [-MS] [-2NMS] [-sse|-avx|-x86|-gpu] [-seed s] [-max-frames n] [-encoder]
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../include/ldpc_b200.h"

#define MAX_GPUS 16

typedef struct {
    double snr_min, snr_max, snr_pas;
    int fe_limit, timer_s, worst_case, ber_limit, fer_limit;
    double ber_limit_value, fer_limit_value;
    int iters, gpus, real_encoder;
    size_t frames, max_frames;
    uint64_t seed;
} sim_t;

typedef struct {
    int gpu;
    ldpc_handle h;
    void* d_llr; uint8_t* d_hard; uint8_t* d_cw; ldpc_encoder enc;
    const sim_t* sim; const ldpc_code_t* code;
    float sigma; int point;
    /* shared accumulators */
    pthread_mutex_t* mu; uint64_t* frames; uint64_t* be; uint64_t* fe; volatile int* stop; uint64_t* next_frame;
    double decode_seconds; uint64_t decoded;
    int rc; char err[256];
} worker_t;

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

static void show_time(unsigned long secondes)   /* (ref: CTerminal::ShowTime code/x86/CTerminal/CTerminal.cpp:27-36) */
{
    int ss = (int)(secondes % 60), mn = (int)((secondes / 60) % 60), hh = (int)(secondes / 3600);
    printf("%2.2dh%2.2d'%2.2d", hh, mn, ss);
}

static void* worker(void* arg)
{
    worker_t* w = (worker_t*)arg;
    const sim_t* s = w->sim;
    const int k_info = w->code->n - w->code->n_checks;
    (void)k_info;
    while (!*w->stop) {
        pthread_mutex_lock(w->mu);
        const uint64_t first = *w->next_frame; *w->next_frame += s->frames;
        pthread_mutex_unlock(w->mu);
        uint64_t cnt[2];
        const double t0 = now_s();
        int rc;
        if (s->real_encoder) {      /* random information bits -> systematic codeword -> BPSK + AWGN (ref: -encoder, main_p.cpp:232-233, GenericEncoder.cpp:38-78) */
            rc = ldpc_b200_encode_device(w->enc, NULL, w->d_cw, s->frames, s->seed * 7919u + (uint64_t)w->point, first, ldpc_b200_stream(w->h, 0));   /* one stream for the whole batch: NULL below = the same slot-0 stream */
            if (!rc) rc = ldpc_b200_awgn_codeword_device(w->h, w->d_llr, w->d_cw, s->frames, w->sigma, s->seed + (uint64_t)w->point, first, NULL);
        } else rc = ldpc_b200_awgn_device(w->h, w->d_llr, s->frames, w->sigma, s->seed + (uint64_t)w->point, first, NULL);
        if (!rc) rc = ldpc_b200_decode_device(w->h, w->d_llr, w->d_hard, s->frames, s->iters, NULL, NULL);
        if (!rc) rc = s->real_encoder ? ldpc_b200_count_errors_ref_device(w->h, w->d_hard, w->d_cw, s->frames, cnt, NULL)
                                      : ldpc_b200_count_errors_device(w->h, w->d_hard, s->frames, cnt, NULL);
        if (rc) { w->rc = rc; snprintf(w->err, sizeof(w->err), "%s", ldpc_b200_last_error(w->h)); *w->stop = 1; break; }
        w->decode_seconds += now_s() - t0; w->decoded += s->frames;
        pthread_mutex_lock(w->mu);
        *w->frames += s->frames; *w->be += cnt[0]; *w->fe += cnt[1];
        if (*w->fe >= (uint64_t)s->fe_limit) *w->stop = 1;                                   /* (ref: main_p.cpp:602-604) */
        if (s->max_frames && *w->frames >= s->max_frames) *w->stop = 1;
        pthread_mutex_unlock(w->mu);
    }
    return NULL;
}

static int bits_range(const char* a) { int b = atoi(a); return (1 << (b - 1)) - 1; }

int main(int argc, char** argv)
{
    sim_t sim = { 0.5, 3.0, 0.1, 100, -1, 0, 0, 0, 0.0, 0.0, 30, 1, 0, 65536, 0, 1 };             /* (ref: defaults main_p.cpp:118-122,134,142-143) */
    ldpc_params_t prm; ldpc_b200_default_params(&prm);
    const char *code_name = "576x288", *header = NULL, *table = NULL;
    int bits_llr = 6, bits_msg = 6, bits_var = 8, fraq = 3;                                     /* (ref: main_p.cpp:90-104) */
    char type[8] = "OMS";
    float nms_float = 0.75f, oms_float = 0.15f;

    if (argc < 2 || (strcmp(argv[1], "-fixed") && strcmp(argv[1], "-float"))) {
        printf("(EE) First argument must be the decoder data format [-float, -fixed]\n");         /* (ref: main_p.cpp:158-161) */
        printf("(EE) -float : floatting   point format decoder\n(EE) -fixed : fixed-point point format decoder\n");
        return 0;
    }
    const int is_float = !strcmp(argv[1], "-float");
    if (is_float) { prm.dtype = LDPC_DTYPE_F32; prm.algo = LDPC_ALGO_NMS; strcpy(type, "NMS"); }
    for (int p = 2; p < argc; p++) {
        const char* a = argv[p];
        const char* v = (p + 1 < argc) ? argv[p + 1] : "0";
        if (!strcmp(a, "-min")) { sim.snr_min = atof(v); p++; }
        else if (!strcmp(a, "-max")) { sim.snr_max = atof(v); p++; }
        else if (!strcmp(a, "-pas")) { sim.snr_pas = atof(v); p++; }
        else if (!strcmp(a, "-fer")) { sim.fe_limit = atoi(v); p++; }
        else if (!strcmp(a, "-wc_fer")) printf("(WW) -wc_fer: the device counters look at the information part only (ref: CErrorAnalyzer.cpp:129-137); ignored\n");
        else if (!strcmp(a, "-timer")) { sim.timer_s = atoi(v); p++; }
        else if (!strcmp(a, "-qef")) { sim.ber_limit = 1; sim.ber_limit_value = atof(v); p++; }
        else if (!strcmp(a, "-tfer")) { sim.fer_limit = 1; sim.fer_limit_value = atof(v); p++; }
        else if (!strcmp(a, "-encoder")) sim.real_encoder = 1;
        else if (!strcmp(a, "-bpsk") || !strcmp(a, "-Eb/N0") || !strcmp(a, "-random") || !strcmp(a, "-histo")) { /* accepted, no effect here */ }
        else if (!strcmp(a, "-thread")) { p++; /* host threads of the CPU simulator: the GPU build shards over -gpus instead */ }
        else if (!strcmp(a, "-sse")) prm.semantics = LDPC_SEM_X86_SSE;
        else if (!strcmp(a, "-avx")) prm.semantics = LDPC_SEM_UNIFORM;
        else if (!strcmp(a, "-x86")) prm.semantics = LDPC_SEM_ARM_SCALAR;
        else if (!strcmp(a, "-gpu")) prm.semantics = LDPC_SEM_GPU_FIXED;
        else if (!strcmp(a, "-NMS")) { strcpy(type, "NMS"); prm.algo = LDPC_ALGO_NMS; nms_float = (float)atof(v); prm.factor_q5 = atoi(v); p++; }
        else if (!strcmp(a, "-OMS")) { strcpy(type, "OMS"); prm.algo = LDPC_ALGO_OMS; oms_float = (float)atof(v); prm.offset = atoi(v); p++; }
        else if (!strcmp(a, "-MS")) { strcpy(type, "MS"); prm.algo = LDPC_ALGO_MS; }
        else if (!strcmp(a, "-2NMS")) { strcpy(type, "2NMS"); prm.algo = LDPC_ALGO_2NMS; }
        else if (!strcmp(a, "-iter")) { sim.iters = atoi(v); p++; }
        else if (!strcmp(a, "-var")) { bits_var = atoi(v); prm.sat_var = bits_range(v); p++; }
        else if (!strcmp(a, "-msg")) { bits_msg = atoi(v); prm.sat_msg = bits_range(v); p++; }
        else if (!strcmp(a, "-llr")) { bits_llr = atoi(v); prm.sat_llr = bits_range(v); fraq = bits_llr / 2; prm.llr_scale = 1 << fraq; p++; }
        else if (!strcmp(a, "-fraq")) { fraq = atoi(v); prm.llr_scale = 1 << fraq; p++; }
        /* ---- options the reference does not have ---- */
        else if (!strcmp(a, "-code")) { code_name = v; p++; }
        else if (!strcmp(a, "-header")) { header = v; p++; }
        else if (!strcmp(a, "-table")) { table = v; p++; }
        else if (!strcmp(a, "-gpus")) { sim.gpus = atoi(v); p++; }
        else if (!strcmp(a, "-frames")) { sim.frames = (size_t)atoll(v); p++; }
        else if (!strcmp(a, "-max-frames")) { sim.max_frames = (size_t)atoll(v); p++; }
        else if (!strcmp(a, "-seed")) { sim.seed = (uint64_t)atoll(v); p++; }
        else if (!strcmp(a, "-flooding")) prm.schedule = LDPC_SCHED_FLOODING;
        else if (!strcmp(a, "-layered")) prm.schedule = LDPC_SCHED_LAYERED;
        else if (!strcmp(a, "-early")) prm.early_term = LDPC_ET_SYNDROME;
        else if (!strcmp(a, "-int16")) prm.dtype = LDPC_DTYPE_I16;
        else { printf("(EE) Unknown argument (%d) => [%s]\n", p, a); return 0; }                /* (ref: main_p.cpp:331-334) */
    }
    if (is_float) {
        if (prm.algo == LDPC_ALGO_NMS || prm.algo == LDPC_ALGO_2NMS) { prm.factor1 = nms_float; if (prm.algo == LDPC_ALGO_NMS) prm.factor2 = nms_float; }
        if (prm.algo == LDPC_ALGO_OMS) { prm.llr_scale = 1000; prm.offset = (int)lroundf(oms_float * 1000.0f); }   /* offset/llr_scale channel units */
    }
    if (prm.dtype == LDPC_DTYPE_I16 && prm.semantics != LDPC_SEM_ARM_SCALAR) prm.semantics = LDPC_SEM_UNIFORM;
    if (prm.schedule == LDPC_SCHED_FLOODING || prm.dtype != LDPC_DTYPE_I8) { /* generic engine picks itself */ }

    ldpc_code_t code; memset(&code, 0, sizeof(code));
    int rc;
    if (header) rc = ldpc_b200_load_code_header(&code, header, table);
    else {
        char path[1024];
        const char* dir = getenv("LDPC_B200_CODES");
        if (strchr(code_name, '/') || strstr(code_name, ".ldpc")) snprintf(path, sizeof(path), "%s", code_name);
        else snprintf(path, sizeof(path), "%s/%s.ldpc", dir ? dir : LDPC_CODES_DIR, code_name);
        rc = ldpc_b200_load_code_table(&code, path);
    }
    if (rc) { printf("(EE) cannot load the code table (%s)\n", ldpc_b200_status_string(rc)); return 1; }
    const int N = code.n, K = code.n_checks, info = N - K;
    const int ndev = ldpc_b200_device_count();
    if (ndev < 1) { printf("(EE) no CUDA device: this simulator has no CPU decoder (use the reference's ldpcX86 for that)\n"); return 1; }
    if (sim.gpus < 1) sim.gpus = 1;
    if (sim.gpus > ndev) { printf("(WW) %d GPUs requested, %d visible => using %d\n", sim.gpus, ndev, ndev); sim.gpus = ndev; }
    if (sim.gpus > MAX_GPUS) sim.gpus = MAX_GPUS;

    const double rendement = (double)info / (double)N;
    printf("(II) LDPC DECODER - B200 (sm_100a) decoder behind the C ABI of libldpc_b200 (ABI %d)\n", ldpc_b200_abi_version());
    printf("(II) NUMBER OF GPUs       : %d (frames per batch and GPU: %zu)\n", sim.gpus, sim.frames);
    printf("(II) Code LDPC (N, K)     : (%d,%d)\n", N, K);                                    /* K = number of checks, as in the reference */
    printf("(II) Rendement du code    : %.3f\n", rendement);
    if (!strcmp(type, "MS")) printf("(II) LDPC HEURISTIC (CN)  : MIN-SUM\n");
    else if (!strcmp(type, "OMS")) printf("(II) LDPC HEURISTIC (CN)  : OFFSET MIN-SUM\n");
    else printf("(II) LDPC HEURISTIC (CN)  : NORMALIZED-MIN-SUM%s\n", !strcmp(type, "2NMS") ? " (2 factors)" : "");
    printf("(II) SCHEDULE             : %s%s\n", prm.schedule == LDPC_SCHED_FLOODING ? "FLOODING" : "HORIZONTAL LAYERED", prm.early_term ? " + SYNDROME STOP" : "");
    printf("(II) # ITERATIONs du CODE : %d\n", sim.iters);
    printf("(II) FER LIMIT FOR SIMU   : %d\n", sim.fe_limit);
    printf("(II) SIMULATION  RANGE    : [%.2f, %.2f], STEP = %.2f\n", sim.snr_min, sim.snr_max, sim.snr_pas);
    if (!is_float) {
        printf("(II) LLR DATA    Q(%d,%d)   : %d bits [%d, %d]\n", bits_llr - fraq, fraq, bits_llr, -prm.sat_llr, prm.sat_llr);
        printf("(II) MESSAGE     Q(%d,%d)   : %d bits [%d, %d]\n", bits_msg - fraq, fraq, bits_msg, -prm.sat_msg, prm.sat_msg);
        printf("(II) VARIABLE    Q(%d,%d)   : %d bits [%d, %d]\n", bits_var - fraq, fraq, bits_var, -prm.sat_var, prm.sat_var);
        if (!strcmp(type, "OMS")) printf("(II) VARIABLE BETA FIX    : %d\n", prm.offset);
        if (!strcmp(type, "NMS")) printf("(II) NORMALIZE FACTOR     : %d\n", prm.factor_q5);
        printf("(II) FACTEUR BETA (LLR)   : %d\n", prm.llr_scale);
    } else {
        if (!strcmp(type, "OMS")) printf("(II) OFFSET FACTOR        : %f\n", oms_float);
        if (!strcmp(type, "NMS")) printf("(II) NORMALIZE FACTOR     : %f\n", nms_float);
    }

    worker_t w[MAX_GPUS]; memset(w, 0, sizeof(w));
    const size_t elem = prm.dtype == LDPC_DTYPE_F32 ? 4 : (prm.dtype == LDPC_DTYPE_I16 ? 2 : 1);
    for (int g = 0; g < sim.gpus; g++) {
        w[g].gpu = g; w[g].sim = &sim; w[g].code = &code;
        rc = ldpc_b200_create(&w[g].h, &code, &prm, g, sim.frames);
        if (rc) { printf("(EE) Requested LDPC decoder does not exist (%s)\n", ldpc_b200_last_error(NULL)); return 1; }   /* (ref: DecoderLibrary.h:129-132) */
        if (ldpc_b200_device_alloc(w[g].h, &w[g].d_llr, sim.frames * (size_t)N * elem) || ldpc_b200_device_alloc(w[g].h, (void**)&w[g].d_hard, sim.frames * (size_t)N)) {
            printf("(EE) device allocation failed\n"); return 1;
        }
        if (sim.real_encoder) {
            if (sim.frames % 32) { printf("(EE) -encoder needs -frames to be a multiple of 32\n"); return 1; }
            rc = ldpc_b200_encoder_create(&w[g].enc, &code, g);
            if (rc) { printf("(EE) no systematic encoder for this table: %s\n", ldpc_b200_encoder_last_error(NULL)); return 1; }
            if (ldpc_b200_device_alloc(w[g].h, (void**)&w[g].d_cw, sim.frames * (size_t)N)) { printf("(EE) device allocation failed\n"); return 1; }
        }
    }
    int64_t kern = 0; ldpc_b200_get_info(w[0].h, LDPC_INFO_KERNEL, &kern);
    printf("(II) ENCODER              : %s\n", sim.real_encoder ? "systematic, derived from H (random information bits)" : "all-zero codeword (CFakeEncoder)");
    printf("(II) DECODE KERNEL        : %s\n", kern == 2 ? "row-parallel, on-chip state" : kern == 3 ? "generic engine (fp32 arithmetic, HBM state)" : kern == 5 ? "generic engine (fp32 arithmetic, on-chip state)" : kern == 6 ? "generic engine (fp32 arithmetic, on-chip state, warp per frame)" : kern == 4 ? "frame-parallel, bulk-copy staged" : "frame-parallel, HBM state");

    const double t_simu = now_s();
    int point = 0;
    for (double ebn0 = sim.snr_min; ebn0 <= sim.snr_max; ebn0 += sim.snr_pas, point++) {
        /* sigma of the reference's channel (ref: code/x86/CChanel/CChanelAWGN_MKL.cpp:97-110) */
        const float sigma = (float)sqrt(pow(10.0, -(ebn0 + 10.0 * log10(rendement)) / 10.0) / 2.0);
        pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
        uint64_t frames = 0, be = 0, fe = 0, next_frame = 0; volatile int stop = 0;
        pthread_t tid[MAX_GPUS];
        const double t0 = now_s();
        for (int g = 0; g < sim.gpus; g++) {
            w[g].sigma = sigma; w[g].point = point; w[g].mu = &mu; w[g].frames = &frames; w[g].be = &be; w[g].fe = &fe; w[g].stop = &stop; w[g].next_frame = &next_frame;
            w[g].decode_seconds = 0.0; w[g].decoded = 0;
            pthread_create(&tid[g], NULL, worker, &w[g]);
        }
        while (!stop) {                                                                      /* (RT) line about once a second (ref: CTerminal::temp_report) */
            struct timespec nap = { 0, 200 * 1000 * 1000 }; nanosleep(&nap, NULL);
            if (sim.timer_s != -1 && now_s() - t_simu >= sim.timer_s) stop = 1;                /* (ref: main_p.cpp:609-611) */
            pthread_mutex_lock(&mu);
            const uint64_t F = frames, BE = be, FE = fe;
            pthread_mutex_unlock(&mu);
            if (!F) continue;
            const unsigned long temps = (unsigned long)fmax(1.0, now_s() - t0);
            const unsigned long fpmn = (unsigned long)(60 * F / temps);
            printf("(RT) FRA: %8lu | FE: %3d | FER: %2.2e | BE : %5d | BER: %2.2e | [BE/FE] : %4f | FPM: %3lu | BPS: %2.2f | ETA: ",
                   (unsigned long)F, (int)FE, FE ? (double)FE / F : 1.0 / F, (int)BE, BE ? (double)BE / F / info : 1.0 / F / info,
                   FE ? (float)BE / (float)FE : (float)BE, fpmn, (double)fpmn * N / 60.0 / 1e6);
            show_time(temps); printf("\r"); fflush(stdout);
        }
        for (int g = 0; g < sim.gpus; g++) pthread_join(tid[g], NULL);
        for (int g = 0; g < sim.gpus; g++) if (w[g].rc) { printf("\n(EE) GPU %d: %s\n", g, w[g].err); return 1; }
        const double elapsed = now_s() - t0;
        const unsigned long temps = (unsigned long)elapsed + 1;
        const double ber = (double)be / (double)frames / (sim.worst_case ? N : info), fer = (double)fe / (double)frames;   /* (ref: CErrorAnalyzer.cpp:183-194) */
        const unsigned long fpmn = (unsigned long)(60 * frames / temps);
        printf("SNR = %.2f | BER =  %2.3e | FER =  %2.3e | BPS =  %2.2f | MATRICES = %10lu| FE = %d | BE = %d | BE/FE = %.1f | RUNTIME = ",
               ebn0, ber, fer, (double)fpmn * N / 60.0 / 1e6, (unsigned long)frames, (int)fe, (int)be, fe ? (double)be / (double)fe : 0.0);
        show_time(temps); printf("\n");
        {   /* throughput of the point: every GPU's own busy time (ref: the (PERF) block main_p.cpp:620-632; air = coded bits) */
            double air = 0.0;
            for (int g = 0; g < sim.gpus; g++) if (w[g].decode_seconds > 0) air += (double)w[g].decoded * N / w[g].decode_seconds / 1e6;
            printf("(PERF) SNR = %.2f, ITERS = %d, %d GPU(s): channel + decode + count, air throughput = %1.3f Mbps, info = %1.3f Mbps\n",
                   ebn0, sim.iters, sim.gpus, air, air * rendement);
        }
        fflush(stdout);
        if (sim.timer_s != -1 && now_s() - t_simu >= sim.timer_s) { printf("(II) THE SIMULATION HAS STOP DUE TO THE (USER) TIME CONTRAINT.\n"); break; }
        if (sim.ber_limit && ber < sim.ber_limit_value) { printf("(II) THE SIMULATION HAS STOP DUE TO THE (USER) QUASI-ERROR FREE CONTRAINT (on BER).\n"); break; }
        if (sim.fer_limit && fer < sim.fer_limit_value) { printf("(II) THE SIMULATION HAS STOP DUE TO THE (USER) QUASI-ERROR FREE CONTRAINT (on FER).\n"); break; }
    }
    for (int g = 0; g < sim.gpus; g++) { ldpc_b200_device_free(w[g].h, w[g].d_llr); ldpc_b200_device_free(w[g].h, w[g].d_hard);
        if (sim.real_encoder) { ldpc_b200_device_free(w[g].h, w[g].d_cw); ldpc_b200_encoder_destroy(w[g].enc); }
        ldpc_b200_destroy(w[g].h); }
    ldpc_b200_free_code(&code);
    return 0;
}
