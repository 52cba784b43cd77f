// code_table.cpp — H-matrix load: the reference's compile-time code tables as run-time data.  Host-only (no CUDA).
//
// The reference selects a code by #include-ing a header that defines _N/_K/_M, NB_DEGRES, DEG_i, DEG_i_COMPUTATIONS and
// the edge list PosNoeudsVariable[_M] (ref: code/x86/Constantes/576x288/constantes_sse.h:26-60; GPU flavour split over
// code/gpu_fixed/matrix/576x288/constantes_gpu.h:6-39 and constantes_decoder.h:3).  ldpc_b200_load_code_header reads exactly
// those files; ldpc_b200_{save,load}_code_table use this library's own compact binary form (what ships under
// ldpcgputegra_b200/codes/, produced by tools/import_codes.py).
#include "../../include/ldpc_b200.h"

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

bool read_file(const char* path, std::string& out)
{
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 0) { fclose(f); return false; }
    out.resize((size_t)sz);
    size_t got = sz ? fread(&out[0], 1, (size_t)sz, f) : 0;
    fclose(f);
    return got == (size_t)sz;
}

// blank out /* */ and // comments, keeping newlines (the tables carry "/* msg = 0, deg = 7 */" on every row)
void strip_comments(std::string& s)
{
    size_t i = 0, n = s.size();
    while (i < n) {
        if (s[i] == '/' && i + 1 < n && s[i + 1] == '*') {
            size_t j = i + 2;
            while (j + 1 < n && !(s[j] == '*' && s[j + 1] == '/')) j++;
            size_t end = (j + 1 < n) ? j + 2 : n;
            for (size_t k = i; k < end; k++) if (s[k] != '\n') s[k] = ' ';
            i = end;
        } else if (s[i] == '/' && i + 1 < n && s[i + 1] == '/') {
            while (i < n && s[i] != '\n') s[i++] = ' ';
        } else i++;
    }
}

// "#define NAME <integer literal>" only; expression-valued macros (SAT_POS_VAR ...) are not needed for the table
void collect_defines(const std::string& s, std::map<std::string, long>& defs)
{
    size_t pos = 0;
    while ((pos = s.find("#define", pos)) != std::string::npos) {
        size_t i = pos + 7;
        while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) i++;
        size_t b = i;
        while (i < s.size() && (isalnum((unsigned char)s[i]) || s[i] == '_')) i++;
        std::string name = s.substr(b, i - b);
        while (i < s.size() && (s[i] == ' ' || s[i] == '\t')) i++;
        if (i < s.size() && isdigit((unsigned char)s[i])) {
            char* endp = nullptr;
            long v = strtol(s.c_str() + i, &endp, 0);
            size_t e = (size_t)(endp - s.c_str());
            while (e < s.size() && (s[e] == ' ' || s[e] == '\t' || s[e] == '\r')) e++;
            if (e >= s.size() || s[e] == '\n') defs[name] = v;
        }
        pos = i;
    }
}

bool collect_table(const std::string& s, std::vector<uint32_t>& out)
{
    size_t p = s.find("PosNoeudsVariable");
    while (p != std::string::npos) {
        size_t br = s.find_first_of("={;", p);
        if (br != std::string::npos && s[br] == '=') {
            size_t open = s.find('{', br);
            if (open == std::string::npos) return false;
            size_t close = s.find('}', open);
            if (close == std::string::npos) return false;
            const char* c = s.c_str() + open + 1;
            const char* end = s.c_str() + close;
            while (c < end) {
                while (c < end && !isdigit((unsigned char)*c)) c++;
                if (c >= end) break;
                char* e = nullptr;
                unsigned long v = strtoul(c, &e, 0);
                out.push_back((uint32_t)v);
                c = e;
            }
            return true;
        }
        p = s.find("PosNoeudsVariable", p + 1);
    }
    return false;
}

int fill_from_defs(ldpc_code_t* out, const std::map<std::string, long>& d)
{
    auto get = [&](const std::string& k, long& v) { auto it = d.find(k); if (it == d.end()) return false; v = it->second; return true; };
    long n, k, m, nb;
    if (!get("_N", n) || !get("_K", k) || !get("_M", m) || !get("NB_DEGRES", nb)) return LDPC_ERR_IO;
    if (nb < 1 || nb > LDPC_MAX_DEG_CLASSES) return LDPC_ERR_UNSUPPORTED;
    out->n = (int32_t)n; out->n_checks = (int32_t)k; out->m = (int32_t)m; out->nb_deg = (int32_t)nb;
    for (int i = 0; i < (int)nb; i++) {
        long dg, rw;
        if (!get("DEG_" + std::to_string(i + 1), dg) || !get("DEG_" + std::to_string(i + 1) + "_COMPUTATIONS", rw)) return LDPC_ERR_IO;
        out->deg[i] = (int32_t)dg; out->rows[i] = (int32_t)rw;
    }
    return LDPC_OK;
}

const char kMagic[8] = { 'L', 'D', 'P', 'C', 'T', 'B', 'L', '1' };

}  // namespace

extern "C" {

int ldpc_b200_check_code(const ldpc_code_t* c)
{
    if (!c || !c->pos || c->n <= 0 || c->n_checks <= 0 || c->m <= 0) return LDPC_ERR_INVALID;
    if (c->nb_deg < 1 || c->nb_deg > LDPC_MAX_DEG_CLASSES) return LDPC_ERR_INVALID;
    long long e = 0, r = 0;
    for (int i = 0; i < c->nb_deg; i++) {
        if (c->deg[i] < 1 || c->deg[i] > 64 || c->rows[i] < 0) return LDPC_ERR_INVALID;
        e += (long long)c->deg[i] * c->rows[i]; r += c->rows[i];
    }
    if (e != c->m || r != c->n_checks) return LDPC_ERR_INVALID;
    // indices in range, no variable twice in one row
    long long base = 0;
    for (int i = 0; i < c->nb_deg; i++)
        for (int row = 0; row < c->rows[i]; row++, base += c->deg[i])
            for (int j = 0; j < c->deg[i]; j++) {
                uint32_t v = c->pos[base + j];
                if (v >= (uint32_t)c->n) return LDPC_ERR_INVALID;
                for (int q = 0; q < j; q++) if (c->pos[base + q] == v) return LDPC_ERR_INVALID;
            }
    return LDPC_OK;
}

void ldpc_b200_free_code(ldpc_code_t* c)
{
    if (!c) return;
    free(c->pos);
    memset(c, 0, sizeof(*c));
}

int ldpc_b200_load_code_header(ldpc_code_t* out, const char* header_path, const char* table_path)
{
    if (!out || !header_path) return LDPC_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    std::string h, t;
    if (!read_file(header_path, h)) return LDPC_ERR_IO;
    strip_comments(h);
    std::map<std::string, long> defs;
    collect_defines(h, defs);
    int rc = fill_from_defs(out, defs);
    if (rc) return rc;
    std::vector<uint32_t> tab;
    bool ok = collect_table(h, tab);
    if (!ok && table_path) {
        if (!read_file(table_path, t)) return LDPC_ERR_IO;
        strip_comments(t);
        ok = collect_table(t, tab);
    }
    if (!ok || (long)tab.size() != out->m) return LDPC_ERR_IO;
    out->pos = (uint32_t*)malloc(sizeof(uint32_t) * tab.size());
    if (!out->pos) return LDPC_ERR_NOMEM;
    memcpy(out->pos, tab.data(), sizeof(uint32_t) * tab.size());
    rc = ldpc_b200_check_code(out);
    if (rc) ldpc_b200_free_code(out);
    return rc;
}

// binary table: magic[8] | int32 n, n_checks, m, nb_deg | int32 deg[8] | int32 rows[8] | int32 index_bytes (2|4) | indices
int ldpc_b200_save_code_table(const ldpc_code_t* c, const char* path)
{
    int rc = ldpc_b200_check_code(c);
    if (rc) return rc;
    FILE* f = fopen(path, "wb");
    if (!f) return LDPC_ERR_IO;
    int32_t hdr[4 + 2 * LDPC_MAX_DEG_CLASSES + 1];
    hdr[0] = c->n; hdr[1] = c->n_checks; hdr[2] = c->m; hdr[3] = c->nb_deg;
    for (int i = 0; i < LDPC_MAX_DEG_CLASSES; i++) { hdr[4 + i] = c->deg[i]; hdr[4 + LDPC_MAX_DEG_CLASSES + i] = c->rows[i]; }
    const int ib = (c->n <= 65536) ? 2 : 4;
    hdr[4 + 2 * LDPC_MAX_DEG_CLASSES] = ib;
    bool ok = fwrite(kMagic, 1, 8, f) == 8 && fwrite(hdr, sizeof(int32_t), sizeof(hdr) / sizeof(hdr[0]), f) == sizeof(hdr) / sizeof(hdr[0]);
    if (ok && ib == 2) {
        std::vector<uint16_t> v((size_t)c->m);
        for (int i = 0; i < c->m; i++) v[(size_t)i] = (uint16_t)c->pos[i];
        ok = fwrite(v.data(), 2, v.size(), f) == v.size();
    } else if (ok) ok = fwrite(c->pos, 4, (size_t)c->m, f) == (size_t)c->m;
    fclose(f);
    return ok ? LDPC_OK : LDPC_ERR_IO;
}

int ldpc_b200_load_code_table(ldpc_code_t* out, const char* path)
{
    if (!out || !path) return LDPC_ERR_INVALID;
    memset(out, 0, sizeof(*out));
    FILE* f = fopen(path, "rb");
    if (!f) return LDPC_ERR_IO;
    char magic[8];
    int32_t hdr[4 + 2 * LDPC_MAX_DEG_CLASSES + 1];
    bool ok = fread(magic, 1, 8, f) == 8 && !memcmp(magic, kMagic, 8) &&
              fread(hdr, sizeof(int32_t), sizeof(hdr) / sizeof(hdr[0]), f) == sizeof(hdr) / sizeof(hdr[0]);
    if (!ok) { fclose(f); return LDPC_ERR_IO; }
    out->n = hdr[0]; out->n_checks = hdr[1]; out->m = hdr[2]; out->nb_deg = hdr[3];
    for (int i = 0; i < LDPC_MAX_DEG_CLASSES; i++) { out->deg[i] = hdr[4 + i]; out->rows[i] = hdr[4 + LDPC_MAX_DEG_CLASSES + i]; }
    const int ib = hdr[4 + 2 * LDPC_MAX_DEG_CLASSES];
    if (out->m <= 0 || out->m > (1 << 28) || (ib != 2 && ib != 4)) { fclose(f); memset(out, 0, sizeof(*out)); return LDPC_ERR_IO; }
    out->pos = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)out->m);
    if (!out->pos) { fclose(f); return LDPC_ERR_NOMEM; }
    if (ib == 2) {
        std::vector<uint16_t> v((size_t)out->m);
        ok = fread(v.data(), 2, v.size(), f) == v.size();
        for (int i = 0; ok && i < out->m; i++) out->pos[i] = v[(size_t)i];
    } else ok = fread(out->pos, 4, (size_t)out->m, f) == (size_t)out->m;
    fclose(f);
    int rc = ok ? ldpc_b200_check_code(out) : LDPC_ERR_IO;
    if (rc) ldpc_b200_free_code(out);
    return rc;
}

// Level schedule: row r gets level 1 + max(level of any earlier row sharing a variable with r).
// Rows of one level touch disjoint variables, so updating them concurrently gives the same result as the reference order.
int ldpc_b200_level_schedule(const ldpc_code_t* c, int32_t* level_of_row)
{
    int rc = ldpc_b200_check_code(c);
    if (rc) return rc;
    std::vector<int32_t> last((size_t)c->n, -1);   // level of the last row that touched each variable
    int32_t levels = 0;
    long long e = 0; int row = 0;
    for (int i = 0; i < c->nb_deg; i++)
        for (int r = 0; r < c->rows[i]; r++, row++, e += c->deg[i]) {
            int32_t lv = 0;
            for (int j = 0; j < c->deg[i]; j++) { int32_t l = last[c->pos[e + j]] + 1; if (l > lv) lv = l; }
            for (int j = 0; j < c->deg[i]; j++) last[c->pos[e + j]] = lv;
            if (level_of_row) level_of_row[row] = lv;
            if (lv + 1 > levels) levels = lv + 1;
        }
    return levels;
}

}  // extern "C"
