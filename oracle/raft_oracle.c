/*
 * raft_oracle.c — CPU restatement of RAFT's fragmentation path (see raft_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY — never linked or called by the product path.
 * Parity status: PINNED against the unmodified reference binary (tests/golden/, oracle/_ref/raft).
 *
 * Written from the reference's observable semantics, not from its code:
 * coverage is a difference array + prefix sum instead of the reference's sorted-event sweep,
 * star removal is a set test instead of the reference's two-pointer walk, and fragments are
 * computed arithmetically.  Citations are to files under /root/reference.
 */
#include "raft_oracle.h"

#include <ctype.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ small utilities */

typedef struct {
    uint8_t *p;
    int64_t  len, cap;
    int      oom;
} buf_t;

static void buf_reserve(buf_t *b, int64_t extra)
{
    if (b->oom || b->len + extra <= b->cap) return;
    int64_t nc = b->cap ? b->cap : 4096;
    while (nc < b->len + extra) nc *= 2;
    uint8_t *np = (uint8_t *)realloc(b->p, (size_t)nc);
    if (!np) { b->oom = 1; return; }
    b->p = np; b->cap = nc;
}
static void buf_put(buf_t *b, const void *s, int64_t n)
{
    buf_reserve(b, n);
    if (b->oom) return;
    memcpy(b->p + b->len, s, (size_t)n);
    b->len += n;
}
static void buf_str(buf_t *b, const char *s) { buf_put(b, s, (int64_t)strlen(s)); }
static void buf_int(buf_t *b, long long v)
{ /* operator<<(int): plain decimal, '-' for negatives */
    char tmp[24];
    int  n = snprintf(tmp, sizeof tmp, "%lld", v);
    buf_put(b, tmp, n);
}

static uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

uint64_t orc_digest(const uint8_t *bytes, int64_t len, int64_t abs_offset)
{
    uint64_t d = 0;
    for (int64_t i = 0; i < len; i++)
        d += mix64((uint64_t)(abs_offset + i) * 257u + bytes[i] + 1u);
    return d;
}

void orc_default_params(orc_params_t *p)
{ /* param.hpp:18-31 */
    p->reso = 50; p->est_cov = 0; p->cov_mul = 1.5;
    p->repeat_length = 10000; p->interval_length = 10000;
    p->read_length = 20000; p->overlap_length = 500; p->flanking_length = 1000;
}

/* ------------------------------------------------------------------ name -> id map
 * chop.hpp:73-85 (addStringToMap) assigns dense ids in FASTA order; full-string keyed. */

typedef struct {
    int64_t        cap; /* power of two */
    int32_t       *slot; /* read id or -1 */
    const uint8_t *names;
    const int64_t *off;
} nmap_t;

static uint64_t hash_bytes(const uint8_t *s, int64_t n)
{
    uint64_t h = 0xcbf29ce484222325ULL;
    for (int64_t i = 0; i < n; i++) { h ^= s[i]; h *= 0x100000001b3ULL; }
    return mix64(h);
}

static int nmap_init(nmap_t *m, const orc_reads_t *r)
{
    int64_t cap = 16;
    while (cap < 2 * r->n_reads) cap *= 2;
    m->cap = cap; m->names = r->names; m->off = r->name_off;
    m->slot = (int32_t *)malloc(sizeof(int32_t) * (size_t)cap);
    if (!m->slot) return ORC_E_NOMEM;
    for (int64_t i = 0; i < cap; i++) m->slot[i] = -1;
    return ORC_OK;
}
static int32_t nmap_find(const nmap_t *m, const uint8_t *s, int64_t n)
{
    uint64_t h = hash_bytes(s, n);
    for (int64_t i = (int64_t)(h & (uint64_t)(m->cap - 1));; i = (i + 1) & (m->cap - 1)) {
        int32_t id = m->slot[i];
        if (id < 0) return -1;
        int64_t a = m->off[id], l = m->off[id + 1] - a;
        if (l == n && memcmp(m->names + a, s, (size_t)n) == 0) return id;
    }
}
/* returns existing id if name already present, else inserts */
static int32_t nmap_insert(nmap_t *m, int32_t id)
{
    const uint8_t *s = m->names + m->off[id];
    int64_t        n = m->off[id + 1] - m->off[id];
    uint64_t       h = hash_bytes(s, n);
    for (int64_t i = (int64_t)(h & (uint64_t)(m->cap - 1));; i = (i + 1) & (m->cap - 1)) {
        int32_t cur = m->slot[i];
        if (cur < 0) { m->slot[i] = id; return id; }
        int64_t a = m->off[cur], l = m->off[cur + 1] - a;
        if (l == n && memcmp(m->names + a, s, (size_t)n) == 0) return cur;
    }
}

/* ------------------------------------------------------------------ PAF numeric field
 * paf.hpp:62-81: strtol(q,&r,10) -> uint32_t; overlap.hpp:14-16 / chop.hpp:157-160: -> int.
 * The field is NUL-terminated at the tab (paf.hpp:59) so the scan never leaves [s, s+n). */
static int32_t paf_num(const uint8_t *s, int64_t n)
{
    int64_t i = 0;
    while (i < n && (s[i] == ' ' || (s[i] >= 9 && s[i] <= 13))) i++; /* isspace, "C" locale */
    int neg = 0;
    if (i < n && (s[i] == '+' || s[i] == '-')) { neg = (s[i] == '-'); i++; }
    unsigned long long acc = 0;
    int sat = 0;
    for (; i < n && s[i] >= '0' && s[i] <= '9'; i++) {
        unsigned d = (unsigned)(s[i] - '0');
        if (!sat) {
            if (acc > (ULLONG_MAX - d) / 10) sat = 1;
            else acc = acc * 10 + d;
        }
    }
    long v;
    if (neg) {
        if (sat || acc > (unsigned long long)LONG_MAX + 1ULL) v = LONG_MIN;
        else v = (long)(0ULL - acc);
    } else {
        if (sat || acc > (unsigned long long)LONG_MAX) v = LONG_MAX;
        else v = (long)acc;
    }
    return (int32_t)(uint32_t)v;
}

/* ------------------------------------------------------------------ simulated-read names
 * chop.hpp:99-106 regex, chop.hpp:14-70 field extraction. */
static int is_sim_name(const uint8_t *s, int64_t n)
{ /* ^read=[0-9]+,[a-z]+,position=[0-9]+-[0-9]+,length=[0-9]+,(.*) */
    int64_t i = 0;
#define LIT(str) do { int64_t l_ = (int64_t)strlen(str); if (i + l_ > n || memcmp(s + i, str, (size_t)l_)) return 0; i += l_; } while (0)
#define PLUS(cond) do { int64_t j_ = i; while (i < n && (cond)) i++; if (i == j_) return 0; } while (0)
    LIT("read=");      PLUS(s[i] >= '0' && s[i] <= '9');
    LIT(",");          PLUS(s[i] >= 'a' && s[i] <= 'z');
    LIT(",position="); PLUS(s[i] >= '0' && s[i] <= '9');
    LIT("-");          PLUS(s[i] >= '0' && s[i] <= '9');
    LIT(",length=");   PLUS(s[i] >= '0' && s[i] <= '9');
    LIT(",");
#undef LIT
#undef PLUS
    /* (.*) with std::regex ECMAScript: '.' does not match line terminators; names hold none */
    return 1;
}
typedef struct { int32_t start_pos, end_pos; int fwd, rev; const uint8_t *chr; int64_t chr_len; const uint8_t *last_comma; int64_t tail_len; const uint8_t *align; int64_t align_len; } sim_t;

static const uint8_t *find_ch(const uint8_t *s, const uint8_t *e, int c)
{ for (; s < e; s++) if (*s == c) return s; return NULL; }

static int sim_parse(const uint8_t *s, int64_t n, sim_t *o)
{
    const uint8_t *e = s + n;
    const uint8_t *c1 = find_ch(s, e, ',');            if (!c1) return -1;
    const uint8_t *c2 = find_ch(c1 + 1, e, ',');       if (!c2) return -1;
    const uint8_t *eq = find_ch(c1, e, '=');           if (!eq) return -1;
    const uint8_t *da = find_ch(eq + 1, e, '-');       if (!da) return -1;
    const uint8_t *d0 = find_ch(s, e, '-');            if (!d0) return -1;
    const uint8_t *c3 = find_ch(d0 + 1, e, ',');       if (!c3) return -1;
    const uint8_t *lc = e; while (lc > s && lc[-1] != ',') lc--; /* lc-1 is last comma */
    o->align = c1 + 1; o->align_len = c2 - (c1 + 1);                   /* chop.hpp:49-59 */
    {   /* atoi over [eq+1, da) (chop.hpp:25-35) and [d0+1, c3): digits only per the regex */
        long v = 0; for (const uint8_t *p = eq + 1; p < da && *p >= '0' && *p <= '9'; p++) v = v * 10 + (*p - '0');
        o->start_pos = (int32_t)v;
        v = 0; for (const uint8_t *p = d0 + 1; p < c3 && *p >= '0' && *p <= '9'; p++) v = v * 10 + (*p - '0');
        o->end_pos = (int32_t)v;                                          /* chop.hpp:37-47 */
    }
    o->fwd = (o->align_len == 7 && memcmp(o->align, "forward", 7) == 0);
    o->rev = (o->align_len == 7 && memcmp(o->align, "reverse", 7) == 0);
    o->chr = lc; o->chr_len = e - lc;                                     /* chop.hpp:61-70 */
    o->last_comma = lc - 1; o->tail_len = e - (lc - 1);                   /* name.substr(find_last_of(',')) */
    return 0;
}

/* ------------------------------------------------------------------ the path */

#define ALLOC(ptr, type, count) do { (ptr) = (type *)calloc((size_t)((count) > 0 ? (count) : 1), sizeof(type)); if (!(ptr)) { out->status = ORC_E_NOMEM; return ORC_E_NOMEM; } } while (0)

static int fail(orc_result_t *out, int st, int64_t idx) { out->status = st; out->bad_index = idx; return st; }

int orc_run(const orc_reads_t *rd, const uint8_t *paf, int64_t paf_len, const orc_params_t *prm,
            int flags, orc_result_t *out)
{
    memset(out, 0, sizeof *out);
    out->bad_index = -1;
    const int64_t n = rd->n_reads;
    const int     r = prm->reso, P = prm->interval_length, p = prm->repeat_length;
    const int     l = prm->read_length, v = prm->overlap_length, f = prm->flanking_length;

    /* repeat.hpp:32 divides by reso; repeat.hpp:125 with p<1 emits a repeat per low bin;
     * chop.hpp:209 divides by interval_length; chop.hpp:248,270 divides by div = l/P. */
    if (r < 1 || p < 1 || P < 1 || l < P) return fail(out, ORC_E_PARAM, -1);

    /* ---- a0: ids in FASTA order (chop.hpp:108); sim-mode detection on the first name (chop.hpp:99-106) */
    nmap_t map;
    if (nmap_init(&map, rd)) return fail(out, ORC_E_NOMEM, -1);
    for (int64_t i = 0; i < n; i++) {
        if (nmap_insert(&map, (int32_t)i) != (int32_t)i) { free(map.slot); return fail(out, ORC_E_DUP_NAME, i); }
    }
    out->real_reads = 1;
    if (n > 0 && is_sim_name(rd->names + rd->name_off[0], rd->name_off[1] - rd->name_off[0])) out->real_reads = 0;

    /* ---- a1: PAF lines (paf.hpp:89-99 + kseq.h:107-193), fields (paf.hpp:50-87) */
    int64_t n_lines = 1;
    for (int64_t i = 0; i < paf_len; i++) n_lines += (paf[i] == '\n');
    ALLOC(out->qid, int32_t, n_lines); ALLOC(out->tid, int32_t, n_lines);
    ALLOC(out->qs, int32_t, n_lines);  ALLOC(out->qe, int32_t, n_lines);
    ALLOC(out->ts, int32_t, n_lines);  ALLOC(out->te, int32_t, n_lines);
    ALLOC(out->strand, uint8_t, n_lines);

    int64_t N = 0;
    for (int64_t ls = 0; ls < paf_len;) {
        int64_t le = ls;
        while (le < paf_len && paf[le] != '\n') le++;
        int64_t next = le + 1;
        int64_t len  = le - ls;
        if (len > 1 && paf[le - 1] == '\r') len--; /* kseq.h:189-190 */
        /* split at tabs (paf.hpp:54-58) */
        int64_t fs[12], fe[12];
        int     t = 0;
        int64_t q = ls;
        for (int64_t i = ls; i <= ls + len; i++) {
            if (i < ls + len && paf[i] != '\t') continue;
            if (t < 12) { fs[t] = q; fe[t] = i; }
            t++; q = i + 1;
        }
        ls = next;
        if (t < 10) continue; /* paf.hpp:84-85, 96-98 */

        /* std::string(r.qn): up to the first NUL (chop.hpp:162-163) */
        int64_t qn_len = fe[0] - fs[0], tn_len = fe[5] - fs[5];
        { const uint8_t *z = (const uint8_t *)memchr(paf + fs[0], 0, (size_t)qn_len); if (z) qn_len = z - (paf + fs[0]); }
        { const uint8_t *z = (const uint8_t *)memchr(paf + fs[5], 0, (size_t)tn_len); if (z) tn_len = z - (paf + fs[5]); }
        int32_t qi = nmap_find(&map, paf + fs[0], qn_len);
        int32_t ti = nmap_find(&map, paf + fs[5], tn_len);
        if (qi < 0 || ti < 0) { free(map.slot); return fail(out, ORC_E_UNKNOWN_NAME, N); }
        out->qid[N] = qi; out->tid[N] = ti;
        out->qs[N] = paf_num(paf + fs[2], fe[2] - fs[2]);
        out->qe[N] = paf_num(paf + fs[3], fe[3] - fs[3]);
        out->strand[N] = (fe[4] > fs[4] && paf[fs[4]] == '-'); /* paf.hpp:68-69 */
        out->ts[N] = paf_num(paf + fs[7], fe[7] - fs[7]);
        out->te[N] = paf_num(paf + fs[8], fe[8] - fs[8]);
        N++;
    }
    free(map.slot);
    out->n_rec = N;

    /* ---- a2: symmetric flag (chop.hpp:171-184): some record k>=1 mirrors record 0 */
    int S = 0;
    for (int64_t k = 1; k < N && !S; k++)
        S = out->qid[0] == out->tid[k] && out->tid[0] == out->qid[k] && out->qs[0] == out->ts[k] &&
            out->qe[0] == out->te[k] && out->ts[0] == out->qs[k] && out->te[0] == out->qe[k];
    out->symmetric = S;

    /* ---- a3: coverage (repeat.hpp:28-79) as difference array per read */
    ALLOC(out->bin_off, int64_t, n + 1);
    for (int64_t i = 0; i < n; i++) {
        int64_t L = rd->seq_off[i + 1] - rd->seq_off[i];
        out->bin_off[i + 1] = out->bin_off[i] + (L + r - 1) / r; /* repeat.hpp:32-37 */
    }
    const int64_t B = out->bin_off[n];
    ALLOC(out->cov, int32_t, B + 1);
    int32_t *diff;
    ALLOC(diff, int32_t, B + n + 1); /* one sentinel slot per read */
/* repeat.hpp:62-77: an event (start, end-1) is handled at bin i = floor(start/reso) (i = 0 for a
 * negative start) and increments every k >= i with end-1 >= k*reso.  Outside [0, nb) the reference
 * writes out of bounds -> ORC_E_RANGE. */
#define ADD_INTERVAL(read, s_, e_, recidx) do {                                                         \
        int64_t nb_ = out->bin_off[(read) + 1] - out->bin_off[(read)];                                 \
        int64_t base_ = out->bin_off[(read)] + (read);                                                  \
        int64_t s__ = (s_), em_ = (int64_t)(e_) - 1;                                                    \
        int64_t lo_ = (s__ < 0 ? 0 : s__) / r;                                                          \
        if (em_ >= lo_ * r) {                                                                           \
            int64_t hi_ = em_ / r;                                                                      \
            if (hi_ >= nb_) { free(diff); return fail(out, ORC_E_RANGE, (recidx)); }                   \
            diff[base_ + lo_]++; diff[base_ + hi_ + 1]--;                                               \
        }                                                                                               \
    } while (0)
    for (int64_t k = 0; k < N; k++) {
        int32_t q = out->qid[k], t = out->tid[k];
        ADD_INTERVAL(q, out->qs[k], out->qe[k], k);                 /* repeat.hpp:50-53, chop.hpp:165 */
        if (!S && t != q) ADD_INTERVAL(t, out->ts[k], out->te[k], k); /* repeat.hpp:54-57, chop.hpp:166-169 */
    }
    for (int64_t i = 0; i < n; i++) {
        int64_t nb = out->bin_off[i + 1] - out->bin_off[i], base = out->bin_off[i] + i;
        int32_t run = 0;
        for (int64_t k = 0; k < nb; k++) { run += diff[base + k]; out->cov[out->bin_off[i] + k] = run; }
    }
    free(diff);

    /* ---- a4: repeats (repeat.hpp:89-171) */
    const int32_t H = (int32_t)(prm->est_cov * prm->cov_mul); /* repeat.hpp:89-90: int * double -> int */
    out->high_cov = H;
    ALLOC(out->rep_off, int64_t, n + 1);
    int64_t rep_cap = 1024, n_rep = 0;
    ALLOC(out->rep_s, int32_t, rep_cap); ALLOC(out->rep_e, int32_t, rep_cap);
    for (int64_t i = 0; i < n; i++) {
        int64_t L = rd->seq_off[i + 1] - rd->seq_off[i];
        int64_t nb = out->bin_off[i + 1] - out->bin_off[i];
        const int32_t *c = out->cov + out->bin_off[i];
        out->total_read_len += L;
        out->rep_off[i] = n_rep;
        for (int64_t k = 0; k < nb;) {
            out->total_cov += c[k]; out->total_windows = (int32_t)((uint32_t)out->total_windows + 1u);
            if (c[k] < H) { k++; continue; }
            int64_t k2 = k;
            while (k2 + 1 < nb && c[k2 + 1] >= H) { k2++; out->total_cov += c[k2]; out->total_windows = (int32_t)((uint32_t)out->total_windows + 1u); }
            int64_t start = k * (int64_t)r, end = (k2 + 1) * (int64_t)r;
            if (end - start >= p) { /* repeat.hpp:125,150 */
                out->total_repeat_len += end - start;
                int64_t s = start - f, e = end + f;
                if (s <= 0) s = 0;
                if (e >= L) e = L;
                if (n_rep == rep_cap) {
                    rep_cap *= 2;
                    out->rep_s = (int32_t *)realloc(out->rep_s, sizeof(int32_t) * (size_t)rep_cap);
                    out->rep_e = (int32_t *)realloc(out->rep_e, sizeof(int32_t) * (size_t)rep_cap);
                    if (!out->rep_s || !out->rep_e) return fail(out, ORC_E_NOMEM, i);
                }
                out->rep_s[n_rep] = (int32_t)s; out->rep_e[n_rep] = (int32_t)e; n_rep++;
            }
            k = k2 + 1;
        }
    }
    out->rep_off[n] = n_rep;

    /* ---- a5: stars and fragments (chop.hpp:198-323) */
    const int div = l / P; /* chop.hpp:248 */
    int64_t frag_cap = n + 1024, G = 0;
    ALLOC(out->frag_read, int32_t, frag_cap); ALLOC(out->frag_a, int32_t, frag_cap); ALLOC(out->frag_b, int32_t, frag_cap);
    int64_t  fin_cap = 1024;
    int32_t *fin = (int32_t *)malloc(sizeof(int32_t) * (size_t)fin_cap);
    if (!fin) return fail(out, ORC_E_NOMEM, -1);
    for (int64_t i = 0; i < n; i++) {
        int64_t L = rd->seq_off[i + 1] - rd->seq_off[i];
        int64_t parts = L / P, nstars = parts + 1 + (L % P != 0); /* chop.hpp:209-223 */
        if (nstars > fin_cap) { fin_cap = nstars * 2; fin = (int32_t *)realloc(fin, sizeof(int32_t) * (size_t)fin_cap); if (!fin) return fail(out, ORC_E_NOMEM, i); }
        int64_t nf = 0;
        for (int64_t j = 0; j < nstars; j++) {
            int64_t x = (j <= parts) ? j * P : L;
            int     keep = 1;
            if (j > 0 && j < nstars - 1) /* chop.hpp:225-246: first and last star always survive */
                for (int64_t q = out->rep_off[i]; q < out->rep_off[i + 1]; q++)
                    if (out->rep_s[q] <= x && x <= out->rep_e[q]) { keep = 0; break; }
            if (keep) fin[nf++] = (int32_t)x;
        }
        int64_t F = (nf <= div + 1) ? 1 : 1 + (nf - div - 1) / div + ((nf - div - 1) % div != 0); /* chop.hpp:250,270-276 */
        if (G + F > frag_cap) {
            frag_cap = (G + F) * 2;
            out->frag_read = (int32_t *)realloc(out->frag_read, sizeof(int32_t) * (size_t)frag_cap);
            out->frag_a = (int32_t *)realloc(out->frag_a, sizeof(int32_t) * (size_t)frag_cap);
            out->frag_b = (int32_t *)realloc(out->frag_b, sizeof(int32_t) * (size_t)frag_cap);
            if (!out->frag_read || !out->frag_a || !out->frag_b) return fail(out, ORC_E_NOMEM, i);
        }
        if (F == 1 && nf <= div + 1) {
            out->frag_read[G] = (int32_t)i; out->frag_a[G] = 0; out->frag_b[G] = (int32_t)L; G++; /* chop.hpp:261-266 */
        } else {
            for (int64_t j = 1; j <= F; j++) { /* chop.hpp:280-320 */
                int64_t a = fin[(j - 1) * div] - (j > 1 ? v : 0);
                int64_t b = (j == F) ? fin[nf - 1] : fin[j * div];
                if (a < 0 || a > L) { free(fin); return fail(out, ORC_E_NEG_START, i); }
                out->frag_read[G] = (int32_t)i; out->frag_a[G] = (int32_t)a; out->frag_b[G] = (int32_t)b; G++;
            }
        }
    }
    free(fin);
    out->n_frag = G;

    if (flags & ORC_NO_TEXT) return ORC_OK;

    /* ---- text outputs */
    buf_t cov = {0}, rep = {0}, bed = {0}, fa = {0};
    for (int64_t i = 0; i < n; i++) { /* repeat.hpp:105-108 */
        int64_t nb = out->bin_off[i + 1] - out->bin_off[i];
        buf_str(&cov, "read "); buf_int(&cov, i); buf_str(&cov, " ");
        for (int64_t k = 0; k < nb; k++) {
            buf_int(&cov, k * (int64_t)r); buf_str(&cov, ",");
            buf_int(&cov, out->cov[out->bin_off[i] + k]); buf_str(&cov, " ");
        }
        buf_str(&cov, "\n");
    }
    for (int64_t i = 0; i < n; i++) { /* repeat.hpp:180-203 */
        sim_t sm; int have_sim = 0;
        if (!out->real_reads) have_sim = sim_parse(rd->names + rd->name_off[i], rd->name_off[i + 1] - rd->name_off[i], &sm) == 0;
        buf_str(&rep, "read "); buf_int(&rep, i); buf_str(&rep, ", ");
        for (int64_t q = out->rep_off[i]; q < out->rep_off[i + 1]; q++) {
            buf_int(&rep, out->rep_s[q]); buf_str(&rep, ","); buf_int(&rep, out->rep_e[q]); buf_str(&rep, "    ");
            if (have_sim && (sm.fwd || sm.rev)) { /* repeat.hpp:187-199 */
                buf_put(&bed, sm.chr, sm.chr_len); buf_str(&bed, "\t");
                buf_int(&bed, sm.fwd ? sm.start_pos + out->rep_s[q] : sm.end_pos - out->rep_e[q]); buf_str(&bed, "\t");
                buf_int(&bed, sm.fwd ? sm.start_pos + out->rep_e[q] : sm.end_pos - out->rep_s[q]); buf_str(&bed, "\n");
            }
        }
        buf_str(&rep, "\n");
    }
    for (int64_t g = 0; g < G; g++) { /* chop.hpp:250-322 */
        int64_t        i = out->frag_read[g], a = out->frag_a[g], b = out->frag_b[g];
        int64_t        L = rd->seq_off[i + 1] - rd->seq_off[i];
        const uint8_t *nm = rd->names + rd->name_off[i];
        int64_t        nl = rd->name_off[i + 1] - rd->name_off[i];
        int            whole = (g == 0 || out->frag_read[g - 1] != i) && (g + 1 == G || out->frag_read[g + 1] != i);
        if (out->real_reads) {
            buf_str(&fa, ">read="); buf_int(&fa, g + 1); buf_str(&fa, ","); buf_put(&fa, nm, nl);
            buf_str(&fa, ",pos_on_original_read="); buf_int(&fa, a); buf_str(&fa, "-"); buf_int(&fa, b); buf_str(&fa, "\n");
        } else {
            sim_t sm;
            if (sim_parse(nm, nl, &sm) == 0) {
                if (whole) { /* chop.hpp:252-258 */
                    buf_str(&fa, ">read="); buf_int(&fa, g + 1); buf_str(&fa, ","); buf_put(&fa, sm.align, sm.align_len);
                    buf_str(&fa, ",position="); buf_int(&fa, sm.start_pos); buf_str(&fa, "-"); buf_int(&fa, sm.end_pos);
                    buf_str(&fa, ",length="); buf_int(&fa, L); buf_put(&fa, sm.last_comma, sm.tail_len); buf_str(&fa, "\n");
                } else if (sm.fwd || sm.rev) { /* chop.hpp:293-310 */
                    buf_str(&fa, ">read="); buf_int(&fa, g + 1); buf_str(&fa, ","); buf_put(&fa, sm.align, sm.align_len);
                    buf_str(&fa, ",position=");
                    buf_int(&fa, sm.fwd ? sm.start_pos + a : sm.end_pos - b); buf_str(&fa, "-");
                    buf_int(&fa, sm.fwd ? sm.start_pos + b : sm.end_pos - a);
                    buf_str(&fa, ",length="); buf_int(&fa, b - a); buf_put(&fa, sm.last_comma, sm.tail_len); buf_str(&fa, "\n");
                }
            }
        }
        int64_t cnt = b - a; /* substr(a, b-a) clips at the end of the string (chop.hpp:318) */
        if (a + cnt > L) cnt = L - a;
        if (cnt < 0) cnt = 0;
        buf_put(&fa, rd->seq + rd->seq_off[i] + a, cnt);
        buf_str(&fa, "\n");
    }
    if (cov.oom || rep.oom || bed.oom || fa.oom) { free(cov.p); free(rep.p); free(bed.p); free(fa.p); return fail(out, ORC_E_NOMEM, -1); }
    out->cov_txt = cov.p; out->cov_txt_len = cov.len;
    out->rep_txt = rep.p; out->rep_txt_len = rep.len;
    out->bed_txt = bed.p; out->bed_txt_len = bed.len;
    out->fasta = fa.p;    out->fasta_len = fa.len;
    return ORC_OK;
}

void orc_free(orc_result_t *r)
{
    free(r->qid); free(r->tid); free(r->qs); free(r->qe); free(r->ts); free(r->te); free(r->strand);
    free(r->bin_off); free(r->cov); free(r->rep_off); free(r->rep_s); free(r->rep_e);
    free(r->frag_read); free(r->frag_a); free(r->frag_b);
    free(r->cov_txt); free(r->rep_txt); free(r->bed_txt); free(r->fasta);
    memset(r, 0, sizeof *r);
}

/* ------------------------------------------------------------------ FASTA/FASTQ records
 * kseq.h:240-298 over an in-memory buffer; chop.hpp:108-121 keeps name (to first isspace) and
 * strlen(seq).  Lines lose a trailing '\r' only when the accumulated string is longer than one
 * byte (kseq.h:189-190). */
static int64_t take_line(const uint8_t *t, int64_t len, int64_t pos, buf_t *dst)
{ /* append the rest of the current line to dst; returns position after the '\n' */
    int64_t e = pos;
    while (e < len && t[e] != '\n') e++;
    buf_put(dst, t + pos, e - pos);
    if (dst->len > 1 && dst->p[dst->len - 1] == '\r') dst->len--;
    return e < len ? e + 1 : len;
}

int64_t orc_parse_fasta(const uint8_t *t, int64_t len, orc_fasta_t *out)
{
    memset(out, 0, sizeof *out);
    buf_t   seq = {0}, names = {0}, soff = {0}, noff = {0}, cur = {0}, qual = {0};
    int64_t zero = 0, n = 0, pos = 0;
    int     last_char = 0;
    int64_t rc = 0;
    buf_put(&soff, &zero, 8); buf_put(&noff, &zero, 8);
    for (;;) {
        if (!last_char) { /* kseq.h:246-252 */
            while (pos < len && t[pos] != '>' && t[pos] != '@') pos++;
            if (pos >= len) break;
            last_char = t[pos++];
        }
        /* name: up to first isspace (kseq.h:254); EOF right after the marker ends the stream */
        if (pos >= len) break;
        int64_t ns = pos;
        while (pos < len && !(t[pos] == ' ' || (t[pos] >= 9 && t[pos] <= 13))) pos++;
        int64_t ne = pos;
        int     delim = pos < len ? t[pos++] : -1;
        if (delim != '\n' && delim != -1) { /* comment: rest of line (kseq.h:256-257) */
            while (pos < len && t[pos] != '\n') pos++;
            if (pos < len) pos++;
        }
        cur.len = 0;
        int c = -1;
        while (pos < len) { /* kseq.h:263-269 */
            c = t[pos++];
            if (c == '>' || c == '+' || c == '@') break;
            if (c == '\n') { c = -1; continue; }
            uint8_t ch = (uint8_t)c;
            buf_put(&cur, &ch, 1);
            pos = take_line(t, len, pos, &cur);
            c = -1;
        }
        last_char = (c == '>' || c == '@') ? c : 0;
        if (c == '+') { /* FASTQ (kseq.h:281-296) */
            while (pos < len && t[pos] != '\n') pos++;
            if (pos >= len) { rc = ORC_E_FASTQ; break; }
            pos++;
            qual.len = 0;
            while (pos < len && qual.len < cur.len) pos = take_line(t, len, pos, &qual);
            last_char = 0;
            if (qual.len != cur.len) { rc = ORC_E_FASTQ; break; }
        }
        /* strlen semantics (chop.hpp:112): stop at an embedded NUL */
        int64_t sl = cur.len;
        { const uint8_t *z = cur.len ? (const uint8_t *)memchr(cur.p, 0, (size_t)cur.len) : NULL; if (z) sl = z - cur.p; }
        int64_t nl = ne - ns;
        { const uint8_t *z = nl ? (const uint8_t *)memchr(t + ns, 0, (size_t)nl) : NULL; if (z) nl = z - (t + ns); }
        buf_put(&seq, cur.p, sl);
        buf_put(&names, t + ns, nl);
        buf_put(&soff, &seq.len, 8); buf_put(&noff, &names.len, 8);
        n++;
    }
    free(cur.p); free(qual.p);
    /* a truncated FASTQ record stops the loop but keeps earlier reads (chop.hpp:97: l >= 0) */
    (void)rc;
    if (seq.oom || names.oom || soff.oom || noff.oom) { free(seq.p); free(names.p); free(soff.p); free(noff.p); return ORC_E_NOMEM; }
    if (!seq.p) seq.p = (uint8_t *)calloc(1, 1);
    if (!names.p) names.p = (uint8_t *)calloc(1, 1);
    out->n_reads = n;
    out->seq = seq.p; out->names = names.p;
    out->seq_off = (int64_t *)soff.p; out->name_off = (int64_t *)noff.p;
    return n;
}

void orc_free_fasta(orc_fasta_t *f)
{
    free(f->seq_off); free(f->seq); free(f->name_off); free(f->names);
    memset(f, 0, sizeof *f);
}

/* ------------------------------------------------------------------ split_naive (split_naive.cpp:22-37) */
int64_t orc_split_naive(const orc_reads_t *rd, int32_t sublen, uint8_t **out)
{
    buf_t b = {0};
    if (sublen < 1) return ORC_E_PARAM;
    for (int64_t i = 0; i < rd->n_reads; i++) {
        int64_t L = rd->seq_off[i + 1] - rd->seq_off[i];
        int64_t k = 1;
        for (int64_t a = 0; a < L; a += sublen, k++) { /* split_naive.cpp:27-29 */
            int64_t n = a + sublen < L ? sublen : L - a;
            buf_str(&b, ">"); buf_put(&b, rd->names + rd->name_off[i], rd->name_off[i + 1] - rd->name_off[i]);
            buf_str(&b, "_"); buf_int(&b, k); buf_str(&b, "\n");          /* split_naive.cpp:32 */
            buf_put(&b, rd->seq + rd->seq_off[i] + a, n); buf_str(&b, "\n");
        }
    }
    if (b.oom) { free(b.p); return ORC_E_NOMEM; }
    if (!b.p) b.p = (uint8_t *)calloc(1, 1);
    *out = b.p;
    return b.len;
}
