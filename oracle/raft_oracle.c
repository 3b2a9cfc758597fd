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

/* Text sink: every output byte of the three files goes through sink_put, whichever way the caller
 * consumes it -- materialised (orc_run), only counted, or digested at its absolute file offset
 * (orc_run_digest).  One emission code path for all three, so pinning orc_run's bytes pins the digests. */
enum { SINK_BUF = 0, SINK_COUNT = 1, SINK_DIGEST = 2 };
typedef struct {
    int      mode;
    buf_t   *b;   /* SINK_BUF */
    int64_t  pos; /* absolute offset of the next byte */
    uint64_t dig; /* SINK_DIGEST */
} sink_t;

static void sink_put(sink_t *s, const void *p, int64_t n)
{
    if (s->mode == SINK_BUF) buf_put(s->b, p, n);
    else if (s->mode == SINK_DIGEST) s->dig += orc_digest((const uint8_t *)p, n, s->pos);
    s->pos += n;
}
static void sink_str(sink_t *s, const char *t) { sink_put(s, t, (int64_t)strlen(t)); }
static void sink_int(sink_t *s, long long v)
{ /* operator<<(int): plain decimal, '-' for negatives */
    char               tmp[24];
    int                k = 24;
    unsigned long long u = v < 0 ? 0ULL - (unsigned long long)v : (unsigned long long)v;
    do { tmp[--k] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) tmp[--k] = '-';
    sink_put(s, tmp + k, 24 - k);
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

/* ------------------------------------------------------------------ the path
 * The stages are written as helpers over one byte range of the PAF / one read, so that the sequential
 * orc_run (which keeps every table and the text) and the multi-threaded orc_run_digest (which keeps
 * only coverage and digests the text as it is produced) execute the same code. */

#define ALLOC(ptr, type, count) do { (ptr) = (type *)calloc((size_t)((count) > 0 ? (count) : 1), sizeof(type)); if (!(ptr)) { out->status = ORC_E_NOMEM; return ORC_E_NOMEM; } } while (0)

static int fail(orc_result_t *out, int st, int64_t idx) { out->status = st; out->bad_index = idx; return st; }

typedef struct { int32_t *qid, *tid, *qs, *qe, *ts, *te; uint8_t *strand; } cols_t;

/* ---- a1: PAF lines (paf.hpp:89-99 + kseq.h:107-193) and fields (paf.hpp:50-87) of the lines that START in [lo, hi).
 * Records are appended at c[*N...]; returns the (range-relative) index of the first record with an unknown name, or -1. */
static int64_t parse_paf_range(const uint8_t *paf, int64_t lo, int64_t hi, int64_t paf_len, const nmap_t *map, cols_t *c, int64_t *N_out)
{
    int64_t N = 0;
    for (int64_t ls = lo; ls < hi;) {
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
        int32_t qi = nmap_find(map, paf + fs[0], qn_len);
        int32_t ti = nmap_find(map, paf + fs[5], tn_len);
        if (qi < 0 || ti < 0) { *N_out = N; return N; }
        c->qid[N] = qi; c->tid[N] = ti;
        c->qs[N] = paf_num(paf + fs[2], fe[2] - fs[2]);
        c->qe[N] = paf_num(paf + fs[3], fe[3] - fs[3]);
        c->strand[N] = (fe[4] > fs[4] && paf[fs[4]] == '-'); /* paf.hpp:68-69 */
        c->ts[N] = paf_num(paf + fs[7], fe[7] - fs[7]);
        c->te[N] = paf_num(paf + fs[8], fe[8] - fs[8]);
        N++;
    }
    *N_out = N;
    return -1;
}

/* ---- a3 (one interval): repeat.hpp:62-77.  An event (start, end-1) is handled at bin i = floor(start/reso) (i = 0
 * for a negative start) and increments every k >= i with end-1 >= k*reso: +1 at the first bin, -1 one past the last
 * in the read's difference slots (nb bins + one sentinel).  Returns 0 when it would leave [0, nb): the reference
 * writes out of bounds there.  `atomic`: several threads add into the same array. */
static int add_interval(int32_t *slots, int64_t nb, int64_t s, int64_t e, int r, int atomic)
{
    int64_t em = e - 1;
    int64_t lo = (s < 0 ? 0 : s) / r;
    if (em < lo * r) return 1;
    int64_t hi = em / r;
    if (hi >= nb) return 0;
    if (atomic) { __atomic_fetch_add(&slots[lo], 1, __ATOMIC_RELAXED); __atomic_fetch_add(&slots[hi + 1], -1, __ATOMIC_RELAXED); }
    else { slots[lo]++; slots[hi + 1]--; }
    return 1;
}

typedef struct { int32_t *s, *e; int64_t n, cap; } reps_t;
static int reps_push(reps_t *v, int32_t s, int32_t e)
{
    if (v->n == v->cap) {
        int64_t  nc = v->cap ? v->cap * 2 : 64;
        int32_t *ns = (int32_t *)realloc(v->s, sizeof(int32_t) * (size_t)nc);
        if (ns) v->s = ns;
        int32_t *ne = (int32_t *)realloc(v->e, sizeof(int32_t) * (size_t)nc);
        if (ne) v->e = ne;
        if (!ns || !ne) return 0;
        v->cap = nc;
    }
    v->s[v->n] = s; v->e[v->n] = e; v->n++;
    return 1;
}
typedef struct { int64_t total_cov; uint32_t total_windows; int64_t total_repeat_len; } rstats_t;

/* ---- a4 (one read): repeats appended to v (repeat.hpp:111-168), stats accumulated (repeat.hpp:93-97).  0 on OOM. */
static int read_repeats(const int32_t *c, int64_t nb, int64_t L, int r, int p, int f, int32_t H, reps_t *v, rstats_t *st)
{
    for (int64_t k = 0; k < nb;) {
        st->total_cov += c[k]; st->total_windows += 1u;
        if (c[k] < H) { k++; continue; }
        int64_t k2 = k;
        while (k2 + 1 < nb && c[k2 + 1] >= H) { k2++; st->total_cov += c[k2]; st->total_windows += 1u; }
        int64_t start = k * (int64_t)r, end = (k2 + 1) * (int64_t)r;
        if (end - start >= p) { /* repeat.hpp:125,150 */
            st->total_repeat_len += end - start;
            int64_t s = start - f, e = end + f;
            if (s <= 0) s = 0;
            if (e >= L) e = L;
            if (!reps_push(v, (int32_t)s, (int32_t)e)) return 0;
        }
        k = k2 + 1;
    }
    return 1;
}

typedef struct { int32_t *fin; int64_t fin_cap; int32_t *a, *b; int64_t n, cap; } frags_t;

/* ---- a5 (one read): stars and fragments (chop.hpp:198-323); the read's fragments REPLACE the contents of fr.
 * Returns ORC_OK, ORC_E_NEG_START or ORC_E_NOMEM. */
static int read_frags(int64_t L, int P, int div, int v, const int32_t *rep_s, const int32_t *rep_e, int64_t nrep, frags_t *fr)
{
    int64_t parts = L / P, nstars = parts + 1 + (L % P != 0); /* chop.hpp:209-223 */
    if (nstars > fr->fin_cap) {
        int32_t *nf = (int32_t *)realloc(fr->fin, sizeof(int32_t) * (size_t)(nstars * 2));
        if (!nf) return ORC_E_NOMEM;
        fr->fin = nf; fr->fin_cap = nstars * 2;
    }
    int32_t *fin = fr->fin;
    int64_t  nf = 0;
    for (int64_t j = 0; j < nstars; j++) {
        int64_t x = (j <= parts) ? j * P : L;
        int     keep = 1;
        if (j > 0 && j < nstars - 1) /* chop.hpp:225-246: first and last star always survive */
            for (int64_t q = 0; q < nrep; q++)
                if (rep_s[q] <= x && x <= rep_e[q]) { keep = 0; break; }
        if (keep) fin[nf++] = (int32_t)x;
    }
    int64_t F = (nf <= div + 1) ? 1 : 1 + (nf - div - 1) / div + ((nf - div - 1) % div != 0); /* chop.hpp:250,270-276 */
    if (F > fr->cap) {
        int32_t *na = (int32_t *)realloc(fr->a, sizeof(int32_t) * (size_t)(F * 2));
        if (na) fr->a = na;
        int32_t *nb_ = (int32_t *)realloc(fr->b, sizeof(int32_t) * (size_t)(F * 2));
        if (nb_) fr->b = nb_;
        if (!na || !nb_) return ORC_E_NOMEM;
        fr->cap = F * 2;
    }
    fr->n = 0;
    if (F == 1 && nf <= div + 1) {
        fr->a[0] = 0; fr->b[0] = (int32_t)L; fr->n = 1; /* chop.hpp:261-266 */
    } else {
        for (int64_t j = 1; j <= F; j++) { /* chop.hpp:280-320 */
            int64_t a = fin[(j - 1) * div] - (j > 1 ? v : 0);
            int64_t b = (j == F) ? fin[nf - 1] : fin[j * div];
            if (a < 0 || a > L) return ORC_E_NEG_START;
            fr->a[fr->n] = (int32_t)a; fr->b[fr->n] = (int32_t)b; fr->n++;
        }
    }
    return ORC_OK;
}

/* ---- text of one read ---- */
static void emit_cov_line(sink_t *s, int64_t i, const int32_t *c, int64_t nb, int r)
{ /* repeat.hpp:105-108 */
    sink_str(s, "read "); sink_int(s, i); sink_str(s, " ");
    for (int64_t k = 0; k < nb; k++) {
        sink_int(s, k * (int64_t)r); sink_str(s, ",");
        sink_int(s, c[k]); sink_str(s, " ");
    }
    sink_str(s, "\n");
}
static void emit_rep_line(sink_t *rep, sink_t *bed, int64_t i, const int32_t *rs, const int32_t *re, int64_t nrep, int real_reads,
                          const uint8_t *name, int64_t name_len)
{ /* repeat.hpp:180-203 */
    sim_t sm; int have_sim = 0;
    if (!real_reads) have_sim = sim_parse(name, name_len, &sm) == 0;
    sink_str(rep, "read "); sink_int(rep, i); sink_str(rep, ", ");
    for (int64_t q = 0; q < nrep; q++) {
        sink_int(rep, rs[q]); sink_str(rep, ","); sink_int(rep, re[q]); sink_str(rep, "    ");
        if (have_sim && (sm.fwd || sm.rev)) { /* repeat.hpp:187-199 */
            sink_put(bed, sm.chr, sm.chr_len); sink_str(bed, "\t");
            sink_int(bed, sm.fwd ? sm.start_pos + rs[q] : sm.end_pos - re[q]); sink_str(bed, "\t");
            sink_int(bed, sm.fwd ? sm.start_pos + re[q] : sm.end_pos - rs[q]); sink_str(bed, "\n");
        }
    }
    sink_str(rep, "\n");
}
/* record g (0-based, global) = bases [a, b) of a read of length L; whole = the read's only record */
static void emit_fasta_record(sink_t *fa, int64_t g, int whole, int64_t a, int64_t b, int64_t L, const uint8_t *nm, int64_t nl,
                              const uint8_t *seq, int real_reads)
{ /* chop.hpp:250-322 */
    if (real_reads) {
        sink_str(fa, ">read="); sink_int(fa, g + 1); sink_str(fa, ","); sink_put(fa, nm, nl);
        sink_str(fa, ",pos_on_original_read="); sink_int(fa, a); sink_str(fa, "-"); sink_int(fa, b); sink_str(fa, "\n");
    } else {
        sim_t sm;
        if (sim_parse(nm, nl, &sm) == 0) {
            if (whole) { /* chop.hpp:252-258 */
                sink_str(fa, ">read="); sink_int(fa, g + 1); sink_str(fa, ","); sink_put(fa, sm.align, sm.align_len);
                sink_str(fa, ",position="); sink_int(fa, sm.start_pos); sink_str(fa, "-"); sink_int(fa, sm.end_pos);
                sink_str(fa, ",length="); sink_int(fa, L); sink_put(fa, sm.last_comma, sm.tail_len); sink_str(fa, "\n");
            } else if (sm.fwd || sm.rev) { /* chop.hpp:293-310 */
                sink_str(fa, ">read="); sink_int(fa, g + 1); sink_str(fa, ","); sink_put(fa, sm.align, sm.align_len);
                sink_str(fa, ",position=");
                sink_int(fa, sm.fwd ? sm.start_pos + a : sm.end_pos - b); sink_str(fa, "-");
                sink_int(fa, sm.fwd ? sm.start_pos + b : sm.end_pos - a);
                sink_str(fa, ",length="); sink_int(fa, b - a); sink_put(fa, sm.last_comma, sm.tail_len); sink_str(fa, "\n");
            }
        }
    }
    int64_t cnt = b - a; /* substr(a, b-a) clips at the end of the string (chop.hpp:318) */
    if (a + cnt > L) cnt = L - a;
    if (cnt < 0) cnt = 0;
    sink_put(fa, seq + a, cnt);
    sink_str(fa, "\n");
}

/* a0: ids in FASTA order (chop.hpp:108); a duplicate name aliases ids in the reference -> error */
static int build_map(nmap_t *map, const orc_reads_t *rd, int64_t *dup)
{
    if (nmap_init(map, rd)) return ORC_E_NOMEM;
    for (int64_t i = 0; i < rd->n_reads; i++)
        if (nmap_insert(map, (int32_t)i) != (int32_t)i) { free(map->slot); *dup = i; return ORC_E_DUP_NAME; }
    return ORC_OK;
}

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

    /* ---- a0: name map; sim-mode detection on the first name (chop.hpp:99-106) */
    nmap_t  map;
    int64_t dup = -1;
    int     st = build_map(&map, rd, &dup);
    if (st) return fail(out, st, dup);
    out->real_reads = 1;
    if (n > 0 && is_sim_name(rd->names + rd->name_off[0], rd->name_off[1] - rd->name_off[0])) out->real_reads = 0;

    /* ---- a1: records in file order */
    int64_t n_lines = 1;
    for (int64_t i = 0; i < paf_len; i++) n_lines += (paf[i] == '\n');
    ALLOC(out->qid, int32_t, n_lines); ALLOC(out->tid, int32_t, n_lines);
    ALLOC(out->qs, int32_t, n_lines);  ALLOC(out->qe, int32_t, n_lines);
    ALLOC(out->ts, int32_t, n_lines);  ALLOC(out->te, int32_t, n_lines);
    ALLOC(out->strand, uint8_t, n_lines);
    cols_t  cols = {out->qid, out->tid, out->qs, out->qe, out->ts, out->te, out->strand};
    int64_t N = 0;
    int64_t bad = parse_paf_range(paf, 0, paf_len, paf_len, &map, &cols, &N);
    free(map.slot);
    if (bad >= 0) return fail(out, ORC_E_UNKNOWN_NAME, bad);
    out->n_rec = N;

    /* ---- a2: symmetric flag (chop.hpp:171-184): some record k>=1 mirrors record 0 */
    int S = 0;
    for (int64_t k = 1; k < N && !S; k++)
        S = out->qid[0] == out->tid[k] && out->tid[0] == out->qid[k] && out->qs[0] == out->ts[k] &&
            out->qe[0] == out->te[k] && out->ts[0] == out->qs[k] && out->te[0] == out->qe[k];
    out->symmetric = S;

    /* ---- a3: coverage (repeat.hpp:28-79) as a difference array per read (nb bins + one sentinel slot) */
    ALLOC(out->bin_off, int64_t, n + 1);
    for (int64_t i = 0; i < n; i++) {
        int64_t L = rd->seq_off[i + 1] - rd->seq_off[i];
        out->bin_off[i + 1] = out->bin_off[i] + (L + r - 1) / r; /* repeat.hpp:32-37 */
    }
    const int64_t B = out->bin_off[n];
    ALLOC(out->cov, int32_t, B + 1);
    int32_t *diff;
    ALLOC(diff, int32_t, B + n + 1);
    for (int64_t k = 0; k < N; k++) {
        int32_t q = out->qid[k], t = out->tid[k];
        /* repeat.hpp:50-53, chop.hpp:165 */
        if (!add_interval(diff + out->bin_off[q] + q, out->bin_off[q + 1] - out->bin_off[q], out->qs[k], out->qe[k], r, 0)) { free(diff); return fail(out, ORC_E_RANGE, k); }
        /* repeat.hpp:54-57, chop.hpp:166-169 */
        if (!S && t != q && !add_interval(diff + out->bin_off[t] + t, out->bin_off[t + 1] - out->bin_off[t], out->ts[k], out->te[k], r, 0)) { free(diff); return fail(out, ORC_E_RANGE, k); }
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
    reps_t   reps = {0};
    rstats_t rs = {0};
    for (int64_t i = 0; i < n; i++) {
        int64_t L = rd->seq_off[i + 1] - rd->seq_off[i];
        out->total_read_len += L;
        out->rep_off[i] = reps.n;
        if (!read_repeats(out->cov + out->bin_off[i], out->bin_off[i + 1] - out->bin_off[i], L, r, p, f, H, &reps, &rs)) { free(reps.s); free(reps.e); return fail(out, ORC_E_NOMEM, i); }
    }
    out->rep_off[n] = reps.n;
    out->rep_s = reps.s; out->rep_e = reps.e;
    if (!out->rep_s) { ALLOC(out->rep_s, int32_t, 1); }
    if (!out->rep_e) { ALLOC(out->rep_e, int32_t, 1); }
    out->total_cov = rs.total_cov; out->total_windows = (int32_t)rs.total_windows; out->total_repeat_len = rs.total_repeat_len;

    /* ---- a5: stars and fragments (chop.hpp:198-323) */
    const int div = l / P; /* chop.hpp:248 */
    int64_t frag_cap = n + 1024, G = 0;
    ALLOC(out->frag_read, int32_t, frag_cap); ALLOC(out->frag_a, int32_t, frag_cap); ALLOC(out->frag_b, int32_t, frag_cap);
    frags_t fr = {0};
    for (int64_t i = 0; i < n; i++) {
        int64_t L = rd->seq_off[i + 1] - rd->seq_off[i];
        st = read_frags(L, P, div, v, out->rep_s + out->rep_off[i], out->rep_e + out->rep_off[i], out->rep_off[i + 1] - out->rep_off[i], &fr);
        if (st) { free(fr.fin); free(fr.a); free(fr.b); return fail(out, st, i); }
        if (G + fr.n > frag_cap) {
            frag_cap = (G + fr.n) * 2;
            out->frag_read = (int32_t *)realloc(out->frag_read, sizeof(int32_t) * (size_t)frag_cap);
            out->frag_a = (int32_t *)realloc(out->frag_a, sizeof(int32_t) * (size_t)frag_cap);
            out->frag_b = (int32_t *)realloc(out->frag_b, sizeof(int32_t) * (size_t)frag_cap);
            if (!out->frag_read || !out->frag_a || !out->frag_b) return fail(out, ORC_E_NOMEM, i);
        }
        for (int64_t j = 0; j < fr.n; j++) { out->frag_read[G] = (int32_t)i; out->frag_a[G] = fr.a[j]; out->frag_b[G] = fr.b[j]; G++; }
    }
    free(fr.fin); free(fr.a); free(fr.b);
    out->n_frag = G;

    if (flags & ORC_NO_TEXT) return ORC_OK;

    /* ---- text outputs */
    buf_t  cov = {0}, rep = {0}, bed = {0}, fa = {0};
    sink_t s_cov = {SINK_BUF, &cov, 0, 0}, s_rep = {SINK_BUF, &rep, 0, 0}, s_bed = {SINK_BUF, &bed, 0, 0}, s_fa = {SINK_BUF, &fa, 0, 0};
    for (int64_t i = 0; i < n; i++) emit_cov_line(&s_cov, i, out->cov + out->bin_off[i], out->bin_off[i + 1] - out->bin_off[i], r);
    for (int64_t i = 0; i < n; i++)
        emit_rep_line(&s_rep, &s_bed, i, out->rep_s + out->rep_off[i], out->rep_e + out->rep_off[i], out->rep_off[i + 1] - out->rep_off[i],
                      out->real_reads, rd->names + rd->name_off[i], rd->name_off[i + 1] - rd->name_off[i]);
    for (int64_t g = 0; g < G; g++) {
        int64_t i = out->frag_read[g];
        int     whole = (g == 0 || out->frag_read[g - 1] != i) && (g + 1 == G || out->frag_read[g + 1] != i);
        emit_fasta_record(&s_fa, g, whole, out->frag_a[g], out->frag_b[g], rd->seq_off[i + 1] - rd->seq_off[i], rd->names + rd->name_off[i],
                          rd->name_off[i + 1] - rd->name_off[i], rd->seq + rd->seq_off[i], out->real_reads);
    }
    if (cov.oom || rep.oom || bed.oom || fa.oom) { free(cov.p); free(rep.p); free(bed.p); free(fa.p); return fail(out, ORC_E_NOMEM, -1); }
    out->cov_txt = cov.p; out->cov_txt_len = cov.len;
    out->rep_txt = rep.p; out->rep_txt_len = rep.len;
    out->bed_txt = bed.p; out->bed_txt_len = bed.len;
    out->fasta = fa.p;    out->fasta_len = fa.len;
    return ORC_OK;
}

/* ------------------------------------------------------------------ whole path, digests only (multi-threaded)
 * Same stages and the same per-read helpers as orc_run, for inputs whose outputs do not fit host memory: the PAF is
 * parsed by byte range, intervals are added atomically, and reads are handled in contiguous chunks whose fragment
 * numbering and file offsets come from prefix sums over the chunks (three passes: counts, sizes, digests).
 * Threads: OpenMP when compiled with -fopenmp, else one. */
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t  n_frag, n_rep;
    int64_t  bytes[4];  /* coverage, long_repeats, bed, reads.fasta */
    rstats_t rs;
    int      status; int64_t bad;
} chunk_t;

/* passes over the reads [r0, r1) of one chunk: pass 1 counts fragments / repeats / text bytes that do not depend on the
 * numbering, pass 2 sizes reads.fasta (first_frag known), pass 3 digests all four streams (offsets known). */
static void chunk_pass(int pass, const orc_reads_t *rd, const orc_params_t *prm, int32_t H, int real_reads, const int64_t *bin_off,
                       const int32_t *slots, int64_t r0, int64_t r1, chunk_t *ck, int64_t first_frag, const int64_t base[4], uint64_t dig[4])
{
    const int r = prm->reso, P = prm->interval_length, p = prm->repeat_length, div = prm->read_length / prm->interval_length;
    const int v = prm->overlap_length, f = prm->flanking_length;
    reps_t   reps = {0};
    frags_t  fr = {0};
    rstats_t rs = {0};
    const int mode = pass == 3 ? SINK_DIGEST : SINK_COUNT;
    sink_t   s_cov = {mode, NULL, pass == 3 ? base[0] : 0, 0}, s_rep = {mode, NULL, pass == 3 ? base[1] : 0, 0};
    sink_t   s_bed = {mode, NULL, pass == 3 ? base[2] : 0, 0}, s_fa = {mode, NULL, pass == 3 ? base[3] : 0, 0};
    int64_t  g = first_frag, n_rep = 0;
    for (int64_t i = r0; i < r1; i++) {
        const int64_t  L = rd->seq_off[i + 1] - rd->seq_off[i], nb = bin_off[i + 1] - bin_off[i];
        const int32_t *c = slots + bin_off[i] + i; /* coverage of read i, in place in its slots */
        const uint8_t *nm = rd->names + rd->name_off[i];
        const int64_t  nl = rd->name_off[i + 1] - rd->name_off[i];
        reps.n = 0;
        if (!read_repeats(c, nb, L, r, p, f, H, &reps, &rs)) { ck->status = ORC_E_NOMEM; ck->bad = i; break; }
        int st = read_frags(L, P, div, v, reps.s, reps.e, reps.n, &fr);
        if (st) { ck->status = st; ck->bad = i; break; }
        n_rep += reps.n;
        if (pass != 2) { emit_cov_line(&s_cov, i, c, nb, r); emit_rep_line(&s_rep, &s_bed, i, reps.s, reps.e, reps.n, real_reads, nm, nl); }
        if (pass != 1) {
            /* a counting sink never touches the bytes, so pass 2 costs no sequence traffic */
            for (int64_t j = 0; j < fr.n; j++)
                emit_fasta_record(&s_fa, g + j, fr.n == 1, fr.a[j], fr.b[j], L, nm, nl, rd->seq + rd->seq_off[i], real_reads);
        }
        g += fr.n;
    }
    free(reps.s); free(reps.e); free(fr.fin); free(fr.a); free(fr.b);
    if (pass == 1) { ck->n_frag = g - first_frag; ck->n_rep = n_rep; ck->rs = rs; ck->bytes[0] = s_cov.pos; ck->bytes[1] = s_rep.pos; ck->bytes[2] = s_bed.pos; }
    if (pass == 2) ck->bytes[3] = s_fa.pos;
    if (pass == 3) { dig[0] = s_cov.dig; dig[1] = s_rep.dig; dig[2] = s_bed.dig; dig[3] = s_fa.dig; }
}

int orc_run_digest(const orc_reads_t *rd, const uint8_t *paf, int64_t paf_len, const orc_params_t *prm, int nthreads,
                   orc_digest_result_t *out)
{
    memset(out, 0, sizeof *out);
    out->bad_index = -1;
#define DFAIL(st_, idx_) do { out->status = (st_); out->bad_index = (idx_); goto done; } while (0)
    const int64_t n = rd->n_reads;
    const int     r = prm->reso, P = prm->interval_length, p = prm->repeat_length, l = prm->read_length;
    int32_t      *slots = NULL;
    int64_t      *bin_off = NULL, *range = NULL, *rN = NULL, *rbad = NULL;
    cols_t       *rc = NULL;
    cols_t        c = {0};
    chunk_t      *cks = NULL;
    int           T = nthreads > 0 ? nthreads : 1, nr = 0;
    int           have_map = 0;
    nmap_t        map;
#ifdef _OPENMP
    omp_set_num_threads(T);
#else
    T = 1;
#endif
    if (r < 1 || p < 1 || P < 1 || l < P) DFAIL(ORC_E_PARAM, -1);
    {
        int64_t dup = -1;
        int     st = build_map(&map, rd, &dup);
        if (st) DFAIL(st, dup);
        have_map = 1;
    }
    out->real_reads = 1;
    if (n > 0 && is_sim_name(rd->names + rd->name_off[0], rd->name_off[1] - rd->name_off[0])) out->real_reads = 0;

    /* ---- a1: byte ranges cut after newlines, parsed independently, concatenated in file order */
    nr = T * 4;
    range = (int64_t *)calloc((size_t)nr + 1, sizeof(int64_t));
    rN = (int64_t *)calloc((size_t)nr + 1, sizeof(int64_t));
    rbad = (int64_t *)calloc((size_t)nr + 1, sizeof(int64_t));
    rc = (cols_t *)calloc((size_t)nr + 1, sizeof(cols_t));
    if (!range || !rN || !rbad || !rc) DFAIL(ORC_E_NOMEM, -1);
    for (int k = 1; k < nr; k++) {
        int64_t pos = paf_len / nr * k;
        if (pos < range[k - 1]) pos = range[k - 1];
        while (pos < paf_len && pos > 0 && paf[pos - 1] != '\n') pos++;
        range[k] = pos;
    }
    range[nr] = paf_len;
    int oom = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < nr; k++) {
        int64_t lines = 1;
        for (int64_t i = range[k]; i < range[k + 1]; i++) lines += (paf[i] == '\n');
        cols_t *q = &rc[k];
        q->qid = (int32_t *)malloc(sizeof(int32_t) * (size_t)lines); q->tid = (int32_t *)malloc(sizeof(int32_t) * (size_t)lines);
        q->qs = (int32_t *)malloc(sizeof(int32_t) * (size_t)lines);  q->qe = (int32_t *)malloc(sizeof(int32_t) * (size_t)lines);
        q->ts = (int32_t *)malloc(sizeof(int32_t) * (size_t)lines);  q->te = (int32_t *)malloc(sizeof(int32_t) * (size_t)lines);
        q->strand = (uint8_t *)malloc((size_t)lines);
        if (!q->qid || !q->tid || !q->qs || !q->qe || !q->ts || !q->te || !q->strand) {
#pragma omp atomic write
            oom = 1;
            rbad[k] = -1;
            continue;
        }
        rbad[k] = parse_paf_range(paf, range[k], range[k + 1], paf_len, &map, q, &rN[k]);
    }
    if (oom) DFAIL(ORC_E_NOMEM, -1);
    int64_t N = 0;
    for (int k = 0; k < nr; k++) {
        if (rbad[k] >= 0) DFAIL(ORC_E_UNKNOWN_NAME, N + rbad[k]);
        N += rN[k];
    }
    out->n_rec = N;
    c.qid = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N + 1)); c.tid = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N + 1));
    c.qs = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N + 1));  c.qe = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N + 1));
    c.ts = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N + 1));  c.te = (int32_t *)malloc(sizeof(int32_t) * (size_t)(N + 1));
    if (!c.qid || !c.tid || !c.qs || !c.qe || !c.ts || !c.te) DFAIL(ORC_E_NOMEM, -1);
    {
        int64_t *pre = rbad; /* reuse: exclusive prefix of the record counts */
        int64_t  acc = 0;
        for (int k = 0; k < nr; k++) { pre[k] = acc; acc += rN[k]; }
#pragma omp parallel for schedule(dynamic, 1)
        for (int k = 0; k < nr; k++) {
            size_t b4 = sizeof(int32_t) * (size_t)rN[k];
            memcpy(c.qid + pre[k], rc[k].qid, b4); memcpy(c.tid + pre[k], rc[k].tid, b4);
            memcpy(c.qs + pre[k], rc[k].qs, b4);   memcpy(c.qe + pre[k], rc[k].qe, b4);
            memcpy(c.ts + pre[k], rc[k].ts, b4);   memcpy(c.te + pre[k], rc[k].te, b4);
            free(rc[k].qid); free(rc[k].tid); free(rc[k].qs); free(rc[k].qe); free(rc[k].ts); free(rc[k].te); free(rc[k].strand);
            memset(&rc[k], 0, sizeof rc[k]);
        }
    }
    free(map.slot); have_map = 0;

    /* ---- a2 */
    int S = 0;
#pragma omp parallel for reduction(| : S)
    for (int64_t k = 1; k < N; k++)
        S |= c.qid[0] == c.tid[k] && c.tid[0] == c.qid[k] && c.qs[0] == c.ts[k] && c.qe[0] == c.te[k] && c.ts[0] == c.qs[k] && c.te[0] == c.qe[k];
    out->symmetric = S;

    /* ---- a3: difference slots (nb + 1 per read), then the prefix sum in place */
    bin_off = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
    if (!bin_off) DFAIL(ORC_E_NOMEM, -1);
    for (int64_t i = 0; i < n; i++) {
        int64_t L = rd->seq_off[i + 1] - rd->seq_off[i];
        bin_off[i + 1] = bin_off[i] + (L + r - 1) / r;
        out->total_read_len += L;
    }
    slots = (int32_t *)calloc((size_t)(bin_off[n] + n + 1), sizeof(int32_t));
    if (!slots) DFAIL(ORC_E_NOMEM, -1);
    int64_t bad_range = INT64_MAX;
#pragma omp parallel for reduction(min : bad_range)
    for (int64_t k = 0; k < N; k++) {
        int32_t q = c.qid[k], t = c.tid[k];
        if (!add_interval(slots + bin_off[q] + q, bin_off[q + 1] - bin_off[q], c.qs[k], c.qe[k], r, 1) && k < bad_range) bad_range = k;
        if (!S && t != q && !add_interval(slots + bin_off[t] + t, bin_off[t + 1] - bin_off[t], c.ts[k], c.te[k], r, 1) && k < bad_range) bad_range = k;
    }
    if (bad_range != INT64_MAX) DFAIL(ORC_E_RANGE, bad_range);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; i++) {
        int32_t *d = slots + bin_off[i] + i;
        int64_t  nb = bin_off[i + 1] - bin_off[i];
        int32_t  run = 0;
        for (int64_t k = 0; k < nb; k++) { run += d[k]; d[k] = run; }
    }
    const int32_t H = (int32_t)(prm->est_cov * prm->cov_mul);
    out->high_cov = H;

    /* ---- a4, a5 and the text, chunk by chunk */
    int64_t nck = (int64_t)T * 16;
    if (nck > n) nck = n > 0 ? n : 1;
    cks = (chunk_t *)calloc((size_t)nck, sizeof(chunk_t));
    if (!cks) DFAIL(ORC_E_NOMEM, -1);
#define CK_LO(k) (n * (k) / nck)
    for (int pass = 1; pass <= 3; pass++) {
        int64_t first = 0, base[4] = {0, 0, 0, 0};
        int64_t *firsts = (int64_t *)calloc((size_t)nck * 5, sizeof(int64_t));
        if (!firsts) DFAIL(ORC_E_NOMEM, -1);
        for (int64_t k = 0; k < nck; k++) {
            firsts[5 * k] = first; first += cks[k].n_frag;
            for (int w = 0; w < 4; w++) { firsts[5 * k + 1 + w] = base[w]; base[w] += cks[k].bytes[w]; }
        }
        uint64_t d0 = 0, d1 = 0, d2 = 0, d3 = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : d0, d1, d2, d3)
        for (int64_t k = 0; k < nck; k++) {
            uint64_t dg[4] = {0, 0, 0, 0};
            if (cks[k].status == 0)
                chunk_pass(pass, rd, prm, H, out->real_reads, bin_off, slots, CK_LO(k), CK_LO(k + 1), &cks[k], firsts[5 * k], &firsts[5 * k + 1], dg);
            d0 += dg[0]; d1 += dg[1]; d2 += dg[2]; d3 += dg[3];
        }
        free(firsts);
        for (int64_t k = 0; k < nck; k++) if (cks[k].status) DFAIL(cks[k].status, cks[k].bad);
        if (pass == 3) { out->digest[0] = d0; out->digest[1] = d1; out->digest[2] = d2; out->digest[3] = d3; }
    }
    {
        uint32_t tw = 0;
        for (int64_t k = 0; k < nck; k++) {
            out->n_frag += cks[k].n_frag; out->n_rep += cks[k].n_rep;
            out->total_cov += cks[k].rs.total_cov; tw += cks[k].rs.total_windows; out->total_repeat_len += cks[k].rs.total_repeat_len;
            for (int w = 0; w < 4; w++) out->bytes[w] += cks[k].bytes[w];
        }
        out->total_windows = (int32_t)tw;
    }
done:
    if (have_map) free(map.slot);
    if (rc) for (int k = 0; k < nr; k++) { free(rc[k].qid); free(rc[k].tid); free(rc[k].qs); free(rc[k].qe); free(rc[k].ts); free(rc[k].te); free(rc[k].strand); }
    free(rc); free(range); free(rN); free(rbad);
    free(c.qid); free(c.tid); free(c.qs); free(c.qe); free(c.ts); free(c.te);
    free(bin_off); free(slots); free(cks);
    return out->status;
#undef DFAIL
#undef CK_LO
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
    buf_t  b = {0};
    sink_t sk = {SINK_BUF, &b, 0, 0};
    if (sublen < 1) return ORC_E_PARAM;
    for (int64_t i = 0; i < rd->n_reads; i++) {
        int64_t L = rd->seq_off[i + 1] - rd->seq_off[i];
        int64_t k = 1;
        for (int64_t a = 0; a < L; a += sublen, k++) { /* split_naive.cpp:27-29 */
            int64_t n = a + sublen < L ? sublen : L - a;
            sink_str(&sk, ">"); sink_put(&sk, rd->names + rd->name_off[i], rd->name_off[i + 1] - rd->name_off[i]);
            sink_str(&sk, "_"); sink_int(&sk, k); sink_str(&sk, "\n");          /* split_naive.cpp:32 */
            sink_put(&sk, rd->seq + rd->seq_off[i] + a, n); sink_str(&sk, "\n");
        }
    }
    if (b.oom) { free(b.p); return ORC_E_NOMEM; }
    if (!b.p) b.p = (uint8_t *)calloc(1, 1);
    *out = b.p;
    return b.len;
}
