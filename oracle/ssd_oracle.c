/*
 * ssd_oracle.c — TEST INFRASTRUCTURE ONLY (see ssd_oracle.h).  Parity status: PINNED against golden
 * vectors generated from the unmodified reference (tests/golden/make_golden.py).
 *
 * Every function names the reference lines it restates.  V2 = python_only_tool/
 * DemultiplexUsingBarcodes_New_V2.py, LIB = python_only_tool/Splitseq_fun_lib.py, CLI3 =
 * python_only_tool/splitseqdemultiplex_0.2.3.py.  The restatement is deliberately plain: 3 x 96
 * byte-wise Hamming comparisons per read exactly like the reference's hot loop, no lookup tables.
 */
#include "ssd_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------- small helpers --------------- */

typedef struct sv {
    const char *p;
    size_t n;
} sv;

/* str.isspace() for ASCII: what str.strip()/rstrip()/split() remove. */
static int py_space(unsigned char c) { return (c >= 0x09 && c <= 0x0D) || (c >= 0x1C && c <= 0x1F) || c == 0x20; }

static sv sv_rstrip(sv s)
{
    while (s.n && py_space((unsigned char)s.p[s.n - 1])) s.n--;
    return s;
}

static sv sv_strip(sv s)
{
    s = sv_rstrip(s);
    while (s.n && py_space((unsigned char)s.p[0])) {
        s.p++;
        s.n--;
    }
    return s;
}

/* name.split()[0] for a non-empty, already rstripped string */
static sv sv_first_token(sv s)
{
    size_t a = 0, b;
    while (a < s.n && py_space((unsigned char)s.p[a])) a++;
    b = a;
    while (b < s.n && !py_space((unsigned char)s.p[b])) b++;
    sv t = {s.p + a, b - a};
    return t;
}

/* Python slice s[a:b] with 0 <= a, 0 <= b */
static sv sv_slice(sv s, int a, int b)
{
    sv r = {s.p, 0};
    size_t ua = (size_t)a, ub = (size_t)b;
    if (ub > s.n) ub = s.n;
    if (ua >= ub) {
        r.p = s.p + (ua < s.n ? ua : s.n);
        return r;
    }
    r.p = s.p + ua;
    r.n = ub - ua;
    return r;
}

typedef struct buf {
    char *p;
    size_t n, cap;
} buf;

static int buf_put(buf *b, const char *src, size_t n)
{
    if (b->n + n > b->cap) {
        size_t cap = b->cap ? b->cap * 2 : 1 << 16;
        while (cap < b->n + n) cap *= 2;
        char *q = (char *)realloc(b->p, cap);
        if (!q) return -1;
        b->p = q;
        b->cap = cap;
    }
    memcpy(b->p + b->n, src, n);
    b->n += n;
    return 0;
}

/* Line cursor: Python `for line in infile` over text with "\n" terminators.  A final unterminated
 * non-empty tail is a line.  The returned view excludes the "\n". */
typedef struct cursor {
    const char *p;
    size_t n, pos;
} cursor;

static int next_line(cursor *c, sv *line)
{
    if (c->pos >= c->n) return 0;
    const char *s = c->p + c->pos;
    const char *e = (const char *)memchr(s, '\n', c->n - c->pos);
    line->p = s;
    if (e) {
        line->n = (size_t)(e - s);
        c->pos += line->n + 1;
    } else {
        line->n = c->n - c->pos;
        c->pos = c->n;
    }
    return 1;
}

static uint64_t count_lines(const char *p, size_t n)
{
    cursor c = {p, n, 0};
    sv l;
    uint64_t k = 0;
    while (next_line(&c, &l)) k++;
    return k;
}

/* V2:72-73  hamming(s1, s2) = sum(c1 != c2 for c1, c2 in zip(s1, s2)) — zip truncates. */
static int hamming(const char *a, size_t la, const char *b, size_t lb)
{
    size_t n = la < lb ? la : lb, i;
    int d = 0;
    for (i = 0; i < n; i++) d += a[i] != b[i];
    return d;
}

/* V2:302-304 + 327-329: [s for s in Eight_BP_barcode if hamming(s, window) <= e] then [0]:
 * the FIRST whitelist entry in file order within distance e, or -1 when the list is empty. */
/* the list of round k (0: bc1, 1: bc2, 2: bc3): the one list of V2:60 unless the extension fields name another */
static void round_list(const ssd_oracle_cfg *cfg, int k, const char **wl, const uint8_t **len, int *n)
{
    *wl = cfg->whitelist; *len = cfg->wl_len; *n = cfg->n_whitelist;
    if (k == 1 && cfg->whitelist2) { *wl = cfg->whitelist2; *len = cfg->wl_len2; *n = cfg->n_whitelist2; }
    if (k == 2 && cfg->whitelist3) { *wl = cfg->whitelist3; *len = cfg->wl_len3; *n = cfg->n_whitelist3; }
}

static int first_hit(const ssd_oracle_cfg *cfg, int round, sv window)
{
    const char *list; const uint8_t *len; int n, i;
    round_list(cfg, round, &list, &len, &n);
    for (i = 0; i < n; i++) {
        size_t wl = len ? len[i] : 8;
        if (hamming(list + 8 * (size_t)i, wl, window.p, window.n) <= cfg->errors) return i;
    }
    return -1;
}

/* ---------------------------------------------------------------- id -> record map ------------ */

typedef struct frec {
    sv name, id, read, qual;
} frec;

typedef struct rrec {
    sv id, umi;
    int b1, b2, b3;
} rrec;

static uint64_t hash_sv(sv s)
{
    uint64_t h = 1469598103934665603ull;
    size_t i;
    for (i = 0; i < s.n; i++) h = (h ^ (unsigned char)s.p[i]) * 1099511628211ull;
    return h;
}

static int sv_eq(sv a, sv b) { return a.n == b.n && memcmp(a.p, b.p, a.n) == 0; }

typedef struct idmap {
    int64_t *slot; /* index into a record array, -1 empty */
    size_t cap;
} idmap;

static int idmap_init(idmap *m, size_t n)
{
    size_t cap = 16;
    while (cap < n * 2 + 1) cap *= 2;
    m->slot = (int64_t *)malloc(cap * sizeof(int64_t));
    if (!m->slot) return -1;
    m->cap = cap;
    memset(m->slot, 0xFF, cap * sizeof(int64_t));
    return 0;
}

/* ---------------------------------------------------------------- demultiplex_fun ------------- */

typedef struct parse_state { /* the function-level locals of V2 that survive across bins */
    sv name, id, read, qual;
    int have_id, have_name, have_read, have_qual;
} parse_state;

static int fail(ssd_oracle_out *out, int code, const char *what, sv arg)
{
    out->error = code;
    snprintf(out->error_msg, sizeof out->error_msg, "%s%.*s", what, (int)(arg.n > 200 ? 200 : arg.n), arg.p ? arg.p : "");
    return code;
}

int ssd_oracle_demultiplex(const ssd_oracle_cfg *cfg, const char *r1, size_t r1_len, const char *r2,
                           size_t r2_len, ssd_oracle_out *out)
{
    const uint64_t bin_lines = (uint64_t)cfg->bin_reads * 4; /* V2:80 binIterator */
    const uint64_t lines_f = count_lines(r1, r1_len);         /* V2:88-90 */
    cursor cf = {r1, r1_len, 0}, cr = {r2, r2_len, 0};
    buf merged = {out->merged1, out->merged1_len, out->merged1_len};
    frec *fr = NULL;
    rrec *rr = NULL;
    size_t fcap = 0, rcap = 0;
    int32_t *rec_cell = out->rec_cell;
    uint64_t rec_n = out->n_records, rec_cap = out->n_records;
    parse_state pf, pr;
    int rc = SSD_ORACLE_OK;
    const int n2 = cfg->whitelist2 ? cfg->n_whitelist2 : cfg->n_whitelist, n3 = cfg->whitelist3 ? cfg->n_whitelist3 : cfg->n_whitelist;
    uint64_t i;
    sv none = {NULL, 0};

    memset(&pf, 0, sizeof pf);
    memset(&pr, 0, sizeof pr);
    if (bin_lines == 0) return fail(out, SSD_ORACLE_EMALFORMED, "numReadsBin must be > 0", none);

    for (i = 0; i < lines_f; i += bin_lines) { /* V2:207 */
        size_t nf = 0, nr = 0, k;
        uint64_t taken;
        sv line;
        int starter, complete;
        uint64_t line_ct;
        int f1 = -1, f2 = -1, f3 = -1;
        sv umi = {NULL, 0};
        idmap fmap = {NULL, 0}, rmap = {NULL, 0};

        /* ---- forward reads of this bin: V2:223-262 ---- */
        starter = 0;
        complete = 0;
        line_ct = 0;
        for (taken = 0; taken < bin_lines && next_line(&cf, &line); taken++) {
            int at = line.n > 0 && line.p[0] == '@';
            if (!starter && !at) continue;                      /* V2:230-231 */
            if (!starter) {                                     /* V2:232-239 */
                starter = 1;
                pf.name = sv_rstrip(line);
                pf.id = sv_first_token(pf.name);
                pf.have_id = pf.have_name = 1;
                complete++;
                line_ct++;
                continue;
            }
            if (line_ct % 4 == 0 && at) {                       /* V2:240-244 */
                pf.name = sv_rstrip(line);
                pf.id = sv_first_token(pf.name);
                pf.have_id = pf.have_name = 1;
                complete++;
            }
            if (line_ct % 4 == 1) {                             /* V2:245-247 */
                pf.read = sv_rstrip(line);
                pf.have_read = 1;
                complete++;
            }
            if (line_ct % 4 == 3) {                             /* V2:248-250 */
                pf.qual = sv_rstrip(line);
                pf.have_qual = 1;
                complete++;
            }
            if (complete == 3) {                                /* V2:251-260 */
                if (!pf.have_id || !pf.have_read || !pf.have_qual) {
                    rc = fail(out, SSD_ORACLE_EMALFORMED, "read 1: record completed without id/seq/qual", none);
                    goto bin_done;
                }
                if (nf == fcap) {
                    fcap = fcap ? fcap * 2 : 1024;
                    fr = (frec *)realloc(fr, fcap * sizeof *fr);
                    if (!fr) { rc = SSD_ORACLE_ENOMEM; goto bin_done; }
                }
                fr[nf].name = pf.name;
                fr[nf].id = pf.id;
                fr[nf].read = pf.read;
                fr[nf].qual = pf.qual;
                nf++;
                pf.have_id = 0;                                 /* del(readIDF) */
                complete = 0;
            }
            line_ct++;                                          /* V2:261-262 (starterF == 1 here) */
        }

        /* ---- reverse reads of this bin: V2:270-337 ---- */
        starter = 0;
        complete = 0;
        line_ct = 0;
        for (taken = 0; taken < bin_lines && next_line(&cr, &line); taken++) {
            int at = line.n > 0 && line.p[0] == '@';
            if (!starter && !at) continue;                      /* V2:277-278 */
            if (!starter) {                                     /* V2:279-286 */
                starter = 1;
                pr.name = sv_rstrip(line);
                pr.id = sv_first_token(pr.name);
                pr.have_id = pr.have_name = 1;
                complete++;
                line_ct++;
                continue;
            }
            if (line_ct % 4 == 0 && at) {                       /* V2:287-291 */
                pr.name = sv_rstrip(line);
                pr.id = sv_first_token(pr.name);
                pr.have_id = pr.have_name = 1;
                complete++;
            }
            if (line_ct % 4 == 1) {                             /* V2:292-305 */
                pr.read = sv_rstrip(line);
                pr.have_read = 1;
                umi = sv_slice(pr.read, cfg->umi_start, cfg->umi_end);
                f1 = first_hit(cfg, 0, sv_slice(pr.read, cfg->bc1_start, cfg->bc1_end));
                f2 = first_hit(cfg, 1, sv_slice(pr.read, cfg->bc2_start, cfg->bc2_end));
                f3 = first_hit(cfg, 2, sv_slice(pr.read, cfg->bc3_start, cfg->bc3_end));
                complete++;
            }
            if (line_ct % 4 == 3) {                             /* V2:306-308 */
                pr.qual = sv_rstrip(line);
                pr.have_qual = 1;
                complete++;
            }
            if (complete == 3) {                                /* V2:309-335 */
                if (rec_n == rec_cap) {
                    rec_cap = rec_cap ? rec_cap * 2 : 1024;
                    rec_cell = (int32_t *)realloc(rec_cell, rec_cap * sizeof *rec_cell);
                    if (!rec_cell) { rc = SSD_ORACLE_ENOMEM; goto bin_done; }
                }
                if (!pr.have_read) {
                    rc = fail(out, SSD_ORACLE_EMALFORMED, "read 2: record completed without a sequence line", none);
                    goto bin_done;
                }
                if (f1 < 0 || f2 < 0 || f3 < 0) {               /* V2:310-321: rejected, id NOT deleted */
                    rec_cell[rec_n++] = -1;
                    line_ct++;
                    complete = 0;
                    continue;
                }
                if (!pr.have_id) {
                    rc = fail(out, SSD_ORACLE_EMALFORMED, "read 2: record completed without an id", none);
                    goto bin_done;
                }
                rec_cell[rec_n++] = (int32_t)((f1 * n2 + f2) * n3 + f3);
                if (nr == rcap) {
                    rcap = rcap ? rcap * 2 : 1024;
                    rr = (rrec *)realloc(rr, rcap * sizeof *rr);
                    if (!rr) { rc = SSD_ORACLE_ENOMEM; goto bin_done; }
                }
                rr[nr].id = pr.id;
                rr[nr].umi = umi;
                rr[nr].b1 = f1;
                rr[nr].b2 = f2;
                rr[nr].b3 = f3;
                nr++;
                pr.have_id = 0;                                 /* del(readIDR) */
                complete = 0;
            }
            line_ct++;
        }

        /* ---- readsF / readsR are dicts keyed by read id: later duplicates overwrite (V2:257, 332) */
        if (idmap_init(&fmap, nf) || idmap_init(&rmap, nr)) { rc = SSD_ORACLE_ENOMEM; goto bin_done; }
        for (k = 0; k < nf; k++) {
            size_t h = (size_t)hash_sv(fr[k].id) & (fmap.cap - 1);
            while (fmap.slot[h] >= 0 && !sv_eq(fr[fmap.slot[h]].id, fr[k].id)) h = (h + 1) & (fmap.cap - 1);
            fmap.slot[h] = (int64_t)k;
        }
        for (k = 0; k < nr; k++) {
            size_t h = (size_t)hash_sv(rr[k].id) & (rmap.cap - 1);
            while (rmap.slot[h] >= 0 && !sv_eq(rr[rmap.slot[h]].id, rr[k].id)) h = (h + 1) & (rmap.cap - 1);
            if (rmap.slot[h] >= 0) { /* same key again: value replaced, dict position kept */
                rrec *first = &rr[rmap.slot[h]];
                first->umi = rr[k].umi;
                first->b1 = rr[k].b1;
                first->b2 = rr[k].b2;
                first->b3 = rr[k].b3;
                rr[k].b1 = -1; /* tombstone */
            } else {
                rmap.slot[h] = (int64_t)k;
            }
        }

        /* ---- join + write: V2:349-358, 134-141, 367-370 (records emitted in read-2 order) ---- */
        for (k = 0; k < nr; k++) {
            const frec *f;
            size_t h, q;
            sv name;
            char hdr[4096];
            size_t hn = 0;
            const char *w;
            size_t wl;
            if (rr[k].b1 < 0) continue;
            h = (size_t)hash_sv(rr[k].id) & (fmap.cap - 1);
            while (fmap.slot[h] >= 0 && !sv_eq(fr[fmap.slot[h]].id, rr[k].id)) h = (h + 1) & (fmap.cap - 1);
            if (fmap.slot[h] < 0) { /* KeyError(key), V2:350 */
                rc = fail(out, SSD_ORACLE_EKEY, "", rr[k].id);
                goto bin_done;
            }
            f = &fr[fmap.slot[h]];
            /* return_fastq, V2:134-141: name.split("/")[0].replace(" ", "").strip() */
            name = f->name;
            for (q = 0; q < name.n && name.p[q] != '/'; q++)
                if (name.p[q] != ' ' && hn < sizeof hdr) hdr[hn++] = name.p[q];
            {
                sv t = {hdr, hn};
                t = sv_strip(t);
                if (buf_put(&merged, t.p, t.n)) { rc = SSD_ORACLE_ENOMEM; goto bin_done; }
            }
            buf_put(&merged, "_", 1);
            {
                const int idx[3] = {rr[k].b1, rr[k].b2, rr[k].b3};
                int rq;
                for (rq = 0; rq < 3; rq++) {
                    const char *list; const uint8_t *len; int nq;
                    round_list(cfg, rq, &list, &len, &nq);
                    w = list + 8 * (size_t)idx[rq]; wl = len ? len[idx[rq]] : 8; buf_put(&merged, w, wl);
                }
            }
            buf_put(&merged, "_", 1);
            buf_put(&merged, rr[k].umi.p, rr[k].umi.n);
            buf_put(&merged, "\n", 1);
            buf_put(&merged, f->read.p, f->read.n);
            buf_put(&merged, "\n+\n", 3);
            buf_put(&merged, f->qual.p, f->qual.n);
            if (buf_put(&merged, "\n", 1)) { rc = SSD_ORACLE_ENOMEM; goto bin_done; }
            out->n_matched++;
        }
    bin_done:
        free(fmap.slot);
        free(rmap.slot);
        if (rc != SSD_ORACLE_OK) break;
    }
    free(fr);
    free(rr);
    out->merged1 = merged.p;
    out->merged1_len = merged.n;
    out->rec_cell = rec_cell;
    out->n_records = rec_n;
    if (rc != SSD_ORACLE_OK && !out->error) out->error = rc;
    return rc;
}

/* ---------------------------------------------------------------- calc_demux_results ---------- */

typedef struct cu {
    sv cell, umi;
} cu;

static int cmp_cu(const void *a, const void *b)
{
    const cu *x = (const cu *)a, *y = (const cu *)b;
    size_t n = x->cell.n < y->cell.n ? x->cell.n : y->cell.n;
    int c = memcmp(x->cell.p, y->cell.p, n);
    if (c) return c;
    if (x->cell.n != y->cell.n) return x->cell.n < y->cell.n ? -1 : 1;
    n = x->umi.n < y->umi.n ? x->umi.n : y->umi.n;
    c = memcmp(x->umi.p, y->umi.p, n);
    if (c) return c;
    if (x->umi.n != y->umi.n) return x->umi.n < y->umi.n ? -1 : 1;
    return 0;
}

/* header.split("_") -> fields [1] and [2]; IndexError when there are fewer than three (V2:396-398) */
static int split_cell_umi(sv hdr, sv *cell, sv *umi)
{
    size_t a = 0, b, c;
    while (a < hdr.n && hdr.p[a] != '_') a++;
    if (a == hdr.n) return -1;
    b = a + 1;
    while (b < hdr.n && hdr.p[b] != '_') b++;
    if (b == hdr.n) return -1;
    c = b + 1;
    while (c < hdr.n && hdr.p[c] != '_') c++;
    cell->p = hdr.p + a + 1;
    cell->n = b - a - 1;
    umi->p = hdr.p + b + 1;
    umi->n = c - b - 1;
    return 0;
}

static int cmp_cell_name(const void *key, const void *elem)
{
    const sv *k = (const sv *)key;
    const ssd_oracle_cell *c = (const ssd_oracle_cell *)elem;
    size_t cn = strlen(c->cell), n = k->n < cn ? k->n : cn;
    int r = memcmp(k->p, c->cell, n);
    if (r) return r;
    return k->n == cn ? 0 : (k->n < cn ? -1 : 1);
}

int ssd_oracle_calc_demux_results(ssd_oracle_out *out, long min_umis)
{
    cursor c = {out->merged1, out->merged1_len, 0};
    sv line, none = {NULL, 0};
    cu *v = NULL;
    size_t nv = 0, cap = 0, i;
    uint64_t line_ct = 0;
    buf passing = {out->passing, out->passing_len, out->passing_len};

    /* pass 1, V2:389-404 */
    while (next_line(&c, &line)) {
        if (line_ct % 4 == 0) {
            sv cell, umi;
            if (split_cell_umi(sv_rstrip(line), &cell, &umi)) {
                free(v);
                return fail(out, SSD_ORACLE_EINDEX, "list index out of range: ", sv_rstrip(line));
            }
            if (cell.n >= sizeof out->cells[0].cell) {
                free(v);
                return fail(out, SSD_ORACLE_EMALFORMED, "cell string too long", none);
            }
            if (nv == cap) {
                cap = cap ? cap * 2 : 4096;
                v = (cu *)realloc(v, cap * sizeof *v);
                if (!v) return SSD_ORACLE_ENOMEM;
            }
            v[nv].cell = cell;
            v[nv].umi = umi;
            nv++;
        }
        line_ct++;
    }
    if (nv) qsort(v, nv, sizeof *v, cmp_cu);

    /* V2:406-412: cells = dict keys; keep those with len(set(umis)) >= int(m) */
    free(out->cells);
    out->cells = NULL;
    out->n_cells = out->n_passing_cells = 0;
    for (i = 0; i < nv;) {
        size_t j = i, distinct = 0;
        ssd_oracle_cell *cell;
        while (j < nv && sv_eq(v[j].cell, v[i].cell)) {
            if (j == i || !sv_eq(v[j].umi, v[j - 1].umi)) distinct++;
            j++;
        }
        out->cells = (ssd_oracle_cell *)realloc(out->cells, (out->n_cells + 1) * sizeof *out->cells);
        if (!out->cells) { free(v); return SSD_ORACLE_ENOMEM; }
        cell = &out->cells[out->n_cells++];
        memset(cell, 0, sizeof *cell);
        memcpy(cell->cell, v[i].cell.p, v[i].cell.n);
        cell->reads = j - i;
        cell->distinct_umis = distinct;
        cell->passing = (long)distinct >= min_umis;
        out->n_passing_cells += (uint64_t)cell->passing;
        i = j;
    }
    free(v);

    /* pass 2, V2:415-454: copy 4-line records of kept cells, each line strip()ped */
    c.pos = 0;
    line_ct = 0;
    {
        int keep = 0;
        while (next_line(&c, &line)) {
            if (line_ct % 4 == 0) {
                sv cell, umi;
                const ssd_oracle_cell *hit;
                split_cell_umi(sv_rstrip(line), &cell, &umi);
                hit = (const ssd_oracle_cell *)bsearch(&cell, out->cells, out->n_cells, sizeof *out->cells, cmp_cell_name);
                keep = hit && hit->passing;
                if (keep) out->n_retained++;
            }
            if (keep) {
                sv s = sv_strip(line);
                buf_put(&passing, s.p, s.n);
                if (buf_put(&passing, "\n", 1)) return SSD_ORACLE_ENOMEM;
            }
            line_ct++;
        }
    }
    out->passing = passing.p;
    out->passing_len = passing.n;
    return SSD_ORACLE_OK;
}

/* ---------------------------------------------------------------- Splitseq_fun_lib ------------ */

/* LIB:11-19 (and 58-66): split_size = lengthFastq / N, bumped until it is a multiple of 4 */
static long tuned_split_size(long length_fastq, int n)
{
    double s = (double)length_fastq / (double)n;
    int tries;
    for (tries = 0; tries < 100; tries++) {
        double r = s - 4.0 * (double)(long)(s / 4.0);
        if (r != 0.0) s = (double)((long)s + 1);
    }
    return (long)s;
}

int ssd_oracle_split_fastq(const char *data, size_t len, int n_chunks, long length_fastq, char **chunk,
                           size_t *chunk_len)
{
    const long split_size = tuned_split_size(length_fastq, n_chunks);
    cursor c = {data, len, 0};
    sv line;
    buf cur = {NULL, 0, 0};
    long counter = 0;
    int produced = 0;
    while (next_line(&c, &line)) { /* LIB:27-50 */
        sv s = sv_strip(line);
        if (counter == split_size && counter != 0) {
            if (produced < n_chunks) {
                chunk[produced] = cur.p;
                chunk_len[produced] = cur.n;
            } else {
                free(cur.p); /* the CLI never reads chunk files beyond numCores */
            }
            produced++;
            cur.p = NULL;
            cur.n = cur.cap = 0;
            counter = 0;
        }
        buf_put(&cur, s.p, s.n);
        buf_put(&cur, "\n", 1);
        counter++;
    }
    if (counter) {
        if (produced < n_chunks) {
            chunk[produced] = cur.p;
            chunk_len[produced] = cur.n;
        } else {
            free(cur.p);
        }
        produced++;
    }
    return produced;
}

/* ---------------------------------------------------------------- CLI3 orchestration ---------- */

int ssd_oracle_run_cli(const ssd_oracle_cfg *cfg, const char *r1, size_t r1_len, const char *r2,
                       size_t r2_len, long length_fastq, int n_workers, long min_umis,
                       ssd_oracle_out *out)
{
    char **cf = (char **)calloc((size_t)n_workers, sizeof(char *));
    char **cr = (char **)calloc((size_t)n_workers, sizeof(char *));
    size_t *lf = (size_t *)calloc((size_t)n_workers, sizeof(size_t));
    size_t *lr = (size_t *)calloc((size_t)n_workers, sizeof(size_t));
    ssd_oracle_out *part = (ssd_oracle_out *)calloc((size_t)n_workers, sizeof *part);
    int nf, nr, w, rc = SSD_ORACLE_OK;
    sv none = {NULL, 0};
    buf merged = {NULL, 0, 0};

    nf = ssd_oracle_split_fastq(r1, r1_len, n_workers, length_fastq, cf, lf); /* CLI3:52 */
    nr = ssd_oracle_split_fastq(r2, r2_len, n_workers, length_fastq, cr, lr); /* CLI3:53 */
    if (nf < n_workers || nr < n_workers) {
        rc = fail(out, SSD_ORACLE_EMALFORMED, "FileNotFoundError: fewer chunk files than workers", none);
        goto done;
    }
    /* CLI3:56-64: one worker per chunk */
#ifdef _OPENMP
#pragma omp parallel for schedule(static, 1) num_threads(n_workers)
#endif
    for (w = 0; w < n_workers; w++) ssd_oracle_demultiplex(cfg, cf[w], lf[w], cr[w], lr[w], &part[w]);

    for (w = 0; w < n_workers; w++) {
        if (part[w].error && rc == SSD_ORACLE_OK) {
            rc = part[w].error;
            out->error = rc;
            memcpy(out->error_msg, part[w].error_msg, sizeof out->error_msg);
        }
        buf_put(&merged, part[w].merged1, part[w].merged1_len);
        out->n_matched += part[w].n_matched;
        out->n_records += part[w].n_records;
    }
    out->merged1 = merged.p;
    out->merged1_len = merged.n;
    if (rc == SSD_ORACLE_OK) rc = ssd_oracle_calc_demux_results(out, min_umis); /* CLI3:70 */
done:
    for (w = 0; w < n_workers; w++) {
        free(cf[w]);
        free(cr[w]);
        ssd_oracle_free(&part[w]);
    }
    free(cf);
    free(cr);
    free(lf);
    free(lr);
    free(part);
    return rc;
}

/* ---------------------------------------------------------------- position learner ------------ */

static long find_sub(sv hay, const char *needle)
{
    size_t m = strlen(needle), i;
    if (hay.n < m) return -1;
    for (i = 0; i + m <= hay.n; i++)
        if (memcmp(hay.p + i, needle, m) == 0) return (long)i;
    return -1;
}

/* V2:164-183: over sequence lines (line index % 4 == 1) keep find() positions >= 17, take the mode */
int ssd_oracle_learn_positions(const char *sample, size_t len, int pos[3])
{
    static const char *anchors[3] = {"CATTCG", "ATCCAC", "GTGGCC"}; /* V2:36-38 */
    enum { MAXPOS = 1 << 16 };
    uint32_t *hist = (uint32_t *)calloc(3 * MAXPOS, sizeof(uint32_t));
    cursor c = {sample, len, 0};
    sv line;
    uint64_t ct = 0;
    int a, rc = SSD_ORACLE_OK;
    if (!hist) return SSD_ORACLE_ENOMEM;
    while (next_line(&c, &line)) {
        if (ct % 4 == 1)
            for (a = 0; a < 3; a++) {
                long p = find_sub(line, anchors[a]);
                if (p >= 17 && p < MAXPOS) hist[a * MAXPOS + p]++;
            }
        ct++;
    }
    for (a = 0; a < 3; a++) {
        uint32_t best = 0;
        int p;
        pos[a] = -1;
        for (p = 0; p < MAXPOS; p++)
            if (hist[a * MAXPOS + p] > best) {
                best = hist[a * MAXPOS + p];
                pos[a] = p;
            }
        if (pos[a] < 0) rc = SSD_ORACLE_EMALFORMED; /* max() of an empty set: ValueError */
    }
    free(hist);
    return rc;
}

/* ---------------------------------------------------------------- collapse rule --------------- */

/* Collapse_RanHex_Odt.py:44-71.  OligoDT[k] is Round-1 whitelist index k, RanHex[k] is index
 * k + n_pairs.  A RanHex cell is folded into its OligoDT partner only when both
 * exist for the same (bc2, bc3). */
void ssd_oracle_collapse_map(const uint8_t *present, int n, int n_pairs, uint32_t *remap)
{
    int i1, i2, i3;
    for (i1 = 0; i1 < n; i1++)
        for (i2 = 0; i2 < n; i2++)
            for (i3 = 0; i3 < n; i3++) {
                uint32_t id = (uint32_t)((i1 * n + i2) * n + i3);
                remap[id] = id;
                if (i1 >= n_pairs && i1 < 2 * n_pairs && present[id]) {
                    uint32_t odt = (uint32_t)(((i1 - n_pairs) * n + i2) * n + i3);
                    if (present[odt]) remap[id] = odt;
                }
            }
}

void ssd_oracle_free(ssd_oracle_out *out)
{
    free(out->merged1);
    free(out->passing);
    free(out->rec_cell);
    free(out->cells);
    memset(out, 0, sizeof *out);
}
