/*
 * bedtools_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never on the product path).
 *
 * Plain-C restatement of the two bedtools sub-commands that mimeo's generated
 * shell script invokes for the annotation half of the hot path:
 *
 *   bedtools genomecov -bg -i BED -g LENS   reference call sites: wrappers.py:1131-1144
 *                                           (self inter), 847-860 (x), 1223-1236 (self intra)
 *   bedtools merge -i BED                   reference call sites: wrappers.py:1150, 866, 1246-1250
 *
 * bedtools itself is a third-party dependency that is NOT vendored in the
 * reference tree (environment.yml:8, no version pin) and is not installed in
 * this image, so this file restates its published algorithm:
 *
 *   genomecov: per chromosome of size N keep starts[N], ends[N] (uint32);
 *     for an interval (s,e):  if s < N: starts[s]++ ;  e' = e-1 ;
 *     if 0 <= e' < N: ends[e']++ else ends[N-1]++ ;
 *     sweep pos = 0..N-1: depth += starts[pos]; if depth != lastDepth:
 *       if lastDepth > 0: emit(chrom,lastStart,pos,lastDepth); lastDepth = depth;
 *       lastStart = pos;  depth -= ends[pos];   flush (lastStart,N) at the end.
 *     Chromosomes are reported in order of first appearance in the (sorted) BED.
 *   merge: per chromosome, sorted by start, fold while next.start <= cur.end
 *     (overlapping AND book-ended intervals merge).
 *
 * "parity unpinned" for this file on its own: the reference holds no golden
 * vectors for bedtools (tests/test_dummy.py only). It is pinned indirectly by the
 * hand-derived known-answer vectors of SURVEY.md 9.3 (tests/test_oracle_annot.py)
 * and by running the reference's literal awk/sed/sort command strings around it
 * (tests/golden/make_golden.py uses this binary as the `bedtools` on the script's
 * command line).
 *
 * Built two ways by oracle/Makefile:
 *   oracle/_build/bedtools          CLI shim (argv contract of the two sub-commands)
 *   oracle/_build/libannot_oracle.so  same loops callable through ctypes
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* library form: one chromosome at a time, arrays in / arrays out      */
/* ------------------------------------------------------------------ */

/* genomecov -bg sweep for ONE chromosome.
 * starts/ends: the n intervals on this chromosome (BED coordinates as parsed).
 * size: chromosome length. out_*: caller-allocated, capacity cap rows.
 * returns number of bedGraph rows, or -1 if cap is too small, -2 on bad size. */
long ora_genomecov_bg(const int64_t *s, const int64_t *e, long n, int64_t size,
                      int64_t *out_start, int64_t *out_end, int64_t *out_depth, long cap)
{
    if (size <= 0) return -2;
    uint32_t *st = (uint32_t *)calloc((size_t)size, sizeof(uint32_t));
    uint32_t *en = (uint32_t *)calloc((size_t)size, sizeof(uint32_t));
    if (!st || !en) { free(st); free(en); return -3; }
    for (long i = 0; i < n; i++) {
        int64_t a = s[i], b = e[i] - 1;
        if (a < size) st[a]++;           /* bedtools AddCoverage: start inside chrom */
        if (b >= 0 && b < size) en[b]++; /* end-1 inside chrom                        */
        else en[size - 1]++;             /* otherwise clipped to the last base        */
    }
    long nout = 0;
    uint32_t depth = 0;
    int64_t lastStart = -1;
    int64_t lastDepth = -1;
    for (int64_t pos = 0; pos < size; pos++) {
        depth += st[pos];
        if ((int64_t)depth != lastDepth) {
            if (lastDepth > 0) {
                if (nout >= cap) { free(st); free(en); return -1; }
                out_start[nout] = lastStart; out_end[nout] = pos; out_depth[nout] = lastDepth; nout++;
            }
            lastDepth = depth;
            lastStart = pos;
        }
        depth -= en[pos];
    }
    if (lastDepth > 0) {
        if (nout >= cap) { free(st); free(en); return -1; }
        out_start[nout] = lastStart; out_end[nout] = size; out_depth[nout] = lastDepth; nout++;
    }
    free(st); free(en);
    return nout;
}

/* merge for ONE chromosome; input must be sorted by start. In-place safe
 * (out_* may alias in_*). returns number of merged rows. */
long ora_merge(const int64_t *s, const int64_t *e, long n, int64_t *out_start, int64_t *out_end)
{
    long nout = 0;
    if (n == 0) return 0;
    int64_t cs = s[0], ce = e[0];
    for (long i = 1; i < n; i++) {
        if (s[i] <= ce) { if (e[i] > ce) ce = e[i]; }
        else { out_start[nout] = cs; out_end[nout] = ce; nout++; cs = s[i]; ce = e[i]; }
    }
    out_start[nout] = cs; out_end[nout] = ce; nout++;
    return nout;
}


/* Whole annotation core on arrays (SURVEY.md 9.2 steps G,T,M,L for every chromosome):
 * hits given as (chrom index, start, end); chromosomes are reported in index
 * order (the caller numbers them in `sort -k 1,1` byte order). Output rows are
 * the merged runs with depth >= cov and end-start >= min_len.
 * returns number of segments, -1 if cap too small, -4 on an invalid hit. */
long ora_coverage_segments(const int32_t *chrom, const int32_t *start, const int32_t *end, long nhits,
                           const int64_t *sizes, int nchrom, int cov, int min_len,
                           int32_t *out_chrom, int32_t *out_start, int32_t *out_end, long cap)
{
    long *cnt = (long *)calloc((size_t)nchrom + 1, sizeof(long));
    for (long i = 0; i < nhits; i++) {
        if (chrom[i] < 0 || chrom[i] >= nchrom || start[i] < 0 || start[i] > end[i]) { free(cnt); return -4; }
        cnt[chrom[i] + 1]++;
    }
    for (int c = 0; c < nchrom; c++) cnt[c + 1] += cnt[c];
    int64_t *s = (int64_t *)malloc((size_t)(nhits ? nhits : 1) * sizeof(int64_t));
    int64_t *e = (int64_t *)malloc((size_t)(nhits ? nhits : 1) * sizeof(int64_t));
    long *fill = (long *)malloc((size_t)(nchrom + 1) * sizeof(long));
    memcpy(fill, cnt, (size_t)(nchrom + 1) * sizeof(long));
    for (long i = 0; i < nhits; i++) { long k = fill[chrom[i]]++; s[k] = start[i]; e[k] = end[i]; }
    long nout = 0;
    for (int c = 0; c < nchrom; c++) {
        long n = cnt[c + 1] - cnt[c];
        if (n == 0) continue;
        long bcap = 2 * n + 2;
        int64_t *bs = (int64_t *)malloc(bcap * sizeof(int64_t));
        int64_t *be = (int64_t *)malloc(bcap * sizeof(int64_t));
        int64_t *bd = (int64_t *)malloc(bcap * sizeof(int64_t));
        long nb = ora_genomecov_bg(s + cnt[c], e + cnt[c], n, sizes[c], bs, be, bd, bcap);
        if (nb < 0) { free(bs); free(be); free(bd); free(s); free(e); free(fill); free(cnt); return nb; }
        long k = 0;                       /* T: keep depth >= cov */
        for (long i = 0; i < nb; i++) if (bd[i] >= cov) { bs[k] = bs[i]; be[k] = be[i]; k++; }
        long nm = ora_merge(bs, be, k, bs, be);   /* M */
        for (long i = 0; i < nm; i++) {
            if (be[i] - bs[i] >= min_len) {       /* L */
                if (nout >= cap) { free(bs); free(be); free(bd); free(s); free(e); free(fill); free(cnt); return -1; }
                out_chrom[nout] = c; out_start[nout] = (int32_t)bs[i]; out_end[nout] = (int32_t)be[i]; nout++;
            }
        }
        free(bs); free(be); free(bd);
    }
    free(s); free(e); free(fill); free(cnt);
    return nout;
}

/* ------------------------------------------------------------------ */
/* CLI shim                                                             */
/* ------------------------------------------------------------------ */
#ifdef ORACLE_MAIN

typedef struct { char *name; int64_t size; } chrom_t;

static void die(const char *msg, const char *arg)
{
    fprintf(stderr, "bedtools-oracle: %s%s%s\n", msg, arg ? ": " : "", arg ? arg : "");
    exit(1);
}

static chrom_t *read_genome(const char *path, long *n)
{
    FILE *f = fopen(path, "r");
    if (!f) die("cannot open genome file", path);
    chrom_t *g = NULL; long cap = 0; *n = 0;
    char *line = NULL; size_t lcap = 0; ssize_t len;
    while ((len = getline(&line, &lcap, f)) > 0) {
        char *tab = strchr(line, '\t');
        if (!tab) continue;
        *tab = 0;
        if (*n == cap) { cap = cap ? cap * 2 : 64; g = (chrom_t *)realloc(g, cap * sizeof(chrom_t)); }
        g[*n].name = strdup(line);
        g[*n].size = strtoll(tab + 1, NULL, 10);
        (*n)++;
    }
    free(line); fclose(f);
    return g;
}

typedef struct { int64_t *s, *e; long n, cap; } ivs_t;
static void ivs_push(ivs_t *v, int64_t s, int64_t e)
{
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 1024;
        v->s = (int64_t *)realloc(v->s, v->cap * sizeof(int64_t));
        v->e = (int64_t *)realloc(v->e, v->cap * sizeof(int64_t));
    }
    v->s[v->n] = s; v->e[v->n] = e; v->n++;
}

static int parse_bed3(char *line, char **chrom, int64_t *s, int64_t *e)
{
    char *p = line;
    char *t1 = strchr(p, '\t'); if (!t1) return 0;
    *t1 = 0; *chrom = p;
    char *endp;
    *s = strtoll(t1 + 1, &endp, 10);
    if (endp == t1 + 1 || *endp != '\t') return 0;
    char *q = endp + 1;
    *e = strtoll(q, &endp, 10);
    if (endp == q) return 0;
    return 1;
}

static void flush_cov(const char *chrom, ivs_t *v, chrom_t *g, long ng)
{
    if (!chrom) return;
    int64_t size = -1;
    for (long i = 0; i < ng; i++) if (!strcmp(g[i].name, chrom)) { size = g[i].size; break; }
    if (size < 0) die("chromosome found in BED but not in genome file", chrom);
    long cap = 2 * v->n + 2;
    int64_t *os = (int64_t *)malloc(cap * sizeof(int64_t));
    int64_t *oe = (int64_t *)malloc(cap * sizeof(int64_t));
    int64_t *od = (int64_t *)malloc(cap * sizeof(int64_t));
    long n = ora_genomecov_bg(v->s, v->e, v->n, size, os, oe, od, cap);
    if (n < 0) die("genomecov sweep failed", chrom);
    for (long i = 0; i < n; i++)
        printf("%s\t%lld\t%lld\t%lld\n", chrom, (long long)os[i], (long long)oe[i], (long long)od[i]);
    free(os); free(oe); free(od);
    v->n = 0;
}

static void flush_merge(const char *chrom, ivs_t *v)
{
    if (!chrom || v->n == 0) return;
    long n = ora_merge(v->s, v->e, v->n, v->s, v->e);
    for (long i = 0; i < n; i++)
        printf("%s\t%lld\t%lld\n", chrom, (long long)v->s[i], (long long)v->e[i]);
    v->n = 0;
}

int main(int argc, char **argv)
{
    if (argc < 2) die("usage: bedtools {genomecov -bg -i BED -g LENS | merge -i BED}", NULL);
    const char *sub = argv[1];
    const char *in = NULL, *gen = NULL; int bg = 0;
    for (int i = 2; i < argc; i++) {
        if (!strcmp(argv[i], "-i") && i + 1 < argc) in = argv[++i];
        else if (!strcmp(argv[i], "-g") && i + 1 < argc) gen = argv[++i];
        else if (!strcmp(argv[i], "-bg")) bg = 1;
        else die("unsupported option", argv[i]);
    }
    if (!in) die("-i is required", NULL);
    FILE *f = fopen(in, "r");
    if (!f) die("cannot open", in);
    int is_cov = !strcmp(sub, "genomecov");
    if (!is_cov && strcmp(sub, "merge")) die("unsupported sub-command", sub);
    chrom_t *g = NULL; long ng = 0;
    if (is_cov) {
        if (!bg || !gen) die("genomecov needs -bg and -g", NULL);
        g = read_genome(gen, &ng);
    }
    ivs_t v = {0};
    char *cur = NULL;
    char *line = NULL; size_t lcap = 0; ssize_t len; long lineno = 0;
    int64_t prev_start = -1;
    while ((len = getline(&line, &lcap, f)) > 0) {
        lineno++;
        while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
        if (len == 0) continue;
        char *chrom; int64_t s, e;
        if (!parse_bed3(line, &chrom, &s, &e)) { fprintf(stderr, "bedtools-oracle: malformed BED entry at line %ld\n", lineno); return 1; }
        if (s > e) { fprintf(stderr, "bedtools-oracle: malformed BED entry at line %ld. Start was greater than end.\n", lineno); return 1; }
        if (s < 0) { fprintf(stderr, "bedtools-oracle: malformed BED entry at line %ld. Negative start.\n", lineno); return 1; }
        if (!cur || strcmp(cur, chrom)) {
            if (is_cov) flush_cov(cur, &v, g, ng); else flush_merge(cur, &v);
            free(cur); cur = strdup(chrom); prev_start = -1;
        }
        if (!is_cov && s < prev_start) { fprintf(stderr, "bedtools-oracle: merge input is not sorted at line %ld\n", lineno); return 1; }
        prev_start = s;
        ivs_push(&v, s, e);
    }
    if (is_cov) flush_cov(cur, &v, g, ng); else flush_merge(cur, &v);
    free(line); fclose(f);
    return 0;
}
#endif
