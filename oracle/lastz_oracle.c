/*
 * lastz_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle; never on the product path).
 *
 * "LASTZ-restatement": a plain-C restatement of the alignment half of mimeo's hot path, i.e.
 * of what the external program LASTZ computes for the one command line mimeo ever issues
 *
 *   lastz T.fa Q.fa --entropy --format=general:name1,strand1,start1,end1,length1,name2,strand2,
 *         start2+,end2+,length2,score,identity --markend --gfextend --chain --gapped --step=1
 *         --strand=both --hspthresh=K
 *   (reference call sites: src/mimeo/wrappers.py:1025-1037, 1070-1082, 786-798, 645-653)
 *
 * LASTZ is a third-party dependency that is absent from /root/reference (environment.yml:7,
 * no version pin) and is not installed in this image or on the GPU boxes, and the reference
 * has no tests, fixtures or golden vectors for it (tests/test_dummy.py only):
 *
 *      ****  PARITY UNPINNED against a real LASTZ binary.  ****
 *
 * What is restated is LASTZ's published algorithm with its default parameters:
 *   scoring   HOXD70 (A/C/G/T rows: 91 -114 -31 -123 | -114 100 -125 -31 | -31 -125 100 -114 |
 *             -123 -31 -114 91), any non-ACGT character scores -100, gap open 400 / extend 30,
 *             x-drop 910, y-drop 9400, hspthresh K (default 3000), gappedthresh = K.
 *   seeds     12-of-19 spaced seed 1110100110010101111, every target position (--step=1), exact
 *             match on the 12 care positions or exactly one transition (A<->G, C<->T) among them;
 *             a window containing a non-ACGT character is never a seed.
 *   HSPs      gap-free x-drop extension (--gfextend): right from the seed end, left from the seed
 *             end through the seed, each keeping its running maximum and stopping when the running
 *             sum falls more than 910 below it; keep if score >= K; --entropy multiplies the score
 *             by the base-4 Shannon entropy of the matched columns before the K test.
 *   chain     --chain with zero penalties: per (target, query, strand) the maximum-total-score
 *             subset of HSPs that is strictly increasing in both sequences.
 *   gapped    each chained HSP is reduced to an anchor (centre of its best 31-column window);
 *             anchors are extended best-first with an affine-gap y-drop DP in both directions;
 *             anchors already inside a reported alignment are skipped; keep if score >= K.
 *
 * Where LASTZ's behaviour is an artefact of its sequential implementation, this file fixes ONE
 * deterministic, order-independent definition, which is the normative spec that the CUDA path
 * reproduces bit-exactly (DESIGN.md lists every such choice as a stated deviation):
 *   D1  seed hits on one diagonal are first reduced to run leaders (a hit whose predecessor
 *       (i-1,j-1) is not a hit); only leaders are extension candidates.
 *   D2  per diagonal, leaders are visited in increasing position; a leader whose 19-mer ends at or
 *       before the end of the last KEPT HSP on that diagonal is skipped. Failed extensions leave
 *       no trace (LASTZ also remembers failed extents; that only saves work).
 *   D3  entropy is evaluated in integer fixed point (Q24 log2), score' = (score * h) >> 24.
 *   D4  y-drop pruning is applied per anti-diagonal: a cell survives iff H >= best - 9400 where
 *       best is the maximum over all earlier anti-diagonals. (LASTZ prunes row by row.)
 *   D5  "inside a reported alignment" is tested on the alignment's bounding box, and extensions
 *       are not clipped by earlier alignments; no traceback-memory truncation exists.
 *   (D6 of round 1 -- soft-masking ignored -- is gone: lower-case bases are excluded from seeding on both sequences and
 *   extended through by their base, as LASTZ does without [unmask].)
 *
 * Coordinates inside this file: 0-based, half-open, on the strand that was aligned (the caller
 * reverse-complements the query for the minus strand and converts back).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 */
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SEED_SPAN 19
#define NEG_INF (INT_MIN / 4)

/* Codes: 0..3 = A,C,G,T, 4 = anything else; bit 3 set = soft-masked (lower case in the input). A soft-masked base is
 * never part of a seed word (every test below that asks "code > 3" rejects it) but is extended through by its base. */
#define BASE(x) ((x) & 7)
static const int SUB[5][5] = {
    {91, -114, -31, -123, -100},
    {-114, 100, -125, -31, -100},
    {-31, -125, 100, -114, -100},
    {-123, -31, -114, 91, -100},
    {-100, -100, -100, -100, -100},
};
/* care positions of 1110100110010101111 */
static const int CARE[12] = {0, 1, 2, 4, 7, 8, 11, 13, 15, 16, 17, 18};

typedef struct {
    int32_t hspthresh, xdrop, ydrop, gap_open, gap_extend, gappedthresh;
    int32_t entropy, chain, gapped, transition; /* flags, 1 = as mimeo runs LASTZ */
} lzo_params;

typedef struct { int32_t s1, s2, len, score; } lzo_hsp; /* target start, query start, length, score */

typedef struct {
    int32_t s1, e1, s2, e2; /* box, half-open */
    int32_t score, nmatch, ncols;
    int32_t a1, a2;         /* anchor point it grew from */
} lzo_aln;

typedef struct {
    int64_t seed_hits, leaders, extended, ungapped_cells, hsps_raw, hsps_kept, chained, anchors_extended, gapped_cells;
} lzo_stats;

void lzo_default_params(lzo_params *p)
{
    p->hspthresh = 3000; p->xdrop = 910; p->ydrop = 9400; p->gap_open = 400; p->gap_extend = 30;
    p->gappedthresh = 3000; p->entropy = 1; p->chain = 1; p->gapped = 1; p->transition = 1;
}

/* ------------------------------------------------------------------ seeds */
int lzo_seed_at(const uint8_t *t, long n, const uint8_t *q, long m, long i, long j, int transition)
{
    if (i < 0 || j < 0 || i + SEED_SPAN > n || j + SEED_SPAN > m) return 0;
    for (int c = 0; c < SEED_SPAN; c++)
        if (t[i + c] > 3 || q[j + c] > 3) return 0;
    int ts = 0;
    for (int k = 0; k < 12; k++) {
        int a = t[i + CARE[k]], b = q[j + CARE[k]];
        if (a != b) {
            if ((a ^ b) == 2 && transition) ts++;
            else return 0;
        }
    }
    return ts <= 1;
}

static inline int seed_key(const uint8_t *s, long i) /* caller guarantees a clean window */
{
    int key = 0;
    for (int k = 0; k < 12; k++) key = (key << 2) | s[i + CARE[k]];
    return key;
}

static inline int window_clean(const uint8_t *s, long n, long i)
{
    if (i < 0 || i + SEED_SPAN > n) return 0;
    for (int c = 0; c < SEED_SPAN; c++) if (s[i + c] > 3) return 0;
    return 1;
}

/* ------------------------------------------------------------------ fixed-point entropy (D3) */
static uint32_t log2_q24(uint32_t x) /* x >= 1 */
{
    int ip = 31 - __builtin_clz(x);
    uint64_t y = (uint64_t)x << (31 - ip); /* Q31 in [1,2) */
    uint32_t frac = 0;
    for (int k = 0; k < 24; k++) {
        y = (y * y) >> 31;
        frac <<= 1;
        if (y >= (1ull << 32)) { y >>= 1; frac |= 1; }
    }
    return ((uint32_t)ip << 24) | frac;
}

/* h = entropy of the matched columns in Q24 (1.0 == 1<<24) */
uint32_t lzo_entropy_q24(const uint32_t cnt[4])
{
    uint32_t n = cnt[0] + cnt[1] + cnt[2] + cnt[3];
    if (n == 0) return 0;
    uint32_t ln = log2_q24(n);
    uint64_t T = 0;
    for (int b = 0; b < 4; b++)
        if (cnt[b]) T += (uint64_t)cnt[b] * (uint64_t)(ln - log2_q24(cnt[b]));
    return (uint32_t)(T / (2ull * n));
}

/* ------------------------------------------------------------------ gap-free extension */
static void xdrop_extend(const uint8_t *t, long n, const uint8_t *q, long m, long i, long j, int X,
                         long *bstart, long *bend, int *score, int64_t *cells)
{
    long c1 = i + SEED_SPAN, c2 = j + SEED_SPAN;
    int run = 0, best = 0;
    long be = c1;
    while (c1 < n && c2 < m) {
        run += SUB[BASE(t[c1])][BASE(q[c2])];
        c1++; c2++; (*cells)++;
        if (run > best) { best = run; be = c1; }
        else if (run < best - X) break;
    }
    c1 = i + SEED_SPAN - 1; c2 = j + SEED_SPAN - 1;
    int runl = 0, bestl = 0;
    long bs = i + SEED_SPAN;
    while (c1 >= 0 && c2 >= 0) {
        runl += SUB[BASE(t[c1])][BASE(q[c2])];
        (*cells)++;
        if (runl > bestl) { bestl = runl; bs = c1; }
        else if (runl < bestl - X) break;
        c1--; c2--;
    }
    *bstart = bs; *bend = be; *score = best + bestl;
}

/* Target position table: counting sort of every clean 19-window start by its 24-bit seed key. */
typedef struct { uint32_t *off; uint32_t *pos; long npos; long n; } lzo_index;

lzo_index *lzo_index_build(const uint8_t *t, long n)
{
    const long NB = 1L << 24;
    lzo_index *ix = (lzo_index *)calloc(1, sizeof(lzo_index));
    ix->n = n;
    ix->off = (uint32_t *)calloc((size_t)NB + 1, sizeof(uint32_t));
    int32_t *keyof = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
    long npos = 0;
    for (long i = 0; i + SEED_SPAN <= n; i++) {
        keyof[i] = window_clean(t, n, i) ? seed_key(t, i) : -1;
        if (keyof[i] >= 0) { ix->off[keyof[i] + 1]++; npos++; }
    }
    for (long k = 0; k < NB; k++) ix->off[k + 1] += ix->off[k];
    ix->pos = (uint32_t *)malloc((size_t)(npos ? npos : 1) * sizeof(uint32_t));
    uint32_t *fill = (uint32_t *)malloc((size_t)NB * sizeof(uint32_t));
    memcpy(fill, ix->off, (size_t)NB * sizeof(uint32_t));
    for (long i = 0; i + SEED_SPAN <= n; i++) if (keyof[i] >= 0) ix->pos[fill[keyof[i]]++] = (uint32_t)i;
    ix->npos = npos;
    free(fill); free(keyof);
    return ix;
}
void lzo_index_free(lzo_index *ix) { if (ix) { free(ix->off); free(ix->pos); free(ix); } }

/* All kept HSPs of one (target, query-strand) tile. ix may be NULL (then the table is built here, which is
 * what one LASTZ process per scaffold pair does). Returns count, -1 if cap too small. */
long lzo_hsps_ix(const lzo_index *ix_in, const uint8_t *t, long n, const uint8_t *q, long m, const lzo_params *p,
                 lzo_hsp *out, long cap, lzo_stats *st)
{
    lzo_stats local; memset(&local, 0, sizeof(local)); if (!st) st = &local;
    if (n < SEED_SPAN || m < SEED_SPAN) return 0;
    lzo_index *own = ix_in ? NULL : lzo_index_build(t, n);
    const lzo_index *ix = ix_in ? ix_in : own;
    const uint32_t *off = ix->off, *pos = ix->pos;

    /* covered_end per diagonal d = i - j, index d + m */
    int32_t *covered = (int32_t *)malloc((size_t)(n + m + 1) * sizeof(int32_t));
    for (long k = 0; k < n + m + 1; k++) covered[k] = -1;

    long nout = 0;
    for (long j = 0; j + SEED_SPAN <= m; j++) {
        if (!window_clean(q, m, j)) continue;
        int key = seed_key(q, j);
        int nprobe = p->transition ? 13 : 1;
        for (int pr = 0; pr < nprobe; pr++) {
            int k = pr == 0 ? key : (key ^ (2 << (2 * (pr - 1)))); /* flip the transition bit of one care base */
            for (uint32_t x = off[k]; x < off[k + 1]; x++) {
                long i = pos[x];
                st->seed_hits++;
                if (lzo_seed_at(t, n, q, m, i - 1, j - 1, p->transition)) continue; /* D1: not a run leader */
                st->leaders++;
                long d = i - j + m;
                if (i + SEED_SPAN <= covered[d]) continue; /* D2 */
                long bs, be; int score;
                st->extended++;
                xdrop_extend(t, n, q, m, i, j, p->xdrop, &bs, &be, &score, &st->ungapped_cells);
                if (score < p->hspthresh) continue;
                st->hsps_raw++;
                if (p->entropy) {
                    uint32_t cnt[4] = {0, 0, 0, 0};
                    for (long c = bs; c < be; c++) if (BASE(t[c]) == BASE(q[c - (i - j)]) && BASE(t[c]) < 4) cnt[BASE(t[c])]++;
                    uint32_t h = lzo_entropy_q24(cnt);
                    score = (int)(((int64_t)score * (int64_t)h) >> 24);
                    if (score < p->hspthresh) continue;
                }
                if (nout >= cap) { lzo_index_free(own); free(covered); return -1; }
                out[nout].s1 = (int32_t)bs; out[nout].s2 = (int32_t)(bs - (i - j));
                out[nout].len = (int32_t)(be - bs); out[nout].score = score; nout++;
                st->hsps_kept++;
                if (be > covered[d]) covered[d] = (int32_t)be;
            }
        }
    }
    lzo_index_free(own); free(covered);
    return nout;
}
long lzo_hsps(const uint8_t *t, long n, const uint8_t *q, long m, const lzo_params *p, lzo_hsp *out, long cap,
              lzo_stats *st)
{
    return lzo_hsps_ix(NULL, t, n, q, m, p, out, cap, st);
}

/* ------------------------------------------------------------------ chain */
static int hsp_cmp(const void *a, const void *b)
{
    const lzo_hsp *x = (const lzo_hsp *)a, *y = (const lzo_hsp *)b;
    if (x->s1 != y->s1) return x->s1 < y->s1 ? -1 : 1;
    if (x->s2 != y->s2) return x->s2 < y->s2 ? -1 : 1;
    if (x->len != y->len) return x->len < y->len ? -1 : 1;
    if (x->score != y->score) return x->score < y->score ? -1 : 1;
    return 0;
}
void lzo_sort_hsps(lzo_hsp *h, long n) { qsort(h, (size_t)n, sizeof(lzo_hsp), hsp_cmp); }

typedef struct { int64_t c; int32_t idx; } chv_t;
static inline int chv_better(chv_t a, chv_t b) /* a strictly better than b */
{
    if (a.c != b.c) return a.c > b.c;
    return a.idx < b.idx;
}
static const long *g_e1;
static int by_e1(const void *a, const void *b)
{
    long x = *(const long *)a, y = *(const long *)b;
    if (g_e1[x] != g_e1[y]) return g_e1[x] < g_e1[y] ? -1 : 1;
    return x < y ? -1 : (x > y);
}
static int cmp_long(const void *a, const void *b)
{
    long x = *(const long *)a, y = *(const long *)b;
    return x < y ? -1 : (x > y);
}

/* Best collinear chain. hsps are sorted in place into canonical order (s1,s2,len,score);
 * in_chain[k] = 1 for members. Returns number of members. */
long lzo_chain(lzo_hsp *h, long n, uint8_t *in_chain)
{
    if (n == 0) return 0;
    lzo_sort_hsps(h, n);
    long *e1 = (long *)malloc(n * sizeof(long)), *e2 = (long *)malloc(n * sizeof(long));
    long *ord = (long *)malloc(n * sizeof(long)), *ue2 = (long *)malloc(n * sizeof(long));
    for (long k = 0; k < n; k++) { e1[k] = (long)h[k].s1 + h[k].len; e2[k] = (long)h[k].s2 + h[k].len; ord[k] = k; ue2[k] = e2[k]; }
    g_e1 = e1; qsort(ord, n, sizeof(long), by_e1);
    qsort(ue2, n, sizeof(long), cmp_long);
    long nu = 0;
    for (long k = 0; k < n; k++) if (k == 0 || ue2[k] != ue2[k - 1]) ue2[nu++] = ue2[k];
    chv_t *bit = (chv_t *)malloc((nu + 1) * sizeof(chv_t));
    for (long k = 0; k <= nu; k++) { bit[k].c = 0; bit[k].idx = INT32_MAX; } /* "no predecessor" */
    int64_t *C = (int64_t *)malloc(n * sizeof(int64_t));
    int32_t *pred = (int32_t *)malloc(n * sizeof(int32_t));
    long ins = 0;
    for (long b = 0; b < n; b++) { /* canonical order = increasing s1 */
        while (ins < n && e1[ord[ins]] <= h[b].s1) { /* make every a with e1_a <= s1_b visible */
            long a = ord[ins++];
            /* rank of e2[a] among unique values (1-based) */
            long lo = 0, hi = nu; while (lo < hi) { long mid = (lo + hi) / 2; if (ue2[mid] < e2[a]) lo = mid + 1; else hi = mid; }
            chv_t v = {C[a], (int32_t)a};
            for (long x = lo + 1; x <= nu; x += x & (-x)) if (chv_better(v, bit[x])) bit[x] = v;
        }
        /* prefix max over e2 <= s2_b */
        long lo = 0, hi = nu; while (lo < hi) { long mid = (lo + hi) / 2; if (ue2[mid] <= h[b].s2) lo = mid + 1; else hi = mid; }
        chv_t best = {0, INT32_MAX};
        for (long x = lo; x > 0; x -= x & (-x)) if (chv_better(bit[x], best)) best = bit[x];
        pred[b] = best.idx == INT32_MAX ? -1 : best.idx;
        C[b] = (int64_t)h[b].score + (pred[b] >= 0 ? best.c : 0);
    }
    /* NOTE: insertion happens lazily, so an HSP a is only visible to later b with s1_b >= e1_a; correct since
     * canonical order is by s1 and e1_a > s1_a. */
    long end = 0;
    for (long k = 1; k < n; k++) if (C[k] > C[end]) end = k;
    memset(in_chain, 0, (size_t)n);
    long cnt = 0;
    for (long k = end; k >= 0; k = pred[k]) { in_chain[k] = 1; cnt++; }
    free(e1); free(e2); free(ord); free(ue2); free(bit); free(C); free(pred);
    return cnt;
}

/* ------------------------------------------------------------------ anchors */
void lzo_anchor(const uint8_t *t, const uint8_t *q, const lzo_hsp *h, int32_t *a1, int32_t *a2)
{
    int off;
    if (h->len <= 31) off = h->len / 2;
    else {
        int sum = 0;
        for (int c = 0; c < 31; c++) sum += SUB[BASE(t[h->s1 + c])][BASE(q[h->s2 + c])];
        int best = sum, bw = 0;
        for (int w = 1; w + 31 <= h->len; w++) {
            sum += SUB[BASE(t[h->s1 + w + 30])][BASE(q[h->s2 + w + 30])] - SUB[BASE(t[h->s1 + w - 1])][BASE(q[h->s2 + w - 1])];
            if (sum > best) { best = sum; bw = w; }
        }
        off = bw + 15;
    }
    *a1 = h->s1 + off; *a2 = h->s2 + off;
}

/* ------------------------------------------------------------------ y-drop gapped extension (D4) */
typedef struct { int h, d, i; int hm, hc, dm, dc, im, ic; } cell_t; /* scores + (nmatch, ncols) payload per state */

typedef struct { int score, di, dj, nmatch, ncols; } ext_t;

/* dir=+1: cell (i,j) consumes t[a1 + i - 1], q[a2 + j - 1];  dir=-1: t[a1 - i], q[a2 - j].
 * tn/qn: bases available in that direction. */
static ext_t ydrop_extend(const uint8_t *t, long a1, long tn, const uint8_t *q, long a2, long qn, int dir,
                          const lzo_params *p, int64_t *cells)
{
    const int O = p->gap_open, E = p->gap_extend, Y = p->ydrop;
    ext_t r = {0, 0, 0, 0, 0};
    long cap = 1024;
    cell_t *A = (cell_t *)malloc(cap * sizeof(cell_t)), *B = (cell_t *)malloc(cap * sizeof(cell_t)),
           *Cc = (cell_t *)malloc(cap * sizeof(cell_t));
    cell_t *p2 = A, *p1 = B, *cur = Cc; /* k-2, k-1, k ; each indexed by i - lo */
    long lo2 = 0, hi2 = -1, lo1 = 0, hi1 = 0; /* empty k-2 ; k-1 is anti-diagonal 0 */
    p1[0].h = 0; p1[0].d = NEG_INF; p1[0].i = NEG_INF; p1[0].hm = p1[0].hc = p1[0].dm = p1[0].dc = p1[0].im = p1[0].ic = 0;
    int best = 0;
    for (long k = 1; k <= tn + qn; k++) {
        long clo = LONG_MAX, chi = LONG_MIN;
        if (hi1 >= lo1) { clo = lo1; chi = hi1 + 1; }
        if (hi2 >= lo2) { if (lo2 + 1 < clo) clo = lo2 + 1; if (hi2 + 1 > chi) chi = hi2 + 1; }
        if (clo == LONG_MAX) break; /* both previous anti-diagonals dead */
        if (clo < 0) clo = 0;
        if (clo < k - qn) clo = k - qn;
        if (chi > k) chi = k;
        if (chi > tn) chi = tn;
        long nlo = 0, nhi = -1;
        if (chi >= clo) {
            long w = chi - clo + 1;
            if (w > cap) {
                long o2 = p2 - A, o1 = p1 - A, oc = cur - A; (void)o2; (void)o1; (void)oc;
                /* grow all three buffers, preserving contents */
                long ncap = w * 2;
                cell_t *nA = (cell_t *)malloc(ncap * sizeof(cell_t)), *nB = (cell_t *)malloc(ncap * sizeof(cell_t)), *nC = (cell_t *)malloc(ncap * sizeof(cell_t));
                memcpy(nA, A, cap * sizeof(cell_t)); memcpy(nB, B, cap * sizeof(cell_t)); memcpy(nC, Cc, cap * sizeof(cell_t));
                cell_t *np2 = p2 == A ? nA : (p2 == B ? nB : nC), *np1 = p1 == A ? nA : (p1 == B ? nB : nC), *ncur = cur == A ? nA : (cur == B ? nB : nC);
                free(A); free(B); free(Cc); A = nA; B = nB; Cc = nC; p2 = np2; p1 = np1; cur = ncur; cap = ncap;
            }
            const int thr = best - Y;
            int curbest = best;
            for (long i = clo; i <= chi; i++) {
                long j = k - i;
                cell_t c; c.h = c.d = c.i = NEG_INF; c.hm = c.hc = c.dm = c.dc = c.im = c.ic = 0;
                /* up: (i-1, j) on k-1 */
                if (i - 1 >= lo1 && i - 1 <= hi1) {
                    const cell_t *u = &p1[i - 1 - lo1];
                    if (u->h > NEG_INF) {
                        int open = u->h - O - E, ext = u->d > NEG_INF ? u->d - E : NEG_INF;
                        if (open >= ext) { c.d = open; c.dm = u->hm; c.dc = u->hc; } else { c.d = ext; c.dm = u->dm; c.dc = u->dc; }
                    }
                }
                /* left: (i, j-1) on k-1 */
                if (i >= lo1 && i <= hi1) {
                    const cell_t *l = &p1[i - lo1];
                    if (l->h > NEG_INF) {
                        int open = l->h - O - E, ext = l->i > NEG_INF ? l->i - E : NEG_INF;
                        if (open >= ext) { c.i = open; c.im = l->hm; c.ic = l->hc; } else { c.i = ext; c.im = l->im; c.ic = l->ic; }
                    }
                }
                /* diag: (i-1, j-1) on k-2 */
                int mval = NEG_INF, mm = 0, mc = 0;
                if (i >= 1 && j >= 1 && i - 1 >= lo2 && i - 1 <= hi2) {
                    const cell_t *dg = &p2[i - 1 - lo2];
                    if (dg->h > NEG_INF) {
                        int a = BASE(dir > 0 ? t[a1 + i - 1] : t[a1 - i]), b = BASE(dir > 0 ? q[a2 + j - 1] : q[a2 - j]);
                        mval = dg->h + SUB[a][b]; mm = dg->hm + (a == b && a < 4); mc = dg->hc + 1;
                    }
                }
                if (mval >= c.d && mval >= c.i) { c.h = mval; c.hm = mm; c.hc = mc; }
                else if (c.d >= c.i) { c.h = c.d; c.hm = c.dm; c.hc = c.dc; }
                else { c.h = c.i; c.hm = c.im; c.hc = c.ic; }
                (*cells)++;
                if (c.h <= NEG_INF || c.h < thr) { c.h = c.d = c.i = NEG_INF; }
                else {
                    if (nhi < nlo) nlo = i;
                    nhi = i;
                    if (c.h > curbest) { curbest = c.h; r.score = c.h; r.di = (int)i; r.dj = (int)j; r.nmatch = c.hm; r.ncols = c.hc; }
                }
                cur[i - clo] = c;
            }
            best = curbest;
        }
        /* rotate: the new anti-diagonal keeps base clo; trim to alive range */
        cell_t *tmp = p2; p2 = p1; lo2 = lo1; hi2 = hi1; p1 = cur; cur = tmp;
        if (nhi >= nlo) {
            if (nlo > clo) memmove(p1, p1 + (nlo - clo), (size_t)(nhi - nlo + 1) * sizeof(cell_t));
            lo1 = nlo; hi1 = nhi;
        } else { lo1 = 0; hi1 = -1; }
    }
    free(A); free(B); free(Cc);
    return r;
}

/* Test hook: one one-sided extension. out = {score, di, dj, nmatch, ncols}; returns the number of DP cells. */
long lzo_ydrop_extend(const uint8_t *t, long a1, long tn, const uint8_t *q, long a2, long qn, int dir, const lzo_params *p,
                      int32_t *out)
{
    int64_t cells = 0;
    ext_t r = ydrop_extend(t, a1, tn, q, a2, qn, dir, p, &cells);
    out[0] = r.score; out[1] = r.di; out[2] = r.dj; out[3] = r.nmatch; out[4] = r.ncols;
    return (long)cells;
}

/* Gapped stage for one tile-strand: chained HSPs in, alignments out. */
long lzo_gapped(const uint8_t *t, long n, const uint8_t *q, long m, const lzo_hsp *h, long nh, const lzo_params *p,
                lzo_aln *out, long cap, lzo_stats *st)
{
    lzo_stats local; if (!st) st = &local;
    long *ord = (long *)malloc((nh ? nh : 1) * sizeof(long));
    for (long k = 0; k < nh; k++) ord[k] = k;
    /* best-first: score desc, then s1 asc, s2 asc (insertion sort is fine for a test oracle up to ~1e4; use qsort) */
    for (long a = 1; a < nh; a++) {
        long v = ord[a], b = a - 1;
        while (b >= 0) {
            const lzo_hsp *x = &h[ord[b]], *y = &h[v];
            int after = (x->score < y->score) || (x->score == y->score && (x->s1 > y->s1 || (x->s1 == y->s1 && x->s2 > y->s2)));
            if (!after) break;
            ord[b + 1] = ord[b]; b--;
        }
        ord[b + 1] = v;
    }
    long nout = 0;
    for (long a = 0; a < nh; a++) {
        const lzo_hsp *hh = &h[ord[a]];
        int32_t a1, a2;
        lzo_anchor(t, q, hh, &a1, &a2);
        int covered = 0;
        for (long k = 0; k < nout && !covered; k++)
            if (a1 >= out[k].s1 && a1 < out[k].e1 && a2 >= out[k].s2 && a2 < out[k].e2) covered = 1; /* D5 */
        if (covered) continue;
        st->anchors_extended++;
        ext_t f = ydrop_extend(t, a1, n - a1, q, a2, m - a2, +1, p, &st->gapped_cells);
        ext_t b = ydrop_extend(t, a1, a1, q, a2, a2, -1, p, &st->gapped_cells);
        int score = f.score + b.score;
        if (score < p->gappedthresh) continue;
        if (nout >= cap) { free(ord); return -1; }
        lzo_aln *o = &out[nout++];
        o->s1 = a1 - b.di; o->e1 = a1 + f.di; o->s2 = a2 - b.dj; o->e2 = a2 + f.dj;
        o->score = score; o->nmatch = f.nmatch + b.nmatch; o->ncols = f.ncols + b.ncols; o->a1 = a1; o->a2 = a2;
    }
    free(ord);
    return nout;
}

/* Whole pipeline for one (target, query-strand) tile. Returns number of alignments (or of HSPs when
 * gapped is off: then each HSP is reported as an ungapped alignment). */
long lzo_align_tile_ix(const lzo_index *ix, const uint8_t *t, long n, const uint8_t *q, long m, const lzo_params *p,
                       lzo_aln *out, long cap, lzo_stats *st)
{
    lzo_stats local; memset(&local, 0, sizeof(local)); if (!st) st = &local;
    long hcap = 1 << 16, nh;
    lzo_hsp *h = NULL;
    for (;;) {
        h = (lzo_hsp *)malloc(hcap * sizeof(lzo_hsp));
        lzo_stats s0 = *st;
        nh = lzo_hsps_ix(ix, t, n, q, m, p, h, hcap, st);
        if (nh >= 0) break;
        *st = s0; free(h); hcap *= 4;
    }
    if (p->chain && nh > 0) {
        uint8_t *in = (uint8_t *)malloc(nh);
        lzo_chain(h, nh, in);
        long k = 0;
        for (long x = 0; x < nh; x++) if (in[x]) h[k++] = h[x];
        nh = k; free(in);
    } else lzo_sort_hsps(h, nh);
    st->chained += nh;
    long nout;
    if (p->gapped) nout = lzo_gapped(t, n, q, m, h, nh, p, out, cap, st);
    else {
        nout = 0;
        for (long x = 0; x < nh; x++) {
            if (nout >= cap) { nout = -1; break; }
            lzo_aln *o = &out[nout++];
            o->s1 = h[x].s1; o->e1 = h[x].s1 + h[x].len; o->s2 = h[x].s2; o->e2 = h[x].s2 + h[x].len; o->score = h[x].score;
            int nm = 0; for (int c = 0; c < h[x].len; c++) nm += (BASE(t[h[x].s1 + c]) == BASE(q[h[x].s2 + c]) && BASE(t[h[x].s1 + c]) < 4);
            o->nmatch = nm; o->ncols = h[x].len; o->a1 = o->s1; o->a2 = o->s2;
        }
    }
    free(h);
    return nout;
}

long lzo_align_tile(const uint8_t *t, long n, const uint8_t *q, long m, const lzo_params *p, lzo_aln *out, long cap,
                    lzo_stats *st)
{
    return lzo_align_tile_ix(NULL, t, n, q, m, p, out, cap, st);
}

/* helpers for the Python side */
void lzo_revcomp(const uint8_t *s, long n, uint8_t *out)
{
    for (long i = 0; i < n; i++) { uint8_t b = s[n - 1 - i]; out[i] = BASE(b) < 4 ? (uint8_t)((3 - BASE(b)) | (b & 8)) : b; }
}
int lzo_sub(int a, int b) { return SUB[BASE(a)][BASE(b)]; }
