/*
 * lastz_faithful.c -- TEST INFRASTRUCTURE ONLY (CPU; never on the product path).
 *
 * A SECOND statement of the LASTZ stage, sequential and order-dependent the way LASTZ's own implementation is, written to
 * MEASURE how far the order-independent spec of lastz_oracle.c (deviations D1-D5, which the CUDA path reproduces bit for
 * bit) moves mimeo's output. It is not a reference for bit-exact parity: LASTZ itself is absent (environment.yml:7, no
 * pin) and this file restates its published behaviour from memory of its documentation and source -- PARITY UNPINNED.
 *
 * Where it differs from lastz_oracle.c (each item undoes one stated deviation):
 *   F1 (D1/D2)  every seed hit is a candidate, visited in LASTZ's order (query position, probe, target position); a hit is
 *               skipped when it starts before the extent reached by the last gap-free extension on its diagonal --
 *               INCLUDING extensions that failed the score threshold (the extent is where the x-drop stopped).
 *   F3 (D3)     --entropy is evaluated in double precision.
 *   F4 (D4)     the gapped extension is evaluated row by row (target rows), the best score is updated cell by cell and the
 *               y-drop test uses the best seen so far at that moment; a row covers the alive columns of the previous row
 *               plus whatever stays alive to the right.
 *   F5 (D5)     an anchor is skipped only if it lies ON the path of a reported alignment (between its first and last column
 *               on the anchor's row); an extension may not cross a reported alignment: per row its columns are clipped at
 *               the paths of the alignments reported before it. Identity comes from the traceback.
 * Shared with lastz_oracle.c (included below): scoring, seeds, the x-drop rule, --chain, anchors, tie-breaks (M > D > I,
 * open beats extend).
 */
#include "lastz_oracle.c"

#include <math.h>

/* ------------------------------------------------------------------ F1 + F3: HSPs */
static double entropy_double(const uint32_t cnt[4])
{
    double n = (double)cnt[0] + cnt[1] + cnt[2] + cnt[3], h = 0.0;
    if (n <= 0) return 0.0;
    for (int b = 0; b < 4; b++)
        if (cnt[b]) { double p = cnt[b] / n; h -= p * log(p) / log(4.0); }
    return h;
}

/* x-drop extension that also reports where the scan stopped (right terminal position), for the diagonal memory */
static void xdrop_extend_stop(const uint8_t *t, long n, const uint8_t *q, long m, long i, long j, int X,
                              long *bstart, long *bend, int *score, long *stop_right, int64_t *cells)
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
    *stop_right = c1;
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

long lzf_hsps(const lzo_index *ix, const uint8_t *t, long n, const uint8_t *q, long m, const lzo_params *p, lzo_hsp *out,
              long cap, lzo_stats *st)
{
    lzo_stats local; memset(&local, 0, sizeof(local)); if (!st) st = &local;
    if (n < SEED_SPAN || m < SEED_SPAN) return 0;
    const uint32_t *off = ix->off, *pos = ix->pos;
    int32_t *diag_end = (int32_t *)malloc((size_t)(n + m + 1) * sizeof(int32_t));
    for (long k = 0; k < n + m + 1; k++) diag_end[k] = -1;
    long nout = 0;
    for (long j = 0; j + SEED_SPAN <= m; j++) {
        if (!window_clean(q, m, j)) continue;
        int key = seed_key(q, j);
        int nprobe = p->transition ? 13 : 1;
        for (int pr = 0; pr < nprobe; pr++) {
            int k = pr == 0 ? key : (key ^ (2 << (2 * (pr - 1))));
            for (uint32_t x = off[k]; x < off[k + 1]; x++) {
                long i = pos[x];
                st->seed_hits++;
                long d = i - j + m;
                if (i < diag_end[d]) continue;                     /* F1: inside the extent of the last extension, kept or not */
                long bs, be, stop; int score;
                st->extended++;
                xdrop_extend_stop(t, n, q, m, i, j, p->xdrop, &bs, &be, &score, &stop, &st->ungapped_cells);
                if (stop > diag_end[d]) diag_end[d] = (int32_t)stop;
                if (score < p->hspthresh) continue;
                st->hsps_raw++;
                if (p->entropy) {
                    uint32_t cnt[4] = {0, 0, 0, 0};
                    for (long c = bs; c < be; c++) if (BASE(t[c]) == BASE(q[c - (i - j)]) && BASE(t[c]) < 4) cnt[BASE(t[c])]++;
                    score = (int)((double)score * entropy_double(cnt));          /* F3 */
                    if (score < p->hspthresh) continue;
                }
                if (nout >= cap) { free(diag_end); return -1; }
                out[nout].s1 = (int32_t)bs; out[nout].s2 = (int32_t)(bs - (i - j));
                out[nout].len = (int32_t)(be - bs); out[nout].score = score; nout++;
                st->hsps_kept++;
            }
        }
    }
    free(diag_end);
    return nout;
}

/* ------------------------------------------------------------------ F4 + F5: gapped extension, row by row, with barriers */
typedef struct {
    int32_t s1, e1;        /* target rows [s1, e1) the alignment touches (0-based) */
    int32_t *jlo, *jhi;    /* per row: first / last query column on the path */
} lzf_path;

typedef struct { int score, di, dj, nmatch, ncols; } fext_t;

/* One-sided extension from the anchor. dir = +1: cell (i,j) consumes t[a1+i-1], q[a2+j-1]; dir = -1: t[a1-i], q[a2-j].
 * lim_lo[i] / lim_hi[i] (i = 0..tn): allowed columns of row i (inclusive), already clipped at earlier alignments; NULL = free.
 * The path is returned as per-row (jmin, jmax) in extension coordinates through pj_lo / pj_hi (caller allocates tn + 1). */
static fext_t rowwise_extend(const uint8_t *t, long a1, long tn, const uint8_t *q, long a2, long qn, int dir, const lzo_params *p,
                             const int32_t *lim_lo, const int32_t *lim_hi, int32_t *pj_lo, int32_t *pj_hi, int64_t *cells)
{
    const int O = p->gap_open, E = p->gap_extend, Y = p->ydrop;
    fext_t r = {0, 0, 0, 0, 0};
    /* rows kept for the traceback: for every row its column range and 2+1+1 direction bits per cell */
    long rows_cap = 1024, nrows = 0;
    long *row_lo = (long *)malloc(rows_cap * sizeof(long)), *row_n = (long *)malloc(rows_cap * sizeof(long));
    uint8_t **row_dir = (uint8_t **)malloc(rows_cap * sizeof(uint8_t *));
    long wcap = 2048;
    int *Hp = (int *)malloc(wcap * sizeof(int)), *Dp = (int *)malloc(wcap * sizeof(int));
    int *Hc = (int *)malloc(wcap * sizeof(int)), *Dc = (int *)malloc(wcap * sizeof(int));
    uint8_t *dbuf = (uint8_t *)malloc(wcap);
    long plo = 0, phi = -1;             /* previous row: alive columns [plo, phi], arrays indexed by j - plo */
    int best = 0;
    for (long i = 0; i <= tn; i++) {
        /* reachable cells of the row: below or diagonally below an alive cell (columns plo .. phi + 1), then rightwards as long
         * as the horizontal gap state stays alive; row 0 starts from the anchor cell alone */
        long lo = i == 0 ? 0 : plo, hi_seed = i == 0 ? 0 : phi + 1;
        long llo = lim_lo ? lim_lo[i] : 0, lhi = lim_hi ? lim_hi[i] : qn;
        if (lhi > qn) lhi = qn;
        if (lo < llo) lo = llo;
        if (lo > lhi || lo > hi_seed) break;
        int icur = NEG_INF, hleft = NEG_INF;
        long first = -1, last = -1, j;
        for (j = lo; j <= lhi; j++) {
            if (j > hi_seed && hleft <= NEG_INF) break;               /* nothing can reach further right */
            if (j - lo + 2 > wcap) {
                wcap *= 2;
                Hp = (int *)realloc(Hp, wcap * sizeof(int)); Dp = (int *)realloc(Dp, wcap * sizeof(int));
                Hc = (int *)realloc(Hc, wcap * sizeof(int)); Dc = (int *)realloc(Dc, wcap * sizeof(int));
                dbuf = (uint8_t *)realloc(dbuf, wcap);
            }
            int hu = NEG_INF, du = NEG_INF, hd = NEG_INF;
            if (i > 0 && j >= plo && j <= phi) { hu = Hp[j - plo]; du = Dp[j - plo]; }
            if (i > 0 && j - 1 >= plo && j - 1 <= phi) hd = Hp[j - 1 - plo];
            int d = NEG_INF, dbit = 0;
            if (hu > NEG_INF) { int open = hu - O - E, ext = du > NEG_INF ? du - E : NEG_INF; if (open >= ext) d = open; else { d = ext; dbit = 1; } }
            int ii = NEG_INF, ibit = 0;
            if (hleft > NEG_INF) { int open = hleft - O - E, ext = icur > NEG_INF ? icur - E : NEG_INF; if (open >= ext) ii = open; else { ii = ext; ibit = 1; } }
            int mval = NEG_INF;
            if (i == 0 && j == 0) mval = 0;
            else if (hd > NEG_INF && i >= 1 && j >= 1) {
                int a = BASE(dir > 0 ? t[a1 + i - 1] : t[a1 - i]), b = BASE(dir > 0 ? q[a2 + j - 1] : q[a2 - j]);
                mval = hd + SUB[a][b];
            }
            int h, hsel;
            if (mval >= d && mval >= ii) { h = mval; hsel = 0; } else if (d >= ii) { h = d; hsel = 1; } else { h = ii; hsel = 2; }
            (*cells)++;
            if (h <= NEG_INF || h < best - Y) { h = NEG_INF; d = NEG_INF; ii = NEG_INF; }   /* F4: against the best seen SO FAR */
            else {
                if (first < 0) first = j;
                last = j;
                if (h > best) { best = h; r.score = h; r.di = (int)i; r.dj = (int)j; }
            }
            Hc[j - lo] = h; Dc[j - lo] = d;
            dbuf[j - lo] = (uint8_t)(hsel | (dbit << 2) | (ibit << 3));
            hleft = h; icur = ii;
        }
        if (first < 0) break;
        if (nrows >= rows_cap) {
            rows_cap *= 2;
            row_lo = (long *)realloc(row_lo, rows_cap * sizeof(long)); row_n = (long *)realloc(row_n, rows_cap * sizeof(long));
            row_dir = (uint8_t **)realloc(row_dir, rows_cap * sizeof(uint8_t *));
        }
        uint8_t *dirs = (uint8_t *)malloc((size_t)(last - first + 1));
        memcpy(dirs, dbuf + (first - lo), (size_t)(last - first + 1));
        row_lo[nrows] = first; row_n[nrows] = last - first + 1; row_dir[nrows] = dirs; nrows++;
        /* next row sees only the alive part */
        memmove(Hc, Hc + (first - lo), (size_t)(last - first + 1) * sizeof(int));
        memmove(Dc, Dc + (first - lo), (size_t)(last - first + 1) * sizeof(int));
        int *tmp = Hp; Hp = Hc; Hc = tmp; tmp = Dp; Dp = Dc; Dc = tmp;
        plo = first; phi = last;
    }
    free(dbuf);
    /* traceback from the best cell: matches, aligned columns, per-row path extent */
    for (long i = 0; i <= tn && pj_lo; i++) { pj_lo[i] = INT32_MAX; pj_hi[i] = -1; }
    {
        long i = r.di, j = r.dj;
        int state = 0;                       /* 0 = H, 1 = D, 2 = I */
        while (i > 0 || j > 0) {
            if (i >= nrows) break;
            if (pj_lo) { if (j < pj_lo[i]) pj_lo[i] = (int32_t)j; if (j > pj_hi[i]) pj_hi[i] = (int32_t)j; }
            uint8_t b = row_dir[i][j - row_lo[i]];
            if (state == 0) {
                int hs = b & 3;
                if (hs == 0) {
                    int a = BASE(dir > 0 ? t[a1 + i - 1] : t[a1 - i]), c = BASE(dir > 0 ? q[a2 + j - 1] : q[a2 - j]);
                    r.ncols++; if (a == c && a < 4) r.nmatch++;
                    i--; j--;
                } else state = hs;
            } else if (state == 1) { state = (b >> 2) & 1 ? 1 : 0; i--; }
            else { state = (b >> 3) & 1 ? 2 : 0; j--; }
        }
        if (pj_lo) { if (0 < pj_lo[0]) pj_lo[0] = 0; if (0 > pj_hi[0]) pj_hi[0] = 0; }
    }
    for (long k = 0; k < nrows; k++) free(row_dir[k]);
    free(row_dir); free(row_lo); free(row_n); free(Hp); free(Dp); free(Hc); free(Dc);
    return r;
}

/* Gapped stage of one tile-strand, sequential: anchors best-first, path-based cover test, barriers at reported alignments. */
long lzf_gapped(const uint8_t *t, long n, const uint8_t *q, long m, const lzo_hsp *h, long nh, const lzo_params *p,
                lzo_aln *out, long cap, lzo_stats *st)
{
    lzo_stats local; memset(&local, 0, sizeof(local)); if (!st) st = &local;
    long *ord = (long *)malloc((nh ? nh : 1) * sizeof(long));
    for (long k = 0; k < nh; k++) ord[k] = k;
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
    lzf_path *paths = (lzf_path *)calloc((size_t)(nh ? nh : 1), sizeof(lzf_path));
    long nout = 0;
    for (long a = 0; a < nh; a++) {
        const lzo_hsp *hh = &h[ord[a]];
        int32_t a1, a2;
        lzo_anchor(t, q, hh, &a1, &a2);
        int covered = 0;
        for (long k = 0; k < nout && !covered; k++)
            if (a1 >= paths[k].s1 && a1 < paths[k].e1 && a2 >= paths[k].jlo[a1 - paths[k].s1] && a2 <= paths[k].jhi[a1 - paths[k].s1]) covered = 1;   /* F5 */
        if (covered) continue;
        st->anchors_extended++;
        fext_t ext[2];
        int32_t *plo[2], *phi[2];
        for (int side = 0; side < 2; side++) {
            const int dir = side == 0 ? +1 : -1;
            const long tn = dir > 0 ? n - a1 : a1, qn = dir > 0 ? m - a2 : a2;
            int32_t *lim_lo = (int32_t *)malloc((size_t)(tn + 1) * sizeof(int32_t)), *lim_hi = (int32_t *)malloc((size_t)(tn + 1) * sizeof(int32_t));
            for (long i = 0; i <= tn; i++) { lim_lo[i] = 0; lim_hi[i] = (int32_t)qn; }
            for (long k = 0; k < nout; k++) {            /* F5: do not cross alignments reported before */
                const lzf_path *P = &paths[k];
                long rr = a1 < P->s1 ? P->s1 : (a1 >= P->e1 ? P->e1 - 1 : a1);
                long pcol = ((long)P->jlo[rr - P->s1] + P->jhi[rr - P->s1]) / 2, mycol = a2 + (rr - a1);
                int right = pcol > mycol;               /* the old alignment runs to the right of this anchor's diagonal */
                for (long i = 0; i <= tn; i++) {
                    long row = dir > 0 ? a1 + i - 1 : a1 - i;       /* target base consumed on extension row i (row 0: none) */
                    if (i == 0) continue;
                    if (row < P->s1 || row >= P->e1) continue;
                    long lo = P->jlo[row - P->s1], hi = P->jhi[row - P->s1];       /* absolute query columns */
                    /* extension column j consumes query base a2 + j - 1 (dir +1) or a2 - j (dir -1) */
                    if (dir > 0) {
                        if (right) { long lim = lo - a2; if (lim < lim_hi[i]) lim_hi[i] = (int32_t)(lim < 0 ? -1 : lim); }
                        else { long lim = hi - a2 + 2; if (lim > lim_lo[i]) lim_lo[i] = (int32_t)lim; }
                    } else {
                        if (right) { long lim = a2 - lo + 1; if (lim > lim_lo[i]) lim_lo[i] = (int32_t)lim; }
                        else { long lim = a2 - hi - 1; if (lim < lim_hi[i]) lim_hi[i] = (int32_t)(lim < 0 ? -1 : lim); }
                    }
                }
            }
            plo[side] = (int32_t *)malloc((size_t)(tn + 1) * sizeof(int32_t)); phi[side] = (int32_t *)malloc((size_t)(tn + 1) * sizeof(int32_t));
            ext[side] = rowwise_extend(t, a1, tn, q, a2, qn, dir, p, lim_lo, lim_hi, plo[side], phi[side], &st->gapped_cells);
            free(lim_lo); free(lim_hi);
        }
        int score = ext[0].score + ext[1].score;
        if (score >= p->gappedthresh && nout < cap) {
            lzo_aln *o = &out[nout];
            o->s1 = a1 - ext[1].di; o->e1 = a1 + ext[0].di; o->s2 = a2 - ext[1].dj; o->e2 = a2 + ext[0].dj;
            o->score = score; o->nmatch = ext[0].nmatch + ext[1].nmatch; o->ncols = ext[0].ncols + ext[1].ncols; o->a1 = a1; o->a2 = a2;
            /* absolute path: row r of the target -> query columns */
            lzf_path *P = &paths[nout];
            P->s1 = o->s1; P->e1 = o->e1 > o->s1 ? o->e1 : o->s1 + 1;
            long nr = P->e1 - P->s1;
            P->jlo = (int32_t *)malloc((size_t)nr * sizeof(int32_t)); P->jhi = (int32_t *)malloc((size_t)nr * sizeof(int32_t));
            for (long k = 0; k < nr; k++) { P->jlo[k] = INT32_MAX; P->jhi[k] = -1; }
            for (long i = 1; i <= ext[0].di; i++) {          /* forward rows: target base a1 + i - 1, query bases a2 + j - 1 */
                long row = a1 + i - 1 - P->s1;
                if (plo[0][i] <= phi[0][i]) { long lo = a2 + (plo[0][i] > 0 ? plo[0][i] - 1 : 0), hi = a2 + (phi[0][i] > 0 ? phi[0][i] - 1 : 0);
                    if (lo < P->jlo[row]) P->jlo[row] = (int32_t)lo; if (hi > P->jhi[row]) P->jhi[row] = (int32_t)hi; }
            }
            for (long i = 1; i <= ext[1].di; i++) {          /* backward rows: target base a1 - i, query bases a2 - j */
                long row = a1 - i - P->s1;
                if (plo[1][i] <= phi[1][i]) { long hi = a2 - (plo[1][i] > 0 ? plo[1][i] : 1), lo = a2 - (phi[1][i] > 0 ? phi[1][i] : 1);
                    if (lo < P->jlo[row]) P->jlo[row] = (int32_t)lo; if (hi > P->jhi[row]) P->jhi[row] = (int32_t)hi; }
            }
            /* rows the path only passes vertically inherit the neighbouring column */
            int32_t lastc = (int32_t)o->s2;
            for (long k = 0; k < nr; k++) { if (P->jhi[k] < 0) { P->jlo[k] = lastc; P->jhi[k] = lastc; } else lastc = P->jhi[k]; }
            nout++;
        }
        for (int side = 0; side < 2; side++) { free(plo[side]); free(phi[side]); }
    }
    for (long k = 0; k < nout; k++) { free(paths[k].jlo); free(paths[k].jhi); }
    free(paths); free(ord);
    return nout;
}

/* Whole faithful pipeline for one (target, query-strand) tile. */
long lzf_align_tile_ix(const lzo_index *ix, const uint8_t *t, long n, const uint8_t *q, long m, const lzo_params *p,
                       lzo_aln *out, long cap, lzo_stats *st)
{
    lzo_stats local; memset(&local, 0, sizeof(local)); if (!st) st = &local;
    long hcap = 1 << 16, nh;
    lzo_hsp *h = NULL;
    for (;;) {
        h = (lzo_hsp *)malloc(hcap * sizeof(lzo_hsp));
        lzo_stats s0 = *st;
        nh = lzf_hsps(ix, t, n, q, m, p, h, hcap, st);
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
    long nout = lzf_gapped(t, n, q, m, h, nh, p, out, cap, st);
    free(h);
    return nout;
}
