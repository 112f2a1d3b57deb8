// hsp.cu -- kernel family (b): warp-cooperative gap-free x-drop extension of the surviving seed hits
// into HSPs, scored against --hspthresh with LASTZ's --entropy adjustment.
//
// Replaces LASTZ's --gfextend stage under the order-independent spec of
// oracle/lastz_oracle.c (D1-D3):
//   1. survivors (diagonal, query position) are radix-sorted, which groups them by diagonal in
//      increasing position;
//   2. one warp walks one diagonal: a leader whose 19-mer ends inside the last KEPT HSP of the
//      diagonal is skipped, otherwise the warp extends it 32 columns per step (prefix-sum and
//      prefix-max by shuffles, exact x-drop termination by ballot);
//   3. HSPs with score >= K are entropy-adjusted in integer fixed point and kept if still >= K.
#include "primitives.cuh"
#include "seq.cuh"
#include <cooperative_groups.h>

#include "internal.cuh"
#include "xdrop_table.cuh"

namespace mb2 {

__device__ __forceinline__ int sub_lut2(uint32_t idx) {
    const uint64_t lo = 0xE183648E85E18E5Bull, hi = 0x5B8EE1858E6483E1ull;
    const uint64_t v = (idx & 8) ? hi : lo;
    return (int)(int8_t)(v >> ((idx & 7) * 8));
}

__device__ __forceinline__ int warp_incl_sum_i(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
    return v;
}
__device__ __forceinline__ int warp_incl_max_i(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v = max(v, t); }
    return v;
}

__device__ __forceinline__ int col_score(const GenomeView& T, const GenomeView& Q, uint32_t ct, uint32_t cq) {
    const uint32_t an = isn_at(T.nm, ct) | isn_at(Q.nm, cq);
    return an ? SCORE_N : sub_lut2((base_at(T.pk, ct) << 2) | base_at(Q.pk, cq));
}

// Exact x-drop extension, 1024 columns per warp step: lane l scores its own block of 32 consecutive columns from
// registers (two unaligned 64-bit windows + N masks), then two warp scans stitch the blocks together.
//   phase 1 (per lane): block sum, block prefix maximum (+ its first position), the deepest drop below the block's own
//            running maximum and the lowest prefix value;
//   phase 2 (warp):     S_in = exclusive sum, M_in = exclusive running best; a block can contain the termination column
//            iff  min_prefix < (M_in - S_in) - X  or  deepest_drop < -X; the first such lane replays its 32 columns with
//            the true incoming state to find the exact column.
// DIR=+1: columns ct0, ct0+1, ... -> best score and exclusive end of the best prefix.
// DIR=-1: columns ct0, ct0-1, ... -> best score and inclusive start of the best prefix.
template <int DIR>
__device__ __forceinline__ void xdrop_side(const GenomeView& T, const GenomeView& Q, uint32_t ct0, uint32_t cq0, int X, int lane,
                                           int& best_out, uint32_t& bpos_out, unsigned long long& cells) {
    int run0 = 0, best = 0;
    uint32_t bpos = DIR > 0 ? ct0 : ct0 + 1;
    for (uint32_t base = 0;; base += 1024) {
        // block of this lane: DIR>0 columns [ct0+base+32l, +32) ascending; DIR<0 columns (ct0-base-32l) descending
        const uint32_t wt_pos = DIR > 0 ? ct0 + base + 32u * lane : ct0 - base - 32u * lane - 31u;
        const uint32_t wq_pos = DIR > 0 ? cq0 + base + 32u * lane : cq0 - base - 32u * lane - 31u;
        const uint64_t wt = window32(T.pk, wt_pos), wq = window32(Q.pk, wq_pos);
        const uint32_t an = nwindow32(T.nm, wt_pos) | nwindow32(Q.nm, wq_pos);
        int sc[32];
        int p = 0, lmax = INT_MIN, lmax_at = 0, mind = INT_MAX, minp = INT_MAX;
#pragma unroll
        for (int c = 0; c < 32; c++) {
            const int k = DIR > 0 ? c : 31 - c;                       // bit position inside the windows
            const int s = ((an >> k) & 1u) ? SCORE_N : sub_lut2((uint32_t)(((wt >> (2 * k)) & 3) << 2 | ((wq >> (2 * k)) & 3)));
            sc[c] = s;
            p += s;
            if (c > 0) mind = min(mind, p - lmax);
            minp = min(minp, p);
            if (p > lmax) { lmax = p; lmax_at = c; }
        }
        const int sum_incl = warp_incl_sum_i(p, lane);
        const int s_in = sum_incl - p + run0;                          // running sum entering this lane's block
        const int cand = s_in + lmax;                                  // best running value reached inside the block
        const int cm = warp_incl_max_i(cand, lane);
        int m_in = __shfl_up_sync(0xffffffffu, cm, 1);
        if (lane == 0) m_in = INT_MIN;
        m_in = max(m_in, best);                                        // best seen before this block
        const int dgap = m_in - s_in;
        const bool may_term = (minp < dgap - X) || (mind < -X);
        const uint32_t tmask = __ballot_sync(0xffffffffu, may_term);
        const int lt = tmask ? __ffs(tmask) - 1 : 32;
        // the terminating lane replays its block with the true incoming state
        int tcol = 32, tbest = INT_MIN, tbest_at = 0;
        if (lane == lt) {
            int q = 0, lm = INT_MIN;
#pragma unroll
            for (int c = 0; c < 32; c++) {
                q += sc[c];
                const int ref = max(dgap, lm);
                if (tcol == 32 && q < ref - X) tcol = c;
                if (tcol == 32 && q > lm) { lm = q; if (q > tbest) { tbest = q; tbest_at = c; } }
            }
        }
        // candidates for a new best: whole blocks of lanes < lt, and the prefix of lane lt before its termination column
        int v = INT_MIN, vat = 0;
        if (lane < lt) { v = cand; vat = lmax_at; }
        else if (lane == lt && tbest != INT_MIN) { v = s_in + tbest; vat = tbest_at; }
        const int mx = __reduce_max_sync(0xffffffffu, v);
        if (mx > best) {
            const int src = __ffs(__ballot_sync(0xffffffffu, v == mx)) - 1;
            const int at = __shfl_sync(0xffffffffu, vat, src);
            const uint32_t col = base + 32u * src + at;                // columns consumed before + this one
            bpos = DIR > 0 ? ct0 + col + 1 : ct0 - col;
            best = mx;
        }
        if (tmask) {
            const int tc = __shfl_sync(0xffffffffu, tcol, lt);
            cells += 32ull * lt + tc + 1;
            if (tc < 32) break;
            // the flagged lane did not actually terminate (conservative test): continue with the lanes after it
            // by restarting the step right after that lane's block
            run0 = __shfl_sync(0xffffffffu, s_in + p, lt);
            // shift so that the next step starts at the block following lane lt
            base += 32u * (lt + 1) - 1024u;
            continue;
        }
        cells += 1024;
        run0 = __shfl_sync(0xffffffffu, sum_incl, 31) + run0;
    }
    best_out = best; bpos_out = bpos;
}

__device__ __forceinline__ uint32_t log2_q24(uint32_t x) {   // x >= 1; same integer algorithm as the oracle (spec D3)
    const int ip = 31 - __clz(x);
    uint64_t y = (uint64_t)x << (31 - ip);
    uint32_t frac = 0;
#pragma unroll 1
    for (int k = 0; k < 24; k++) {
        y = (y * y) >> 31;
        frac <<= 1;
        if (y >= (1ull << 32)) { y >>= 1; frac |= 1; }
    }
    return ((uint32_t)ip << 24) | frac;
}

__device__ __forceinline__ uint32_t entropy_q24(const uint32_t cnt[4]) {
    const uint32_t n = cnt[0] + cnt[1] + cnt[2] + cnt[3];
    if (n == 0) return 0;
    const uint32_t ln = log2_q24(n);
    uint64_t Tt = 0;
#pragma unroll
    for (int b = 0; b < 4; b++)
        if (cnt[b]) Tt += (uint64_t)cnt[b] * (uint64_t)(ln - log2_q24(cnt[b]));
    return (uint32_t)(Tt / (2ull * n));
}

__device__ __forceinline__ int scaf_of(const uint32_t* __restrict__ off, int n, uint32_t p) {
    int lo = 0, hi = n;
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid] <= p) lo = mid; else hi = mid; }
    return lo;
}

// A segment = consecutive survivors on one global diagonal inside one target scaffold. (Global diagonals are shared by
// different scaffold pairs, e.g. every scaffold's trivial self-alignment lies on diagonal 0; splitting at scaffold
// boundaries is exact because an HSP never crosses the pad between scaffolds, and it lets those walks run in parallel.)
__global__ void __launch_bounds__(256)
diag_heads_kernel(const uint64_t* __restrict__ surv, uint32_t n, GenomeView T, uint32_t diag_bias, uint32_t* __restrict__ flag) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t head = 1;
    if (k > 0) {
        const uint32_t dg = (uint32_t)(surv[k] >> 32), dgp = (uint32_t)(surv[k - 1] >> 32);
        if (dg == dgp) {
            const uint32_t i = (uint32_t)surv[k] + dg - diag_bias, ip = (uint32_t)surv[k - 1] + dg - diag_bias;
            head = scaf_of(T.off, T.nscaf, i) != scaf_of(T.off, T.nscaf, ip) ? 1u : 0u;
        }
    }
    flag[k] = head;
}
__global__ void __launch_bounds__(256)
diag_starts_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ flag_off, uint32_t n, uint32_t* __restrict__ seg_start) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (flag[k]) seg_start[flag_off[k]] = k;
}

constexpr uint32_t HSP_CLOSED_ITEM = 0xFFFFFFFFu;   // = HSP_CLOSED below
// Warp-per-diagonal walk (1024 columns per warp step): takes the diagonals the thread-per-diagonal kernel handed over
// because one of their extensions outgrew a single thread. Item k = {segment, first survivor to process, covered}.
__global__ void __launch_bounds__(128)
hsp_extend_kernel(GenomeView T, GenomeView Q, const uint64_t* __restrict__ surv, uint32_t nsurv,
                  const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ nseg_p,
                  const uint32_t* __restrict__ items, const uint32_t* __restrict__ nitems_p, uint32_t diag_bias,
                  int X, int K, int entropy, uint32_t cap,
                  uint32_t* __restrict__ o_tile, int32_t* __restrict__ o_s1, int32_t* __restrict__ o_s2,
                  int32_t* __restrict__ o_len, int32_t* __restrict__ o_score, unsigned long long* __restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const uint32_t nseg = *nseg_p, nitems = *nitems_p;
    unsigned long long cells = 0, extended = 0;
    for (;;) {
        uint32_t it = 0;
        if (lane == 0) it = (uint32_t)atomicAdd(&counters[CNT_WORK], 1ull);
        it = __shfl_sync(0xffffffffu, it, 0);
        if (it >= nitems) break;
        const uint32_t seg = items[3 * it];
        const uint32_t a = seg_start[seg];
        const uint32_t b = (seg + 1 < nseg) ? seg_start[seg + 1] : nsurv;
        const uint32_t dg = (uint32_t)(surv[a] >> 32);          // i - j + bias
        // Trivial self-diagonal (an N-free scaffold against itself, flagged by the thread kernel): every column is a match
        // scoring s(b,b) > 0, so the extension from any seed covers the whole scaffold and stops in the pads:
        // score = 91 * #(A,T) + 100 * #(C,G), no walk needed.
        const bool closed = items[3 * it + 2] == HSP_CLOSED_ITEM;
        uint32_t covered = closed ? 0u : items[3 * it + 2];      // exclusive end (target coords) of the last kept HSP
        for (uint32_t x = items[3 * it + 1]; x < b; x++) {
            const uint32_t j = (uint32_t)surv[x];
            const uint32_t i = j + dg - diag_bias;
            if (i + SEED_SPAN <= covered) continue;               // spec D2
            extended++;
            int best_r = 0, best_l = 0; uint32_t be, bs;
            if (closed) {
                const int ts = scaf_of(T.off, T.nscaf, i);
                bs = T.off[ts]; be = bs + T.len[ts];
                int cgn = 0;
                const uint32_t w0 = bs >> 5, w1 = (be - 1) >> 5;
#pragma unroll 8
                for (uint32_t w = w0 + lane; w <= w1; w += 32) {       // independent loads: unrolled so that several are in flight
                    const uint64_t xw = T.pk[w];
                    uint64_t m = (xw ^ (xw >> 1)) & 0x5555555555555555ull;       // one bit per C or G base
                    if (w == w0 && (bs & 31)) m &= ~0ull << (2 * (bs & 31));
                    if (w == w1 && (be & 31)) m &= ~0ull >> (64 - 2 * (be & 31));
                    cgn += __popcll(m);
                }
                cgn = __reduce_add_sync(0xffffffffu, cgn);
                best_r = 91 * (int)(be - bs - (uint32_t)cgn) + 100 * cgn;
                cells += be - bs;
            } else {
                xdrop_side<+1>(T, Q, i + SEED_SPAN, j + SEED_SPAN, X, lane, best_r, be, cells);
                xdrop_side<-1>(T, Q, i + SEED_SPAN - 1, j + SEED_SPAN - 1, X, lane, best_l, bs, cells);
            }
            int score = best_r + best_l;
            if (score < K) continue;
            const uint32_t qs = bs - (i - j);
            if (entropy) {
                uint32_t cnt[4] = {0, 0, 0, 0};
#pragma unroll 4
                for (uint32_t c = bs + 32u * lane; c < be; c += 1024u) {      // each lane owns 32-column words
                    const uint64_t wt = window32(T.pk, c), wq = window32(Q.pk, qs + (c - bs));
                    const uint32_t an = nwindow32(T.nm, c) | nwindow32(Q.nm, qs + (c - bs));
                    const uint64_t x = wt ^ wq;
                    uint64_t m = ~(x | (x >> 1)) & 0x5555555555555555ull;      // bit 2k set iff column k matches
                    // drop N columns and columns beyond the HSP end
                    if (an) {
                        uint64_t nsp = 0;
#pragma unroll
                        for (int k = 0; k < 32; k++) nsp |= (uint64_t)((an >> k) & 1u) << (2 * k);
                        m &= ~nsp;
                    }
                    const uint32_t left = be - c;
                    if (left < 32) m &= (~0ull) >> (64 - 2 * left);
                    const uint64_t lo = wt & 0x5555555555555555ull, hi = (wt >> 1) & 0x5555555555555555ull;
                    cnt[0] += __popcll(m & ~lo & ~hi); cnt[1] += __popcll(m & lo & ~hi);
                    cnt[2] += __popcll(m & ~lo & hi);  cnt[3] += __popcll(m & lo & hi);
                }
#pragma unroll
                for (int bb = 0; bb < 4; bb++) cnt[bb] = __reduce_add_sync(0xffffffffu, cnt[bb]);
                const uint32_t h = entropy_q24(cnt);
                score = (int)(((long long)score * (long long)h) >> 24);
                if (score < K) continue;
            }
            covered = be;
            if (lane == 0) {
                const unsigned long long slot = atomicAdd(&counters[CNT_HSPS], 1ull);
                if (slot < cap) {
                    const int ts = scaf_of(T.off, T.nscaf, bs), qsf = scaf_of(Q.off, Q.nscaf, qs);
                    o_tile[slot] = (uint32_t)ts * (uint32_t)Q.nscaf + (uint32_t)qsf;
                    o_s1[slot] = (int32_t)(bs - T.off[ts]);
                    o_s2[slot] = (int32_t)(qs - Q.off[qsf]);
                    o_len[slot] = (int32_t)(be - bs);
                    o_score[slot] = score;
                }
            }
        }
    }
    if (lane == 0) {
        if (cells) atomicAdd(&counters[CNT_S2_CELLS], cells);
        if (extended) atomicAdd(&counters[CNT_EXTENDED], extended);
    }
}


// ---- thread-per-diagonal walk ---------------------------------------------------------------------------------
// Most extensions are a few hundred columns long, far too short to feed a whole warp. Here every THREAD walks one
// diagonal segment with the exact x-drop at three columns per table lookup (xdrop_table.cuh). An extension that is
// still alive after HT_LIMIT windows of 30 columns on one side hands the rest of its diagonal (from that survivor on,
// with the current `covered`) to the warp kernel above.
constexpr int HT_LIMIT = 16;
constexpr uint32_t HSP_CLOSED = 0xFFFFFFFFu;     // `covered` value of a handed-over item that is a trivial self-diagonal

// one side, one thread. DIR=+1: columns ct0, ct0+1, ...; DIR=-1: ct0, ct0-1, ... Returns false if the limit was reached.
// bcol = number of columns of the best prefix (0 = empty prefix).
template <int DIR>
__device__ __forceinline__ bool xdrop_thread(const uint32_t* __restrict__ tab, const GenomeView& T, const GenomeView& Q, uint32_t ct0, uint32_t cq0,
                                             int X, int& best, uint32_t& bcol, unsigned long long& cells) {
    best = 0; bcol = 0;
    int D = 375 - X;
    const int c2 = 250 - X;
    for (int w = 0; w < HT_LIMIT; w++) {
        const uint32_t ct = DIR > 0 ? ct0 + 30u * w : ct0 - 30u * w - 31u;
        const uint32_t cq = DIR > 0 ? cq0 + 30u * w : cq0 - 30u * w - 31u;
        uint64_t wt = window32(T.pk, ct), wq = window32(Q.pk, cq);
        uint32_t an = nwindow32(T.nm, ct) | nwindow32(Q.nm, cq);
        if (DIR < 0) { wt = rev2groups(wt); wq = rev2groups(wq); an = __brev(an); }
        an &= (1u << 30) - 1u;
        if (an) {                                             // a non-ACGT column in the window: column by column
            int run = best - (D - 375 + X);
            for (int c = 0; c < 30; c++) {
                const int sc = (an & 1u) ? SCORE_N : sub_lut((uint32_t)((wt & 3) << 2 | (wq & 3)));
                wt >>= 2; wq >>= 2; an >>= 1;
                run += sc; cells++;
                if (run > best) { best = run; bcol = 30u * w + c + 1; }
                else if (run < best - X) return true;
            }
            D = (best - run) + 375 - X;
            continue;
        }
        const uint32_t tl = (uint32_t)wt, th = (uint32_t)(wt >> 32), ql = (uint32_t)wq, qh = (uint32_t)(wq >> 32);
#pragma unroll
        for (int k = 0; k < 10; k++) {
            const uint32_t e = tab[xt_index(tl, th, ql, qh, k)];
            cells += 3;
            if (xt_minf(e) < D) return true;
            const int dm = max(D, xt_maxf(e) + c2);
            if (dm > D) { best += dm - D; bcol = 30u * w + 3u * k + (uint32_t)xt_argmax(e) + 1u; }
            D = dm - xt_sum(e);
        }
    }
    return false;
}

__global__ void __launch_bounds__(128)
hsp_extend_thread_kernel(GenomeView T, GenomeView Q, const uint64_t* __restrict__ surv, uint32_t nsurv,
                         const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ nseg_p, uint32_t diag_bias,
                         int X, int K, int entropy, uint32_t cap,
                         uint32_t* __restrict__ o_tile, int32_t* __restrict__ o_s1, int32_t* __restrict__ o_s2,
                         int32_t* __restrict__ o_len, int32_t* __restrict__ o_score,
                         uint32_t* __restrict__ items, uint32_t* __restrict__ nitems_p, const int32_t* __restrict__ same_q,
                         unsigned long long* __restrict__ counters) {
    namespace cg = cooperative_groups;
    __shared__ uint32_t tab[XT_SIZE];
    for (int e = threadIdx.x; e < XT_SIZE; e += blockDim.x) tab[e] = xt_entry((uint32_t)e);
    __syncthreads();
    const uint32_t nseg = *nseg_p;
    unsigned long long cells = 0, extended = 0;
    for (;;) {
        uint32_t seg;
        {   // the threads that need work right now fetch it with one atomic
            cg::coalesced_group g = cg::coalesced_threads();
            unsigned long long base = 0;
            if (g.thread_rank() == 0) base = atomicAdd(&counters[CNT_WORK], (unsigned long long)g.size());
            seg = (uint32_t)g.shfl(base, 0) + g.thread_rank();
        }
        if (seg >= nseg) break;
        const uint32_t a = seg_start[seg];
        const uint32_t b = (seg + 1 < nseg) ? seg_start[seg + 1] : nsurv;
        const uint32_t dg = (uint32_t)(surv[a] >> 32);          // i - j + bias
        uint32_t covered = 0;                                     // exclusive end (target coords) of the last kept HSP
        if (same_q) {
            // main diagonal of an N-free scaffold against itself: the whole scaffold is one HSP (closed form in the warp kernel)
            const uint32_t j0 = (uint32_t)surv[a], i0 = j0 + dg - diag_bias;
            const int ts = scaf_of(T.off, T.nscaf, i0), qs = scaf_of(Q.off, Q.nscaf, j0);
            if (same_q[ts] == qs && i0 - T.off[ts] == j0 - Q.off[qs] && T.nfree[ts]) {
                const uint32_t slot = atomicAdd(nitems_p, 1u);
                items[3 * slot] = seg; items[3 * slot + 1] = a; items[3 * slot + 2] = HSP_CLOSED;
                continue;
            }
        }
        for (uint32_t x = a; x < b; x++) {
            const uint32_t j = (uint32_t)surv[x];
            const uint32_t i = j + dg - diag_bias;
            if (i + SEED_SPAN <= covered) continue;               // spec D2
            int best_r, best_l; uint32_t cr, cl;
            const unsigned long long cells0 = cells;
            const bool done = xdrop_thread<+1>(tab, T, Q, i + SEED_SPAN, j + SEED_SPAN, X, best_r, cr, cells) &&
                              xdrop_thread<-1>(tab, T, Q, i + SEED_SPAN - 1, j + SEED_SPAN - 1, X, best_l, cl, cells);
            if (!done) {                                          // too long for one thread: the warp kernel redoes it and finishes the diagonal
                cells = cells0;
                const uint32_t slot = atomicAdd(nitems_p, 1u);
                items[3 * slot] = seg; items[3 * slot + 1] = x; items[3 * slot + 2] = covered;
                break;
            }
            extended++;
            int score = best_r + best_l;
            if (score < K) continue;
            const uint32_t be = i + SEED_SPAN + cr, bs = i + SEED_SPAN - cl;     // [bs, be) in target coordinates
            const uint32_t qs = bs - (i - j);
            if (entropy) {
                uint32_t cnt[4] = {0, 0, 0, 0};
                for (uint32_t c = bs; c < be; c += 32u) {
                    const uint64_t wt = window32(T.pk, c), wq = window32(Q.pk, qs + (c - bs));
                    const uint32_t an = nwindow32(T.nm, c) | nwindow32(Q.nm, qs + (c - bs));
                    const uint64_t xr = wt ^ wq;
                    uint64_t m = ~(xr | (xr >> 1)) & 0x5555555555555555ull;      // bit 2k set iff column k matches
                    if (an) {
                        uint64_t nsp = 0;
                        for (int k = 0; k < 32; k++) nsp |= (uint64_t)((an >> k) & 1u) << (2 * k);
                        m &= ~nsp;
                    }
                    const uint32_t left = be - c;
                    if (left < 32) m &= (~0ull) >> (64 - 2 * left);
                    const uint64_t lo = wt & 0x5555555555555555ull, hi = (wt >> 1) & 0x5555555555555555ull;
                    cnt[0] += __popcll(m & ~lo & ~hi); cnt[1] += __popcll(m & lo & ~hi);
                    cnt[2] += __popcll(m & ~lo & hi);  cnt[3] += __popcll(m & lo & hi);
                }
                const uint32_t h = entropy_q24(cnt);
                score = (int)(((long long)score * (long long)h) >> 24);
                if (score < K) continue;
            }
            covered = be;
            unsigned long long slot;
            {
                cg::coalesced_group g = cg::coalesced_threads();
                unsigned long long base = 0;
                if (g.thread_rank() == 0) base = atomicAdd(&counters[CNT_HSPS], (unsigned long long)g.size());
                slot = g.shfl(base, 0) + g.thread_rank();
            }
            if (slot < cap) {
                const int ts = scaf_of(T.off, T.nscaf, bs), qsf = scaf_of(Q.off, Q.nscaf, qs);
                o_tile[slot] = (uint32_t)ts * (uint32_t)Q.nscaf + (uint32_t)qsf;
                o_s1[slot] = (int32_t)(bs - T.off[ts]);
                o_s2[slot] = (int32_t)(qs - Q.off[qsf]);
                o_len[slot] = (int32_t)(be - bs);
                o_score[slot] = score;
            }
        }
    }
    if (cells) atomicAdd(&counters[CNT_S2_CELLS], cells);
    if (extended) atomicAdd(&counters[CNT_EXTENDED], extended);
}

// ---- canonical ordering (tile, s1, s2, len) by two stable radix sorts on packed keys
__global__ void __launch_bounds__(256)
hsp_key1_kernel(const int32_t* __restrict__ s2, const int32_t* __restrict__ len, uint32_t n, int lb, uint64_t* __restrict__ key, uint32_t* __restrict__ idx) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    key[k] = ((uint64_t)(uint32_t)s2[k] << lb) | (uint64_t)(uint32_t)len[k];
    idx[k] = k;
}
__global__ void __launch_bounds__(256)
hsp_key2_kernel(const uint32_t* __restrict__ tile, const int32_t* __restrict__ s1, const uint32_t* __restrict__ perm, uint32_t n, int lb, uint64_t* __restrict__ key) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t src = perm[k];
    key[k] = ((uint64_t)tile[src] << lb) | (uint64_t)(uint32_t)s1[src];
}
__global__ void __launch_bounds__(256)
hsp_gather_kernel(const uint32_t* __restrict__ perm, uint32_t n, const uint32_t* __restrict__ tile, const int32_t* __restrict__ s1,
                  const int32_t* __restrict__ s2, const int32_t* __restrict__ len, const int32_t* __restrict__ score,
                  uint32_t* __restrict__ o_tile, int32_t* __restrict__ o_s1, int32_t* __restrict__ o_s2, int32_t* __restrict__ o_len,
                  int32_t* __restrict__ o_score) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t src = perm[k];
    o_tile[k] = tile[src]; o_s1[k] = s1[src]; o_s2[k] = s2[src]; o_len[k] = len[src]; o_score[k] = score[src];
}

static int bits_for(uint64_t v) { int b = 1; while (b < 64 && (v >> b)) b++; return b; }

void find_hsps(const Genome& T, const Genome& Q, uint64_t* surv0, uint64_t* surv1, uint32_t nsurv, const AlignParams& p,
               HspSet& out, unsigned long long* counters, const int32_t* d_same_q) {
    out.n = 0;
    if (nsurv == 0) return;
    Ctx& cx = ctx();
    NoVal* nv = nullptr;
    uint64_t* sorted;
    {
        ProfScope ps("surv_sort");
        int w = radix_sort_bits<uint64_t, NoVal>(surv0, surv1, nv, nv, nsurv, 0, bits_for(Q.G));
        uint64_t* a = w ? surv1 : surv0; uint64_t* b = w ? surv0 : surv1;
        w = radix_sort_bits<uint64_t, NoVal>(a, b, nv, nv, nsurv, 32, 32 + bits_for(T.G + Q.G));
        sorted = w ? b : a;
    }
    DevBuf<uint32_t> flag(nsurv), flag_off(nsurv), seg_start(nsurv), d_nseg(1);
    launch(diag_heads_kernel, cdiv(nsurv, 256), 256, 0, sorted, nsurv, view(T), (uint32_t)Q.G, flag.get());
    exclusive_scan_u32(flag.get(), flag_off.get(), nsurv, d_nseg.get());
    launch(diag_starts_kernel, cdiv(nsurv, 256), 256, 0, flag.get(), flag_off.get(), nsurv, seg_start.get());

    // every survivor yields at most one HSP, but usually far fewer: start with a quarter and redo the (deterministic)
    // walk with the exact size if that was too small
    uint32_t hcap = std::min<uint32_t>(nsurv, std::max<uint32_t>(1u << 20, nsurv / 4));
    DevBuf<uint32_t> r_tile;
    DevBuf<int32_t> r_s1, r_s2, r_len, r_score;
    unsigned long long h_n = 0;
    uint32_t h_nseg = 0;
    MB2_CUDA(cudaMemcpyAsync(&h_nseg, d_nseg.get(), sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
    MB2_CUDA(cudaStreamSynchronize(cx.stream));
    DevBuf<uint32_t> items(3 * (size_t)std::max<uint32_t>(h_nseg, 1u)), d_nitems(1);   // at most one hand-over per diagonal segment
    for (int attempt = 0; attempt < 2; attempt++) {
        r_tile.alloc(hcap); r_s1.alloc(hcap); r_s2.alloc(hcap); r_len.alloc(hcap); r_score.alloc(hcap);
        MB2_CUDA(cudaMemsetAsync(counters + CNT_WORK, 0, sizeof(unsigned long long), cx.stream));
        MB2_CUDA(cudaMemsetAsync(counters + CNT_HSPS, 0, sizeof(unsigned long long), cx.stream));
        if (attempt) {   // the first walk already counted these
            MB2_CUDA(cudaMemsetAsync(counters + CNT_S2_CELLS, 0, sizeof(unsigned long long), cx.stream));
            MB2_CUDA(cudaMemsetAsync(counters + CNT_EXTENDED, 0, sizeof(unsigned long long), cx.stream));
        }
        {
            ProfScope ps("hsp_extend");
            MB2_REQUIRE(p.xdrop >= XT_MIN_XDROP, -2, "x-drop below 251 is not supported by the three-column extension table");
            MB2_CUDA(cudaMemsetAsync(d_nitems.get(), 0, sizeof(uint32_t), cx.stream));
            const unsigned grid = (unsigned)cx.sm_count * 8;
            launch(hsp_extend_thread_kernel, grid, 128, 0, view(T), view(Q), sorted, nsurv, seg_start.get(), d_nseg.get(), (uint32_t)Q.G,
                   p.xdrop, p.hspthresh, p.entropy, hcap, r_tile.get(), r_s1.get(), r_s2.get(), r_len.get(), r_score.get(),
                   items.get(), d_nitems.get(), d_same_q, counters);
            MB2_CUDA(cudaMemsetAsync(counters + CNT_WORK, 0, sizeof(unsigned long long), cx.stream));
            launch(hsp_extend_kernel, grid, 128, 0, view(T), view(Q), sorted, nsurv, seg_start.get(), d_nseg.get(),
                   (const uint32_t*)items.get(), (const uint32_t*)d_nitems.get(), (uint32_t)Q.G,
                   p.xdrop, p.hspthresh, p.entropy, hcap, r_tile.get(), r_s1.get(), r_s2.get(), r_len.get(), r_score.get(), counters);
        }
        MB2_CUDA(cudaMemcpyAsync(&h_n, counters + CNT_HSPS, sizeof(h_n), cudaMemcpyDeviceToHost, cx.stream));
        MB2_CUDA(cudaStreamSynchronize(cx.stream));
        if (h_n <= hcap) break;
        MB2_REQUIRE(h_n <= nsurv && attempt == 0, -5, "hsp stage: internal overflow");
        hcap = (uint32_t)h_n;
    }
    const uint32_t n = (uint32_t)h_n;
    out.n = n;
    if (n == 0) return;
    // canonical order
    uint32_t maxlen = 0;
    for (int s = 0; s < T.nscaf; s++) maxlen = std::max(maxlen, T.len[s]);
    for (int s = 0; s < Q.nscaf; s++) maxlen = std::max(maxlen, Q.len[s]);
    const int lb = bits_for(maxlen);
    const int tb = bits_for((uint64_t)T.nscaf * (uint64_t)Q.nscaf);
    MB2_REQUIRE(2 * lb <= 64 && tb + lb <= 64, -3, "hsp stage: key does not fit 64 bits");
    DevBuf<uint64_t> k0(n), k1(n);
    DevBuf<uint32_t> i0(n), i1(n);
    ProfScope ps("hsp_sort");
    launch(hsp_key1_kernel, cdiv(n, 256), 256, 0, r_s2.get(), r_len.get(), n, lb, k0.get(), i0.get());
    int w = radix_sort_bits<uint64_t, uint32_t>(k0.get(), k1.get(), i0.get(), i1.get(), n, 0, 2 * lb);
    uint32_t* perm = w ? i1.get() : i0.get();
    uint32_t* perm_other = w ? i0.get() : i1.get();
    launch(hsp_key2_kernel, cdiv(n, 256), 256, 0, r_tile.get(), r_s1.get(), perm, n, lb, k0.get());
    w = radix_sort_bits<uint64_t, uint32_t>(k0.get(), k1.get(), perm, perm_other, n, 0, tb + lb);
    const uint32_t* fin = w ? perm_other : perm;
    out.tile.alloc(n); out.s1.alloc(n); out.s2.alloc(n); out.len.alloc(n); out.score.alloc(n);
    launch(hsp_gather_kernel, cdiv(n, 256), 256, 0, fin, n, r_tile.get(), r_s1.get(), r_s2.get(), r_len.get(), r_score.get(),
           out.tile.get(), out.s1.get(), out.s2.get(), out.len.get(), out.score.get());
}

}  // namespace mb2
