// gapped.cu -- kernel family (c): anchors + affine-gap y-drop extension, warp-shuffle anti-diagonal
// dynamic programming, integer only (no tensor cores: this is not a dense contraction).
//
// Replaces LASTZ's --gapped stage under spec D4/D5 of oracle/lastz_oracle.c:
//   * every chained HSP is reduced to an anchor = centre of its best 31-column window;
//   * per tile, anchors are taken best-first; an anchor inside the bounding box of an alignment
//     already reported for the tile is skipped;
//   * from the anchor the alignment is extended forwards and backwards by an affine-gap DP
//     (open 400, extend 30) evaluated one anti-diagonal at a time: a warp computes 32 cells of the
//     anti-diagonal per step (each cell needs only its up/left neighbours on the previous anti-diagonal
//     and its diagonal neighbour two back); a cell survives iff H >= best - ydrop, where best is the
//     maximum over all earlier anti-diagonals. The three DP states each carry (matches, aligned
//     columns) of their arg-max path, so identity needs no traceback;
//   * keep the alignment if forward + backward score >= gappedthresh.
// One warp owns one tile (its anchors are inherently sequential); tiles are scheduled dynamically.
#include "primitives.cuh"
#include "seq.cuh"
#include "internal.cuh"

namespace mb2 {

constexpr int GP_NT = 256;             // threads per CTA: one cell per thread for bands up to 512, 4 warps per scheduler hide latency
constexpr int GP_WARPS = GP_NT / 32;
constexpr int GP_CAP = 1024;           // circular row capacity of one anti-diagonal; the y-drop band must stay below it
constexpr int GP_MASK = GP_CAP - 1;
constexpr int NEG_INF = INT_MIN / 4;
// shared memory: A={h,hm,hc,d} for three anti-diagonals, B={i,dm,dc,im} and C={ic} for two
constexpr int GP_SMEM_INTS = (3 * 4 + 2 * 4 + 2 * 1) * GP_CAP;
constexpr size_t GP_SMEM_BYTES = (size_t)GP_SMEM_INTS * sizeof(int);

__device__ __forceinline__ int sub_lut3(uint32_t idx) {
    const uint64_t lo = 0xE183648E85E18E5Bull, hi = 0x5B8EE1858E6483E1ull;
    const uint64_t v = (idx & 8) ? hi : lo;
    return (int)(int8_t)(v >> ((idx & 7) * 8));
}

// ---- ordering of the chain members: per tile, score descending, then canonical (s1, s2)
__global__ void __launch_bounds__(256)
anchor_keys_kernel(const uint32_t* __restrict__ tile, const int32_t* __restrict__ score, const uint8_t* __restrict__ in_chain,
                   uint32_t n, uint64_t* __restrict__ key, uint32_t* __restrict__ idx) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    // non-members sort behind every real tile
    key[k] = in_chain[k] ? (((uint64_t)tile[k] << 31) | (uint64_t)(0x7fffffff - score[k])) : ~0ull;
    idx[k] = k;
}
__global__ void __launch_bounds__(256)
anchor_heads_kernel(const uint64_t* __restrict__ key, uint32_t n, uint32_t* __restrict__ flag) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const bool member = key[k] != ~0ull;
    flag[k] = (member && (k == 0 || (key[k] >> 31) != (key[k - 1] >> 31))) ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
anchor_member_kernel(const uint64_t* __restrict__ key, uint32_t n, uint32_t* __restrict__ flag) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    flag[k] = key[k] != ~0ull ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
anchor_starts_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ flag_off, uint32_t n, uint32_t* __restrict__ seg_start) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (flag[k]) seg_start[flag_off[k]] = k;
}

struct Ext { int score, di, dj, nmatch, ncols; };

struct WarpRec { int wmax, wi, hm, hc, alo, ahi, pad0, pad1; };

// One-sided y-drop extension by the whole CTA. DIR=+1: cell (i,j) consumes T[ta+i-1], Q[qa+j-1]; DIR=-1: T[ta-i], Q[qa-j].
// All DP state lives in shared memory, indexed circularly by the row i, as 16-byte records so that one cell costs
// 6 vector loads and 3 vector stores:  A = {h, hm, hc, d} (three anti-diagonals kept: k, k-1, k-2),
// B = {i, dm, dc, im} and C = {ic} (two kept). One __syncthreads per anti-diagonal.
template <int DIR>
__device__ Ext ydrop_extend_cta(const GenomeView& T, const GenomeView& Q, uint32_t ta, uint32_t qa, int tn, int qn, int O, int E, int Y,
                                int* __restrict__ sm, WarpRec (*rec)[GP_WARPS], const int* __restrict__ sub5,
                                unsigned long long& cells, int& err) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int4* As = reinterpret_cast<int4*>(sm);                       // [3][CAP]
    int4* Bs = reinterpret_cast<int4*>(sm + 12 * GP_CAP);         // [2][CAP]
    int* Cs = sm + 20 * GP_CAP;                                   // [2][CAP]
    Ext r = {0, 0, 0, 0, 0};
    __syncthreads();                    // previous users of the buffers are done
    if (tid == 0) {                     // anti-diagonal 0 = the single cell (0,0), generation 0
        As[0] = make_int4(0, 0, 0, NEG_INF); Bs[0] = make_int4(NEG_INF, 0, 0, 0); Cs[0] = 0;
    }
    __syncthreads();
    int lo2 = 0, hi2 = -1, lo1 = 0, hi1 = 0;
    int best = 0;
    const uint32_t kmax = (uint32_t)tn + (uint32_t)qn;
    int g0 = 1, g1 = 0, g2 = 2;         // generation slots of anti-diagonals k, k-1, k-2
    for (uint32_t k = 1; k <= kmax; k++) {
        int clo = INT_MAX, chi = INT_MIN;
        if (hi1 >= lo1) { clo = lo1; chi = hi1 + 1; }
        if (hi2 >= lo2) { clo = min(clo, lo2 + 1); chi = max(chi, hi2 + 1); }
        if (clo == INT_MAX) break;                                   // two dead anti-diagonals in a row
        clo = max(clo, 0);
        if (k > (uint32_t)qn) clo = max(clo, (int)(k - (uint32_t)qn));
        chi = min(chi, (int)min(k, (uint32_t)tn));
        if (chi - clo + 1 > GP_CAP - 2) { err = 1; break; }
        const int e0 = (int)(k & 1), e1 = e0 ^ 1;
        int4* Ac = As + g0 * GP_CAP; const int4* A1 = As + g1 * GP_CAP; const int4* A2 = As + g2 * GP_CAP;
        int4* Bc = Bs + e0 * GP_CAP; const int4* B1 = Bs + e1 * GP_CAP;
        int* Cc = Cs + e0 * GP_CAP; const int* C1 = Cs + e1 * GP_CAP;
        const int thr = best - Y;
        int tmax = INT_MIN, ti = 0, thm = 0, thc = 0, talo = INT_MAX, tahi = INT_MIN;
        for (int i = clo + tid; i <= chi; i += GP_NT) {
            const int j = (int)k - i;
            const int x = i & GP_MASK, xu = (i - 1) & GP_MASK;
            // the substitution score does not depend on the DP state: fetch it first (padding keeps these reads in bounds)
            const uint32_t ct = DIR > 0 ? ta + (uint32_t)i - 1u : ta - (uint32_t)i;
            const uint32_t cq = DIR > 0 ? qa + (uint32_t)j - 1u : qa - (uint32_t)j;
            const uint32_t tb = T.codes[ct], qb = Q.codes[cq];
            const int sc = sub5[tb * 5 + qb];
            const int ismatch = (tb == qb && tb < 4) ? 1 : 0;
            int h, d = NEG_INF, ii = NEG_INF, hm, hc, dm = 0, dc = 0, im = 0, ic = 0;
            if (i - 1 >= lo1 && i - 1 <= hi1) {                       // up: (i-1, j) on k-1
                const int4 ua = A1[xu];                               // {h, hm, hc, d}
                if (ua.x > NEG_INF) {
                    const int open = ua.x - O - E, ext = ua.w > NEG_INF ? ua.w - E : NEG_INF;
                    if (open >= ext) { d = open; dm = ua.y; dc = ua.z; }
                    else { const int4 ub = B1[xu]; d = ext; dm = ub.y; dc = ub.z; }
                }
            }
            if (i >= lo1 && i <= hi1) {                               // left: (i, j-1) on k-1
                const int4 la = A1[x];
                if (la.x > NEG_INF) {
                    const int4 lb = B1[x];                            // {i, dm, dc, im}
                    const int open = la.x - O - E, ext = lb.x > NEG_INF ? lb.x - E : NEG_INF;
                    if (open >= ext) { ii = open; im = la.y; ic = la.z; }
                    else { ii = ext; im = lb.w; ic = C1[x]; }
                }
            }
            int mval = NEG_INF, mm = 0, mc = 0;
            if (i >= 1 && j >= 1 && i - 1 >= lo2 && i - 1 <= hi2) {   // diagonal: (i-1, j-1) on k-2
                const int4 da = A2[xu];
                if (da.x > NEG_INF) { mval = da.x + sc; mm = da.y + ismatch; mc = da.z + 1; }
            }
            if (mval >= d && mval >= ii) { h = mval; hm = mm; hc = mc; }
            else if (d >= ii) { h = d; hm = dm; hc = dc; }
            else { h = ii; hm = im; hc = ic; }
            if (h <= NEG_INF || h < thr) { h = NEG_INF; d = NEG_INF; ii = NEG_INF; }
            else {
                talo = min(talo, i); tahi = max(tahi, i);
                if (h > tmax) { tmax = h; ti = i; thm = hm; thc = hc; }   // i increases along the loop: first = smallest row
            }
            Ac[x] = make_int4(h, hm, hc, d); Bc[x] = make_int4(ii, dm, dc, im); Cc[x] = ic;
        }
        // per-warp summary: best cell (max h, then smallest row) and alive row range
        const int wmax = __reduce_max_sync(0xffffffffu, tmax);
        const int walo = __reduce_min_sync(0xffffffffu, talo), wahi = __reduce_max_sync(0xffffffffu, tahi);
        int wi = 0, whm = 0, whc = 0;
        if (wmax > best) {
            wi = __reduce_min_sync(0xffffffffu, tmax == wmax ? ti : INT_MAX);
            const int src = __ffs(__ballot_sync(0xffffffffu, tmax == wmax && ti == wi)) - 1;
            whm = __shfl_sync(0xffffffffu, thm, src); whc = __shfl_sync(0xffffffffu, thc, src);
        }
        const int par = (int)(k & 1);
        if (lane == 0) {
            int4* rp = reinterpret_cast<int4*>(&rec[par][warp]);
            rp[0] = make_int4(wmax, wi, whm, whc); rp[1] = make_int4(walo, wahi, 0, 0);
        }
        __syncthreads();
        // every warp combines the per-warp records lane-parallel (lane w reads record w)
        int4 q0 = make_int4(INT_MIN, INT_MAX, 0, 0), q1 = make_int4(INT_MAX, INT_MIN, 0, 0);
        if (lane < GP_WARPS) {
            const int4* rp = reinterpret_cast<const int4*>(&rec[par][lane]);
            q0 = rp[0]; q1 = rp[1];
        }
        const int alo = __reduce_min_sync(0xffffffffu, q1.x), ahi = __reduce_max_sync(0xffffffffu, q1.y);
        const int bmax = __reduce_max_sync(0xffffffffu, q0.x);
        int bi = 0, bhm = 0, bhc = 0;
        if (bmax > best) {
            bi = __reduce_min_sync(0xffffffffu, q0.x == bmax ? q0.y : INT_MAX);
            const int src = __ffs(__ballot_sync(0xffffffffu, q0.x == bmax && q0.y == bi)) - 1;
            bhm = __shfl_sync(0xffffffffu, q0.z, src); bhc = __shfl_sync(0xffffffffu, q0.w, src);
        }
        if (bmax > best) { best = bmax; r.score = bmax; r.di = bi; r.dj = (int)k - bi; r.nmatch = bhm; r.ncols = bhc; }
        cells += (unsigned long long)(chi >= clo ? chi - clo + 1 : 0);
        lo2 = lo1; hi2 = hi1;
        if (ahi >= alo && alo != INT_MAX) { lo1 = alo; hi1 = ahi; } else { lo1 = 0; hi1 = -1; }
        const int gt = g2; g2 = g1; g1 = g0; g0 = gt;
    }
    return r;
}

// Exact shortcut for the trivial alignment of a scaffold with itself. When target and query are the SAME N-free
// sequence and the anchor lies on the main diagonal, the y-drop DP has a closed form: every column (x,x) is a match
// scoring s(b,b) > 0, and any other path to an anti-diagonal k uses at most min(i,j) <= k/2 aligned columns, each
// worth at most s(b,b) of its row base, minus gap costs -- so the main-diagonal cell is the strict maximum of every
// even anti-diagonal, is never pruned, and the extension ends at the scaffold end with score = sum of s(b,b),
// matches = columns = length. (Proof in DESIGN.md; sequences with any non-ACGT base take the general DP.)
__device__ Ext selfdiag_extend_cta(const GenomeView& T, uint32_t p0, uint32_t p1, int* __restrict__ sm) {
    // count C/G bases in padded-coordinate range [p0, p1): 2-bit code has exactly one bit set for C (01) and G (10)
    int cg = 0;
    if (p1 > p0) {
        const uint32_t w0 = p0 >> 5, w1 = (p1 - 1) >> 5;
        for (uint32_t w = w0 + threadIdx.x; w <= w1; w += blockDim.x) {
            uint64_t x = T.pk[w];
            uint64_t m = (x ^ (x >> 1)) & 0x5555555555555555ull;
            if (w == w0 && (p0 & 31)) m &= ~0ull << (2 * (p0 & 31));
            if (w == w1 && (p1 & 31)) m &= ~0ull >> (64 - 2 * (p1 & 31));
            cg += __popcll(m);
        }
    }
    cg = __reduce_add_sync(0xffffffffu, cg);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = cg;
    __syncthreads();
    int tot = 0;
    for (int w = 0; w < GP_WARPS; w++) tot += sm[w];
    __syncthreads();
    const int n = (int)(p1 - p0);
    Ext r;
    r.score = 91 * (n - tot) + 100 * tot; r.di = n; r.dj = n; r.nmatch = n; r.ncols = n;
    return r;
}

// best 31-column window of an HSP (first maximum); returns the anchor offset inside the HSP
__device__ int anchor_offset(const GenomeView& T, const GenomeView& Q, uint32_t ts, uint32_t qs, int len, int lane) {
    if (len <= 31) return len / 2;
    int best = INT_MIN, bw = 0;
    for (int base = 0; base + 31 <= len; base += 32) {
        // scores of columns base+lane and base+32+lane (the second only below len)
        const int c0 = base + lane, c1 = base + 32 + lane;
        int s0 = 0, s1 = 0;
        if (c0 < len) {
            const uint32_t an = isn_at(T.nm, ts + c0) | isn_at(Q.nm, qs + c0);
            s0 = an ? SCORE_N : sub_lut3((base_at(T.pk, ts + c0) << 2) | base_at(Q.pk, qs + c0));
        }
        if (c1 < len) {
            const uint32_t an = isn_at(T.nm, ts + c1) | isn_at(Q.nm, qs + c1);
            s1 = an ? SCORE_N : sub_lut3((base_at(T.pk, ts + c1) << 2) | base_at(Q.pk, qs + c1));
        }
        int p0 = s0, p1 = s1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t0 = __shfl_up_sync(0xffffffffu, p0, d), t1 = __shfl_up_sync(0xffffffffu, p1, d);
            if (lane >= d) { p0 += t0; p1 += t1; }
        }
        p1 += __shfl_sync(0xffffffffu, p0, 31);
        // window w = base + lane covers columns [w, w+30]: P[lane+30] - P[lane-1]
        const int hi_idx = lane + 30;
        const int a0 = __shfl_sync(0xffffffffu, p0, hi_idx & 31), a1 = __shfl_sync(0xffffffffu, p1, hi_idx & 31);
        const int top = hi_idx < 32 ? a0 : a1;
        int bot = __shfl_up_sync(0xffffffffu, p0, 1);
        if (lane == 0) bot = 0;
        const bool valid = base + lane + 31 <= len;
        const int wsum = valid ? top - bot : INT_MIN;
        const int mx = __reduce_max_sync(0xffffffffu, wsum);
        if (mx > best) {
            best = mx;
            bw = base + __ffs(__ballot_sync(0xffffffffu, wsum == mx)) - 1;
        }
    }
    return bw + 15;
}

__global__ void __launch_bounds__(GP_NT, 1)
gapped_kernel(GenomeView T, GenomeView Q, const uint32_t* __restrict__ tile, const int32_t* __restrict__ hs1, const int32_t* __restrict__ hs2,
              const int32_t* __restrict__ hlen, const uint32_t* __restrict__ order, uint32_t nmember,
              const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ nseg_p,
              int O, int E, int Y, int gthr,
              int32_t* __restrict__ o_s1, int32_t* __restrict__ o_e1, int32_t* __restrict__ o_s2, int32_t* __restrict__ o_e2,
              int32_t* __restrict__ o_score, int32_t* __restrict__ o_nm, int32_t* __restrict__ o_nc, uint32_t* __restrict__ o_tile,
              uint32_t* __restrict__ o_keep, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) int gp_smem[];
    __shared__ __align__(16) WarpRec rec[2][GP_WARPS];
    __shared__ uint32_t sh_seg;
    __shared__ int sh_off;
    __shared__ int sub5[25];
    if (threadIdx.x < 25) {
        const int a = threadIdx.x / 5, b = threadIdx.x % 5;
        sub5[threadIdx.x] = (a == 4 || b == 4) ? SCORE_N : c_sub[a * 4 + b];
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t nseg = *nseg_p;
    unsigned long long cells = 0, anchors = 0;
    int err = 0;
    for (;;) {
        if (tid == 0) sh_seg = (uint32_t)atomicAdd(&counters[CNT_WORK], 1ull);
        __syncthreads();
        const uint32_t seg = sh_seg;
        if (seg >= nseg) break;
        const uint32_t a = seg_start[seg];
        const uint32_t b = (seg + 1 < nseg) ? seg_start[seg + 1] : nmember;
        const uint32_t tl = tile[order[a]];
        const uint32_t tsc = tl / (uint32_t)Q.nscaf, qsc = tl % (uint32_t)Q.nscaf;
        const uint32_t toff = T.off[tsc], qoff = Q.off[qsc];
        const int tlen = (int)T.len[tsc], qlen = (int)Q.len[qsc];
        uint32_t nkept = 0;                                      // kept alignments of this tile live in slots [a, a+nkept)
        for (uint32_t x = a; x < b; x++) {
            const uint32_t g = order[x];
            const int s1 = hs1[g], s2 = hs2[g], len = hlen[g];
            if (warp == 0) {
                const int off = anchor_offset(T, Q, toff + s1, qoff + s2, len, lane);
                if (lane == 0) sh_off = off;
            }
            __syncthreads();
            const int a1 = s1 + sh_off, a2 = s2 + sh_off;
            int cov = 0;                                          // spec D5: bounding-box test against reported alignments
            for (uint32_t kk = tid; kk < nkept; kk += GP_NT)
                cov |= (a1 >= o_s1[a + kk] && a1 < o_e1[a + kk] && a2 >= o_s2[a + kk] && a2 < o_e2[a + kk]) ? 1 : 0;
            if (__syncthreads_or(cov)) continue;
            anchors++;
            Ext f, r;
            if (T.pk == Q.pk && tsc == qsc && a1 == a2 && T.nfree[tsc]) {
                f = selfdiag_extend_cta(T, toff + a1, toff + tlen, gp_smem);
                r = selfdiag_extend_cta(T, toff, toff + a1, gp_smem);
            } else {
                f = ydrop_extend_cta<+1>(T, Q, toff + a1, qoff + a2, tlen - a1, qlen - a2, O, E, Y, gp_smem, rec, sub5, cells, err);
                r = ydrop_extend_cta<-1>(T, Q, toff + a1, qoff + a2, a1, a2, O, E, Y, gp_smem, rec, sub5, cells, err);
            }
            const int score = f.score + r.score;
            if (score < gthr) continue;
            if (tid == 0) {
                const uint32_t o = a + nkept;
                o_s1[o] = a1 - r.di; o_e1[o] = a1 + f.di; o_s2[o] = a2 - r.dj; o_e2[o] = a2 + f.dj;
                o_score[o] = score; o_nm[o] = f.nmatch + r.nmatch; o_nc[o] = f.ncols + r.ncols; o_tile[o] = tl; o_keep[o] = 1;
                __threadfence_block();
            }
            nkept++;
            __syncthreads();
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (cells) atomicAdd(&counters[CNT_GAPPED_CELLS], cells);
        if (anchors) atomicAdd(&counters[CNT_ANCHORS], anchors);
        if (err) atomicAdd(&counters[CNT_ERR], 1ull);
    }
}

__global__ void __launch_bounds__(256)
aln_gather_kernel(const uint32_t* __restrict__ keep, const uint32_t* __restrict__ keep_off, uint32_t n,
                  const uint32_t* __restrict__ i_tile, const int32_t* __restrict__ i_s1, const int32_t* __restrict__ i_e1,
                  const int32_t* __restrict__ i_s2, const int32_t* __restrict__ i_e2, const int32_t* __restrict__ i_score,
                  const int32_t* __restrict__ i_nm, const int32_t* __restrict__ i_nc,
                  uint32_t* __restrict__ o_tile, int32_t* __restrict__ o_s1, int32_t* __restrict__ o_e1, int32_t* __restrict__ o_s2,
                  int32_t* __restrict__ o_e2, int32_t* __restrict__ o_score, int32_t* __restrict__ o_nm, int32_t* __restrict__ o_nc) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || !keep[k]) return;
    const uint32_t o = keep_off[k];
    o_tile[o] = i_tile[k]; o_s1[o] = i_s1[k]; o_e1[o] = i_e1[k]; o_s2[o] = i_s2[k]; o_e2[o] = i_e2[k];
    o_score[o] = i_score[k]; o_nm[o] = i_nm[k]; o_nc[o] = i_nc[k];
}

// ungapped mode (--gapped off): every chain member becomes an alignment row; matches counted here
__global__ void __launch_bounds__(128)
ungapped_rows_kernel(GenomeView T, GenomeView Q, const uint32_t* __restrict__ tile, const int32_t* __restrict__ hs1,
                     const int32_t* __restrict__ hs2, const int32_t* __restrict__ hlen, const int32_t* __restrict__ hscore,
                     const uint8_t* __restrict__ in_chain, uint32_t n,
                     int32_t* __restrict__ o_s1, int32_t* __restrict__ o_e1, int32_t* __restrict__ o_s2, int32_t* __restrict__ o_e2,
                     int32_t* __restrict__ o_score, int32_t* __restrict__ o_nm, int32_t* __restrict__ o_nc, uint32_t* __restrict__ o_tile,
                     uint32_t* __restrict__ o_keep) {
    const int lane = threadIdx.x & 31;
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= n || !in_chain[g]) return;
    const uint32_t tl = tile[g];
    const uint32_t ts = T.off[tl / (uint32_t)Q.nscaf] + hs1[g], qs = Q.off[tl % (uint32_t)Q.nscaf] + hs2[g];
    int nm = 0;
    for (int c = lane; c < hlen[g]; c += 32)
        nm += (!(isn_at(T.nm, ts + c) | isn_at(Q.nm, qs + c)) && base_at(T.pk, ts + c) == base_at(Q.pk, qs + c)) ? 1 : 0;
    nm = __reduce_add_sync(0xffffffffu, nm);
    if (lane == 0) {
        o_s1[g] = hs1[g]; o_e1[g] = hs1[g] + hlen[g]; o_s2[g] = hs2[g]; o_e2[g] = hs2[g] + hlen[g];
        o_score[g] = hscore[g]; o_nm[g] = nm; o_nc[g] = hlen[g]; o_tile[g] = tl; o_keep[g] = 1;
    }
}

void gapped_extend(const Genome& T, const Genome& Q, const HspSet& h, const DevBuf<uint8_t>& in_chain, const AlignParams& p,
                   AlnSet& out, unsigned long long* counters) {
    out.n = 0;
    const uint32_t n = h.n;
    if (n == 0) return;
    Ctx& cx = ctx();
    DevBuf<uint32_t> r_tile(n), keep(n), keep_off(n), d_nout(1);
    DevBuf<int32_t> r_s1(n), r_e1(n), r_s2(n), r_e2(n), r_score(n), r_nm(n), r_nc(n);
    MB2_CUDA(cudaMemsetAsync(keep.get(), 0, (size_t)n * sizeof(uint32_t), cx.stream));
    if (!p.gapped) {
        launch(ungapped_rows_kernel, cdiv((size_t)n * 32, 128), 128, 0, view(T), view(Q), h.tile.get(), h.s1.get(), h.s2.get(),
               h.len.get(), h.score.get(), in_chain.get(), n, r_s1.get(), r_e1.get(), r_s2.get(), r_e2.get(), r_score.get(),
               r_nm.get(), r_nc.get(), r_tile.get(), keep.get());
    } else {
        // order the chain members: (tile, score desc), stable over the canonical order
        int tb = 1; while (tb < 33 && (((uint64_t)T.nscaf * (uint64_t)Q.nscaf) >> tb)) tb++;
        DevBuf<uint64_t> k0(n), k1(n);
        DevBuf<uint32_t> i0(n), i1(n);
        launch(anchor_keys_kernel, cdiv(n, 256), 256, 0, h.tile.get(), h.score.get(), in_chain.get(), n, k0.get(), i0.get());
        const int w = radix_sort_bits<uint64_t, uint32_t>(k0.get(), k1.get(), i0.get(), i1.get(), n, 0, std::min(64, tb + 31));
        const uint64_t* skey = w ? k1.get() : k0.get();
        const uint32_t* order = w ? i1.get() : i0.get();
        DevBuf<uint32_t> flag(n), flag_off(n), seg_start(n), d_nseg(1);
        launch(anchor_heads_kernel, cdiv(n, 256), 256, 0, skey, n, flag.get());
        exclusive_scan_u32(flag.get(), flag_off.get(), n, d_nseg.get());
        launch(anchor_starts_kernel, cdiv(n, 256), 256, 0, flag.get(), flag_off.get(), n, seg_start.get());
        // number of chain members = first index whose key is the non-member sentinel; computed on the device side by
        // passing n and letting segments end at the next head; the last segment must stop at the member count:
        uint32_t h_nseg = 0;
        MB2_CUDA(cudaMemcpyAsync(&h_nseg, d_nseg.get(), sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
        // member count = number of in_chain flags; reuse scan on a temporary
        DevBuf<uint32_t> mflag(n), moff(n), d_nmember(1);
        launch(anchor_member_kernel, cdiv(n, 256), 256, 0, skey, n, mflag.get());
        exclusive_scan_u32(mflag.get(), moff.get(), n, d_nmember.get());
        uint32_t h_nmember = 0;
        MB2_CUDA(cudaMemcpyAsync(&h_nmember, d_nmember.get(), sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
        MB2_CUDA(cudaStreamSynchronize(cx.stream));
        if (h_nmember) {
            static bool attr_set = false;
            if (!attr_set) {
                MB2_CUDA(cudaFuncSetAttribute(gapped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GP_SMEM_BYTES));
                attr_set = true;
            }
            const unsigned blocks = std::min<unsigned>((unsigned)cx.sm_count * 2, std::max<unsigned>(1u, h_nseg));   // 88 KB smem + 512 threads: 2 CTAs/SM
            MB2_CUDA(cudaMemsetAsync(counters + CNT_WORK, 0, sizeof(unsigned long long), cx.stream));
            ProfScope ps("gapped");
            launch(gapped_kernel, blocks, GP_NT, GP_SMEM_BYTES, view(T), view(Q), h.tile.get(), h.s1.get(), h.s2.get(), h.len.get(),
                   order, h_nmember, seg_start.get(), d_nseg.get(), p.gap_open, p.gap_extend, p.ydrop, p.gappedthresh,
                   r_s1.get(), r_e1.get(), r_s2.get(), r_e2.get(), r_score.get(), r_nm.get(), r_nc.get(), r_tile.get(), keep.get(), counters);
        }
    }
    exclusive_scan_u32(keep.get(), keep_off.get(), n, d_nout.get());
    uint32_t h_nout = 0; unsigned long long h_err = 0;
    MB2_CUDA(cudaMemcpyAsync(&h_nout, d_nout.get(), sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
    MB2_CUDA(cudaMemcpyAsync(&h_err, counters + CNT_ERR, sizeof(h_err), cudaMemcpyDeviceToHost, cx.stream));
    MB2_CUDA(cudaStreamSynchronize(cx.stream));
    MB2_REQUIRE(h_err == 0, -3, "gapped stage: y-drop band exceeded the anti-diagonal buffer capacity");
    out.n = h_nout;
    if (h_nout == 0) return;
    out.tile.alloc(h_nout); out.s1.alloc(h_nout); out.e1.alloc(h_nout); out.s2.alloc(h_nout); out.e2.alloc(h_nout);
    out.score.alloc(h_nout); out.nmatch.alloc(h_nout); out.ncols.alloc(h_nout);
    launch(aln_gather_kernel, cdiv(n, 256), 256, 0, keep.get(), keep_off.get(), n, r_tile.get(), r_s1.get(), r_e1.get(), r_s2.get(),
           r_e2.get(), r_score.get(), r_nm.get(), r_nc.get(), out.tile.get(), out.s1.get(), out.e1.get(), out.s2.get(), out.e2.get(),
           out.score.get(), out.nmatch.get(), out.ncols.get());
}

}  // namespace mb2
