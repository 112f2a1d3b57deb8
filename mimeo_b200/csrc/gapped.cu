// gapped.cu -- kernel family (c): anchors + affine-gap y-drop extension, anti-diagonal dynamic
// programming, integer only (no tensor cores: this is not a dense contraction).
//
// Replaces LASTZ's --gapped stage under spec D4/D5 of oracle/lastz_oracle.c:
//   * every chained HSP is reduced to an anchor = centre of its best 31-column window;
//   * per tile, anchors are taken best-first; an anchor inside the bounding box of an alignment
//     already reported for the tile is skipped;
//   * from the anchor the alignment is extended forwards and backwards by an affine-gap DP
//     (open 400, extend 30) evaluated one anti-diagonal at a time; a cell survives iff
//     H >= best - ydrop, where best is the maximum over all earlier anti-diagonals. The three DP
//     states each carry (matches, aligned columns) of their arg-max path, so identity needs no
//     traceback;
//   * keep the alignment if forward + backward score >= gappedthresh.
//
// Schedule. Under spec D5 an extension depends only on its anchor, never on the alignments reported
// before it; only the SKIP decision is sequential. So the stage runs in rounds:
//   gp_select_kernel   one warp per tile walks the tile's anchors best-first as far as results exist
//                      (accept / skip exactly as the sequential rule says), then schedules extensions
//                      for anchors that are still undecided: the first undecided one, plus -- as
//                      speculation -- the best anchor of every other "cluster" of the chain (run of
//                      chain members with small gaps, which one alignment usually swallows whole);
//   gp_extend_kernel   one CTA per (anchor, direction), all tiles in one launch.
// Speculation changes which extensions are computed early, never the result: a speculative result is
// only used when the sequential walk reaches its anchor and finds it uncovered.
#include "primitives.cuh"
#include "seq.cuh"
#include "internal.cuh"

namespace mb2 {

// CTA shape of the extension kernel: NT threads, each owning SLOTS consecutive diagonals of a circular window of
// ND = NT * SLOTS diagonal slots, slot = (i - j) & (ND - 1); a thread computes SLOTS/2 independent cells per anti-diagonal.
template <int NT_, int SLOTS_> struct GpShape {
    static constexpr int NT = NT_, SLOTS = SLOTS_, WARPS = NT_ / 32, ND = NT_ * SLOTS_, DMASK = ND - 1;
    static constexpr int MAXBAND = ND - 64;   // widest alive diagonal range the circular window can hold
};
constexpr int NEG_INF = INT_MIN / 4;
constexpr int GP_CLUSTER_GAP = 1000;     // chain members further apart than this (either axis) start a new speculation cluster
constexpr int GP_NARROW_ABORT_K = 140000;  // 16-bit payload run: beyond this anti-diagonal every cell has >= 65536 columns behind it

__device__ __forceinline__ int sub_lut3(uint32_t idx) {
    const uint64_t lo = 0xE183648E85E18E5Bull, hi = 0x5B8EE1858E6483E1ull;
    const uint64_t v = (idx & 8) ? hi : lo;
    return (int)(int8_t)(v >> ((idx & 7) * 8));
}

// ---- ordering of the chain members: per tile, score descending, then canonical (s1, s2)
__global__ void __launch_bounds__(256)
anchor_keys_kernel(const uint32_t* __restrict__ tile, const int32_t* __restrict__ score, const uint8_t* __restrict__ in_chain,
                   uint32_t n, uint64_t* __restrict__ key, uint32_t* __restrict__ idx, uint32_t* __restrict__ mflag) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    // non-members sort behind every real tile
    key[k] = in_chain[k] ? (((uint64_t)tile[k] << 31) | (uint64_t)(0x7fffffff - score[k])) : ~0ull;
    idx[k] = k;
    mflag[k] = in_chain[k] ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
anchor_heads_kernel(const uint64_t* __restrict__ key, uint32_t n, uint32_t* __restrict__ flag) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const bool member = key[k] != ~0ull;
    flag[k] = (member && (k == 0 || (key[k] >> 31) != (key[k - 1] >> 31))) ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
anchor_starts_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ flag_off, uint32_t n, uint32_t* __restrict__ seg_start) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (flag[k]) seg_start[flag_off[k]] = k;
}
// speculation clusters over the members in canonical (= chain) order: mlist[r] = HSP index of the r-th member
__global__ void __launch_bounds__(256)
member_list_kernel(const uint32_t* __restrict__ mflag, const uint32_t* __restrict__ mrank, uint32_t n, uint32_t* __restrict__ mlist) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || !mflag[k]) return;
    mlist[mrank[k]] = k;
}
__global__ void __launch_bounds__(256)
cluster_heads_kernel(const uint32_t* __restrict__ mlist, uint32_t nmember, const uint32_t* __restrict__ tile, const int32_t* __restrict__ hs1,
                     const int32_t* __restrict__ hs2, const int32_t* __restrict__ hlen, uint32_t* __restrict__ head) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nmember) return;
    uint32_t h = 1;
    if (r > 0) {
        const uint32_t g = mlist[r], pg = mlist[r - 1];
        const int g1 = hs1[g] - (hs1[pg] + hlen[pg]), g2 = hs2[g] - (hs2[pg] + hlen[pg]);
        h = (tile[g] != tile[pg] || g1 > GP_CLUSTER_GAP || g2 > GP_CLUSTER_GAP || g1 < -GP_CLUSTER_GAP || g2 < -GP_CLUSTER_GAP) ? 1u : 0u;
    }
    head[r] = h;
}
__global__ void __launch_bounds__(256)
cluster_ids_kernel(const uint32_t* __restrict__ mlist, uint32_t nmember, const uint32_t* __restrict__ head, const uint32_t* __restrict__ head_off,
                   uint32_t* __restrict__ cluster_of) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nmember) return;
    cluster_of[mlist[r]] = head_off[r] + head[r] - 1u;
}

struct Ext { int score, di, dj, nmatch, ncols; };

// ---- payload (matches, aligned columns) of a DP state, packed in one register (16+16 bits) or two (32+32)
template <typename P> struct Pay;
template <> struct Pay<uint32_t> {
    static __device__ __forceinline__ uint32_t step(uint32_t p, int match) { return p + 0x10000u + (uint32_t)match; }
    static __device__ __forceinline__ int nm(uint32_t p) { return (int)(p & 0xffffu); }
    static __device__ __forceinline__ int nc(uint32_t p) { return (int)(p >> 16); }
};
template <> struct Pay<uint64_t> {
    static __device__ __forceinline__ uint64_t step(uint64_t p, int match) { return p + (1ull << 32) + (uint64_t)(uint32_t)match; }
    static __device__ __forceinline__ int nm(uint64_t p) { return (int)(uint32_t)p; }
    static __device__ __forceinline__ int nc(uint64_t p) { return (int)(p >> 32); }
};

template <typename P> struct Cell { int h, d, i; P hp, dp, ip; };
// A dead cell holds NEG_INF in all three scores; arithmetic on it stays far below any threshold (thr >= -ydrop), so the
// recurrence needs no "is this predecessor alive" branches: a cell whose predecessors are all dead evaluates to ~NEG_INF,
// fails H >= thr and is reset to exactly NEG_INF.
template <typename P> __device__ __forceinline__ void cell_dead(Cell<P>& c) { c.h = c.d = c.i = NEG_INF; c.hp = c.dp = c.ip = 0; }

// what a thread publishes for its neighbours: slot 0 is read as the LEFT neighbour (h, i) of thread t-1's last slot,
// the last slot as the UP neighbour (h, d) of thread t+1's slot 0
template <typename P> struct Edge { int h, x; P hp, xp; };
template <typename P, int NT> struct EdgeBuf;
template <int NT> struct EdgeBuf<uint32_t, NT> {
    int4 v[NT];
    __device__ __forceinline__ void put(int t, const Edge<uint32_t>& e) { v[t] = make_int4(e.h, e.x, (int)e.hp, (int)e.xp); }
    __device__ __forceinline__ Edge<uint32_t> get(int t) const { const int4 a = v[t]; return Edge<uint32_t>{a.x, a.y, (uint32_t)a.z, (uint32_t)a.w}; }
};
template <int NT> struct EdgeBuf<uint64_t, NT> {
    int4 v[NT]; int2 w[NT];
    __device__ __forceinline__ void put(int t, const Edge<uint64_t>& e) {
        v[t] = make_int4(e.h, e.x, (int)(uint32_t)e.hp, (int)(uint32_t)(e.hp >> 32)); w[t] = make_int2((int)(uint32_t)e.xp, (int)(uint32_t)(e.xp >> 32));
    }
    __device__ __forceinline__ Edge<uint64_t> get(int t) const {
        const int4 a = v[t]; const int2 b = w[t];
        return Edge<uint64_t>{a.x, a.y, (uint64_t)(uint32_t)a.z | ((uint64_t)(uint32_t)a.w << 32), (uint64_t)(uint32_t)b.x | ((uint64_t)(uint32_t)b.y << 32)};
    }
};

template <typename P, typename S> struct GpShared {
    EdgeBuf<P, S::NT> left;             // slot 0 of every thread: {h, i, hp, ip}
    EdgeBuf<P, S::NT> up;               // last slot of every thread: {h, d, hp, dp}
    int smax[2][S::WARPS];              // per parity, per warp: maximum score of the anti-diagonal (NEG_INF = nothing alive)
    int2 rng[S::WARPS];                 // end of an epoch: per warp {lowest, highest} alive diagonal
    int4 fin[S::WARPS]; int2 finp[S::WARPS];   // end of the extension: per warp {score, k, i, -} and payload of its first maximum
    int2 lut[25];                       // {substitution score, is-match} for codes 0..4 x 0..4
    int red[S::WARPS];
};

// One DP cell (i,j) of anti-diagonal k on diagonal delta = i - j.  self = this diagonal's cell two anti-diagonals ago
// (updated in place), (uh, ud) = H and D of cell (i-1,j), (lh, li) = H and I of cell (i,j-1), both of the previous anti-diagonal.
// ok = cell lies inside both sequences (always true away from the sequence ends). hmax collects the maximum of the
// thread's cells on this anti-diagonal: every thread remembers the first cell (anti-diagonal, then row) that reached its own
// maximum, and the alignment end is picked among those records once, at the end of the extension, so an anti-diagonal only
// has to share ONE number (its maximum score) for the y-drop threshold.
template <typename P>
__device__ __forceinline__ void gp_cell(Cell<P>& self, int uh, int ud, P uhp, P udp, int lh, int li, P lhp, P lip, int2 sc, bool ok,
                                        int OE, int E, int thr, int& hmax) {
    // D: vertical gap state, I: horizontal gap state (ties prefer opening from H, as in the oracle)
    const int dopen = uh - OE, dext = ud - E;
    const bool dsel = dopen >= dext;
    const int nd = dsel ? dopen : dext; const P ndp = dsel ? uhp : udp;
    const int iopen = lh - OE, iext = li - E;
    const bool isel = iopen >= iext;
    const int ni = isel ? iopen : iext; const P nip = isel ? lhp : lip;
    const int mval = self.h + sc.x; const P mp = Pay<P>::step(self.hp, sc.y);
    // H = max(M, D, I) with ties M > D > I
    const bool pickm = mval >= nd && mval >= ni;
    const bool pickd = nd >= ni;
    const int nh = pickm ? mval : (pickd ? nd : ni);
    const P nhp = pickm ? mp : (pickd ? ndp : nip);
    const bool alive = ok && nh >= thr;
    self.h = alive ? nh : NEG_INF; self.d = alive ? nd : NEG_INF; self.i = alive ? ni : NEG_INF;
    self.hp = nhp; self.dp = ndp; self.ip = nip;
    hmax = max(hmax, self.h);
}

// One-sided y-drop extension by the whole CTA in diagonal-major coordinates. DIR=+1: cell (i,j) consumes T[ta+i-1],
// Q[qa+j-1]; DIR=-1: T[ta-i], Q[qa-j]. Diagonals delta = i-j live in a circular window of ND slots, slot = delta mod ND;
// thread t owns SLOTS consecutive slots for the whole extension, so every cell and all but one of its neighbours stay in
// registers. Per anti-diagonal a thread computes its cells of the right parity (independent of each other), reads ONE
// cell of a neighbouring thread from shared memory and publishes one; one __syncthreads per anti-diagonal, which shares
// a single number per warp (the anti-diagonal's maximum, for the y-drop threshold and the "all dead" test). Dead cells
// are -inf by value, so the window follows the alignment without bookkeeping: a slot that re-enters the band on another
// diagonal is already dead.
//
// The band is re-measured only every GP_EPOCH anti-diagonals: the exact range [lo, hi] of alive diagonals is taken at the
// end of an epoch, and because an alive cell needs an alive neighbour one anti-diagonal earlier (or itself two earlier),
// the alive range grows by at most one diagonal per side and step, so computing [lo - GP_EPOCH, hi + GP_EPOCH] for the
// whole next epoch covers every cell that can be alive (the extra cells evaluate to dead). Which threads and warps work,
// the window base and the sequence-end test are therefore constants of an epoch; warps outside the range only keep
// the barrier company.
// Returns false if the payload type cannot represent the result (16-bit columns overflowed): rerun with P = uint64_t.
constexpr int GP_EPOCH = 16;
template <typename P, typename S, int DIR>
__device__ bool ydrop_extend_cta(const GenomeView& T, const GenomeView& Q, uint32_t ta, uint32_t qa, int tn, int qn, int O, int E, int Y,
                                 GpShared<P, S>& sm, Ext& r, unsigned& ncell, int& err) {
    constexpr int GP_NT = S::NT, GP_SLOTS = S::SLOTS, GP_WARPS = S::WARPS, GP_DMASK = S::DMASK, GP_MAXBAND = S::MAXBAND;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint8_t* __restrict__ tcodes = T.codes;
    const uint8_t* __restrict__ qcodes = Q.codes;
    const int OE = O + E;
    r = Ext{0, 0, 0, 0, 0};
    Cell<P> st[GP_SLOTS];
#pragma unroll
    for (int s = 0; s < GP_SLOTS; s++) cell_dead(st[s]);
    if (tid == 0) st[0].h = 0;             // cell (0,0) on diagonal 0 = slot 0 of thread 0
    __syncthreads();                        // previous users of the buffers are done
    sm.left.put(tid, Edge<P>{st[0].h, NEG_INF, 0, 0});
    sm.up.put(tid, Edge<P>{NEG_INF, NEG_INF, 0, 0});
    __syncthreads();
    int best = 0, dead_steps = 0;
    int lo_e = 0, hi_e = 0;                     // exact alive diagonal range over the last two anti-diagonals
    int tbest = 0, tk = 0, ti = 0;              // this thread's first maximum; (0, 0, 0) = the anchor cell itself
    P tp = 0;
    bool narrow_ok = true, finished = false;
    const uint32_t kmax = (uint32_t)tn + (uint32_t)qn;
    uint32_t k = 1;
    while (!finished && k <= kmax) {
        if (sizeof(P) == 4 && k >= (uint32_t)GP_NARROW_ABORT_K) { narrow_ok = false; break; }
        if (hi_e - lo_e + 2 * GP_EPOCH > GP_MAXBAND) { err = 1; break; }
        // ---------------- constants of the epoch
        const int rlo = lo_e - GP_EPOCH, rhi = hi_e + GP_EPOCH;
        const int base = (rlo - 16) & ~(GP_SLOTS - 1);     // window base: a multiple of SLOTS so a thread's slots stay consecutive
        const int d0 = base + ((GP_SLOTS * tid - base) & GP_DMASK);  // diagonal of slot 0; slot s holds d0 + s
        const bool active = d0 + GP_SLOTS - 1 >= rlo && d0 <= rhi;
        const bool wactive = __any_sync(0xffffffffu, active);
        const uint32_t kend = min(kmax, k + (uint32_t)GP_EPOCH - 1u);
        // can any cell of the epoch lie beyond the end of a sequence? (uniform over the CTA)
        const bool edge = (((int)kend + rhi + GP_SLOTS + 1) >> 1) > tn || (((int)kend - rlo + GP_SLOTS + 1) >> 1) > qn;
        if (!wactive && lane == 0) { sm.smax[0][warp] = NEG_INF; sm.smax[1][warp] = NEG_INF; }
        for (; k <= kend; k++) {
            const int thr = best - Y;
            const int par = (int)(k & 1);
            int hmax = NEG_INF;
            if (active) {
                const int i0 = ((int)k + d0 + par) >> 1, j0 = ((int)k - d0 - par) >> 1;   // cell c of this thread: (i0 + c, j0 - c)
                int2 sc[GP_SLOTS / 2];
                bool ok[GP_SLOTS / 2];
                if (!edge) {
                    const uint8_t* tp_ = DIR > 0 ? tcodes + ta + i0 - 1 : tcodes + ta - i0;
                    const uint8_t* qp_ = DIR > 0 ? qcodes + qa + j0 - 1 : qcodes + qa - j0;
#pragma unroll
                    for (int c = 0; c < GP_SLOTS / 2; c++) {
                        const uint32_t tb = DIR > 0 ? tp_[c] : tp_[-c], qb = DIR > 0 ? qp_[-c] : qp_[c];
                        sc[c] = sm.lut[tb * 5 + qb]; ok[c] = true;
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < GP_SLOTS / 2; c++) {
                        const int i = i0 + c, j = j0 - c;
                        ok[c] = i >= 0 && j >= 0 && i <= tn && j <= qn;
                        const int ic_ = min(max(i, 1), max(tn, 1)), jc_ = min(max(j, 1), max(qn, 1));
                        const uint32_t tb = tcodes[DIR > 0 ? ta + (uint32_t)ic_ - 1u : ta - (uint32_t)ic_];
                        const uint32_t qb = qcodes[DIR > 0 ? qa + (uint32_t)jc_ - 1u : qa - (uint32_t)jc_];
                        sc[c] = sm.lut[tb * 5 + qb];
                    }
                }
                if (par) {
                    // odd anti-diagonal: odd slots; up = slot s-1 (own), left = slot s+1 (own, or slot 0 of thread t+1 for the last slot)
                    const Edge<P> fl = sm.left.get((tid + 1) & (GP_NT - 1));
#pragma unroll
                    for (int c = 0; c < GP_SLOTS / 2; c++) {
                        const int s = 2 * c + 1;
                        if (s + 1 < GP_SLOTS) {
                            const Cell<P>& L = st[s + 1 < GP_SLOTS ? s + 1 : s];
                            gp_cell<P>(st[s], st[s - 1].h, st[s - 1].d, st[s - 1].hp, st[s - 1].dp, L.h, L.i, L.hp, L.ip, sc[c], ok[c], OE, E, thr, hmax);
                        } else {
                            gp_cell<P>(st[s], st[s - 1].h, st[s - 1].d, st[s - 1].hp, st[s - 1].dp, fl.h, fl.x, fl.hp, fl.xp, sc[c], ok[c], OE, E, thr, hmax);
                        }
                    }
                    const Cell<P>& e = st[GP_SLOTS - 1];
                    sm.up.put(tid, Edge<P>{e.h, e.d, e.hp, e.dp});
                } else {
                    // even anti-diagonal: even slots; up = slot s-1 (own, or the last slot of thread t-1 for slot 0), left = slot s+1 (own)
                    const Edge<P> fu = sm.up.get((tid - 1) & (GP_NT - 1));
#pragma unroll
                    for (int c = 0; c < GP_SLOTS / 2; c++) {
                        const int s = 2 * c;
                        if (s > 0) {
                            const Cell<P>& U = st[s > 0 ? s - 1 : 0];
                            gp_cell<P>(st[s], U.h, U.d, U.hp, U.dp, st[s + 1].h, st[s + 1].i, st[s + 1].hp, st[s + 1].ip, sc[c], ok[c], OE, E, thr, hmax);
                        } else {
                            gp_cell<P>(st[s], fu.h, fu.x, fu.hp, fu.xp, st[s + 1].h, st[s + 1].i, st[s + 1].hp, st[s + 1].ip, sc[c], ok[c], OE, E, thr, hmax);
                        }
                    }
                    const Cell<P>& e = st[0];
                    sm.left.put(tid, Edge<P>{e.h, e.i, e.hp, e.ip});
                }
                ncell += GP_SLOTS / 2;
                if (hmax > tbest) {
                    // a new maximum of this thread (rare): its first cell in row order on this anti-diagonal. Strict > across
                    // anti-diagonals keeps the earliest one.
                    tbest = hmax; tk = (int)k;
                    bool found = false;
#pragma unroll
                    for (int c = 0; c < GP_SLOTS / 2; c++) {
                        const int hv = par ? st[2 * c + 1].h : st[2 * c].h;
                        if (!found && hv == hmax) { found = true; ti = i0 + c; tp = par ? st[2 * c + 1].hp : st[2 * c].hp; }
                    }
                }
            }
            if (wactive) {
                const int wmax = __reduce_max_sync(0xffffffffu, hmax);
                if (lane == 0) sm.smax[par][warp] = wmax;
            }
            __syncthreads();                    // the one barrier of this anti-diagonal
            int bmax = NEG_INF;
#pragma unroll
            for (int w = 0; w < GP_WARPS; w++) bmax = max(bmax, sm.smax[par][w]);
            best = max(best, bmax);
            dead_steps = bmax > NEG_INF ? 0 : dead_steps + 1;
            if (dead_steps >= 2) { finished = true; break; }   // two dead anti-diagonals in a row: nothing can revive
        }
        if (finished) break;
        // ---------------- end of the epoch: exact range of alive diagonals (every slot holds its diagonal's latest cell)
        int dlo = INT_MAX, dhi = INT_MIN;
        if (active) {
#pragma unroll
            for (int s = 0; s < GP_SLOTS; s++)
                if (st[s].h > NEG_INF) { dlo = min(dlo, d0 + s); dhi = max(dhi, d0 + s); }
        }
        const int wlo = __reduce_min_sync(0xffffffffu, dlo), whi = __reduce_max_sync(0xffffffffu, dhi);
        if (lane == 0) sm.rng[warp] = make_int2(wlo, whi);
        __syncthreads();
        int alo = INT_MAX, ahi = INT_MIN;
#pragma unroll
        for (int w = 0; w < GP_WARPS; w++) { const int2 q = sm.rng[w]; alo = min(alo, q.x); ahi = max(ahi, q.y); }
        if (ahi < alo) break;                   // nothing alive (the dead-step test catches this first)
        lo_e = alo; hi_e = ahi;
    }
    // the alignment end: among the threads whose own maximum equals the global one, the earliest anti-diagonal, then the
    // smallest row (each thread's record already is its first such cell)
    {
        __syncthreads();
        const bool cand = tbest == best;
        const int wk = __reduce_min_sync(0xffffffffu, cand ? tk : INT_MAX);
        const int wi = __reduce_min_sync(0xffffffffu, (cand && tk == wk) ? ti : INT_MAX);
        if (cand && tk == wk && ti == wi) {
            const uint64_t p64 = (uint64_t)tp;
            sm.fin[warp] = make_int4(best, wk, wi, 0); sm.finp[warp] = make_int2((int)(uint32_t)p64, (int)(uint32_t)(p64 >> 32));
        }
        if (wk == INT_MAX && lane == 0) sm.fin[warp] = make_int4(INT_MIN, INT_MAX, INT_MAX, 0);
        __syncthreads();
        int bk = INT_MAX, bi = INT_MAX; uint64_t bp = 0;
#pragma unroll
        for (int w = 0; w < GP_WARPS; w++) {
            const int4 f = sm.fin[w];
            if (f.x == best && (f.y < bk || (f.y == bk && f.z < bi))) {
                bk = f.y; bi = f.z; const int2 q = sm.finp[w];
                bp = (uint64_t)(uint32_t)q.x | ((uint64_t)(uint32_t)q.y << 32);
            }
        }
        r.score = best; r.di = bi; r.dj = bk - bi;
        r.nmatch = Pay<P>::nm((P)bp); r.ncols = Pay<P>::nc((P)bp);
        __syncthreads();
    }
    // 16-bit columns are exact iff the optimal path has fewer than 65536 columns, which min(di, dj) bounds from above
    if (sizeof(P) == 4 && min(r.di, r.dj) >= 65536) narrow_ok = false;
    return narrow_ok;
}

// Exact shortcut for the trivial alignment of a scaffold with itself. When target and query are the SAME N-free
// sequence and the anchor lies on the main diagonal, the y-drop DP has a closed form: every column (x,x) is a match
// scoring s(b,b) > 0, and any other path to an anti-diagonal k uses at most min(i,j) <= k/2 aligned columns, each
// worth at most s(b,b) of its row base, minus gap costs -- so the main-diagonal cell is the strict maximum of every
// even anti-diagonal, is never pruned, and the extension ends at the scaffold end with score = sum of s(b,b),
// matches = columns = length. (Proof in DESIGN.md; sequences with any non-ACGT base take the general DP.)
__device__ Ext selfdiag_extend_cta(const GenomeView& T, uint32_t p0, uint32_t p1, int* __restrict__ sm, int nwarps) {
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
    for (int w = 0; w < nwarps; w++) tot += sm[w];
    __syncthreads();
    const int n = (int)(p1 - p0);
    Ext r;
    r.score = 91 * (n - tot) + 100 * tot; r.di = n; r.dj = n; r.nmatch = n; r.ncols = n;
    return r;
}

// Anchors of all chain members at once (slot x = position in the (tile, score desc) order): centre of the best
// 31-column window of the HSP (first maximum), the whole HSP if it is shorter. One CTA per member; every thread slides a
// window over its own contiguous share of the window positions, then the CTA keeps the first maximum.
constexpr int AP_NT = 128;
constexpr int AP_LONG = 16384;       // HSPs with more windows than this are split over AP_PARTS CTAs (second kernel)
constexpr int AP_PARTS = 32;
constexpr int AP_LONG_SLOTS = 64;

// windows [w0, w1) of one HSP by one thread: best window sum and its first position (strict > keeps the first)
__device__ __forceinline__ void ap_scan(const GenomeView& T, const GenomeView& Q, uint32_t ts, uint32_t qs, int w0, int w1, int& best, int& bw) {
    // 32 column scores at a time from registers: words of 32 bases / 32 N flags starting at column c
    auto score_at = [](uint64_t wt, uint64_t wq, uint32_t an, int t) -> int {
        return ((an >> t) & 1u) ? SCORE_N : sub_lut3((uint32_t)(((wt >> (2 * t)) & 3) << 2) | (uint32_t)((wq >> (2 * t)) & 3));
    };
    best = INT_MIN; bw = INT_MAX;
    if (w0 >= w1) return;
    int sum = 0;
    {
        const uint64_t wt = window32(T.pk, ts + w0), wq = window32(Q.pk, qs + w0);
        const uint32_t an = nwindow32(T.nm, ts + w0) | nwindow32(Q.nm, qs + w0);
#pragma unroll
        for (int t = 0; t < 31; t++) sum += score_at(wt, wq, an, t);
    }
    best = sum; bw = w0;
    // window w = w0 + 1 + 32 b + t: column w + 30 enters, column w - 1 leaves
    for (int wb = w0 + 1; wb < w1; wb += 32) {
        const uint64_t et = window32(T.pk, ts + wb + 30), eq = window32(Q.pk, qs + wb + 30);
        const uint32_t en = nwindow32(T.nm, ts + wb + 30) | nwindow32(Q.nm, qs + wb + 30);
        const uint64_t lt = window32(T.pk, ts + wb - 1), lq = window32(Q.pk, qs + wb - 1);
        const uint32_t ln = nwindow32(T.nm, ts + wb - 1) | nwindow32(Q.nm, qs + wb - 1);
#pragma unroll
        for (int t = 0; t < 32; t++) {
            sum += score_at(et, eq, en, t) - score_at(lt, lq, ln, t);
            if (wb + t < w1 && sum > best) { best = sum; bw = wb + t; }
        }
    }
}
// first maximum over the CTA (highest sum, then lowest window); valid in thread 0
__device__ __forceinline__ void ap_reduce(int& best, int& bw, int* sh_best, int* sh_w) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wbest = __reduce_max_sync(0xffffffffu, best);
    const int wbw = __reduce_min_sync(0xffffffffu, best == wbest ? bw : INT_MAX);
    if (lane == 0) { sh_best[warp] = wbest; sh_w[warp] = wbw; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int b = sh_best[0], w = sh_w[0];
        for (int k = 1; k < AP_NT / 32; k++)
            if (sh_best[k] > b || (sh_best[k] == b && sh_w[k] < w)) { b = sh_best[k]; w = sh_w[k]; }
        best = b; bw = w;
    }
}
// (sum, window) packed so that "higher sum, then lower window" is plain unsigned max
__device__ __forceinline__ unsigned long long ap_pack(int sum, int w) {
    return ((unsigned long long)((uint32_t)sum ^ 0x80000000u) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)w);
}

__global__ void __launch_bounds__(AP_NT)
anchor_points_kernel(GenomeView T, GenomeView Q, const uint32_t* __restrict__ tile, const int32_t* __restrict__ hs1, const int32_t* __restrict__ hs2,
                     const int32_t* __restrict__ hlen, const uint32_t* __restrict__ order, uint32_t nmember,
                     int32_t* __restrict__ a1, int32_t* __restrict__ a2, uint32_t* __restrict__ long_list, uint32_t* __restrict__ nlong,
                     unsigned long long* __restrict__ long_key) {
    __shared__ int sh_best[AP_NT / 32], sh_w[AP_NT / 32];
    const int tid = threadIdx.x;
    const uint32_t x = blockIdx.x;
    const uint32_t g = order[x], tl = tile[g];
    const uint32_t ts = T.off[tl / (uint32_t)Q.nscaf] + (uint32_t)hs1[g], qs = Q.off[tl % (uint32_t)Q.nscaf] + (uint32_t)hs2[g];
    const int len = hlen[g];
    if (len <= 31) {
        if (tid == 0) { a1[x] = hs1[g] + len / 2; a2[x] = hs2[g] + len / 2; }
        return;
    }
    const int nwin = len - 30;                       // window w covers columns [w, w + 30]
    if (nwin > AP_LONG) {                            // split over several CTAs by anchor_points_long_kernel
        if (tid == 0) { const uint32_t k = atomicAdd(nlong, 1u); long_list[k] = x; long_key[k] = 0ull; }
        return;
    }
    const int chunk = (nwin + AP_NT - 1) / AP_NT;
    int best, bw;
    ap_scan(T, Q, ts, qs, tid * chunk, min(nwin, tid * chunk + chunk), best, bw);
    ap_reduce(best, bw, sh_best, sh_w);
    if (tid == 0) { a1[x] = hs1[g] + bw + 15; a2[x] = hs2[g] + bw + 15; }
}
// grid (AP_PARTS, AP_LONG_SLOTS): CTA (p, s) scans part p of the long HSPs s, s + AP_LONG_SLOTS, ...
__global__ void __launch_bounds__(AP_NT)
anchor_points_long_kernel(GenomeView T, GenomeView Q, const uint32_t* __restrict__ tile, const int32_t* __restrict__ hs1, const int32_t* __restrict__ hs2,
                          const int32_t* __restrict__ hlen, const uint32_t* __restrict__ order, const uint32_t* __restrict__ long_list,
                          const uint32_t* __restrict__ nlong, unsigned long long* __restrict__ long_key) {
    __shared__ int sh_best[AP_NT / 32], sh_w[AP_NT / 32];
    const int tid = threadIdx.x;
    const uint32_t n = *nlong;
    for (uint32_t k = blockIdx.y; k < n; k += gridDim.y) {
        const uint32_t x = long_list[k];
        const uint32_t g = order[x], tl = tile[g];
        const uint32_t ts = T.off[tl / (uint32_t)Q.nscaf] + (uint32_t)hs1[g], qs = Q.off[tl % (uint32_t)Q.nscaf] + (uint32_t)hs2[g];
        const int nwin = hlen[g] - 30;
        const int part = (nwin + AP_PARTS - 1) / AP_PARTS;
        const int p0 = (int)blockIdx.x * part, p1 = min(nwin, p0 + part);
        const int chunk = (max(p1 - p0, 0) + AP_NT - 1) / AP_NT;
        int best, bw;
        ap_scan(T, Q, ts, qs, p0 + tid * chunk, min(p1, p0 + tid * chunk + chunk), best, bw);
        ap_reduce(best, bw, sh_best, sh_w);
        if (tid == 0 && best != INT_MIN) atomicMax(&long_key[k], ap_pack(best, bw));
        __syncthreads();
    }
}
__global__ void __launch_bounds__(64)
anchor_points_finish_kernel(const int32_t* __restrict__ hs1, const int32_t* __restrict__ hs2, const uint32_t* __restrict__ order,
                            const uint32_t* __restrict__ long_list, const uint32_t* __restrict__ nlong,
                            const unsigned long long* __restrict__ long_key, int32_t* __restrict__ a1, int32_t* __restrict__ a2) {
    const uint32_t n = *nlong;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t x = long_list[k], g = order[x];
        const int w = (int)(0xFFFFFFFFu - (uint32_t)(long_key[k] & 0xFFFFFFFFull));
        a1[x] = hs1[g] + w + 15; a2[x] = hs2[g] + w + 15;
    }
}

// Per (slot, direction) status of an extension.
enum : uint8_t { GX_NONE = 0, GX_NARROW = 1, GX_NEED_WIDE = 2, GX_WIDE = 3 };

struct GpWork {
    // per slot x (position in the (tile, score desc) order of the chain members)
    const uint32_t* order; const int32_t *a1, *a2; const uint32_t* cluster_of;
    uint8_t* status;                 // [2 * nmember]: slot * 2 + direction (0 forward, 1 backward)
    int32_t *e_score, *e_di, *e_dj, *e_nm, *e_nc;   // [2 * nmember] one-sided extension results
    uint32_t* stamp;                 // per cluster: last round that scheduled one of its members
    uint32_t *items_narrow, *items_wide;   // work lists of this round: slot * 2 + direction
    uint32_t* counts;                // {narrow, wide}
};

constexpr int SEL_PEND = 64;         // pending (computed, not yet accepted) boxes a warp remembers while scheduling

// One warp per tile. Phase 1 replays the sequential rule as far as results exist; phase 2 schedules.
__global__ void __launch_bounds__(128)
gp_select_kernel(GpWork W, const uint32_t* __restrict__ tile, uint32_t nmember, const uint32_t* __restrict__ seg_start, uint32_t nseg,
                 uint32_t* __restrict__ resume, uint32_t* __restrict__ nkept_arr, uint32_t round, int gthr,
                 int32_t* __restrict__ o_s1, int32_t* __restrict__ o_e1, int32_t* __restrict__ o_s2, int32_t* __restrict__ o_e2,
                 int32_t* __restrict__ o_score, int32_t* __restrict__ o_nm, int32_t* __restrict__ o_nc, uint32_t* __restrict__ o_tile,
                 uint32_t* __restrict__ o_keep, unsigned long long* __restrict__ counters) {
    __shared__ int pend[128 / 32][SEL_PEND][4];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (seg >= nseg) return;
    const uint32_t a = seg_start[seg];
    const uint32_t b = (seg + 1 < nseg) ? seg_start[seg + 1] : nmember;
    uint32_t x = a + resume[seg];
    uint32_t nk = nkept_arr[seg];
    if (x >= b) return;
    unsigned long long accepted = 0;
    auto covered = [&](int p1, int p2) -> bool {                 // spec D5: bounding-box test against reported alignments
        int cov = 0;
        for (uint32_t kk = lane; kk < nk; kk += 32)
            cov |= (p1 >= o_s1[a + kk] && p1 < o_e1[a + kk] && p2 >= o_s2[a + kk] && p2 < o_e2[a + kk]) ? 1 : 0;
        return __any_sync(0xffffffffu, cov);
    };
    auto ready = [&](uint32_t y) -> bool {
        const uint8_t f = W.status[2 * y], r = W.status[2 * y + 1];
        return (f == GX_NARROW || f == GX_WIDE) && (r == GX_NARROW || r == GX_WIDE);
    };
    // ---------------- phase 1: resolve in best-first order while results exist
    for (; x < b; x++) {
        const int p1 = W.a1[x], p2 = W.a2[x];
        if (covered(p1, p2)) continue;
        if (!ready(x)) break;
        accepted++;
        const int score = W.e_score[2 * x] + W.e_score[2 * x + 1];
        if (score >= gthr) {
            if (lane == 0) {
                const uint32_t o = a + nk;
                o_s1[o] = p1 - W.e_di[2 * x + 1]; o_e1[o] = p1 + W.e_di[2 * x]; o_s2[o] = p2 - W.e_dj[2 * x + 1]; o_e2[o] = p2 + W.e_dj[2 * x];
                o_score[o] = score; o_nm[o] = W.e_nm[2 * x] + W.e_nm[2 * x + 1]; o_nc[o] = W.e_nc[2 * x] + W.e_nc[2 * x + 1];
                o_tile[o] = tile[W.order[x]]; o_keep[o] = 1;
            }
            nk++;
            __syncwarp();
        }
    }
    if (lane == 0) {
        resume[seg] = x - a; nkept_arr[seg] = nk;
        if (accepted) atomicAdd(&counters[CNT_ANCHORS], accepted);
    }
    __syncwarp();
    if (x >= b) return;
    // ---------------- phase 2: schedule. Boxes of results that wait for their turn count as covered FOR SCHEDULING only.
    int npend = 0;
    for (uint32_t y0 = x; y0 < b; y0 += 32) {
        const uint32_t y = y0 + lane;
        const bool have = y < b && ready(y);
        const uint32_t m = __ballot_sync(0xffffffffu, have);
        if (have) {
            const int slot = npend + __popc(m & ((1u << lane) - 1u));
            if (slot < SEL_PEND) {
                const int p1 = W.a1[y], p2 = W.a2[y];
                pend[wib][slot][0] = p1 - W.e_di[2 * y + 1]; pend[wib][slot][1] = p1 + W.e_di[2 * y];
                pend[wib][slot][2] = p2 - W.e_dj[2 * y + 1]; pend[wib][slot][3] = p2 + W.e_dj[2 * y];
            }
        }
        npend = min(SEL_PEND, npend + __popc(m));
    }
    __syncwarp();
    for (uint32_t y = x; y < b; y++) {
        const uint8_t sf = W.status[2 * y], sr = W.status[2 * y + 1];
        const bool is_ready = (sf == GX_NARROW || sf == GX_WIDE) && (sr == GX_NARROW || sr == GX_WIDE);
        if (is_ready) continue;
        const uint32_t cl = W.cluster_of[W.order[y]];
        if (y != x && W.stamp[cl] == round) continue;             // one speculative extension per cluster and round
        const int p1 = W.a1[y], p2 = W.a2[y];
        if (y != x) {
            if (covered(p1, p2)) continue;                        // will be skipped when its turn comes
            int cov = 0;
            for (int kk = lane; kk < npend; kk += 32)
                cov |= (p1 >= pend[wib][kk][0] && p1 < pend[wib][kk][1] && p2 >= pend[wib][kk][2] && p2 < pend[wib][kk][3]) ? 1 : 0;
            if (__any_sync(0xffffffffu, cov)) continue;           // probably inside an alignment that is waiting for its turn
        }
        if (lane < 2) {
            const uint8_t s = lane == 0 ? sf : sr;
            if (s == GX_NONE) {
                W.items_narrow[atomicAdd(&W.counts[0], 1u)] = 2 * y + lane;
                W.status[2 * y + lane] = GX_NARROW;
            } else if (s == GX_NEED_WIDE) {
                W.items_wide[atomicAdd(&W.counts[1], 1u)] = 2 * y + lane;
                W.status[2 * y + lane] = GX_WIDE;
            }
        }
        if (lane == 0) W.stamp[cl] = round;
        __syncwarp();
    }
}

// One CTA per work item = (anchor, direction). P = uint32_t: 16+16-bit payload (fast path); uint64_t: always exact.
template <typename P, typename S, int MINB>
__global__ void __launch_bounds__(S::NT, MINB)
gp_extend_kernel(GenomeView T, GenomeView Q, GpWork W, const uint32_t* __restrict__ items, const uint32_t* __restrict__ tile,
                 int O, int E, int Y, const int32_t* __restrict__ same_q, unsigned long long* __restrict__ counters) {
    __shared__ __align__(16) GpShared<P, S> sm;
    const int tid = threadIdx.x;
    if (tid < 25) {
        const int a = tid / 5, b = tid % 5;
        sm.lut[tid] = make_int2((a == 4 || b == 4) ? SCORE_N : c_sub[a * 4 + b], (a == b && a < 4) ? 1 : 0);
    }
    const uint32_t item = items[blockIdx.x];
    const uint32_t x = item >> 1, dir = item & 1u;
    const uint32_t tl = tile[W.order[x]];
    const uint32_t tsc = tl / (uint32_t)Q.nscaf, qsc = tl % (uint32_t)Q.nscaf;
    const uint32_t toff = T.off[tsc], qoff = Q.off[qsc];
    const int tlen = (int)T.len[tsc], qlen = (int)Q.len[qsc];
    const int a1 = W.a1[x], a2 = W.a2[x];
    const bool closed = same_q[tsc] == (int)qsc && a1 == a2 && T.nfree[tsc];
    Ext r;
    unsigned ncell = 0;
    int err = 0;
    bool ok = true;
    if (closed) {
        r = dir == 0 ? selfdiag_extend_cta(T, toff + a1, toff + tlen, sm.red, S::WARPS) : selfdiag_extend_cta(T, toff, toff + a1, sm.red, S::WARPS);
    } else if (dir == 0) {
        ok = ydrop_extend_cta<P, S, +1>(T, Q, toff + a1, qoff + a2, tlen - a1, qlen - a2, O, E, Y, sm, r, ncell, err);
    } else {
        ok = ydrop_extend_cta<P, S, -1>(T, Q, toff + a1, qoff + a2, a1, a2, O, E, Y, sm, r, ncell, err);
    }
    if (tid == 0) {
        W.e_score[item] = r.score; W.e_di[item] = r.di; W.e_dj[item] = r.dj; W.e_nm[item] = r.nmatch; W.e_nc[item] = r.ncols;
        if (!ok) W.status[item] = GX_NEED_WIDE;
        if (err) atomicAdd(&counters[CNT_ERR], 1ull);
    }
    ncell = __reduce_add_sync(0xffffffffu, ncell);
    if ((tid & 31) == 0 && ncell) atomicAdd(&counters[CNT_GAPPED_CELLS], (unsigned long long)ncell);
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
                   const int32_t* h_same_q, AlnSet& out, unsigned long long* counters) {
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
        ProfScope ps("gapped");
        // order the chain members: (tile, score desc), stable over the canonical order
        int tb = 1; while (tb < 33 && (((uint64_t)T.nscaf * (uint64_t)Q.nscaf) >> tb)) tb++;
        DevBuf<uint64_t> k0(n), k1(n);
        DevBuf<uint32_t> i0(n), i1(n), mflag(n), mrank(n), d_nmember(1);
        launch(anchor_keys_kernel, cdiv(n, 256), 256, 0, h.tile.get(), h.score.get(), in_chain.get(), n, k0.get(), i0.get(), mflag.get());
        const int w = radix_sort_bits<uint64_t, uint32_t>(k0.get(), k1.get(), i0.get(), i1.get(), n, 0, std::min(64, tb + 31));
        const uint64_t* skey = w ? k1.get() : k0.get();
        const uint32_t* order = w ? i1.get() : i0.get();
        DevBuf<uint32_t> flag(n), flag_off(n), seg_start(n), d_nseg(1);
        launch(anchor_heads_kernel, cdiv(n, 256), 256, 0, skey, n, flag.get());
        exclusive_scan_u32(flag.get(), flag_off.get(), n, d_nseg.get());
        launch(anchor_starts_kernel, cdiv(n, 256), 256, 0, flag.get(), flag_off.get(), n, seg_start.get());
        exclusive_scan_u32(mflag.get(), mrank.get(), n, d_nmember.get());
        uint32_t h_nseg = 0, h_nmember = 0;
        MB2_CUDA(cudaMemcpyAsync(&h_nseg, d_nseg.get(), sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
        MB2_CUDA(cudaMemcpyAsync(&h_nmember, d_nmember.get(), sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
        // same_q[t] = query scaffold that is the identical sequence as target scaffold t (or -1): enables the closed form
        std::vector<int32_t> same(T.nscaf, -1);
        for (int t = 0; t < T.nscaf; t++) {
            if (h_same_q) same[t] = h_same_q[t];
            else if (Q.fwd_src_id != 0 && Q.fwd_src_id == T.id && t < Q.nfwd) same[t] = t;
        }
        DevBuf<int32_t> d_same(T.nscaf);
        MB2_CUDA(cudaMemcpyAsync(d_same.get(), same.data(), T.nscaf * sizeof(int32_t), cudaMemcpyHostToDevice, cx.stream));
        MB2_CUDA(cudaStreamSynchronize(cx.stream));
        if (h_nmember) {
            const uint32_t nm = h_nmember;
            // speculation clusters along the chains
            DevBuf<uint32_t> mlist(nm), chead(nm), chead_off(nm), cluster_of(n), stamp(nm);
            launch(member_list_kernel, cdiv(n, 256), 256, 0, mflag.get(), mrank.get(), n, mlist.get());
            launch(cluster_heads_kernel, cdiv(nm, 256), 256, 0, mlist.get(), nm, h.tile.get(), h.s1.get(), h.s2.get(), h.len.get(), chead.get());
            exclusive_scan_u32(chead.get(), chead_off.get(), nm);
            launch(cluster_ids_kernel, cdiv(nm, 256), 256, 0, mlist.get(), nm, chead.get(), chead_off.get(), cluster_of.get());
            MB2_CUDA(cudaMemsetAsync(stamp.get(), 0, (size_t)nm * sizeof(uint32_t), cx.stream));
            // anchors, per-slot state
            DevBuf<int32_t> a1(nm), a2(nm), e_score(2 * (size_t)nm), e_di(2 * (size_t)nm), e_dj(2 * (size_t)nm), e_nm(2 * (size_t)nm), e_nc(2 * (size_t)nm);
            DevBuf<uint8_t> status(2 * (size_t)nm);
            DevBuf<uint32_t> items_n(2 * (size_t)nm), items_w(2 * (size_t)nm), counts(2), resume(h_nseg), nkept(h_nseg);
            MB2_CUDA(cudaMemsetAsync(status.get(), 0, 2 * (size_t)nm, cx.stream));
            MB2_CUDA(cudaMemsetAsync(resume.get(), 0, (size_t)h_nseg * sizeof(uint32_t), cx.stream));
            MB2_CUDA(cudaMemsetAsync(nkept.get(), 0, (size_t)h_nseg * sizeof(uint32_t), cx.stream));
            const GenomeView tv = view(T), qv = view(Q);
            DevBuf<uint32_t> long_list(nm), d_nlong(1);
            DevBuf<unsigned long long> long_key(nm);
            MB2_CUDA(cudaMemsetAsync(d_nlong.get(), 0, sizeof(uint32_t), cx.stream));
            launch(anchor_points_kernel, nm, AP_NT, 0, tv, qv, h.tile.get(), h.s1.get(), h.s2.get(), h.len.get(), order, nm,
                   a1.get(), a2.get(), long_list.get(), d_nlong.get(), long_key.get());
            launch(anchor_points_long_kernel, dim3(AP_PARTS, AP_LONG_SLOTS), AP_NT, 0, tv, qv, h.tile.get(), h.s1.get(), h.s2.get(), h.len.get(),
                   order, (const uint32_t*)long_list.get(), (const uint32_t*)d_nlong.get(), long_key.get());
            launch(anchor_points_finish_kernel, 4, 64, 0, h.s1.get(), h.s2.get(), order, (const uint32_t*)long_list.get(),
                   (const uint32_t*)d_nlong.get(), (const unsigned long long*)long_key.get(), a1.get(), a2.get());
            GpWork W;
            W.order = order; W.a1 = a1.get(); W.a2 = a2.get(); W.cluster_of = cluster_of.get(); W.status = status.get();
            W.e_score = e_score.get(); W.e_di = e_di.get(); W.e_dj = e_dj.get(); W.e_nm = e_nm.get(); W.e_nc = e_nc.get();
            W.stamp = stamp.get(); W.items_narrow = items_n.get(); W.items_wide = items_w.get(); W.counts = counts.get();
            static const bool dbg = getenv("MB2_GP_DEBUG") != nullptr;
            cudaEvent_t ev0 = nullptr, ev1 = nullptr;
            if (dbg) { cudaEventCreate(&ev0); cudaEventCreate(&ev1); }
            for (uint32_t round = 1;; round++) {
                MB2_CUDA(cudaMemsetAsync(counts.get(), 0, 2 * sizeof(uint32_t), cx.stream));
                launch(gp_select_kernel, cdiv((size_t)h_nseg * 32, 128), 128, 0, W, h.tile.get(), nm, seg_start.get(), h_nseg, resume.get(), nkept.get(),
                       round, p.gappedthresh, r_s1.get(), r_e1.get(), r_s2.get(), r_e2.get(), r_score.get(), r_nm.get(), r_nc.get(), r_tile.get(),
                       keep.get(), counters);
                uint32_t h_counts[2] = {0, 0};
                MB2_CUDA(cudaMemcpyAsync(h_counts, counts.get(), sizeof(h_counts), cudaMemcpyDeviceToHost, cx.stream));
                MB2_CUDA(cudaStreamSynchronize(cx.stream));
                if (dbg) {
                    float ms = 0;
                    if (round > 1) { cudaEventSynchronize(ev1); cudaEventElapsedTime(&ms, ev0, ev1); }
                    fprintf(stderr, "[gapped] members %u tiles %u | extend of round %u took %.3f ms | round %u schedules narrow %u wide %u\n", nm, h_nseg,
                            round - 1, ms, round, h_counts[0], h_counts[1]);
                    cudaEventRecord(ev0, cx.stream);
                }
                if (h_counts[0] == 0 && h_counts[1] == 0) break;
                auto go = [&](auto kern, uint32_t count, int nt, const uint32_t* items) {
                    launch(kern, count, nt, 0, tv, qv, W, items, h.tile.get(), p.gap_open, p.gap_extend, p.ydrop, d_same.get(), counters);
                };
                // 128 threads x 8 diagonals (window of 1024): measured best against 64 x 16 and 256 x 4; 5 CTAs per SM
                using S0 = GpShape<128, 8>;
                if (h_counts[0]) go(gp_extend_kernel<uint32_t, S0, 5>, h_counts[0], S0::NT, items_n.get());
                if (h_counts[1]) go(gp_extend_kernel<uint64_t, S0, 3>, h_counts[1], S0::NT, items_w.get());
                if (dbg) cudaEventRecord(ev1, cx.stream);
            }
            if (dbg) { cudaEventDestroy(ev0); cudaEventDestroy(ev1); }
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
