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
#include <cooperative_groups.h>

#include "primitives.cuh"
#include "seq.cuh"
#include "internal.cuh"

namespace mb2 {

constexpr int GP_NT = 256;             // threads per CTA (measured sweet spot between per-warp overhead and per-thread chain length)
constexpr int GP_WARPS = GP_NT / 32;
constexpr int GP_SLOTS = 4;            // diagonals per thread: two independent cells per thread per anti-diagonal
constexpr int GP_ND = GP_SLOTS * GP_NT;  // circular diagonal slots: thread t owns GP_SLOTS consecutive slots, slot = (i - j) & (GP_ND - 1)
constexpr int GP_DMASK = GP_ND - 1;
constexpr int GP_MAXBAND = GP_ND - 64;   // widest alive diagonal range the circular window can hold
constexpr int NEG_INF = INT_MIN / 4;
// shared memory: only the two boundary slots of every thread are exchanged: A = {h, hm, hc, d}, B = {i, dm, dc, im}, C = {ic}
constexpr int GP_SMEM_INTS = 2 * (4 + 4 + 1) * GP_NT;
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


struct CellState { int h, d, i, hm, hc, dm, dc, im, ic; };
// A dead cell holds NEG_INF in all three scores; arithmetic on it stays far below any threshold (thr >= -ydrop), so the
// recurrence needs no "is this predecessor alive" branches: a cell whose predecessors are all dead evaluates to ~NEG_INF,
// fails H >= thr and is reset to exactly NEG_INF. Out-of-range cells (i or j outside the sequences) are forced dead, which
// also makes the (i-1) / (j-1) predecessors of the first row and column dead without special cases.
__device__ __forceinline__ void cs_dead(CellState& c) { c.h = c.d = c.i = NEG_INF; c.hm = c.hc = c.dm = c.dc = c.im = c.ic = 0; }

// One DP cell (i,j) of anti-diagonal k on diagonal delta = i - j.  self = this diagonal's cell two anti-diagonals ago
// (updated in place), up = cell (i-1,j) and left = cell (i,j-1) of the previous anti-diagonal. Returns alive.
template <int DIR>
__device__ __forceinline__ bool gp_cell(const uint8_t* __restrict__ tcodes, const uint8_t* __restrict__ qcodes, uint32_t ta, uint32_t qa,
                                        int tn, int qn, int OE, int E, int thr, int k, int delta, CellState& self, const CellState& up,
                                        const CellState& left, const int* __restrict__ sub5, int& tmax, int& ti, int& thm, int& thc) {
    const int i2 = k + delta, j2 = k - delta;                    // 2i, 2j (same parity as k by construction)
    const int i = i2 >> 1, j = j2 >> 1;
    const bool inb = i2 >= 0 && j2 >= 0 && i <= tn && j <= qn;
    // substitution score of (i,j); clamped addresses keep the loads in bounds for cells that are forced dead anyway
    const int ic_ = min(max(i, 1), tn), jc_ = min(max(j, 1), qn);
    const uint32_t ct = DIR > 0 ? ta + (uint32_t)ic_ - 1u : ta - (uint32_t)ic_;
    const uint32_t cq = DIR > 0 ? qa + (uint32_t)jc_ - 1u : qa - (uint32_t)jc_;
    const uint32_t tb = tcodes[ct], qb = qcodes[cq];
    const int sc = sub5[tb * 5 + qb];
    // D: vertical gap state, I: horizontal gap state (ties prefer opening from H, as in the oracle)
    const int dopen = up.h - OE, dext = up.d - E;
    const bool dsel = dopen >= dext;
    const int nd = dsel ? dopen : dext, ndm = dsel ? up.hm : up.dm, ndc = dsel ? up.hc : up.dc;
    const int iopen = left.h - OE, iext = left.i - E;
    const bool isel = iopen >= iext;
    const int ni = isel ? iopen : iext, nim = isel ? left.hm : left.im, nic = isel ? left.hc : left.ic;
    const int mval = self.h + sc, mm = self.hm + ((tb == qb && tb < 4) ? 1 : 0), mc = self.hc + 1;
    // H = max(M, D, I) with ties M > D > I
    const bool pickm = mval >= nd && mval >= ni;
    const bool pickd = nd >= ni;
    const int nh = pickm ? mval : (pickd ? nd : ni);
    const int nhm = pickm ? mm : (pickd ? ndm : nim);
    const int nhc = pickm ? mc : (pickd ? ndc : nic);
    const bool alive = inb && nh >= thr;
    self.h = alive ? nh : NEG_INF; self.d = alive ? nd : NEG_INF; self.i = alive ? ni : NEG_INF;
    self.hm = nhm; self.hc = nhc; self.dm = ndm; self.dc = ndc; self.im = nim; self.ic = nic;
    if (alive && nh > tmax) { tmax = nh; ti = i; thm = nhm; thc = nhc; }
    return alive;
}

struct WarpRec { int wmax, wi, hm, hc, dlo, dhi, pad0, pad1; };

// One-sided y-drop extension by the whole CTA in diagonal-major coordinates. DIR=+1: cell (i,j) consumes T[ta+i-1],
// Q[qa+j-1]; DIR=-1: T[ta-i], Q[qa-j]. Diagonals delta = i-j live in a circular window of GP_ND slots, slot = delta mod
// GP_ND; thread t owns GP_SLOTS consecutive slots for the whole extension, so every cell and all but one of its neighbours stay in
// registers. Per anti-diagonal a thread computes its four cells of the right parity (independent of each other), reads ONE
// cell of a neighbouring thread from shared memory and publishes one; one __syncthreads per anti-diagonal. Dead cells are
// -inf by value, so the window follows the alignment without bookkeeping: a slot that re-enters the band on another
// diagonal is already dead.
template <int DIR>
__device__ Ext ydrop_extend_cta(const GenomeView& T, const GenomeView& Q, uint32_t ta, uint32_t qa, int tn, int qn, int O, int E, int Y,
                                int* __restrict__ sm, WarpRec (*rec)[GP_WARPS], const int* __restrict__ sub5,
                                unsigned long long& cells, int& err) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int4* A0 = reinterpret_cast<int4*>(sm);                 // published first slot of every thread
    int4* B0 = A0 + GP_NT;
    int4* AL = B0 + GP_NT;                                  // published last slot
    int4* BL = AL + GP_NT;
    int* C0 = reinterpret_cast<int*>(BL + GP_NT);
    int* CL = C0 + GP_NT;
    const uint8_t* __restrict__ tcodes = T.codes;
    const uint8_t* __restrict__ qcodes = Q.codes;
    const int OE = O + E;
    Ext r = {0, 0, 0, 0, 0};
    CellState st[GP_SLOTS];
#pragma unroll
    for (int s = 0; s < GP_SLOTS; s++) cs_dead(st[s]);
    if (tid == 0) st[0].h = 0;              // the cell (0,0): delta 0 -> slot 0 of thread 0
    __syncthreads();                        // previous users of the buffers are done
    A0[tid] = make_int4(st[0].h, 0, 0, NEG_INF); B0[tid] = make_int4(NEG_INF, 0, 0, 0); C0[tid] = 0;
    AL[tid] = make_int4(NEG_INF, 0, 0, NEG_INF); BL[tid] = make_int4(NEG_INF, 0, 0, 0); CL[tid] = 0;
    __syncthreads();
    int best = 0, dead_steps = 0;
    int lo1 = 0, hi1 = 0, lo2 = 1, hi2 = 0;     // alive diagonal ranges of anti-diagonals k-1 and k-2 (empty when lo > hi)
    unsigned ncell = 0;
    const uint32_t kmax = (uint32_t)tn + (uint32_t)qn;
    for (uint32_t k = 1; k <= kmax; k++) {
        // candidate diagonals of this anti-diagonal; the window base is a multiple of GP_SLOTS so a thread's slots stay consecutive
        int clo = INT_MAX, chi = INT_MIN;
        if (hi1 >= lo1) { clo = lo1 - 1; chi = hi1 + 1; }
        if (hi2 >= lo2) { clo = min(clo, lo2); chi = max(chi, hi2); }
        if (chi - clo > GP_MAXBAND) { err = 1; break; }
        const int base = (clo - 16) & ~(GP_SLOTS - 1);
        const int d0 = base + ((GP_SLOTS * tid - base) & GP_DMASK);  // diagonal of slot 0; slot s holds d0 + s
        const int thr = best - Y;
        int tmax = INT_MIN, ti = 0, thm = 0, thc = 0, dlo = INT_MAX, dhi = INT_MIN;
        const bool active = d0 + GP_SLOTS - 1 >= clo && d0 <= chi;
        if (active) {
            if (k & 1) {
                // odd anti-diagonal: odd slots; up = slot s-1 (own), left = slot s+1 (own, or slot 0 of thread t+1 for the last slot)
                const int f = (tid + 1) & (GP_NT - 1);
                CellState fl; const int4 fa = A0[f]; const int4 fb = B0[f];
                fl.h = fa.x; fl.hm = fa.y; fl.hc = fa.z; fl.d = fa.w; fl.i = fb.x; fl.dm = fb.y; fl.dc = fb.z; fl.im = fb.w; fl.ic = C0[f];
#pragma unroll
                for (int s = 1; s < GP_SLOTS; s += 2) {
                    const bool a = gp_cell<DIR>(tcodes, qcodes, ta, qa, tn, qn, OE, E, thr, (int)k, d0 + s, st[s], st[s - 1],
                                                s + 1 < GP_SLOTS ? st[s + 1 < GP_SLOTS ? s + 1 : s] : fl, sub5, tmax, ti, thm, thc);
                    if (a) { dlo = min(dlo, d0 + s); dhi = d0 + s; }
                }
                const CellState& e = st[GP_SLOTS - 1];
                AL[tid] = make_int4(e.h, e.hm, e.hc, e.d); BL[tid] = make_int4(e.i, e.dm, e.dc, e.im); CL[tid] = e.ic;
            } else {
                // even anti-diagonal: even slots; up = slot s-1 (own, or the last slot of thread t-1 for slot 0), left = slot s+1 (own)
                const int f = (tid - 1) & (GP_NT - 1);
                CellState fu; const int4 fa = AL[f]; const int4 fb = BL[f];
                fu.h = fa.x; fu.hm = fa.y; fu.hc = fa.z; fu.d = fa.w; fu.i = fb.x; fu.dm = fb.y; fu.dc = fb.z; fu.im = fb.w; fu.ic = CL[f];
#pragma unroll
                for (int s = 0; s < GP_SLOTS; s += 2) {
                    const bool a = gp_cell<DIR>(tcodes, qcodes, ta, qa, tn, qn, OE, E, thr, (int)k, d0 + s, st[s], s > 0 ? st[s > 0 ? s - 1 : 0] : fu,
                                                st[s + 1], sub5, tmax, ti, thm, thc);
                    if (a) { dlo = min(dlo, d0 + s); dhi = d0 + s; }
                }
                const CellState& e = st[0];
                A0[tid] = make_int4(e.h, e.hm, e.hc, e.d); B0[tid] = make_int4(e.i, e.dm, e.dc, e.im); C0[tid] = e.ic;
            }
            ncell += GP_SLOTS / 2;
        }
        const int par = (int)(k & 1);
        if (__any_sync(0xffffffffu, active)) {
            const int wmax = __reduce_max_sync(0xffffffffu, tmax);
            const int wlo = __reduce_min_sync(0xffffffffu, dlo), whi = __reduce_max_sync(0xffffffffu, dhi);
            int wi = 0, whm = 0, whc = 0;
            if (wmax > best) {
                wi = __reduce_min_sync(0xffffffffu, tmax == wmax ? ti : INT_MAX);
                const int src = __ffs(__ballot_sync(0xffffffffu, tmax == wmax && ti == wi)) - 1;
                whm = __shfl_sync(0xffffffffu, thm, src); whc = __shfl_sync(0xffffffffu, thc, src);
            }
            if (lane == 0) {
                int4* rp = reinterpret_cast<int4*>(&rec[par][warp]);
                rp[0] = make_int4(wmax, wi, whm, whc); rp[1] = make_int4(wlo, whi, 0, 0);
            }
        } else if (lane == 0) {
            int4* rp = reinterpret_cast<int4*>(&rec[par][warp]);
            rp[0] = make_int4(INT_MIN, 0, 0, 0); rp[1] = make_int4(INT_MAX, INT_MIN, 0, 0);
        }
        __syncthreads();                    // the one barrier of this anti-diagonal
        int4 q0 = make_int4(INT_MIN, INT_MAX, 0, 0), q1 = make_int4(INT_MAX, INT_MIN, 0, 0);
        if (lane < GP_WARPS) {
            const int4* rp = reinterpret_cast<const int4*>(&rec[par][lane]);
            q0 = rp[0]; q1 = rp[1];
        }
        const int bmax = __reduce_max_sync(0xffffffffu, q0.x);
        const int alo = __reduce_min_sync(0xffffffffu, q1.x), ahi = __reduce_max_sync(0xffffffffu, q1.y);
        if (bmax > best) {
            const int bi = __reduce_min_sync(0xffffffffu, q0.x == bmax ? q0.y : INT_MAX);
            const int src = __ffs(__ballot_sync(0xffffffffu, q0.x == bmax && q0.y == bi)) - 1;
            best = bmax; r.score = bmax; r.di = bi; r.dj = (int)k - bi;
            r.nmatch = __shfl_sync(0xffffffffu, q0.z, src); r.ncols = __shfl_sync(0xffffffffu, q0.w, src);
        }
        lo2 = lo1; hi2 = hi1; lo1 = alo; hi1 = ahi;          // empty ranges arrive as (INT_MAX, INT_MIN)
        if (hi1 < lo1) { lo1 = 1; hi1 = 0; }
        dead_steps = (hi1 < lo1) ? dead_steps + 1 : 0;
        if (dead_steps >= 2) break;         // two dead anti-diagonals in a row: nothing can revive
    }
    cells += ncell;                         // per-thread; summed by the caller
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

// One thread-block CLUSTER of two CTAs owns one tile: CTA 0 extends every anchor forwards, CTA 1 backwards, concurrently.
// They meet twice per anchor at a cluster barrier: once to swap their one-sided results through distributed shared memory,
// once after CTA 0 has published the alignment record that later anchors of the tile are tested against.
__global__ void __launch_bounds__(GP_NT, 3)
gapped_kernel(GenomeView T, GenomeView Q, const uint32_t* __restrict__ tile, const int32_t* __restrict__ hs1, const int32_t* __restrict__ hs2,
              const int32_t* __restrict__ hlen, const uint32_t* __restrict__ order, uint32_t nmember,
              const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ nseg_p,
              int O, int E, int Y, int gthr, const int32_t* __restrict__ same_q,
              int32_t* __restrict__ o_s1, int32_t* __restrict__ o_e1, int32_t* __restrict__ o_s2, int32_t* __restrict__ o_e2,
              int32_t* __restrict__ o_score, int32_t* __restrict__ o_nm, int32_t* __restrict__ o_nc, uint32_t* __restrict__ o_tile,
              uint32_t* __restrict__ o_keep, unsigned long long* __restrict__ counters) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned crank = cluster.block_rank();          // 0: forward extension, 1: backward extension
    extern __shared__ __align__(16) int gp_smem[];
    __shared__ __align__(16) WarpRec rec[2][GP_WARPS];
    __shared__ uint32_t sh_seg;
    __shared__ int sh_off;
    __shared__ Ext sh_ext;
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
        if (crank == 0 && tid == 0) sh_seg = (uint32_t)atomicAdd(&counters[CNT_WORK], 1ull);
        cluster.sync();
        const uint32_t seg = *cluster.map_shared_rank(&sh_seg, 0);
        cluster.sync();                                     // everyone has read it before CTA 0 may overwrite it
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
            if (__syncthreads_or(cov)) continue;                  // both CTAs take the same decision
            if (crank == 0) anchors++;
            Ext mine;
            const bool closed = same_q[tsc] == (int)qsc && a1 == a2 && T.nfree[tsc];
            if (crank == 0)
                mine = closed ? selfdiag_extend_cta(T, toff + a1, toff + tlen, gp_smem)
                              : ydrop_extend_cta<+1>(T, Q, toff + a1, qoff + a2, tlen - a1, qlen - a2, O, E, Y, gp_smem, rec, sub5, cells, err);
            else
                mine = closed ? selfdiag_extend_cta(T, toff, toff + a1, gp_smem)
                              : ydrop_extend_cta<-1>(T, Q, toff + a1, qoff + a2, a1, a2, O, E, Y, gp_smem, rec, sub5, cells, err);
            if (tid == 0) sh_ext = mine;
            cluster.sync();                                       // swap the one-sided results
            const Ext other = *cluster.map_shared_rank(&sh_ext, crank ^ 1);
            const Ext f = crank == 0 ? mine : other, r = crank == 0 ? other : mine;
            const int score = f.score + r.score;
            const bool keep = score >= gthr;
            if (keep && crank == 0 && tid == 0) {
                const uint32_t o = a + nkept;
                o_s1[o] = a1 - r.di; o_e1[o] = a1 + f.di; o_s2[o] = a2 - r.dj; o_e2[o] = a2 + f.dj;
                o_score[o] = score; o_nm[o] = f.nmatch + r.nmatch; o_nc[o] = f.ncols + r.ncols; o_tile[o] = tl; o_keep[o] = 1;
                __threadfence();
            }
            if (keep) nkept++;
            cluster.sync();                                       // record visible to both CTAs; sh_ext reusable
        }
    }
    if (cells) atomicAdd(&counters[CNT_GAPPED_CELLS], cells);
    if (tid == 0) {
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
            // same_q[t] = query scaffold that is the identical sequence as target scaffold t (or -1): enables the closed form
            std::vector<int32_t> same(T.nscaf, -1);
            for (int t = 0; t < T.nscaf; t++) {
                if (h_same_q) same[t] = h_same_q[t];
                else if (Q.fwd_src_id != 0 && Q.fwd_src_id == T.id && t < Q.nfwd) same[t] = t;
            }
            DevBuf<int32_t> d_same(T.nscaf);
            MB2_CUDA(cudaMemcpyAsync(d_same.get(), same.data(), T.nscaf * sizeof(int32_t), cudaMemcpyHostToDevice, cx.stream));
            MB2_CUDA(cudaStreamSynchronize(cx.stream));
            static bool attr_set = false;
            if (!attr_set) {
                MB2_CUDA(cudaFuncSetAttribute(gapped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GP_SMEM_BYTES));
                attr_set = true;
            }
            // one cluster (2 CTAs) per tile in flight; 3 CTAs of 256 threads per SM
            const unsigned nclusters = std::min<unsigned>((unsigned)cx.sm_count * 3 / 2, std::max<unsigned>(1u, h_nseg));
            MB2_CUDA(cudaMemsetAsync(counters + CNT_WORK, 0, sizeof(unsigned long long), cx.stream));
            ProfScope ps("gapped");
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2 * nclusters); cfg.blockDim = dim3(GP_NT); cfg.dynamicSmemBytes = GP_SMEM_BYTES; cfg.stream = cx.stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            const GenomeView tv = view(T), qv = view(Q);
            MB2_CUDA(cudaLaunchKernelEx(&cfg, gapped_kernel, tv, qv, (const uint32_t*)h.tile.get(), (const int32_t*)h.s1.get(),
                                        (const int32_t*)h.s2.get(), (const int32_t*)h.len.get(), (const uint32_t*)order, h_nmember,
                                        (const uint32_t*)seg_start.get(), (const uint32_t*)d_nseg.get(), p.gap_open, p.gap_extend, p.ydrop,
                                        p.gappedthresh, (const int32_t*)d_same.get(), r_s1.get(), r_e1.get(), r_s2.get(), r_e2.get(),
                                        r_score.get(), r_nm.get(), r_nc.get(), r_tile.get(), keep.get(), counters));
            cx.launches++;
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
