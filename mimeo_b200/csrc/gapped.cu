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

constexpr int GAP_W = 2048;            // circular capacity (cells) of one anti-diagonal buffer; band must stay below it
constexpr int GAP_WM = GAP_W - 1;
constexpr int NEG_INF = INT_MIN / 4;
constexpr int GAP_FIELDS = 9;          // h,d,i,hm,hc,dm,dc,im,ic
constexpr size_t GAP_SCRATCH_INTS = (size_t)3 * GAP_FIELDS * GAP_W;   // three rotating anti-diagonals

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

// One-sided y-drop extension. DIR=+1: cell (i,j) consumes T[ta + i - 1], Q[qa + j - 1]; DIR=-1: T[ta - i], Q[qa - j].
// ta/qa are padded-genome coordinates of the anchor, tn/qn the bases available in that direction.
template <int DIR>
__device__ Ext ydrop_extend(const GenomeView& T, const GenomeView& Q, uint32_t ta, uint32_t qa, int tn, int qn, int O, int E, int Y,
                            int* __restrict__ scratch, int lane, unsigned long long& cells, int& err) {
    int* buf[3] = {scratch, scratch + GAP_FIELDS * GAP_W, scratch + 2 * GAP_FIELDS * GAP_W};
    int* p2 = buf[0]; int* p1 = buf[1]; int* cur = buf[2];
#define F(b, f, i) (b)[(f) * GAP_W + ((i) & GAP_WM)]
    Ext r = {0, 0, 0, 0, 0};
    int lo2 = 0, hi2 = -1, lo1 = 0, hi1 = 0;
    if (lane == 0) {
        F(p1, 0, 0) = 0; F(p1, 1, 0) = NEG_INF; F(p1, 2, 0) = NEG_INF;
        for (int f = 3; f < GAP_FIELDS; f++) F(p1, f, 0) = 0;
    }
    __syncwarp();
    int best = 0;
    const long long kmax = (long long)tn + (long long)qn;
    for (long long k = 1; k <= kmax; k++) {
        int clo = INT_MAX, chi = INT_MIN;
        if (hi1 >= lo1) { clo = lo1; chi = hi1 + 1; }
        if (hi2 >= lo2) { clo = min(clo, lo2 + 1); chi = max(chi, hi2 + 1); }
        if (clo == INT_MAX) break;
        clo = max(clo, 0);
        if ((long long)clo < k - qn) clo = (int)(k - qn);
        if ((long long)chi > k) chi = (int)k;
        chi = min(chi, tn);
        int nlo = INT_MAX, nhi = INT_MIN;
        if (chi >= clo) {
            if (chi - clo + 1 > GAP_W - 2) { err = 1; break; }
            const int thr = best - Y;
            int curbest = best;
            for (int base = clo; base <= chi; base += 32) {
                const int i = base + lane;
                const int j = (int)(k - i);
                int h = NEG_INF, d = NEG_INF, ii = NEG_INF, hm = 0, hc = 0, dm = 0, dc = 0, im = 0, ic = 0;
                const bool in = i <= chi;
                if (in) {
                    if (i - 1 >= lo1 && i - 1 <= hi1) {           // up: (i-1, j) on k-1
                        const int uh = F(p1, 0, i - 1);
                        if (uh > NEG_INF) {
                            const int ud = F(p1, 1, i - 1);
                            const int open = uh - O - E, ext = ud > NEG_INF ? ud - E : NEG_INF;
                            if (open >= ext) { d = open; dm = F(p1, 3, i - 1); dc = F(p1, 4, i - 1); }
                            else { d = ext; dm = F(p1, 5, i - 1); dc = F(p1, 6, i - 1); }
                        }
                    }
                    if (i >= lo1 && i <= hi1) {                   // left: (i, j-1) on k-1
                        const int lh = F(p1, 0, i);
                        if (lh > NEG_INF) {
                            const int li = F(p1, 2, i);
                            const int open = lh - O - E, ext = li > NEG_INF ? li - E : NEG_INF;
                            if (open >= ext) { ii = open; im = F(p1, 3, i); ic = F(p1, 4, i); }
                            else { ii = ext; im = F(p1, 7, i); ic = F(p1, 8, i); }
                        }
                    }
                    int mval = NEG_INF, mm = 0, mc = 0;
                    if (i >= 1 && j >= 1 && i - 1 >= lo2 && i - 1 <= hi2) {   // diagonal: (i-1, j-1) on k-2
                        const int dh = F(p2, 0, i - 1);
                        if (dh > NEG_INF) {
                            const uint32_t ct = DIR > 0 ? ta + (uint32_t)i - 1u : ta - (uint32_t)i;
                            const uint32_t cq = DIR > 0 ? qa + (uint32_t)j - 1u : qa - (uint32_t)j;
                            const uint32_t an = isn_at(T.nm, ct) | isn_at(Q.nm, cq);
                            const uint32_t tb = base_at(T.pk, ct), qb = base_at(Q.pk, cq);
                            mval = dh + (an ? SCORE_N : sub_lut3((tb << 2) | qb));
                            mm = F(p2, 3, i - 1) + ((!an && tb == qb) ? 1 : 0);
                            mc = F(p2, 4, i - 1) + 1;
                        }
                    }
                    if (mval >= d && mval >= ii) { h = mval; hm = mm; hc = mc; }
                    else if (d >= ii) { h = d; hm = dm; hc = dc; }
                    else { h = ii; hm = im; hc = ic; }
                    if (h <= NEG_INF || h < thr) { h = NEG_INF; d = NEG_INF; ii = NEG_INF; }
                    F(cur, 0, i) = h; F(cur, 1, i) = d; F(cur, 2, i) = ii;
                    F(cur, 3, i) = hm; F(cur, 4, i) = hc; F(cur, 5, i) = dm; F(cur, 6, i) = dc; F(cur, 7, i) = im; F(cur, 8, i) = ic;
                }
                const bool alive = in && h > NEG_INF;
                const uint32_t amask = __ballot_sync(0xffffffffu, alive);
                if (amask) {
                    if (nlo == INT_MAX) nlo = base + __ffs(amask) - 1;
                    nhi = base + 31 - __clz(amask);
                    const int mx = __reduce_max_sync(0xffffffffu, alive ? h : INT_MIN);
                    if (mx > curbest) {
                        const int src = __ffs(__ballot_sync(0xffffffffu, alive && h == mx)) - 1;
                        curbest = mx;
                        r.score = mx; r.di = base + src; r.dj = (int)(k - (base + src));
                        r.nmatch = __shfl_sync(0xffffffffu, hm, src); r.ncols = __shfl_sync(0xffffffffu, hc, src);
                    }
                }
            }
            cells += (unsigned long long)(chi - clo + 1);
            best = curbest;
        }
        __syncwarp();
        int* tmp = p2; p2 = p1; p1 = cur; cur = tmp;
        lo2 = lo1; hi2 = hi1;
        if (nhi >= nlo && nlo != INT_MAX) { lo1 = nlo; hi1 = nhi; } else { lo1 = 0; hi1 = -1; }
    }
#undef F
    __syncwarp();
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

__global__ void __launch_bounds__(128)
gapped_kernel(GenomeView T, GenomeView Q, const uint32_t* __restrict__ tile, const int32_t* __restrict__ hs1, const int32_t* __restrict__ hs2,
              const int32_t* __restrict__ hlen, const uint32_t* __restrict__ order, uint32_t nmember,
              const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ nseg_p,
              int O, int E, int Y, int gthr, int* __restrict__ scratch_all,
              int32_t* __restrict__ o_s1, int32_t* __restrict__ o_e1, int32_t* __restrict__ o_s2, int32_t* __restrict__ o_e2,
              int32_t* __restrict__ o_score, int32_t* __restrict__ o_nm, int32_t* __restrict__ o_nc, uint32_t* __restrict__ o_tile,
              uint32_t* __restrict__ o_keep, unsigned long long* __restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int* scratch = scratch_all + (size_t)gwarp * GAP_SCRATCH_INTS;
    const uint32_t nseg = *nseg_p;
    unsigned long long cells = 0, anchors = 0;
    int err = 0;
    for (;;) {
        uint32_t seg = 0;
        if (lane == 0) seg = (uint32_t)atomicAdd(&counters[CNT_WORK], 1ull);
        seg = __shfl_sync(0xffffffffu, seg, 0);
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
            const int off = anchor_offset(T, Q, toff + s1, qoff + s2, len, lane);
            const int a1 = s1 + off, a2 = s2 + off;
            bool covered = false;                                 // spec D5: bounding-box test against reported alignments
            for (uint32_t kb = 0; kb < nkept; kb += 32) {
                const uint32_t kk = kb + lane;
                const bool c = kk < nkept && a1 >= o_s1[a + kk] && a1 < o_e1[a + kk] && a2 >= o_s2[a + kk] && a2 < o_e2[a + kk];
                if (__any_sync(0xffffffffu, c)) { covered = true; break; }
            }
            if (covered) continue;
            anchors++;
            const Ext f = ydrop_extend<+1>(T, Q, toff + a1, qoff + a2, tlen - a1, qlen - a2, O, E, Y, scratch, lane, cells, err);
            const Ext r = ydrop_extend<-1>(T, Q, toff + a1, qoff + a2, a1, a2, O, E, Y, scratch, lane, cells, err);
            const int score = f.score + r.score;
            if (score < gthr) continue;
            if (lane == 0) {
                const uint32_t o = a + nkept;
                o_s1[o] = a1 - r.di; o_e1[o] = a1 + f.di; o_s2[o] = a2 - r.dj; o_e2[o] = a2 + f.dj;
                o_score[o] = score; o_nm[o] = f.nmatch + r.nmatch; o_nc[o] = f.ncols + r.ncols; o_tile[o] = tl; o_keep[o] = 1;
            }
            nkept++;
            __syncwarp();
        }
    }
    if (lane == 0) {
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
            const unsigned blocks = std::min<unsigned>((unsigned)cx.sm_count * 4, std::max<unsigned>(1u, (h_nseg + 3) / 4));
            const unsigned nwarps = blocks * 4;
            DevBuf<int> scratch((size_t)nwarps * GAP_SCRATCH_INTS);
            MB2_CUDA(cudaMemsetAsync(counters + CNT_WORK, 0, sizeof(unsigned long long), cx.stream));
            ProfScope ps("gapped");
            launch(gapped_kernel, blocks, 128, 0, view(T), view(Q), h.tile.get(), h.s1.get(), h.s2.get(), h.len.get(), order, h_nmember,
                   seg_start.get(), d_nseg.get(), p.gap_open, p.gap_extend, p.ydrop, p.gappedthresh, scratch.get(),
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
