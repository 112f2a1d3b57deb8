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
#include "internal.cuh"

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

// dir = +1: columns ct0, ct0+1, ... ; returns best score and the exclusive end of the best prefix.
// dir = -1: columns ct0, ct0-1, ... ; returns best score and the inclusive start of the best prefix.
template <int DIR>
__device__ __forceinline__ void xdrop_side(const GenomeView& T, const GenomeView& Q, uint32_t ct0, uint32_t cq0, int X, int lane,
                                           int& best_out, uint32_t& bpos_out, unsigned long long& cells) {
    int run0 = 0, best = 0;
    uint32_t bpos = DIR > 0 ? ct0 : ct0 + 1;
    for (uint32_t base = 0;; base += 32) {
        const uint32_t ct = DIR > 0 ? ct0 + base + lane : ct0 - base - lane;
        const uint32_t cq = DIR > 0 ? cq0 + base + lane : cq0 - base - lane;
        const int s = col_score(T, Q, ct, cq);
        const int ps = warp_incl_sum_i(s, lane) + run0;
        const int pm = warp_incl_max_i(ps, lane);
        int pm_excl = __shfl_up_sync(0xffffffffu, pm, 1);
        if (lane == 0) pm_excl = INT_MIN;
        const int best_prev = max(best, pm_excl);
        const uint32_t tmask = __ballot_sync(0xffffffffu, ps < best_prev - X);
        const int nvalid = tmask ? __ffs(tmask) - 1 : 32;
        const int cand = lane < nvalid ? ps : INT_MIN;
        const int mx = __reduce_max_sync(0xffffffffu, cand);
        if (mx > best) {
            const uint32_t eq = __ballot_sync(0xffffffffu, cand == mx);
            const uint32_t first = __ffs(eq) - 1;
            bpos = DIR > 0 ? ct0 + base + first + 1 : ct0 - base - first;
            best = mx;
        }
        cells += tmask ? nvalid + 1 : 32;
        if (tmask) break;
        run0 = __shfl_sync(0xffffffffu, ps, 31);
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

__global__ void __launch_bounds__(256)
diag_heads_kernel(const uint64_t* __restrict__ surv, uint32_t n, uint32_t* __restrict__ flag) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    flag[k] = (k == 0 || (uint32_t)(surv[k] >> 32) != (uint32_t)(surv[k - 1] >> 32)) ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
diag_starts_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ flag_off, uint32_t n, uint32_t* __restrict__ seg_start) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (flag[k]) seg_start[flag_off[k]] = k;
}

__global__ void __launch_bounds__(128)
hsp_extend_kernel(GenomeView T, GenomeView Q, const uint64_t* __restrict__ surv, uint32_t nsurv,
                  const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ nseg_p, uint32_t diag_bias,
                  int X, int K, int entropy, uint32_t cap,
                  uint32_t* __restrict__ o_tile, int32_t* __restrict__ o_s1, int32_t* __restrict__ o_s2,
                  int32_t* __restrict__ o_len, int32_t* __restrict__ o_score, unsigned long long* __restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const uint32_t nseg = *nseg_p;
    unsigned long long cells = 0, extended = 0;
    for (;;) {
        uint32_t seg = 0;
        if (lane == 0) seg = (uint32_t)atomicAdd(&counters[CNT_WORK], 1ull);
        seg = __shfl_sync(0xffffffffu, seg, 0);
        if (seg >= nseg) break;
        const uint32_t a = seg_start[seg];
        const uint32_t b = (seg + 1 < nseg) ? seg_start[seg + 1] : nsurv;
        const uint32_t dg = (uint32_t)(surv[a] >> 32);          // i - j + bias
        uint32_t covered = 0;                                     // exclusive end (target coords) of the last kept HSP
        for (uint32_t x = a; x < b; x++) {
            const uint32_t j = (uint32_t)surv[x];
            const uint32_t i = j + dg - diag_bias;
            if (i + SEED_SPAN <= covered) continue;               // spec D2
            extended++;
            int best_r, best_l; uint32_t be, bs;
            xdrop_side<+1>(T, Q, i + SEED_SPAN, j + SEED_SPAN, X, lane, best_r, be, cells);
            xdrop_side<-1>(T, Q, i + SEED_SPAN - 1, j + SEED_SPAN - 1, X, lane, best_l, bs, cells);
            int score = best_r + best_l;
            if (score < K) continue;
            const uint32_t qs = bs - (i - j);
            if (entropy) {
                uint32_t cnt[4] = {0, 0, 0, 0};
                for (uint32_t c = bs; c < be; c += 32) {
                    const uint32_t ct = c + lane, cq = qs + (c - bs) + lane;
                    const bool in = ct < be;
                    const uint32_t tb = in ? base_at(T.pk, ct) : 0u, qb = in ? base_at(Q.pk, cq) : 1u;
                    const bool m = in && tb == qb && !(isn_at(T.nm, ct) | isn_at(Q.nm, cq));
#pragma unroll
                    for (uint32_t bb = 0; bb < 4; bb++) cnt[bb] += __popc(__ballot_sync(0xffffffffu, m && tb == bb));
                }
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
               HspSet& out, unsigned long long* counters) {
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
    launch(diag_heads_kernel, cdiv(nsurv, 256), 256, 0, sorted, nsurv, flag.get());
    exclusive_scan_u32(flag.get(), flag_off.get(), nsurv, d_nseg.get());
    launch(diag_starts_kernel, cdiv(nsurv, 256), 256, 0, flag.get(), flag_off.get(), nsurv, seg_start.get());

    DevBuf<uint32_t> r_tile(nsurv);
    DevBuf<int32_t> r_s1(nsurv), r_s2(nsurv), r_len(nsurv), r_score(nsurv);
    MB2_CUDA(cudaMemsetAsync(counters + CNT_WORK, 0, sizeof(unsigned long long), cx.stream));
    MB2_CUDA(cudaMemsetAsync(counters + CNT_HSPS, 0, sizeof(unsigned long long), cx.stream));
    {
        ProfScope ps("hsp_extend");
        const unsigned grid = (unsigned)cx.sm_count * 8;
        launch(hsp_extend_kernel, grid, 128, 0, view(T), view(Q), sorted, nsurv, seg_start.get(), d_nseg.get(), (uint32_t)Q.G,
               p.xdrop, p.hspthresh, p.entropy, nsurv, r_tile.get(), r_s1.get(), r_s2.get(), r_len.get(), r_score.get(), counters);
    }
    unsigned long long h_n = 0;
    MB2_CUDA(cudaMemcpyAsync(&h_n, counters + CNT_HSPS, sizeof(h_n), cudaMemcpyDeviceToHost, cx.stream));
    MB2_CUDA(cudaStreamSynchronize(cx.stream));
    MB2_REQUIRE(h_n <= nsurv, -5, "hsp stage: internal overflow");
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
