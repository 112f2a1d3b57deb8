// chain.cu -- kernel family (b'): LASTZ --chain with zero penalties. Per (target scaffold, query
// scaffold, strand) tile, keep the maximum-total-score subset of HSPs that is strictly increasing in
// both sequences (spec and tie-breaks as in oracle/lastz_oracle.c lzo_chain()).
//
// Sparse dynamic programming, O(n log n) per tile: HSPs are visited in canonical order (increasing
// s1); an HSP a becomes visible once e1_a <= s1_b; visibility is a Fenwick tree over the rank of e2,
// holding (chain score, index) maxima with the deterministic order "larger score, then smaller
// index". The three orders needed (canonical, by e1, by e2) come from stable radix sorts, so one warp
// per tile only walks arrays; tiles are independent and scheduled dynamically.
#include "primitives.cuh"
#include "internal.cuh"

namespace mb2 {

__global__ void __launch_bounds__(256)
chain_keys_kernel(const uint32_t* __restrict__ tile, const int32_t* __restrict__ s, const int32_t* __restrict__ len, uint32_t n,
                  int lb, uint64_t* __restrict__ key, uint32_t* __restrict__ idx) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    key[k] = ((uint64_t)tile[k] << lb) | (uint64_t)(uint32_t)(s[k] + len[k]);
    idx[k] = k;
}
__global__ void __launch_bounds__(256)
tile_heads_kernel(const uint32_t* __restrict__ tile, uint32_t n, uint32_t* __restrict__ flag) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    flag[k] = (k == 0 || tile[k] != tile[k - 1]) ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
heads_scatter_kernel(const uint32_t* __restrict__ flag, const uint32_t* __restrict__ flag_off, uint32_t n, uint32_t* __restrict__ seg_start) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (flag[k]) seg_start[flag_off[k]] = k;
}
// rank2[global idx] = position (global) of that HSP in the e2-sorted order; e2s[pos] = its e2 value
__global__ void __launch_bounds__(256)
chain_rank_kernel(const uint32_t* __restrict__ ord2, const int32_t* __restrict__ s2, const int32_t* __restrict__ len, uint32_t n,
                  uint32_t* __restrict__ rank2, int32_t* __restrict__ e2s) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t x = ord2[k];
    rank2[x] = k;
    e2s[k] = s2[x] + len[x];
}

// (chain score, index) packed so that "larger score, then smaller index" is plain unsigned max:
// score in the top 40 bits, (0xFFFFFF - index) in the low 24. 0 = "no predecessor".
__device__ __forceinline__ unsigned long long chv_pack(long long c, uint32_t idx) {
    return ((unsigned long long)c << 24) | (unsigned long long)(0xFFFFFFu - idx);
}

// Per HSP g (parallel): upto1[g] = #HSPs of its tile with e1 <= s1[g] (how many are visible to it),
//                       cnt2[g]  = #HSPs of its tile with e2 <= s2[g] (Fenwick prefix it may query).
__global__ void __launch_bounds__(256)
chain_bounds_kernel(const int32_t* __restrict__ s1, const int32_t* __restrict__ s2, uint32_t n,
                    const uint32_t* __restrict__ flag, const uint32_t* __restrict__ flag_off,
                    const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ nseg_p,
                    const int32_t* __restrict__ e1s, const int32_t* __restrict__ e2s,
                    uint32_t* __restrict__ upto1, uint32_t* __restrict__ cnt2) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const uint32_t seg = flag_off[g] + flag[g] - 1;
    const uint32_t a = seg_start[seg];
    const uint32_t b = (seg + 1 < *nseg_p) ? seg_start[seg + 1] : n;
    const uint32_t m = b - a;
    uint32_t lo = 0, hi = m;
    const int x1 = s1[g];
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (e1s[a + mid] <= x1) lo = mid + 1; else hi = mid; }
    upto1[g] = lo;
    lo = 0; hi = m;
    const int x2 = s2[g];
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (e2s[a + mid] <= x2) lo = mid + 1; else hi = mid; }
    cnt2[g] = lo;
}

// e1s[k] = e1 of the k-th HSP in (tile, e1) order
__global__ void __launch_bounds__(256)
chain_e1s_kernel(const uint32_t* __restrict__ ord1, const int32_t* __restrict__ s1, const int32_t* __restrict__ len, uint32_t n,
                 int32_t* __restrict__ e1s) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t x = ord1[k];
    e1s[k] = s1[x] + len[x];
}

// One warp per tile, the HSPs of a tile in canonical order (increasing s1), a BATCH at a time. A batch is a run of
// consecutive HSPs none of which can be the predecessor of another: every member starts before the earliest end of the
// members before it. Their queries are then independent of each other:
//   1. every HSP that ends at or before the FIRST member's start becomes visible: lane l inserts the l-th pending one
//      (update chain p, p + lowbit(p), ... as 64-bit atomic maxima -- "larger score, then smaller index" is plain
//      unsigned max of the packed value -- so any number of inserts may run at once);
//   2. lane b answers member b's prefix-maximum query; the positions of a query chain (q, q - lowbit(q), ...) do not
//      depend on the stored values, so a lane issues its loads back to back, eight at a time;
//   3. HSPs that end inside the batch's window (after the first member's start, before the last one's) are visible to
//      the later members only; there are about as many of them as members, so they are handed round the warp by
//      shuffles and compared directly, and enter the tree with the next batch's inserts.
// The result is the same dynamic programme as the one-HSP-at-a-time walk (same maxima, same tie-breaks).
constexpr int CHAIN_SMEM_ENTRIES = 1024;     // tiles up to this many HSPs keep their tree in shared memory

template <bool SMEM>
__device__ __forceinline__ unsigned long long chain_load(const unsigned long long* p) {
    if (SMEM) return *reinterpret_cast<const volatile unsigned long long*>(p);
    return __ldcg(p);                                  // the atomics live in L2: never read the tree through L1
}

template <bool SMEM>
__device__ void chain_tile(const int32_t* __restrict__ score, const int32_t* __restrict__ s1, const int32_t* __restrict__ len,
                           uint32_t a, uint32_t m, const uint32_t* __restrict__ ord1, const uint32_t* __restrict__ rank2,
                           const uint32_t* __restrict__ upto1, const uint32_t* __restrict__ cnt2, unsigned long long* bit,
                           long long* __restrict__ C, int* __restrict__ pred, uint8_t* __restrict__ in_chain, int lane) {
    for (uint32_t k = lane; k < m; k += 32) { bit[k] = 0ull; in_chain[a + k] = 0; }
    if (!SMEM) __threadfence();
    __syncwarp();
    uint32_t ins = 0;
    long long cbest = -1; uint32_t end = 0;
    for (uint32_t x0 = 0; x0 < m;) {
        // ---- the batch: longest prefix of the next 32 HSPs in which nobody ends before a later member starts
        const uint32_t g = a + x0 + lane;
        const bool have = x0 + lane < m;
        const int my_s1 = have ? s1[g] : INT_MAX, my_e1 = have ? my_s1 + len[g] : INT_MAX;
        int pmin = my_e1;                                   // inclusive prefix minimum of e1
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pmin, d); if (lane >= d) pmin = min(pmin, t); }
        int before = __shfl_up_sync(0xffffffffu, pmin, 1);
        if (lane == 0) before = INT_MAX;
        const uint32_t bad = __ballot_sync(0xffffffffu, !have || my_s1 >= before);
        const int B = bad ? __ffs(bad) - 1 : 32;            // lane 0 is never bad while x0 < m
        const bool member = lane < B;
        const uint32_t ub = member ? upto1[g] : 0u;
        const uint32_t u0 = __shfl_sync(0xffffffffu, ub, 0), umax = __shfl_sync(0xffffffffu, ub, B - 1);
        // ---- 1. inserts visible to the whole batch
        for (; ins < u0; ins += 32) {
            const uint32_t t = ins + lane;
            if (t < u0) {
                const uint32_t y = ord1[a + t];
                const unsigned long long key = chv_pack(C[y], y - a);
                for (uint32_t pos = rank2[y] - a + 1; pos <= m; pos += pos & (0u - pos)) atomicMax(&bit[pos - 1], key);
            }
        }
        ins = u0;
        if (!SMEM) __threadfence();
        __syncwarp();
        // ---- 2. one query per lane
        unsigned long long v = 0ull;
        const uint32_t c2 = member ? cnt2[g] : 0u;
        {
            uint32_t q = c2;
            while (q > 0) {
                uint32_t qs[8];
#pragma unroll
                for (int t = 0; t < 8; t++) { qs[t] = q; q -= q & (0u - q); }       // q stays 0 once it reaches 0
                unsigned long long w[8];
#pragma unroll
                for (int t = 0; t < 8; t++) w[t] = qs[t] ? chain_load<SMEM>(&bit[qs[t] - 1]) : 0ull;
#pragma unroll
                for (int t = 0; t < 8; t++) v = w[t] > v ? w[t] : v;
            }
        }
        // ---- 3. HSPs that end inside the batch window: visible to the members that start after them
        for (uint32_t p0 = u0; p0 < umax; p0 += 32) {
            const uint32_t t = p0 + lane;
            unsigned long long pk = 0ull; uint32_t prank = 0xffffffffu;
            if (t < umax) { const uint32_t y = ord1[a + t]; pk = chv_pack(C[y], y - a); prank = rank2[y] - a; }
            const int cnt = (int)min(32u, umax - p0);
            for (int l = 0; l < cnt; l++) {
                const unsigned long long ok = __shfl_sync(0xffffffffu, pk, l);
                const uint32_t orank = __shfl_sync(0xffffffffu, prank, l);
                if (member && p0 + (uint32_t)l < ub && orank < c2 && ok > v) v = ok;
            }
        }
        // ---- 4. results of the batch
        long long c = -1;
        if (member) {
            c = (long long)score[g] + (v ? (long long)(v >> 24) : 0);
            C[g] = c;
            pred[g] = v ? (int)(0xFFFFFFu - (uint32_t)(v & 0xFFFFFFu)) : -1;
        }
        long long bc = c;                                    // best chain end so far: larger score, then smaller index
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { const long long o = __shfl_xor_sync(0xffffffffu, bc, d); bc = o > bc ? o : bc; }
        if (bc > cbest) {
            cbest = bc;
            end = x0 + (uint32_t)(__ffs(__ballot_sync(0xffffffffu, member && c == bc)) - 1);
        }
        __syncwarp();
        x0 += (uint32_t)B;
    }
    if (lane == 0)
        for (int k = (int)end; k >= 0; k = pred[a + k]) in_chain[a + k] = 1;
    __syncwarp();
}

// The same dynamic programme one HSP at a time (sparse tiles, where a batch would hold one or two members): every
// Fenwick-tree operation is done by the warp in ONE memory round trip -- the positions of an update chain and of a query
// chain do not depend on the stored values, so lane l takes the l-th position of the chain.
__device__ void chain_tile_seq(const int32_t* __restrict__ score, uint32_t a, uint32_t m, const uint32_t* __restrict__ ord1,
                               const uint32_t* __restrict__ rank2, const uint32_t* __restrict__ upto1, const uint32_t* __restrict__ cnt2,
                               unsigned long long* bit, long long* __restrict__ C, int* __restrict__ pred, uint8_t* __restrict__ in_chain, int lane) {
    for (uint32_t k = lane; k < m; k += 32) { bit[k] = 0ull; in_chain[a + k] = 0; }
    __syncwarp();
    uint32_t ins = 0;
    long long cbest = -1; uint32_t end = 0;
    for (uint32_t x = 0; x < m; x++) {              // canonical order inside the tile
        const uint32_t g = a + x;
        const uint32_t upto = upto1[g];
        for (; ins < upto; ins++) {                  // make every HSP with e1 <= s1_x visible
            const uint32_t y = ord1[a + ins];
            const unsigned long long key = chv_pack(C[y], y - a);
            uint32_t pos = rank2[y] - a + 1;
            for (int l = 0; l < lane && pos <= m; l++) pos += pos & (0u - pos);      // lane l takes the l-th node of the update chain
            if (pos <= m && bit[pos - 1] < key) bit[pos - 1] = key;
            __syncwarp();
        }
        uint32_t q = cnt2[g];
        for (int l = 0; l < lane && q > 0; l++) q -= q & (0u - q);                   // l-th node of the query chain
        unsigned long long v = q > 0 ? bit[q - 1] : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d); v = o > v ? o : v; }
        if (lane == 0) {
            const long long bc = (long long)(v >> 24);
            const long long c = (long long)score[g] + (v ? bc : 0);
            C[g] = c;
            pred[g] = v ? (int)(0xFFFFFFu - (uint32_t)(v & 0xFFFFFFu)) : -1;
            if (c > cbest) { cbest = c; end = x; }
        }
        __syncwarp();
    }
    if (lane == 0)
        for (int k = (int)end; k >= 0; k = pred[a + k]) in_chain[a + k] = 1;
    __syncwarp();
}

__global__ void __launch_bounds__(128)
chain_kernel(const int32_t* __restrict__ score, const int32_t* __restrict__ s1, const int32_t* __restrict__ len, uint32_t n,
             const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ nseg_p,
             const uint32_t* __restrict__ ord1, const uint32_t* __restrict__ rank2, const uint32_t* __restrict__ upto1,
             const uint32_t* __restrict__ cnt2, unsigned long long* __restrict__ bit_g, long long* __restrict__ C, int* __restrict__ pred,
             uint8_t* __restrict__ in_chain, unsigned long long* __restrict__ work, int* __restrict__ err, int mode) {
    __shared__ unsigned long long bit_s[4][CHAIN_SMEM_ENTRIES];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nseg = *nseg_p;
    for (;;) {
        uint32_t seg = 0;
        if (lane == 0) seg = (uint32_t)atomicAdd(work, 1ull);
        seg = __shfl_sync(0xffffffffu, seg, 0);
        if (seg >= nseg) break;
        const uint32_t a = seg_start[seg];
        const uint32_t b = (seg + 1 < nseg) ? seg_start[seg + 1] : n;
        const uint32_t m = b - a;
        if (m >= 0xFFFFFFu) { if (lane == 0) atomicOr(err, 1); continue; }
        // small tiles (tree in shared memory) are sparse: one HSP at a time; big tiles are dense: batches (mode: 0 = this
        // rule, 1 = always one at a time, 2 = always batches; for measurements)
        const bool small = m <= (uint32_t)CHAIN_SMEM_ENTRIES;
        const bool batched = mode == 2 || (mode == 0 && !small);
        if (!batched) chain_tile_seq(score, a, m, ord1, rank2, upto1, cnt2, small ? bit_s[warp] : bit_g + a, C, pred, in_chain, lane);
        else if (small) chain_tile<true>(score, s1, len, a, m, ord1, rank2, upto1, cnt2, bit_s[warp], C, pred, in_chain, lane);
        else chain_tile<false>(score, s1, len, a, m, ord1, rank2, upto1, cnt2, bit_g + a, C, pred, in_chain, lane);
    }
}

void chain_hsps(const HspSet& h, int lb, int tb, DevBuf<uint8_t>& in_chain) {
    const uint32_t n = h.n;
    in_chain.alloc(n ? n : 1);
    if (n == 0) return;
    Ctx& cx = ctx();
    ProfScope ps("chain");
    const int mode = getenv("MB2_CHAIN_MODE") ? atoi(getenv("MB2_CHAIN_MODE")) : 0;   // read per call: tests switch it
    MB2_REQUIRE(tb + lb <= 64, -3, "chain: key does not fit 64 bits");
    DevBuf<uint64_t> k0(n), k1(n);
    DevBuf<uint32_t> i0(n), i1(n), j0(n), j1(n);
    launch(chain_keys_kernel, cdiv(n, 256), 256, 0, h.tile.get(), h.s1.get(), h.len.get(), n, lb, k0.get(), i0.get());
    int w = radix_sort_bits<uint64_t, uint32_t>(k0.get(), k1.get(), i0.get(), i1.get(), n, 0, tb + lb);
    const uint32_t* ord1 = w ? i1.get() : i0.get();
    launch(chain_keys_kernel, cdiv(n, 256), 256, 0, h.tile.get(), h.s2.get(), h.len.get(), n, lb, k0.get(), j0.get());
    w = radix_sort_bits<uint64_t, uint32_t>(k0.get(), k1.get(), j0.get(), j1.get(), n, 0, tb + lb);
    const uint32_t* ord2 = w ? j1.get() : j0.get();
    DevBuf<uint32_t> rank2(n);
    DevBuf<int32_t> e2s(n);
    launch(chain_rank_kernel, cdiv(n, 256), 256, 0, ord2, h.s2.get(), h.len.get(), n, rank2.get(), e2s.get());

    DevBuf<uint32_t> flag(n), flag_off(n), seg_start(n), d_nseg(1);
    launch(tile_heads_kernel, cdiv(n, 256), 256, 0, h.tile.get(), n, flag.get());
    exclusive_scan_u32(flag.get(), flag_off.get(), n, d_nseg.get());
    launch(heads_scatter_kernel, cdiv(n, 256), 256, 0, flag.get(), flag_off.get(), n, seg_start.get());

    DevBuf<int32_t> e1s(n);
    launch(chain_e1s_kernel, cdiv(n, 256), 256, 0, ord1, h.s1.get(), h.len.get(), n, e1s.get());
    DevBuf<uint32_t> upto1(n), cnt2(n);
    launch(chain_bounds_kernel, cdiv(n, 256), 256, 0, h.s1.get(), h.s2.get(), n, flag.get(), flag_off.get(), seg_start.get(), d_nseg.get(),
           e1s.get(), e2s.get(), upto1.get(), cnt2.get());
    DevBuf<unsigned long long> bit(n), work(1);
    DevBuf<long long> C(n);
    DevBuf<int> pred(n), d_err(1);
    MB2_CUDA(cudaMemsetAsync(work.get(), 0, sizeof(unsigned long long), cx.stream));
    MB2_CUDA(cudaMemsetAsync(d_err.get(), 0, sizeof(int), cx.stream));
    launch(chain_kernel, (unsigned)cx.sm_count * 4, 128, 0, h.score.get(), h.s1.get(), h.len.get(), n, seg_start.get(), d_nseg.get(), ord1, rank2.get(), upto1.get(),
           cnt2.get(), bit.get(), C.get(), pred.get(), in_chain.get(), work.get(), d_err.get(), mode);
    int h_err = 0;
    MB2_CUDA(cudaMemcpyAsync(&h_err, d_err.get(), sizeof(int), cudaMemcpyDeviceToHost, cx.stream));
    MB2_CUDA(cudaStreamSynchronize(cx.stream));
    MB2_REQUIRE(h_err == 0, -3, "chain: a tile holds more than 2^24 HSPs");
}

}  // namespace mb2
