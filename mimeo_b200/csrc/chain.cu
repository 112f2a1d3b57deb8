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

struct ChV { long long c; int idx; };
__device__ __forceinline__ bool chv_better(long long ac, int ai, long long bc, int bi) {
    return ac != bc ? ac > bc : ai < bi;
}

__global__ void __launch_bounds__(128)
chain_kernel(const int32_t* __restrict__ s1, const int32_t* __restrict__ s2, const int32_t* __restrict__ len,
             const int32_t* __restrict__ score, uint32_t n, const uint32_t* __restrict__ seg_start, const uint32_t* __restrict__ nseg_p,
             const uint32_t* __restrict__ ord1, const uint32_t* __restrict__ rank2, const int32_t* __restrict__ e2s,
             long long* __restrict__ bitC, int* __restrict__ bitI, long long* __restrict__ C, int* __restrict__ pred,
             uint8_t* __restrict__ in_chain, unsigned long long* __restrict__ work) {
    const int lane = threadIdx.x & 31;
    const uint32_t nseg = *nseg_p;
    for (;;) {
        uint32_t seg = 0;
        if (lane == 0) seg = (uint32_t)atomicAdd(work, 1ull);
        seg = __shfl_sync(0xffffffffu, seg, 0);
        if (seg >= nseg) break;
        const uint32_t a = seg_start[seg];
        const uint32_t b = (seg + 1 < nseg) ? seg_start[seg + 1] : n;
        const uint32_t m = b - a;
        for (uint32_t k = lane; k < m; k += 32) { bitC[a + k] = 0; bitI[a + k] = INT_MAX; in_chain[a + k] = 0; }
        __syncwarp();
        if (lane == 0) {
            uint32_t ins = 0;
            for (uint32_t x = 0; x < m; x++) {              // canonical order inside the tile
                const uint32_t g = a + x;
                const int xs1 = s1[g], xs2 = s2[g];
                while (ins < m) {
                    const uint32_t y = ord1[a + ins];        // global idx, increasing e1 inside the tile
                    if (s1[y] + len[y] > xs1) break;
                    const long long cy = C[y];
                    const int yi = (int)(y - a);
                    for (uint32_t pos = rank2[y] - a + 1; pos <= m; pos += pos & (~pos + 1)) {
                        if (chv_better(cy, yi, bitC[a + pos - 1], bitI[a + pos - 1])) { bitC[a + pos - 1] = cy; bitI[a + pos - 1] = yi; }
                    }
                    ins++;
                }
                // number of HSPs of the tile with e2 <= xs2
                uint32_t lo = 0, hi = m;
                while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (e2s[a + mid] <= xs2) lo = mid + 1; else hi = mid; }
                long long bc = 0; int bi = INT_MAX;
                for (uint32_t pos = lo; pos > 0; pos -= pos & (~pos + 1)) {
                    if (chv_better(bitC[a + pos - 1], bitI[a + pos - 1], bc, bi)) { bc = bitC[a + pos - 1]; bi = bitI[a + pos - 1]; }
                }
                pred[g] = bi == INT_MAX ? -1 : bi;
                C[g] = (long long)score[g] + (bi == INT_MAX ? 0 : bc);
            }
            uint32_t end = 0;
            for (uint32_t x = 1; x < m; x++) if (C[a + x] > C[a + end]) end = x;
            for (int k = (int)end; k >= 0; k = pred[a + k]) in_chain[a + k] = 1;
        }
        __syncwarp();
    }
}

void chain_hsps(const HspSet& h, int lb, int tb, DevBuf<uint8_t>& in_chain) {
    const uint32_t n = h.n;
    in_chain.alloc(n ? n : 1);
    if (n == 0) return;
    Ctx& cx = ctx();
    ProfScope ps("chain");
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

    DevBuf<long long> bitC(n), C(n);
    DevBuf<int> bitI(n), pred(n);
    DevBuf<unsigned long long> work(1);
    MB2_CUDA(cudaMemsetAsync(work.get(), 0, sizeof(unsigned long long), cx.stream));
    launch(chain_kernel, (unsigned)cx.sm_count * 8, 128, 0, h.s1.get(), h.s2.get(), h.len.get(), h.score.get(), n, seg_start.get(),
           d_nseg.get(), ord1, rank2.get(), e2s.get(), bitC.get(), bitI.get(), C.get(), pred.get(), in_chain.get(), work.get());
}

}  // namespace mb2
