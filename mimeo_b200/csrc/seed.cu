// seed.cu -- kernel family (a): 12-of-19 spaced-seed position table and lookup, fused with the
// leader test and a bounded first-stage gap-free extension that discards the (overwhelmingly
// random) seed hits which provably cannot reach --hspthresh.
//
// Replaces LASTZ's seed position table + seed_hit_search + the start of its gap-free extension
// (reached from mimeo through wrappers.py:1025-1037 / 786-798 / 645-653).
//
//   seed_keys_kernel     every target position -> 24-bit key of its 19-mer window (or "invalid")
//   radix_sort_bits      positions sorted by key (stable, deterministic)
//   bucket_offsets_kernel  key -> [begin,end) in the sorted position list (2^24+1 offsets)
//   seed_scan_kernel     per query position: 13 probes (exact + 12 single-transition variants);
//                        per candidate target position: run-leader test (spec D1), then an exact
//                        x-drop extension bounded to 64 columns right / 96 left. Hits whose
//                        extension terminated inside the bounds with score < K are dead (they can
//                        never become an HSP and, by spec D2, leave no trace); everything else is
//                        a survivor and is appended as (diagonal, query position) for stage 2.
#include "primitives.cuh"
#include "seq.cuh"
#include "internal.cuh"

namespace mb2 {

constexpr uint32_t KEY_INVALID = 1u << 24;
constexpr int S1_RIGHT_BLOCKS = 2;   // 64 columns
constexpr int S1_LEFT_BLOCKS = 3;    // 96 columns (19 of them are the seed itself)

__global__ void __launch_bounds__(256)
seed_keys_kernel(GenomeView T, uint32_t p_lo, uint32_t n, uint32_t* __restrict__ keys, uint32_t* __restrict__ pos) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t p = p_lo + k;
    const uint32_t nw = nwindow32(T.nm, p) & SEED_WINDOW_MASK19;
    keys[k] = nw ? KEY_INVALID : seed_key(window32(T.pk, p));
    pos[k] = p;
}

// off[k] = first index in sorted keys with key >= k, for k in [0, 2^24]; sorted keys end with KEY_INVALID entries
__global__ void __launch_bounds__(256)
bucket_offsets_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ off) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx > n) return;
    const int64_t prev = idx == 0 ? -1 : (int64_t)keys[idx - 1];
    const int64_t cur = idx == n ? (int64_t)KEY_INVALID : (int64_t)min(keys[idx], KEY_INVALID);
    for (int64_t k = prev + 1; k <= cur; k++) off[k] = idx;
}

void build_seed_table(const Genome& T, uint32_t p_lo, uint32_t p_hi, SeedTable& tab) {
    const uint32_t n = p_hi - p_lo;
    ProfScope ps("seed_table_build");
    DevBuf<uint32_t> k0(n), k1(n), v0(n), v1(n);
    launch(seed_keys_kernel, cdiv(n, 256), 256, 0, view(T), p_lo, n, k0.get(), v0.get());
    const int w = radix_sort_bits<uint32_t, uint32_t>(k0.get(), k1.get(), v0.get(), v1.get(), n, 0, 25);
    tab.off.alloc((size_t)KEY_INVALID + 1);
    launch(bucket_offsets_kernel, cdiv((size_t)n + 1, 256), 256, 0, w ? k1.get() : k0.get(), n, tab.off.get());
    tab.pos = std::move(w ? v1 : v0);
    tab.p_lo = p_lo; tab.p_hi = p_hi;
}

// 16 HOXD70 scores as int8 in two 64-bit registers, index t*4+q
__device__ __forceinline__ int sub_lut(uint32_t idx) {
    // {91,-114,-31,-123, -114,100,-125,-31} , {-31,-125,100,-114, -123,-31,-114,91}
    const uint64_t lo = 0xE183648E85E18E5Bull, hi = 0x5B8EE1858E6483E1ull;
    const uint64_t v = (idx & 8) ? hi : lo;
    return (int)(int8_t)(v >> ((idx & 7) * 8));
}

struct S1Result { bool survivor; };

// exact gap-free x-drop extension of the seed hit (i,j), bounded; returns whether it may still yield an HSP
__device__ __forceinline__ bool stage1_survives(const GenomeView& T, const GenomeView& Q, uint32_t i, uint32_t j, int X, int K,
                                                unsigned long long& cells) {
    // ---- right of the seed
    int run = 0, best = 0;
    bool term_r = false;
#pragma unroll 1
    for (int b = 0; b < S1_RIGHT_BLOCKS && !term_r; b++) {
        const uint32_t ct = i + SEED_SPAN + 32 * b, cq = j + SEED_SPAN + 32 * b;
        uint64_t wt = window32(T.pk, ct), wq = window32(Q.pk, cq);
        uint32_t an = nwindow32(T.nm, ct) | nwindow32(Q.nm, cq);
#pragma unroll 8
        for (int c = 0; c < 32; c++) {
            const int s = (an & 1u) ? SCORE_N : sub_lut((uint32_t)((wt & 3) << 2 | (wq & 3)));
            wt >>= 2; wq >>= 2; an >>= 1;
            run += s;
            cells++;
            if (run > best) best = run;
            else if (run < best - X) { term_r = true; break; }
        }
    }
    // ---- left, from the last seed column downwards
    int runl = 0, bestl = 0;
    bool term_l = false;
#pragma unroll 1
    for (int b = 0; b < S1_LEFT_BLOCKS && !term_l; b++) {
        const uint32_t ct = i + SEED_SPAN - 32 * (b + 1), cq = j + SEED_SPAN - 32 * (b + 1);   // window = 32 columns ending at the previous start
        uint64_t wt = window32(T.pk, ct), wq = window32(Q.pk, cq);
        uint32_t an = nwindow32(T.nm, ct) | nwindow32(Q.nm, cq);
#pragma unroll 8
        for (int c = 31; c >= 0; c--) {
            const int s = ((an >> c) & 1u) ? SCORE_N : sub_lut((uint32_t)(((wt >> (2 * c)) & 3) << 2 | ((wq >> (2 * c)) & 3)));
            runl += s;
            cells++;
            if (runl > bestl) bestl = runl;
            else if (runl < bestl - X) { term_l = true; break; }
        }
    }
    if (term_r && term_l) return best + bestl >= K;
    return true;
}

__global__ void __launch_bounds__(256)
seed_scan_kernel(GenomeView T, GenomeView Q, const uint32_t* __restrict__ off, const uint32_t* __restrict__ pos,
                 uint32_t q_lo, uint32_t q_n, int X, int K, int transition, uint32_t diag_bias,
                 uint64_t* __restrict__ surv, uint32_t surv_cap, unsigned long long* __restrict__ counters) {
    // counters: [0] survivors, [1] seed hits, [2] leaders, [3] stage-1 cells
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long n_hits = 0, n_lead = 0, n_cells = 0;
    if (k < q_n) {
        const uint32_t j = q_lo + k;
        const uint32_t nq = nwindow32(Q.nm, j) & SEED_WINDOW_MASK19;
        if (nq == 0) {
            const uint64_t wq = window32(Q.pk, j);
            const uint32_t key = seed_key(wq);
            const uint64_t wq1 = window32(Q.pk, j - 1);
            const bool q1_clean = (nwindow32(Q.nm, j - 1) & SEED_WINDOW_MASK19) == 0;
            const int nprobe = transition ? 13 : 1;
            for (int pr = 0; pr < nprobe; pr++) {
                const uint32_t kk = pr == 0 ? key : key ^ (2u << (2 * (pr - 1)));
                const uint32_t b0 = off[kk], b1 = off[kk + 1];
                for (uint32_t x = b0; x < b1; x++) {
                    const uint32_t i = pos[x];
                    n_hits++;
                    // spec D1: only run leaders are candidates
                    if (q1_clean && (nwindow32(T.nm, i - 1) & SEED_WINDOW_MASK19) == 0 &&
                        seed_match(window32(T.pk, i - 1), wq1, transition != 0))
                        continue;
                    n_lead++;
                    if (!stage1_survives(T, Q, i, j, X, K, n_cells)) continue;
                    const unsigned long long slot = atomicAdd(&counters[0], 1ull);
                    if (slot < surv_cap) surv[slot] = ((uint64_t)(i - j + diag_bias) << 32) | j;
                }
            }
        }
    }
    // block-level reduction of the statistics
    __shared__ unsigned long long sh[3];
    if (threadIdx.x < 3) sh[threadIdx.x] = 0;
    __syncthreads();
    if (n_hits) atomicAdd(&sh[0], n_hits);
    if (n_lead) atomicAdd(&sh[1], n_lead);
    if (n_cells) atomicAdd(&sh[2], n_cells);
    __syncthreads();
    if (threadIdx.x < 3 && sh[threadIdx.x]) atomicAdd(&counters[1 + threadIdx.x], sh[threadIdx.x]);
}

// Enqueue the scan of query positions [q_lo, q_hi) against a built table. Survivors are appended to surv.
void seed_scan(const Genome& T, const Genome& Q, const SeedTable& tab, uint32_t q_lo, uint32_t q_hi, const AlignParams& p,
               uint64_t* surv, uint32_t surv_cap, unsigned long long* counters) {
    const uint32_t n = q_hi - q_lo;
    if (n == 0) return;
    ProfScope ps("seed_scan");
    launch(seed_scan_kernel, cdiv(n, 256), 256, 0, view(T), view(Q), tab.off.get(), tab.pos.get(), q_lo, n, p.xdrop,
           p.hspthresh, p.transition, (uint32_t)Q.G, surv, surv_cap, counters);
}

}  // namespace mb2
