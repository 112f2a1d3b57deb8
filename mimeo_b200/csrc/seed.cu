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

// ---- load-balanced scan --------------------------------------------------------------------------------------
// A CTA takes SC_NQ consecutive query positions per round. Phase A: every thread looks up the 13 buckets of its own
// position and the CTA prefix-sums the 13*SC_NQ bucket sizes, which numbers the round's candidate hits 0..H-1.
// Phase B: hit number h goes to thread h mod SC_NT, which locates the owning bucket by binary search in the prefix
// array and runs the leader test and the bounded x-drop in a fixed, fully unrolled shape (32-column windows held in
// registers); a warp skips the remaining windows as soon as all of its lanes have terminated.
constexpr int SC_NT = 256;                 // threads per CTA
constexpr int SC_NQ = 256;                 // query positions per round
constexpr int SC_NPROBE = 13;
constexpr int SC_NDESC = SC_NQ * SC_NPROBE;

__global__ void __launch_bounds__(SC_NT)
seed_scan_kernel(GenomeView T, GenomeView Q, const uint32_t* __restrict__ off, const uint32_t* __restrict__ pos,
                 uint32_t q_lo, uint32_t q_n, int X, int K, int transition, uint32_t diag_bias,
                 uint64_t* __restrict__ surv, uint32_t surv_cap, unsigned long long* __restrict__ counters) {
    __shared__ uint32_t d_start[SC_NDESC + 1];     // exclusive prefix of bucket sizes (hit number of each bucket's first entry)
    __shared__ uint32_t d_b0[SC_NDESC];            // first index of each bucket in pos[]
    __shared__ uint32_t sh_scan[SC_NT / 32 + 1];
    __shared__ unsigned long long sh_stat[3];
    const int tid = threadIdx.x, lane = tid & 31;
    const int nprobe = transition ? SC_NPROBE : 1;
    if (tid < 3) sh_stat[tid] = 0;
    unsigned long long n_lead = 0, n_cells = 0, n_hits_cta = 0;
    const uint32_t nrounds = (q_n + SC_NQ - 1) / SC_NQ;
    for (uint32_t round = blockIdx.x; round < nrounds; round += gridDim.x) {
        // ---------------- phase A: bucket ranges of this thread's query position
        const uint32_t j = q_lo + round * SC_NQ + tid;
        uint32_t cnt[SC_NPROBE];
        uint32_t mine = 0;
        const bool jvalid = (round * SC_NQ + tid) < q_n && (nwindow32(Q.nm, j) & SEED_WINDOW_MASK19) == 0;
        uint32_t key = 0;
        if (jvalid) key = seed_key(window32(Q.pk, j));
#pragma unroll
        for (int pr = 0; pr < SC_NPROBE; pr++) {
            uint32_t c = 0, b0 = 0;
            if (jvalid && pr < nprobe) {
                const uint32_t kk = pr == 0 ? key : key ^ (2u << (2 * (pr - 1)));
                b0 = off[kk];
                c = off[kk + 1] - b0;
            }
            cnt[pr] = c;
            d_b0[tid * SC_NPROBE + pr] = b0;
            mine += c;
        }
        uint32_t total;
        uint32_t run = block_excl_scan<SC_NT>(mine, sh_scan, total);
#pragma unroll
        for (int pr = 0; pr < SC_NPROBE; pr++) { d_start[tid * SC_NPROBE + pr] = run; run += cnt[pr]; }
        if (tid == 0) d_start[SC_NDESC] = total;
        __syncthreads();
        n_hits_cta += (tid == 0) ? total : 0;
        // ---------------- phase B: hits of the round, strided over the CTA; fixed-shape, convergent extension
        for (uint32_t hbase = 0; hbase < total; hbase += SC_NT) {
            const uint32_t h = hbase + tid;
            bool live = h < total;                                   // lane has a hit that still needs an answer
            uint32_t hi = 0, hj = 0;
            if (live) {
                int lo = 0, hi_d = SC_NDESC;                          // bucket that owns hit h: last descriptor with d_start <= h
                while (hi_d - lo > 1) { const int mid = (lo + hi_d) >> 1; if (d_start[mid] <= h) lo = mid; else hi_d = mid; }
                hi = pos[d_b0[lo] + (h - d_start[lo])];
                hj = q_lo + round * SC_NQ + (uint32_t)(lo / SC_NPROBE);
                // spec D1: only run leaders are candidates
                const bool prev_seed = (nwindow32(Q.nm, hj - 1) & SEED_WINDOW_MASK19) == 0 &&
                                       (nwindow32(T.nm, hi - 1) & SEED_WINDOW_MASK19) == 0 &&
                                       seed_match(window32(T.pk, hi - 1), window32(Q.pk, hj - 1), transition != 0);
                if (prev_seed) live = false; else n_lead++;
            }
            // right of the seed: up to S1_RIGHT_BLOCKS windows of 32 columns; a warp stops early once all its lanes are done
            int runv = 0, best_r = 0;
            bool term_r = !live;
#pragma unroll 1
            for (int b = 0; b < S1_RIGHT_BLOCKS; b++) {
                if (__all_sync(0xffffffffu, term_r)) break;
                if (!term_r) {
                    const uint32_t ct = hi + SEED_SPAN + 32 * b, cq = hj + SEED_SPAN + 32 * b;
                    uint64_t wt = window32(T.pk, ct), wq = window32(Q.pk, cq);
                    uint32_t an = nwindow32(T.nm, ct) | nwindow32(Q.nm, cq);
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        const int sc = (an & 1u) ? SCORE_N : sub_lut((uint32_t)((wt & 3) << 2 | (wq & 3)));
                        wt >>= 2; wq >>= 2; an >>= 1;
                        if (!term_r) {
                            runv += sc; n_cells++;
                            if (runv > best_r) best_r = runv; else if (runv < best_r - X) term_r = true;
                        }
                    }
                }
            }
            const bool open_r = live && !term_r;
            // left, from the last seed column downwards: up to S1_LEFT_BLOCKS windows
            int best_l = 0;
            runv = 0;
            bool term_l = !live;
#pragma unroll 1
            for (int b = 0; b < S1_LEFT_BLOCKS; b++) {
                if (__all_sync(0xffffffffu, term_l)) break;
                if (!term_l) {
                    const uint32_t ct = hi + SEED_SPAN - 32 * (b + 1), cq = hj + SEED_SPAN - 32 * (b + 1);
                    uint64_t wt = window32(T.pk, ct), wq = window32(Q.pk, cq);
                    uint32_t an = nwindow32(T.nm, ct) | nwindow32(Q.nm, cq);
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        const int sc = (an >> 31) ? SCORE_N : sub_lut((uint32_t)(((wt >> 62) & 3) << 2 | ((wq >> 62) & 3)));
                        wt <<= 2; wq <<= 2; an <<= 1;
                        if (!term_l) {
                            runv += sc; n_cells++;
                            if (runv > best_l) best_l = runv; else if (runv < best_l - X) term_l = true;
                        }
                    }
                }
            }
            // dead iff both sides terminated inside their bounds with a total below K (spec D2: failed extensions leave no trace)
            const bool survivor = live && (open_r || !term_l || (best_r + best_l >= K));
            const uint32_t smask = __ballot_sync(0xffffffffu, survivor);
            if (smask) {
                unsigned long long basev = 0;
                if (lane == __ffs(smask) - 1) basev = atomicAdd(&counters[0], (unsigned long long)__popc(smask));
                basev = __shfl_sync(0xffffffffu, basev, __ffs(smask) - 1);
                if (survivor) {
                    const unsigned long long slot = basev + __popc(smask & ((1u << lane) - 1u));
                    if (slot < surv_cap) surv[slot] = ((uint64_t)(hi - hj + diag_bias) << 32) | hj;
                }
            }
        }
        __syncthreads();
    }
    // statistics
    if (n_lead) atomicAdd(&sh_stat[1], n_lead);
    if (n_cells) atomicAdd(&sh_stat[2], n_cells);
    if (n_hits_cta) atomicAdd(&sh_stat[0], n_hits_cta);
    __syncthreads();
    if (tid < 3 && sh_stat[tid]) atomicAdd(&counters[1 + tid], sh_stat[tid]);
}

// Enqueue the scan of query positions [q_lo, q_hi) against a built table. Survivors are appended to surv.
void seed_scan(const Genome& T, const Genome& Q, const SeedTable& tab, uint32_t q_lo, uint32_t q_hi, const AlignParams& p,
               uint64_t* surv, uint32_t surv_cap, unsigned long long* counters) {
    const uint32_t n = q_hi - q_lo;
    if (n == 0) return;
    ProfScope ps("seed_scan");
    const unsigned nrounds = cdiv(n, SC_NQ);
    const unsigned grid = std::min<unsigned>(nrounds, (unsigned)ctx().sm_count * 8);
    launch(seed_scan_kernel, grid, SC_NT, 0, view(T), view(Q), tab.off.get(), tab.pos.get(), q_lo, n, p.xdrop,
           p.hspthresh, p.transition, (uint32_t)Q.G, surv, surv_cap, counters);
}

}  // namespace mb2
