// primitives.cuh -- device-wide exclusive scan and LSD radix sort, hand-written for sm_100a.
//
// Both are global-atomic-free: the radix pass is histogram -> scan -> stable scatter with
// per-warp digit counters in shared memory (lanes of a warp with the same digit find each other
// through a shared-memory OR of lane bits, ranks follow lane order), so results are deterministic. Used by the coverage stage (bin interval events by genome tile) and by the
// alignment stage (group seed hits by diagonal, HSPs by scaffold pair).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace mb2 {

// ------------------------------------------------------------------------------------------
// exclusive scan (uint32), any n; out may alias in. If total != nullptr the grand total is
// written there (device pointer).
// ------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Block-wide exclusive scan of one value per thread; returns exclusive prefix, total via ref.
// NT = threads per block (multiple of 32, <= 1024). sh must hold NT/32 uint32.
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* sh, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan(v, lane);
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < NT / 32) ? sh[lane] : 0;
        uint32_t wi = warp_incl_scan(w, lane);
        if (lane < NT / 32) sh[lane] = wi - w;
        if (lane == NT / 32 - 1) sh[NT / 32] = wi;
    }
    __syncthreads();
    uint32_t res = incl - v + sh[warp];
    total = sh[NT / 32];
    __syncthreads();
    return res;
}

// Same, without the trailing barrier: the caller guarantees a block barrier before sh is written again.
template <int NT>
__device__ __forceinline__ uint32_t block_excl_scan_open(uint32_t v, uint32_t* sh, uint32_t& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = warp_incl_scan(v, lane);
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < NT / 32) ? sh[lane] : 0;
        uint32_t wi = warp_incl_scan(w, lane);
        if (lane < NT / 32) sh[lane] = wi - w;
        if (lane == NT / 32 - 1) sh[NT / 32] = wi;
    }
    __syncthreads();
    total = sh[NT / 32];
    return incl - v + sh[warp];
}

static __global__ void __launch_bounds__(SCAN_THREADS)
scan_block_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n, uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t sh[SCAN_THREADS / 32 + 1];
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        sum += v[i];
    }
    uint32_t total;
    uint32_t run = block_excl_scan<SCAN_THREADS>(sum, sh, total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
    if (threadIdx.x == 0 && block_sums) block_sums[blockIdx.x] = total;
}

static __global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(uint32_t* __restrict__ out, size_t n, const uint32_t* __restrict__ block_offsets) {
    const uint32_t off = block_offsets[blockIdx.x];
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) out[base + i] += off;
}

// Whole scan in ONE launch for small inputs (the histograms of the radix passes on a few hundred thousand keys): one CTA
// walks the tiles and carries the running total, instead of block scan + scan of block sums + add (three launches whose
// fixed cost dominates at these sizes).
constexpr size_t SCAN_SINGLE_MAX = 8 * 1024;
static __global__ void __launch_bounds__(SCAN_THREADS)
scan_single_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n, uint32_t* __restrict__ total_out) {
    __shared__ uint32_t sh[SCAN_THREADS / 32 + 1];
    uint32_t carry = 0;
    for (size_t tile = 0; tile < n; tile += SCAN_TILE) {
        const size_t base = tile + (size_t)threadIdx.x * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS];
        uint32_t sum = 0;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) {
            v[i] = (base + i < n) ? in[base + i] : 0u;
            sum += v[i];
        }
        uint32_t total;
        uint32_t run = block_excl_scan<SCAN_THREADS>(sum, sh, total) + carry;
#pragma unroll
        for (int i = 0; i < SCAN_ITEMS; i++) {
            if (base + i < n) out[base + i] = run;
            run += v[i];
        }
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

static __global__ void scan_total_kernel(const uint32_t* __restrict__ last_excl, const uint32_t* __restrict__ last_in, uint32_t* __restrict__ total) {
    *total = *last_excl + *last_in;
}

// sums_scratch (optional, cdiv(n, SCAN_TILE) words): block sums go there instead of a fresh allocation.
inline void exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* total = nullptr, uint32_t* sums_scratch = nullptr) {
    if (n == 0) {
        if (total) MB2_CUDA(cudaMemsetAsync(total, 0, sizeof(uint32_t), ctx().stream));
        return;
    }
    if (n <= SCAN_SINGLE_MAX) {
        launch(scan_single_kernel, 1, SCAN_THREADS, 0, in, out, n, total);
        return;
    }
    // total must be derived before `in` is overwritten when aliasing: keep the last input element.
    DevBuf<uint32_t> last_in;
    if (total) {
        last_in.alloc(1);
        MB2_CUDA(cudaMemcpyAsync(last_in.get(), in + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx().stream));
    }
    const unsigned nb = cdiv(n, SCAN_TILE);
    if (nb == 1) {
        launch(scan_block_kernel, 1, SCAN_THREADS, 0, in, out, n, (uint32_t*)nullptr);
    } else {
        DevBuf<uint32_t> sums_own;
        if (!sums_scratch) sums_own.alloc(nb);
        uint32_t* const sums = sums_scratch ? sums_scratch : sums_own.get();
        launch(scan_block_kernel, nb, SCAN_THREADS, 0, in, out, n, sums);
        exclusive_scan_u32(sums, sums, nb, nullptr);
        launch(scan_add_kernel, nb, SCAN_THREADS, 0, out, n, sums);
    }
    if (total) launch(scan_total_kernel, 1, 1, 0, out + (n - 1), last_in.get(), total);
}

// ------------------------------------------------------------------------------------------
// LSD radix sort on a bit range of the key. K = uint32_t or uint64_t, V = payload type
// (use NoVal for keys only). Stable. n < 2^32.
// ------------------------------------------------------------------------------------------
struct NoVal {};

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 4096 keys per CTA, 512 contiguous keys per warp

// One launch serves up to two independent key arrays of the same length (blockIdx.y picks the array): the coverage stage
// bins its +1 and -1 event arrays together, which halves the launches and doubles the CTAs in flight per launch.
// hist layout: [array][digit][CTA].
template <typename K>
__global__ void __launch_bounds__(RS_THREADS)
radix_hist_kernel(const K* __restrict__ in0, const K* __restrict__ in1, uint32_t n, int shift, int bits, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[RS_WARPS][256];   // per-warp counters: no contention between warps
    const K* __restrict__ keys = blockIdx.y ? in1 : in0;
    const int nbins = 1 << bits;
    const uint32_t mask = nbins - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&sh[0][0])[i] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * RS_TILE;
    uint32_t* my = sh[warp];
    if (n - base >= (uint32_t)RS_TILE) {
        K key[RS_ITEMS];
#pragma unroll
        for (int i = 0; i < RS_ITEMS; i++) key[i] = keys[base + i * RS_THREADS + threadIdx.x];   // all loads in flight at once
#pragma unroll
        for (int i = 0; i < RS_ITEMS; i++) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
            int same;
            __match_all_sync(0xffffffffu, d, &same);
            if (same) {                       // (nearly) sorted input: one update per warp instead of 32 colliding ones
                if (lane == 0) atomicAdd(&my[d], 32u);
            } else {
                atomicAdd(&my[d], 1u);
            }
        }
    } else {
        for (int i = 0; i < RS_ITEMS; i++) {
            const uint32_t idx = base + i * RS_THREADS + threadIdx.x;
            if (idx < n) atomicAdd(&my[(uint32_t)(keys[idx] >> shift) & mask], 1u);
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < nbins; d += RS_THREADS) {
        uint32_t c = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) c += sh[w][d];
        hist[((size_t)blockIdx.y * nbins + d) * gridDim.x + blockIdx.x] = c;
    }
}

// Lanes of `act` holding the same (at most 8-bit) digit. Not __match_any_sync: MATCH.ANY issues on the ADU pipe, which ncu
// showed 55-60 % busy and the limiter of the scatter kernel (profiles/r1_cov_v18_ncu_summary.txt); votes are cheap, and a
// warp of (nearly) sorted keys takes the one-vote exit. Used for the one partial tile of a sort; full tiles use the
// shared-memory OR in radix_scatter_kernel (an eighth of the instructions on unsorted keys).
__device__ __forceinline__ uint32_t match_digit(uint32_t act, uint32_t d) {
    const uint32_t d0 = __shfl_sync(act, d, __ffs(act) - 1);
    if (__all_sync(act, d == d0)) return act;
    uint32_t peers = act;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const bool bit = (d >> b) & 1u;
        const uint32_t bal = __ballot_sync(act, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

template <typename K, typename V, bool HAS_V>
__global__ void __launch_bounds__(RS_THREADS, sizeof(K) == 4 ? 4 : 3)
radix_scatter_kernel(const K* __restrict__ in0, const K* __restrict__ in1, K* __restrict__ out0, K* __restrict__ out1,
                     const V* __restrict__ vals_in, V* __restrict__ vals_out, uint32_t n, int shift, int bits,
                     const uint32_t* __restrict__ hist_scanned) {
    // Keys-only sorts stage the tile in shared memory in digit order and write it out in runs of consecutive addresses
    // (a warp's direct stores would hit 32 different sectors on unsorted input; ncu: same time with a third of the
    // instructions, i.e. bound by scattered L2 writes). With a payload the keys and values go straight to their slots.
    constexpr bool STAGE = !HAS_V;
    constexpr int MATCH_BYTES = RS_WARPS * 256 * 4, STAGE_BYTES = STAGE ? RS_TILE * (int)sizeof(K) : 0;
    __shared__ uint32_t wcount[RS_WARPS][256];   // per-warp running digit counters, then output bases
    __shared__ __align__(16) unsigned char raw[MATCH_BYTES > STAGE_BYTES ? MATCH_BYTES : STAGE_BYTES];
    uint32_t (*wmatch)[256] = reinterpret_cast<uint32_t (*)[256]>(raw);   // lane masks of the digit groups of the current round
    K* const stage = reinterpret_cast<K*>(raw);                           // later: the tile in digit order
    __shared__ uint32_t gofs[256];
    __shared__ uint32_t sh_scan[RS_THREADS / 32 + 1];
    const K* __restrict__ keys_in = blockIdx.y ? in1 : in0;
    K* __restrict__ keys_out = blockIdx.y ? out1 : out0;
    const int nbins = 1 << bits;
    const uint32_t mask = nbins - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) { (&wcount[0][0])[i] = 0; (&wmatch[0][0])[i] = 0; }
    static_assert(RS_THREADS >= 256, "one digit per thread in the staged write-out");

    const uint32_t wbase = blockIdx.x * RS_TILE + warp * (RS_ITEMS * 32);
    K key[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {   // every load of the thread in flight before the ranking rounds (their warp
        const uint32_t idx = wbase + r * 32 + lane;   // barriers would otherwise serialise 16 global-memory latencies)
        key[r] = idx < n ? keys_in[idx] : (K)0;
    }
    __syncthreads();
    uint16_t rank[RS_ITEMS];
    const uint32_t lt = (1u << lane) - 1u;
    if (n - blockIdx.x * (uint32_t)RS_TILE >= (uint32_t)RS_TILE) {
        // Full tile (all but the last CTA): every lane holds a key, all warp primitives take the full mask.
        // Lanes with the same digit find each other through a shared-memory OR of lane bits (one atomic, one load)
        // instead of one vote per digit bit; a warp whose 32 keys share the digit (sorted input) skips even that.
#pragma unroll
        for (int r = 0; r < RS_ITEMS; r++) {
            const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
            const bool uniform = __all_sync(0xffffffffu, d == __shfl_sync(0xffffffffu, d, 0));
            uint32_t peers = 0xffffffffu;
            if (!uniform) {
                atomicOr(&wmatch[warp][d], 1u << lane);
                __syncwarp();
                peers = wmatch[warp][d];
            }
            const uint32_t prev = wcount[warp][d];
            __syncwarp();                          // every peer has read the mask and the count
            if ((peers & lt) == 0) {               // lowest lane of the group
                wcount[warp][d] = prev + __popc(peers);
                if (!uniform) wmatch[warp][d] = 0;   // clean for the next round
            }
            __syncwarp();
            rank[r] = (uint16_t)(prev + __popc(peers & lt));
        }
    } else {
#pragma unroll
        for (int r = 0; r < RS_ITEMS; r++) {
            const uint32_t idx = wbase + r * 32 + lane;
            const bool valid = idx < n;
            const uint32_t act = __ballot_sync(0xffffffffu, valid);
            uint32_t rk = 0;
            if (valid) {
                const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
                const uint32_t peers = match_digit(act, d);
                const uint32_t prev = wcount[warp][d];
                __syncwarp(act);
                if ((peers & lt) == 0) wcount[warp][d] = prev + __popc(peers);
                rk = prev + __popc(peers & lt);
            }
            __syncwarp();
            rank[r] = (uint16_t)rk;
        }
    }
    __syncthreads();
    // the scan ran over both arrays back to back, so the second array's offsets carry the first array's n keys
    if (STAGE) {
        const int d = threadIdx.x;   // RS_THREADS >= 256 bins: one digit per thread
        uint32_t c[RS_WARPS], tot = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) { c[w] = d < nbins ? wcount[w][d] : 0u; tot += c[w]; }
        uint32_t tile_total;
        const uint32_t lstart = block_excl_scan<RS_THREADS>(tot, sh_scan, tile_total);   // where the digit starts inside the tile
        if (d < nbins) {
            uint32_t run = lstart;
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) { wcount[w][d] = run; run += c[w]; }
            gofs[d] = hist_scanned[((size_t)blockIdx.y * nbins + d) * gridDim.x + blockIdx.x] - blockIdx.y * n - lstart;
        }
        __syncthreads();   // also: nobody touches wmatch any more, raw[] becomes the staging tile
#pragma unroll
        for (int r = 0; r < RS_ITEMS; r++) {
            const uint32_t idx = wbase + r * 32 + lane;
            if (idx < n) stage[wcount[warp][(uint32_t)(key[r] >> shift) & mask] + rank[r]] = key[r];
        }
        __syncthreads();
        const uint32_t ntile = min((uint32_t)RS_TILE, n - blockIdx.x * (uint32_t)RS_TILE);
#pragma unroll 4
        for (uint32_t i = threadIdx.x; i < ntile; i += RS_THREADS) {
            const K k = stage[i];
            keys_out[gofs[(uint32_t)(k >> shift) & mask] + i] = k;
        }
    } else {
        // per digit: exclusive prefix over warps on top of the block's global base: wcount becomes the output base of (warp, digit)
        for (int d = threadIdx.x; d < nbins; d += RS_THREADS) {
            uint32_t run = hist_scanned[((size_t)blockIdx.y * nbins + d) * gridDim.x + blockIdx.x] - blockIdx.y * n;
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) { uint32_t c = wcount[w][d]; wcount[w][d] = run; run += c; }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RS_ITEMS; r++) {
            const uint32_t idx = wbase + r * 32 + lane;
            if (idx < n) {
                const uint32_t d = (uint32_t)(key[r] >> shift) & mask;
                const uint32_t pos = wcount[warp][d] + rank[r];
                keys_out[pos] = key[r];
                if (HAS_V) vals_out[pos] = vals_in[idx];
            }
        }
    }
}

// Sort keys (and payload) on bits [begin_bit, end_bit). Ping-pongs between (k0,v0) and (k1,v1);
// returns 0 if the result is in (k0,v0), 1 if in (k1,v1). With a second key array (j0, j1; keys only, same n) both
// arrays are sorted by the same launches and end up on the same side.
// radix_sort_scratch_words(n, narr) words at `scratch` (optional) replace the per-call allocations of the histogram and scan sums.
inline size_t radix_sort_scratch_words(size_t n, unsigned narr) {
    const size_t hist = (size_t)256 * cdiv(n, 4096) * narr;
    return hist + cdiv(hist, SCAN_TILE) + 64;
}
template <typename K, typename V>
inline int radix_sort_bits(K* k0, K* k1, V* v0, V* v1, size_t n, int begin_bit, int end_bit, K* j0 = nullptr, K* j1 = nullptr,
                           uint32_t* scratch = nullptr) {
    constexpr bool HAS_V = !std::is_same<V, NoVal>::value;
    if (n == 0 || end_bit <= begin_bit) return 0;
    MB2_REQUIRE(n < 0xffffffffull, -3, "radix sort: n must be < 2^32");
    const unsigned narr = j0 ? 2 : 1;
    MB2_REQUIRE(narr == 1 || (!HAS_V && j1 && 2 * n < 0xffffffffull), -3, "radix sort: a second array is keys-only and needs 2n < 2^32");
    const int total_bits = end_bit - begin_bit;
    const int passes = (total_bits + 7) / 8;
    const unsigned nb = cdiv(n, RS_TILE);
    static_assert(RS_TILE == 4096, "radix_sort_scratch_words assumes 4096 keys per CTA");
    DevBuf<uint32_t> hist_own;
    if (!scratch) hist_own.alloc((size_t)256 * nb * narr);
    uint32_t* const hist = scratch ? scratch : hist_own.get();
    uint32_t* const sums = scratch ? scratch + (size_t)256 * nb * narr : nullptr;
    int cur = 0;
    int bit = begin_bit;
    for (int p = 0; p < passes; p++) {
        const int bits = (total_bits - (bit - begin_bit) + (passes - p) - 1) / (passes - p);   // spread evenly
        K* kin = cur ? k1 : k0; K* kout = cur ? k0 : k1;
        K* jin = cur ? j1 : j0; K* jout = cur ? j0 : j1;
        V* vin = cur ? v1 : v0; V* vout = cur ? v0 : v1;
        launch(radix_hist_kernel<K>, dim3(nb, narr), RS_THREADS, 0, kin, jin, (uint32_t)n, bit, bits, hist);
        exclusive_scan_u32(hist, hist, (size_t)(1 << bits) * nb * narr, nullptr, sums);
        launch(radix_scatter_kernel<K, V, HAS_V>, dim3(nb, narr), RS_THREADS, 0, kin, jin, kout, jout, vin, vout, (uint32_t)n, bit,
               bits, hist);
        cur ^= 1;
        bit += bits;
    }
    return cur;
}

}  // namespace mb2
