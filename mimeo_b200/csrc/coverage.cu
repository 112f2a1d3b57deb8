// coverage.cu -- kernel family (d) of the hot path: per-base coverage -> threshold -> merged runs.
//
// Replaces, bit-exactly, the reference's text pipeline
//   awk BED projection | sort | bedtools genomecov -bg | awk '0+$4 >= cov' | sort | bedtools merge | awk minLen
// (wrappers.py:1120-1167, 827-885, 1201-1258; semantics restated in SURVEY.md 9.2 and oracle/).
//
// B200 design (HBM-bound, integer only, no global atomics):
//   1. cov_events_kernel   : every hit (chrom,start,end) becomes a +1 event at its start and a -1 event
//                            at its clipped stop, both in one concatenated coordinate space (one pad
//                            base between scaffolds so runs can never join across scaffolds).
//   2. radix_sort_bits     : the two event arrays are binned by genome tile (LSD radix on the tile-id
//                            bits only; order inside a tile is irrelevant).
//   3. cov_tile_kernel     : each CTA owns a contiguous range of 8192-base tiles. The per-scaffold
//                            difference array lives ONLY in shared memory: zero, apply the tile's events,
//                            block prefix-sum, compare with cov, emit the positions where the
//                            (depth >= cov) flag flips. The depth entering a tile is simply
//                            (#starts before it) - (#stops before it), i.e. two array indices, so tiles
//                            are independent and the dense array never touches HBM.
//   4. cov_gather_kernel   : per-CTA flip lists -> one dense, globally ordered flip list.
//   5. cov_runs_kernel     : consecutive (rise, fall) flips are the merged runs; map back to
//                            (scaffold, local start, local end), apply minLen, ordered compaction.
#include "primitives.cuh"
#include "internal.cuh"
#include "mimeo_b200.h"

namespace mb2 {

constexpr int COV_TILE_BITS = 13;
constexpr int COV_TILE = 1 << COV_TILE_BITS;          // positions per tile
constexpr int COV_THREADS = 512;
constexpr int COV_PER_THREAD = COV_TILE / COV_THREADS; // 16
constexpr int COV_PAD_TILE = COV_TILE + COV_TILE / COV_PER_THREAD;   // +1 word per 16: conflict-free strided reads
constexpr uint32_t COV_SENTINEL = 0xffffffffu;

// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cov_events_kernel(const int32_t* __restrict__ chrom, const int32_t* __restrict__ start, const int32_t* __restrict__ end,
                  uint32_t nhits, const uint32_t* __restrict__ chrom_off, const int32_t* __restrict__ chrom_size,
                  int nchrom, uint32_t* __restrict__ ev_start, uint32_t* __restrict__ ev_stop, int* __restrict__ err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nhits) return;
    const int c = chrom[i];
    const int s = start[i], e = end[i];
    uint32_t ks = COV_SENTINEL, ke = COV_SENTINEL;
    if (c < 0 || c >= nchrom || s < 0 || s > e) {
        atomicOr(err, 1);   // error path only; never taken on valid input
    } else {
        const int N = chrom_size[c];
        if (s < N) {                                   // bedtools: a start beyond the scaffold is never counted
            const int stop = (e >= 1 && e <= N) ? e : N;   // end-1 outside [0,N) is clipped to the last base
            ks = chrom_off[c] + (uint32_t)s;
            ke = chrom_off[c] + (uint32_t)stop;
        }
    }
    ev_start[i] = ks;
    ev_stop[i] = ke;
}

// lower_bound on a sorted-by-tile array: first index whose key >= value (keys compared on full value
// is fine because every key of an earlier tile is smaller than `value` = tile start).
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* __restrict__ a, uint32_t n, uint32_t value) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (a[mid] < value) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ int pad_idx(int i) { return i + (i >> 4); }

constexpr int COV_MAX_TILES_PER_CTA = 32;
constexpr int COV_PF = 2;   // events per thread and array fetched one tile ahead (covers tiles of up to 1024 + 1024 events)

// Events [r0, r1) of one array for one tile: the first COV_PF * COV_THREADS of them into registers.
__device__ __forceinline__ void cov_fetch(const uint32_t* __restrict__ ev, uint32_t r0, uint32_t r1, int tid, uint32_t (&q)[COV_PF]) {
#pragma unroll
    for (int k = 0; k < COV_PF; k++) {
        const uint32_t i = r0 + tid + k * COV_THREADS;
        q[k] = i < r1 ? ev[i] : COV_SENTINEL;
    }
}

__global__ void __launch_bounds__(COV_THREADS, 3)
cov_tile_kernel(const uint32_t* __restrict__ ev_start, const uint32_t* __restrict__ ev_stop, uint32_t nev,
                uint32_t num_tiles, uint32_t tiles_per_cta, int cov,
                uint32_t* __restrict__ flips_staging, uint32_t* __restrict__ cta_count, uint32_t* __restrict__ cta_base,
                int* __restrict__ err) {
    __shared__ int diff[COV_PAD_TILE];
    __shared__ uint32_t sh_scan[COV_THREADS / 32 + 1];
    __shared__ uint32_t s_off[COV_MAX_TILES_PER_CTA + 2], e_off[COV_MAX_TILES_PER_CTA + 2];
    const int tid = threadIdx.x;
    const uint32_t t0 = blockIdx.x * tiles_per_cta;
    const uint32_t t1 = min(t0 + tiles_per_cta, num_tiles);
    const uint32_t nt = t1 - t0;
    // event ranges of all the CTA's tiles at once: 2 (nt + 1) independent binary searches, one latency chain in total
    if (tid <= (int)nt) s_off[tid] = lower_bound_u32(ev_start, nev, (t0 + tid) << COV_TILE_BITS);
    if (tid >= 64 && tid - 64 <= (int)nt) e_off[tid - 64] = lower_bound_u32(ev_stop, nev, (t0 + tid - 64) << COV_TILE_BITS);
    for (int i = tid; i < COV_PAD_TILE; i += COV_THREADS) diff[i] = 0;   // once: every tile re-zeroes what it read
    __syncthreads();
    if (tid == 0) { s_off[nt + 1] = s_off[nt]; e_off[nt + 1] = e_off[nt]; }   // "tile nt" is empty: the look-ahead of the last tile fetches nothing
    __syncthreads();
    const uint32_t out_base = s_off[0] + e_off[0];   // #flips in this CTA's range <= #events in it: disjoint staging regions
    uint32_t out_n = 0;

    uint32_t ps[COV_PF], pe[COV_PF];   // first events of the tile about to be processed
    cov_fetch(ev_start, s_off[0], s_off[1], tid, ps);
    cov_fetch(ev_stop, e_off[0], e_off[1], tid, pe);
    for (uint32_t j = 0; j < nt; j++) {
        const uint32_t s0 = s_off[j], s1 = s_off[j + 1], e0 = e_off[j], e1 = e_off[j + 1];
        uint32_t ns[COV_PF], ne[COV_PF];   // the next tile's events stay in flight while this tile is scanned
        cov_fetch(ev_start, s1, s_off[j + 2], tid, ns);
        cov_fetch(ev_stop, e1, e_off[j + 2], tid, ne);
        if (s0 != s1 || e0 != e1) {         // no event: depth is constant across the tile, so no flip can occur
            const uint32_t lo = (t0 + j) << COV_TILE_BITS;
#pragma unroll
            for (int k = 0; k < COV_PF; k++) {
                if (ps[k] != COV_SENTINEL) {
                    const uint32_t r = ps[k] - lo;
                    if (r < (uint32_t)COV_TILE) atomicAdd(&diff[pad_idx((int)r)], 1); else atomicOr(err, 2);   // binning invariant
                }
                if (pe[k] != COV_SENTINEL) {
                    const uint32_t r = pe[k] - lo;
                    if (r < (uint32_t)COV_TILE) atomicAdd(&diff[pad_idx((int)r)], -1); else atomicOr(err, 2);
                }
            }
            for (uint32_t i = s0 + tid + COV_PF * COV_THREADS; i < s1; i += COV_THREADS) {   // crowded tiles: the rest
                const uint32_t r = ev_start[i] - lo;
                if (r < (uint32_t)COV_TILE) atomicAdd(&diff[pad_idx((int)r)], 1); else atomicOr(err, 2);
            }
            for (uint32_t i = e0 + tid + COV_PF * COV_THREADS; i < e1; i += COV_THREADS) {
                const uint32_t r = ev_stop[i] - lo;
                if (r < (uint32_t)COV_TILE) atomicAdd(&diff[pad_idx((int)r)], -1); else atomicOr(err, 2);
            }
            __syncthreads();
            // each thread owns 16 consecutive positions
            int* const mine = diff + tid * (COV_PER_THREAD + 1);
            int tsum = 0;
#pragma unroll
            for (int k = 0; k < COV_PER_THREAD; k++) tsum += mine[k];
            uint32_t total;
            const uint32_t excl = block_excl_scan_open<COV_THREADS>((uint32_t)tsum, sh_scan, total);
            int depth = (int)s0 - (int)e0 + (int)excl;   // depth at lo-1 = (#starts before) - (#stops before), plus the threads before
            bool prev = depth >= cov;
            uint32_t mask = 0;
#pragma unroll
            for (int k = 0; k < COV_PER_THREAD; k++) {
                depth += mine[k];
                mine[k] = 0;                   // leave the difference array clean for the next tile
                const bool f = depth >= cov;
                if (f != prev) mask |= 1u << k;
                prev = f;
            }
            if (__syncthreads_or(mask != 0)) {   // most tiles of a deep pile-up hold no flip at all
                uint32_t ftotal;
                uint32_t off = block_excl_scan<COV_THREADS>((uint32_t)__popc(mask), sh_scan, ftotal);
                uint32_t widx = out_base + out_n + off;
                while (mask) {
                    const int k = __ffs(mask) - 1;
                    mask &= mask - 1;
                    if (widx < 2u * nev) flips_staging[widx] = lo + tid * COV_PER_THREAD + k; else atomicOr(err, 4);   // #flips <= #events
                    widx++;
                }
                out_n += ftotal;
            }
            // both paths end with a block barrier after the last access to diff[] and sh_scan[]
        }
#pragma unroll
        for (int k = 0; k < COV_PF; k++) { ps[k] = ns[k]; pe[k] = ne[k]; }
    }
    if (tid == 0) { cta_count[blockIdx.x] = out_n; cta_base[blockIdx.x] = out_base; }
}

// Per-CTA flip lists -> one dense, globally ordered flip list. Every CTA sums the counts of the CTAs before it (a few
// thousand words at most), so no separate scan launch; the last CTA publishes the total.
__global__ void __launch_bounds__(256)
cov_gather_kernel(const uint32_t* __restrict__ flips_staging, const uint32_t* __restrict__ cta_count,
                  const uint32_t* __restrict__ cta_base, uint32_t* __restrict__ flips, uint32_t cap,
                  uint32_t* __restrict__ nflips_out, int* __restrict__ err) {
    __shared__ uint32_t sh[256 / 32 + 1];
    uint32_t part = 0;
    for (uint32_t i = threadIdx.x; i < blockIdx.x; i += 256) part += cta_count[i];
    uint32_t o0;
    block_excl_scan<256>(part, sh, o0);
    const uint32_t n = cta_count[blockIdx.x];
    const uint32_t b0 = cta_base[blockIdx.x];
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *nflips_out = o0 + n;
    if ((uint64_t)b0 + n > cap || (uint64_t)o0 + n > cap) { if (threadIdx.x == 0) atomicOr(err, 8); return; }
    for (uint32_t i = threadIdx.x; i < n; i += 256) flips[o0 + i] = flips_staging[b0 + i];
}

// Runs -> segments in ONE launch when there are at most cap_seg runs (one CTA walks them with a running output offset):
// minLen filter, ordered compaction, scaffold lookup. More runs than that: *nseg_out = 0xffffffff and the host takes the
// flag / scan / write path below.
constexpr uint32_t COV_RUNS_FUSED_MAX = 1u << 18;
__global__ void __launch_bounds__(1024)
cov_runs_fused_kernel(const uint32_t* __restrict__ flips, const uint32_t* __restrict__ nflips_p, int min_len,
                      const uint32_t* __restrict__ chrom_off, int nchrom, int32_t* __restrict__ seg_chrom,
                      int32_t* __restrict__ seg_start, int32_t* __restrict__ seg_end, uint32_t cap_seg, uint32_t* __restrict__ nseg_out) {
    __shared__ uint32_t sh[1024 / 32 + 1];
    const uint32_t nruns = *nflips_p >> 1;
    if (nruns > cap_seg) { if (threadIdx.x == 0) *nseg_out = 0xffffffffu; return; }
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nruns; base += 1024) {
        const uint32_t k = base + threadIdx.x;
        uint32_t s = 0, e = 0;
        bool keep = false;
        if (k < nruns) {
            const uint2 se = reinterpret_cast<const uint2*>(flips)[k];
            s = se.x; e = se.y;
            keep = (int64_t)e - (int64_t)s >= (int64_t)min_len;
        }
        uint32_t total;
        const uint32_t off = block_excl_scan<1024>(keep ? 1u : 0u, sh, total);
        if (keep) {
            int lo = 0, hi = nchrom;            // last scaffold whose offset <= s
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (chrom_off[mid] <= s) lo = mid; else hi = mid; }
            const uint32_t o = carry + off;
            seg_chrom[o] = lo;
            seg_start[o] = (int32_t)(s - chrom_off[lo]);
            seg_end[o] = (int32_t)(e - chrom_off[lo]);
        }
        carry += total;
    }
    if (threadIdx.x == 0) *nseg_out = carry;
}

// flips[2k], flips[2k+1] = rise/fall of run k in concatenated coordinates
__global__ void __launch_bounds__(256)
cov_runs_flag_kernel(const uint32_t* __restrict__ flips, const uint32_t* __restrict__ nflips_p, int min_len, uint32_t* __restrict__ keep,
                     uint32_t cap_runs) {
    const uint32_t nruns = min(*nflips_p >> 1, cap_runs);
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nruns) return;
    keep[k] = ((int64_t)flips[2 * k + 1] - (int64_t)flips[2 * k] >= (int64_t)min_len) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
cov_runs_write_kernel(const uint32_t* __restrict__ flips, const uint32_t* __restrict__ nflips_p, int min_len,
                      const uint32_t* __restrict__ keep_off, const uint32_t* __restrict__ chrom_off, int nchrom,
                      int32_t* __restrict__ seg_chrom, int32_t* __restrict__ seg_start, int32_t* __restrict__ seg_end,
                      uint32_t cap_runs, uint32_t cap_seg) {
    const uint32_t nruns = min(*nflips_p >> 1, cap_runs);
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nruns) return;
    const uint32_t s = flips[2 * k], e = flips[2 * k + 1];
    if ((int64_t)e - (int64_t)s < (int64_t)min_len) return;
    int lo = 0, hi = nchrom;            // last scaffold whose offset <= s
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (chrom_off[mid] <= s) lo = mid; else hi = mid; }
    const uint32_t o = keep_off[k];
    if (o >= cap_seg) return;
    seg_chrom[o] = lo;
    seg_start[o] = (int32_t)(s - chrom_off[lo]);
    seg_end[o] = (int32_t)(e - chrom_off[lo]);
}

// pinned landing / staging zone: the few counters the host reads back (pageable targets would stage every copy) and the
// per-scaffold offsets and sizes it sends (no wait for a pageable upload). Reused by every call: each call ends with a
// stream synchronisation, so the previous call's transfers are complete.
struct HostInfo { int err; uint32_t nflips, nseg; };   // mirrors d_info[0..2]
struct HostZone {
    HostInfo* info = nullptr;
    uint32_t* meta = nullptr;   // [nchrom] offsets, then [nchrom] sizes
    size_t meta_cap = 0;
};
static HostZone& host_zone(size_t nchrom) {
    static HostZone z;
    if (!z.info) MB2_CUDA(cudaMallocHost((void**)&z.info, sizeof(HostInfo)));
    if (2 * nchrom > z.meta_cap) {
        if (z.meta) cudaFreeHost(z.meta);
        z.meta = nullptr; z.meta_cap = 0;
        const size_t cap = std::max<size_t>(2 * nchrom, 4096);
        MB2_CUDA(cudaMallocHost((void**)&z.meta, cap * sizeof(uint32_t)));
        z.meta_cap = cap;
    }
    return z;
}

// Grow-only device scratch of this stage, carved per call: a call makes no allocator traffic once the arena is large enough
// (a dozen stream-ordered allocations and frees per call were a visible share of a 0.6 ms step). Safe to reuse: everything
// runs on the one library stream and every call ends with a stream synchronisation.
struct Arena {
    uint8_t* p = nullptr;
    size_t cap = 0, used = 0;
    void reserve(size_t bytes) {
        used = 0;
        if (bytes <= cap) return;
        if (p) { scratch_free(p); p = nullptr; cap = 0; }
        const size_t want = bytes + bytes / 8;
        p = static_cast<uint8_t*>(scratch_alloc(want));
        cap = want;
    }
    template <typename T> T* take(size_t count) {
        T* r = (T*)(p + used);
        used += (count * sizeof(T) + 255) & ~(size_t)255;
        MB2_REQUIRE(used <= cap, -5, "coverage: scratch arena overrun");
        return r;
    }
};
static Arena g_arena;
void coverage_release_scratch() {   // mb2_shutdown
    if (g_arena.p) scratch_free(g_arena.p);
    g_arena.p = nullptr; g_arena.cap = g_arena.used = 0;
}
static size_t arena_slot(size_t count, size_t elem) { return (count * elem + 255) & ~(size_t)255; }

// -------------------------------------------------------------------------------------------------
// Host driver. All pointers are device pointers; everything is enqueued on the library stream.
// One host round trip (error bits, flip count, segment count) unless there are more than COV_RUNS_FUSED_MAX runs.
// -------------------------------------------------------------------------------------------------
void coverage_segments_device(const int32_t* d_chrom, const int32_t* d_start, const int32_t* d_end, uint64_t nhits,
                              const int64_t* h_sizes, int nchrom, int min_cov, int min_len, CoverageResult& res) {
    MB2_REQUIRE(nchrom > 0, -2, "coverage: need at least one scaffold");
    MB2_REQUIRE(nhits < 0x7fffffffull, -2, "coverage: more than 2^31-1 hits in one call");
    HostZone& hz = host_zone((size_t)nchrom);
    uint64_t G = 0;
    for (int c = 0; c < nchrom; c++) {
        MB2_REQUIRE(h_sizes[c] > 0 && h_sizes[c] < 0x7fffffffll, -2, "coverage: scaffold size out of range");
        hz.meta[c] = (uint32_t)G;
        hz.meta[nchrom + c] = (uint32_t)h_sizes[c];
        G += (uint64_t)h_sizes[c] + 1;   // one pad base after every scaffold
        MB2_REQUIRE(G < 0xfff00000ull, -2, "coverage: concatenated genome exceeds 2^32 positions; split the call by scaffold groups");
    }
    res.n = 0;
    if (nhits == 0) return;
    Ctx& cx = ctx();
    struct DebugScope {   // MB2_DEBUG_COV=1: synchronise after every launch of this stage only
        bool prev; DebugScope() : prev(ctx().debug_sync) { if (getenv("MB2_DEBUG_COV")) { cudaError_t e = cudaStreamSynchronize(ctx().stream); if (e != cudaSuccess) throw Error(-100, std::string("fault BEFORE the coverage stage: ") + cudaGetErrorString(e)); ctx().debug_sync = true; } }
        ~DebugScope() { ctx().debug_sync = prev; }
    } debug_scope;
    const uint32_t H = (uint32_t)nhits;
    static int occ = 0;
    if (occ == 0) {
        MB2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, cov_tile_kernel, COV_THREADS, 0));
        if (occ < 1) occ = 1;
    }
    // one wave of CTAs (the kernel is a chain of short barrier-separated phases: a second, partial wave would idle most SMs),
    // more only when a CTA would otherwise own more than COV_MAX_TILES_PER_CTA tiles
    static int waves = 0;   // MB2_COV_WAVES: CTAs per resident slot (measurement knob; default 1)
    if (waves == 0) { const char* e = getenv("MB2_COV_WAVES"); waves = e ? std::max(1, atoi(e)) : 1; }
    const uint32_t num_tiles = (uint32_t)((G + COV_TILE - 1) >> COV_TILE_BITS);
    const uint32_t max_ctas = (uint32_t)cx.sm_count * (uint32_t)occ * (uint32_t)waves;
    const uint32_t tiles_per_cta = std::min<uint32_t>((num_tiles + max_ctas - 1) / max_ctas, COV_MAX_TILES_PER_CTA);
    const uint32_t nctas = (num_tiles + tiles_per_cta - 1) / tiles_per_cta;
    const size_t sort_words = radix_sort_scratch_words(H, 2);
    g_arena.reserve(arena_slot((size_t)2 * nchrom, 4) + arena_slot(4, 4) + 4 * arena_slot(H, 4) + arena_slot(sort_words, 4) +
                    2 * arena_slot((size_t)2 * H, 4) + 2 * arena_slot(nctas, 4));
    uint32_t* const d_meta = g_arena.take<uint32_t>((size_t)2 * nchrom);
    uint32_t* const d_info = g_arena.take<uint32_t>(4);   // [0] error bits, [1] flip count, [2] segment count: one read-back
    uint32_t* const evs0 = g_arena.take<uint32_t>(H);
    uint32_t* const evs1 = g_arena.take<uint32_t>(H);
    uint32_t* const eve0 = g_arena.take<uint32_t>(H);
    uint32_t* const eve1 = g_arena.take<uint32_t>(H);
    uint32_t* const sort_scratch = g_arena.take<uint32_t>(sort_words);
    uint32_t* const staging = g_arena.take<uint32_t>((size_t)2 * H);
    uint32_t* const flips = g_arena.take<uint32_t>((size_t)2 * H);
    uint32_t* const cta_count = g_arena.take<uint32_t>(nctas);
    uint32_t* const cta_base = g_arena.take<uint32_t>(nctas);
    MB2_CUDA(cudaMemcpyAsync(d_meta, hz.meta, (size_t)2 * nchrom * sizeof(uint32_t), cudaMemcpyHostToDevice, cx.stream));
    const uint32_t* const d_off = d_meta;
    const int32_t* const d_size = (const int32_t*)(d_meta + nchrom);
    MB2_CUDA(cudaMemsetAsync(d_info, 0, 3 * sizeof(uint32_t), cx.stream));
    int* const d_err_p = (int*)d_info;
    uint32_t* const d_nflips_p = d_info + 1;
    uint32_t* const d_nseg_p = d_info + 2;
    { ProfScope ps("cov_events");
    launch(cov_events_kernel, cdiv(H, 256), 256, 0, d_chrom, d_start, d_end, H, d_off, d_size, nchrom,
           evs0, eve0, d_err_p); }

    int top = COV_TILE_BITS;
    while (top < 32 && ((G + COV_TILE) >> top) != 0) top++;       // sentinel (all ones) must stay the largest tile id
    NoVal* nv = nullptr;
    int w;
    { ProfScope ps("cov_bin_events");   // both event arrays through the same launches
      w = radix_sort_bits<uint32_t, NoVal>(evs0, evs1, nv, nv, H, COV_TILE_BITS, top, eve0, eve1, sort_scratch); }
    const uint32_t* s_sorted = w ? evs1 : evs0;
    const uint32_t* e_sorted = w ? eve1 : eve0;

    { ProfScope ps("cov_tile");
    launch(cov_tile_kernel, nctas, COV_THREADS, 0, s_sorted, e_sorted, H, num_tiles, tiles_per_cta,
           min_cov < 1 ? 1 : min_cov, staging, cta_count, cta_base, d_err_p); }
    launch(cov_gather_kernel, nctas, 256, 0, staging, cta_count, cta_base, flips, 2u * H, d_nflips_p, d_err_p);

    // runs: at most H of them (every run needs at least one start event)
    const bool ext = res.ext_chrom != nullptr;
    const uint32_t cap_fused = (uint32_t)std::min<uint64_t>(std::min<uint32_t>(H, COV_RUNS_FUSED_MAX), ext ? res.ext_cap : ~0ull);
    if (!ext) { res.chrom.alloc(cap_fused); res.start.alloc(cap_fused); res.end.alloc(cap_fused); }
    int32_t* o_chrom = ext ? res.ext_chrom : res.chrom.get();
    int32_t* o_start = ext ? res.ext_start : res.start.get();
    int32_t* o_end = ext ? res.ext_end : res.end.get();
    launch(cov_runs_fused_kernel, 1, 1024, 0, flips, d_nflips_p, min_len, d_off, nchrom, o_chrom, o_start, o_end, cap_fused, d_nseg_p);
    MB2_CUDA(cudaMemcpyAsync(hz.info, d_info, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
    MB2_CUDA(cudaStreamSynchronize(cx.stream));
    const int h_err = hz.info->err;
    const uint32_t h_nflips = hz.info->nflips;
    uint32_t h_nseg = hz.info->nseg;
    MB2_REQUIRE((h_err & 1) == 0, -4, "coverage: invalid hit (scaffold index out of range, negative start, or start > end)");
    MB2_REQUIRE(h_err == 0, -5, std::string("coverage: internal invariant violated, code ") + std::to_string(h_err));
    MB2_REQUIRE((h_nflips & 1u) == 0 && h_nflips <= 2ull * H, -5, std::string("coverage: internal error, inconsistent flip count ") + std::to_string(h_nflips));
    const uint32_t nruns = h_nflips >> 1;   // consecutive (rise, fall) flips
    if (nruns > cap_fused) {                // very many runs: flag, device-wide scan, write at the exact size
        MB2_REQUIRE(h_nseg == 0xffffffffu, -5, "coverage: internal error, fused run stage ignored its capacity");
        DevBuf<uint32_t> keep(nruns), keep_off(nruns);
        launch(cov_runs_flag_kernel, cdiv(nruns, 256), 256, 0, flips, d_nflips_p, min_len, keep.get(), nruns);
        exclusive_scan_u32(keep.get(), keep_off.get(), nruns, d_nseg_p);
        MB2_CUDA(cudaMemcpyAsync(&hz.info->nseg, d_nseg_p, sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
        MB2_CUDA(cudaStreamSynchronize(cx.stream));
        h_nseg = hz.info->nseg;
        MB2_REQUIRE(h_nseg <= nruns, -5, "coverage: internal error, more segments than runs");
        if (ext) {
            if (h_nseg > res.ext_cap) { res.n = h_nseg; throw Error(-6, "coverage: " + std::to_string(h_nseg) + " segments do not fit the caller's arrays of " + std::to_string(res.ext_cap)); }
        } else {
            res.chrom.alloc(h_nseg); res.start.alloc(h_nseg); res.end.alloc(h_nseg);
            o_chrom = res.chrom.get(); o_start = res.start.get(); o_end = res.end.get();
        }
        if (h_nseg)
            launch(cov_runs_write_kernel, cdiv(nruns, 256), 256, 0, flips, d_nflips_p, min_len, keep_off.get(),
                   d_off, nchrom, o_chrom, o_start, o_end, nruns, h_nseg);
    }
    MB2_REQUIRE(h_nseg <= nruns, -5, "coverage: internal error, more segments than runs");
    res.n = h_nseg;
}

}  // namespace mb2
