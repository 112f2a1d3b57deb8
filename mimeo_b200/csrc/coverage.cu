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

__global__ void __launch_bounds__(COV_THREADS, 3)
cov_tile_kernel(const uint32_t* __restrict__ ev_start, const uint32_t* __restrict__ ev_stop, uint32_t nev,
                uint32_t num_tiles, uint32_t tiles_per_cta, int cov,
                uint32_t* __restrict__ flips_staging, uint32_t* __restrict__ cta_count, uint32_t* __restrict__ cta_base,
                int* __restrict__ err) {
    __shared__ int diff[COV_PAD_TILE];
    __shared__ uint32_t sh_scan[COV_THREADS / 32 + 1];
    __shared__ uint32_t s_off[COV_MAX_TILES_PER_CTA + 1], e_off[COV_MAX_TILES_PER_CTA + 1];
    const int tid = threadIdx.x;
    const uint32_t t0 = blockIdx.x * tiles_per_cta;
    const uint32_t t1 = min(t0 + tiles_per_cta, num_tiles);
    const uint32_t nt = t1 - t0;
    // event ranges of all the CTA's tiles at once: 2 (nt + 1) independent binary searches, one latency chain in total
    if (tid <= (int)nt) s_off[tid] = lower_bound_u32(ev_start, nev, (t0 + tid) << COV_TILE_BITS);
    if (tid >= 64 && tid - 64 <= (int)nt) e_off[tid - 64] = lower_bound_u32(ev_stop, nev, (t0 + tid - 64) << COV_TILE_BITS);
    __syncthreads();
    const uint32_t out_base = s_off[0] + e_off[0];   // #flips in this CTA's range <= #events in it: disjoint staging regions
    uint32_t out_n = 0;

    for (uint32_t j = 0; j < nt; j++) {
        const uint32_t s0 = s_off[j], s1 = s_off[j + 1], e0 = e_off[j], e1 = e_off[j + 1];
        if (s0 == s1 && e0 == e1) continue;         // no event: depth is constant across the tile, so no flip can occur
        const uint32_t lo = (t0 + j) << COV_TILE_BITS;
        for (int i = tid; i < COV_PAD_TILE; i += COV_THREADS) diff[i] = 0;
        __syncthreads();
        for (uint32_t i = s0 + tid; i < s1; i += COV_THREADS) {   // +1 events of this tile (independent loads)
            const uint32_t r = ev_start[i] - lo;
            if (r < (uint32_t)COV_TILE) atomicAdd(&diff[pad_idx((int)r)], 1); else atomicOr(err, 2);   // binning invariant
        }
        for (uint32_t i = e0 + tid; i < e1; i += COV_THREADS) {   // -1 events
            const uint32_t r = ev_stop[i] - lo;
            if (r < (uint32_t)COV_TILE) atomicAdd(&diff[pad_idx((int)r)], -1); else atomicOr(err, 2);
        }
        __syncthreads();
        const int base_depth = (int)s0 - (int)e0;   // depth at position lo-1 = (#starts before) - (#stops before)
        // each thread owns 16 consecutive positions
        int d[COV_PER_THREAD];
        int tsum = 0;
#pragma unroll
        for (int k = 0; k < COV_PER_THREAD; k++) { d[k] = diff[tid * (COV_PER_THREAD + 1) + k]; tsum += d[k]; }
        uint32_t total;
        const uint32_t excl = block_excl_scan<COV_THREADS>((uint32_t)tsum, sh_scan, total);
        int depth = base_depth + (int)excl;
        bool prev = depth >= cov;
        uint32_t mask = 0;
#pragma unroll
        for (int k = 0; k < COV_PER_THREAD; k++) {
            depth += d[k];
            const bool f = depth >= cov;
            if (f != prev) mask |= 1u << k;
            prev = f;
        }
        if (!__syncthreads_or(mask != 0)) continue;   // most tiles of a deep pile-up hold no flip at all
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
        // every path to the next tile ends with a block barrier after the last read of diff[], so it can be zeroed again
    }
    if (tid == 0) { cta_count[blockIdx.x] = out_n; cta_base[blockIdx.x] = out_base; }
}

__global__ void __launch_bounds__(256)
cov_gather_kernel(const uint32_t* __restrict__ flips_staging, const uint32_t* __restrict__ cta_count,
                  const uint32_t* __restrict__ cta_base, const uint32_t* __restrict__ cta_out, uint32_t* __restrict__ flips,
                  uint32_t cap, int* __restrict__ err) {
    const uint32_t n = cta_count[blockIdx.x];
    const uint32_t b0 = cta_base[blockIdx.x], o0 = cta_out[blockIdx.x];
    if ((uint64_t)b0 + n > cap || (uint64_t)o0 + n > cap) { if (threadIdx.x == 0) atomicOr(err, 8); return; }
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) flips[o0 + i] = flips_staging[b0 + i];
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

// pinned landing zone for the few counters the host reads back (pageable targets would stage every 4-byte copy)
struct HostInfo { int err; uint32_t nflips, nseg; };   // err and nflips mirror d_info[0..1]
static HostInfo* host_info() {
    static HostInfo* p = nullptr;
    if (!p) MB2_CUDA(cudaMallocHost((void**)&p, sizeof(HostInfo)));
    return p;
}

// -------------------------------------------------------------------------------------------------
// Host driver. All pointers are device pointers; everything is enqueued on the library stream.
// Returns the number of segments (device->host read of two counters is the only sync).
// -------------------------------------------------------------------------------------------------
void coverage_segments_device(const int32_t* d_chrom, const int32_t* d_start, const int32_t* d_end, uint64_t nhits,
                              const int64_t* h_sizes, int nchrom, int min_cov, int min_len, CoverageResult& res) {
    MB2_REQUIRE(nchrom > 0, -2, "coverage: need at least one scaffold");
    MB2_REQUIRE(nhits < 0x7fffffffull, -2, "coverage: more than 2^31-1 hits in one call");
    std::vector<uint32_t> off(nchrom);
    std::vector<int32_t> size32(nchrom);
    uint64_t G = 0;
    for (int c = 0; c < nchrom; c++) {
        MB2_REQUIRE(h_sizes[c] > 0 && h_sizes[c] < 0x7fffffffll, -2, "coverage: scaffold size out of range");
        off[c] = (uint32_t)G;
        size32[c] = (int32_t)h_sizes[c];
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
    DevBuf<uint32_t> d_off(nchrom);
    DevBuf<int32_t> d_size(nchrom);
    MB2_CUDA(cudaMemcpyAsync(d_off.get(), off.data(), nchrom * sizeof(uint32_t), cudaMemcpyHostToDevice, cx.stream));
    MB2_CUDA(cudaMemcpyAsync(d_size.get(), size32.data(), nchrom * sizeof(int32_t), cudaMemcpyHostToDevice, cx.stream));
    MB2_CUDA(cudaStreamSynchronize(cx.stream));   // off/size32 are stack-lifetime host buffers

    const uint32_t H = (uint32_t)nhits;
    DevBuf<uint32_t> evs0(H), evs1(H), eve0(H), eve1(H);
    DevBuf<uint32_t> d_info(2);   // [0] error bits, [1] flip count: one read-back
    MB2_CUDA(cudaMemsetAsync(d_info.get(), 0, 2 * sizeof(uint32_t), cx.stream));
    int* const d_err_p = (int*)d_info.get();
    uint32_t* const d_nflips_p = d_info.get() + 1;
    { ProfScope ps("cov_events");
    launch(cov_events_kernel, cdiv(H, 256), 256, 0, d_chrom, d_start, d_end, H, d_off.get(), d_size.get(), nchrom,
           evs0.get(), eve0.get(), d_err_p); }

    int top = COV_TILE_BITS;
    while (top < 32 && ((G + COV_TILE) >> top) != 0) top++;       // sentinel (all ones) must stay the largest tile id
    NoVal* nv = nullptr;
    int w;
    { ProfScope ps("cov_bin_events");   // both event arrays through the same launches
      w = radix_sort_bits<uint32_t, NoVal>(evs0.get(), evs1.get(), nv, nv, H, COV_TILE_BITS, top, eve0.get(), eve1.get()); }
    const uint32_t* s_sorted = w ? evs1.get() : evs0.get();
    const uint32_t* e_sorted = w ? eve1.get() : eve0.get();

    const uint32_t num_tiles = (uint32_t)((G + COV_TILE - 1) >> COV_TILE_BITS);
    const uint32_t max_ctas = (uint32_t)cx.sm_count * 4;
    const uint32_t tiles_per_cta = std::min<uint32_t>((num_tiles + max_ctas - 1) / max_ctas, COV_MAX_TILES_PER_CTA);
    const uint32_t nctas = (num_tiles + tiles_per_cta - 1) / tiles_per_cta;
    DevBuf<uint32_t> staging((size_t)2 * H), cta_count(nctas), cta_base(nctas), cta_out(nctas);
    { ProfScope ps("cov_tile");
    launch(cov_tile_kernel, nctas, COV_THREADS, 0, s_sorted, e_sorted, H, num_tiles, tiles_per_cta,
           min_cov < 1 ? 1 : min_cov, staging.get(), cta_count.get(), cta_base.get(), d_err_p); }
    exclusive_scan_u32(cta_count.get(), cta_out.get(), nctas, d_nflips_p);
    DevBuf<uint32_t> flips((size_t)2 * H);
    launch(cov_gather_kernel, nctas, 256, 0, staging.get(), cta_count.get(), cta_base.get(), cta_out.get(), flips.get(), 2u * H, d_err_p);

    // first (and usually only large) host round trip: the flip count sizes everything downstream, so the run stage
    // costs O(runs), not O(hits)
    HostInfo* hi = host_info();
    MB2_CUDA(cudaMemcpyAsync(hi, d_info.get(), 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
    MB2_CUDA(cudaStreamSynchronize(cx.stream));
    const int h_err = hi->err;
    const uint32_t h_nflips = hi->nflips;
    MB2_REQUIRE((h_err & 1) == 0, -4, "coverage: invalid hit (scaffold index out of range, negative start, or start > end)");
    MB2_REQUIRE(h_err == 0, -5, std::string("coverage: internal invariant violated, code ") + std::to_string(h_err));
    MB2_REQUIRE((h_nflips & 1u) == 0 && h_nflips <= 2ull * H, -5, std::string("coverage: internal error, inconsistent flip count ") + std::to_string(h_nflips));
    const uint32_t nruns = h_nflips >> 1;   // consecutive (rise, fall) flips
    if (nruns == 0) return;
    DevBuf<uint32_t> keep(nruns), keep_off(nruns), d_nseg(1);
    launch(cov_runs_flag_kernel, cdiv(nruns, 256), 256, 0, flips.get(), d_nflips_p, min_len, keep.get(), nruns);
    exclusive_scan_u32(keep.get(), keep_off.get(), nruns, d_nseg.get());
    MB2_CUDA(cudaMemcpyAsync(&hi->nseg, d_nseg.get(), sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
    MB2_CUDA(cudaStreamSynchronize(cx.stream));
    const uint32_t h_nseg = hi->nseg;
    MB2_REQUIRE(h_nseg <= nruns, -5, "coverage: internal error, more segments than runs");
    res.chrom.alloc(h_nseg); res.start.alloc(h_nseg); res.end.alloc(h_nseg);
    res.n = h_nseg;
    if (h_nseg)
        launch(cov_runs_write_kernel, cdiv(nruns, 256), 256, 0, flips.get(), d_nflips_p, min_len, keep_off.get(),
               d_off.get(), nchrom, res.chrom.get(), res.start.get(), res.end.get(), nruns, h_nseg);
}

}  // namespace mb2
