// hits.cu -- the hit table of the alignment stage kept in HBM, and kernel family (d) part 1 on it: the identity / length
// filter, compaction and sort that follow every LASTZ call in the reference's script (wrappers.py:1044-1056: awk
// '0+$5 >= minLen', awk '0+$13 >= minIdt', column projection, sort -k 1,1 -k 3n,4n), plus the projection of the surviving
// rows onto the coverage stage (wrappers.py:1120-1128) without a round trip through the host.
#include <algorithm>

#include "primitives.cuh"
#include "seq.cuh"
#include "internal.cuh"

namespace mb2 {

// strand-local alignments of one query chunk -> LASTZ's output columns (origin-one closed coordinates, query coordinates
// on the + strand). nq = scaffolds of the ORIGINAL query genome; nq2 = scaffolds of the genome that was aligned (2 nq when
// both strands went through one pass); strands as in mb2_align.
__global__ void __launch_bounds__(256)
rows_from_alns_kernel(const uint32_t* __restrict__ tile, const int32_t* __restrict__ s1, const int32_t* __restrict__ e1,
                      const int32_t* __restrict__ s2, const int32_t* __restrict__ e2, const int32_t* __restrict__ score,
                      const int32_t* __restrict__ nm, const int32_t* __restrict__ nc, uint32_t n, uint32_t nq2, int nq, int strands,
                      const uint32_t* __restrict__ qlen, HitCols out, uint32_t base) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int32_t q2 = (int32_t)(tile[k] % nq2), ti = (int32_t)(tile[k] / nq2);
    const int st = strands == 3 ? (q2 >= nq ? 1 : 0) : (strands == 2 ? 1 : 0);
    const int32_t qi = q2 >= nq ? q2 - nq : q2;
    const int32_t m = (int32_t)qlen[q2];
    const uint32_t o = base + k;
    out.c[0][o] = ti; out.c[1][o] = qi; out.c[2][o] = st;
    out.c[3][o] = s1[k] + 1; out.c[4][o] = e1[k];
    out.c[5][o] = st == 0 ? s2[k] + 1 : m - e2[k] + 1;
    out.c[6][o] = st == 0 ? e2[k] : m - s2[k];
    out.c[7][o] = score[k]; out.c[8][o] = nm[k]; out.c[9][o] = nc[k];
}

void DevHits::reserve(size_t want) {
    if (want <= cap) return;
    const size_t ncap = std::max<size_t>(want, std::max<size_t>(1024, cap * 2));
    for (int c = 0; c < 10; c++) {
        DevBuf<int32_t> nb(ncap);
        if (n) MB2_CUDA(cudaMemcpyAsync(nb.get(), col[c].get(), n * sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx().stream));
        col[c] = std::move(nb);
    }
    cap = ncap;
}
HitCols DevHits::view() const {
    HitCols v;
    for (int c = 0; c < 10; c++) v.c[c] = col[c].get();
    return v;
}
void DevHits::append(const AlnSet& a, const Genome& Q, int nq, int strands) {
    if (a.n == 0) return;
    reserve(n + a.n);
    launch(rows_from_alns_kernel, cdiv(a.n, 256), 256, 0, a.tile.get(), a.s1.get(), a.e1.get(), a.s2.get(), a.e2.get(), a.score.get(),
           a.nmatch.get(), a.ncols.get(), a.n, (uint32_t)Q.nscaf, nq, strands, Q.d_len.get(), view(), (uint32_t)n);
    n += a.n;
}

// The printed identity: LASTZ writes '%.1f' of the double 100.0 * nmatch / ncols and awk compares that text as a number.
// tenths(nm, nc) = the printed value times ten, exactly as printf rounds it: the quotient of integers is at least 2^-32 away
// from a rounding boundary unless it sits exactly on one; exact ties are decided by the double that was actually formed
// (representable: round half to even; otherwise by the side the correctly rounded double fell on).
__device__ __forceinline__ int printed_tenths(int nm, int nc) {
    if (nc <= 0) return 0;
    const long long num = 1000ll * nm;
    const long long k = num / nc, rem = num - k * nc;           // value = k + rem / nc tenths
    if (2 * rem < nc) return (int)k;
    if (2 * rem > nc) return (int)k + 1;
    const double q = 100.0 * (double)nm / (double)nc;            // the double LASTZ formats; the tie point is (2k+1)/20
    const double err = fma(q, 20.0, -(double)(2 * k + 1));       // exact sign of q - (2k+1)/20
    if (err > 0.0) return (int)k + 1;
    if (err < 0.0) return (int)k;
    return (int)(k + (k & 1));                                   // exactly representable tie: round half to even
}

// keep[k] = 1 iff the awk filters keep the row: length1 >= min_len and the printed identity >= min_idt; with map_rule also
// import_Align's test int(end1) - int(start1) >= min_len (wrappers.py:76)
__global__ void __launch_bounds__(256)
hit_filter_kernel(HitCols h, uint32_t n, double min_len, double min_idt, int map_rule, uint32_t* __restrict__ keep) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int s1 = h.c[3][k], e1 = h.c[4][k];
    bool ok = (double)(e1 - s1 + 1) >= min_len;
    ok = ok && ((double)printed_tenths(h.c[8][k], h.c[9][k]) / 10.0 >= min_idt);
    if (map_rule) ok = ok && ((double)(e1 - s1) >= min_len);
    keep[k] = ok ? 1u : 0u;
}
// compaction: source rows that passed the filter, in input order, with their (start1, end1) key
__global__ void __launch_bounds__(256)
hit_pos_keys_kernel(HitCols h, uint32_t n, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ keep_off,
                    uint64_t* __restrict__ key_pos, uint32_t* __restrict__ idx) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || !keep[k]) return;
    const uint32_t o = keep_off[k];
    key_pos[o] = ((uint64_t)(uint32_t)h.c[3][k] << 32) | (uint64_t)(uint32_t)h.c[4][k];
    idx[o] = k;
}
// (t_id, q_id) key of the rows in their current order
__global__ void __launch_bounds__(256)
hit_pair_keys_kernel(HitCols h, const uint32_t* __restrict__ idx, uint32_t n, int qbits, uint64_t* __restrict__ key_pair) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t g = idx[k];
    key_pair[k] = ((uint64_t)(uint32_t)h.c[0][g] << qbits) | (uint64_t)(uint32_t)h.c[1][g];
}
__global__ void __launch_bounds__(256)
hit_gather_kernel(HitCols in, HitCols out, const uint32_t* __restrict__ idx, uint32_t n) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const uint32_t g = idx[k];
#pragma unroll
    for (int c = 0; c < 10; c++) out.c[c][k] = in.c[c][g];
}

// Filter + compaction + sort by (t_id, q_id, start1, end1), in place. Two stable LSD radix sorts over the source-row index:
// by (start1, end1) first, then by the pair; rows equal in all four keys keep their input order (the text formatter breaks
// such ties by the whole line, as `sort` does).
void hits_filter_sort(DevHits& h, double min_len, double min_idt, bool map_rule, int nt, int nq) {
    ProfScope ps("hit_filter_sort");
    const uint32_t n = (uint32_t)h.n;
    if (n == 0) return;
    Ctx& cx = ctx();
    DevBuf<uint32_t> keep(n), keep_off(n), d_nk(1);
    launch(hit_filter_kernel, cdiv(n, 256), 256, 0, h.view(), n, min_len, min_idt, map_rule ? 1 : 0, keep.get());
    exclusive_scan_u32(keep.get(), keep_off.get(), n, d_nk.get());
    uint32_t nk = 0;
    MB2_CUDA(cudaMemcpyAsync(&nk, d_nk.get(), sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
    MB2_CUDA(cudaStreamSynchronize(cx.stream));
    if (nk == 0) { h.n = 0; return; }
    int qbits = 1; while (qbits < 32 && (nq >> qbits)) qbits++;
    int tbits = 1; while (tbits < 32 && (nt >> tbits)) tbits++;
    DevBuf<uint64_t> k0(nk), k1(nk);
    DevBuf<uint32_t> i0(nk), i1(nk);
    launch(hit_pos_keys_kernel, cdiv(n, 256), 256, 0, h.view(), n, keep.get(), keep_off.get(), k0.get(), i0.get());
    const int w1 = radix_sort_bits<uint64_t, uint32_t>(k0.get(), k1.get(), i0.get(), i1.get(), nk, 0, 64);
    uint32_t* cur = w1 ? i1.get() : i0.get();
    uint32_t* oth = w1 ? i0.get() : i1.get();
    launch(hit_pair_keys_kernel, cdiv(nk, 256), 256, 0, h.view(), (const uint32_t*)cur, nk, qbits, k0.get());
    const int w2 = radix_sort_bits<uint64_t, uint32_t>(k0.get(), k1.get(), cur, oth, nk, 0, std::min(64, tbits + qbits));
    const uint32_t* order = w2 ? oth : cur;
    DevHits out;
    out.reserve(nk);
    launch(hit_gather_kernel, cdiv(nk, 256), 256, 0, h.view(), out.view(), order, nk);
    out.n = nk;
    for (int c = 0; c < 10; c++) h.col[c] = std::move(out.col[c]);
    h.n = nk; h.cap = out.cap;
}

// rows of a device hit table selected by `which` (0 all, 1 t_id != q_id, 2 t_id == q_id) as (chrom = t_id, start1, end1)
__global__ void __launch_bounds__(256)
hit_select_kernel(HitCols h, uint32_t n, int which, uint32_t* __restrict__ flag) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const bool same = h.c[0][k] == h.c[1][k];
    flag[k] = (which == 0 || (which == 1 && !same) || (which == 2 && same)) ? 1u : 0u;
}
__global__ void __launch_bounds__(256)
hit_project_kernel(HitCols h, uint32_t n, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ off,
                   int32_t* __restrict__ chrom, int32_t* __restrict__ start, int32_t* __restrict__ end) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || !flag[k]) return;
    const uint32_t o = off[k];
    chrom[o] = h.c[0][k]; start[o] = h.c[3][k]; end[o] = h.c[4][k];
}

// coverage stage straight from the device table (the BED projection of wrappers.py:1120-1128 never leaves HBM)
void hits_coverage(const DevHits& h, int which, const int64_t* h_sizes, int nchrom, int min_cov, int min_len, CoverageResult& res) {
    const uint32_t n = (uint32_t)h.n;
    Ctx& cx = ctx();
    DevBuf<int32_t> chrom(n ? n : 1), start(n ? n : 1), end(n ? n : 1);
    uint32_t m = 0;
    if (n) {
        DevBuf<uint32_t> flag(n), off(n), d_m(1);
        launch(hit_select_kernel, cdiv(n, 256), 256, 0, h.view(), n, which, flag.get());
        exclusive_scan_u32(flag.get(), off.get(), n, d_m.get());
        launch(hit_project_kernel, cdiv(n, 256), 256, 0, h.view(), n, (const uint32_t*)flag.get(), (const uint32_t*)off.get(), chrom.get(), start.get(), end.get());
        MB2_CUDA(cudaMemcpyAsync(&m, d_m.get(), sizeof(uint32_t), cudaMemcpyDeviceToHost, cx.stream));
        MB2_CUDA(cudaStreamSynchronize(cx.stream));
    }
    coverage_segments_device(chrom.get(), start.get(), end.get(), m, h_sizes, nchrom, min_cov, min_len, res);
}

}  // namespace mb2
