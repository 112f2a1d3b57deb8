// align.cu -- host driver of the alignment half: target table -> seed scan -> HSPs -> chain -> gapped.
// One call handles one query strand; the C ABI layer runs both strands and converts coordinates.
#include "primitives.cuh"
#include "seq.cuh"
#include "internal.cuh"

namespace mb2 {

// Kept HSPs of every (target scaffold, query scaffold) tile of T x Q (Q already strand-oriented).
void align_hsps(const Genome& T, const Genome& Q, const AlignParams& p, HspSet& hsps, unsigned long long* h_counters) {
    Ctx& cx = ctx();
    MB2_REQUIRE(T.G + Q.G < 0xffffffffull, -3, "align: target + query exceed 2^32 padded positions");
    DevBuf<unsigned long long> counters(CNT_N);
    MB2_CUDA(cudaMemsetAsync(counters.get(), 0, CNT_N * sizeof(unsigned long long), cx.stream));
    SeedTable tab;
    build_seed_table(T, 0, (uint32_t)T.G, tab);
    uint32_t cap = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1u << 20, (T.nbases + Q.nbases) / 2), 0x7fffffffull);
    DevBuf<uint64_t> s0, s1;
    unsigned long long nsurv = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        s0.alloc(cap);
        MB2_CUDA(cudaMemsetAsync(counters.get(), 0, 4 * sizeof(unsigned long long), cx.stream));
        seed_scan(T, Q, tab, GENOME_END_PAD, (uint32_t)Q.G - GENOME_END_PAD, p, s0.get(), cap, counters.get());
        MB2_CUDA(cudaMemcpyAsync(&nsurv, counters.get() + CNT_SURV, sizeof(nsurv), cudaMemcpyDeviceToHost, cx.stream));
        MB2_CUDA(cudaStreamSynchronize(cx.stream));
        if (nsurv <= cap) break;
        MB2_REQUIRE(nsurv < 0x7fffffffull && attempt == 0, -3, "align: too many surviving seed hits for one pass");
        cap = (uint32_t)nsurv;
    }
    s1.alloc(nsurv ? nsurv : 1);
    find_hsps(T, Q, s0.get(), s1.get(), (uint32_t)nsurv, p, hsps, counters.get());
    if (h_counters) {
        MB2_CUDA(cudaMemcpyAsync(h_counters, counters.get(), CNT_N * sizeof(unsigned long long), cudaMemcpyDeviceToHost, cx.stream));
        MB2_CUDA(cudaStreamSynchronize(cx.stream));
    }
}

// Full pipeline for one strand orientation of Q: alignments in strand-local, scaffold-local coordinates.
void align_strand(const Genome& T, const Genome& Q, const AlignParams& p, const int32_t* h_same_q, AlnSet& alns, unsigned long long* h_counters) {
    Ctx& cx = ctx();
    HspSet hsps;
    unsigned long long c1[CNT_N];
    align_hsps(T, Q, p, hsps, c1);
    uint32_t maxlen = 0;
    for (int s = 0; s < T.nscaf; s++) maxlen = std::max(maxlen, T.len[s]);
    for (int s = 0; s < Q.nscaf; s++) maxlen = std::max(maxlen, Q.len[s]);
    int lb = 1; while (lb < 32 && (maxlen >> lb)) lb++;
    lb += 1;   // e = s + len can reach maxlen exactly
    int tb = 1; while (tb < 33 && (((uint64_t)T.nscaf * (uint64_t)Q.nscaf) >> tb)) tb++;
    DevBuf<uint8_t> in_chain;
    if (p.chain) chain_hsps(hsps, lb, tb, in_chain);
    else {
        in_chain.alloc(hsps.n ? hsps.n : 1);
        if (hsps.n) MB2_CUDA(cudaMemsetAsync(in_chain.get(), 1, hsps.n, cx.stream));
    }
    DevBuf<unsigned long long> counters(CNT_N);
    MB2_CUDA(cudaMemsetAsync(counters.get(), 0, CNT_N * sizeof(unsigned long long), cx.stream));
    gapped_extend(T, Q, hsps, in_chain, p, h_same_q, alns, counters.get());
    unsigned long long c2[CNT_N];
    MB2_CUDA(cudaMemcpyAsync(c2, counters.get(), sizeof(c2), cudaMemcpyDeviceToHost, cx.stream));
    MB2_CUDA(cudaStreamSynchronize(cx.stream));
    if (h_counters) {
        for (int k = 0; k < CNT_N; k++) h_counters[k] = c1[k];
        h_counters[CNT_GAPPED_CELLS] = c2[CNT_GAPPED_CELLS];
        h_counters[CNT_ANCHORS] = c2[CNT_ANCHORS];
        h_counters[CNT_ALNS] = alns.n;
    }
}

}  // namespace mb2
