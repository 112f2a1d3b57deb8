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
        seed_scan(T, Q, tab, GENOME_PAD, (uint32_t)Q.G, p, s0.get(), cap, counters.get());
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

}  // namespace mb2
