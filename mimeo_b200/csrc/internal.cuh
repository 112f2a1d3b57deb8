// internal.cuh -- declarations shared between the translation units of libmimeo_b200.
#pragma once
#include <vector>

#include "common.cuh"

namespace mb2 {

// ---- coverage.cu
struct CoverageResult {
    DevBuf<int32_t> chrom, start, end;
    uint64_t n = 0;
};
void coverage_segments_device(const int32_t* d_chrom, const int32_t* d_start, const int32_t* d_end, uint64_t nhits,
                              const int64_t* h_sizes, int nchrom, int min_cov, int min_len, CoverageResult& res);

}  // namespace mb2
