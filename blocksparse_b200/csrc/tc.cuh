// tcgen05 kernel families (filled in below); returns TC_NOT_APPLICABLE when the
// configuration has no tensor-core kernel so that the caller uses the CUDA-core family.
#pragma once
#include "common.cuh"

namespace bsmm {
constexpr int TC_NOT_APPLICABLE = -1000;

inline int tc_xprop(int, int, int, int, const int32_t*, int, int, int, const void*, const void*, void*, int,
                    const float*, const int32_t*, int, cudaStream_t) { return TC_NOT_APPLICABLE; }
inline int tc_updat(int, int, int, int, const int32_t*, int, int, int, const void* const*, const void* const*, int,
                    void*, int, float, float, const float*, int, const int32_t*, int, cudaStream_t) { return TC_NOT_APPLICABLE; }
inline int tc_bst_nt(int, int, int, const int32_t*, int, int, const void*, const void*, void*, int, int, int, int, int,
                     cudaStream_t) { return TC_NOT_APPLICABLE; }
inline int tc_bst_xn(int, int, int, int, const int32_t*, int, int, int, const void*, const void*, void*, int, int, int,
                     int, int, cudaStream_t) { return TC_NOT_APPLICABLE; }
}  // namespace bsmm
