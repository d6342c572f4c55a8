// TEMPORARY link stubs for the tcgen05 entry points (replaced by dense_tc.cu / retrieval_tc.cu / topk_tc.cu).
#include "common.cuh"
namespace tt {
int tc_retrieval_fwd(const void*, const void*, int64_t, int64_t, int64_t, float, int64_t, const float*, const float*, const int64_t*, float*, float*, float*, void*, int64_t, cudaStream_t) { return set_error(TT_ERR_UNSUPPORTED, "bf16 retrieval path not built"); }
int tc_retrieval_bwd(const void*, const void*, const void*, const void*, int64_t, int64_t, int64_t, float, int64_t, const float*, const float*, const int64_t*, const float*, float, float*, float*, uint16_t*, uint16_t*, uint16_t*, uint16_t*, void*, int64_t, cudaStream_t) { return set_error(TT_ERR_UNSUPPORTED, "bf16 retrieval path not built"); }
int64_t tc_retrieval_workspace_bytes(int64_t, int64_t, int64_t) { return 256; }
int tc_topk(const void*, const void*, int64_t, int64_t, int64_t, int, int64_t, const int64_t*, float*, int64_t*, void*, int64_t, cudaStream_t) { return set_error(TT_ERR_UNSUPPORTED, "bf16 top-k path not built"); }
int tc_topk_num_splits(int64_t, int64_t, int64_t, int) { return 1; }
}
