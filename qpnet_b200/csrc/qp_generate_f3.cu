// placeholder until the tcgen05 generator lands: reports "unsupported" so the dispatch uses fold2 / generic
#include "qp_gen_common.cuh"
namespace qp {
int f3_generate(const QpArch*, const float* const*, const QpGenerateArgs*, void*, size_t, cudaStream_t) { return 1; }
size_t f3_workspace_bytes(const QpArch*, int, int) { return 0; }
bool f3_supported(const QpArch*, int) { return false; }
int f3_trace_copy(const QpArch*, int, int, void*, size_t, long long*, int, cudaStream_t) { return 0; }
}  // namespace qp
