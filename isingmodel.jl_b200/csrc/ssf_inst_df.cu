#include "ssf_inst.cuh"
namespace isb {
cudaError_t launch_ssf_df(const SsfParams &p, int npl, bool list, bool tma, int cl, int grid, int threads, size_t smem,
                          cudaStream_t st) {
    return launch_pair<double, float>(p, npl, list, tma, cl, grid, threads, smem, st);
}
}  // namespace isb
