// float64 with algebraic shortcuts and FMA contraction (TFG_F64_FAST)
#include "tfg_run.cuh"
namespace tfg {
cudaError_t launch_run_fast(const RunParams<double>& p, bool rec, bool agg, bool vol, cudaStream_t stream) {
  return launch_run<FastF64>(p, rec, agg, vol, stream);
}
}  // namespace tfg
