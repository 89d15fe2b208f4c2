// float64, reference-order IEEE arithmetic (TFG_F64_STRICT)
#include "tfg_run.cuh"
namespace tfg {
cudaError_t launch_run_strict(const RunParams<double>& p, bool rec, bool agg, bool vol, cudaStream_t stream) {
  return launch_run<StrictF64>(p, rec, agg, vol, stream);
}
}  // namespace tfg
