// float32 with MUFU intrinsics (TFG_F32)
#include "tfg_run.cuh"
namespace tfg {
cudaError_t launch_run_f32(const RunParams<float>& p, bool rec, bool agg, bool vol, cudaStream_t stream) {
  return launch_run<FastF32>(p, rec, agg, vol, stream);
}
}  // namespace tfg
