// float64 with algebraic shortcuts and written-out FMAs (TFG_F64_FAST)
#ifdef TFG_LEAN_W2
// Experimental two-cells-per-thread kernel with 128-bit forcing / state / window access (tfg_lean.cuh).  Bit-identical
// results, 8 % fewer instructions per cell-step, but 168 registers (3 blocks per SM) and a hot loop of 26 KB that misses
// the instruction cache: 22.5 G cell-steps/s against 27.9 G for the one-cell kernel (profiles/r2_w2_experiment.txt).
#include "tfg_lean.cuh"
#else
#include "tfg_run.cuh"
#endif
namespace tfg {
cudaError_t launch_run_fast(const RunParams<double>& p, bool rec, bool agg, bool vol, cudaStream_t stream) {
#ifdef TFG_LEAN_W2
  return launch_run_lean(p, rec, agg, vol, stream);
#else
  return launch_run<FastF64>(p, rec, agg, vol, stream);
#endif
}
}  // namespace tfg
