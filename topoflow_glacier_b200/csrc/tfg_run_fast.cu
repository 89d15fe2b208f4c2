// float64 with algebraic shortcuts and written-out FMAs (TFG_F64_FAST)
#ifndef TFG_PIPELINE   // 1: experimental software-pipelined time loop (tfg_pipe.cuh, needs -DTFG_SPLIT_STEP=1 for bit-identical
#define TFG_PIPELINE 0 //    results across routes; measured slower); 0: single-stream loop of tfg_run.cuh (default)
#endif
#ifdef TFG_LEAN_W2
// Experimental two-cells-per-thread kernel with 128-bit forcing / state / window access (tfg_lean.cuh).  Bit-identical
// results, 8 % fewer instructions per cell-step, but 168 registers (3 blocks per SM) and a hot loop of 26 KB that misses
// the instruction cache: 22.5 G cell-steps/s against 27.9 G for the one-cell kernel (profiles/r2_w2_experiment.txt).
#include "tfg_lean.cuh"
#else
#include "tfg_pipe.cuh"
#endif
namespace tfg {
cudaError_t launch_run_fast(const RunParams<double>& p, bool rec, bool agg, bool vol, cudaStream_t stream) {
#ifdef TFG_LEAN_W2
  return launch_run_lean(p, rec, agg, vol, stream);
#else
  // one-step launches re-sum the snowfall window exactly (the literal update()), TMA staging is an option of the
  // single-stream kernel; both evaluate the same two device functions, so results do not depend on the route
  if (TFG_PIPELINE && !p.exact_ring && !p.use_tma) return launch_run_pipe(p, rec, agg, vol, stream);
  return launch_run<FastF64>(p, rec, agg, vol, stream);
#endif
}
}  // namespace tfg
