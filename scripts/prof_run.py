"""Small fixed workload for ncu: synthetic cells, one fused launch per mode (after a warm-up launch)."""
import argparse, sys, torch
sys.path.insert(0, '.')
from topoflow_glacier_b200.engine import MeltEngine
from topoflow_glacier_b200.config import default_constants
from topoflow_glacier_b200.synthetic import synthetic_cells

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="f64_fast")
ap.add_argument("--cells", type=int, default=1 << 21)
ap.add_argument("--steps", type=int, default=16)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--agg", type=int, default=0)
ap.add_argument("--storm", type=int, default=1)
ap.add_argument("--start", type=int, default=2400, help="absolute first step (2400 = January: snow everywhere)")
ap.add_argument("--sum", action="store_true", help="print bit-pattern checksums of the final state / window")
a = ap.parse_args()
dev = torch.device("cuda", 0)
tabs = synthetic_cells(a.cells, 4096, dev)
raw = tabs.pop("raw")
basin = (torch.arange(a.cells, device=dev) // max(1, a.cells // 4096)).to(torch.int32) if a.agg else None
eng = MeltEngine(None, default_constants(), "2012100100", zones=[-8.0], mode=a.mode, horizon_steps=a.steps * (a.launches + 1) + a.start,
                 device_statics=tabs, basin_id=basin, n_basin=4096 if a.agg else 0)
f = torch.empty(a.steps, 5, a.cells, dtype=eng.dtype, device=dev)
eng.step_index = a.start
eng.synth_forcing(f, a.start, a.steps, raw["elev"].to(eng.dtype), 7, a.storm)
agg = torch.zeros(a.steps, 4096, 3, dtype=torch.float64, device=dev) if a.agg == 1 else None
if a.agg == 2:  # order-independent fixed-point accumulators (TFG_OPT_EXACT_AGG)
    from topoflow_glacier_b200.sharding import BasinAggregates
    agg = BasinAggregates(a.steps, 4096, device=dev, exponents=eng.agg_exponents()).accumulator
torch.cuda.synchronize()
for i in range(a.launches):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); eng.run(f, a.steps, basin_agg=agg); e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    print(f"{a.mode} launch {i}: {ms:.3f} ms  {a.cells * a.steps / ms / 1e6:.2f} G cell-steps/s")
if a.sum:  # order-independent checksums of the raw bits: equal sums <=> (practically) bit-identical results
    it = torch.int64 if eng.dtype == torch.float64 else torch.int32
    print("checksum state", eng.state.view(it).to(torch.int64).sum(dim=1).tolist())
    print("checksum ring", int(eng.ring.view(it).to(torch.int64).sum().item()))
    if agg is not None:
        print("checksum agg", float(agg.double().sum().item()) if agg.dtype != torch.int64 else int(agg.sum().item()))
