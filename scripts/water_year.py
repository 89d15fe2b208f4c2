"""BASELINE configs[3] in full: the synthetic 4096 x 4096 raster over ONE WATER YEAR of hourly forcing (8 760 steps,
69 launches of <= 128 steps).  Forcing chunks are generated on the device (Philox, keyed by cell and absolute step)
right before each launch; the kernel of every chunk is timed with CUDA events.  Prints the per-chunk range and the
year's mean rate as one JSON line (bench.py times the first chunk only).

    python scripts/water_year.py [--mode f64_fast|f64|f32] [--cells N]
"""
import argparse, json, sys, time
import torch
sys.path.insert(0, '.')
from topoflow_glacier_b200.config import default_constants
from topoflow_glacier_b200.engine import MeltEngine
from topoflow_glacier_b200.sharding import BasinAggregates
from topoflow_glacier_b200.synthetic import synthetic_cells

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="f64_fast")
ap.add_argument("--cells", type=int, default=4096 * 4096)
ap.add_argument("--steps", type=int, default=8760)
ap.add_argument("--chunk", type=int, default=128)
ap.add_argument("--columns", type=int, default=0,
                help="> 0: catchment forcing -- that many forcing series shared by consecutive cells (tfg_bind_forcing_map; the "
                     "fast float64 mode then evaluates the forcing-only terms once per column, TFG_OPT_COLUMN_TERMS)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
NB = 4096
tabs = synthetic_cells(a.cells, 4096, dev)
elev = tabs.pop("raw")["elev"]
basin = (torch.arange(a.cells, device=dev) // max(1, a.cells // NB)).clamp_(max=NB - 1).to(torch.int32)
eng = MeltEngine(None, default_constants(), "2012100100", zones=[-8.0], mode=a.mode, horizon_steps=a.steps + 1,
                 device_statics=tabs, basin_id=basin, n_basin=NB)
elev = elev.to(eng.dtype)
forcing = torch.empty(a.chunk, 5, a.cells, dtype=eng.dtype, device=dev)
fcol = None
if a.columns > 0:   # the series of the first `columns` cells serve as the catchments' met series
    per = -(-a.cells // a.columns)
    eng.set_forcing_map((torch.arange(a.cells, device=dev) // per).to(torch.int32), a.columns)
    fcol = torch.empty(a.chunk, 5, a.columns, dtype=eng.dtype, device=dev)
agg = BasinAggregates(a.chunk, NB, device=dev, exponents=eng.agg_exponents())
runoff = torch.zeros(NB, dtype=torch.float64, device=dev)   # m3 per basin over the year
rates, t_wall = [], time.perf_counter()
for t0 in range(0, a.steps, a.chunk):
    n = min(a.chunk, a.steps - t0)
    eng.synth_forcing(forcing, t0, n, elev, seed=20121001)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tgt = agg.zero()
    if fcol is not None:
        fcol[:n].copy_(forcing[:n, :, :a.columns])
    s.record(); eng.run(forcing if fcol is None else fcol, n, basin_agg=tgt); e.record()
    agg.reduce()
    runoff += agg.buffer[:n, :, 0].sum(dim=0) * 3600.0
    torch.cuda.synchronize()
    rates.append(a.cells * n / (s.elapsed_time(e) * 1e-3))
wall = time.perf_counter() - t_wall
kernel_s = sum(a.cells * min(a.chunk, a.steps - t0) / r for t0, r in zip(range(0, a.steps, a.chunk), rates))
print(json.dumps({
    "workload": f"{a.cells} cells x {a.steps} hourly steps, mode {a.mode}"
                + (f", {a.columns} shared forcing columns (column terms: {eng.column_term_launches} launches)" if a.columns else ""),
    "launches": len(rates),
    "cell_steps_per_s_year_mean": a.cells * a.steps / kernel_s, "kernel_seconds": kernel_s, "wall_seconds_incl_forcing_synthesis": wall,
    "chunk_rate_min": min(rates), "chunk_rate_max": max(rates),
    "chunk_rates_G": [round(r / 1e9, 2) for r in rates],
    "swe_m3_end": float(agg.buffer[min(a.chunk, a.steps - (a.steps - 1) // a.chunk * a.chunk) - 1, :, 1].sum()),
    "runoff_m3_year": float(runoff.sum()), "aggregate_contributions_left_out": agg.n_left_out}))
