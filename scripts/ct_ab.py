"""A/B of the column-term path: 16 777 216 cells sharing 4096 forcing columns, 128 steps per launch, exact aggregates.

    TFG_LIBRARY=<variant.so> python scripts/ct_ab.py [--off] [--start STEP]
"""
import argparse
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

ap = argparse.ArgumentParser()
ap.add_argument("--off", action="store_true", help="TFG_OPT_COLUMN_TERMS = 0: the per-cell evaluation")
ap.add_argument("--cells", type=int, default=1 << 24)
ap.add_argument("--start", type=int, default=2400)
ap.add_argument("--mode", default="f64_fast", choices=["f64_fast", "f64"])
a = ap.parse_args()
if a.off:
    os.environ["TFG_COLUMN_TERMS"] = "0"

import bench  # noqa: E402
from topoflow_glacier_b200.config import default_constants  # noqa: E402
from topoflow_glacier_b200.engine import MeltEngine  # noqa: E402
from topoflow_glacier_b200.sharding import BasinAggregates  # noqa: E402
from topoflow_glacier_b200.synthetic import synthetic_cells  # noqa: E402

dev = torch.device("cuda:0")
T, NB = 128, 4096
tabs = synthetic_cells(a.cells, 4096, dev)
elev = tabs.pop("raw")["elev"]
basin = (torch.arange(a.cells, device=dev, dtype=torch.int64) // (-(-a.cells // NB))).to(torch.int32)
eng = MeltEngine(None, default_constants(), "2012100100", dt_hours=1, zones=[-8.0], basin_id=basin, n_basin=NB, mode=a.mode,
                 horizon_steps=a.start + 8 * T + 64, device_statics=tabs)
eng.step_index = a.start
agg = BasinAggregates(T, NB, device=dev, exponents=eng.agg_exponents())
f = torch.empty(T, 5, eng.N, dtype=eng.dtype, device=dev)
eng.synth_forcing(f, a.start, T, elev.to(eng.dtype), seed=20121001)
fc = f[:, :, :NB].contiguous()
del f
eng.set_forcing_map(basin % NB, NB)
for _ in range(2):
    eng.run(fc, T, basin_agg=agg.zero())
k, _ = bench.time_launches(eng, fc, T, agg.zero, agg.reduce, 3, 1, dev)
print(f"{a.mode} column_terms={eng.column_term_launches > 0}: {k:.3f} ms  {eng.N * T / k / 1e6:.2f} G cell-steps/s  "
      f"state {int(eng.state.view(torch.int64).sum().item())} ring {int(eng.ring.view(torch.int64).sum().item())}")
