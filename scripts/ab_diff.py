"""Run the fixed workload under library $TFG_LIBRARY and save the final state; with --against FILE compare bit for bit."""
import argparse, os, sys, torch
sys.path.insert(0, '.')
from topoflow_glacier_b200.engine import MeltEngine
from topoflow_glacier_b200.config import default_constants
from topoflow_glacier_b200.synthetic import synthetic_cells
ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=2097150); ap.add_argument("--steps", type=int, default=24)
ap.add_argument("--launches", type=int, default=2); ap.add_argument("--start", type=int, default=6000)
ap.add_argument("--save"); ap.add_argument("--against")
a = ap.parse_args()
dev = torch.device("cuda", 0)
tabs = synthetic_cells(a.cells, 4096, dev); raw = tabs.pop("raw")
eng = MeltEngine(None, default_constants(), "2012100100", zones=[-8.0], mode="f64_fast", horizon_steps=a.steps * (a.launches + 1) + a.start, device_statics=tabs)
f = torch.empty(a.steps, 5, a.cells, dtype=eng.dtype, device=dev)
eng.step_index = a.start
eng.synth_forcing(f, a.start, a.steps, raw["elev"].to(eng.dtype), 7, 1)
hist = []
for i in range(a.launches):
    rec = eng.run(f, a.steps, record=("h_swe", "SM", "Q_sum", "Eccs", "P_snow", "RH", "albedo", "n", "snow3day", "T_surf", "Qn_SW", "Qh", "Qe", "Qn_LW"))
    hist.append({k: v.clone() for k, v in rec.items()})
torch.cuda.synchronize()
if a.save:
    torch.save({"state": eng.state.cpu(), "hist": [{k: v.cpu() for k, v in h.items()} for h in hist]}, a.save)
if a.against:
    ref = torch.load(a.against)
    bad = (eng.state.cpu().view(torch.int64) != ref["state"].view(torch.int64)).any(dim=0).nonzero().flatten()
    print("cells that differ:", bad.tolist()[:20], "of", a.cells)
    for c in bad.tolist()[:3]:
        for li, (h, hr) in enumerate(zip(hist, ref["hist"])):
            for k in h:
                x, y = h[k][:, c].cpu(), hr[k][:, c]
                d = (x.view(torch.int64) != y.view(torch.int64)).nonzero().flatten()
                if len(d):
                    t = int(d[0]); print(f"cell {c} launch {li} {k}: first differs at step {t}: {x[t].item()!r} vs {y[t].item()!r}")
        print("forcing at cell", c, f[:, :, c].cpu()[:3].tolist())
