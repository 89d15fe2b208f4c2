"""Condense gpurun_out/ artefacts into the tracked profiles/ directory (round-tagged)."""
import collections, csv, io, json, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT, SRC = ROOT / "profiles", ROOT / "gpurun_out"
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT.mkdir(exist_ok=True)
# workload of the ncu --set full captures: round 1 used the small fixed workload, round 2 the BENCH shape
CELLS, STEPS = (2097152, 24) if tag == "r1" else (16777216, 128)
WORK = f"scripts/prof_run.py: {CELLS:,} cells x {STEPS} steps".replace(",", " ")


def launches(csv_path: Path, dst: Path):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0][:110]
        if "run_kernel" in r[4]:
            name = r[4].split("(RunParams")[0]
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += float(r[14].replace(",", "")) / (1e3 if r[13] == "us" else 1e6 if r[13] == "ns" else 1.0)  # -> ms
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({csv_path.name}); durations are cold-cache, serialised: compare SHARES\n")
        f.write("kernel,launches,total_ms,share\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{ms:.3f},{ms / tot:.4f}\n")


def parity(dst: Path):
    rep = json.loads((SRC / "parity_report.json").read_text())
    with open(dst, "w") as f:
        f.write("# worst |gpu-oracle| / tolerance per case and mode (tolerance = 1e-12*|ref| + atol, tests/helpers.py)\n")
        f.write("case/mode,worst_quantity,err_over_tol,max_rel_err_of_that_quantity,extras\n")
        for tagk, r in sorted(rep.items()):
            q = {k: v for k, v in r.items() if isinstance(v, dict)}
            k = max(q, key=lambda x: q[x]["err_over_tol"])
            extras = {a: b for a, b in r.items() if not isinstance(b, dict)}
            f.write(f"{tagk},{k},{q[k]['err_over_tol']:.3g},{q[k]['max_rel']:.3g},\"{extras}\"\n")


if (SRC / f"launches_{tag}.csv").exists():
    launches(SRC / f"launches_{tag}.csv", OUT / f"{tag}_launches.csv")
if (SRC / "parity_report.json").exists():
    parity(OUT / f"{tag}_parity.csv")
tcsv = SRC / f"traffic_{tag}.csv"
if tcsv.exists():
    rows = [r for r in csv.reader(open(tcsv)) if len(r) > 14 and r[0].isdigit()]
    tot = sum(float(r[14]) for r in rows if r[12].startswith("dram__bytes"))
    dur = [float(r[14]) for r in rows if r[12].startswith("gpu__time")]
    (OUT / "traffic.json").write_text(json.dumps({
        "f64_fast:16777216x128": tot, "_source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum, {tcsv.name}, one launch "
        "of tfg::run_kernel<FastF64,0,1,1> at the bench configuration", "_kernel_ns_under_ncu": dur}, indent=1) + "\n")
for rep in sorted(SRC.glob(f"prof_{tag}_*.ncu-rep")):
    cs = sys.argv[2] if len(sys.argv) > 2 else str(CELLS * STEPS)
    txt = subprocess.run([sys.executable, str(ROOT / "scripts" / "ncu_summary.py"), str(rep), cs], capture_output=True, text=True).stdout
    (OUT / f"{tag}_{rep.stem}.txt").write_text(f"# ncu --set full, {rep.name}; workload {WORK}\n" + txt)
# instruction mix of the hot kernel -> profiles/kernel_mix.json (bench.py turns it into the FP64-pipe ceiling)
mix = {}
for mode in ("f64_fast", "f64", "f32"):
    rep = SRC / f"prof_{tag}_{mode}.ncu-rep"
    if not rep.exists():
        continue
    src = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    iS, iE = h.index("Source"), h.index("Instructions Executed")
    import re
    tot = fp64 = 0
    for r in rows[2:]:
        mm = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
        if mm:
            tot += int(r[iE])
            fp64 += int(r[iE]) if mm.group(1) in ("DFMA", "DMUL", "DADD", "DSETP") else 0
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    m = dict(zip(rr[0], rr[2]))
    warp_steps = CELLS * STEPS / 32
    if mode == "f64_fast" and tag != "r1":   # DRAM traffic of the benched instantiation at the bench shape, same capture
        def num(key):
            v, u = m.get(key, "0"), dict(zip(rr[0], rr[1])).get(key, "byte")
            return float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
        (OUT / "traffic.json").write_text(json.dumps({
            f"f64_fast:{CELLS}x{STEPS}": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
            "_source": f"ncu --set full ({rep.name}): dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of "
                       "tfg::run_kernel<FastF64,0,1,1,0> at the bench shape", "_kernel": m.get("Kernel Name", "")}, indent=1) + "\n")
    mix[mode] = {"warp_inst_per_warp_step": tot / warp_steps, "fp64_warp_inst_per_warp_step": fp64 / warp_steps,
                 "fp64_pipe_active_pct": float(m.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 0)),
                 "issue_active_pct": float(m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0)),
                 "kernel": m.get("Kernel Name", ""),
                 "source": f"ncu --set full, {rep.name}, {WORK}"}
if mix:
    (OUT / "kernel_mix.json").write_text(json.dumps(mix, indent=1) + "\n")
for b in sorted(SRC.glob(f"bench_{tag}_*.json")):   # only this round's lines (gpurun_out/ keeps older rounds' files)
    try:
        line = json.loads(b.read_text().strip().splitlines()[-1])
        (OUT / f"{tag}_{b.stem}.json").write_text(json.dumps(line, indent=1) + "\n")
    except Exception:
        pass
print(sorted(p.name for p in OUT.iterdir()))
