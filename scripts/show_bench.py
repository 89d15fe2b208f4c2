"""One-screen summary of a bench.py JSON line: python scripts/show_bench.py FILE.json"""
import json
import sys

txt = open(sys.argv[1]).read().strip()
try:
    d = json.loads(txt)                      # pretty-printed copy under profiles/
except json.JSONDecodeError:
    d = json.loads(txt.splitlines()[-1])     # the one-line original
if d.get("impl") == "reference":
    print("reference", d.get("value"), d.get("cpu_baseline"))
    sys.exit(0)
G = 1e9
print("value %.2f G  frac %.4f  n_gpus %d  ms/step %.2f  launches %s  wall %s" % (
    d["value"] / G, d["roofline"]["frac"], d["n_gpus"], d["ms_per_step"], d["gpu_launches"], d.get("wall_s")), d["clocks"])
e = d["e2e"]
print("e2e %.2f G  bound %s  %s" % (e["value"] / G, e.get("bound"), {k: round(v, 1) for k, v in e["stage_ms_per_block"].items() if v}))
c = d.get("e2e_catchment_forcing")
if c:
    print("catchment forcing: e2e %.2f G  launch %.2f G  column_terms %s" % (
        c["value"] / G, c.get("launch_cell_steps_per_s_per_gpu", 0) / G, c.get("column_terms")))
if d.get("coherent_weather"):
    print("coherent weather %.2f G" % (d["coherent_weather"]["cell_steps_per_s_per_gpu"] / G))
s = d.get("strong_regional")
if s:
    print("strong regional %.2f G  steps %d x %d  %s" % (s["value"] / G, s["steps"], s["timesteps_per_step"], s["clocks"]))
for k, v in (d.get("modes") or {}).items():
    print("mode %s %.2f G  frac %.4f  steps %d  %s" % (k, v["value"] / G, v["roofline"]["frac"], v["steps"], v["clocks"]))
if d.get("cpu_baseline"):
    print("cpu", d["cpu_baseline"]["kind"], d["cpu_baseline"]["value"], "port", (d.get("cpu_baseline_port") or {}).get("value"))
