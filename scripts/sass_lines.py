"""Attribute the SASS of one melt-kernel instantiation to lines of tfg_physics.cuh / tfg_run.cuh.

    python scripts/sass_lines.py [--kernel FastF64ELb0ELb0ELb0ELb0] [--obj tfg_run_fast] [--ncu-csv FILE]

Static view: `nvdisasm -gi` of the cubin inside the object file gives every instruction its inline chain; an
instruction is booked on the outermost frame that lies in tfg_physics.cuh (else tfg_run.cuh).  With --ncu-csv (the
`ncu -i X.ncu-rep --page source --csv` export of a capture of the SAME binary) the per-instruction "Instructions
Executed" counts are joined by code offset, which turns the table into the dynamic instruction mix per source line.
"""
from __future__ import annotations

import argparse
import collections
import csv
import re
import subprocess
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
FP64 = ("DFMA", "DMUL", "DADD", "DSETP", "MUFU.RCP64H", "MUFU.RSQ64H", "F2F.F64", "I2F.F64", "F2I.F64", "DMNMX")


def disasm(obj: str) -> str:
    o = ROOT / "topoflow_glacier_b200" / "lib" / "obj" / f"{obj}.o"
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", str(o)], cwd=d, check=True, capture_output=True)
        cubin = next(Path(d).glob("*.cubin"))
        return subprocess.run(["nvdisasm", "-gi", "-c", str(cubin)], check=True, capture_output=True, text=True).stdout


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernel", default="FastF64ELb0ELb0ELb0ELb0")
    ap.add_argument("--obj", default="tfg_run_fast")
    ap.add_argument("--ncu-csv")
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("--via", help="only instructions inlined through a tfg_run.cuh line containing this text")
    ap.add_argument("--ops", action="store_true", help="also print the opcode histogram of the selected instructions")
    ap.add_argument("--by-samples", action="store_true", help="rank source lines by warp-stall samples (time) instead of FP64 count")
    a = ap.parse_args()

    text = disasm(a.obj)
    run_src = (ROOT / "topoflow_glacier_b200" / "csrc" / "tfg_run.cuh").read_text().splitlines()
    via_lines = {i + 1 for i, l in enumerate(run_src) if a.via and a.via in l}
    sect = None
    chain: list[tuple[str, int]] = []
    fresh = True
    insts = []  # (offset, opcode, booked (file, line))
    for ln in text.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            sect = m.group(1)
            continue
        if sect is None or a.kernel not in sect:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            if fresh:
                chain = []
                fresh = False
            chain.append((Path(m.group(1)).name, int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            fresh = True
            if a.via and not any(f == "tfg_run.cuh" and l in via_lines for f, l in chain):
                continue
            book = None
            for f, l in reversed(chain):  # outermost first
                if f == "tfg_physics.cuh":
                    book = (f, l)
                    break
            if book is None:
                for f, l in reversed(chain):
                    if f == "tfg_run.cuh" and "cell_step<" not in run_src[l - 1]:
                        book = (f, l)
                        break
            if book is None:
                book = chain[-1] if chain else ("?", 0)
            insts.append((int(m.group(1), 16), m.group(2), book))

    dyn = smp = None
    if a.ncu_csv:
        rows = list(csv.reader(open(a.ncu_csv)))
        hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
        col = rows[hdr].index("Instructions Executed")
        base = int(rows[hdr + 1][0], 16)
        dyn = {int(r[0], 16) - base: int(r[col]) for r in rows[hdr + 1:] if r and r[0].startswith("0x")}
        scol = rows[hdr].index("# Samples")
        smp = {int(r[0], 16) - base: int(r[scol]) for r in rows[hdr + 1:] if r and r[0].startswith("0x")}

    tot = collections.Counter()
    stall = collections.Counter()
    fp = collections.Counter()
    ops = collections.Counter()
    for off, op, book in insts:
        w = dyn.get(off, 0) if dyn is not None else 1
        tot[book] += w
        if smp is not None:
            stall[book] += smp.get(off, 0)
        ops[op.split(".")[0]] += w
        if op.startswith(FP64):
            fp[book] += w
    all_t, all_f = sum(tot.values()), sum(fp.values())
    src = {}
    for f in ("tfg_physics.cuh", "tfg_run.cuh"):
        src[f] = (ROOT / "topoflow_glacier_b200" / "csrc" / f).read_text().splitlines()
    print(f"kernel {a.kernel}: {len(insts)} SASS instructions; {'dynamic' if dyn else 'static'} totals: all={all_t} fp64={all_f}")
    ranked = fp if all_f > 0.05 * all_t else tot  # float32 kernels have (almost) no FP64 instructions: rank by all
    if a.by_samples and smp is not None:
        ranked = stall
    all_s = max(sum(stall.values()), 1)
    for book, _ in sorted(ranked.items(), key=lambda kv: -kv[1])[: a.top]:
        n = fp[book]
        f, l = book
        s = src.get(f, [""] * (l + 1))[l - 1].strip()[:90] if f in src else ""
        extra = f" | {100 * stall[book] / all_s:5.1f}% of warp-stall samples" if smp is not None else ""
        print(f"{100 * n / max(all_f, 1):5.1f}% fp64 {n:>12} | all {tot[book]:>12} ({100 * tot[book] / max(all_t, 1):4.1f}%){extra} | {f}:{l}  {s}")
    if a.ops:
        print("opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(25)))




if __name__ == "__main__":
    main()
