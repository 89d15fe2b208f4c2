"""Copy the reference's DATA fixtures (no code) into this repository: the five catchment configurations and the one
forcing sample that exists upstream.

    python scripts/import_reference_data.py        (needs /root/reference; the outputs are committed)

* `config/cat-*.yaml`  <- `/root/reference/config/*.yaml`, values verbatim (three of them keep their unquoted integer
  `start_time` / `end_time`, which the reference's own schema rejects and this schema coerces).  Two lines change:
  `forcing_file` points at the sample below -- the per-catchment CSVs named upstream (`data/cat-30627xx.csv`) are
  large blobs that are not part of the upstream checkout (`.MISSING_LARGE_BLOBS`) -- and `tz_name` (extension key) names
  the zone the reference would look up with `timezonefinder`.
* `tests/data/sample-cat-3062920.csv`, `tests/data/output_m_total.npy` <- the reference's test sample and its golden
  vector (`tests/integration_test.py:81,151`).
* `tests/data/mock-forcing-june2012.csv` <- the 168-row forcing embedded in the reference's `tests/conftest.py:10-178`,
  whose columns come in a different order (ingestion is header-keyed).
"""

import re
import shutil
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


def main():
    (ROOT / "config").mkdir(exist_ok=True)
    (ROOT / "tests" / "data").mkdir(exist_ok=True)
    for src in sorted((REF / "config").glob("*.yaml")):
        out = []
        for line in src.read_text().splitlines():
            if line.startswith("forcing_file:"):
                out.append(f"forcing_file: tests/data/sample-cat-3062920.csv  # upstream: {line.split(':', 1)[1].strip()} "
                           "(not in the upstream checkout)")
            else:
                out.append(line)
        out.append("tz_name: America/Los_Angeles  # extension key: replaces the reference's timezonefinder lookup")
        (ROOT / "config" / src.name).write_text("\n".join(out) + "\n")
        print("wrote", ROOT / "config" / src.name)
    for name in ("sample-cat-3062920.csv", "output_m_total.npy"):
        shutil.copy(REF / "tests" / "data" / name, ROOT / "tests" / "data" / name)
        print("copied tests/data/" + name)
    text = (REF / "tests" / "conftest.py").read_text()
    m = re.search(r'"""(Time,RAINRATE.*?)"""', text, re.S)
    if m:
        rows = [ln.strip() for ln in m.group(1).strip().splitlines()]
        (ROOT / "tests" / "data" / "mock-forcing-june2012.csv").write_text("\n".join(rows) + "\n")
        print("wrote tests/data/mock-forcing-june2012.csv")


if __name__ == "__main__":
    main()
