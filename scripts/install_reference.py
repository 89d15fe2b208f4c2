"""Put the UNMODIFIED reference where `bench.py --impl reference` can import it on the GPU box: `baseline/_ref/`.

    python scripts/install_reference.py          (run by __graft_entry__.build() when /root/reference exists)

`pip install --target baseline/_ref /root/reference` is not possible offline (the build backend hatchling + hatch-vcs
and the runtime dependencies bmipy / timezonefinder / pyprojroot are not in the wheelhouse), so this script does what
that install would do for a pure-Python package: it copies `src/topoflow_glacier` as it is, writes the `_version.py`
hatch-vcs would generate, and adds three stub modules for the missing dependencies:

* `bmipy.Bmi`            -- an empty base class (the reference only inherits from it, bmi_base.py:1,7);
* `timezonefinder`       -- `TimezoneFinder().timezone_at()` answers `TFG_REF_TZ` (default America/Los_Angeles, the zone
                            of all shipped catchments); the DST arithmetic stays the reference's own zoneinfo code;
* `pyprojroot.here`      -- the directory that holds this copy.

`baseline/_ref/` is git-ignored (nothing of the reference enters the history) but travels to the GPU box with the
repository snapshot.  The sample forcing the reference's own test uses is copied next to it.
"""

from __future__ import annotations

import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
DST = ROOT / "baseline" / "_ref"

STUBS = {
    "bmipy.py": 'class Bmi:  # stand-in for bmipy.Bmi (abstract interface; the reference only derives from it)\n    pass\n',
    "timezonefinder.py": (
        "import os\n\n\nclass TimezoneFinder:  # stand-in: the polygon database is not installable offline\n"
        "    def timezone_at(self, lat=None, lng=None):\n        return os.environ.get('TFG_REF_TZ', 'America/Los_Angeles')\n\n"
        "    certain_timezone_at = timezone_at\n"),
    "pyprojroot.py": "from pathlib import Path\n\n\ndef here():\n    return Path(__file__).resolve().parent\n",
}


def install(force: bool = False) -> Path | None:
    if not (REF / "src" / "topoflow_glacier").is_dir():
        return DST if (DST / "topoflow_glacier").is_dir() else None
    if (DST / "topoflow_glacier").is_dir() and not force:
        return DST
    if DST.exists():
        shutil.rmtree(DST)
    DST.mkdir(parents=True)
    shutil.copytree(REF / "src" / "topoflow_glacier", DST / "topoflow_glacier",
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    (DST / "topoflow_glacier" / "_version.py").write_text('__version__ = "0+reference"\nversion = __version__\n')
    for name, text in STUBS.items():
        (DST / name).write_text(text)
    (DST / "data").mkdir()
    shutil.copy(REF / "tests" / "data" / "sample-cat-3062920.csv", DST / "data" / "sample-cat-3062920.csv")
    (DST / "README").write_text("Unmodified copy of NGWPC/topoflow-glacier src/ + stub dependencies, made by "
                                "scripts/install_reference.py; git-ignored.\n")
    return DST


if __name__ == "__main__":
    p = install(force="--force" in sys.argv)
    print(p if p else "reference checkout not present; nothing installed")
