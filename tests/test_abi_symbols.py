"""CPU: the C-ABI library loads and exports every symbol include/tfglacier.h declares (no compute calls)."""

import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "tfglacier.h").read_text()
    return sorted(set(re.findall(r"TFG_API[^;(]*?\b(tfg_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("tfg_create", "tfg_destroy", "tfg_set_constants", "tfg_bind_static", "tfg_bind_state", "tfg_bind_time",
              "tfg_run", "tfg_ingest_async", "tfg_convert_forcing", "tfg_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from topoflow_glacier_b200 import _lib, build

    build.build()  # no-op when current; nvcc cross-compiles without a GPU
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for s in declared_symbols():
        assert hasattr(lib, s), s
    assert set(declared_symbols()) == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert _lib.load().tfg_abi_version() == 1


def test_struct_layouts_match_header():
    from topoflow_glacier_b200 import _lib

    assert ctypes.sizeof(_lib.TimeRow) == 8 * 8
    assert ctypes.sizeof(_lib.Constants) == 26 * 8 + 2 * 4
    assert ctypes.sizeof(_lib.Statics) == 14 * 8
    assert ctypes.sizeof(_lib.State) == 19 * 8
    text = (ROOT / "include" / "tfglacier.h").read_text()
    enum = re.search(r"enum tfg_rec \{(.*?)\};", text, re.S).group(1)
    names = [n for n in re.findall(r"TFG_REC_(\w+)", enum) if n != "COUNT"]
    assert len(names) == len(_lib.REC_NAMES)


def test_no_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from topoflow_glacier_b200 import _lib

    ctx = ctypes.c_void_p()
    assert _lib.load().tfg_create(ctypes.byref(ctx), 0, 0) != 0
    assert _lib.load().tfg_last_error()
    from topoflow_glacier_b200.bmi import BmiTopoflowGlacier

    m = BmiTopoflowGlacier()
    cfg = ROOT / "tests" / "golden" / "cat-test.yaml"
    with pytest.raises(RuntimeError):
        m._build([__import__("topoflow_glacier_b200.config", fromlist=["x"]).TopoflowGlacierConfig.model_validate(
            {"site_prefix": "x", "forcing_file": "x", "dt": 1, "start_time": 2013032000, "end_time": "2013033100",
             "da": 1.0, "slope": 1.0, "lat": 46.0, "lon": -121.0, "h0_snow": 0, "h0_ice": 0, "h0_swe": 0,
             "h0_iwe": 0, "elev": 100.0})])


def test_build_flags_that_results_depend_on():
    """The fast float64 unit must not contract multiply-adds implicitly (bit-identical template instantiations), the
    float32 unit is built flush-to-zero, and everything targets sm_100a."""
    from topoflow_glacier_b200 import build

    assert "arch=compute_100a,code=sm_100a" in build.NVCC_FLAGS
    assert "-fmad=false" in build.EXTRA_FLAGS["tfg_run_fast.cu"]
    assert "-ftz=true" in build.EXTRA_FLAGS["tfg_run_f32.cu"]
    assert "tfg_run_strict.cu" not in build.EXTRA_FLAGS      # strict: single-rounding intrinsics, no special flags
