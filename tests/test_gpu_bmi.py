"""GPU: the BMI surface, ported from the reference's own tests (reference tests/integration_test.py:64-243),
plus the multi-step fast path, forcing ingestion and basin aggregates -- all through the C ABI."""

import numpy as np
import pytest
import yaml

from helpers import ATOL, GOLDEN, err_report, load_case, make_engine, make_oracle

pytestmark = pytest.mark.gpu

SAMPLE_CONFIG = {
    "site_prefix": "cat-3062920", "forcing_file": "data/sample-cat-3062920.csv", "dt": 1, "start_time": "2013032000",
    "end_time": "2013033100", "da": 11.418749923500716, "slope": 88.582729, "aspect": 242.8644693769529,
    "lon": -121.81418, "lat": 46.81953220, "elev": 2446.3922737596167, "h_active_layer": 0.125, "h0_snow": 5.0,
    "h0_ice": 2.0, "h0_swe": 0.25, "h0_iwe": 1.834, "T_rain_snow": 0.0,
}
SET_ORDER = ("atmosphere_water__liquid_equivalent_precipitation_rate", "land_surface_air__temperature",
             "land_surface_air__pressure", "atmosphere_air_water~vapor__relative_saturation", "wind_speed_UV")


def _model(tmp_path, cfg):
    from topoflow_glacier import BmiTopoflowGlacier  # the reference's import path

    p = tmp_path / "cfg.yaml"
    p.write_text(yaml.dump(cfg))
    m = BmiTopoflowGlacier()
    m.initialize(str(p))
    return m


def test_full_model_workflow(tmp_path, cuda_device):
    """integration_test.py:67-153: per-step set/update/get driver loop against the reference's golden vector."""
    case = load_case("sample265")
    model = _model(tmp_path, SAMPLE_CONFIG)
    dest = np.zeros(1)
    assert model.get_value("snowpack__depth", dest).item() == 5.0
    assert model.get_value("glacier_ice__thickness", dest).item() == 2.0
    f = case["forcing"]
    out = np.zeros(f.shape[0])
    for i in range(f.shape[0]):
        for name, v in zip(SET_ORDER, f[i, :, 0]):
            model.set_value(name, np.array([v]))
        model.set_value("land_surface_radiation~incoming~longwave__energy_flux", np.array([300.0]))
        model.set_value("land_surface_radiation~incoming~shortwave__energy_flux", np.array([100.0]))
        model.update()
        for name in ("snowpack__melt_volume_flux", "glacier_ice__melt_volume_flux", "snowpack__depth",
                     "glacier_ice__thickness"):
            assert model.get_value(name, np.zeros(1)).item() >= 0
        out[i] = model.get_value("land_surface_water__runoff_volume_flux", np.zeros(1))[0]
    model.finalize()
    out = out * model.da_m2
    np.testing.assert_allclose(out, case["upstream_output_m_total"], rtol=1e-12, atol=3e-18 * model.da_m2)
    assert abs(out.sum() - 287.4577) < 1e-3


def test_bmi_variable_access(tmp_path, cuda_device):
    """integration_test.py:155-186."""
    model = _model(tmp_path, SAMPLE_CONFIG)
    assert "land_surface_air__temperature" in model.get_input_var_names()
    assert "atmosphere_water__liquid_equivalent_precipitation_rate" in model.get_input_var_names()
    assert "snowpack__depth" in model.get_output_var_names()
    assert "glacier_ice__thickness" in model.get_output_var_names()
    assert "float" in model.get_var_type("snowpack__depth")
    assert model.get_var_itemsize("snowpack__depth") == 8
    assert model.get_var_nbytes("snowpack__depth") == 8
    model.set_value("land_surface_air__temperature", np.array([273.15]))
    got = np.zeros(1)
    model.get_value("land_surface_air__temperature", got)
    assert np.allclose(got, [273.15])
    # superset: what raises in the reference is defined here
    assert model.get_component_name() == "Topoflow-Glacier"
    assert model.get_time_units() == "s" and model.get_time_step() == 3600.0 and model.get_current_time() == 0
    assert model.get_var_units("snowpack__melt_volume_flux") == "m s-1"
    assert model.get_input_item_count() == 7 and model.get_output_item_count() == 8
    assert model.get_grid_size(model.get_var_grid("snowpack__depth")) == 1
    with pytest.raises(KeyError):
        model.get_value_ptr("no_such_variable")
    ptr = model.get_value_ptr("snowpack__depth")
    assert ptr.is_cuda and ptr.data_ptr() == model.get_value_ptr("snowpack__depth").data_ptr()
    model.set_value("wind_speed_UV", 3.5)  # scalars broadcast like ndarray[:] = scalar
    assert model.get_value("wind_speed_UV", np.zeros(1))[0] == 3.5
    model.finalize()


def test_no_snow_no_ice(tmp_path, cuda_device):
    """integration_test.py:192-243."""
    cfg = dict(SAMPLE_CONFIG, h0_snow=0.0, h0_ice=0.0, h0_swe=0.0, h0_iwe=0.0)
    model = _model(tmp_path, cfg)
    for name, v in zip(SET_ORDER, (0.0, 5.0, 88000.0, 0.003, 2.0)):
        model.set_value(name, np.array([v]))
    model.update()
    assert model.get_value("snowpack__melt_volume_flux", np.zeros(1)).item() == 0.0
    assert model.get_value("glacier_ice__melt_volume_flux", np.zeros(1)).item() == 0.0
    model.finalize()


def test_int_start_time_yaml_and_ensemble(tmp_path, cuda_device):
    """Configs with unquoted integer times load (3 of the 5 shipped yamls); an ensemble advances N cells at once."""
    case = load_case("cats288")
    cfgs = []
    for i in range(case["N"]):
        c = dict(SAMPLE_CONFIG, start_time=2013032000, end_time=2013033123)
        c.update({k: float(v[i]) for k, v in case["statics"].items()})
        cfgs.append(c)
    from topoflow_glacier_b200 import BmiTopoflowGlacier

    m = BmiTopoflowGlacier()
    m.initialize_ensemble(cfgs)
    import torch

    T = 48
    forcing = torch.as_tensor(case["forcing"][:T]).cuda()
    m.load_forcing(forcing)
    m.update_until(24 * 3600.0)
    assert m.get_current_time() == 24 * 3600.0
    m.update_until(T * 3600.0)
    ora = make_oracle(case, strict_pow=True)
    want = ora.run(case["forcing"][:T], record=("M_total", "h_swe", "h_iwe", "RH"))
    for bmi_name, k in (("land_surface_water__runoff_volume_flux", "M_total"),
                        ("snowpack__liquid-equivalent_depth", "h_swe"), ("glacier__liquid_equivalent_depth", "h_iwe"),
                        ("atmosphere_bottom_air_water-vapor__relative_saturation", "RH")):
        got = m.get_value(bmi_name, np.zeros(case["N"]))
        ok, *rest = err_report(got, want[k][-1], ATOL[k])
        assert ok, (k, rest)
    assert m.get_grid_size(0) == case["N"] and m.get_var_nbytes("snowpack__depth") == 8 * case["N"]
    m.finalize()


def test_update_until_holds_inputs(tmp_path, cuda_device):
    """update_until without a forcing block == repeated update() with unchanged inputs (reference :489-490)."""
    a, b = _model(tmp_path, SAMPLE_CONFIG), _model(tmp_path, SAMPLE_CONFIG)
    for m in (a, b):
        for name, v in zip(SET_ORDER, (0.0004, -2.0, 88000.0, 0.003, 4.0)):
            m.set_value(name, v)
    for _ in range(30):
        a.update()
    b.update_until(30 * 3600.0)
    for name in a.get_output_var_names():
        x, y = a.get_value(name, np.zeros(1)), b.get_value(name, np.zeros(1))
        assert x[0] == y[0], name


def test_forcing_streamer_equals_host_conversion(cuda_device):
    """pinned -> async H2D -> device unit conversion == the driver's NumPy conversion, bit for bit; float32
    sources widen exactly; streaming in chunks == one resident block."""
    import torch

    from topoflow_glacier_b200.forcing import ForcingStreamer, convert_on_host

    case = load_case("cats288")
    N, T = case["N"], 100
    rng = np.random.default_rng(3)
    raw = np.stack([rng.exponential(0.4, (T, N)), 273.15 + rng.normal(0, 6, (T, N)), 88900 + rng.normal(0, 300, (T, N)),
                    rng.uniform(1e-3, 6e-3, (T, N)), rng.normal(0, 3, (T, N)), rng.normal(0, 3, (T, N))], axis=1)
    raw32 = raw.astype(np.float32)
    want = convert_on_host(raw32.astype(np.float64))
    for src, dtype in ((raw32.astype(np.float64), "float64"), (raw32, "float32"),
                       (torch.as_tensor(raw32).pin_memory(), "float32")):
        eng = make_engine(case, mode="f64")
        st = ForcingStreamer(eng, chunk_steps=32, raw_dtype=dtype)
        got = torch.cat([c.clone() for c in st.chunks(src)]).cpu().numpy()
        assert np.array_equal(got, want)
        eng.close()
    ref = make_engine(case, mode="f64")
    ref.run(torch.as_tensor(want).cuda())
    eng = make_engine(case, mode="f64")
    ForcingStreamer(eng, chunk_steps=17, raw_dtype="float32").drive(raw32)
    torch.cuda.synchronize()
    assert torch.equal(eng.state, ref.state) and torch.equal(eng.ring, ref.ring)


def test_basin_aggregates_match_host_sums(cuda_device):
    """Area-weighted per-basin sums from the kernel (warp shuffles + RED) == NumPy bincount of the recorded series."""
    import torch

    import bench
    from helpers import default_constants
    from topoflow_glacier_b200.engine import MeltEngine
    from topoflow_glacier_b200.sharding import basin_sums_host

    N, T, NB = 5000, 12, 37
    statics, forcing = bench.synthetic_host_sample(N, T, seed=5)
    rng = np.random.default_rng(1)
    for basin_id in (np.sort(rng.integers(0, NB, N)).astype(np.int32), rng.integers(0, NB, N).astype(np.int32)):
        eng = MeltEngine(statics, default_constants(), "2013040100", zones=[-8.0], mode="f64_fast", basin_id=basin_id,
                         n_basin=NB, horizon_steps=T + 1)
        agg = torch.zeros(T, NB, 3, dtype=torch.float64, device=cuda_device)
        rec = eng.run(torch.as_tensor(forcing).cuda(), record=("M_total", "h_swe", "h_iwe"), basin_agg=agg)
        agg = agg.cpu().numpy()
        da_m2 = statics["da"] * 1e6
        for t in range(T):
            for j, k in enumerate(("M_total", "h_swe", "h_iwe")):
                want = basin_sums_host(rec[k][t].cpu().numpy(), da_m2, basin_id, NB)
                np.testing.assert_allclose(agg[t, :, j], want, rtol=1e-12, atol=1e-18)
        eng.close()


def test_exact_basin_aggregates_are_order_and_sharding_independent(cuda_device):
    """TFG_OPT_EXACT_AGG: fixed-point integer accumulators.  The sums equal the host sums, repeat bit for bit, and
    one engine over all cells == two engines over 128-aligned shards whose integer words are added (what the
    int64 all-reduce does on N GPUs), bit for bit; the float path only agrees to rounding."""
    import torch

    import bench
    from helpers import default_constants
    from topoflow_glacier_b200.engine import MeltEngine
    from topoflow_glacier_b200.sharding import BasinAggregates, basin_sums_host, shard_bounds

    N, T, NB = 6000, 10, 23
    statics, forcing = bench.synthetic_host_sample(N, T, seed=11)
    rng = np.random.default_rng(2)
    basin_id = rng.integers(0, NB, N).astype(np.int32)      # warps straddle basins: per-lane atomics
    basin_id[1024:4096] = np.sort(basin_id[1024:4096])      # ... and long uniform runs: warp butterflies
    f = torch.as_tensor(forcing).cuda()

    def run(lo, hi, exps=None):
        st = {k: v[lo:hi] for k, v in statics.items()}
        eng = MeltEngine(st, default_constants(), "2013040100", zones=[-8.0], mode="f64_fast", basin_id=basin_id[lo:hi],
                         n_basin=NB, horizon_steps=T + 1)
        agg = BasinAggregates(T, NB, device=cuda_device, exponents=exps or eng.agg_exponents())
        rec = eng.run(f[:, :, lo:hi].contiguous(), record=("M_total", "h_swe", "h_iwe"), basin_agg=agg.zero())
        torch.cuda.synchronize()
        out = (agg, {k: v.cpu().numpy() for k, v in rec.items()}, eng.agg_exponents())
        eng.close()
        return out

    whole, rec, exps = run(0, N)
    assert whole.n_left_out == 0
    again, _, _ = run(0, N)
    assert torch.equal(whole.accumulator, again.accumulator)           # independent of the order of the atomics
    acc = torch.zeros_like(whole.accumulator)
    for r in range(2):
        lo, hi = shard_bounds(N, 2, r)
        part, _, _ = run(lo, hi, exps)                                  # same exponents on every shard
        acc += part.accumulator
    assert torch.equal(acc, whole.accumulator)                         # sharding-independent, bit for bit
    got = whole.reduce() or whole.buffer.cpu().numpy()
    da_m2 = statics["da"] * 1e6
    for t in range(T):
        for j, k in enumerate(("M_total", "h_swe", "h_iwe")):
            want = basin_sums_host(rec[k][t], da_m2, basin_id, NB)
            np.testing.assert_allclose(got[t, :, j], want, rtol=1e-13, atol=1e-20)


@pytest.mark.parametrize("exact", [False, True])
def test_long_run_is_split_into_launches_transparently(exact, cuda_device):
    """A run longer than one launch's 128 clock rows (tfg_run splits it): recorded series, basin aggregates (float
    and fixed-point) and final state equal those of caller-side pieces of 100 + 100 + 88 steps, bit for bit."""
    import torch

    from topoflow_glacier_b200.sharding import BasinAggregates

    case = load_case("cats288")
    T = case["forcing"].shape[0]
    assert T > 2 * 128
    f = torch.as_tensor(case["forcing"]).cuda()
    basin = np.array([0, 0, 1, 1], dtype=np.int32)

    def run(pieces):
        eng = make_engine(case, mode="f64_fast", basin_id=basin, n_basin=2)
        recs, aggs, done = [], [], 0
        for n in pieces:
            agg = BasinAggregates(n, 2, device=cuda_device, exponents=eng.agg_exponents() if exact else None)
            recs.append(eng.run(f[done:done + n].contiguous(), n, record=("M_total", "h_swe"), basin_agg=agg.zero()))
            agg.reduce()
            aggs.append(agg.buffer.clone())
            done += n
        torch.cuda.synchronize()
        out = (torch.cat([r["M_total"] for r in recs]), torch.cat([r["h_swe"] for r in recs]), torch.cat(aggs),
               eng.state.clone(), eng.ring.clone())
        eng.close()
        return out

    whole, parts = run([T]), run([100, 100, T - 200])
    for a, b in zip(whole, parts):
        assert torch.equal(a, b)
    assert float(whole[2][:, :, 1].abs().sum()) > 0


@pytest.mark.parametrize("mode", ["f64", "f64_fast", "f32"])
def test_forcing_map_equals_replicated_forcing(mode, cuda_device):
    """tfg_bind_forcing_map: cells that share a catchment's forcing column == the same forcing replicated per cell,
    bit for bit, through the fused run, the per-step path and the host streamer."""
    import torch

    from helpers import default_constants
    from topoflow_glacier_b200.engine import MeltEngine
    from topoflow_glacier_b200.forcing import ForcingStreamer
    import bench

    N, M, T = 700, 9, 40
    statics, _ = bench.synthetic_host_sample(N, 1, seed=3)
    _, fcols = bench.synthetic_host_sample(M, T, seed=4)           # [T, 5, M]: one series per catchment
    rng = np.random.default_rng(0)
    col = rng.integers(0, M, N).astype(np.int32)
    col[100:400] = 3                                               # a long run of cells in one catchment
    kw = dict(zones=[-8.0], mode=mode, horizon_steps=T + 1)
    a = MeltEngine(statics, default_constants(), "2013040100", **kw)
    b = MeltEngine(statics, default_constants(), "2013040100", forcing_index=col, n_forcing_cols=M, **kw)
    c = MeltEngine(statics, default_constants(), "2013040100", forcing_index=col, n_forcing_cols=M, **kw)
    f_cols = torch.as_tensor(fcols).to(cuda_device, a.dtype).contiguous()
    f_full = f_cols[:, :, torch.as_tensor(col, dtype=torch.int64, device=cuda_device)].contiguous()
    a.run(f_full)
    b.run(f_cols)
    # host streamer on [T, 6, M] raw columns, then per-step update() from the [7, M] input block
    raw = np.stack([fcols[:, 0] * 1e3, fcols[:, 1] + 273.15, fcols[:, 2], fcols[:, 3], fcols[:, 4], np.zeros_like(fcols[:, 4])], axis=1)
    d = MeltEngine(statics, default_constants(), "2013040100", forcing_index=col, n_forcing_cols=M, **kw)
    ForcingStreamer(d, chunk_steps=7, raw_dtype="float64").drive(np.ascontiguousarray(raw))
    ref = MeltEngine(statics, default_constants(), "2013040100", forcing_index=col, n_forcing_cols=M, **kw)
    from topoflow_glacier_b200.forcing import convert_on_host
    ref.run(torch.as_tensor(convert_on_host(raw)).to(cuda_device, a.dtype).contiguous())
    for t in range(T):
        c.inputs[:5].copy_(f_cols[t])
        c.step()
    torch.cuda.synchronize()
    assert torch.equal(a.state, b.state) and torch.equal(a.ring, b.ring)
    assert torch.equal(a.state, c.state) and torch.equal(a.ring, c.ring)
    assert torch.equal(d.state, ref.state)
    with pytest.raises(ValueError):
        b.run(f_full)                                              # a per-cell block no longer fits the map
    for e in (a, b, c, d, ref):
        e.close()


def test_bmi_ensemble_with_shared_forcing_series(cuda_device):
    """BMI surface on an ensemble whose members share forcing series: inputs hold one value per series."""
    from topoflow_glacier import BmiTopoflowGlacier

    z = np.load(GOLDEN / "cats288.npz")
    keys = ("da", "slope", "aspect", "lon", "lat", "elev", "h0_snow", "h0_ice", "h0_swe", "h0_iwe", "T_rain_snow")
    cfgs = [dict({k: float(z[f"static_{k}"][i % 4]) for k in keys}, site_prefix=f"c{i}", forcing_file="-", dt=1,
                 start_time="2013032000", end_time="2013033123") for i in range(6)]
    col = np.array([0, 0, 1, 1, 1, 0], dtype=np.int32)
    shared = BmiTopoflowGlacier(); shared.initialize_ensemble(cfgs, forcing_index=col, n_forcing_cols=2)
    plain = BmiTopoflowGlacier(); plain.initialize_ensemble(cfgs)
    f = z["forcing"][:12, :, :2]                                  # two series
    names = ("atmosphere_water__liquid_equivalent_precipitation_rate", "land_surface_air__temperature",
             "land_surface_air__pressure", "atmosphere_air_water~vapor__relative_saturation", "wind_speed_UV")
    dest_s, dest_p = np.zeros(6), np.zeros(6)
    for t in range(f.shape[0]):
        for j, name in enumerate(names):
            shared.set_value(name, f[t, j])                       # 2 values
            plain.set_value(name, f[t, j][col])                   # 6 values
        shared.update(); plain.update()
        for name in shared.get_output_var_names():
            assert np.array_equal(shared.get_value(name, dest_s), plain.get_value(name, dest_p)), (t, name)
    shared.finalize(); plain.finalize()


def test_checkpoint_resume_is_bit_identical(tmp_path, cuda_device):
    """state + snowfall window + step counter saved mid-run, resumed in a fresh model == uninterrupted run."""
    import torch

    case = load_case("rand64")
    T = case["forcing"].shape[0]
    forcing = torch.as_tensor(case["forcing"]).cuda()
    a = make_engine(case, mode="f64_fast")
    a.run(forcing)
    b = make_engine(case, mode="f64_fast")
    b.run(forcing[:20].contiguous(), 20)
    torch.save(b.state_dict(), tmp_path / "ckpt.pt")
    c = make_engine(case, mode="f64_fast")
    c.load_state_dict(torch.load(tmp_path / "ckpt.pt", weights_only=False))
    assert c.step_index == 20
    c.run(forcing[20:].contiguous(), T - 20)
    torch.cuda.synchronize()
    assert torch.equal(a.state, c.state) and torch.equal(a.ring, c.ring)
    d = make_engine(load_case("cats288"), mode="f64_fast")
    with pytest.raises(ValueError):
        d.load_state_dict(b.state_dict())
    # through the BMI surface
    m = _model(tmp_path, SAMPLE_CONFIG)
    for name, v in zip(SET_ORDER, (0.0004, -2.0, 88000.0, 0.003, 4.0)):
        m.set_value(name, v)
    for _ in range(5):
        m.update()
    m.save_state(tmp_path / "bmi.pt")
    n = _model(tmp_path, SAMPLE_CONFIG)
    n.load_state(tmp_path / "bmi.pt")
    assert n.get_current_time() == m.get_current_time() == 5 * 3600.0
    m.update(), n.update()
    for name in m.get_output_var_names():
        assert m.get_value(name, np.zeros(1))[0] == n.get_value(name, np.zeros(1))[0]


def test_mock_routing_fir_matches_numpy_convolve(cuda_device):
    """20-tap 0.05 box filter of the reference example (examples/run_topoflow_glacier.py:129-131) on the device."""
    import torch

    case = load_case("cats288")
    eng = make_engine(case, mode="f64_fast")
    q = eng.run(torch.as_tensor(case["forcing"]).cuda(), record=("M_total",))["M_total"].to(torch.float64)
    q = q * torch.as_tensor(case["statics"]["da"] * 1e6).cuda()
    routed = eng.route_fir(q).cpu().numpy()
    qh = q.cpu().numpy()
    w = np.zeros(20) + 0.05
    for j in range(qh.shape[1]):
        want = np.convolve(qh[:, j], w, mode="full")[: qh.shape[0]]
        np.testing.assert_allclose(routed[:, j], want, rtol=1e-13, atol=1e-18)
    eng.close()


def _shared_ensemble(n_members=6):
    from topoflow_glacier import BmiTopoflowGlacier

    z = np.load(GOLDEN / "cats288.npz")
    keys = ("da", "slope", "aspect", "lon", "lat", "elev", "h0_snow", "h0_ice", "h0_swe", "h0_iwe", "T_rain_snow")
    cfgs = [dict({k: float(z[f"static_{k}"][i % 4]) for k in keys}, site_prefix=f"c{i}", forcing_file="-", dt=1,
                 start_time="2013032000", end_time="2013033123") for i in range(n_members)]
    col = np.array([0, 0, 1, 1, 1, 0], dtype=np.int32)[:n_members]
    m = BmiTopoflowGlacier()
    m.initialize_ensemble(cfgs, forcing_index=col, n_forcing_cols=2)
    return m, cfgs, col, z["forcing"]


def test_bulk_bmi_path_with_forcing_map(cuda_device):
    """update_steps / load_forcing + update_until on an ensemble built with forcing_index:
    the block has one column per SERIES; equal to per-step update() bit for bit; held inputs work too; a block of
    the wrong shape is rejected BEFORE the model advances; an exhausted queue raises instead of silently holding
    the inputs constant."""
    import torch

    bulk, cfgs, col, forcing = _shared_ensemble()
    step, _, _, _ = _shared_ensemble()
    T = 30
    f = np.ascontiguousarray(forcing[:T, :, :2])
    f_dev = torch.as_tensor(f).cuda()
    bulk.update_steps(10, f_dev[:10].contiguous())
    bulk.load_forcing(f_dev[10:].contiguous())
    bulk.update_until(25 * 3600.0)
    names = ("atmosphere_water__liquid_equivalent_precipitation_rate", "land_surface_air__temperature",
             "land_surface_air__pressure", "atmosphere_air_water~vapor__relative_saturation", "wind_speed_UV")
    for t in range(25):
        for j, name in enumerate(names):
            step.set_value(name, f[t, j])
        step.update()
    a, b = np.zeros(6), np.zeros(6)
    for name in bulk.get_output_var_names():
        assert np.array_equal(bulk.get_value(name, a), step.get_value(name, b)), name
    # the inputs reflect the last consumed step (one value per series)
    assert np.array_equal(bulk.get_value("land_surface_air__temperature", np.zeros(2)), f[24, 1])
    # held inputs over an interval == repeated update()
    bulk._forcing_block = None
    bulk.update_until(28 * 3600.0)
    for _ in range(3):
        step.update()
    for name in bulk.get_output_var_names():
        assert np.array_equal(bulk.get_value(name, a), step.get_value(name, b)), name
    # wrong shape: rejected, model not advanced
    now = bulk.get_current_time()
    with pytest.raises(ValueError):
        bulk.update_steps(2, torch.zeros(2, 5, 6, dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        bulk.load_forcing(torch.zeros(2, 5, 6, dtype=torch.float64, device="cuda"))
    assert bulk.get_current_time() == now
    # exhausted queue: the remaining steps are consumed, then a loud error
    bulk.load_forcing(f_dev[28:30].contiguous())
    with pytest.raises(RuntimeError, match="held 2 more steps"):
        bulk.update_until(now + 5 * 3600.0)
    assert bulk.get_current_time() == now + 2 * 3600.0
    bulk.finalize(), step.finalize()


def test_back_to_back_set_update_without_reads(cuda_device):
    """A driver that never reads outputs between steps (set_value x5, update, repeat): the pinned staging block is
    re-written while the previous upload may still be in flight (bmi.py:157).  Large N (initialize_cells)
    so that copies and kernels take long enough to run behind the host loop; compared with the fused run."""
    import torch

    import bench
    from helpers import default_constants
    from topoflow_glacier import BmiTopoflowGlacier
    from topoflow_glacier_b200.engine import MeltEngine

    N, T = 400_000, 40
    statics, forcing = bench.synthetic_host_sample(N, T, seed=21)
    m = BmiTopoflowGlacier()
    m.initialize_cells(dict(SAMPLE_CONFIG, start_time="2013040100", utc_offset_hours=-8.0, precision="f64_fast"), statics)
    assert m.get_grid_size(0) == N and m.get_var_nbytes("snowpack__depth") == 8 * N
    ref = MeltEngine(statics, default_constants(), "2013040100", zones=[-8.0], mode="f64_fast", horizon_steps=T + 2)
    ref.run(torch.as_tensor(forcing).cuda())
    for t in range(T):
        for j, name in enumerate(SET_ORDER):
            m.set_value(name, forcing[t, j])
        m.update()
    got = {name: m.get_value(name, np.zeros(N)).copy() for name in m.get_output_var_names()}
    # update() re-sums the snowfall window exactly every step, the fused run incrementally: decisions are identical
    for name, k in (("snowpack__liquid-equivalent_depth", "h_swe"), ("glacier__liquid_equivalent_depth", "h_iwe"),
                    ("land_surface_water__runoff_volume_flux", "M_total"), ("snowpack__depth", "h_snow"),
                    ("atmosphere_bottom_air_water-vapor__relative_saturation", "RH")):
        assert np.array_equal(got[name], ref.row(k).cpu().numpy()), name
    m.finalize(), ref.close()


def test_bmi_index_access_grid_and_time_queries(tmp_path, cuda_device):
    """get/set_value_at_indices, get_grid_shape, get_end_time, get_var_location, get_var_grid (VERDICT r1 weak #10;
    reference bmi_topoflow_glacier.py:1804-1808, context.py:47-59; they raise in the reference's base class)."""
    m, cfgs, col, forcing = _shared_ensemble()
    n = len(cfgs)
    assert m.get_end_time() == (11 * 24 + 23) * 3600.0 and m.get_start_time() == 0
    assert m.get_grid_rank(0) == 1 and m.get_grid_size(0) == n and m.get_grid_node_count(0) == n
    assert m.get_grid_type(0) == "points"
    shape = np.zeros(1, dtype=np.int64)
    assert m.get_grid_shape(0, shape) is shape and shape[0] == n
    assert m.get_var_location("snowpack__depth") == "node" and m.get_var_grid("wind_speed_UV") == 0
    with pytest.raises(KeyError):
        m.get_var_location("nope")
    # outputs: set single members, read them back through both accessors
    name = "snowpack__liquid-equivalent_depth"
    before = m.get_value(name, np.zeros(n)).copy()
    m.set_value_at_indices(name, np.array([1, 4]), np.array([0.125, 0.5]))
    after = m.get_value(name, np.zeros(n))
    want = before.copy()
    want[[1, 4]] = (0.125, 0.5)
    assert np.array_equal(after, want)
    got = m.get_value_at_indices(name, np.zeros(3), np.array([4, 0, 1]))
    assert np.array_equal(got, want[[4, 0, 1]])
    assert np.array_equal(m.get_value_ptr(name).cpu().numpy(), want)      # the live device tensor saw the write
    # inputs hold one value per forcing series (2)
    t_name = "land_surface_air__temperature"
    m.set_value(t_name, np.array([1.5, -3.0]))
    m.set_value_at_indices(t_name, np.array([1]), np.array([-7.25]))
    assert np.array_equal(m.get_value(t_name, np.zeros(2)), [1.5, -7.25])
    assert np.array_equal(m.get_value_at_indices(t_name, np.zeros(1), np.array([1])), [-7.25])
    m.update()                                                             # and the kernel consumed it
    assert m.get_current_time() == 3600.0
    single = _model(tmp_path, SAMPLE_CONFIG)
    assert single.get_grid_type(0) == "scalar" and single.get_end_time() == 11 * 24 * 3600.0
    single.finalize(), m.finalize()


def test_sharded_engine_single_rank_equals_plain_engine(cuda_device):
    """ShardedMeltEngine without a process group == MeltEngine + BasinAggregates; cells from a factory too."""
    import torch

    import bench
    from helpers import default_constants
    from topoflow_glacier_b200.engine import MeltEngine
    from topoflow_glacier_b200.sharding import BasinAggregates, ShardedMeltEngine

    N, T, NB = 3000, 9, 11
    statics, forcing = bench.synthetic_host_sample(N, T, seed=8)
    basin = (np.arange(N) * NB // N).astype(np.int32)
    f = torch.as_tensor(forcing).cuda()
    kw = dict(zones=[-8.0], mode="f64_fast", horizon_steps=T + 1)
    plain = MeltEngine(statics, default_constants(), "2013040100", basin_id=basin, n_basin=NB, **kw)
    agg = BasinAggregates(T, NB, device=cuda_device, exponents=plain.agg_exponents())
    plain.run(f, basin_agg=agg.zero())
    agg.reduce()
    sh = ShardedMeltEngine(statics, default_constants(), "2013040100", basin_id=basin, n_basin=NB, **kw)
    assert sh.bounds == (0, N) and sh.world == 1
    _, got = sh.run(sh.local(f).contiguous())
    fac = ShardedMeltEngine(lambda lo, hi: {k: v[lo:hi] for k, v in statics.items()}, default_constants(), "2013040100",
                            n_total=N, basin_id=lambda lo, hi: basin[lo:hi], n_basin=NB, **kw)
    _, got2 = fac.run(f)
    torch.cuda.synchronize()
    assert torch.equal(got, agg.buffer) and torch.equal(got2, agg.buffer)
    assert torch.equal(sh.state, plain.state)
    np.testing.assert_allclose(sh.basin_area().cpu().numpy(), np.bincount(basin, weights=statics["da"] * 1e6), rtol=1e-14)
    for e in (plain, sh, fac):
        e.close()


@pytest.mark.timeout(600)
def test_sharded_engine_two_ranks_equal_one(tmp_path, cuda_device):
    """Two ranks (one process each; gloo carries the collectives so that both may share the one GPU of the test box)
    driving ShardedMeltEngine / BmiTopoflowGlacier.initialize_cells(shard=True): the global exact aggregates are
    bit-identical to a single engine over all cells, and the shards' states are the single engine's rows."""
    import os
    import subprocess
    import sys

    import torch

    import bench
    from helpers import default_constants
    from topoflow_glacier_b200.engine import MeltEngine
    from topoflow_glacier_b200.sharding import BasinAggregates

    N, T, NB = 5000, 14, 13
    statics, forcing = bench.synthetic_host_sample(N, T, seed=12)
    rng = np.random.default_rng(4)
    basin = rng.integers(0, NB, N).astype(np.int32)
    basin[500:3000] = np.sort(basin[500:3000])
    np.savez(tmp_path / "in.npz", forcing=forcing, basin=basin, **{f"s_{k}": v for k, v in statics.items()})
    one = MeltEngine(statics, default_constants(), "2013040100", zones=[-8.0], mode="f64_fast", basin_id=basin,
                     n_basin=NB, horizon_steps=T + 1)
    agg = BasinAggregates(T, NB, device=cuda_device, exponents=one.agg_exponents())
    one.run(torch.as_tensor(forcing).cuda(), basin_agg=agg.zero())
    agg.reduce()
    torch.cuda.synchronize()
    root = str(__import__("pathlib").Path(__file__).resolve().parent.parent)
    script = tmp_path / "w.py"
    script.write_text(f'''
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
os.environ["NGEN_EWTS_LOGGING"] = "DISABLED"
import numpy as np, torch, torch.distributed as dist
from topoflow_glacier_b200.config import default_constants
from topoflow_glacier_b200.sharding import ShardedMeltEngine
from topoflow_glacier import BmiTopoflowGlacier
dist.init_process_group("gloo")
rank = dist.get_rank()
z = np.load({str(tmp_path / "in.npz")!r})
statics = {{k[2:]: z[k] for k in z.files if k.startswith("s_")}}
f = torch.as_tensor(z["forcing"]).cuda()
T = f.shape[0]
sh = ShardedMeltEngine(statics, default_constants(), "2013040100", zones=[-8.0], mode="f64_fast", basin_id=z["basin"],
                       n_basin={NB}, horizon_steps=T + 1, device=0)
done, parts = 0, []
for k in (5, T - 5):            # two launches: the aggregate buffers are per launch length
    _, g = sh.run(sh.local(f[done:done + k]).contiguous())
    parts.append(g.clone()); done += k
# the same through the BMI surface
cfg = dict(site_prefix="x", forcing_file="-", dt=1, start_time="2013040100", end_time="2013041000", da=1.0, slope=1.0,
           lat=46.0, lon=-121.0, h0_snow=0.0, h0_ice=0.0, h0_swe=0.0, h0_iwe=0.0, elev=1.0, utc_offset_hours=-8.0,
           precision="f64_fast")
m = BmiTopoflowGlacier()
m.initialize_cells(cfg, statics, shard=True, basin_id=z["basin"], n_basin={NB}, device=0)
lo, hi = m.sharded.bounds
assert m.get_grid_size(0) == hi - lo and len(m.da_m2) == hi - lo
m.update_steps(T, m.sharded.local(f).contiguous())
torch.cuda.synchronize()
assert torch.equal(m._engine.state, sh.state)
np.savez({str(tmp_path)!r} + f"/out{{rank}}.npz", agg=torch.cat(parts).cpu().numpy(), state=sh.state.cpu().numpy(),
         lo=lo, hi=hi)
dist.destroy_process_group()
sys.stdout.write("rank %d ok\\n" % rank)  # one write: two ranks share the pipe
''')
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       capture_output=True, text=True, env=env, timeout=560)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    outs = [np.load(tmp_path / f"out{k}.npz") for k in range(2)]
    want = agg.buffer.cpu().numpy()
    for o in outs:
        assert np.array_equal(o["agg"], want)                      # global sums on every rank, bit for bit
        assert np.array_equal(o["state"], one.state[:, int(o["lo"]):int(o["hi"])].cpu().numpy())
    assert int(outs[0]["hi"]) == int(outs[1]["lo"]) and int(outs[1]["hi"]) == N
    one.close()


@pytest.mark.parametrize("mode", ["f64", "f64_fast"])
def test_csv_file_to_streamer_to_kernel_matches_reference_golden(mode, tmp_path, cuda_device):
    """f4 end to end (VERDICT r1 missing #5): the forcing FILE the reference's test reads -> header-keyed CSV reader ->
    pinned block -> cudaMemcpyAsync on the side stream -> device unit conversion -> fused launches, against the
    reference's own golden vector tests/data/output_m_total.npy (reference tests/integration_test.py:81-153)."""
    import torch

    from topoflow_glacier import BmiTopoflowGlacier
    from topoflow_glacier_b200.forcing import ForcingStreamer, read_forcing_csv
    from topoflow_glacier_b200.timebase import parse_start

    root = GOLDEN.parent.parent
    cfg_path = tmp_path / "sample.yaml"
    cfg_path.write_text(yaml.dump(dict(SAMPLE_CONFIG, forcing_file="tests/data/sample-cat-3062920.csv", precision=mode)))
    model = BmiTopoflowGlacier()
    model.initialize(str(cfg_path))
    raw = read_forcing_csv(root / model.cfg.forcing_file, parse_start(model.cfg.start_time), parse_start(model.cfg.end_time))
    assert raw.shape == (265, 6)
    gold = np.load(root / "tests" / "data" / "output_m_total.npy")
    eng = model._engine
    for chunk_steps, dtype in ((64, "float64"), (265, "float64")):
        if eng.step_index:
            model.initialize(str(cfg_path))
            eng = model._engine
        st = ForcingStreamer(eng, chunk_steps=chunk_steps, raw_dtype=dtype)
        parts = [eng.run(c, c.shape[0], record=("M_total",))["M_total"] for c in st.chunks(raw[:, :, None])]
        out = torch.cat(parts).cpu().numpy()[:, 0] * model.da_m2
        np.testing.assert_allclose(out, gold, rtol=1e-12, atol=3e-18 * model.da_m2)
        assert st.h2d_bytes == raw.size * 8
    assert model.get_current_time() == 265 * 3600.0
    model.finalize()


def test_example_driver_on_shipped_configs(cuda_device):
    """BASELINE configs[0] and [1] as the reference drives them: examples/run_topoflow_glacier.py opens config/*.yaml and
    the forcing CSV it names (VERDICT r1 missing #3); hydrographs against the golden runs of the unmodified reference."""
    import importlib.util

    root = GOLDEN.parent.parent
    spec = importlib.util.spec_from_file_location("example_driver", root / "examples" / "run_topoflow_glacier.py")
    drv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(drv)
    # cfg 1: cat-3062784 (integer start_time / end_time in the yaml), 288 rows, strict mode, + the ensemble of all four
    out = drv.run(root / "config" / "cat-3062784.yaml", None, const=False, mode="f64", ensemble=True)
    case = load_case("cats288")
    want = case["ref"]["M_total"] * (case["statics"]["da"] * 1e6)[None, :]
    np.testing.assert_allclose(out["runoff_m3s"], want[:, 0], rtol=1e-12, atol=1e-10)
    assert np.array_equal(out["fused_runoff_m3s"], out["runoff_m3s"])           # streamed + fused == per-step loop
    np.testing.assert_allclose(out["ensemble_runoff_m3s"], want, rtol=1e-12, atol=1e-10)
    routed = np.convolve(out["runoff_m3s"], np.zeros(20) + 0.05, mode="full")[:288]
    np.testing.assert_allclose(out["fused_routed_m3s"], routed, rtol=1e-12, atol=1e-12)
    # cfg 2: the constant-forcing example (RAINRATE = 3, T2D = 10 degC), fast mode
    out = drv.run(root / "config" / "cat-3062920-const.yaml", None, const=True, mode="f64_fast", ensemble=False)
    case = load_case("const")
    want = case["ref"]["M_total"][:, 0] * case["statics"]["da"][0] * 1e6
    np.testing.assert_allclose(out["runoff_m3s"], want, rtol=1e-12, atol=1e-10)
    np.testing.assert_allclose(out["fused_runoff_m3s"], want, rtol=1e-12, atol=1e-10)


@pytest.mark.parametrize("mode", ["f64", "f32"])
def test_packed_int16_forcing_equals_host_unpack(mode, cuda_device):
    """NetCDF-style packed met columns (int16, scale_factor / add_offset): the device unpack + unit conversion equals
    the host statement `int16 * scale + offset` followed by the driver's conversions, bit for bit; 12 B per cell-step
    cross PCIe; quantisation stays inside the resolution of the packing."""
    import torch

    from topoflow_glacier_b200.forcing import (DEFAULT_PACKING, ForcingStreamer, convert_on_host, pack_forcing,
                                                 unpack_forcing)

    case = load_case("cats288")
    N, T = case["N"], 50
    rng = np.random.default_rng(9)
    raw = np.stack([rng.exponential(0.4, (T, N)), 273.15 + rng.normal(0, 8, (T, N)), 88900 + rng.normal(0, 300, (T, N)),
                    rng.uniform(5e-4, 1.2e-2, (T, N)), rng.normal(0, 3, (T, N)), rng.normal(0, 3, (T, N))], axis=1)
    packed = pack_forcing(raw)
    assert packed.dtype == np.int16 and np.abs(unpack_forcing(packed) - raw).max(axis=(0, 2)).tolist() <= (DEFAULT_PACKING[0] / 2 + 1e-12).tolist()
    want = convert_on_host(unpack_forcing(packed))
    eng = make_engine(case, mode=mode)
    st = ForcingStreamer(eng, chunk_steps=16, raw_dtype="int16")
    got = torch.cat([c.clone() for c in st.chunks(packed)]).cpu().numpy()
    assert st.h2d_bytes == packed.size * 2
    if mode == "f64":
        assert np.array_equal(got, want)
    else:
        assert np.array_equal(got, want.astype(np.float32))
    eng.close()


def test_netcdf_file_to_packed_streamer_to_kernel(tmp_path, cuda_device):
    """A NetCDF forcing file with int16-packed variables (scale_factor / add_offset) goes to the device UNCHANGED
    (12 B per cell-step), is unpacked and converted there, and drives the fused kernel: the result equals the oracle run
    on the host statement of the same unpacking (`unpack_forcing` -> `convert_on_host`) within the float64 tolerance."""
    import pandas as pd
    import torch

    from topoflow_glacier_b200.forcing import (ForcingStreamer, convert_on_host, read_forcing_csv, read_forcing_netcdf,
                                                 unpack_forcing, write_forcing_netcdf)

    root = GOLDEN.parent.parent
    case = load_case("cats288")
    csv = root / "tests" / "data" / "sample-cat-3062920.csv"
    raw = read_forcing_csv(csv)
    when = pd.DatetimeIndex(pd.to_datetime(pd.read_csv(csv)["Time"]))
    nc = tmp_path / "cats.nc"
    write_forcing_netcdf(nc, when, np.repeat(raw[:, :, None], case["N"], axis=2))
    packed, packing = read_forcing_netcdf(nc)
    assert packed.dtype == np.int16 and packed.shape == (288, 6, case["N"])
    forcing = convert_on_host(unpack_forcing(packed, packing))
    want = make_oracle(dict(case, forcing=forcing), strict_pow=True).run(forcing, record=("M_total", "h_swe", "h_iwe", "Q_sum"))
    eng = make_engine(case, mode="f64_fast")
    st = ForcingStreamer(eng, chunk_steps=100, raw_dtype="int16", packing=packing)
    parts = [eng.run(c, c.shape[0], record=("M_total", "h_swe", "h_iwe", "Q_sum")) for c in st.chunks(packed)]
    assert st.h2d_bytes == packed.size * 2
    for k in want:
        got = torch.cat([pp[k] for pp in parts]).cpu().numpy()
        ok, ratio, dabs, drel = err_report(got, want[k], ATOL[k])
        assert ok, (k, ratio, dabs, drel)
    eng.close()


@pytest.mark.parametrize("mode", ["f64_fast", "f64"])
def test_column_terms_pass_is_bit_identical_and_taken(mode, cuda_device):
    """TFG_OPT_COLUMN_TERMS: with a forcing map of few columns a float64 engine (fast or strict) evaluates the forcing-only part
    of update() once per column and timestep (column_terms_kernel) -- every recorded quantity, the state, the snowfall
    window, the diagnostic integrals and the exact basin sums must equal the per-cell evaluation bit for bit, also where
    a column holds missing or absurd forcing (those cell-steps take the strict step from the raw values in both)."""
    import torch

    import bench
    from helpers import default_constants
    from topoflow_glacier_b200 import _lib
    from topoflow_glacier_b200.engine import MeltEngine
    from topoflow_glacier_b200.sharding import BasinAggregates

    N, M, T = 3000, 24, 150                                        # 150 steps: two launches (128 + 22)
    statics, _ = bench.synthetic_host_sample(N, 1, seed=11)
    _, fcols = bench.synthetic_host_sample(M, T, seed=12)          # [T, 5, M]
    fcols = np.ascontiguousarray(fcols)
    fcols[17, 1, 3] = np.nan                                       # missing air temperature in column 3
    fcols[40:44, 4, 5] = 0.0                                       # calm
    fcols[60, 2, 7] = 5.0e6                                        # absurd pressure: strict step
    fcols[61, 3, 7] = 0.5                                          # absurd humidity
    fcols[90, 0, 9] = -1.0e-3                                      # negative precipitation
    col = (np.arange(N) * M // N).astype(np.int32)                 # contiguous catchments ...
    col[::7] = np.random.default_rng(5).integers(0, M, col[::7].size)   # ... with foreign cells inside the warps
    basin = (np.arange(N) // 100).astype(np.int32)
    kw = dict(zones=[-8.0], mode=mode, horizon_steps=T + 1, forcing_index=col, n_forcing_cols=M,
              basin_id=basin, n_basin=30)
    f = torch.as_tensor(fcols).to(cuda_device, torch.float64).contiguous()
    out = {}
    for on in (True, False):
        e = MeltEngine(statics, default_constants(), "2013020100", column_terms=on, **kw)
        agg = BasinAggregates(T, 30, device=cuda_device, exponents=e.agg_exponents())
        rec = e.run(f, record=_lib.REC_NAMES, basin_agg=agg.accumulator)
        half = e.run(f[:T // 2].contiguous())                      # and a launch without recording / aggregates
        torch.cuda.synchronize()
        out[on] = (rec, e.state.clone(), e.ring.clone(), agg.accumulator.clone(), e.column_term_launches)
        assert half == {}
        e.close()
    assert out[True][4] == 3 and out[False][4] == 0                # 2 + 1 launches took the pass
    for name in _lib.REC_NAMES:
        a, b = out[True][0][name], out[False][0][name]
        assert torch.equal(a.view(torch.int64), b.view(torch.int64)), name   # bit patterns: NaN-poisoned cells included
    for i in (1, 2):   # state (incl. the diagnostic integrals) and snowfall window
        assert torch.equal(out[True][i].view(torch.int64), out[False][i].view(torch.int64))
    assert torch.equal(out[True][3], out[False][3])                # exact basin sums and the left-out counter
    assert torch.isnan(out[True][1]).any()                         # the missing value did poison its cells


@pytest.mark.parametrize("mode", ["f64_fast", "f64"])
def test_column_terms_on_the_per_step_path(mode, cuda_device):
    """The literal update() (one-step launches, exact window re-sum) with the column-term pass in front: same state as
    without it, and as the fused launch, bit for bit (large enough for the pass to be taken: >= 65536 cell-steps)."""
    import torch

    import bench
    from helpers import default_constants
    from topoflow_glacier_b200.engine import MeltEngine

    N, M, T = 66000, 16, 5
    statics, _ = bench.synthetic_host_sample(N, 1, seed=21)
    _, fcols = bench.synthetic_host_sample(M, T, seed=22)
    col = (np.arange(N) % M).astype(np.int32)
    kw = dict(zones=[-8.0], mode=mode, horizon_steps=T + 1, forcing_index=col, n_forcing_cols=M)
    f = torch.as_tensor(np.ascontiguousarray(fcols)).to(cuda_device, torch.float64).contiguous()
    states = []
    for on, fused in ((True, False), (False, False), (True, True)):
        e = MeltEngine(statics, default_constants(), "2013020100", column_terms=on, **kw)
        if fused:
            e.run(f)
        else:
            for t in range(T):
                e.inputs[:5].copy_(f[t])
                e.step()
        torch.cuda.synchronize()
        assert (e.column_term_launches > 0) == on
        states.append((e.state.clone(), e.ring.clone()))
        e.close()
    for s, r in states[1:]:
        assert torch.equal(states[0][0], s) and torch.equal(states[0][1], r)


def test_write_combined_host_block_streams_like_a_pinned_one(cuda_device):
    """forcing.pinned_block (tfg_host_alloc, write-combined pages): recognised as page-locked memory, taken by the
    streamer without a staging copy, same device forcing and state as the NumPy source; freed with the tensor."""
    import gc

    import torch

    import bench
    from helpers import default_constants
    from topoflow_glacier_b200.engine import MeltEngine
    from topoflow_glacier_b200.forcing import ForcingStreamer, pack_forcing, pinned_block

    N, T = 512, 12
    statics, f = bench.synthetic_host_sample(N, T, seed=31)
    raw = np.stack([f[:, 0] * 1e3, f[:, 1] + 273.15, f[:, 2], f[:, 3], f[:, 4] * 0.6, f[:, 4] * 0.8], axis=1)
    packed = pack_forcing(raw)
    blk = pinned_block(packed.shape, torch.int16)
    from topoflow_glacier_b200 import _lib

    assert _lib.load().tfg_host_is_pinned(blk.data_ptr()) == 1 and blk.dtype == torch.int16 and tuple(blk.shape) == packed.shape
    assert _lib.load().tfg_host_is_pinned(packed.ctypes.data) == 0
    blk.copy_(torch.as_tensor(packed))
    states = []
    for src in (blk, packed):
        e = MeltEngine(statics, default_constants(), "2013020100", zones=[-8.0], mode="f64_fast", horizon_steps=T + 1)
        s = ForcingStreamer(e, chunk_steps=5, raw_dtype="int16")
        s.drive(src)
        torch.cuda.synchronize()
        if src is blk:
            assert all(p is None for p in s.pinned)            # no staging copy: the block is DMA-ed from where it lies
        states.append(e.state.clone())
        e.close()
    assert torch.equal(states[0], states[1])
    del blk
    gc.collect()
