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
