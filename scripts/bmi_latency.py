"""Per-step BMI driver latency (the reference's usage pattern): 7 x set_value, update, 8 x get_value, N = 1."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
from topoflow_glacier import BmiTopoflowGlacier
z = np.load('tests/golden/cats288.npz')
keys = ("da", "slope", "aspect", "lon", "lat", "elev", "h0_snow", "h0_ice", "h0_swe", "h0_iwe", "T_rain_snow")
cfg = dict({k: float(z[f"static_{k}"][1]) for k in keys}, site_prefix="c", forcing_file="-", dt=1, start_time="2013032000", end_time="2013033123")
m = BmiTopoflowGlacier(); m.initialize_ensemble([cfg])
f = z["forcing"][:, :, 1]
names_in = ("atmosphere_water__liquid_equivalent_precipitation_rate", "land_surface_air__temperature", "land_surface_air__pressure",
            "atmosphere_air_water~vapor__relative_saturation", "wind_speed_UV", "land_surface_radiation~incoming~longwave__energy_flux",
            "land_surface_radiation~incoming~shortwave__energy_flux")
names_out = m.get_output_var_names()
dest = np.zeros(1)
def loop(n, sets=True, upd=True, gets=True):
    t0 = time.perf_counter()
    for i in range(n):
        if sets:
            for name, v in zip(names_in, list(f[i % 288]) + [300.0, 100.0]): m.set_value(name, v)
        if upd: m.update()
        if gets:
            for name in names_out: m.get_value(name, dest)
    return (time.perf_counter() - t0) / n * 1e6
loop(50)
print("full step  %.1f us" % loop(500))
print("sets only  %.1f us" % loop(500, True, False, False))
print("set+update %.1f us" % loop(500, True, True, False))
print("update+get %.1f us" % loop(500, False, True, True))

# BASELINE configs[2]: the four shipped catchments as one device-resident ensemble, one water year, one launch
import torch
zy = np.load('tests/golden/year4.npz')
cfgs = [dict({k: float(zy[f"static_{k}"][i]) for k in keys}, site_prefix=f"c{i}", forcing_file="-", dt=1,
             start_time="2012100100", end_time="2013093023") for i in range(4)]
for prec in ("f64", "f64_fast"):
    ens = BmiTopoflowGlacier(); ens.initialize_ensemble(cfgs, mode=prec)
    fy = torch.as_tensor(np.repeat(zy["forcing"], 4, axis=2)).cuda().contiguous()
    ens.update_steps(24, fy[:24].contiguous()); torch.cuda.synchronize()
    t0 = time.perf_counter(); ens.update_steps(8736, fy[24:].contiguous()); torch.cuda.synchronize()
    dt_ = time.perf_counter() - t0
    print(f"4-catchment ensemble, 8736 hourly steps, one launch ({prec}): {dt_*1e3:.1f} ms = {dt_/8736*1e6:.2f} us/step")
    ens.finalize()
