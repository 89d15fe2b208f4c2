import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from bench import synthetic_host_sample
from oracle.np_ref import CellStatics, Constants, OracleModel
from topoflow_glacier_b200.engine import MeltEngine
from topoflow_glacier_b200.config import default_constants
from helpers import ATOL, err_report
N, T = 8192, 24
st, f = synthetic_host_sample(N, T)
REC = ["h_snow","h_swe","SM","h_ice","h_iwe","IM","M_total","RH","p0","T_dew","T_surf","Ri","Dn","Dh","Qh","W_p","Qe","TSN_offset","albedo","n","Qn_SW","em_air","Qn_LW","Q_sum","Eccs","Ecci","snow3day","P_rain","P_snow"]
ora = OracleModel(CellStatics(**st, tz=[-8.0]), Constants(), "2012100100", strict_pow=False)
want = {k: np.empty((T, N)) for k in REC}
for t in range(T):
    d = ora.step(*f[t])
    for k in REC: want[k][t] = d[k]
for mode in ("f64", "f64_fast"):
    eng = MeltEngine(st, default_constants(), "2012100100", zones=[-8.0], mode=mode, horizon_steps=T+1)
    got = eng.run(torch.as_tensor(f).cuda(), record=REC)
    got = {k: v.cpu().numpy() for k, v in got.items()}
    print("==", mode)
    for k in REC:
        ok, ratio, dabs, drel = err_report(got[k], want[k], ATOL[k])
        if not ok:
            idx = np.unravel_index(np.nanargmax(np.abs(got[k]-want[k])), got[k].shape)
            print(f"{k:10s} ratio={ratio:.3g} abs={dabs:.3g} rel={drel:.3g} at t,c={idx} got={got[k][idx]!r} want={want[k][idx]!r}")
    bad = np.argwhere(~np.isclose(got["Q_sum"], want["Q_sum"], rtol=1e-9, atol=1e-6))
    print("bad Q_sum count", len(bad), bad[:5])
    if len(bad):
        t, c = bad[0]
        print({k: (got[k][t, c], want[k][t, c]) for k in REC})
        print({k: v[c] for k, v in st.items()}, f[t, :, c])
    eng.close()
