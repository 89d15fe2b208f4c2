"""Synthetic glacierised grids generated on the device (benchmark workloads of BASELINE.json cfg 4/5).

Large rasters (16.7 M - 100 M cells) are too big to build on the host and ship over PCIe, so the
per-cell attributes are drawn with ``torch`` on the GPU and the cell-only tables of ``statics.cell_tables``
are evaluated there with the same formulas (device libm, so the last bit may differ from the host tables --
irrelevant for a synthetic grid; real catchments always go through the host path).  The raw attributes are
kept so that a sample of cells can be handed to the CPU oracle.
"""

from __future__ import annotations

import math

import torch

__all__ = ["synthetic_cells"]


def synthetic_cells(n: int, seed: int, device, *, M_mass_air: float = 0.0289644, g: float = 9.81,
                    rho_H2O: float = 1000.0, rho_snow: float = 50.0, rho_ice: float = 917.0,
                    glacier_fraction: float = 0.4, cell_km2: float = 9e-4) -> dict:
    """SURVEY.md 8(d) cfg 4: lat U(46.5,47.1), lon U(-122.1,-121.4), elev U(1200,4300) m, slope U(0,120),
    aspect U(0,360), 30 m cells, SWE U(0,1.5) m, IWE U(0,60) m on a ``glacier_fraction`` mask."""
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    f64 = dict(dtype=torch.float64, device=device)

    def U(lo, hi):
        return lo + (hi - lo) * torch.rand(n, generator=gen, **f64)

    raw = {"lat": U(46.5, 47.1), "lon": U(-122.1, -121.4), "elev": U(1200.0, 4300.0), "slope": U(0.0, 120.0),
           "aspect": U(0.0, 360.0)}
    raw["da"] = torch.full((n,), cell_km2, **f64)
    raw["h0_swe"] = U(0.0, 1.5)
    raw["h0_snow"] = raw["h0_swe"] * (rho_H2O / rho_snow)
    raw["h0_iwe"] = U(0.0, 60.0) * (torch.rand(n, generator=gen, **f64) < glacier_fraction)
    raw["h0_ice"] = raw["h0_iwe"] * (rho_H2O / rho_ice)
    raw["T_rain_snow"] = torch.zeros(n, **f64)

    twopi = 2.0 * math.pi
    omega = (360.0 / 24.0) * (math.pi / 180.0)
    alpha = torch.remainder(twopi + ((math.pi / 2) - raw["aspect"]), twopi)
    beta = torch.remainder(twopi + torch.atan(raw["slope"]), twopi)
    lat_rad = raw["lat"] * (math.pi / 180.0)
    sin_b, cos_b = torch.sin(beta), torch.cos(beta)
    lat_eq = torch.asin((sin_b * torch.cos(alpha) * torch.cos(lat_rad)) + (cos_b * torch.sin(lat_rad)))
    dlon = torch.atan((sin_b * torch.sin(alpha)) /
                      ((cos_b * torch.cos(lat_rad)) - (sin_b * torch.sin(lat_rad) * torch.cos(alpha))))
    lat_eq_rt = (lat_eq * (180.0 / math.pi)) * (math.pi / 180.0)
    tabs = {
        "a_elev": (-M_mass_air * g) * raw["elev"],
        "sin_lat": torch.sin(lat_rad), "cos_lat": torch.cos(lat_rad), "neg_tan_lat": -torch.tan(lat_rad),
        "lon": raw["lon"], "sin_lat_eq": torch.sin(lat_eq), "cos_lat_eq": torch.cos(lat_eq),
        "neg_tan_lat_eq": -torch.tan(lat_eq_rt), "dlon": dlon, "t_noon": -dlon / omega,
        "da_m2": raw["da"] * 1e6, "t_rain_snow": raw["T_rain_snow"],
    }
    tabs.update({k: raw[k] for k in ("h0_snow", "h0_ice", "h0_swe", "h0_iwe")})
    tabs["raw"] = raw
    return tabs
