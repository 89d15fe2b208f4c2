"""Cell-only tables (kernel input ``tfg_statics``), evaluated once on the host.

The reference recomputes these every step although they never change:
``set_aspect_angle`` / ``set_slope_angle`` (reference ``bmi_topoflow_glacier.py:1082-1113``),
``Equivalent_Latitude`` (``solar_funcs.py:741-767``), ``Longitude_Offset`` (``:718-736``),
``Noon_Offset_Slope`` (``:772-778``), the latitude trigonometry of ``Zenith_Angle`` / ``Sunrise_Offset``
(``:280-284``, ``:320-326``) and the numerator of the barometric exponent (``bmi_topoflow_glacier.py:552``).
Expressions are kept in the reference's order so the float64 tables carry the same bits.
"""

from __future__ import annotations

import numpy as np

__all__ = ["cell_tables", "CELL_KEYS"]

CELL_KEYS = ("da", "slope", "aspect", "lon", "lat", "elev", "h0_snow", "h0_ice", "h0_swe", "h0_iwe", "T_rain_snow")


def cell_tables(lat, lon, slope, aspect, elev, da_km2, t_rain_snow, *, M_mass_air: float, g: float) -> dict:
    """float64 ``[N]`` tables named after the fields of ``tfg_statics``."""
    f = lambda a: np.atleast_1d(np.asarray(a, dtype=np.float64))  # noqa: E731
    lat, lon, slope, aspect, elev, da_km2, t_rain_snow = map(f, (lat, lon, slope, aspect, elev, da_km2, t_rain_snow))
    twopi = np.float64(2) * np.pi
    omega = (np.float64(360) / np.float64(24)) * (np.pi / np.float64(180))
    # aspect is used as radians and slope is fed to arctan as is -- reference behaviour, not a unit fix
    alpha = (np.pi / 2) - aspect
    alpha = (twopi + alpha) % twopi
    alpha = np.where(np.isfinite(alpha), alpha, 0.0)
    beta = np.arctan(slope)
    beta = (twopi + beta) % twopi
    beta = np.where(np.isfinite(beta), beta, 0.0)
    if np.any((beta < 0) | (beta > np.pi / 2)):
        raise ValueError("slope angle outside [0, pi/2]")
    lat_rad = lat * (np.pi / np.float64(180))
    sin_b, cos_b = np.sin(beta), np.cos(beta)
    lat_eq = np.arcsin((sin_b * np.cos(alpha) * np.cos(lat_rad)) + (cos_b * np.sin(lat_rad)))
    with np.errstate(divide="ignore", invalid="ignore"):
        dlon = np.arctan((sin_b * np.sin(alpha)) / ((cos_b * np.cos(lat_rad)) - (sin_b * np.sin(lat_rad) * np.cos(alpha))))
    t_noon = -np.float64(1) * dlon / omega
    lat_eq_rt = (lat_eq * (np.float64(180) / np.pi)) * (np.pi / np.float64(180))
    return {
        "a_elev": -M_mass_air * g * elev,
        "sin_lat": np.sin(lat_rad),
        "cos_lat": np.cos(lat_rad),
        "neg_tan_lat": -np.float64(1) * np.tan(lat_rad),
        "lon": lon,
        "sin_lat_eq": np.sin(lat_eq),
        "cos_lat_eq": np.cos(lat_eq),
        "neg_tan_lat_eq": -np.float64(1) * np.tan(lat_eq_rt),
        "dlon": dlon,
        "t_noon": t_noon,
        "da_m2": da_km2 * 1e6,
        "t_rain_snow": t_rain_snow,
    }
