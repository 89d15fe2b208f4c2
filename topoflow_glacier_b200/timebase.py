"""Clock-only quantities of ``update()`` as ``[T]`` host tables (kernel input ``tfg_time_row``).

Everything here depends on the timestamp alone, so it is evaluated once per step on the host instead of
once per cell-step on the device -- with the reference's own expressions, in the same order, so the bits
match what every reference instance would compute:

* the clock advance + decimal day of year + clock hour of ``update_julian_day``
  (reference ``bmi_topoflow_glacier.py:957-991``; step ``k`` uses ``start + (k+1)*dt`` because the
  clock is advanced before it is used, ``:962``);
* ``Day_Angle``, ``Declination``, ``Eccentricity_Correction`` (``solar_funcs.py:156-247``), ``Solar_Constant`` (``:151``);
* ``Equation_Of_Time`` (``solar_funcs.py:1301-1429``) incl. ``Earth_Perihelion`` (``:1142-1256``) and
  ``Vernal_Equinox`` (``:1111-1137``);
* the UTC offset of ``gmt_offset_hours`` (``solar_funcs.py:1616-1637``) for a given IANA zone or fixed
  offset (the polygon lookup ``timezonefinder`` does is replaced by configuration data).
"""

from __future__ import annotations

from datetime import datetime

import numpy as np
import pandas as pd

__all__ = ["parse_start", "time_tables", "utc_offsets", "default_timezone", "PERIHELION"]

# year -> (day of January, hour UTC) of Earth's perihelion, 1981..2060 (astropixels ephemeris, as tabulated
# at solar_funcs.py:1167-1248); stored compactly as day*100 + hour
_PERI = """
202 411 215 322 320 205 423 300 122 417 303 315 403 206 411 407 200 421 313 305
409 214 405 418 201 415 320 300 415 300 319 500 205 412 407 223 414 306 305 508
214 407 416 301 413 317 303 512 218 310 421 305 412 405 301 514 304 305 507 312
322 409 222 513 315 301 512 318 310 420 306 509 322 218 512 404 303 504 311 423
"""
PERIHELION = {1981 + i: divmod(int(tok), 100) for i, tok in enumerate(_PERI.split())}
assert len(PERIHELION) == 80 and PERIHELION[2013] == (2, 5) and PERIHELION[2060] == (4, 23)


def parse_start(s) -> pd.Timestamp:
    """'YYYYMMDDHH' or 'YYYYMMDD-HH' -> Timestamp (reference ``_parse_yyyymmddhh``, ``:512-517``)."""
    s = str(s).strip()
    d = datetime.strptime(s, "%Y%m%d-%H" if "-" in s else "%Y%m%d%H")  # ValueError if malformed
    return pd.Timestamp(year=d.year, month=d.month, day=d.day, hour=d.hour)


def _perihelion_jd(years: np.ndarray) -> np.ndarray:
    this_year = datetime.now().year  # reference quirk for years outside the table (solar_funcs.py:1158-1162)
    out = np.empty(years.shape, dtype=np.float64)
    for y in np.unique(years):
        d, h = PERIHELION[int(y) if 1981 <= int(y) <= 2060 else this_year]
        out[years == y] = np.float64(max(d - 1, 0)) + (h / np.float64(24))
    return out


def time_tables(start: pd.Timestamp, dt_hours, n_steps: int) -> dict[str, np.ndarray]:
    """Tables for steps ``0..n_steps-1``; step ``k`` is evaluated at ``start + (k+1)*dt``."""
    when = start + pd.to_timedelta(np.arange(1, n_steps + 1) * dt_hours, unit="h")
    when = pd.DatetimeIndex(when)
    years = when.year.values.astype(np.int64)
    jd = (when.dayofyear.values - 1 + when.hour.values / 24 + when.minute.values / 1440 + when.second.values / 86400)
    clock_hour = (jd - np.floor(jd)) * np.float64(24)

    twopi = np.float64(2) * np.pi
    # Equation_Of_Time
    e = np.float64(0.016713)
    eps = np.float64(23.4397) * (np.pi / np.float64(180))
    dpy = np.float64(365.2425)
    tp_jd = _perihelion_jd(years)
    M = (twopi / dpy) * (jd - tp_jd)
    M = (M + twopi) % twopi
    ve_jd = np.float64(79.3125) + dpy * (years - np.float64(2000))
    PT = (np.float64(365) + tp_jd) - ve_jd
    omega_p = twopi * (PT / dpy)
    L = M + omega_p
    half_eps_sq = float(eps / 2) ** 2.0
    TE = (-2.0 * e * np.sin(M)) + (np.sin(2 * L) * half_eps_sq)
    TE = TE / (np.float64(2) * np.pi / np.float64(24))

    # Day_Angle / Declination / Eccentricity_Correction
    G = (2 * np.pi) * jd / np.float64(365)
    delta = (np.float64(0.006918) - (np.float64(0.399912) * np.cos(G)) + (np.float64(0.070257) * np.sin(G))
             - (np.float64(0.006758) * np.cos(np.float64(2) * G)) + (np.float64(0.000907) * np.sin(np.float64(2) * G))
             - (np.float64(0.002697) * np.cos(np.float64(3) * G)) + (np.float64(0.001480) * np.sin(np.float64(3) * G)))
    E0 = (np.float64(1.000110) + (np.float64(0.034221) * np.cos(G)) + (np.float64(0.001280) * np.sin(G))
          + (np.float64(0.000719) * np.cos(np.float64(2) * G)) + (np.float64(0.000077) * np.sin(np.float64(2) * G)))
    omega = (np.float64(360) / np.float64(24)) * (np.pi / np.float64(180))
    hour_angle = omega * ((clock_hour - np.float64(12)) - TE)  # omega*th + omega*LC (fast modes)
    return {
        "cos_hour": np.cos(hour_angle), "sin_hour": np.sin(hour_angle),
        "when": when, "julian_day": jd, "clock_hour": clock_hour, "TE": TE, "delta": delta, "E0": E0,
        "sin_decl": np.sin(delta), "cos_decl": np.cos(delta), "tan_decl": np.tan(delta),
        "isc_e0": np.float64(1361.5) * E0,
    }


def utc_offsets(when: pd.DatetimeIndex, zones) -> np.ndarray:
    """``[T, n_zones]`` UTC offsets in hours (DST included); a zone is an IANA name or a number of hours."""
    cols = []
    for z in zones:
        if isinstance(z, (int, float, np.integer, np.floating)):
            cols.append(np.full(len(when), float(z)))
        else:
            local = when.tz_localize("UTC").tz_convert(str(z)).tz_localize(None)
            cols.append(np.asarray((local - when) / pd.Timedelta(hours=1), dtype=np.float64))
    return np.ascontiguousarray(np.stack(cols, axis=1))


def default_timezone(lat: float, lon: float):
    """Zone used when the configuration names none.

    The reference asks ``timezonefinder`` (not installable here, unpinned upstream).  If that package is
    importable it is used; otherwise a coarse table covers the conterminous US + Alaska (where the shipped
    catchments lie) and everything else falls back to the nautical offset ``round(lon/15)``.
    """
    try:  # pragma: no cover - optional dependency
        from timezonefinder import TimezoneFinder

        name = TimezoneFinder().timezone_at(lat=lat, lng=lon)
        if name:
            return name
    except Exception:  # noqa: BLE001
        pass
    if 24.0 <= lat <= 50.0 and -125.0 <= lon <= -66.0:
        if lon < -114.0:
            return "America/Los_Angeles"
        if lon < -102.0:
            return "America/Denver"
        if lon < -87.0:
            return "America/Chicago"
        return "America/New_York"
    if 51.0 <= lat <= 72.0 and -170.0 <= lon <= -130.0:
        return "America/Anchorage"
    return float(round(lon / 15.0))
