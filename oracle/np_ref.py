"""CPU oracle: vectorised NumPy float64 restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``topoflow_glacier_b200/`` may import
this module; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / CPU baseline -- never as the thing shipped.

What it restates (paths relative to the upstream checkout):

* ``src/topoflow_glacier/bmi/bmi_topoflow_glacier.py:413-465``  ``update()``
  call order, and the bodies at ``:519-1777``;
* ``src/topoflow_glacier/physics/solar_funcs.py:141-953`` clear-sky shortwave,
  ``:958-1009`` Julian day, ``:1111-1137`` vernal equinox, ``:1142-1256``
  perihelion table, ``:1301-1480`` equation of time / true solar noon,
  ``:1616-1637`` UTC offset;
* defaults of ``src/topoflow_glacier/bmi/config.py:27-101``;
* initial state of ``bmi_topoflow_glacier.py:281-411``.

The reference only supports arrays of shape ``(1,)`` (one BMI instance per
catchment).  Here every per-cell quantity is an ``[N]`` array and cell ``i``
evolves exactly like an independent reference instance configured with cell
``i``'s static attributes.  Floating-point operations are kept in the
reference's order and association; no algebraic simplification is applied.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified
reference (imported from the upstream checkout with three stub modules) and
this oracle on the same inputs and requires bit-equality of every dumped
intermediate; the resulting vectors are committed under ``tests/golden/`` and
re-checked by ``tests/test_oracle_golden.py`` together with the reference's own
golden vector ``tests/data/output_m_total.npy`` (tolerance 1e-13 relative: that
file was produced on another machine/libm, see ``DESIGN.md``).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field, fields
from datetime import datetime
from zoneinfo import ZoneInfo

import numpy as np
import pandas as pd

__all__ = ["Constants", "CellStatics", "OracleModel", "TimeRow", "time_row", "FORCING_VARS", "OUTPUT_VARS"]

# order of the five live forcings in every [.., 5, N] forcing block
FORCING_VARS = ("P", "T_air", "P_air", "Hum_sp", "uz")
# order of the eight BMI outputs (bmi_topoflow_glacier.py:28-37), internal names
OUTPUT_VARS = ("h_snow", "h_swe", "SM", "h_ice", "h_iwe", "IM", "M_total", "RH")

RING_SLOTS_HOURS = 3 * 24  # bmi_topoflow_glacier.py:296  int(3*hours_per_day/dt) slots


@dataclass
class Constants:
    """Physical constants the path reads (config.py:27-101 defaults)."""

    dt: int = 1
    dust_atten: float = 0.08
    canopy_factor: float = 0.0
    cloud_factor: float = 0.0
    rho_air: float = 1.2614
    rho_snow: float = 50.0
    rho_ice: float = 917.0
    rho_H2O: float = 1000.0
    h_active_layer: float = 0.125
    T0: float = -0.2
    Cp_air: float = 1005.7
    Cp_ice: float = 2060.0
    Cp_snow: float = 2090.0
    g: float = 9.81
    Lf: float = 334000.0
    eps: float = 0.622
    kappa: float = 0.408
    latent_heat_constant: float = 0.622
    Lv: float = 2500000.0
    sigma: float = 5.67 * 10 ** (-8)
    sea_level_p0: float = 101325.0
    uni_gas_const: float = 8.3144598
    M_mass_air: float = 0.0289644
    z0_air: float = 0.01
    em_surf: float = 0.985
    SATTERLUND: bool = False

    @classmethod
    def from_mapping(cls, m) -> "Constants":
        names = {f.name for f in fields(cls)}
        return cls(**{k: v for k, v in dict(m).items() if k in names})


@dataclass
class CellStatics:
    """Per-cell static attributes + initial state, each an ``[N]`` float64 array."""

    da: np.ndarray  # km2
    slope: np.ndarray
    aspect: np.ndarray
    lon: np.ndarray
    lat: np.ndarray
    elev: np.ndarray
    h0_snow: np.ndarray
    h0_ice: np.ndarray
    h0_swe: np.ndarray
    h0_iwe: np.ndarray
    T_rain_snow: np.ndarray
    tz: list = field(default_factory=lambda: ["America/Los_Angeles"])  # one name, or one per cell

    @classmethod
    def from_configs(cls, cfgs, tz="America/Los_Angeles") -> "CellStatics":
        def col(k, default=None):
            return np.array([float(c.get(k, default)) for c in cfgs], dtype=np.float64)

        return cls(
            da=col("da"), slope=col("slope"), aspect=col("aspect", 0.0), lon=col("lon"), lat=col("lat"),
            elev=col("elev"), h0_snow=col("h0_snow"), h0_ice=col("h0_ice"), h0_swe=col("h0_swe"),
            h0_iwe=col("h0_iwe"), T_rain_snow=col("T_rain_snow", 1.0),
            tz=[tz] if isinstance(tz, (str, int, float)) else list(tz),
        )

    @property
    def n(self) -> int:
        return int(np.size(self.lat))


# ---------------------------------------------------------------------------------------------
# time-only quantities (scalar path of the reference, evaluated with the same scalar expressions)
# ---------------------------------------------------------------------------------------------

_TP_TABLE = {  # solar_funcs.py:1167-1248  (day of January, hour) of perihelion
    1981: (2, 2), 1982: (4, 11), 1983: (2, 15), 1984: (3, 22), 1985: (3, 20), 1986: (2, 5), 1987: (4, 23),
    1988: (3, 0), 1989: (1, 22), 1990: (4, 17), 1991: (3, 3), 1992: (3, 15), 1993: (4, 3), 1994: (2, 6),
    1995: (4, 11), 1996: (4, 7), 1997: (2, 0), 1998: (4, 21), 1999: (3, 13), 2000: (3, 5), 2001: (4, 9),
    2002: (2, 14), 2003: (4, 5), 2004: (4, 18), 2005: (2, 1), 2006: (4, 15), 2007: (3, 20), 2008: (3, 0),
    2009: (4, 15), 2010: (3, 0), 2011: (3, 19), 2012: (5, 0), 2013: (2, 5), 2014: (4, 12), 2015: (4, 7),
    2016: (2, 23), 2017: (4, 14), 2018: (3, 6), 2019: (3, 5), 2020: (5, 8), 2021: (2, 14), 2022: (4, 7),
    2023: (4, 16), 2024: (3, 1), 2025: (4, 13), 2026: (3, 17), 2027: (3, 3), 2028: (5, 12), 2029: (2, 18),
    2030: (3, 10), 2031: (4, 21), 2032: (3, 5), 2033: (4, 12), 2034: (4, 5), 2035: (3, 1), 2036: (5, 14),
    2037: (3, 4), 2038: (3, 5), 2039: (5, 7), 2040: (3, 12), 2041: (3, 22), 2042: (4, 9), 2043: (2, 22),
    2044: (5, 13), 2045: (3, 15), 2046: (3, 1), 2047: (5, 12), 2048: (3, 18), 2049: (3, 10), 2050: (4, 20),
    2051: (3, 6), 2052: (5, 9), 2053: (3, 22), 2054: (2, 18), 2055: (5, 12), 2056: (4, 4), 2057: (3, 3),
    2058: (5, 4), 2059: (3, 11), 2060: (4, 23),
}


def _perihelion_jd(year: int):
    """solar_funcs.py:1142-1256 (+ Julian_Day :958-1009 for January, no year given)."""
    if (year < 1981) or (year > 2060):
        year = int(datetime.now().year)  # reference quirk: silently uses the current year (:1158-1162)
    d, h = _TP_TABLE[year]
    jd = np.int64(0) + np.maximum(d - 1, 0)  # np.sum(month_days[:1]) == 0
    return jd + (h / np.float64(24))


def _equation_of_time_hours(julian_day: float, year: int):
    """solar_funcs.py:1301-1429 with DEGREES=DMS=False."""
    e = np.float64(0.016713)  # :1106
    eps = np.float64(23.4397) * (np.pi / np.float64(180))  # :1086-1089
    days_per_year = np.float64(365.2425)  # :1051
    tp_jd = _perihelion_jd(year)
    twopi = np.float64(2) * np.pi
    M = (twopi / days_per_year) * (julian_day - tp_jd)
    M = (M + twopi) % twopi
    ve_jd = np.float64(79.3125) + days_per_year * (year - np.float64(2000))  # :1133-1135
    PT = (np.float64(365) + tp_jd) - ve_jd
    omega = twopi * (PT / days_per_year)
    L = M + omega
    TE = (-2.0 * e * np.sin(M)) + (np.sin(2 * L) * (eps / 2) ** 2.0)
    spin_rate = np.float64(2) * np.pi / np.float64(24)  # :1070
    return TE / spin_rate


def _declination(day_angle):
    """solar_funcs.py:221-229."""
    return (
        np.float64(0.006918)
        - (np.float64(0.399912) * np.cos(day_angle))
        + (np.float64(0.070257) * np.sin(day_angle))
        - (np.float64(0.006758) * np.cos(np.float64(2) * day_angle))
        + (np.float64(0.000907) * np.sin(np.float64(2) * day_angle))
        - (np.float64(0.002697) * np.cos(np.float64(3) * day_angle))
        + (np.float64(0.001480) * np.sin(np.float64(3) * day_angle))
    )


def _eccentricity(day_angle):
    """solar_funcs.py:192-198."""
    return (
        np.float64(1.000110)
        + (np.float64(0.034221) * np.cos(day_angle))
        + (np.float64(0.001280) * np.sin(day_angle))
        + (np.float64(0.000719) * np.cos(np.float64(2) * day_angle))
        + (np.float64(0.000077) * np.sin(np.float64(2) * day_angle))
    )


OMEGA = (np.float64(360) / np.float64(24)) * (np.pi / np.float64(180))  # solar_funcs.py:257-258
I_SC = np.float64(1361.5)  # solar_funcs.py:151


def utc_offset_hours(tz, when_utc: pd.Timestamp) -> float:
    """solar_funcs.py:1616-1637 with the polygon lookup replaced by a given zone.

    ``tz`` is an IANA name or a fixed offset in hours (number).
    """
    if isinstance(tz, (int, float)):
        return float(tz)
    local_time = when_utc.tz_localize("UTC").astimezone(ZoneInfo(tz))
    return local_time.utcoffset().total_seconds() / 3600.0


@dataclass
class TimeRow:
    """Everything in one ``update()`` that depends on the clock only."""

    when: pd.Timestamp
    year: int
    julian_day: float
    clock_hour: float
    TE: float
    delta: float
    E0: float


def time_row(when: pd.Timestamp) -> TimeRow:
    """bmi_topoflow_glacier.py:964-991 for the already-advanced clock ``when``."""
    julian_day = when.day_of_year - 1 + when.hour / 24 + when.minute / 1440 + when.second / 86400
    dec_part = julian_day - int(julian_day)
    clock_hour = dec_part * np.float64(24)
    TE = _equation_of_time_hours(julian_day, when.year)
    Gamma = (2 * np.pi) * julian_day / np.float64(365)  # solar_funcs.py:176
    return TimeRow(when, when.year, julian_day, clock_hour, TE, _declination(Gamma), _eccentricity(Gamma))


def _pairwise72(ring: np.ndarray) -> np.ndarray:
    """np.sum over the 72-slot snowfall window, one row per cell.

    bmi_topoflow_glacier.py:1035-1037 sums a contiguous 1-D array of 72 doubles;
    NumPy evaluates that with eight interleaved accumulators.  Summing a
    C-contiguous ``[N, 72]`` array along axis 1 uses the same inner kernel per row.
    """
    return np.sum(np.ascontiguousarray(ring), axis=1)


def pairwise72_explicit(ring: np.ndarray) -> np.ndarray:
    """Spelled-out order of the above (what the CUDA kernel implements); used by tests."""
    r = [ring[:, j].copy() for j in range(8)]
    for i in range(8, ring.shape[1], 8):
        for j in range(8):
            r[j] = r[j] + ring[:, i + j]
    return ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))


class OracleModel:
    """N independent reference instances advanced in lock step."""

    def __init__(self, cells: CellStatics, consts: Constants | None = None, start_time: str = "2013032000",
                 strict_pow: bool = True):
        self.c = consts or Constants()
        self.cells = cells
        self.strict_pow = strict_pow
        c, s = self.c, cells
        n = s.n
        self.N = n
        self.dt = c.dt
        self.days_per_dt = self.dt / 86400  # :287 (sic)
        self.da_m2 = s.da * 1e6  # :294
        self.one_seventh = np.float64(1) / 7
        self.twopi = np.float64(2) * np.pi
        self.z = np.full(n, 10.0)  # :301
        self.rho_H2O = np.float64(c.rho_H2O)
        self.rho_snow = np.float64(c.rho_snow)
        self.rho_ice = np.float64(c.rho_ice)
        self.ws_density_ratio = self.rho_H2O / self.rho_snow  # :385
        self.wi_density_ratio = self.rho_H2O / self.rho_ice
        # state (:350-369)
        self.h_snow = s.h0_snow.astype(np.float64).copy()
        self.h_ice = s.h0_ice.astype(np.float64).copy()
        self.h_swe = s.h0_swe.astype(np.float64).copy()
        self.h_iwe = s.h0_iwe.astype(np.float64).copy()
        self.albedo = np.full(n, 0.3)
        self.n = np.zeros(n)
        slots = int(3 * np.float64(24) / self.dt)
        self.ring = np.zeros((n, slots))  # logical order: column 0 oldest ... column -1 newest
        # cold content (:389-395) with T_surf = 0 at initialize
        T0 = np.full(n, c.T0)
        self.T0_cc = T0
        del_T = T0 - np.zeros(n)
        self.Eccs = np.maximum((self.rho_snow * np.float64(c.Cp_snow)) * self.h_snow * del_T, 0.0)
        self.Ecci = np.maximum((self.rho_ice * np.float64(c.Cp_ice)) * np.full(n, c.h_active_layer) * del_T, 0.0)
        # outputs that exist before the first update
        self.SM = np.zeros(n)
        self.IM = np.zeros(n)
        self.M_total = np.zeros(n)
        self.RH = np.zeros(n)
        # diagnostic integrals (:314-317, :362-363)
        self.vol_P = np.zeros(n)
        self.vol_PR = np.zeros(n)
        self.vol_PS = np.zeros(n)
        self.vol_SM = np.zeros(n)
        self.vol_IM = np.zeros(n)
        self.P_max = np.zeros(n)
        # clock (:398-411)
        st = str(start_time).strip()
        fmt = "%Y%m%d-%H" if "-" in st else "%Y%m%d%H"
        d0 = datetime.strptime(st, fmt)
        self.now = pd.Timestamp(year=d0.year, month=d0.month, day=d0.day, hour=d0.hour)
        self.step_index = 0
        # cell-only solar geometry (set_aspect_angle :1082, set_slope_angle :1095, solar_funcs :718-778)
        alpha = (np.pi / 2) - s.aspect
        alpha = (self.twopi + alpha) % self.twopi
        alpha = np.where(np.isfinite(alpha), alpha, 0.0)
        beta = np.arctan(s.slope)
        beta = (self.twopi + beta) % self.twopi
        beta = np.where(np.isfinite(beta), beta, 0.0)
        if np.any((beta < 0) | (beta > np.pi / 2)):
            raise ValueError("slope angle out of range (reference logs an error and keeps a stale beta)")
        self.alpha, self.beta = alpha, beta
        lat_rad = s.lat * (np.pi / np.float64(180))
        self.lat_rad = lat_rad
        # Equivalent_Latitude :753-757
        t1 = np.sin(beta) * np.cos(alpha) * np.cos(lat_rad)
        t2 = np.cos(beta) * np.sin(lat_rad)
        self.lat_eq = np.arcsin(t1 + t2)
        # Longitude_Offset :730-734
        u1 = np.sin(beta) * np.sin(alpha)
        u2 = np.cos(beta) * np.cos(lat_rad)
        u3 = np.sin(beta) * np.sin(lat_rad) * np.cos(alpha)
        with np.errstate(divide="ignore", invalid="ignore"):
            self.dlon = np.arctan(u1 / (u2 - u3))
        self.t_noon = -np.float64(1) * self.dlon / OMEGA  # :776
        eq_lat_deg = self.lat_eq * (np.float64(180) / np.pi)  # :764
        self.lat_eq_rt = eq_lat_deg * (np.pi / np.float64(180))  # round trip through degrees (:320)
        tz = s.tz
        self._tz = tz if len(tz) == n else [tz[0]] * n
        self._tz_unique = sorted(set(map(str, self._tz)), key=str)
        self.diag: dict[str, np.ndarray] = {}

    # -- helpers ------------------------------------------------------------------------------
    def _gmt_offsets(self, when: pd.Timestamp) -> np.ndarray:
        cache = {}
        out = np.empty(self.N)
        for i, tz in enumerate(self._tz):
            k = str(tz)
            if k not in cache:
                cache[k] = utc_offset_hours(tz, when)
            out[i] = cache[k]
        return out

    def _e_sat(self, T):
        """update_saturation_vapor_pressure :784-802 with MBAR=True."""
        if not self.c.SATTERLUND:
            term1 = (np.float64(17.3) * T) / (T + np.float64(237.3))
            e_sat = np.float64(0.611) * np.exp(term1)
        else:
            term1 = np.float64(2353) / (T + np.float64(273.15))
            e_sat = np.float64(10) ** (np.float64(11.4) - term1)
            e_sat = e_sat / np.float64(1000)
        return e_sat * np.float64(10)

    def _pow_scalar_path(self, base, expo):
        """``np.float64 ** float`` goes through libm ``pow`` (solar_funcs.py:567)."""
        if self.strict_pow:
            return np.array([math.pow(float(b), expo) for b in np.asarray(base).ravel()]).reshape(np.shape(base))
        return base**expo

    # -- one update() -------------------------------------------------------------------------
    def step(self, P, T_air, P_air, Hum_sp, uz) -> dict[str, np.ndarray]:
        c = self.c
        dt = self.dt
        P = np.asarray(P, dtype=np.float64) + np.zeros(self.N)
        T_air = np.asarray(T_air, dtype=np.float64) + np.zeros(self.N)
        P_air = np.asarray(P_air, dtype=np.float64) + np.zeros(self.N)
        Hum_sp = np.asarray(Hum_sp, dtype=np.float64) + np.zeros(self.N)
        uz = np.asarray(uz, dtype=np.float64) + np.zeros(self.N)
        d = self.diag = {}
        err = np.errstate(all="ignore")
        err.__enter__()

        # update_atm_pressure_from_elevation(T_C=True, MBAR=True) :550-556
        T_K = T_air + 273.15
        p0 = c.sea_level_p0 * np.exp(-c.M_mass_air * c.g * self.cells.elev / (c.uni_gas_const * T_K))
        p0 = p0 / np.float64(1000)
        p0 = p0 * np.float64(10.0)
        d["p0"] = p0
        # update_P_integral / P_max :567-576
        self.vol_P = self.vol_P + (P * self.da_m2 * dt)
        self.P_max = np.maximum(self.P_max, P)
        # update_P_rain / P_snow :585, :604
        P_rain = P * (T_air > self.cells.T_rain_snow)
        P_snow = P * (T_air <= self.cells.T_rain_snow)
        d["P_rain"], d["P_snow"] = P_rain, P_snow
        self.vol_PR = self.vol_PR + (P_rain * self.da_m2 * dt)
        self.vol_PS = self.vol_PS + (P_snow * self.da_m2 * dt)
        # vapour pressures :423-425
        e_sat_air = self._e_sat(T_air)
        e = Hum_sp * P_air / (c.eps + ((1 - c.eps) * Hum_sp))  # :817
        e = e / np.float64(1000)
        e_air = e * np.float64(10)
        RH = e_air / e_sat_air  # :838
        d["e_sat_air"], d["e_air"], d["RH"] = e_sat_air, e_air, RH
        # update_dew_point :888-893
        log_term = np.log(e_air / 6.1121)
        T_dew = 257.14 * log_term / (18.678 - log_term)
        d["T_dew"] = T_dew
        # update_T_surf :906-911
        T_surf = np.where((self.h_snow > 0) | (self.h_ice > 0), np.minimum(T_dew, np.float64(0)), T_dew)
        d["T_surf"] = T_surf
        e_sat_surf = self._e_sat(T_surf)
        d["e_sat_surf"] = e_sat_surf
        # update_bulk_richardson_number :640-644
        top = c.g * self.z * (T_air - T_surf)
        bot = (uz) ** 2.0 * (T_air + np.float64(273.15))
        bot = np.where(bot == 0.0, 0.01, bot)
        Ri = top / bot
        d["Ri"] = Ri
        # update_bulk_aero_conductance :670-733 (scalar-Ri branch, per cell)
        arg = c.kappa / np.log(np.maximum((self.z - self.h_snow) / c.z0_air, 0.01))
        Dn = uz * (arg) ** 2.0
        Dh = np.where(
            T_air == T_surf,
            Dn,
            np.where(Ri > 0, Dn / (np.float64(1) + (np.float64(10) * Ri)), Dn * (np.float64(1) - (np.float64(10) * Ri))),
        )
        De = Dh
        d["Dn"], d["Dh"] = Dn, Dh
        # update_sensible_heat_flux :744-745
        Qh = (c.rho_air * c.Cp_air) * Dh * (T_air - T_surf)
        d["Qh"] = Qh
        # update_precipitable_water_content :919-920
        W_p = np.float64(1.12) * np.exp(0.0614 * T_dew)
        d["W_p"] = W_p
        # update_vapor_pressure(SURFACE=True) :853
        e_surf = RH * e_sat_surf
        d["e_surf"] = e_surf
        # update_latent_heat_flux :931-934
        factor = c.rho_air * c.Lv * De
        Qe = factor * (e_air - e_surf) * (c.latent_heat_constant / p0)
        d["Qe"] = Qe
        Qa = np.zeros(self.N)
        Qc = np.zeros(self.N)

        # update_julian_day("hour") :957-1004 -- the clock advances BEFORE it is used
        self.now = self.now + pd.to_timedelta(dt, unit="h")
        tr = time_row(self.now)
        GMT = self._gmt_offsets(self.now)
        LC = (GMT * np.float64(15) - self.cells.lon) / np.float64(15)  # solar_funcs.py:1466-1468
        solar_noon = np.float64(12) + LC + tr.TE  # :1471
        th = tr.clock_hour - solar_noon  # bmi :1004
        d["TSN_offset"] = th

        # update_albedo("aging") :1021-1059
        r = np.where(T_air > 0, 0.12, 0.05)
        self.ring = np.roll(self.ring, -1, axis=1)
        self.ring[:, -1] = P_snow * dt * (self.rho_H2O / self.rho_snow)
        tot = _pairwise72(self.ring)
        d["snow3day"] = tot
        n = np.where(tot >= 0.03, 0, self.n)
        n = np.where(tot < 0.03, n + self.days_per_dt, n)
        self.n = n
        snow_albedo = 0.4 + 0.44 * np.exp(-n * r)
        albedo = np.where(self.h_snow > 0, snow_albedo, self.albedo)
        albedo = np.where((self.h_snow == 0) & (self.h_ice > 0), np.float64(0.3), albedo)
        albedo = np.where((self.h_snow == 0) & (self.h_ice == 0), np.float64(0.15), albedo)
        self.albedo = albedo
        d["albedo"], d["n"] = albedo, n

        # update_net_shortwave_radiation :1122-1139 -> Clear_Sky_Radiation solar_funcs.py:894-953
        K_cs = self._clear_sky(tr, W_p, th, albedo)
        Qn_SW = K_cs * (1 - albedo)
        d["K_cs"], d["Qn_SW"] = K_cs, Qn_SW

        # update_em_air :1167-1192
        T_air_K = T_air + 273.15
        if not c.SATTERLUND:
            e_air_kPa = e_air / np.float64(10)
            F, C = c.canopy_factor, c.cloud_factor
            term1 = (1.0 - F) * 1.72 * (e_air_kPa / T_air_K) ** self.one_seventh
            term2 = 1.0 + (0.22 * C**2.0)
            em_air = (term1 * term2) + F
        else:
            eterm = np.exp(-1 * (e_air) ** (T_air_K / 2016))
            em_air = 1.08 * (1.0 - eterm)
        d["em_air"] = em_air
        # update_net_longwave_radiation :1231-1248
        T_surf_K = T_surf + 273.15
        LW_in = em_air * c.sigma * (T_air_K) ** 4.0
        LW_out = c.em_surf * c.sigma * (T_surf_K) ** 4.0
        LW_out = LW_out + (1.0 - c.em_surf) * LW_in
        Qn_LW = LW_in - LW_out
        d["Qn_LW"] = Qn_LW
        # update_net_energy_flux :1314
        Q_sum = Qn_SW + Qn_LW + Qh + Qe + Qa + Qc
        d["Q_sum"] = Q_sum

        # ---- snow & ice ----
        previous_swe = self.h_swe.copy()  # :1571
        # update_snow_meltrate :1364-1368 ; enforce_max_snow_meltrate :1465
        E_in = Q_sum * dt
        E_rem = np.maximum(E_in - self.Eccs, np.float64(0))
        Qm = E_rem / dt
        SM = Qm / (self.rho_H2O * np.float64(c.Lf))
        SM = np.maximum(SM, np.float64(0))
        # update_SM_integral :1486-1487
        self.vol_SM = self.vol_SM + (SM * self.da_m2 * dt * 3600)
        # update_swe :1594-1606
        h_swe = self.h_swe + (P_snow * dt)
        SM_one_hour = np.minimum(SM * 3600, h_swe)
        SM = SM_one_hour / 3600
        h_swe = h_swe - (SM * dt * 3600)
        h_swe = np.maximum(h_swe, np.float64(0))
        self.h_swe = h_swe
        # update_snowfall_cold_content :1507-1537
        new_h_snow = (P_snow * dt) * self.ws_density_ratio
        T_wb = (
            T_air * np.arctan(0.151977 * ((RH + 8.313659) ** 0.5))
            + np.arctan(T_air + RH)
            - np.arctan(RH - 1.676331)
            + ((0.00391838 * (RH**1.5)) * np.arctan(0.023101 * RH))
            - 4.86035
        )
        d["T_wb"] = T_wb
        del_T = self.T0_cc - T_wb
        Eccs = np.where(
            P_snow > 0,
            np.maximum((self.Eccs + ((self.rho_snow * np.float64(c.Cp_snow)) * new_h_snow * del_T) - E_in), np.float64(0)),
            self.Eccs,
        )
        # update_ice_meltrate :1418-1428
        E_rem_i = np.maximum(E_in - self.Ecci, np.float64(0))
        M = (E_rem_i / dt) / (self.rho_H2O * np.float64(c.Lf))
        IM = np.maximum(M, np.float64(0))
        IM = np.where((h_swe == 0) & (previous_swe == 0), IM, np.float64(0))
        Ecci = np.maximum(self.Ecci - E_in, np.float64(0))
        Ecci = np.where(self.h_ice == 0, np.float64(0), Ecci)
        self.Ecci = Ecci
        # enforce_max_ice_meltrate :1473-1480
        IM = np.minimum(IM, self.h_iwe / dt)
        IM = np.maximum(IM, np.float64(0))
        # update_IM_integral :1493-1494
        self.vol_IM = self.vol_IM + (IM * self.da_m2 * dt * 3600)
        # update_iwe :1612-1617
        IM_one_hour = np.minimum(IM * 3600, self.h_iwe)
        IM = IM_one_hour / 3600
        h_iwe = self.h_iwe - (IM * dt * 3600)
        h_iwe = np.maximum(h_iwe, np.float64(0))
        self.h_iwe = h_iwe
        # update_combined_meltrate :1441-1445
        M_total = IM + SM + P_rain / 3600
        # update_snow_depth :1711 ; update_ice_depth :1726
        self.h_snow = h_swe * self.ws_density_ratio
        self.h_ice = h_iwe * self.wi_density_ratio
        # update_snowpack_cold_content :1552-1558
        Eccs = np.where(P_snow <= 0, np.maximum(Eccs - E_in, np.float64(0)), Eccs)
        Eccs = np.where(self.h_snow == 0, np.float64(0), Eccs)
        self.Eccs = Eccs

        self.SM, self.IM, self.M_total, self.RH = SM, IM, M_total, RH
        d.update(SM=SM, IM=IM, M_total=M_total, h_swe=h_swe, h_iwe=h_iwe, h_snow=self.h_snow,
                 h_ice=self.h_ice, Eccs=Eccs, Ecci=Ecci)
        self.step_index += 1
        err.__exit__(None, None, None)
        return d

    # -- Clear_Sky_Radiation with its redundant recomputation removed ------------------------------
    def _clear_sky(self, tr: TimeRow, W_p, th, albedo):
        """solar_funcs.py:894-953.

        Every callee recomputes Day_Angle / Declination / Optical_Air_Mass from the
        same arguments, so evaluating each distinct value once is bit-identical.
        """
        c = self.c
        gamma_dust = c.dust_atten
        delta, E0 = tr.delta, tr.E0
        lat_rad = self.lat_rad
        # Zenith_Angle :279-284
        term1 = np.sin(lat_rad) * np.sin(delta)
        term2 = np.cos(lat_rad) * np.cos(delta) * np.cos(OMEGA * th)
        Z = np.arccos(term1 + term2)
        # Optical_Air_Mass :540-568
        a, b, cc = 0.50572, 6.07995, 1.6364
        Z_deg = Z * (180 / np.pi)
        gamma = 90.0 - Z_deg
        gamma = np.where(0 > gamma, 0.0, gamma)  # python max(gamma, 0)
        t1 = np.sin(gamma * (np.pi / 180))
        t2 = a / self._pow_scalar_path(gamma + b, cc)
        M_opt = np.float64(1) / (t1 + t2)
        self.diag["M_opt"] = M_opt
        # Atmospheric_Transmissivity :608-614
        a_sa = -np.float64(0.1240) - (np.float64(0.0207) * W_p)
        b_sa = -np.float64(0.0682) - (np.float64(0.0248) * W_p)
        tau_sa = np.exp(a_sa + (b_sa * M_opt))
        tau = np.minimum(np.maximum(tau_sa - gamma_dust, 0), 1)
        # Scattering_Attenuation :649-653
        a_s = -np.float64(0.0363) - (np.float64(0.0084) * W_p)
        b_s = -np.float64(0.0572) - (np.float64(0.0173) * W_p)
        tau_s = np.exp(a_s + (b_s * M_opt))
        gam_s = (1 - tau_s) + gamma_dust
        # ET_Radiation_Flux :391-412
        h1 = np.cos(delta) * np.cos(lat_rad) * np.cos(OMEGA * th)
        h2 = np.sin(delta) * np.sin(lat_rad)
        K_h = np.maximum(I_SC * E0 * (h1 + h2), 0.0)
        # ET_Radiation_Flux_Slope :866-887
        s1 = np.cos(delta) * np.cos(self.lat_eq)
        s2 = np.cos((OMEGA * th) + self.dlon)
        s3 = np.sin(self.lat_eq) * np.sin(delta)
        K_s = np.maximum(I_SC * E0 * ((s1 * s2) + s3), 0)
        K_dif = np.float64(0.5) * gam_s * K_h  # :667
        K_dir = tau * K_h  # :634
        K_global = K_dir + K_dif  # :683
        K_bs = np.float64(0.5) * gam_s * albedo * K_global  # :711
        K_cs = (tau * K_s) + K_dif + K_bs  # :909
        self.diag.update(tau=tau, gam_s=gam_s, K_h=K_h, K_s=K_s)
        # sunrise / sunset on the slope :783-830 and horizontal :305-358
        arg_eq = np.minimum(np.maximum(-1, -np.float64(1) * np.tan(self.lat_eq_rt) * np.tan(delta)), 1)
        arg_h = np.minimum(np.maximum(-1, -np.float64(1) * np.tan(lat_rad) * np.tan(delta)), 1)
        T_sr = np.maximum((-np.float64(1) * np.arccos(arg_eq) / OMEGA) + self.t_noon,
                          -np.float64(1) * np.arccos(arg_h) / OMEGA)
        T_ss = np.minimum((np.arccos(arg_eq) / OMEGA) + self.t_noon, np.arccos(arg_h) / OMEGA)
        dark = np.logical_or(th <= T_sr, th >= T_ss)
        self.diag.update(T_sr=T_sr, T_ss=T_ss)
        return np.where(dark, np.float64(0), K_cs)

    # -- convenience ------------------------------------------------------------------------------
    def outputs(self) -> dict[str, np.ndarray]:
        return {k: getattr(self, k) for k in OUTPUT_VARS}

    def run(self, forcing: np.ndarray, record=OUTPUT_VARS) -> dict[str, np.ndarray]:
        """Advance ``T`` steps with ``forcing[T, 5, N]`` (FORCING_VARS order); returns ``{name: [T, N]}``."""
        T = forcing.shape[0]
        out = {k: np.empty((T, self.N)) for k in record}
        for t in range(T):
            dd = self.step(*forcing[t])
            for k in record:
                out[k][t] = dd[k] if k in dd else getattr(self, k)
        return out
