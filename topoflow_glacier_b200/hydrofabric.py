"""NextGen hydrofabric (GeoPackage) -> ensemble of cells + basin topology.   [SURVEY.md 8f, rank 1]

The reference ships ``data/12082500.gpkg`` (43 divides upstream of USGS gage 12082500) but no code reads it;
``config/*.yaml:da`` simply repeats ``divides.areasqkm``.  Here the GeoPackage (plain SQLite, read with the
standard library) supplies what the aggregate outputs need:

* per-divide drainage area (``divides.areasqkm`` -> ``da``),
* the flow topology ``divide -> nexus -> waterbody -> nexus ...`` (``divides.toid``, ``network.id/toid``),
* from it a ``basin_id`` per cell: the index of the chosen outlet (nexus / gage) the cell drains to.
"""

from __future__ import annotations

import sqlite3
from dataclasses import dataclass
from pathlib import Path
from typing import Iterable, Optional, Sequence

import numpy as np

__all__ = ["Hydrofabric", "read_hydrofabric"]


@dataclass
class Hydrofabric:
    divide_id: list            # 'cat-…'
    toid: list                 # nexus each divide drains to
    areasqkm: np.ndarray
    tot_drainage_areasqkm: np.ndarray
    downstream: dict           # any id ('cat-', 'wb-', 'nex-') -> id it flows to

    def area_of(self, divide_ids: Iterable[str]) -> np.ndarray:
        idx = {d: i for i, d in enumerate(self.divide_id)}
        return np.array([self.areasqkm[idx[d]] for d in divide_ids], dtype=np.float64)

    def path_to_outlet(self, start: str) -> list:
        """Ids visited when following the flow from ``start`` to the terminal node."""
        seen, cur = [start], start
        while cur in self.downstream and self.downstream[cur] not in seen:
            cur = self.downstream[cur]
            if cur is None:
                break
            seen.append(cur)
        return seen

    def terminal(self, start: str) -> str:
        return self.path_to_outlet(start)[-1]

    def basin_ids(self, divide_ids: Sequence[str], outlets: Optional[Sequence[str]] = None):
        """``(basin_id[int32 N], outlet_names)``: for every cell the index of the FIRST listed outlet on its
        flow path (most upstream match wins); cells reaching none get the terminal node of their path, appended
        to the outlet list.  ``outlets=None`` groups by terminal node."""
        names = list(outlets or [])
        out = np.empty(len(divide_ids), dtype=np.int32)
        for i, d in enumerate(divide_ids):
            path = self.path_to_outlet(d)
            hit = next((p for p in path if p in names), None) if outlets else None
            if hit is None:
                hit = path[-1]
                if hit not in names:
                    names.append(hit)
            out[i] = names.index(hit)
        return out, names

    def upstream_divides(self, outlet: str) -> list:
        return [d for d in self.divide_id if outlet in self.path_to_outlet(d)]


def read_hydrofabric(path) -> Hydrofabric:
    p = Path(path)
    if not p.exists():
        raise FileNotFoundError(p)
    con = sqlite3.connect(f"file:{p}?mode=ro&immutable=1", uri=True)
    try:
        div = con.execute("select divide_id, toid, areasqkm, tot_drainage_areasqkm from divides order by fid").fetchall()
        net = con.execute("select id, toid, divide_id from network").fetchall()
    finally:
        con.close()
    down = {}
    for did, toid, _, _ in div:
        down[did] = toid
    for wid, toid, did in net:
        if wid is not None and toid is not None:
            down.setdefault(wid, toid)
    # a nexus flows into the waterbody that carries the same numeric suffix, when the network has one
    wb_ids = {wid for wid, _, _ in net if wid}
    for _, toid, _ in net:
        if toid and toid.startswith("nex-") and toid not in down:
            wb = "wb-" + toid.split("-", 1)[1]
            if wb in wb_ids:
                down[toid] = wb
    for _, toid, _, _ in div:
        if toid and toid.startswith("nex-") and toid not in down:
            wb = "wb-" + toid.split("-", 1)[1]
            if wb in wb_ids:
                down[toid] = wb
    return Hydrofabric([d[0] for d in div], [d[1] for d in div], np.array([d[2] for d in div], dtype=np.float64),
                       np.array([d[3] if d[3] is not None else np.nan for d in div], dtype=np.float64), down)
