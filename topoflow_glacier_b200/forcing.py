"""Forcing ingestion: per-catchment met series -> pinned host blocks -> SoA device chunks.

In the reference the driver script owns this step: it reads one CSV per catchment, converts units and
calls ``set_value`` seven times per timestep (reference ``examples/run_topoflow_glacier.py:30-73``,
``tests/integration_test.py:81-116``).  Here whole blocks of timesteps move at once:

    CSV (header-keyed)  ->  raw block [T, 6, N] in pinned memory
                        ->  cudaMemcpyAsync on a side stream            (tfg_ingest_async)
                        ->  unit conversion into [T, 5, N] on the device  (tfg_convert_forcing)
                        ->  event hand-off to the compute stream          (tfg_run consumes it)

with two buffers so that the copy of chunk k+1 overlaps the melt kernel of chunk k.
Only five of the seven BMI inputs are live (SW_in / LW_in are never read by the reference physics,
``bmi_topoflow_glacier.py:1122-1139``, ``:1234-1235``), so only the columns they derive from are moved.
"""

from __future__ import annotations

from pathlib import Path
from typing import Iterator, Optional, Sequence

import numpy as np
import pandas as pd

__all__ = ["RAW_COLUMNS", "read_forcing_csv", "stack_catchments", "convert_on_host", "ForcingStreamer",
           "DEFAULT_PACKING", "pack_forcing", "unpack_forcing", "bind_host_to_gpu", "read_forcing_netcdf",
           "write_forcing_netcdf"]

# raw met columns moved to the device, in kernel order (tfg_convert_forcing)
RAW_COLUMNS = ("RAINRATE", "T2D", "PSFC", "Q2D", "U2D", "V2D")


def read_forcing_csv(path, start: Optional[pd.Timestamp] = None, end: Optional[pd.Timestamp] = None) -> np.ndarray:
    """``[T, 6]`` float64 raw columns of one catchment, rows with ``start <= Time <= end``.

    Columns are located by header name: the reference's two sample files order them differently
    (``tests/data/sample-cat-3062920.csv:1`` vs ``tests/conftest.py:10``).
    """
    df = pd.read_csv(Path(path))
    missing = [c for c in ("Time",) + RAW_COLUMNS if c not in df.columns]
    if missing:
        raise KeyError(f"{path}: missing forcing columns {missing}")
    when = pd.to_datetime(df["Time"])
    keep = np.ones(len(df), dtype=bool)
    if start is not None:
        keep &= (when >= start).values
    if end is not None:
        keep &= (when <= end).values
    return np.ascontiguousarray(df.loc[keep, list(RAW_COLUMNS)].to_numpy(dtype=np.float64))


def stack_catchments(series: Sequence[np.ndarray]) -> np.ndarray:
    """``N`` arrays ``[T, 6]`` -> one block ``[T, 6, N]`` (cell index fastest), truncated to the shortest."""
    T = min(s.shape[0] for s in series)
    return np.ascontiguousarray(np.stack([s[:T] for s in series], axis=2))


def convert_on_host(raw: np.ndarray) -> np.ndarray:
    """Host statement of the unit conversions (``[T, 6, N]`` -> ``[T, 5, N]``); the device kernel must equal it.

    ``P = RAINRATE * 10**-3`` [m/h], ``T_air = -273.15 + T2D`` [degC], ``uz = (U2D**2 + V2D**2) ** 0.5``
    (``examples/run_topoflow_glacier.py:47-49,66-73``).
    """
    rain, t2d, psfc, q2d, u, v = (raw[:, i] for i in range(6))
    return np.ascontiguousarray(np.stack([rain * 10 ** (-3), -273.15 + t2d, psfc, q2d, (u**2 + v**2) ** 0.5], axis=1))


# NetCDF-style packing of the six raw columns (value = int16 * scale + offset), wide enough for any met record:
# RAINRATE 0..163 mm/h in 0.005 steps, T2D 109..437 K in 0.005 K, PSFC 34 464..165 536 Pa in 2 Pa, Q2D 0..0.065 in
# 1e-6, U2D / V2D +-163 m/s in 0.005.  (scale, offset) per column of RAW_COLUMNS.
DEFAULT_PACKING = (np.array([0.005, 0.005, 2.0, 1e-6, 0.005, 0.005]), np.array([163.84, 273.15, 100000.0, 0.0325, 0.0, 0.0]))


def pack_forcing(raw: np.ndarray, packing=DEFAULT_PACKING) -> np.ndarray:
    """``[T, 6, N]`` float raw columns -> int16 (round to nearest, saturating)."""
    scale, offset = (np.asarray(a, dtype=np.float64).reshape(1, 6, 1) for a in packing)
    q = np.rint((np.asarray(raw, dtype=np.float64) - offset) / scale)
    return np.ascontiguousarray(np.clip(q, -32768, 32767).astype(np.int16))


def unpack_forcing(packed: np.ndarray, packing=DEFAULT_PACKING) -> np.ndarray:
    """Host statement of what the device computes from packed columns: ``int16 * scale + offset`` in float64."""
    scale, offset = (np.asarray(a, dtype=np.float64).reshape(1, 6, 1) for a in packing)
    return packed.astype(np.float64) * scale + offset


def read_forcing_netcdf(path, start: Optional[pd.Timestamp] = None, end: Optional[pd.Timestamp] = None, packed: bool = True):
    """One NetCDF (classic / 64-bit offset) forcing file -> the block contract of ``ForcingStreamer``.

    Layout expected (what NWM / AORC style per-catchment archives hold): a ``time`` variable (seconds since the epoch, or
    with a CF ``units = "<seconds|minutes|hours> since YYYY-MM-DD HH:MM:SS"`` attribute) and the six ``RAW_COLUMNS`` as
    variables ``[time, cell]`` (or ``[time]`` for one catchment), stored either as floats or as int16 with the CF
    attributes ``scale_factor`` / ``add_offset``.

    ``packed=True`` and all six variables int16: returns ``(int16 [T, 6, N], (scale[6], offset[6]))`` -- the file's own
    packing goes to the device unchanged (12 B per cell-step over PCIe, ``ForcingStreamer(raw_dtype="int16",
    packing=...)``).  Otherwise returns ``(float64 [T, 6, N], None)`` with the packing applied on the host.
    Read with ``scipy.io.netcdf_file`` (NetCDF-4 / HDF5 files need a converter; no HDF5 library is assumed)."""
    from scipy.io import netcdf_file

    with netcdf_file(str(path), "r", mmap=False) as nc:
        missing = [c for c in ("time",) + RAW_COLUMNS if c not in nc.variables]
        if missing:
            raise KeyError(f"{path}: missing forcing variables {missing}")
        tv = nc.variables["time"]
        units = getattr(tv, "units", b"seconds since 1970-01-01 00:00:00")
        units = units.decode() if isinstance(units, bytes) else str(units)
        step, _, origin = units.partition(" since ")
        when = pd.Timestamp(origin.strip() or "1970-01-01") + pd.to_timedelta(np.asarray(tv[:], dtype=np.float64),
                                                                            unit={"seconds": "s", "minutes": "min", "hours": "h"}
                                                                            .get(step.strip().lower(), "s"))
        keep = np.ones(len(when), dtype=bool)
        if start is not None:
            keep &= np.asarray(when >= start)
        if end is not None:
            keep &= np.asarray(when <= end)
        cols, scale, offset, all_i16 = [], [], [], True
        for name in RAW_COLUMNS:
            v = nc.variables[name]
            a = np.array(v[:])
            a = a.reshape(a.shape[0], -1)[keep]
            if a.dtype.kind == "i" and a.dtype.itemsize == 2:
                a = a.astype(np.int16)          # NetCDF stores big-endian; the device wants native int16
            scale.append(float(getattr(v, "scale_factor", 1.0)))
            offset.append(float(getattr(v, "add_offset", 0.0)))
            cols.append(a)
            all_i16 &= a.dtype == np.int16
    if packed and all_i16:
        return np.ascontiguousarray(np.stack(cols, axis=1)), (np.array(scale), np.array(offset))
    out = np.stack([c.astype(np.float64) * s + o if (s != 1.0 or o != 0.0) else c.astype(np.float64)
                    for c, s, o in zip(cols, scale, offset)], axis=1)
    return np.ascontiguousarray(out), None


def write_forcing_netcdf(path, when: pd.DatetimeIndex, raw: np.ndarray, packing=DEFAULT_PACKING) -> None:
    """Write ``raw`` ([T, 6, N] float) as a packed NetCDF-3 forcing file (int16 + scale_factor / add_offset) -- the
    counterpart of ``read_forcing_netcdf``, used by the tests and by drivers that convert CSV archives once."""
    from scipy.io import netcdf_file

    packed = pack_forcing(raw, packing)
    T, _, N = packed.shape
    with netcdf_file(str(path), "w", version=2) as nc:
        nc.createDimension("time", T)
        nc.createDimension("cell", N)
        tv = nc.createVariable("time", "f8", ("time",))
        tv.units = "seconds since 1970-01-01 00:00:00"
        tv[:] = (pd.DatetimeIndex(when) - pd.Timestamp("1970-01-01")) / pd.Timedelta(seconds=1)
        for j, name in enumerate(RAW_COLUMNS):
            v = nc.createVariable(name, "i2", ("time", "cell"))
            v.scale_factor = np.float64(packing[0][j])    # (a Python float would be stored as float32)
            v.add_offset = np.float64(packing[1][j])
            v[:] = packed[:, j, :]


def pinned_block(shape, dtype, write_combined: bool = True):
    """A page-locked host tensor for ``ForcingStreamer`` sources (``tfg_host_alloc``).  ``write_combined`` (default)
    asks for write-combined pages: the host only FILLS such a block (reads from it are slow), and the device's reads
    over PCIe then need no cache snoop.  The memory is released when the tensor is garbage-collected."""
    import ctypes as C
    import weakref

    import torch

    from . import _lib

    lib = _lib.load()
    dtype = {np.dtype("int16"): torch.int16, np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64}.get(
        np.dtype(dtype) if not isinstance(dtype, torch.dtype) else None, dtype)
    n = int(np.prod(shape))
    nbytes = n * torch.empty((), dtype=dtype).element_size()
    ptr = C.c_void_p()
    _lib.check(lib.tfg_host_alloc(C.byref(ptr), nbytes, int(bool(write_combined))), "tfg_host_alloc")
    buf = (C.c_char * nbytes).from_address(ptr.value)
    t = torch.frombuffer(buf, dtype=dtype, count=n).view(*shape)
    weakref.finalize(buf, lib.tfg_host_free, C.c_void_p(ptr.value))   # the tensor keeps `buf` alive
    return t


def bind_host_to_gpu(device_index: int) -> bool:
    """Pin the calling process to the CPU cores next to GPU ``device_index`` (NVML's ideal affinity), so that pinned
    staging buffers allocated afterwards are first-touched on that GPU's NUMA node.  With several ranks streaming
    forcing at once the host->device copies otherwise cross the socket interconnect.  Returns False when NVML (or
    the permission to change the affinity) is missing."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:  # noqa: BLE001
        return False


class ForcingStreamer:
    """Double-buffered host -> device forcing pipeline for one ``MeltEngine``.

    ``raw_dtype``: ``"float64"`` / ``"float32"`` raw columns, or ``"int16"`` for packed columns (``packing=(scale,
    offset)``, default ``DEFAULT_PACKING``; see ``pack_forcing``) -- 12 instead of 24 / 48 bytes per cell-step over PCIe."""

    def __init__(self, engine, chunk_steps: int, n_buffers: int = 2, raw_dtype="float64", packing=None):
        import torch

        from . import _lib

        self.torch, self._lib = torch, _lib
        self.e = engine
        self.Tc = int(chunk_steps)
        self.nb = int(n_buffers)
        if self.nb < 2:  # piece k+1 is uploaded while piece k is being consumed: one buffer would be overwritten
            raise ValueError("ForcingStreamer needs n_buffers >= 2")
        N, dev = engine.n_cols, engine.device   # forcing columns (= cells unless the engine has a forcing map)
        self.N = N
        self.side = torch.cuda.Stream(device=dev)
        name = str(raw_dtype)
        self.raw_dtype = torch.int16 if name.endswith("int16") else torch.float32 if name.endswith("32") else torch.float64
        self.raw_es = {torch.int16: 2, torch.float32: 4, torch.float64: 8}[self.raw_dtype]
        self.packing = None
        if self.raw_dtype == torch.int16:
            sc, of = packing if packing is not None else DEFAULT_PACKING
            self.packing = (np.ascontiguousarray(sc, dtype=np.float64), np.ascontiguousarray(of, dtype=np.float64))
            if self.packing[0].shape != (6,) or self.packing[1].shape != (6,):
                raise ValueError("packing must be (scale[6], offset[6])")
        self.pinned = [None] * self.nb  # staging for NumPy sources, allocated on first use
        self.d_raw = [torch.empty(self.Tc, 6, N, dtype=self.raw_dtype, device=dev) for _ in range(self.nb)]
        self.d_out = [torch.empty(self.Tc, 5, N, dtype=engine.dtype, device=dev) for _ in range(self.nb)]
        self.h2d_done = [torch.cuda.Event() for _ in range(self.nb)]
        self.ready = [torch.cuda.Event() for _ in range(self.nb)]
        self.consumed = [torch.cuda.Event() for _ in range(self.nb)]
        self.h2d_bytes = 0

    def _submit(self, b: int, block) -> int:
        """Stage ``block`` ([Tk, 6, N] host) in buffer ``b``: (pinned copy,) async H2D, device conversion.

        A ``torch`` tensor that already lives in pinned memory is copied from where it is; a NumPy
        array is first staged in this streamer's own pinned buffer.
        """
        torch, lib, e = self.torch, self.e.lib, self.e
        Tk = block.shape[0]
        # page-locked sources are copied from where they lie (torch's own pinned tensors, or any cudaHostAlloc'ed block such
        # as forcing.pinned_block: torch only recognises the former, the library asks the driver)
        if (torch.is_tensor(block) and block.device.type == "cpu" and block.is_contiguous() and block.dtype == self.raw_dtype
                and (block.is_pinned() or lib.tfg_host_is_pinned(block.data_ptr()) == 1)):
            src_ptr = block.data_ptr()
            self._keepalive = block
        else:
            if self.pinned[b] is None:
                self.pinned[b] = torch.empty(self.Tc, 6, self.N, dtype=self.raw_dtype).pin_memory()
            self.h2d_done[b].synchronize()  # pinned buffer free again
            self.pinned[b][:Tk].numpy()[...] = block.numpy() if torch.is_tensor(block) else block
            src_ptr = self.pinned[b].data_ptr()
        self.side.wait_event(self.consumed[b])  # device buffers free again
        nbytes = Tk * 6 * self.N * self.raw_es
        with torch.cuda.device(e.device):
            self._lib.check(lib.tfg_ingest_async(e.ctx, src_ptr, self.d_raw[b].data_ptr(), nbytes,
                                                 self.side.cuda_stream, None), "tfg_ingest_async")
            self.h2d_done[b].record(self.side)
            if self.packing is not None:
                self._lib.check(lib.tfg_convert_forcing_packed(e.ctx, self.d_raw[b].data_ptr(), self.packing[0].ctypes.data,
                                                               self.packing[1].ctypes.data, self.d_out[b].data_ptr(), Tk, self.N,
                                                               self.side.cuda_stream), "tfg_convert_forcing_packed")
            else:
                self._lib.check(lib.tfg_convert_forcing(e.ctx, self.d_raw[b].data_ptr(), self.raw_es, self.d_out[b].data_ptr(), Tk,
                                                        self.N, self.side.cuda_stream), "tfg_convert_forcing")
            self.ready[b].record(self.side)
        self.h2d_bytes += nbytes
        return Tk

    def chunks(self, raw) -> Iterator:
        """Yield device forcing chunks ``[Tk, 5, N]`` in order; the consumer must launch its kernel on the
        current stream before asking for the next chunk.

        ``raw`` is one host block ``[T, 6, N]`` (cut into pieces of ``chunk_steps``) or a sequence of host blocks
        (each at most ``chunk_steps`` long) -- e.g. successive hours arriving from a met feed.  The copy of
        piece k+1 is in flight while piece k is being computed.
        """
        torch = self.torch
        if isinstance(raw, (list, tuple)):
            pieces = list(raw)
        else:
            pieces = [raw[s:s + self.Tc] for s in range(0, raw.shape[0], self.Tc)]
        if not pieces:
            return
        sizes = {0: self._submit(0, pieces[0])}
        for k in range(len(pieces)):
            b = k % self.nb
            if k + 1 < len(pieces):  # prefetch the next piece while this one computes
                sizes[k + 1] = self._submit((k + 1) % self.nb, pieces[k + 1])
            cur = torch.cuda.current_stream(self.e.device)
            cur.wait_event(self.ready[b])
            yield self.d_out[b][:sizes[k]]
            self.consumed[b].record(cur)

    def drive(self, raw: np.ndarray, **run_kw) -> None:
        """Stream ``raw`` ([T, 6, N] host) through the engine."""
        for chunk in self.chunks(raw):
            self.e.run(chunk, chunk.shape[0], **run_kw)
