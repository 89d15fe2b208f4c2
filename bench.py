#!/usr/bin/env python
"""bench.py -- cell-timesteps/sec of the fused energy-balance + melt kernel on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode f64_fast|f64|f32] [--impl reference]

Workload (BASELINE.json configs[3], SURVEY.md 8d cfg 4): a synthetic 4096 x 4096 glacierised raster
(16 777 216 cells) PER GPU, hourly forcing, advanced in chunks of ``--chunk`` (default 128) timesteps.
One bench "step" = one launch of the fused kernel over one forcing chunk for every cell of the rank
(+ the all-reduce of the per-basin aggregates when N > 1).  Weak scaling: per-GPU work is fixed.

* ``value``      whole-job cell-steps/s with the forcing chunk resident in HBM (86 GB in float64: far larger than
                 L2, so no flush is needed between launches);
* ``roofline``   the melt kernel alone, CUDA events on the launching stream; algorithmic bytes = 5 live forcings x
                 element size per cell-step (40 B float64 / 20 B float32), peak = MEASURED_PEAKS.json hbm_gbs;
* ``e2e``        the same metric through the public API with HOST buffers: pinned raw met columns ->
                 ForcingStreamer (cudaMemcpyAsync on a side stream + unit-conversion kernel) -> tfg_run ->
                 device->host copy of the eight BMI outputs and the basin aggregates, all inside the timed region;
* ``cpu_baseline``  the NumPy oracle (a port of the reference: kind "port") on all host cores, on a bounded
                 sample of the same cells and forcing; the sample is also used to check the GPU result.

``--impl reference`` times that CPU port alone (the reference itself is a pure-Python package whose build
backend is not installable offline) and prints the line with "impl": "reference".
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "cell_timesteps_per_sec"
UNIT = "cell-steps/s"
GRID_CELLS = 4096 * 4096
REGIONAL_CELLS = 100_000_000
N_BASIN = 4096


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:  # noqa: BLE001
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- CPU arm
def _oracle_worker(args):
    import numpy as np

    from oracle.np_ref import CellStatics, Constants, OracleModel

    statics, forcing, start, reps = args
    cells = CellStatics(**statics, tz=[-8.0])
    model = OracleModel(cells, Constants(), start, strict_pow=False)
    t0 = time.perf_counter()
    out = None
    for _ in range(reps):
        for t in range(forcing.shape[0]):
            out = model.step(*forcing[t])
    dt = time.perf_counter() - t0
    return dt, {k: np.array(out[k]) for k in ("M_total", "h_swe", "h_iwe", "RH")}


def cpu_port_throughput(statics: dict, forcing, start: str, cores: int, reps: int = 1):
    """All host cores, one oracle instance per core on its slice of the sample cells."""
    import multiprocessing as mp

    import numpy as np

    n = statics["lat"].size
    bounds = np.linspace(0, n, cores + 1).astype(int)
    jobs = []
    for i in range(cores):
        lo, hi = bounds[i], bounds[i + 1]
        if hi > lo:
            jobs.append(({k: v[lo:hi] for k, v in statics.items()}, np.ascontiguousarray(forcing[:, :, lo:hi]), start, reps))
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        res = pool.map(_oracle_worker, jobs)
    wall = time.perf_counter() - t0
    busy = max(r[0] for r in res)
    last = {k: np.concatenate([r[1][k] for r in res]) for k in res[0][1]}
    return n * forcing.shape[0] * reps / busy, wall, last


def synthetic_host_sample(n_cells: int, n_steps: int, seed: int = 4096):
    """Host-only synthetic cells + forcing with the cfg-4 distributions (used when no GPU is involved)."""
    import numpy as np

    rng = np.random.Generator(np.random.PCG64(seed))
    U = lambda lo, hi: rng.uniform(lo, hi, n_cells)  # noqa: E731
    swe = U(0, 1.5)
    iwe = U(0, 60) * (rng.uniform(size=n_cells) < 0.4)
    statics = {"da": np.full(n_cells, 9e-4), "slope": U(0, 120), "aspect": U(0, 360), "lon": U(-122.1, -121.4),
               "lat": U(46.5, 47.1), "elev": U(1200, 4300), "h0_snow": swe * 20.0, "h0_ice": iwe * (1000.0 / 917.0),
               "h0_swe": swe, "h0_iwe": iwe, "T_rain_snow": np.zeros(n_cells)}
    hours = np.arange(n_steps)[:, None]
    T2D = (275.15 + 9 * np.sin(2 * np.pi * ((274 + hours // 24) % 365 - 105) / 365)
           + 4 * np.sin(2 * np.pi * (hours % 24 - 15) / 24) + rng.normal(0, 2, (n_steps, n_cells))
           - 6.5e-3 * (statics["elev"][None, :] - 2400.0))
    PSFC = 88900 + rng.normal(0, 400, (n_steps, n_cells))
    Tc = T2D - 273.15
    esat = 611.0 * np.exp(17.3 * Tc / (Tc + 237.3))
    q = np.clip(0.8 * (0.622 * esat / (PSFC - 0.378 * esat)) * rng.uniform(0.5, 1, (n_steps, n_cells)), 5e-4, 0.012)
    rain = np.where(rng.uniform(size=(n_steps, n_cells)) < 0.12, rng.exponential(0.5, (n_steps, n_cells)), 0.0)
    uz = 3.0 * np.hypot(rng.normal(size=(n_steps, n_cells)), rng.normal(size=(n_steps, n_cells)))
    forcing = np.stack([rain * 1e-3, T2D - 273.15, PSFC, q, uz], axis=1)
    return statics, np.ascontiguousarray(forcing)


def run_reference_arm(args):
    """The reference's CPU implementation of the path (NumPy port), all host cores, bounded sample per step."""
    rank, _, world = env_rank()
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_cells = min(args.cpu_cells or 131072 * cores, 1 << 21)
    n_steps = args.cpu_steps
    statics, forcing = synthetic_host_sample(n_cells, n_steps)
    times = []
    for i in range(args.warmup + args.steps):
        thr, wall, _ = cpu_port_throughput(statics, forcing, "2012100100", cores)
        if i >= args.warmup:
            times.append((thr, wall))
    cs = n_cells * n_steps
    # like the GPU arm (device time, max over ranks): the slowest worker's compute time per step, not the wall
    # clock around fork + pickling of the sample
    busy = sum(cs / thr for thr, _ in times)
    value = cs * len(times) / busy
    wall = busy
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic glacierised raster (cfg-4 distributions), CPU sample", "cells": n_cells,
                   "timesteps_per_step": n_steps},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_cells} cells x {n_steps} timesteps per step, oracle/np_ref.py, one process per core; "
                                   "timed as the slowest worker's compute time (pool start-up and pickling excluded)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------- GPU arm
def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from topoflow_glacier_b200 import _lib
    from topoflow_glacier_b200.engine import MeltEngine
    from topoflow_glacier_b200.forcing import ForcingStreamer
    from topoflow_glacier_b200.sharding import BasinAggregates
    from topoflow_glacier_b200.synthetic import synthetic_cells
    from topoflow_glacier_b200.config import default_constants

    rank, local_rank, world = env_rank()
    # ---- CPU baseline first (rank 0, N == 1), before CUDA is initialised so that fork() is safe ------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        ns = min(args.cpu_cells or 131072 * cores, 1 << 21)
        statics_h, forcing_h = synthetic_host_sample(ns, args.cpu_steps)
        thr, wall_cpu, last = cpu_port_throughput(statics_h, forcing_h, "2012100100", cores)
        cpu = {"value": thr, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{ns} cells (cfg-4 distributions) x {args.cpu_steps} timesteps, oracle/np_ref.py, one process "
                         f"per core ({wall_cpu:.1f} s wall)", "_sample": (statics_h, forcing_h, last)}
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the melt path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    Tc, mode = args.chunk, args.mode
    regional = args.workload == "regional"
    if regional:  # strong scaling: the grid is fixed, every rank owns a contiguous 128-aligned block of it
        from topoflow_glacier_b200.sharding import shard_bounds

        total_cells = args.cells
        lo, hi = shard_bounds(total_cells, world, rank)
        n_cells, first_cell = hi - lo, lo
    else:         # weak scaling: one raster per GPU
        n_cells, first_cell, total_cells = args.cells, rank * args.cells, args.cells * world
    es = 4 if mode == "f32" else 8
    consts = default_constants()
    tabs = synthetic_cells(n_cells, seed=(100 if regional else 4096) + rank, device=dev)
    raw_attrs = {"elev": tabs.pop("raw")["elev"]}
    per_basin = -(-total_cells // N_BASIN)  # 24 415 cells per basin on the regional grid (SURVEY.md 8d cfg 5)
    basin_id = ((torch.arange(n_cells, device=dev, dtype=torch.int64) + first_cell) // per_basin).to(torch.int32)
    horizon = (args.warmup + args.steps + 4) * Tc + (args.e2e_steps + args.warmup + 2) * args.e2e_chunk + 64
    eng = MeltEngine(None, consts, "2012100100", dt_hours=1, zones=[-8.0], basin_id=basin_id, n_basin=N_BASIN,
                     mode=mode, device=local_rank, horizon_steps=horizon, device_statics=tabs)
    elev = raw_attrs["elev"].to(eng.dtype)
    forcing = torch.empty(Tc, 5, n_cells, dtype=eng.dtype, device=dev)
    eng.synth_forcing(forcing, 0, Tc, elev, seed=20121001 + rank)
    # --agg exact: order-independent fixed-point accumulators + integer all-reduce (bit-identical for any N)
    agg_exps = eng.agg_exponents() if args.agg == "exact" else None
    agg = BasinAggregates(Tc, N_BASIN, device=dev, exponents=agg_exps)
    fp64_peak = eng.measure_fp64_peak()  # DFMA microbenchmark, before the timed region
    torch.cuda.synchronize()

    # ---- GPU result on the CPU sample (same host-generated cells and forcing): the correctness check -----
    if cpu is not None:
        statics_h, forcing_h, last = cpu.pop("_sample")
        nt = forcing_h.shape[0]
        chk = MeltEngine(statics_h, consts, "2012100100", zones=[-8.0], mode=mode, device=local_rank,
                         horizon_steps=nt + 1)
        chk.run(torch.as_tensor(forcing_h).to(dev, chk.dtype).contiguous(), nt)
        torch.cuda.synchronize()
        # cells that crossed a melt-out knife edge, or sit on the log-law singularity,
        # legitimately leave the oracle's trajectory; report the distribution instead of a single worst case
        g = {k: chk.row(k).to(torch.float64).cpu().numpy() for k in ("M_total", "h_swe", "h_iwe", "RH")}
        rel = np.zeros(g["RH"].shape)
        for k in g:
            rel = np.maximum(rel, np.abs(g[k] - last[k]) / (np.abs(last[k]) + 1e-12))
        cpu["gpu_vs_cpu_on_sample"] = {
            "cells": int(rel.size), "within_1e-12": int((rel <= 1e-12).sum()), "within_1e-9": int((rel <= 1e-9).sum()),
            "left_trajectory_at_knife_edge": int((rel > 1e-9).sum()), "median_rel_err": float(np.median(rel))}
        chk.close()

    def one_step():
        eng.run(forcing, Tc, basin_agg=agg.zero())
        agg.reduce()

    for _ in range(args.warmup):
        one_step()
    # ---- kernel-only timing (CUDA events around each launch, launching stream) -----------------------------
    evs = []
    sampler = ClockSampler(local_rank)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.start()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    start_all, end_all = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start_all.record()
    for _ in range(args.steps):
        tgt = agg.zero()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.run(forcing, Tc, basin_agg=tgt)
        b.record()
        agg.reduce()
        evs.append((a, b))
    end_all.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    dev_ms = start_all.elapsed_time(end_all)
    kern_ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    cell_steps = n_cells * Tc  # this rank's launch
    value = total_cells * Tc * args.steps / (dev_ms * 1e-3)

    # ---- same launch on spatially coherent weather (precipitation shared by 4096 consecutive cells) --------------
    # The headline above uses per-cell independent precipitation: the worst case for warp divergence
    # in the snowfall branch.  Real rasters see storms that cover whole warps; report that case beside it.
    coherent = None
    if not args.no_coherent:
        eng.synth_forcing(forcing, 0, Tc, elev, seed=20121001 + rank, storm_cells=4096)
        one_step()
        torch.cuda.synchronize()
        ev2 = []
        for _ in range(max(2, args.steps // 2)):
            tgt = agg.zero()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.run(forcing, Tc, basin_agg=tgt)
            b.record()
            ev2.append((a, b))
        torch.cuda.synchronize()
        ms2 = sum(a.elapsed_time(b) for a, b in ev2) / len(ev2)
        coherent = {"cell_steps_per_s_per_gpu": cell_steps / (ms2 * 1e-3), "kernel_ms": ms2,
                    "forcing": "as above, precipitation occurrence shared by 4096 consecutive cells"}

    # ---- end to end through the public API with host buffers -----------------------------------------------
    Te = args.e2e_chunk
    if regional:  # keep only the block the e2e leg streams; the 64 GB kernel-leg chunk goes back to the allocator
        forcing = forcing[:Te].clone()
        torch.cuda.empty_cache()
    raw_dtype = torch.float32 if args.e2e_raw == "float32" else torch.float64
    raw_host = torch.empty(Te, 6, n_cells, dtype=raw_dtype).pin_memory()
    # raw met columns derived from the synthetic chunk (mm/h, K, Pa, kg/kg, U, V)
    blk = forcing[:Te].to(torch.float64)
    raw_dev = torch.stack([blk[:, 0] * 1e3, blk[:, 1] + 273.15, blk[:, 2], blk[:, 3], blk[:, 4] * 0.6, blk[:, 4] * 0.8],
                          dim=1).to(raw_dtype)
    raw_host.copy_(raw_dev)
    del raw_dev, blk
    streamer = ForcingStreamer(eng, Te, raw_dtype=args.e2e_raw)
    out_host = [torch.empty(8, n_cells, dtype=eng.dtype).pin_memory() for _ in range(2)]
    agg_e = [BasinAggregates(Te, N_BASIN, device=dev, exponents=agg_exps) for _ in range(2)]
    agg_host = [torch.empty(Te, N_BASIN, 3, dtype=torch.float64).pin_memory() for _ in range(2)]
    out_rows = torch.tensor([0, 1, 8, 2, 3, 9, 10, 11], device=dev)  # h_snow,h_swe,SM,h_ice,h_iwe,IM,M_total,RH
    drain = torch.cuda.Stream(device=dev)
    drained = [torch.cuda.Event() for _ in range(2)]

    def e2e_run(k_steps):
        """k_steps host blocks streamed back to back: H2D of block i+1, the kernel of block i and the D2H of the
        results of block i-1 overlap (PCIe is full duplex); every block's outputs reach pinned host memory."""
        cur = torch.cuda.current_stream()
        for i, chunk in enumerate(streamer.chunks([raw_host] * k_steps)):
            j = i % 2
            cur.wait_event(drained[j])  # result buffers of two blocks ago have left the device
            eng.run(chunk, chunk.shape[0], basin_agg=agg_e[j].zero())
            agg_e[j].reduce()
            snap = eng.state.index_select(0, out_rows)
            done = torch.cuda.Event()
            done.record(cur)
            with torch.cuda.stream(drain):
                drain.wait_event(done)
                out_host[j].copy_(snap, non_blocking=True)
                agg_host[j].copy_(agg_e[j].buffer, non_blocking=True)
                snap.record_stream(drain)
                drained[j].record(drain)
        drain.synchronize()
        cur.synchronize()

    e2e_run(max(2, args.warmup // 2))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_run(args.e2e_steps)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_cells * Te * args.e2e_steps / float(t.item())
    h2d = raw_host.numel() * raw_host.element_size()
    d2h = out_host[0].numel() * out_host[0].element_size() + agg_host[0].numel() * 8

    # ---- the same end-to-end path with CATCHMENT forcing (tfg_bind_forcing_map) --------------------------------
    # One forcing series per basin (4096 columns) instead of one per 30 m cell, as the reference's one-CSV-per-
    # catchment drivers supply it: the host block shrinks from 24 B to 0.006 B per cell-step and PCIe no longer
    # binds.  Reported beside `e2e`, which stays the per-cell-forcing number.
    shared = None
    if not regional and not args.no_shared:
        del raw_host, streamer
        Ts = Tc
        eng.set_forcing_map(basin_id % N_BASIN, N_BASIN)
        raw_cols = torch.empty(Ts, 6, N_BASIN, dtype=raw_dtype).pin_memory()
        blk = forcing[:Ts, :, :N_BASIN].to(torch.float64)
        raw_cols.copy_(torch.stack([blk[:, 0] * 1e3, blk[:, 1] + 273.15, blk[:, 2], blk[:, 3], blk[:, 4] * 0.6,
                                    blk[:, 4] * 0.8], dim=1).to(raw_dtype))
        streamer2 = ForcingStreamer(eng, Ts, raw_dtype=args.e2e_raw)
        agg_s = [BasinAggregates(Ts, N_BASIN, device=dev, exponents=agg_exps) for _ in range(2)]
        agg_sh = [torch.empty(Ts, N_BASIN, 3, dtype=torch.float64).pin_memory() for _ in range(2)]

        def shared_run(k_steps):
            cur = torch.cuda.current_stream()
            for i, chunk in enumerate(streamer2.chunks([raw_cols] * k_steps)):
                j = i % 2
                cur.wait_event(drained[j])
                eng.run(chunk, chunk.shape[0], basin_agg=agg_s[j].zero())
                agg_s[j].reduce()
                snap = eng.state.index_select(0, out_rows)
                done = torch.cuda.Event()
                done.record(cur)
                with torch.cuda.stream(drain):
                    drain.wait_event(done)
                    out_host[j].copy_(snap, non_blocking=True)
                    agg_sh[j].copy_(agg_s[j].buffer, non_blocking=True)
                    snap.record_stream(drain)
                    drained[j].record(drain)
            drain.synchronize()
            cur.synchronize()

        shared_run(2)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        shared_run(args.e2e_steps)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        shared = {"value": total_cells * Ts * args.e2e_steps / float(t.item()), "unit": UNIT,
                  "h2d_bytes_per_step": raw_cols.numel() * raw_cols.element_size(),
                  "d2h_bytes_per_step": out_host[0].numel() * out_host[0].element_size() + agg_sh[0].numel() * 8,
                  "timesteps_per_step": Ts, "forcing_columns": N_BASIN,
                  "path": "as e2e, one forcing series per basin (tfg_bind_forcing_map) instead of one per cell"}

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = es * _lib.N_FORCING * cell_steps / (kern_ms * 1e-3) / 1e9
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                traffic = json.loads(tp.read_text()).get(f"{mode}:{n_cells}x{Tc}")
            except Exception:  # noqa: BLE001
                traffic = None
        # the unit that actually binds: FP64 pipe at 1 warp-instruction / 2 cycles / scheduler
        compute = None
        mp = ROOT / "profiles" / "kernel_mix.json"
        if mp.exists() and mode != "f32":
            try:
                mix = json.loads(mp.read_text())[mode]
                # ceiling = measured DFMA thread-ops/s / FP64 instructions per cell-step (a warp-instruction is 32 of them)
                ceiling = fp64_peak / mix["fp64_warp_inst_per_warp_step"]
                # issue ceiling: one warp-instruction per cycle and scheduler (4 per SM)
                sms = torch.cuda.get_device_properties(dev).multi_processor_count
                mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
                issue_ceiling = sms * 4 * mhz * 1e6 * 32 / mix["warp_inst_per_warp_step"]
                compute = {"bound": "fp64_pipe", "fp64_warp_inst_per_warp_step": mix["fp64_warp_inst_per_warp_step"],
                           "warp_inst_per_warp_step": mix["warp_inst_per_warp_step"],
                           "issue_ceiling_cell_steps_per_s": issue_ceiling,
                           "frac_of_issue_ceiling": (cell_steps / (kern_ms * 1e-3)) / issue_ceiling,
                           "fp64_peak_tflops_measured": 2 * fp64_peak / 1e12,
                           "ceiling_cell_steps_per_s": ceiling, "frac_of_ceiling": (cell_steps / (kern_ms * 1e-3)) / ceiling,
                           "ncu_fp64_pipe_active_pct": mix["fp64_pipe_active_pct"],
                           "ncu_issue_active_pct": mix["issue_active_pct"], "source": mix["source"]}
            except Exception:  # noqa: BLE001
                compute = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong" if regional else "weak",
            "vs_baseline": None,
            "dtype": "f32" if mode == "f32" else "f64", "data": "synthetic",
            "config": {"workload": ("synthetic 100M-cell regional grid sharded over the GPUs (BASELINE configs[4]), hourly forcing"
                                    if regional else
                                    "synthetic 4096x4096 glacierised raster per GPU (BASELINE configs[3]), hourly forcing"),
                       "cells_total": total_cells, "cells_per_gpu": n_cells, "timesteps_per_step": Tc, "arithmetic_mode": mode,
                       "basin_aggregates": N_BASIN, "aggregate_sums": args.agg, "forcing": "device-resident chunk, Philox synthetic, reused each step",
                       "l2": f"inputs {es * 5 * cell_steps / 1e9:.1f} GB per launch >> 126 MB L2 (no flush needed)",
                       "parallelism": f"cells sharded x{world}, all_reduce of basin aggregates"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "tfg::run_kernel",
                         "kernel_ms": kern_ms, "algorithmic_bytes_per_cell_step": es * _lib.N_FORCING,
                         "note": "FP64-pipe / issue bound, not HBM bound: see compute_roofline, profiles/ and DESIGN.md"},
            "compute_roofline": compute,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "timesteps_per_step": Te, "raw_dtype": args.e2e_raw,
                    "path": "pinned host raw met -> ForcingStreamer (tfg_ingest_async + tfg_convert_forcing) -> tfg_run "
                            "-> D2H of 8 BMI outputs + basin aggregates"},
            "e2e_catchment_forcing": shared,
            "gpu_launches": args.steps, "clocks": clocks, "wall_s": wall, "coherent_weather": coherent,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="f64_fast", choices=["f64", "f64_fast", "f32"])
    ap.add_argument("--workload", default="raster", choices=["raster", "regional"],
                    help="raster: 4096x4096 cells PER GPU (BASELINE configs[3], weak scaling); regional: 100 M cells in "
                         "total, sharded over the GPUs (configs[4], strong scaling, 16 timesteps per launch)")
    ap.add_argument("--cells", type=int, default=0, help="cells per GPU (raster) / in total (regional)")
    ap.add_argument("--chunk", type=int, default=0, help="timesteps per launch (default 128 raster, 16 regional)")
    ap.add_argument("--e2e-chunk", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--e2e-raw", default="float32", choices=["float32", "float64"])
    ap.add_argument("--cpu-cells", type=int, default=0)
    ap.add_argument("--cpu-steps", type=int, default=24)
    ap.add_argument("--agg", default="float", choices=["float", "exact"],
                    help="basin sums: float64 atomics, or order-independent fixed-point accumulators (TFG_OPT_EXACT_AGG)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-coherent", action="store_true")
    ap.add_argument("--no-shared", action="store_true", help="skip the catchment-forcing end-to-end leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    regional = args.workload == "regional"
    args.cells = args.cells or (REGIONAL_CELLS if regional else GRID_CELLS)
    args.chunk = args.chunk or (16 if regional else 128)
    if regional:  # 100 M cells fill the HBM of one GPU: no second forcing realisation, one-timestep e2e blocks
        args.no_coherent, args.no_cpu, args.e2e_chunk = True, True, 1
    return run_reference_arm(args) if args.impl == "reference" else run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
