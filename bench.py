#!/usr/bin/env python
"""bench.py -- D8 flow direction + flow accumulation throughput on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--size S]

One step = one pass of the hot path (flow direction, then flow accumulation) over one synthetic
float32 DEM.  N=1: the S x S raster (default 65536, BASELINE.json configs[2]) is resident in HBM
and processed by one GPU.  N>1 (launched by torchrun, one rank per GPU): the same raster split into
N row strips (strong scaling), halo rows and the strip-boundary graph exchanged over NCCL.
`--impl reference` times the CPU oracle port (oracle/d8_oracle.c, the reference's algorithm in C)
on the host cores on a bounded sample of the same DEM.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NODATA = -9999.0
METRIC = "d8_flowdir_flowacc_throughput"
UNIT = "Gcells/s"
DIR_BYTES_PER_CELL = 5.0   # 4 B float32 read + 1 B code write            
ACC_BYTES_PER_CELL = 9.0   # 1 B code read + 8 B int64 count write         
# accumulation: 9 B/cell = 1 B code read (pass A) + 8 B count write (final pass)
PHASE_BYTES = {"direction": 5.0, "acc_tile_a": 1.0, "acc_tile_b": 8.0}


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_pass(oracle, dem_pad):
    """Reference algorithm on the host: direction on all threads (prange), accumulation serial."""
    t0 = time.perf_counter()
    fdr = oracle.flow_direction_for_tile(dem_pad, NODATA)[1:-1, 1:-1]
    t1 = time.perf_counter()
    fac = oracle.flow_accumulation(fdr)
    t2 = time.perf_counter()
    return fdr, fac, t1 - t0, t2 - t1


def sample_dem_host(args, sample):
    """`sample` x `sample` window of the benchmark DEM (+ nodata ring), generated on the GPU when there
    is one (same generator, same seed as the native arm), else the numpy fractal."""
    import numpy as np

    try:
        import torch

        if torch.cuda.is_available():
            from overflow_b200 import device as dev

            d = dev.synth_dem(sample, sample, row0=0, total_rows=args.size, seed=args.seed, kind=args.kind,
                              holes_permille=args.holes)
            h = d.cpu().numpy()
            del d
            out = np.full((sample + 2, sample + 2), NODATA, dtype=np.float32)
            out[1:-1, 1:-1] = h
            return out, "device generator window"
    except Exception:
        pass
    from oracle import synth

    return synth.pad_nodata(synth.fractal(sample, sample, beta=2.0, seed=args.seed)), "numpy fractal (no GPU)"


def run_reference(args):
    """--impl reference: the CPU port of the reference's path on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle

    oracle.build()
    # every host thread this process may use, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for N > 1)
    try:
        oracle.set_num_threads(len(os.sched_getaffinity(0)))
    except AttributeError:
        oracle.set_num_threads(os.cpu_count() or 1)
    sample = min(args.size, args.cpu_sample)
    dem_pad, how = sample_dem_host(args, sample)
    cells = sample * sample
    for _ in range(args.warmup):
        cpu_pass(oracle, dem_pad)
    t0 = time.perf_counter()
    td = ta = 0.0
    for _ in range(args.steps):
        _, _, a, b = cpu_pass(oracle, dem_pad)
        td += a
        ta += b
    el = time.perf_counter() - t0
    value = cells * args.steps / el / 1e9
    sample_txt = (f"{sample}x{sample} window of the {args.size}x{args.size} DEM ({how}); direction on "
                  f"{oracle.num_threads()} threads {td / args.steps * 1e3:.0f} ms, accumulation serial "
                  f"{ta / args.steps * 1e3:.0f} ms per step")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32->u8->int64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port", "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n):
    kinds = {0: "fractal value-noise", 1: "terraced fractal", 2: "tilted plane"}
    return {
        "workload": f"flow_direction + flow_accumulation on synthetic {args.size}x{args.size} float32 DEM "
                    f"({kinds[args.kind]}, {args.holes} permille nodata holes, seed {args.seed})",
        "rows": args.size, "cols": args.size, "partition": f"{n} row strip(s)",
        "cache": "inputs (4 B/cell DEM) far exceed the 126 MB L2, no flush needed",
    }


def run_native(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from overflow_b200 import _native, device as dev

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one rank per GPU)")
    torch.cuda.set_device(local)
    _native.init(local)
    if world > 1:
        # NCCL announces its version on stdout when the first communicator comes up; stdout carries exactly
        # one JSON line, so that goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    S = args.size
    peak, peak_src = measured_hbm_peak()

    if world > 1:
        from overflow_b200 import strips

        runner = strips.StripPipeline(S, S, rank, world, nodata=NODATA, device=torch.device("cuda", local))
        runner.load_synthetic(seed=args.seed, kind=args.kind, holes_permille=args.holes)
        step = runner.step
        cells_total = S * S
    else:
        dem = dev.synth_dem(S, S, seed=args.seed, kind=args.kind, holes_permille=args.holes, nodata=NODATA)
        fdr = torch.empty((S, S), dtype=torch.uint8, device="cuda")
        fac = torch.empty((S, S), dtype=torch.int64, device="cuda")

        def step():
            dev.flow_routing(dem, NODATA, out_fdr=fdr, out_fac=fac)  # one C-ABI call: direction, then accumulation

        cells_total = S * S

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    _native.phase_timing_read(reset=True)
    _native.phase_timing_enable(True)
    _native.launch_count_reset()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _native.launch_count()
    _native.phase_timing_enable(False)
    phases = _native.phase_timing_read(reset=True)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_per_step = ms / args.steps
    value = cells_total / (ms_per_step * 1e-3) / 1e9

    # ---- roofline (live CUDA-event times from the timed region, this rank).  SURVEY 8(d) states the algorithmic
    #      bytes per stage: direction 5 B/cell, accumulation 9 B/cell (1 B code read + 8 B count write).  The
    #      accumulation stage is three kernels -- tile pass A, the perimeter-graph solve, the final tile pass --
    #      so its launch time is the sum of theirs; `kernels` also gives every kernel on its own share of those
    #      bytes (pass A: the 1 B code read, final pass: the 8 B count write).
    cells_rank = cells_total // world
    per_step = {k: v[0] / args.steps for k, v in phases.items()}          # ms per step, all launches of the phase
    acc_ms = per_step["acc_tile_a"] + per_step["acc_solve"] + per_step["acc_tile_b"] + per_step["strip_edge"]
    stage_ms = {"flow_direction": per_step["direction"], "flow_accumulation": acc_ms}
    stage_bytes = {"flow_direction": DIR_BYTES_PER_CELL, "flow_accumulation": ACC_BYTES_PER_CELL}
    stage_kernels = {"flow_direction": "direction_kernel",
                     "flow_accumulation": "acc_tile_kernel + pj_* (perimeter-graph solve) + acc_final_kernel"}

    def gbs(bytes_per_cell, ms_):
        return cells_rank * bytes_per_cell / (ms_ * 1e-3) / 1e9 if ms_ > 0 else None

    dom = max(stage_ms, key=stage_ms.get)
    achieved = gbs(stage_bytes[dom], stage_ms[dom]) or 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        if tj.get("rows") == S and tj.get("cols") == S and world == 1:
            keys = ["direction"] if dom == "flow_direction" else ["acc_tile_a", "acc_solve", "acc_tile_b"]
            traffic = float(sum(tj.get(k, 0.0) for k in keys))
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": f"{dom}: {stage_kernels[dom]}", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src, "bytes_per_cell": stage_bytes[dom],
        "avg_launch_ms": stage_ms[dom],
        "phases_ms_per_step": {k: round(v, 4) for k, v in per_step.items()},
        "stages": {k: {"ms": stage_ms[k], "bytes_per_cell": stage_bytes[k], "gbs": gbs(stage_bytes[k], stage_ms[k]),
                       "frac": (gbs(stage_bytes[k], stage_ms[k]) or 0.0) / peak} for k in stage_ms},
        "kernels": {k: {"ms": per_step[k], "bytes_per_cell": PHASE_BYTES[k], "gbs": gbs(PHASE_BYTES[k], per_step[k]),
                        "frac": (gbs(PHASE_BYTES[k], per_step[k]) or 0.0) / peak} for k in PHASE_BYTES},
        "whole_step_13B_per_cell": {"ms": ms_per_step, "gbs": gbs(13.0, ms_per_step),
                                    "frac": (gbs(13.0, ms_per_step) or 0.0) / peak},
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32->u8->int64", "data": "synthetic", "config": workload_config(args, world),
        "clocks": clocks.summary(), "gpu_launches": launches, "roofline": roofline,
    }

    # ---- parity on the benchmarked result (outside the timed region)
    if world == 1 and not args.no_check:
        import oracle

        bad = dev.check_accumulation(fdr, fac)
        rng = np.random.default_rng(0)
        wsz = min(512, S - 2)
        dir_ok = True
        for _ in range(4):
            r, c = int(rng.integers(1, S - wsz - 1)), int(rng.integers(1, S - wsz - 1))
            win = dem[r - 1 : r + wsz + 1, c - 1 : c + wsz + 1].cpu().numpy()
            want = oracle.flow_direction_for_tile(win, NODATA)[1:-1, 1:-1]
            dir_ok &= bool(np.array_equal(fdr[r : r + wsz, c : c + wsz].cpu().numpy(), want))
        line["parity"] = {"accumulation_recurrence_violations": bad, "direction_windows_vs_oracle": dir_ok,
                          "max_fac": int(fac.max().item())}

    # ---- CPU baseline: the oracle port on a bounded sample of the same DEM (rank 0, N=1 only)
    if world == 1 and not args.no_cpu:
        import oracle

        sample = min(S, args.cpu_sample)
        dem_pad = np.full((sample + 2, sample + 2), NODATA, dtype=np.float32)
        dem_pad[1:-1, 1:-1] = dem[:sample, :sample].cpu().numpy()
        cpu_pass(oracle, dem_pad[:66, :66])  # warm-up (thread pool)
        best = None
        for _ in range(2):
            _, _, td, ta = cpu_pass(oracle, dem_pad)
            if best is None or td + ta < best[0] + best[1]:
                best = (td, ta)
        line["cpu_baseline"] = {
            "value": sample * sample / (best[0] + best[1]) / 1e9, "unit": UNIT, "cores": oracle.num_threads(),
            "kind": "port",
            "sample": f"{sample}x{sample} window of the benchmark DEM; direction {best[0] * 1e3:.0f} ms on "
                      f"{oracle.num_threads()} threads, accumulation {best[1] * 1e3:.0f} ms on 1 thread (serial algorithm)",
        }

    # ---- end to end through the public host API: pinned host buffers, H2D + D2H inside the timed region
    if world == 1 and not args.no_e2e:
        from overflow_b200.flow_routing import flow_routing_for_raster

        h_dem = torch.empty((S, S), dtype=torch.float32, pin_memory=True)
        h_dem.copy_(dem)
        del dem, fdr, fac
        torch.cuda.empty_cache()
        h_fdr = torch.empty((S, S), dtype=torch.uint8, pin_memory=True)
        h_fac = torch.empty((S, S), dtype=torch.int64, pin_memory=True)
        n_dem, n_fdr, n_fac = h_dem.numpy(), h_fdr.numpy(), h_fac.numpy()

        def e2e_step():
            flow_routing_for_raster(n_dem, NODATA, out_fdr=n_fdr, out_fac=n_fac)

        e2e_step()  # warm-up: allocates the library's device staging
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        el = (time.perf_counter() - t0) / args.e2e_steps
        line["e2e"] = {
            "value": cells_total / el / 1e9, "unit": UNIT,
            "h2d_bytes_per_step": int(cells_total * 4), "d2h_bytes_per_step": int(cells_total + cells_total * 8),
            "ms_per_step": el * 1e3, "steps": args.e2e_steps,
            "api": "overflow_b200.flow_routing.flow_routing_for_raster (C ABI ofl_flow_routing_f32, OFL_MEM_HOST) on pinned "
                   "host arrays: DEM in, codes and counts out",
        }
    elif world > 1 and not args.no_e2e:
        # every rank streams its strip from / to pinned host memory around the distributed step
        h_dem = torch.empty(tuple(runner.dem.shape), dtype=torch.float32, pin_memory=True)
        h_dem.copy_(runner.dem)
        h_fdr = torch.empty(tuple(runner.fdr.shape), dtype=torch.uint8, pin_memory=True)
        h_fac = torch.empty(tuple(runner.fac.shape), dtype=torch.int64, pin_memory=True)

        def e2e_step():
            runner.load_dem(h_dem)
            runner.step()
            h_fdr.copy_(runner.fdr, non_blocking=True)
            h_fac.copy_(runner.fac, non_blocking=True)
            torch.cuda.synchronize()

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        el = torch.tensor([(time.perf_counter() - t0) / args.e2e_steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
        el = float(el.item())
        line["e2e"] = {
            "value": cells_total / el / 1e9, "unit": UNIT,
            "h2d_bytes_per_step": int(cells_total * 4), "d2h_bytes_per_step": int(cells_total * 9),
            "ms_per_step": el * 1e3, "steps": args.e2e_steps,
            "api": "StripPipeline.load_dem(pinned host strip) + step() + copy of fdr/fac strips to pinned host memory, per rank",
        }

    # ---- SURVEY 8(f) row 2, reported next to the headline: flat resolution on a flat-heavy DEM (config 4 shape)
    if world == 1 and not args.no_flats:
        line["next_rows"] = {"fix_flats": flats_leg(torch, dev, peak), "breach_single_cell_pits": pits_leg(torch, dev, peak)}

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def flats_leg(torch, dev, peak, size=8192, window=1024):
    """ofl_fix_flats_f32 (resolve_flats + d8_masked_flow_dirs) on device buffers: terraced synthetic DEM, CUDA-event
    time of the call, a window checked against the CPU oracle, and the oracle timed on that window."""
    import numpy as np

    import oracle

    dem = dev.synth_dem(size, size, seed=3, kind=1, relief=200.0, holes_permille=5)
    fdr0 = dev.flow_direction(dem, NODATA).contiguous()
    work = dev.flats_workspace(size, size)
    flat_mask = torch.empty((size, size), dtype=torch.int32, device="cuda")
    labels = torch.empty((size, size), dtype=torch.int32, device="cuda")
    fdr = fdr0.clone()
    times, info = [], None
    for it in range(4):
        fdr.copy_(fdr0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, info = dev.fix_flats(dem, fdr, workspace=work, flat_mask=flat_mask, labels=labels)
        e1.record()
        torch.cuda.synchronize()
        if it:
            times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    cells = size * size
    hd, hf = dem[:window, :window].contiguous().cpu().numpy(), fdr0[:window, :window].contiguous().cpu().numpy()
    t0 = time.perf_counter()
    want_mask, want_labels = oracle.resolve_flats(hd, hf)
    want = oracle.d8_masked_flow_dirs(want_mask, hf, want_labels)
    cpu_s = time.perf_counter() - t0
    got = dev.fix_flats(torch.from_numpy(hd).cuda(), torch.from_numpy(hf).cuda())[0].cpu().numpy()
    return {
        "api": "ofl_fix_flats_f32 on device buffers (reference fix_flats.py: resolve_flats + d8_masked_flow_dirs)",
        "workload": f"synthetic {size}x{size} terraced float32 DEM (1 m steps, 5 permille nodata holes, seed 3)",
        "ms": ms, "value": cells / ms / 1e6, "unit": UNIT,
        "cells_without_direction": int((fdr0 == 8).sum().item()), "left_without_direction": int((fdr == 8).sum().item()),
        "flats": info[2], "sweep_levels": [info[3], info[4]],
        "roofline": {"bound": "hbm", "bytes_per_cell": 14.0, "achieved": cells * 14.0 / ms / 1e6, "peak": peak,
                     "unit": "GB/s", "frac": cells * 14.0 / ms / 1e6 / peak,
                     "note": "graph work (union-find + breadth-first sweeps): bound by dependent scattered accesses, not by streaming"},
        "parity": {"window": window, "codes_equal_oracle": bool(np.array_equal(got, want))},
        "cpu_baseline": {"value": window * window / cpu_s / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{window}x{window} window of the same DEM (the reference algorithm is serial)"},
    }


def pits_leg(torch, dev, peak, size=16384, window=2048):
    """ofl_breach_single_cell_pits_f32 on device buffers: fractal synthetic DEM, CUDA-event time of the call, a window
    breached on its own and compared with the CPU oracle, and the oracle timed on that window."""
    import numpy as np

    import oracle

    dem0 = dev.synth_dem(size, size, seed=3, kind=0, holes_permille=5)
    dem = dem0.clone()
    times, info = [], None
    for it in range(4):
        dem.copy_(dem0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, info = dev.breach_single_cell_pits(dem, NODATA)
        e1.record()
        torch.cuda.synchronize()
        if it:
            times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    cells = size * size
    hw = dem0[:window, :window].contiguous().cpu().numpy()
    t0 = time.perf_counter()
    want, want_unsolved = oracle.breach_single_cell_pits_in_chunk(hw, NODATA)
    cpu_s = time.perf_counter() - t0
    d_w = torch.from_numpy(hw.copy()).cuda()
    got_unsolved, _ = dev.breach_single_cell_pits(d_w, NODATA)
    ok = bool(np.array_equal(d_w.cpu().numpy().view(np.uint32), want.view(np.uint32))
              and np.array_equal(got_unsolved.cpu().numpy(), want_unsolved))
    return {
        "api": "ofl_breach_single_cell_pits_f32 on device buffers (reference breach_single_cell_pits_in_chunk)",
        "workload": f"synthetic {size}x{size} float32 DEM (fractal value-noise, 5 permille nodata holes, seed 3)",
        "ms": ms, "value": cells / ms / 1e6, "unit": UNIT, "pits": info[0], "unsolved": info[1], "rounds": info[2],
        "roofline": {"bound": "hbm", "bytes_per_cell": 5.0, "achieved": cells * 5.0 / ms / 1e6, "peak": peak, "unit": "GB/s",
                     "frac": cells * 5.0 / ms / 1e6 / peak,
                     "note": "4 B DEM read + 1 B unsolved raster written per cell; breached cells (about 1 %) are rewritten in place"},
        "parity": {"window": window, "bits_equal_oracle": ok},
        "cpu_baseline": {"value": window * window / cpu_s / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{window}x{window} window of the same DEM (the reference's second pass is serial)"},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--size", type=int, default=65536)
    ap.add_argument("--kind", type=int, default=0)
    ap.add_argument("--holes", type=int, default=5, help="nodata holes, permille of the raster")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=8192)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-flats", action="store_true", help="skip the fix_flats leg (next_rows)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
