#!/usr/bin/env python
"""bench.py — pivots/sec + tableau-update HBM GB/s on the BASELINE.json workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): cfg4 of BASELINE.json — the dense 16384 x 32768 fp64 LP
D(16384, 32768, seed 0) of SURVEY.md §8d.  One *step* = PIVOTS_PER_STEP (1000) consecutive
pivots of that tableau.  The full solve needs ~1e6 pivots, so timed steps simply continue the same
solve; the pivot sequence of the run is checked against the 2000-pivot golden prefix
(tests/golden/cfg_digests.json) before anything is printed.

  value   pivots/s with the tableau resident in HBM (CUDA events, max over ranks); the loop is the fused
          one: 8 pivots priced from the stored table, then ONE stream over the body applies them all
  e2e     pivots/s through the public API (SimplexMethod(rows, c).solve(...)) with the
          4.3 GB tableau in PINNED HOST memory: upload + pivots + result read-back
  roofline  the dominant kernel (update_fused_kernel): 16 B x cells per LAUNCH / its CUDA-event duration, plus
          its fp64-issue fraction; roofline_single_pivot_kernel: the one-pivot-per-pass streaming kernel
          (16 B x cells per pivot — the north star's roofline)
  cpu_baseline  oracle/spx_oracle.c (a C port of the reference's loop, OpenMP) on this host
  batched / l2_resident  the other BASELINE configs that fit one line: cfg3 (65,536 small LPs) and cfg2 (1000 x 2000)

N > 1 (torchrun): the body is column-sharded, one process per GPU; per pivot the ranks exchange their entering
keys and candidate columns over NVLink peer memory from inside the pricing kernel (strong scaling: the tableau is
fixed); the cfg3 batch is split over the ranks with no collective.
--impl reference: the reference's algorithm on the host cores (the oracle port; the
reference itself is pure Python and cannot travel to the GPU box), same config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ROWS, M_COLS, SEED = 16384, 32768, 0
PIVOTS_PER_STEP = 1000
METRIC = "pivots/sec (16k x 32k fp64 tableau)"
UNIT = "pivots/s"


def cells(n, m):
    return n * (m + 1) + m


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy_ burst)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def golden_trace():
    p = os.path.join(ROOT, "tests", "golden", "cfg_digests.json")
    with open(p) as fh:
        g = json.load(fh)["cfg4"]
    return g["input_sha256"], np.asarray(g["trace"], dtype=np.int32)


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed region runs (NVML)."""

    def __init__(self, index: int, period: float = 0.1):
        self.index, self.period = index, period
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {
                getattr(nv, "nvmlClocksEventReasonGpuIdle", 0x1): "gpu_idle",
                getattr(nv, "nvmlClocksEventReasonApplicationsClocksSetting", 0x2): "applications_clocks_setting",
                getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksEventReasonSyncBoost", 0x10): "sync_boost",
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
            }
            while not self._stop.is_set():
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for bit, name in names.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
                self._stop.wait(self.period)
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=2)

    def summary(self):
        return {"sm_mhz": int(statistics.median(self.sm)) if self.sm else None,
                "sm_max_mhz": self.max_mhz, "samples": len(self.sm), "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's algorithm on the host cores: oracle/spx_oracle.c (C port, OpenMP)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    from simplex_method_solver_b200 import workloads as W
    oracle.build()
    threads = oracle.num_threads()
    per_step = max(1, args.ref_pivots_per_step)
    log(f"[reference] generating D({N_ROWS},{M_COLS},{SEED}) ...")
    rows, c = W.dense_lp(N_ROWS, M_COLS, SEED)
    T = np.concatenate([rows.reshape(-1), c])
    del rows
    Nn = np.empty_like(T)
    L = oracle.lib()
    _, gold = golden_trace()
    k = 0

    def step():
        nonlocal T, Nn, k
        for _ in range(per_step):
            st, r, cc, _e = oracle.pick(T, N_ROWS, M_COLS)
            assert st == oracle.PIVOT
            if k < len(gold):
                assert (r, cc) == tuple(gold[k]), "oracle diverged from its own golden trace"
            L.orc_update(oracle._dp(T), oracle._dp(Nn), N_ROWS, M_COLS, r, cc)
            T, Nn = Nn, T
            k += 1

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{per_step} pivots per step x {args.steps} steps of the same 16384x32768 tableau "
                                   f"(oracle/spx_oracle.c, OpenMP {threads} threads; the reference itself is "
                                   f"single-threaded pure Python, ~1.5e6 cells/s)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(ngpu, exchange="fused"):
    return {"workload": "cfg4: dense LP D(n=16384, m=32768, seed=0), fp64 tableau 4.295 GB (x2 ping-pong)",
            "n": N_ROWS, "m": M_COLS, "cells": cells(N_ROWS, M_COLS),
            "pivots_per_step": PIVOTS_PER_STEP,
            "loop": "fused: 8 pivots priced from the stored table, then ONE stream over the body applies them (csrc/spx_fused.cu)"
            if ngpu == 1 else "fused, column-sharded: cooperative pricing with the in-kernel NVLink exchange of keys and "
            "the pivot column, then one stream over the local columns per 8 pivots" if exchange == "fused" else
            "pivot at a time, column-sharded: look-ahead pricing of pivot k+1 during update k (csrc/spx_shard.cu, spx_pick.cu)",
            "parallelism": "single GPU" if ngpu == 1 else
            f"column-sharded x{ngpu}, one key + candidate-column exchange per pivot over NVLink peer memory",
            "l2_policy": "inputs (8.6 GB per pivot) far exceed the 126 MB L2; no flush needed",
            "rule": "reference (first-negative entering, max-negative-ratio leaving)"}


# ------------------------------------------------------------------------------ our arm
def cpu_baseline_sample(rows, c, gold, budget_s=20.0):
    """oracle port timed on a bounded sample of the same workload (rank 0, N=1 only)."""
    import oracle
    oracle.build()
    threads = oracle.num_threads()
    T = np.concatenate([rows.reshape(-1), c])
    Nn = np.empty_like(T)
    L = oracle.lib()
    done, t_used = 0, 0.0
    # one untimed pivot to fault the pages in
    while True:
        t0 = time.perf_counter()
        st, r, cc, _e = oracle.pick(T, N_ROWS, M_COLS)
        assert st == oracle.PIVOT and (r, cc) == tuple(gold[done])
        L.orc_update(oracle._dp(T), oracle._dp(Nn), N_ROWS, M_COLS, r, cc)
        T, Nn = Nn, T
        dt = time.perf_counter() - t0
        done += 1
        if done > 1:
            t_used += dt
        if t_used > budget_s or done >= 64:
            break
    timed = done - 1
    return {"value": timed / t_used, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {timed} pivots (after 1 untimed) of the same 16384x32768 tableau, "
                      f"oracle/spx_oracle.c with OpenMP on {threads} threads"}


def batched_leg(dev, rank, world, dist=None, reps=20):
    """cfg3 of BASELINE.json: 65,536 independent 2-var / 8-constraint LPs, one warp per LP, the batch
    split contiguously over the ranks with no collective (weak unit: LPs; total work fixed).
    Returns the dict reported under "batched" (rank 0) — LPs/s resident and end to end."""
    import torch
    from simplex_method_solver_b200 import _native as N
    from simplex_method_solver_b200 import workloads as W
    from simplex_method_solver_b200.batched import DeviceBatch
    from simplex_method_solver_b200.parallel import shard_range
    B, n, m = 65536, 8, 2
    T, C = W.gui_batch(B, 0)
    tabs = W.batch_flat(T, C)
    start, count = shard_range(B, rank, world)
    pinned = torch.from_numpy(tabs[start:start + count].copy()).pin_memory()
    db = DeviceBatch(count, n, m, max_pivots=64, trace=True, device=dev)
    out_x = torch.empty((count, m), dtype=torch.float64).pin_memory()
    out_st = torch.empty(count, dtype=torch.int32).pin_memory()

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def timed(fn):
        for _ in range(3):
            fn()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def resident():
        db.T[:count].copy_(staged)          # device-to-device restore (the solver works in place)
        db.run()

    def e2e():
        db.upload(pinned, non_blocking=True)
        db.run()
        out_x.copy_(db.x[:count], non_blocking=True)
        out_st.copy_(db.status[:count], non_blocking=True)
        torch.cuda.current_stream().synchronize()

    staged = pinned.to(dev)
    ms_res = timed(resident)
    # the restore copy alone, subtracted so the resident figure is the solver kernel
    ms_copy = timed(lambda: db.T[:count].copy_(staged))
    ms_e2e = timed(e2e)
    res = db.result()
    piv = torch.tensor([int(res.npiv.sum()), int((res.status == 0).sum())], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(piv)
    total_piv, n_opt = int(piv[0].item()), int(piv[1].item())
    assert total_piv == 408212 and n_opt == B, (total_piv, n_opt)     # golden: tests/golden/cfg_digests.json
    ker_ms = max(ms_res - ms_copy, 1e-6)
    return {"workload": "cfg3: 65,536 LPs (8 constraints x 2 vars), one warp per LP, split over ranks, no collective",
            "lps_per_s": B / (ker_ms * 1e-3), "pivots_per_s": total_piv / (ker_ms * 1e-3), "kernel_ms": ker_ms,
            "e2e_lps_per_s": B / (ms_e2e * 1e-3), "e2e_ms": ms_e2e,
            "h2d_bytes": int(pinned.numel() * 8), "d2h_bytes": int(out_x.numel() * 8 + out_st.numel() * 4),
            "northstar_convention_GBps": 16.0 * 26 * total_piv / (ker_ms * 1e-3) / 1e9,
            "parity": "408,212 pivots, all optimal == golden"}


def resident_leg(dev):
    """cfg2 of BASELINE.json: dense 1000 x 2000 LP (16 MB, L2-resident), solved to optimality by the
    persistent cooperative kernel (csrc/spx_resident.cu); full pivot sequence checked against the golden digest."""
    import torch
    from simplex_method_solver_b200 import workloads as W
    from simplex_method_solver_b200.engine import DeviceTableau
    with open(os.path.join(ROOT, "tests", "golden", "cfg_digests.json")) as fh:
        g = json.load(fh)["cfg2"]["oracle_full"]
    rows, c = W.dense_lp(1000, 2000, 0)
    best = None
    for it in range(3):
        tab = DeviceTableau(1000, 2000, device=dev, trace_capacity=20000)
        tab.load(rows, c, max_pivots=20000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st, npiv = tab.solve(lookahead="resident")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    assert st == 0 and npiv == g["npiv"] == 13579
    assert W.pivot_digest(tab.trace[:npiv].cpu().numpy()) == g["pivot_sha256_final"]
    cells2 = cells(1000, 2000)
    return {"workload": "cfg2: dense LP D(1000, 2000, seed 0) to optimality, 13,579 pivots, persistent L2-resident kernel",
            "pivots_per_s": npiv / (best * 1e-3), "ms": best, "us_per_pivot": 1e3 * best / npiv,
            "northstar_convention_GBps": 16.0 * cells2 * npiv / (best * 1e-3) / 1e9,
            "parity": "13,579 pivots, sha256 of the pivot sequence == golden"}


def fused_depth_for(world: int) -> int:
    """Pivots per pass of the sharded fused loop: 8 everywhere (measured; see profiles/README.md)."""
    return 8


def run_ours(args):
    import torch
    import torch.distributed as dist
    from simplex_method_solver_b200 import _native as N
    from simplex_method_solver_b200 import workloads as W
    from simplex_method_solver_b200.engine import DeviceTableau
    from simplex_method_solver_b200.simplex import SimplexMethod

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    N.lib()
    peak, peak_src = measured_peak()
    in_sha, gold = golden_trace()

    log(f"[rank {rank}] generating D({N_ROWS},{M_COLS},{SEED}) ...")
    t0 = time.perf_counter()
    rows, c = W.dense_lp(N_ROWS, M_COLS, SEED)
    log(f"[rank {rank}] generated in {time.perf_counter() - t0:.1f}s")
    P = PIVOTS_PER_STEP
    need = (args.warmup + args.steps) * P

    if world == 1:
        # ---------------- value: tableau resident in HBM --------------------------------
        tab = DeviceTableau(N_ROWS, M_COLS, device=dev, trace_capacity=need + 64)
        tab.load(rows, c, max_pivots=need + 64)   # the cap is never reached inside the timed region
        for _ in range(args.warmup):
            st, npiv = tab.solve(chunk=P, stop_after=P)
        torch.cuda.synchronize()
        N.load().spx_launch_count(1)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(args.steps):
                st, npiv = tab.solve(chunk=P, stop_after=P)
            ev1.record()
            torch.cuda.synchronize()
        launches = int(N.load().spx_launch_count(0))
        total_ms = ev0.elapsed_time(ev1)
        assert npiv == need and st == N.PIVOT, (st, npiv)
        tr = tab.trace[:need].cpu().numpy()
        k = min(need, len(gold))
        assert (tr[:k] == gold[:k]).all(), "pivot sequence differs from the golden prefix"
        value = args.steps * P / (total_ms * 1e-3)

        # ---------------- roofline 1: the dominant kernel of the timed loop = update_fused_kernel (K6):
        # one launch streams the body once (16 B per cell) and applies FUSE_DEPTH pivots to every cell
        nmeas = 24
        F = int(N.load().spx_get_option(N.OPT_FUSE_DEPTH)) or 8
        st_obj = tab.read_state()
        st_obj.max_pivots = need + (nmeas + 3) * F + 64
        st_obj.reserved[0] = need & 1
        tab.write_state(st_obj)
        tab.trace = None
        fe = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(nmeas)]
        for q in range(nmeas):
            fe[q][0].record()
            tab.fused_pass(F, 1)                         # the pricing kernel (whole-GPU cooperative)
            fe[q][1].record()
            tab.fused_pass(F, 2)                         # the fused streaming update
            fe[q][2].record()
        torch.cuda.synchronize()
        price_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in fe[2:])
        fused_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in fe[2:])
        st_obj = tab.read_state()
        assert st_obj.status == N.PIVOT and st_obj.npiv == need + nmeas * F, (st_obj.status, st_obj.npiv)
        fused_npiv, fused_cur = int(st_obj.npiv), int(st_obj.reserved[0]) & 1
        alg_bytes = 16.0 * cells(N_ROWS, M_COLS)
        achieved = alg_bytes / (fused_ms * 1e-3) / 1e9
        sm_mhz = clk.summary()["sm_mhz"] or 0
        dp_ops = 6.0 * cells(N_ROWS, M_COLS) * F          # 2 DMUL + DADD + DMUL + 2 DFMA per cell per pivot
        dp_peak = 148 * 64 * sm_mhz * 1e6                 # fp64 issue slots/s at the SM clock seen under load
        traffic = None
        tp = os.path.join(ROOT, "profiles", "update_kernel_traffic.json")
        if os.path.exists(tp):
            with open(tp) as fh:
                traffic = json.load(fh).get("fused_dram_bytes_per_launch")
        step_ms = total_ms / args.steps
        roofline = {"bound": "hbm", "kernel": f"update_fused_kernel (K6: {F} pivots per pass over the body)",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "peak_source": peak_src, "traffic": traffic,
                    "algorithmic_bytes_per_launch": alg_bytes, "pivots_per_launch": F,
                    "kernel_ms": fused_ms, "pricing_kernel_ms": price_ms,
                    "kernel_share_of_step": fused_ms * (P / F) / step_ms,
                    "northstar_convention_GBps": alg_bytes * F / (fused_ms * 1e-3) / 1e9,
                    "fp64_pipe": {"achieved_Gops": dp_ops / (fused_ms * 1e-3) / 1e9,
                                  "peak_Gops_at_measured_clock": dp_peak / 1e9,
                                  "frac": (dp_ops / (fused_ms * 1e-3)) / dp_peak if dp_peak else None,
                                  "sm_mhz": sm_mhz},
                    "note": "F dependent rank-1 updates per cell in registers: HBM moves 16 B per cell per PASS, so at F=8 "
                            "the kernel is fp64-issue bound, not HBM bound; the single-pivot streaming kernel is below"}

        # ---------------- roofline 2: the single-pivot streaming kernel K3 (the 16 B/cell/pivot roofline)
        st_obj.max_pivots = fused_npiv + 64
        st_obj.reserved[0] = 0
        tab.write_state(st_obj)
        if fused_cur != (fused_npiv & 1):                 # restore "table k lives in buffer k & 1"
            tab.A[fused_npiv & 1].copy_(tab.A[fused_cur])
            tab.b[fused_npiv & 1].copy_(tab.b[fused_cur])
        nk3 = 40
        tab2_npiv = fused_npiv
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nk3)]
        pick_evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nk3)]
        for q in range(nk3):
            pick_evs[q][0].record()
            tab.pick(tab2_npiv, sticky=True)
            pick_evs[q][1].record()
            evs[q][0].record()
            tab.update(tab2_npiv)
            evs[q][1].record()
            tab2_npiv += 1
        torch.cuda.synchronize()
        # the first launches pay CUDA's lazy module load of kernels the fused loop never used: drop them
        upd_ms = statistics.mean(a.elapsed_time(b) for a, b in evs[5:])
        pick_ms = statistics.mean(a.elapsed_time(b) for a, b in pick_evs[5:])
        k3 = alg_bytes / (upd_ms * 1e-3) / 1e9
        k3_traffic = None
        if os.path.exists(tp):
            with open(tp) as fh:
                k3_traffic = json.load(fh).get("dram_bytes_per_launch")
        roofline_k3 = {"bound": "hbm", "kernel": "update_tiled_kernel (K3: one pivot per pass, 16 B per cell per pivot)",
                       "achieved": k3, "peak": peak, "unit": "GB/s", "frac": k3 / peak, "traffic": k3_traffic,
                       "algorithmic_bytes_per_launch": alg_bytes, "update_ms": upd_ms, "pick_ms": pick_ms,
                       "frac_of_nominal_8TBs": k3 / 8000.0}
        del tab
        torch.cuda.empty_cache()

        # ---------------- e2e: public API, pinned host input, upload inside the timed region
        pinned = torch.empty((N_ROWS, M_COLS + 1), dtype=torch.float64).pin_memory()
        pinned.numpy()[...] = rows
        host_rows = pinned.numpy()
        h2d = host_rows.nbytes + c.nbytes
        e2e_steps = max(1, min(args.steps, 3))
        e2e_t, d2h = [], 0
        for it in range(1 + e2e_steps):            # first iteration is a warm-up
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sm = SimplexMethod(host_rows, c, device=dev, engine="stream")
            sol = sm.solve(max_pivots=P, chunk=P)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert sol.npiv == P and (sol.trace == gold[:P]).all()
            d2h = sol.trace.nbytes + sol.x.nbytes + 16 + sol.rowlab.nbytes + sol.collab.nbytes
            if it > 0:
                e2e_t.append(dt)
            del sm
        e2e_val = P / statistics.mean(e2e_t)

        cpu = cpu_baseline_sample(rows, c, gold) if not args.no_cpu_baseline else None
        batched = batched_leg(dev, 0, 1) if not args.no_batched else None
        resident = resident_leg(dev) if not args.no_batched else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(1), "clocks": clk.summary(),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": 1e3 * statistics.mean(e2e_t),
                    "api": f"SimplexMethod(pinned_rows, c).solve(max_pivots={P})"},
            "gpu_launches": launches, "roofline": roofline, "roofline_single_pivot_kernel": roofline_k3,
            "cpu_baseline": cpu,
            "northstar_convention_GBps_whole_step": alg_bytes * args.steps * P / (total_ms * 1e-3) / 1e9,
            "batched": batched, "l2_resident": resident,
            "parity": f"pivot sequence == golden prefix for the first {k} pivots",
        }
        print(json.dumps(line), flush=True)
        return

    # ---------------- N > 1: column-sharded, one process per GPU ---------------------------
    from simplex_method_solver_b200.parallel import FusedShardedTableau, PeerShardedTableau, ShardedTableau
    fallback_note = None
    if args.exchange == "fused":
        # passes of 8 pivots: cooperative pricing with the in-kernel NVLink exchange, then ONE stream
        # over the local columns applies them all (csrc/spx_fused.cu).  If peer memory cannot be mapped on
        # this box (no CUDA IPC / P2P between the GPUs) every rank falls back to the NCCL all-gather flow.
        sh, err = None, ""
        try:
            sh = FusedShardedTableau(N_ROWS, M_COLS, rank, world, device=dev, trace_capacity=need + 64,
                                     depth=args.depth or fused_depth_for(world))
        except Exception as e:                       # noqa: BLE001 - reported below, then the documented fallback
            err = f"{type(e).__name__}: {e}"
        okf = torch.tensor([1 if sh is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(okf, op=dist.ReduceOp.MIN)
        if int(okf.item()) == 0:
            log(f"[rank {rank}] peer-memory exchange unavailable ({err or 'a peer failed'}); falling back to --exchange nccl")
            fallback_note = f"peer memory unavailable ({err or 'on a peer'})"
            if sh is not None:
                sh.close()
            args.exchange = "nccl"
            sh = ShardedTableau(N_ROWS, M_COLS, rank, world, device=dev, trace_capacity=need + 64,
                                lookahead=not args.no_lookahead)
        else:
            # preflight: a few passes against the golden prefix on EVERY rank before anything is timed; a rank
            # that times out on a peer (SPX_PEER_TIMEOUT) or diverges sends all ranks to the pivot-at-a-time
            # peer-mailbox loop instead (csrc/spx_shard.cu), and the printed config says so
            pre, why = 96, ""
            try:
                sh.load(rows, c, max_pivots=pre + 64)
                sh.run(pre)
                stp = sh.sync()
                trp = sh.trace[:pre].cpu().numpy()
                if stp.status != N.PIVOT or stp.npiv != pre:
                    why = f"status {stp.status} after {stp.npiv} pivots"
                elif not (trp == gold[:pre]).all():
                    why = "pivot sequence differs from the golden prefix"
            except Exception as e:                   # noqa: BLE001
                why = f"{type(e).__name__}: {e}"
            okf = torch.tensor([0 if why else 1], dtype=torch.int32, device=dev)
            dist.all_reduce(okf, op=dist.ReduceOp.MIN)
            if int(okf.item()) == 0:
                log(f"[rank {rank}] fused exchange failed its preflight ({why or 'on a peer'}); falling back to --exchange p2p")
                fallback_note = f"fused preflight failed ({why or 'on a peer'})"
                sh.close()
                args.exchange = "p2p"
                sh = PeerShardedTableau(N_ROWS, M_COLS, rank, world, device=dev, trace_capacity=need + 64)
    elif args.exchange == "p2p":
        # C-side look-ahead loop, candidates exchanged by NVLink peer stores (csrc/spx_shard.cu)
        sh = PeerShardedTableau(N_ROWS, M_COLS, rank, world, device=dev, trace_capacity=need + 64)
    else:
        sh = ShardedTableau(N_ROWS, M_COLS, rank, world, device=dev, trace_capacity=need + 64,
                            lookahead=not args.no_lookahead)
    sh.load(rows, c, max_pivots=need + 64)
    for _ in range(args.warmup):
        sh.run(P)
    torch.cuda.synchronize()
    dist.barrier()
    N.load().spx_launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        dist.barrier()
        ev0.record()
        for _ in range(args.steps):
            sh.run(P)
        ev1.record()
        torch.cuda.synchronize()
        dist.barrier()
    launches = int(N.load().spx_launch_count(0))
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    st = sh.sync()
    assert st.npiv == need and st.status == N.PIVOT, (st.status, st.npiv)
    tr = sh.trace[:need].cpu().numpy()
    k = min(need, len(gold))
    assert (tr[:k] == gold[:k]).all(), "sharded pivot sequence differs from the golden prefix"
    value = args.steps * P / (total_ms * 1e-3)
    alg_bytes = 16.0 * cells(N_ROWS, M_COLS)

    # ---------------- e2e at N GPUs: every rank uploads ITS column block from pinned host memory, the
    # ranks pivot together, every rank reads its state and trace back; wall clock between barriers, max over ranks
    e2e = None
    if args.exchange in ("fused", "p2p"):
        blk = torch.empty((N_ROWS, sh.m_loc + 1), dtype=torch.float64).pin_memory()
        blk.numpy()[:, : sh.m_loc] = rows[:, sh.col0: sh.col0 + sh.m_loc]
        blk.numpy()[:, sh.m_loc] = rows[:, M_COLS]
        cblk = np.ascontiguousarray(c[sh.col0: sh.col0 + sh.m_loc])
        e2e_steps = max(1, min(args.steps, 3))
        ts = []
        for it in range(1 + e2e_steps):                     # first iteration is a warm-up
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            sh.load_local(blk.numpy(), cblk, max_pivots=P + 64)
            sh.run(P)
            st2 = sh.sync()
            tr2 = sh.trace[:P].cpu().numpy()
            torch.cuda.synchronize()
            dist.barrier()
            dt = time.perf_counter() - t0
            assert st2.npiv == P and (tr2 == gold[:P]).all(), "e2e: sharded pivot sequence differs from the golden prefix"
            if it > 0:
                ts.append(dt)
        tmax = torch.tensor([statistics.mean(ts)], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e = {"value": P / float(tmax.item()), "unit": UNIT,
               "h2d_bytes_per_step": int(8 * (N_ROWS * (M_COLS + world) + M_COLS)),
               "d2h_bytes_per_step": int(world * (tr2.nbytes + 128)), "steps": e2e_steps,
               "ms_per_step": 1e3 * float(tmax.item()),
               "api": f"{type(sh).__name__}.load_local(pinned_block, c_block); run({P}); sync() on every rank"}
        del blk
    del rows
    if args.exchange in ("p2p", "fused"):    batched = batched_leg(dev, rank, world, dist) if not args.no_batched else None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config_dict(world, args.exchange), exchange=args.exchange,
                           **({"exchange_fallback": fallback_note} if fallback_note else {})),
            "clocks": clk.summary(),
            "e2e": e2e, "gpu_launches": launches,
            "roofline": {"bound": "hbm",
                         "kernel": ("update_fused_kernel (K6)" if args.exchange == "fused" else "update_tiled_kernel (K3)") +
                                   ": WHOLE step incl. pricing and exchange, 16 B x cells per pivot / step time / N",
                         "achieved": alg_bytes * args.steps * P / (total_ms * 1e-3) / 1e9 / world,
                         "peak": peak, "unit": "GB/s per GPU", "peak_source": peak_src,
                         "frac": alg_bytes * args.steps * P / (total_ms * 1e-3) / 1e9 / world / peak,
                         "traffic": None},
            "cpu_baseline": None, "batched": batched,
            "parity": f"sharded pivot sequence == golden prefix for the first {k} pivots",
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-pivots-per-step", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batched", action="store_true", help="skip the cfg3 batched-LP leg")
    ap.add_argument("--depth", type=int, default=0, help="N>1 fused loop: pivots per pass (0 = default)")
    ap.add_argument("--exchange", default="fused", choices=["fused", "p2p", "nccl"],
                    help="N>1: fused passes with the in-kernel NVLink exchange (default); pivot-at-a-time look-ahead "
                         "with NVLink peer mailboxes (p2p) or an NCCL all-gather (nccl)")
    ap.add_argument("--no-lookahead", action="store_true",
                    help="classic pick->update order instead of pricing pivot k+1 during update k")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("note: the timing rules ask for >= 3 warm-up steps")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
