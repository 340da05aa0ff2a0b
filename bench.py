#!/usr/bin/env python
"""bench.py -- throughput of the N-body hot path on B200 (BASELINE.json metric), one JSON line.

    python bench.py --gpus 1 --steps 20 --warmup 3              # 1 GPU
    torchrun --nproc-per-node N ... bench.py --gpus N ...       # N GPUs, one rank per GPU
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU path (oracle)

A "step" is one nb_step_brute_force over the whole particle set (forces from old positions + Euler).
Workloads (BASELINE.json configs; default c3 = the configuration the headline metric and the north-star
target are quoted on, the 1,048,576-body all-pairs system -- it fits one GPU, and using the same fixed
problem at every N makes the 1/2/4/8-GPU line a strong-scaling curve):
    c2  65,536-body Plummer (2-D projection), all-pairs
    c3  1,048,576-body Plummer, all-pairs                     [default]
    c4  262,144-body uniform disk, Barnes-Hut theta=0.5
    c5  4,194,304-body uniform disk, Barnes-Hut theta=0.75
metric  = pair interactions per second, N(N-1) ordered pairs per step (rs-src/nbody.rs:108-110 evaluates both
          directions); for the Barnes-Hut workloads the line reports body-steps/s and steps/s instead.
value   = whole-job device throughput, state resident in HBM, CUDA events, max over ranks.
e2e     = the same steps through the reference-facing C ABI with HOST buffers: every step does
          nb_set_particles (H2D from pinned memory) + nb_step_* + nb_get_particles (D2H), wall clock.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR = 12  # <2,REF>: 2 sub + 2 FMA + rcp + mul + 2 FMA
FLOP_PER_PAIR_3D_NEWTON = 20  # <3,NEWTON>: 3 sub + 3 FMA + rsqrt + 2 mul (cube) + mul + 3 FMA (GPU Gems 3 ch.31 convention)
WORKLOADS = {
    "c2": dict(n=65536, kind="allpairs", ic="plummer", seed=2, desc="65,536-body Plummer (2-D), all-pairs"),
    "c3": dict(n=1 << 20, kind="allpairs", ic="plummer", seed=3, desc="1,048,576-body Plummer (2-D), all-pairs"),
    "c2n": dict(n=65536, kind="allpairs3", ic="plummer3", seed=2, desc="65,536-body Plummer sphere, 3-D Newtonian (rsqrt) all-pairs [nbx3 extension]"),
    "c3n": dict(n=1 << 20, kind="allpairs3", ic="plummer3", seed=3, desc="1,048,576-body Plummer sphere, 3-D Newtonian (rsqrt) all-pairs [nbx3 extension]"),
    "c4": dict(n=262144, kind="bh", theta=0.5, ic="disk", seed=4, desc="262,144-body uniform disk, Barnes-Hut theta=0.5"),
    "c5": dict(n=1 << 22, kind="bh", theta=0.75, ic="disk", seed=5, desc="4,194,304-body uniform disk, Barnes-Hut theta=0.75"),
}
DT = 0.01


def make_ic(w):
    from rust_exp_b200 import ic

    if w["ic"] == "plummer3":
        return ic.plummer_3d(w["n"], seed=w["seed"])
    if w["ic"] == "plummer":
        return ic.plummer_2d(w["n"], seed=w["seed"])
    return ic.random_disk(w["n"], seed=w["seed"])


def peaks():
    p = {}
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            inside = t0 - 0.05 <= t <= t1 + 0.15
            try:
                if inside:
                    sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
                    for nm, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(nm)
            except ValueError:
                pass
        if not sm:  # region shorter than the sampling period: use everything we saw
            for t, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
                except Exception:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline_allpairs(state, seconds=12.0, all_cores=True):
    """The reference's CPU all-pairs force loop (oracle restatement, rs-src/nbody.rs:132-144, ONE thread as
    in the reference) on a bounded sample of rows of the same particle set."""
    import oracle

    o = oracle.get()
    o.set_particles(state)
    n = state.shape[0]
    t = time.perf_counter(); o.brute_forces_rows(0, min(n, 8)); probe = (time.perf_counter() - t) / min(n, 8)
    rows = int(max(8, min(n, seconds / max(probe, 1e-9))))
    t = time.perf_counter(); o.brute_forces_rows(0, rows); dt = time.perf_counter() - t
    out = {"value": rows * (n - 1) / dt, "unit": "pair-interactions/s", "cores": 1, "kind": "port",
           "sample": f"forces on the first {rows} of {n} bodies against all {n} (rs-src/nbody.rs:132-144, single thread as in the reference), {dt:.1f} s"}
    if all_cores:
        nc = os.cpu_count() or 1
        rows2 = min(n, rows * nc // 2)
        t = time.perf_counter(); o.brute_forces_rows(0, rows2, nthreads=nc); dt2 = time.perf_counter() - t
        out["all_cores_not_in_reference"] = {"value": rows2 * (n - 1) / dt2, "cores": nc,
                                             "note": "i loop split over host threads; the reference's brute force is single-threaded"}
    return out


def cpu_baseline_bh(state, theta, seconds=12.0):
    import oracle

    o = oracle.get()
    nc = os.cpu_count() or 1
    o.set_particles(state)
    t = time.perf_counter(); o.step_barnes_hut(theta, DT, nc); dt = time.perf_counter() - t
    steps = 1
    while dt < seconds / 3 and steps < 4:
        t = time.perf_counter(); o.step_barnes_hut(theta, DT, nc); dt = min(dt, time.perf_counter() - t); steps += 1
    n = state.shape[0]
    return {"value": n / dt, "unit": "body-steps/s", "cores": nc, "kind": "port", "steps_per_s": 1.0 / dt,
            "sample": f"{steps} full nb_step_barnes_hut steps of {n} bodies, nthreads={nc} (serial tree build + threaded walk, rs-src/nbody.rs:413-478), best step {dt:.2f} s"}


def ref_bh_rate(o, state, theta, max_steps, budget_s):
    """Oracle Barnes-Hut steps (serial tree build + nthreads walk, rs-src/nbody.rs:413-478) on all host cores."""
    nc = os.cpu_count() or 1
    o.set_particles(state)
    t0 = time.perf_counter()
    k = 0
    while k < max_steps and (k == 0 or (time.perf_counter() - t0) * (k + 1) / k < budget_s):
        o.step_barnes_hut(theta, DT, nc)
        k += 1
    el = time.perf_counter() - t0
    n = state.shape[0]
    return {"steps_per_s": k / el, "body_steps_per_s": n * k / el, "ms_per_step": el / k * 1e3, "steps": k, "cores": nc,
            "kind": "port", "sample": f"{k} full nb_step_barnes_hut steps of {n} bodies, theta={theta}, nthreads={nc} "
                                      "(serial tree build + threaded walk, rs-src/nbody.rs:413-478)"}


def run_reference(args, w):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the reference is
    Rust and cannot be built in this image), timed on the box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if w["kind"] == "allpairs3":
        print(json.dumps({"impl": "reference", "unavailable": "the reference (rs-src/nbody.rs) is 2-D: the nbx3 workloads have no reference arm"}))
        return 0
    import oracle

    o = oracle.get()
    state = make_ic(w)
    n = w["n"]
    o.set_particles(state)
    nc = os.cpu_count() or 1
    extra = {}
    if w["kind"] == "allpairs":
        # bounded sample per step: rows sized for ~2 s of single-thread work
        t = time.perf_counter(); o.brute_forces_rows(0, 8); probe = (time.perf_counter() - t) / 8
        rows = int(max(8, min(n, 2.0 / probe)))
        for _ in range(args.warmup):
            o.brute_forces_rows(0, max(8, rows // 8))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            o.brute_forces_rows(0, rows)
        el = time.perf_counter() - t0
        value, unit, cores = rows * (n - 1) * args.steps / el, "pair-interactions/s", 1
        sample = f"per step: forces on {rows} of {n} bodies against all {n}, single thread (the reference's brute force is single-threaded, rs-src/nbody.rs:132-144)"
        metric = "pair-interactions/s"
        t = time.perf_counter(); o.brute_forces_rows(0, min(n, rows * nc // 2), nthreads=nc); dtm = time.perf_counter() - t
        all_cores = {"value": min(n, rows * nc // 2) * (n - 1) / dtm, "cores": nc,
                     "note": "NOT the reference: its i loop split over all host threads, for context only"}
        if args.workload == DEFAULT_WORKLOAD and not args.no_bh:
            # the metric also names "BH steps/s": the reference arm's Barnes-Hut rates for the same sub-workloads
            bh = {}
            for name in bh_sub_workloads(args.gpus):
                wb = WORKLOADS[name]
                bh[name] = dict(ref_bh_rate(o, make_ic(wb), wb["theta"], 3, 20.0), workload=f"{name}: {wb['desc']}")
            extra["bh"] = bh
    else:
        for _ in range(min(args.warmup, 1)):
            o.step_barnes_hut(w["theta"], DT, nc)
        k = max(1, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(k):
            o.step_barnes_hut(w["theta"], DT, nc)
        el = time.perf_counter() - t0
        value, unit, cores = n * k / el, "body-steps/s", nc
        sample = f"{k} full Barnes-Hut steps, nthreads={nc}"
        metric = "body-steps/s"
        args.steps = k
        all_cores = None
        extra["bh_steps_per_s"] = k / el
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['desc']}", "n_bodies": n, "dt": DT, "ic": w["ic"], "seed": w["seed"]},
            "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample,
                             **({"all_cores_not_in_reference": all_cores} if all_cores else {})},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line.update(extra)
    print(json.dumps(line))
    return 0


DEFAULT_WORKLOAD = "c3"


def bh_sub_workloads(n_gpus):
    """Barnes-Hut sub-records of the default line: c4 (BASELINE.json configs[3], quoted on 1 GPU) at N=1, c5
    (configs[4], quoted on 8 GPUs, fits one) at every N."""
    return ["c4", "c5"] if n_gpus == 1 else ["c5"]


class Ctx:
    """Everything a measurement needs: library handle, torch stream, process-group facts."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import rust_exp_b200 as pkg
        from rust_exp_b200 import binding
        from rust_exp_b200 import dist as nbdist

        self.torch, self.dist, self.binding = torch, dist, binding
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- libnbody_b200 has no CPU path (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        assert self.world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={self.world}: launch with torchrun --nproc-per-node {args.gpus}"
        self.lib = pkg.load()
        self.lib.init(self.local_rank)
        # a real (non-default) torch stream: handle 0 would mean "library's own stream" to nbx_set_stream,
        # and torch.cuda.Event only sees the stream it is recorded on
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        assert self.stream.cuda_stream != 0
        self.lib.set_stream(self.stream.cuda_stream)  # torch CUDA events then bracket the library's launches
        self.lib.tune(args.bodies_per_thread, args.waves, args.ctas_per_sm)
        self.transport_name = args.transport
        names = [args.workload] + (bh_sub_workloads(self.world) if (args.workload == DEFAULT_WORKLOAD and not args.no_bh) else [])
        self.max_n = max(max(WORKLOADS[k]["n"] for k in names), 65536)
        if self.world > 1:
            # the sharded nbx3 extension exchanges positions with the library's own ncclAllGather: it needs the communicator
            nbdist.wire(self.lib, self.max_n, {"p2p_direct": 0, "p2p_gather": 1, "nccl": 2}[args.transport],
                        nccl=WORKLOADS[args.workload]["kind"] == "allpairs3")
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
        pk = peaks()
        self.pk = pk
        self.sm_max = pk.get("sm_max_mhz", 1965.0)
        self.sms = torch.cuda.get_device_properties(self.local_rank).multi_processor_count
        self.fp32_peak = 2 * 128 * self.sms * self.sm_max * 1e6 / 1e12  # TFLOP/s, SURVEY.md H9

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allreduce(self, vals, op="sum", dtype=None):
        torch = self.torch
        t = torch.tensor(vals, dtype=dtype or torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op={"sum": self.dist.ReduceOp.SUM, "max": self.dist.ReduceOp.MAX}[op])
        return t.tolist()


def measure(cx, name, steps, warmup, e2e_steps, with_cpu_baseline, sample_clocks=True):
    """One workload: device-timed K steps (value), the same through the host-buffer C ABI (e2e), roofline, phases."""
    torch, lib, world, rank = cx.torch, cx.lib, cx.world, cx.rank
    w = WORKLOADS[name]
    n = w["n"]
    is3 = w["kind"] == "allpairs3"
    width = 7 if is3 else 5
    # pinned host state (the e2e leg copies from / to it every step)
    host = torch.empty((n, width), dtype=torch.float32, pin_memory=True)
    host.numpy()[:] = make_ic(w)
    host_out = torch.empty((n, width), dtype=torch.float32, pin_memory=True)
    if is3:
        lib.configure3(cx.binding.LAW3_NEWTON, 1e-4)
        lib.set_particles3(host.numpy())
    else:
        lib.set_particles(host.numpy())

    def step():
        if w["kind"] == "allpairs":
            lib.step_brute_force(DT)
        elif is3:
            lib.step3(DT)
        else:
            lib.step_barnes_hut(w["theta"], DT, 1)

    stream, flush = cx.stream, cx.flush
    # device-timed region: state resident, steps enqueued back to back (nbx_set_async) so that the CUDA events
    # bracket GPU work only; the e2e leg below uses the default, reference-like synchronous calls
    lib.set_async(True)
    for _ in range(warmup):
        step()
        flush.zero_()
    cx.barrier()
    bh_stats = None
    if w["kind"] == "bh":
        # one instrumented step (outside the timed region): interactions, node visits and lane efficiency per step
        lib.bh_count_interactions(True)
        lib.reset_counters()
        lib.bh_pop_histogram(True)
        step()
        lib.synchronize()
        c = lib.counters()
        hist = lib.bh_pop_histogram(True)
        bh_stats = {"interactions_per_step": c["bh_interactions"], "nodes_visited_per_step": c["bh_nodes_visited"],
                    "tree_nodes": c["bh_nodes_built"], "pops": c["bh_pops"], "pop_lanes": c["bh_pop_lanes"]}
        bh_hist = [int(v) for v in hist]
        lib.bh_count_interactions(False)
        cx.barrier()

    lib.reset_counters()
    # All-pairs: the library's per-phase events ride along in the timed region (2 kernels per step, the extra
    # event pairs are noise).  Barnes-Hut: a step is ~15 short kernels, so the timed region runs WITHOUT them
    # and the per-phase times come from a second, instrumented pass of the same steps afterwards.
    phases_in_timed_region = w["kind"] != "bh"
    lib.phase_timing(phases_in_timed_region)
    sampler = ClockSampler(cx.local_rank) if sample_clocks else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    # K steps, each bracketed by its own CUDA-event pair on the library's stream; the L2 flush (a 256 MiB
    # memset, ~70 us) runs BETWEEN the timed iterations, outside the event pairs.  ms = sum of the K step times.
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    cx.barrier()
    t0 = time.perf_counter()
    for a, b in evs:
        a.record(stream)
        step()
        b.record(stream)
        flush.zero_()
    cx.barrier()
    t1 = time.perf_counter()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    clocks = sampler.stop(t0, t1) if sampler else None
    ctr = lib.counters()
    sort_levels_timed = ctr.get("bh_sort_levels")   # key levels the timed steps sorted on (the e2e leg below resets it: new set each step)
    if phases_in_timed_region:
        phases = lib.phase_ms()
    else:
        lib.phase_timing(True)
        for _ in range(min(steps, 20)):
            step()
            flush.zero_()
        phases = lib.phase_ms()
    lib.phase_timing(False)
    lib.set_async(False)
    ms = cx.allreduce([ms], "max")[0]
    launches = int(cx.allreduce([ctr["kernel_launches"]], "sum")[0])
    ms_per_step = ms / steps

    # ---- e2e through the C ABI with host buffers (synchronous step calls, as the reference host makes them) ----
    e2e_steps = max(1, e2e_steps)
    cx.barrier()
    te0 = time.perf_counter()
    for _ in range(e2e_steps):
        if is3:
            lib.set_particles3(host.numpy())
            step()
            lib.L.nbx3_get_particles(host_out.numpy().ctypes.data, n)
            continue
        lib.set_particles(host.numpy())        # H2D (each rank uploads its shard of the pinned array)
        step()
        if world > 1:
            lib.get_particles_local(host_out.numpy())   # D2H of this rank's shard into its rows of the host array
        else:
            lib.get_particles(host_out.numpy())          # D2H of the state (synchronises)
    torch.cuda.synchronize()
    e2e_s = cx.allreduce([time.perf_counter() - te0], "max")[0] / e2e_steps
    b, c = lib.dist_local_range() if world > 1 else (0, n)
    if is3 and world > 1:   # rows of the nbx3 set this rank evaluates
        from rust_exp_b200.dist import nbx3_local_rows
        b, c = nbx3_local_rows(n, rank, world)
    h2d = int(cx.allreduce([4 * width * c], "sum")[0])
    d2h = h2d
    if is3 and world > 1:   # every rank uploads and reads back the whole nbx3 set
        h2d = d2h = 4 * width * n * world

    fp32_peak, pk, sms, sm_max = cx.fp32_peak, cx.pk, cx.sms, cx.sm_max
    out = {"workload": f"{name}: {w['desc']}", "n_bodies": n, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
           "gpu_launches": launches, "phases_ms": phases, "ic": w["ic"], "seed": w["seed"]}
    if clocks is not None:
        out["clocks"] = clocks
    e2e_what = ("nb_set_particles(pinned host AoS) + nb_step (synchronous) + " +
                ("nbx_get_particles_local (each rank reads back its own shard's rows)" if world > 1 else "nb_get_particles(pinned host AoS)") +
                ", wall clock, max over ranks")
    if w["kind"] in ("allpairs", "allpairs3"):
        fpp = FLOP_PER_PAIR_3D_NEWTON if is3 else FLOP_PER_PAIR
        pairs_per_step = n * (n - 1)
        out["value"] = pairs_per_step * steps / (ms * 1e-3)
        e2e_value = pairs_per_step / e2e_s
        out["metric"] = out["unit"] = "pair-interactions/s"
        force_ms = phases["force"] if not is3 else ms_per_step   # nbx3 has no phase events: 2 kernels, integrate ~0.01 ms
        # dominant kernel = allpairs_fast_kernel: each rank's launch evaluates n_local*(n-1) pairs
        per_launch_pairs = c * (n - 1)
        achieved = per_launch_pairs * fpp / (force_ms * 1e-3) / 1e12 if force_ms > 0 else None
        if world > 1:   # the slowest rank's kernel defines the step; ranks hold equal shards
            achieved = cx.allreduce([-(achieved or 0.0)], "max")[0] * -1.0 or None
        kname = "allpairs3_kernel<NEWTON>" if is3 else "allpairs_fast_kernel"
        out["roofline"] = {"bound": "fp32", "kernel": kname, "achieved": achieved, "peak": fp32_peak,
                           "unit": "TFLOP/s", "frac": achieved / fp32_peak if achieved else None,
                           **ncu_traffic(kname, name, world),
                           "flop_per_pair": fpp, "kernel_ms": force_ms, "kernel_share_of_step": force_ms / ms_per_step,
                           "peak_source": f"computed 2*128*{sms} SMs*{sm_max:.0f} MHz (MEASURED_PEAKS.json has no FP32 figure; sm_max_mhz is from it); "
                                          "the kernel is co-limited by the MUFU pipe at 75% of this (DESIGN.md section 4)",
                           "mufu_bound_frac": (achieved / (0.75 * fp32_peak)) if achieved else None,
                           # measured on this pool's B200 with rust_exp_b200/nb_microbench (profiles/r01_microbench_pipes.jsonl):
                           # packed-FP32 FFMA2 sustains 71.3 TFLOP/s (95.7 % of nominal), MUFU.RCP 4.62e12/s
                           "peak_measured_ffma2": 71.3, "frac_of_measured_ffma2": (achieved / 71.3) if achieved else None}
        # the same launch against the HBM roof, for completeness: algorithmic bytes = every j position read once
        # (12 or 16 B per body of the whole set) + one acceleration written per local body.  It is ~1e-5 of the HBM
        # peak -- the position set lives in L2 and is re-read N/512 times from there; HBM is not the binding roof.
        hbm = pk.get("hbm_gbs", 6650.0)
        alg_bytes = (16 if is3 else 12) * n + (12 if is3 else 8) * c
        hbm_achieved = alg_bytes / (force_ms * 1e-3) / 1e9 if force_ms > 0 else None
        out.update({"fp32_tflops": out["value"] * fpp / 1e12, "fp32_frac_of_peak_all_gpus": out["value"] * fpp / 1e12 / (fp32_peak * world),
                    "roofline_hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm, "unit": "GB/s",
                                     "frac": hbm_achieved / hbm if hbm_achieved else None,
                                     "note": "not the binding roof: compute-bound kernel over an L2-resident set"}})
    else:
        out["value"] = n * steps / (ms * 1e-3)
        e2e_value = n / e2e_s
        out["metric"] = out["unit"] = "body-steps/s"
        trav_ms = phases["force"]
        tot = cx.allreduce([bh_stats["nodes_visited_per_step"], bh_stats["interactions_per_step"], bh_stats["tree_nodes"],
                            bh_stats["pops"], bh_stats["pop_lanes"]], "sum")
        vis, inter, nodes, pops, pop_lanes = (int(v) for v in tot)
        hist = [int(v) for v in cx.allreduce(bh_hist, "sum")]
        trav_all = cx.allreduce([trav_ms], "sum")[0]
        trav_max = cx.allreduce([trav_ms], "max")[0]
        # Algorithmic (compulsory) bytes of one traversal launch: every tree node record once (16 B data + 4 B
        # child index), the sorted body coordinates + permutation once (12 B), one acceleration out (8 B).
        # SURVEY.md section 8d's per-visit figure (32 B x visited nodes) counts cache-served re-reads: it is
        # reported separately as visit_bytes -- node records stay in L1/L2 (ncu: DRAM < 1 %), so the walk is
        # bound by instruction issue, not by HBM, and a low HBM fraction is the expected reading.
        alg_bytes = (20 * nodes + 20 * n) / world
        achieved = alg_bytes / (trav_max * 1e-3) / 1e9 if trav_max > 0 else None
        hbm = pk.get("hbm_gbs", 6650.0)
        flops = (10 * vis + 6 * inter) / world
        out["roofline"] = {"bound": "hbm", "kernel": "bh_traverse_fast_kernel", "achieved": achieved, "peak": hbm, "unit": "GB/s",
                           "frac": achieved / hbm if achieved else None, **ncu_traffic("bh_traverse_fast_kernel", name, world),
                           "kernel_ms": trav_max, "kernel_share_of_step": trav_max / ms_per_step,
                           "visit_bytes_per_s_cache_served": 32 * vis / world / (trav_max * 1e-3) if trav_max > 0 else None,
                           # the same launch against the FP32 roofline: 10 flop per node visit (2 sub, mul + FMA for d^2,
                           # theta^2 d^2, compare, d^2+eps) + 6 per interaction (rcp, m*inv, 2 FMA)
                           "fp32_tflops": flops / (trav_max * 1e-3) / 1e12 if trav_max > 0 else None,
                           "fp32_frac": (flops / (trav_max * 1e-3) / 1e12 / fp32_peak) if trav_max > 0 else None,
                           "peak_source": ("MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in pk else "fallback 6.65 TB/s") +
                                          "; issue-bound walk over cache-resident nodes (DESIGN.md 4.3)"}
        out.update({"steps_per_s": steps / (ms * 1e-3), "interactions_per_s": inter * steps / (ms * 1e-3),
                    "fp32_frac": out["roofline"]["fp32_frac"],
                    # every popped (node block, lane mask) entry is evaluated by all 32 lanes of the warp: the useful share
                    "lane_efficiency": pop_lanes / (32.0 * pops) if pops else None,
                    "pops_per_step": pops, "lanes_per_pop_histogram": hist,
                    "interactions_per_step": inter, "nodes_visited_per_step": vis, "tree_nodes": nodes,
                    "walk_ms_max_over_ranks": trav_max, "walk_imbalance_max_over_mean": (trav_max / (trav_all / world)) if trav_all > 0 else None,
                    "theta": w["theta"], "sort_levels": sort_levels_timed})
        if world > 1:
            # per-rank view of the partitioned step: bodies in each rank's domain part and its phase times
            mine = {"rank": rank, "part_bodies": lib.counters()["bh_part_bodies"], **{k: round(v, 4) for k, v in phases.items()},
                    "xrank_boxes_bodies_trees_walks": [round(v, 4) for v in lib.phase_sub_ms("xrank")],
                    "com_merge_top": [round(v, 4) for v in lib.phase_sub_ms("com")]}
            allr = [None] * world
            cx.dist.all_gather_object(allr, mine)
            out["per_rank"] = allr
            out["ordering_point_ms"] = cx.allreduce([lib.dist_sync_test(200)], "max")[0]   # floor of one of the 4 per-step syncs
    out["e2e"] = {"value": e2e_value, "unit": out["unit"], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                  "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "what": e2e_what}
    if w["kind"] == "bh":
        out["e2e"]["steps_per_s"] = 1.0 / e2e_s
    if rank == 0 and with_cpu_baseline:
        st = host.numpy().copy()
        if w["kind"] == "allpairs":
            out["cpu_baseline"] = cpu_baseline_allpairs(st)
        elif w["kind"] == "bh":
            out["cpu_baseline"] = cpu_baseline_bh(st, w["theta"], seconds=6.0 if n > (1 << 20) else 3.0)
        else:
            out["cpu_baseline"] = None   # the reference has no 3-D path
    del host, host_out
    return out


def ncu_traffic(kernel, workload, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` summary (profiles/ncu_traffic.json, written by tools/ncu_summary.py).  An entry only counts
    while the kernel's source file is byte-identical to the one that was profiled (sha1 recorded with it)."""
    import hashlib

    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        ent = tab[kernel][f"{workload}@{world}"]
        src = os.path.join(ROOT, ent["source_file"])
        sha = hashlib.sha1(open(src, "rb").read()).hexdigest()
        if sha != ent["source_sha1"]:
            return {"traffic": None, "traffic_note": f"stale: {ent['source_file']} changed since {ent['capture']}"}
        return {"traffic": ent["dram_bytes_read"] + ent["dram_bytes_write"], "traffic_source": ent["capture"]}
    except Exception:
        return {"traffic": None}


def parity_check(cx):
    """Driver-visible correctness after the timed regions, at every N: 2 steps of 65,536 bodies through the C ABI
    (sharded over the N ranks, Barnes-Hut through the domain-partitioned path when N > 1) against rank 0's oracle --
    EXACT mode bit for bit, FAST mode within the stated tolerance -- and all ranks must read back identical state."""
    import hashlib

    lib, binding, world, rank = cx.lib, cx.binding, cx.world, cx.rank
    from rust_exp_b200 import ic

    n, steps, theta = 65536, 2, 0.5
    sa, sb = ic.plummer_2d(n, seed=21), ic.random_disk(n, seed=22)
    os.environ["NB_BH_PARTS_MIN_N"] = "0"   # N > 1: the FAST Barnes-Hut steps below take the partitioned multi-GPU path
    res = {}

    def gpu(mode, s, bh):
        lib.set_mode(mode)
        lib.set_particles(s)
        for _ in range(steps):
            lib.step_barnes_hut(theta, DT, 1) if bh else lib.step_brute_force(DT)
        g = lib.get_particles()
        lib.set_mode(binding.MODE_FAST)
        return g

    got = {(m, b): gpu(m, s, b) for m in (binding.MODE_EXACT, binding.MODE_FAST) for s, b in ((sa, False), (sb, True))}
    os.environ.pop("NB_BH_PARTS_MIN_N", None)
    digest = hashlib.sha1(b"".join(got[k].tobytes() for k in sorted(got))).digest()
    same = True
    if world > 1:
        allg = [None] * world
        cx.dist.all_gather_object(allg, digest)
        same = all(d == allg[0] for d in allg)
    ok = same
    if rank == 0:
        import oracle

        o = oracle.get()
        nc = os.cpu_count() or 1
        o.set_particles(sa)
        for _ in range(steps):
            o.step_brute_force(DT, nc)
        ra = o.get_particles()
        o.set_particles(sb)
        for _ in range(steps):
            o.step_barnes_hut(theta, DT, nc)
        rb = o.get_particles()

        def bits(a):
            return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)

        def rel(g, r):
            return float(np.abs(g[:, :2].astype(np.float64) - r[:, :2]).max() / np.abs(r[:, :2]).max())

        res = {"n_bodies": n, "steps": steps, "theta": theta,
               "exact_bitwise": bool(np.array_equal(bits(got[(binding.MODE_EXACT, False)]), bits(ra))),
               "exact_bh_bitwise": bool(np.array_equal(bits(got[(binding.MODE_EXACT, True)]), bits(rb))),
               "fast_rel_pos_err": rel(got[(binding.MODE_FAST, False)], ra),
               "fast_bh_rel_pos_err": rel(got[(binding.MODE_FAST, True)], rb),
               "tolerance": 1e-4, "ranks_agree": bool(same),
               "what": f"all-pairs (Plummer) and Barnes-Hut theta={theta} (disk), {steps} steps, sharded over {world} rank(s), vs the oracle on rank 0"}
        ok = (res["exact_bitwise"] and res["exact_bh_bitwise"] and res["fast_rel_pos_err"] <= 1e-4 and
              res["fast_bh_rel_pos_err"] <= 1e-4 and same)
        res["pass"] = bool(ok)
    ok = bool(cx.allreduce([0.0 if ok else 1.0], "max")[0] == 0.0)
    return res, ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("NB_WORKLOAD", DEFAULT_WORKLOAD), choices=sorted(WORKLOADS))
    ap.add_argument("--transport", default=os.environ.get("NB_TRANSPORT", "p2p_direct"),
                    choices=["p2p_direct", "p2p_gather", "nccl"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bh", action="store_true", help="skip the Barnes-Hut sub-records of the default line")
    ap.add_argument("--no-parity", action="store_true", help="skip the post-run parity check against the oracle")
    ap.add_argument("--bodies-per-thread", type=int, default=0)
    ap.add_argument("--waves", type=int, default=0)
    ap.add_argument("--ctas-per-sm", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--bh-steps", type=int, default=20)
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)
    if args.warmup < 3:
        args.warmup = 3  # timing rule: W >= 3
    # stdout carries exactly ONE JSON line: libraries (NCCL prints its version banner to fd 1) are sent to
    # stderr for the whole run and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    cx = Ctx(args)
    world, rank = cx.world, cx.rank
    m = measure(cx, args.workload, args.steps, args.warmup, args.e2e_steps, world == 1 and not args.no_cpu_baseline)
    line = {
        "metric": m["metric"], "value": m["value"], "unit": m["unit"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": m["workload"], "n_bodies": m["n_bodies"], "dt": DT, "mode": "fast",
                   "parallelism": f"index-sharded x{world}" + (f", transport={args.transport}" if world > 1 else ""),
                   "l2": "flushed between timed steps (256 MiB memset between the per-step CUDA-event pairs); within a step the position set is re-read from L2 by design",
                   "ic": m["ic"], "seed": m["seed"]},
        "e2e": m["e2e"], "gpu_launches": m["gpu_launches"], "clocks": m.get("clocks"), "roofline": m["roofline"],
        "phases_ms": m["phases_ms"],
    }
    for k in ("fp32_tflops", "fp32_frac_of_peak_all_gpus", "roofline_hbm", "steps_per_s", "interactions_per_s", "lane_efficiency",
              "lanes_per_pop_histogram", "pops_per_step", "interactions_per_step", "nodes_visited_per_step", "tree_nodes",
              "walk_imbalance_max_over_mean", "cpu_baseline", "per_rank", "sort_levels", "ordering_point_ms"):
        if k in m:
            line[k] = m[k]
    if w["kind"] == "bh":
        line["bh_steps_per_s"] = m["steps_per_s"]
    if args.workload == DEFAULT_WORKLOAD and not args.no_bh:
        # the metric also names "BH steps/s": Barnes-Hut sub-records measured in the same run (same timing rules)
        line["bh"] = {}
        for name in bh_sub_workloads(world):
            sub = measure(cx, name, args.bh_steps if name == "c4" else max(5, args.bh_steps // 2), 3, 3,
                          world == 1 and not args.no_cpu_baseline, sample_clocks=False)
            line["bh"][name] = sub
            line["gpu_launches"] += sub["gpu_launches"]
    if not args.no_parity:
        pc, ok = parity_check(cx)
        line["parity_check"] = pc
    else:
        ok = True
    if world > 1:
        cx.dist.barrier()
        cx.dist.destroy_process_group()
    sys.stdout.flush()
    if rank == 0:
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    os.close(real_stdout)
    return 0 if ok else 3


if __name__ == "__main__":
    sys.exit(main())
