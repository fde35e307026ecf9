#!/usr/bin/env python
"""bench.py -- agent-steps/s of BIDInet's per-timestep hierarchy update on B200.

Workload (BASELINE.json configs[1]/[2], SURVEY section 8 "C3"): the neo::Agent
hierarchy RunnerMain.cpp:141-154 builds (input 8x8, layers 8x8 -> 4x4, input feedback
radius 8, 144 columns), --agents-per-gpu independent agents per GPU, learning on,
driven by the headless RunnerMain loop (sine state vector + the four
clock inputs of RunnerMain.cpp:235-236 + the ten fed-back actions of :241-242,
reward = mean(action) - 0.5).  One "step" = one simStep of every agent.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm
  python bench.py --impl reference [--gpus N] --steps K --warmup W   the reference's CPU code

Our arm prints one JSON line with
  value        agent-steps/s, inputs resident in HBM (frames preloaded, action
               feedback and reward computed on the device), CUDA-event timed
  e2e          the same metric through the C ABI with HOST buffers: per step
               bidi_set_inputs (H2D) + bidi_step (H2D rewards) + bidi_get_predictions
               (D2H), the host doing the action feedback like RunnerMain does
  roofline     the dominant kernel's algorithmic bytes / its CUDA-event duration,
               against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline the reference's own CPU code (oracle/_ref) on this host's cores, on a
               bounded sample of the same workload
Multi-GPU (torchrun, one rank per GPU): agents are sharded over ranks, no data-path
collective; time = max over ranks; scaling = weak (agents per GPU fixed).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from bidinet_b200.sharding import (FB_IDX, base_frames, max_over_ranks as _max_over_ranks, rank_info,  # noqa: E402
                                   seed_base_for, whole_job_rate)


class ClockSampler:
    """nvidia-smi clocks and throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smmax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smmax.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_bytes_per_agent_step(eng, rho_c=1.0):
    """SURVEY 8d, NEO_AGENT: every persistent float read once (4 B), plus 4 B more for
    every float the learning rules write back: coder w+trace, lateral, thresholds,
    prediction weights, column lateral/q/action w+trace/threshold, and rho_c x the
    column feed-forward weights (rho_c = fraction of cells with state > 0; the survey
    measured ~1)."""
    L = eng.n_layers
    read = write = 0
    for l in range(L + 1):
        for name in ("FF_W", "REC_W", "LAT_W", "FF_TRACE", "REC_TRACE", "FB_W", "PRED_W", "HID_THRESHOLD",
                     "COL_LAT_W", "COL_Q_W", "COL_Q_TRACE", "COL_ACT_W", "COL_ACT_TRACE", "COL_THRESHOLD"):
            n = eng.array_len(name, l)
            read += n
            write += n
        n = eng.array_len("COL_FF_W", l)
        read += n
        write += int(rho_c * n)
    return 4 * (read + write)


def column_bytes_per_agent_step(eng, rho_c=1.0):
    """The share of algorithmic_bytes_per_agent_step the column kernel moves."""
    tot = 0
    for l in range(eng.n_layers + 1):
        for name in ("COL_LAT_W", "COL_Q_W", "COL_Q_TRACE", "COL_ACT_W", "COL_ACT_TRACE", "COL_THRESHOLD"):
            tot += 2 * eng.array_len(name, l)
        tot += int((1 + rho_c) * eng.array_len("COL_FF_W", l))
    return 4 * tot


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_rate(cfg, ld, cores, steps_per_thread, warmup=10):
    """agent-steps/s of the reference's CPU code: `cores` threads, one agent each, the
    headless RunnerMain loop (ref_run_runner releases the GIL: it is one C call)."""
    from oracle import pyoracle
    kind = "reference" if pyoracle.have_ref() else "port"
    cls = pyoracle.RefHierarchy if kind == "reference" else pyoracle.OracleHierarchy
    hs = [cls(cfg, ld, 1234 + a) for a in range(cores)]
    frames = [base_frames(64, 1, a)[:, 0, :] for a in range(cores)]
    for h, f in zip(hs, frames):
        h.run_runner(f, warmup, FB_IDX, FB_IDX)

    def work(k):
        hs[k].run_runner(frames[k], steps_per_thread, FB_IDX, FB_IDX)

    def timed():
        th = [threading.Thread(target=work, args=(k,)) for k in range(cores)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0
    return kind, hs, timed


def cpu_baseline_dict(value, cores, kind, sample):
    """per_thread: the figure that is comparable between boxes with different core counts."""
    return {"value": value, "unit": "agent-steps/s", "cores": cores, "per_thread": value / max(1, cores), "kind": kind, "sample": sample}


def hierarchy_cpu_rate(cfg, ld, frames, steps, warmup=3):
    """agent-steps/s of the reference's CPU code on ONE thread for a single hierarchy (C1 / C4): the reference is
    single threaded and a single agent cannot use more."""
    from oracle import pyoracle
    kind = "reference" if pyoracle.have_ref() else "port"
    cls = pyoracle.RefHierarchy if kind == "reference" else pyoracle.OracleHierarchy
    h = cls(cfg, ld, 1234)
    h.run(frames, warmup)

    def timed(n):
        t0 = time.perf_counter()
        h.run(frames, n)
        return time.perf_counter() - t0
    return kind, timed


def workload_label(args):
    from bidinet_b200 import workloads as wl
    if args.workload in wl.WORKLOADS:
        return "%s, %s (%s), 1 agent" % (wl.WORKLOADS[args.workload][2], args.variant, wl.CLASS_NAMES[wl.VARIANTS[args.variant]])
    if args.workload == "c2":
        return "C2: neo::Agent RunnerMain hierarchy (in 8x8, 8x8->4x4, 144 columns), headless runner loop, 1 agent"
    return "C3: neo::Agent RunnerMain hierarchy (in 8x8, 8x8->4x4, 144 columns), headless runner loop"


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    from bidinet_b200 import config as cf
    from bidinet_b200 import workloads as wl
    if args.workload in wl.WORKLOADS:  # one hierarchy, one thread
        build, frames_fn, _ = wl.WORKLOADS[args.workload]
        cfg, ld = build(wl.VARIANTS[args.variant])
        inner = args.ref_inner_steps
        kind, timed = hierarchy_cpu_rate(cfg, ld, frames_fn(), inner, warmup=min(3, inner))
        for _ in range(args.warmup):
            timed(inner)
        total = sum(timed(inner) for _ in range(args.steps))
        value = inner * args.steps / total
        print(json.dumps({
            "impl": "reference", "metric": "agent_steps_per_s", "value": value, "unit": "agent-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_label(args), "agents": 1, "learn": True},
            "cpu_baseline": cpu_baseline_dict(value, 1, kind, "1 thread x 1 agent x %d simSteps per bench step" % inner),
            "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return
    cfg, ld = cf.config_c2()
    cores = 1 if args.workload == "c2" else host_cores()
    inner = args.ref_inner_steps
    kind, hs, timed = cpu_reference_rate(cfg, ld, cores, inner)
    for _ in range(args.warmup):
        timed()
    dts = [timed() for _ in range(args.steps)]
    total = sum(dts)
    value = cores * inner * args.steps / total
    sample = "%d threads x 1 agent x %d simSteps per bench step" % (cores, inner)
    print(json.dumps({
        "impl": "reference", "metric": "agent_steps_per_s", "value": value, "unit": "agent-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_label(args), "agents": cores, "learn": True},
        "cpu_baseline": cpu_baseline_dict(value, cores, kind, sample),
        "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


from bidinet_b200.workloads import c5_frames  # noqa: E402


def c5_bytes_per_step(eng, rho=0.0):
    """SURVEY 8d, KWINNERS: 4 B per ff/rec/pred weight read once + 4*rho*(ff+rec) written back by
    the winners-only learning rule; rho is negligible at sparsity 0.01 (the prediction delta rule
    only touches units whose prediction was wrong)."""
    n_ff, n_rec, n_pred = (eng.array_len(k, 0) for k in ("FF_W", "REC_W", "PRED_W"))
    return 4 * (n_ff + n_rec + n_pred) + int(4 * rho * (n_ff + n_rec))


def run_c5(args, rank, world, local_rank):
    """--workload c5: ONE k-winners layer size x size over a size x size input, every radius 6
    (BASELINE config 5), cut into `world` row strips, one per GPU, halos over NCCL send/recv.
    Strong scaling: the layer is fixed, each GPU streams 1/world of the weights."""
    import torch
    import torch.distributed as dist
    from bidinet_b200 import Engine, nccl_unique_id
    from bidinet_b200 import config as cf

    size, K, W = args.c5_size, args.steps, max(3, args.warmup)
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg, ld = cf.config_c5(size)
    eng = Engine(cfg, ld, n_agents=1, device=local_rank)
    for kv in filter(None, os.environ.get("BIDI_OPTS", "").split(",")):  # experiments only
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    if world > 1:
        eng.strip_configure(rank, world)
    eng.init_random(1234)  # before linking: once linked over peer memory, the neighbours write into this part's state
    frames = c5_frames(size)
    T = len(frames)
    eng.load_frames(frames[:, None, :])
    if world > 1:
        if args.c5_transport == "nccl":
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, 0)
            eng.strip_link_nccl(bytes(uid.cpu().numpy().tobytes()))
        else:  # peer memory: gather every part's CUDA IPC handles (the gather is the barrier the linking needs)
            mine = torch.frombuffer(bytearray(eng.strip_ipc_export()), dtype=torch.uint8).cuda()
            allh = [torch.zeros(128, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(allh, mine)
            eng.strip_link_ipc([bytes(t.cpu().numpy().tobytes()) for t in allh])
    barrier()
    t = 0
    for _ in range(W):
        eng.step_frame(t)
        t += 1
    eng.sync()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = eng.launch_count
    eng.timer_start()
    for _ in range(K):
        eng.step_frame(t)
        t += 1
    ms = eng.timer_stop_ms()
    eng.sync()
    barrier()
    launches = eng.launch_count - launches0
    ms = _max_over_ranks(ms, dist if world > 1 else None, "cuda")
    value = K / (ms * 1e-3)

    # e2e: the whole frame goes to every part from host memory, the owned prediction rows come back -- through the
    # engine's page-locked staging buffers used in place (bidi_host_inputs / bidi_host_predictions), as in the other
    # workloads: the host<->device copies are made every step, a second host-side copy of the 4 MB frame is not
    # (a part of a split layer passes the caller's frame as it is: bidi_set_inputs then stages only the rows its windows reach)
    x, pred = eng.host_inputs(), eng.host_predictions()

    def e2e_step(tt):
        if world == 1:
            np.copyto(x[0], frames[tt % T])
            eng.set_inputs(x)
        else:
            eng.set_inputs(frames[tt % T])
        eng.step()
        eng.get_predictions(out=pred)

    for _ in range(W):
        e2e_step(t)
        t += 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        e2e_step(t)
        t += 1
    eng.sync()
    e2e_s = _max_over_ranks(time.perf_counter() - t0, dist if world > 1 else None, "cuda")
    barrier()
    clk = clocks.stop() if rank == 0 else None
    step_bytes = c5_bytes_per_step(eng)
    halo = eng.strip_halo_floats * 4 if world > 1 else 0
    # input rows a part uploads per step: the rows its windows reach (all of them when unsplit)
    h2d_rows = size if world == 1 else (max(eng.spans["VIS_RECON"][1], eng.spans["OWN_VIS"][1]) - min(eng.spans["VIS_RECON"][0], eng.spans["OWN_VIS"][0]))
    dev_bytes = eng.device_bytes
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = hbm_peak()
    per_gpu = step_bytes / world
    achieved = per_gpu / (ms / K * 1e-3) / 1e9
    n_units = size * size
    print(json.dumps({
        "metric": "agent_steps_per_s", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C5: one k-winners layer %dx%d over a %dx%d input, every radius 6, 1 agent" % (size, size, size, size),
                   "units": n_units, "unit_updates_per_s": value * n_units, "learn": True,
                   "parallelism": ("%d row strips, one per GPU, halo rows %s" % (world, "pushed into peer memory over NVLink by one kernel per exchange point"
                                   if args.c5_transport == "ipc" else "over NCCL send/recv")) if world > 1 else "one GPU, unsplit",
                   "halo_bytes_sent_per_gpu_step": halo, "l2": "weights streamed per step: %.0f MB per GPU" % (per_gpu / 1e6),
                   "state_bytes_per_gpu": dev_bytes},
        "e2e": {"value": K / e2e_s, "unit": "agent-steps/s", "ms_per_step": 1e3 * e2e_s / K,
                "h2d_bytes_per_step": int(4 * n_units * h2d_rows / size), "d2h_bytes_per_step": int(4 * n_units / world)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "whole step (6 phase kernels + halo exchanges)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "bytes_per_launch": per_gpu, "avg_launch_ms": ms / K, "launches_per_step": 1},
        "clocks": clk,
    }))


def run_hierarchy(args, rank, world, local_rank):
    """--workload c1 | c4 [--variant kwinners|iterative]: ONE hierarchy (sdr::PredictiveRSDR or sdr::IPredictiveRSDR)
    stepping on a repeating synthetic sequence, learning on.  These single-agent configurations fit the L2 and are
    bound by the latency of their dependent launches, not by HBM: the roofline object is reported for
    completeness, the figures to read are value / e2e against the one-thread cpu_baseline.  Multi-GPU: replicas only
    (every rank steps its own copy; no collective)."""
    import torch
    from bidinet_b200 import Engine
    from bidinet_b200 import workloads as wl

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    build, frames_fn, _ = wl.WORKLOADS[args.workload]
    variant = wl.VARIANTS[args.variant]
    cfg, ld = build(variant)
    frames = frames_fn()
    T, n_in = frames.shape
    K, W = args.steps, max(3, args.warmup)
    eng = Engine(cfg, ld, n_agents=1, device=local_rank, seed=1234)
    for kv in filter(None, os.environ.get("BIDI_OPTS", "").split(",")):
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    eng.load_frames(frames[:, None, :])
    t = 0
    for _ in range(W + args.burn_in):
        eng.step_frame(t)
        t += 1
    eng.sync()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = eng.launch_count
    eng.timer_start()
    for _ in range(K):
        eng.step_frame(t)
        t += 1
    ms = eng.timer_stop_ms()
    eng.sync()
    barrier()
    launches = eng.launch_count - launches0
    ms = _max_over_ranks(ms, dist, "cuda")
    value = whole_job_rate(1, world, K, ms * 1e-3)

    # e2e: the frame goes up from (pinned) host memory, the prediction comes back, every step
    x, pred = eng.host_inputs(), eng.host_predictions()
    for _ in range(W):
        np.copyto(x[0], frames[t % T])
        eng.set_inputs(x)
        eng.step()
        eng.get_predictions(out=pred)
        t += 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        np.copyto(x[0], frames[t % T])
        eng.set_inputs(x)
        eng.step()
        eng.get_predictions(out=pred)
        t += 1
    eng.sync()
    e2e_s = _max_over_ranks(time.perf_counter() - t0, dist, "cuda")
    barrier()
    clk = clocks.stop() if rank == 0 else None
    step_bytes = wl.hierarchy_bytes_per_step(eng, variant)
    dev_bytes = eng.device_bytes
    if dist is not None:
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = hbm_peak()
    achieved = step_bytes / (ms / K * 1e-3) / 1e9
    out = {
        "metric": "agent_steps_per_s", "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_label(args), "agents": world, "learn": True,
                   "parallelism": "replicas only (one hierarchy per GPU, no collective)" if world > 1 else "one GPU",
                   "l2": "working set %.1f MB fits the 126 MB L2: bound by the latency of %d dependent launches per step, not by HBM"
                         % (dev_bytes / 1e6, launches // max(1, K)),
                   "state_bytes_per_gpu": dev_bytes, "burn_in_steps": args.burn_in},
        "e2e": {"value": whole_job_rate(1, world, K, e2e_s), "unit": "agent-steps/s", "ms_per_step": 1e3 * e2e_s / K,
                "h2d_bytes_per_step": int(4 * n_in), "d2h_bytes_per_step": int(4 * n_in)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "whole step (%d launches)" % (launches // max(1, K)), "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "bytes_per_launch": step_bytes, "avg_launch_ms": ms / K, "launches_per_step": 1,
                     "note": "algorithmic bytes per step (SURVEY 8d) over the step time; the state is L2-resident"},
        "clocks": clk,
    }
    if not args.no_cpu_baseline:
        kind, timed = hierarchy_cpu_rate(cfg, ld, frames, 2)
        n = max(3, min(20000, int(args.cpu_baseline_seconds / max(1e-6, timed(2) / 2))))
        dt = timed(n)
        out["cpu_baseline"] = cpu_baseline_dict(n / dt, 1, kind, "1 thread x 1 agent x %d simSteps of the same sequence (%.1f s)" % (n, dt))
    print(json.dumps(out))


def run_c3_single_process(args):
    """--single-process: C3 over --gpus N devices WITHOUT torchrun -- one host thread drives all devices through
    bidi_create_sharded (agents partitioned contiguously, one engine per device, no collective).  Same metric;
    time = the slowest shard's CUDA-event time."""
    import torch
    from bidinet_b200 import ShardedEngine
    from bidinet_b200 import config as cf
    N, A, K, W = args.gpus, args.agents_per_gpu, args.steps, max(3, args.warmup)
    if torch.cuda.device_count() < N:
        raise SystemExit("bench.py --single-process: %d devices asked, %d visible" % (N, torch.cuda.device_count()))
    cfg, ld = cf.config_c2()
    eng = ShardedEngine(cfg, ld, N * A, devices=list(range(N)), seed=1234)
    n_in = eng.n_in
    T = min(W + K, 64)
    frames = base_frames(T, N * A, 0, n_in)
    eng.load_frames(frames)
    eng.set_feedback(FB_IDX, FB_IDX, 0.5, 0.5, 0.05, True)
    t = 0
    for _ in range(args.burn_in + W):
        eng.step_frame(t)
        t += 1
    eng.sync()
    clocks = ClockSampler(0)
    clocks.start()
    launches0 = eng.launch_count
    eng.timer_start()
    for _ in range(K):
        eng.step_frame(t)
        t += 1
    ms = eng.timer_stop_ms()
    eng.sync()
    launches = eng.launch_count - launches0
    value = N * A * K / (ms * 1e-3)
    # e2e: host buffers of the WHOLE population through the sharded calls
    import ctypes
    from bidinet_b200.build import build_runner_env
    env_lib = ctypes.CDLL(build_runner_env())
    fp = ctypes.POINTER(ctypes.c_float)
    env_lib.runner_env_step.argtypes = [ctypes.c_int] * 4 + [fp] * 4
    env_lib.runner_env_step.restype = None
    rng = np.random.RandomState(0)
    env_noise = rng.normal(0.0, 0.05, (W + K, N * A, len(FB_IDX))).astype(np.float32)
    pred = eng.get_predictions()
    x = frames[t % T].copy()
    rewards = np.zeros(N * A, np.float32)
    fb0, nfb = int(FB_IDX[0]), len(FB_IDX)

    def host_step(i, tt):
        nonlocal pred
        env_lib.runner_env_step(N * A, n_in, fb0, nfb, pred.ctypes.data_as(fp), env_noise[i].ctypes.data_as(fp), x.ctypes.data_as(fp),
                                rewards.ctypes.data_as(fp))
        eng.set_inputs(x)
        eng.step(rewards=rewards)
        np.copyto(x, frames[(tt + 1) % T])
        pred = eng.get_predictions()

    for i in range(W):
        host_step(i, t)
        t += 1
    t0 = time.perf_counter()
    for i in range(K):
        host_step(W + i, t)
        t += 1
    eng.sync()
    e2e_s = time.perf_counter() - t0
    clk = clocks.stop()
    step_bytes = algorithmic_bytes_per_agent_step(eng.shards[0])
    peak, peak_src = hbm_peak()
    achieved = step_bytes * A / (ms / K * 1e-3) / 1e9   # per GPU
    print(json.dumps({
        "metric": "agent_steps_per_s", "value": value, "unit": "agent-steps/s", "n_gpus": N, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_label(args), "agents_per_gpu": A, "agents": N * A, "learn": True,
                   "parallelism": "ONE process, %d devices through bidi_create_sharded: agents partitioned contiguously, no collective" % N,
                   "l2": "per-step working set %.1f GB per GPU exceeds the 126 MB L2" % (eng.device_bytes / N / 1e9),
                   "exploration_noise": "engine counter-based generator", "burn_in_steps": args.burn_in},
        "e2e": {"value": N * A * K / e2e_s, "unit": "agent-steps/s", "ms_per_step": 1e3 * e2e_s / K,
                "h2d_bytes_per_step": int(N * A * (n_in + 1) * 4), "d2h_bytes_per_step": int(N * A * n_in * 4)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "whole step, per GPU", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "bytes_per_launch": step_bytes * A,
                     "avg_launch_ms": ms / K, "launches_per_step": 1},
        "clocks": clk,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 30 for c3/c5; enough for a ~1 s timed region for the single-agent workloads)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--agents-per-gpu", type=int, default=4096)
    ap.add_argument("--ref-inner-steps", type=int, default=200)
    ap.add_argument("--cpu-baseline-steps", type=int, default=1500)
    ap.add_argument("--profile-kernel", default="columns")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--burn-in", type=int, default=400,
                    help="untimed steps before the warm-up: by then every column cell spikes once per step and every "
                         "learning rule does its full work (a freshly initialised network does less)")
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--variant", default="kwinners", choices=["kwinners", "iterative"], help="c1 / c4: which reference class")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=12.0, help="c1 / c4: CPU time of the one-thread baseline sample")
    ap.add_argument("--e2e-groups", type=int, default=3,
                    help="c3: agent groups the end-to-end loop steps in turn (host I/O of one overlaps the GPU step of the other)")
    ap.add_argument("--single-process", action="store_true",
                    help="c3 over --gpus N devices from ONE process (bidi_create_sharded) instead of one torchrun rank per GPU")
    ap.add_argument("--c5-size", type=int, default=1024)
    ap.add_argument("--c5-transport", default="ipc", choices=["ipc", "nccl"])
    args = ap.parse_args()
    rank, world, local_rank = rank_info()
    if args.steps is None:  # a timed region of about a second where a step is a fraction of a millisecond
        args.steps = {"c1": 4000 if args.variant == "kwinners" else 1000, "c2": 2000,
                      "c4": 2000 if args.variant == "kwinners" else 200, "c5": 1000}.get(args.workload, 30)
        if args.impl == "reference":
            args.steps = 5
    if args.workload in ("c1", "c4") and args.burn_in == 400:
        args.burn_in = 50
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "c5":
        run_c5(args, rank, world, local_rank)
        return
    if args.workload in ("c1", "c4"):
        run_hierarchy(args, rank, world, local_rank)
        return
    if args.workload == "c2":  # the RunnerMain hierarchy with ONE agent (per GPU: replicas only)
        args.agents_per_gpu = 1
    if args.single_process and world == 1 and args.workload == "c3":
        run_c3_single_process(args)
        return

    import torch
    from bidinet_b200 import Engine
    from bidinet_b200 import config as cf

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        return _max_over_ranks(x, dist, "cuda")

    A, K, W = args.agents_per_gpu, args.steps, max(3, args.warmup)
    cfg, ld = cf.config_c2()
    eng = Engine(cfg, ld, n_agents=A, device=local_rank)
    for kv in filter(None, os.environ.get("BIDI_OPTS", "").split(",")):  # experiments only, e.g. BIDI_OPTS=early_learn=0
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    eng.init_random(seed_base_for(rank, A))  # global agent g <- std::mt19937(1234 + g)
    eng.set_noise_stream(0x5DEECE66D, rank * A)  # every global agent explores on its own noise stream
    n_in = eng.n_in
    T = min(W + K, 64)
    frames = base_frames(T, A, rank * A, n_in)
    eng.load_frames(frames)
    eng.set_feedback(FB_IDX, FB_IDX, 0.5, 0.5, 0.05, True)

    # ---- value: device-resident inputs, CUDA events on the engine's stream
    t = 0
    for _ in range(args.burn_in):  # part of the set-up, like the weight initialisation
        eng.step_frame(t)
        t += 1
    eng.sync()
    for _ in range(W):
        eng.step_frame(t)
        t += 1
    eng.sync()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = eng.launch_count
    ncu_range = os.environ.get("BIDI_NCU_RANGE") == "1"   # ncu --profile-from-start off: profile the timed steps only
    if ncu_range:
        torch.cuda.cudart().cudaProfilerStart()
    eng.timer_start()
    for _ in range(K):
        eng.step_frame(t)
        t += 1
    ms = eng.timer_stop_ms()
    eng.sync()
    if ncu_range:
        torch.cuda.cudart().cudaProfilerStop()
    barrier()
    launches = eng.launch_count - launches0
    ms = max_over_ranks(ms)
    value = whole_job_rate(A, world, K, ms * 1e-3)

    # ---- e2e: host buffers through the C ABI, host<->device copies inside the timed region.
    # The population is driven as G agent groups of one sharded handle on this device (bidi_create_sharded with the
    # device listed G times: agents never interact, RunnerMain.cpp:100-106), stepped in turn: while the GPU runs one
    # group's step the host reads the other group's results, runs its environment and uploads its next inputs.  Every
    # step of every agent still has its host->device copy of the inputs and its device->host read of the predictions.
    import ctypes
    from bidinet_b200 import ShardedEngine
    from bidinet_b200.build import build_runner_env
    G = max(1, min(args.e2e_groups, A))
    pipe = ShardedEngine(cfg, ld, A, devices=[local_rank] * G)
    pipe.init_random(seed_base_for(rank, A))
    for g, sh in enumerate(pipe.shards):
        sh.set_noise_stream(0x5DEECE66D, rank * A + pipe.first[g])
    pipe.load_frames(frames)
    pipe.set_feedback(FB_IDX, FB_IDX, 0.5, 0.5, 0.05, True)
    tp = 0
    for _ in range(args.burn_in):     # the same regime as the device-resident measurement
        pipe.step_frame(tp)
        tp += 1
    pipe.sync()
    rng = np.random.RandomState(rank)
    # the synthetic environment's noise draws (RunnerMain.cpp:241-242) are data, made before the clock starts
    env_noise = rng.normal(0.0, 0.05, (W + K, A, len(FB_IDX))).astype(np.float32)
    fb0, fb1 = int(FB_IDX[0]), int(FB_IDX[-1]) + 1     # the action inputs are one contiguous run of cells
    assert np.array_equal(FB_IDX, np.arange(fb0, fb1))
    env_lib = ctypes.CDLL(build_runner_env())  # tools/runner_env.c: the environment in compiled code, as in the reference
    fp = ctypes.POINTER(ctypes.c_float)
    env_lib.runner_env_step.argtypes = [ctypes.c_int] * 4 + [fp] * 4
    env_lib.runner_env_step.restype = None
    # the engines' page-locked staging buffers, used in place (bidi_host_inputs / bidi_host_predictions): the
    # host<->device copies are still made every step, the extra host-side staging memcpy is not
    grp = []
    for g, sh in enumerate(pipe.shards):
        lo, n = pipe.first[g], pipe.count[g]
        pred = sh.get_predictions(out=sh.host_predictions())
        x = sh.host_inputs()
        np.copyto(x, frames[tp % T][lo:lo + n])
        grp.append({"sh": sh, "lo": lo, "n": n, "pred": pred, "x": x, "rew": np.zeros(n, dtype=np.float32)})

    def host_step(i, tt):
        """the caller's side of RunnerMain.cpp:235-251, group by group: read the group's predictions (device -> host,
        waits for its previous step), state inputs + fed-back noisy actions + reward, upload, enqueue the step."""
        for q in grp:
            sh, lo, n = q["sh"], q["lo"], q["n"]
            sh.get_predictions(out=q["pred"])
            nz = np.ascontiguousarray(env_noise[i][lo:lo + n])
            env_lib.runner_env_step(n, n_in, fb0, fb1 - fb0, q["pred"].ctypes.data_as(fp), nz.ctypes.data_as(fp),
                                    q["x"].ctypes.data_as(fp), q["rew"].ctypes.data_as(fp))
            sh.set_inputs(q["x"])
            sh.step(rewards=q["rew"])        # asynchronous: returns once the step is enqueued
            sh.wait_inputs()                 # the host->device copy has read x
            np.copyto(q["x"], frames[(tt + 1) % T][lo:lo + n])

    for i in range(W):
        host_step(i, tp)
        tp += 1
    pipe.sync()
    barrier()
    t0 = time.perf_counter()
    for i in range(K):
        host_step(W + i, tp)
        tp += 1
    for q in grp:                            # the last step's results, read like every other step's
        q["sh"].get_predictions(out=q["pred"])
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = whole_job_rate(A, world, K, e2e_s)
    e2e_launches = pipe.launch_count if hasattr(pipe, "launch_count") else None
    pipe.close()
    if rank == 0:
        clk = clocks.stop()

    # ---- roofline: the dominant kernel's own duration, CUDA events inside the step graph
    kern = args.profile_kernel
    eng.profile_select(kern)
    for _ in range(3):
        eng.step_frame(t)
        t += 1
    eng.profile_collect()
    k_ms0, k_n0 = eng.profile_collect()
    for _ in range(K):
        eng.step_frame(t)
        t += 1
        eng.profile_collect()
    k_ms, k_n = eng.profile_collect()
    k_ms, k_n = k_ms - k_ms0, k_n - k_n0
    eng.profile_select("")
    step_bytes = algorithmic_bytes_per_agent_step(eng)
    kern_bytes_step = column_bytes_per_agent_step(eng) if kern == "columns" else step_bytes
    launches_per_step = max(1, k_n // max(1, K))
    avg_launch_s = (k_ms * 1e-3) / max(1, k_n)
    bytes_per_launch = kern_bytes_step * A / launches_per_step
    achieved = bytes_per_launch / avg_launch_s / 1e9 if avg_launch_s > 0 else 0.0
    peak, peak_src = hbm_peak()
    traffic, traffic_stale = None, None
    try:  # measured DRAM bytes of this kernel (ncu dram__bytes_read.sum + dram__bytes_write.sum), per launch like `achieved`
        tj = json.load(open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)))
        tr = tj[kern]
        traffic = tr["dram_bytes_per_agent_step"] * A / tr["launches_per_step"]
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from steady_launches import kernel_source_sha
        traffic_stale = tj.get("kernel_source_sha") != kernel_source_sha()  # the capture was taken on other kernel sources
    except Exception:
        pass

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = 1 if args.workload == "c2" else host_cores()
        kind, hs, timed = cpu_reference_rate(cfg, ld, cores, args.cpu_baseline_steps)
        dt = timed()
        cpu = cpu_baseline_dict(cores * args.cpu_baseline_steps / dt, cores, kind,
                                "%d threads x 1 agent x %d simSteps of the same runner loop (%.1f s)" % (cores, args.cpu_baseline_steps, dt))

    out = {
        "metric": "agent_steps_per_s", "value": value, "unit": "agent-steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_label(args),
                   "agents_per_gpu": A, "agents": world * A, "learn": True, "parallelism": "agents sharded over ranks, no collective",
                   "l2": ("per-step working set %.1f GB per GPU exceeds the 126 MB L2" % (eng.device_bytes / 1e9)) if A > 64 else
                         ("working set %.1f MB fits the L2: launch-latency bound" % (eng.device_bytes / 1e6)),
                   "exploration_noise": "engine counter-based generator", "state_bytes_per_gpu": eng.device_bytes,
                   "burn_in_steps": args.burn_in},
        "e2e": {"value": e2e_value, "unit": "agent-steps/s", "ms_per_step": 1e3 * e2e_s / K,
                "h2d_bytes_per_step": int(A * n_in * 4 + A * 4), "d2h_bytes_per_step": int(A * n_in * 4),
                "agent_groups": G,
                "note": "every step copies every agent's inputs host->device and its predictions device->host; the agents are "
                        "stepped as %d groups in turn so that one group's host work overlaps the other's GPU step" % G},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": kern, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "traffic_source": "profiles/%s (ncu launch list of 4096 agents; scaled by agents)" % TRAFFIC_FILE,
                     "traffic_stale": traffic_stale,
                     "bytes_per_launch": bytes_per_launch, "avg_launch_ms": 1e3 * avg_launch_s,
                     "launches_per_step": int(launches_per_step),
                     "whole_step": {"algorithmic_bytes_per_agent_step": step_bytes,
                                    "achieved": step_bytes * A / (ms / K * 1e-3) / 1e9,
                                    "frac": step_bytes * A / (ms / K * 1e-3) / 1e9 / peak}},
        "clocks": clk,
    }
    if cpu:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


TRAFFIC_FILE = "r2_traffic.json"


def _emit_one_line():
    """The contract is ONE JSON line on stdout, but native libraries write banners to file descriptor 1
    (NCCL prints its version there when NCCL_DEBUG is set in the environment).  Everything written to fd 1
    while the bench runs goes to stderr; only the lines this script prints reach the real stdout."""
    import io
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    buf = io.StringIO()
    py_stdout, sys.stdout = sys.stdout, buf
    try:
        main()
    finally:
        sys.stdout = py_stdout
        os.write(real, buf.getvalue().encode())
        os.close(real)


if __name__ == "__main__":
    _emit_one_line()
