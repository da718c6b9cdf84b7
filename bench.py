#!/usr/bin/env python
"""bench.py -- particle-force evaluations per second of the rebx_additional_forces hot path.

Default workload (BASELINE.json configs[1]): synthetic debris disk, 10^7 test particles per
GPU, radiation_forces with Poynting-Robertson drag, beta per particle.  One "step" is one
force evaluation (one dispatch) over the whole ensemble.

  python bench.py --gpus 1 --steps K --warmup W                  # this build
  torchrun ... bench.py --gpus N ...                             # one rank per GPU, weak scaling
  python bench.py --impl reference ...                           # the reference's CPU loop

JSON keys (one line, rank 0): value = evals/s with state resident in HBM (CUDA events, max over
ranks); e2e = evals/s through the REBOUND-facing call sim->additional_forces(sim) on HOST
particle arrays, host<->device copies inside the timed region; roofline = the sweep kernel's
algorithmic HBM bytes / its event-timed duration vs the measured copy bandwidth;
cpu_baseline = the reference's own C loop (oracle/_ref) on one host core, bounded sample (+ parallel_upper_bound:
all host cores running independent single-threaded dispatches, labelled not-the-reference);
resident_step / resident_wh_step = whole integrator steps with the state kept in HBM (fused drift-kick-drift
sweep; Wisdom-Holman with Kepler drifts); --coordinates JACOBI|BARYCENTRIC runs the damping trio in the
centre-of-mass frames of rebx_com_force (sharded across ranks with prefix / suffix exchanges).
"""
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

METRIC = "particle-force evals/sec"
WORKLOADS = {
    # name: (builder, forces in ADD order, default particles per GPU)
    "debris_disk": ("debris_disk", ["radiation_forces"], 10_000_000),
    "circumplanetary_ring": ("circumplanetary_ring", ["gravitational_harmonics", "gr_potential"], 10_000_000),
    "asteroid_ensemble": ("asteroid_ensemble", ["yarkovsky_effect", "gr_potential"], 1_000_000),
    "planetesimal_disk": ("planetesimal_disk", ["modify_orbits_forces", "gas_damping_timescale", "type_I_migration"], 1_000_000),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="debris_disk", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="test particles per GPU (0 = the config's size)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-facing leg (0 = min(steps, 10))")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=2_000_000)
    ap.add_argument("--no-parallel-bound", action="store_true", help="skip the all-cores upper bound of the CPU baseline")
    ap.add_argument("--math", default="exact", choices=["exact", "tolerance"],
                    help="exact: bit-identical to the reference where + - * / sqrt (default); tolerance: kernels_tol.cu (FMA, shared reciprocals)")
    ap.add_argument("--coordinates", default="PARTICLE", choices=["PARTICLE", "JACOBI", "BARYCENTRIC"],
                    help="reference frame of the rebx_com_force effects (planetesimal_disk only)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines, self.windows = index, None, [], []

    def window(self, t0, t1):
        """Marks [t0, t1] (time.time()) as a timed region; the summary prefers samples taken inside one."""
        self.windows.append((t0, t1))

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # a sample printed at time t describes the ~20 ms before it
        inside = [ln for t, ln in self.lines if any(t0 <= t <= t1 + 0.05 for t0, t1 in self.windows)]
        scope = "timed regions" if inside else "whole run (timed regions shorter than the 20 ms sampling period)"
        for line in (inside or [ln for _, ln in self.lines]):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm), "scope": scope}


def cpu_reference_leg(workload, add_order, sample_n, reps, coordinates="PARTICLE"):
    """The reference's own C implementation (oracle/_ref, kind "reference") or, if that library
    was never built, the C restatement (kind "port"), one host thread, bounded sample."""
    from reboundx_b200 import workloads
    from oracle import oracle
    w = getattr(workloads, workload)(sample_n)
    n = w["x"].shape[0]
    nf = len(add_order)
    if oracle.have_ref():
        _, t = oracle.ref_apply(w, add_order, None, coordinates, reps=reps)
        kind = "reference"
    else:
        best = float("inf")
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            oracle.port_apply(w, add_order[::-1], None, coordinates)
            best = min(best, time.perf_counter() - t0)
        t, kind = best, "port"
    return {"value": (n - 1)*nf/t, "unit": METRIC, "cores": 1, "kind": kind,
            "sample": f"{workload}: {n - 1} particles x {nf} force(s), best of {reps} dispatches of rebx_additional_forces, 1 thread (the reference dispatch is single-threaded)",
            "ms_per_dispatch": t*1e3}


def cpu_parallel_upper_bound(workload, per_worker):
    """What the box's host cores could do if the reference's dispatch were run as `nproc` independent shards at
    once (it is not: REBOUNDx's force path is single-threaded).  Reported next to the 1-core
    reference figure as an upper bound for a hypothetical embarrassingly-parallel CPU version -- the number the
    PCIe-bound end-to-end figure has to be read against; the resident figures are three orders above both."""
    n = len(os.sched_getaffinity(0))
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload, "--steps", "5",
           "--cpu-sample", str(per_worker), "--no-parallel-bound"]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    t0 = time.perf_counter()
    procs = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env) for _ in range(n)]
    vals = []
    for p in procs:
        out, _ = p.communicate(timeout=300)
        for line in out.splitlines():
            if line.startswith("{"):
                vals.append(json.loads(line)["value"])
    if len(vals) != n:
        return None
    return {"value": float(sum(vals)), "unit": "evals/s", "cores": n, "kind": "not the reference: %d concurrent single-threaded dispatches, each on its own sample" % n,
            "sample": f"{per_worker} particles per worker", "seconds": round(time.perf_counter() - t0, 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    builder, add_order, n_default = WORKLOADS[args.workload]
    sample = min(args.cpu_sample if args.coordinates == "PARTICLE" else 20000, args.n or n_default)   # the COM frames are O(N^2) in the reference
    steps = max(1, args.steps)
    base = cpu_reference_leg(builder, add_order, sample, steps, args.coordinates)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_dispatch"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # the same workload naming as the GPU arm's line; each step here evaluates a bounded sample of it
            "config": {"workload": f"{args.workload}: {args.n or n_default} test particles per GPU, {'+'.join(add_order)}" + (f", REBX_COORDINATES_{args.coordinates}" if args.coordinates != "PARTICLE" else ""),
                       "particles_per_gpu": args.n or n_default, "forces": add_order,
                       "sample": f"each step = one dispatch of the reference's rebx_additional_forces over a bounded sample of {sample} test particles of this workload, 1 host thread (the reference is single-threaded)"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["cpu_baseline"]["unit"] = "evals/s"
    if not args.no_parallel_bound and args.coordinates == "PARTICLE":
        line["cpu_baseline"]["parallel_upper_bound"] = cpu_parallel_upper_bound(args.workload, min(500_000, sample))
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index):
    """Pins this rank to the CPUs NVML reports as local to GPU `index` BEFORE the host particle arrays are
    allocated, so that their pages (first touch) and the DMA of the host-facing leg stay on the GPU's own
    NUMA node.  Returns a short description for the JSON line (None when NVML has nothing to say)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63)//64)
        cpus = {64*w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return f"rank bound to the {len(cpus)} CPUs local to GPU {index}"
    except Exception:
        pass
    return None


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    import torch
    import torch.distributed as dist
    from reboundx_b200 import device, scenarios, workloads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the force path has no CPU fallback")
    torch.cuda.set_device(local)
    os.environ["REBX_B200_DEVICE"] = str(local)
    # stdout carries ONE JSON line: anything a library prints to fd 1 meanwhile (NCCL's version banner under
    # NCCL_DEBUG=VERSION, ...) is sent to stderr, and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = f"cuda:{local}"
    builder, add_order, n_default = WORKLOADS[args.workload]
    n = args.n or n_default
    exec_order = add_order[::-1]
    nf = len(add_order)
    com = args.coordinates != "PARTICLE"
    if com and args.workload != "planetesimal_disk":
        raise SystemExit("--coordinates applies to the rebx_com_force effects of --workload planetesimal_disk")
    clocks = ClockSampler(local)
    clocks.__enter__()          # sampled from here to the end of the e2e leg; the timed regions are marked as windows

    # Weak scaling: every rank owns n test particles of one global ensemble of world*n; the
    # central body (global index 0) is replicated and broadcast from rank 0 every evaluation.
    w = getattr(workloads, builder)(n, seed=3 + rank)
    lo = 0 if rank == 0 else 1
    shard = device.Shard(w, exec_order, dev, lo=lo, index_base=(0 if rank == 0 else 1 + rank*n), coordinates=args.coordinates, math=args.math)
    if args.math == "tolerance":
        os.environ["REBX_B200_MATH"] = "tolerance"          # the host-facing leg reads it when its device context is created
    n_local = shard.n - (1 if rank == 0 else 0)          # the body itself is not an evaluation

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    group = None
    for _ in range(max(3, args.warmup)):
        shard.evaluate(group)
    barrier()

    # ---- resident leg: K evaluations, CUDA events on the launching stream, max over ranks -------
    launches0 = shard.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        shard.evaluate(group)
    e1.record()
    barrier()
    gpu_launches = shard.launches - launches0                    # launches of the timed `value` region only
    # ---- roofline of the dominant kernel (the fused sweep), timed alone per launch -----------
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(args.steps):
        shard.launch_local()
    k1.record()
    torch.cuda.synchronize()
    clocks.window(w0, time.time())
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item())/args.steps
    value = world*n*nf/(ms_step*1e-3)
    kernel_ms = k0.elapsed_time(k1)/args.steps
    peak, peak_src = peaks()
    achieved = shard.algorithmic_bytes/(kernel_ms*1e-3)/1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved/peak,
                "traffic": None, "kernel": ("pass_kernel(%s)" if not com else args.coordinates.lower() + " pipeline: moments, scans, sweep(%s), apply") % ",".join(exec_order),
                "algorithmic_bytes_per_particle": shard.algorithmic_bytes//shard.n,
                "kernel_ms": kernel_ms, "peak_source": peak_src}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof) and n == n_default and not com:       # ncu dram__bytes_read+write per launch, captured at the config's size
        with open(prof) as f:
            captured = json.load(f)
        roofline["traffic"] = captured.get(args.workload)
        # the secondary bound of the FLOP-heavy sweeps: how busy ncu saw the FP64 pipe in that capture (its peak,
        # measured with tools/fp64_probe.cu: 36.5 TFLOP/s of DFMA)
        roofline["fp64_pipe_busy_pct"] = captured.get("_fp64_pipe_busy_pct", {}).get(args.workload)

    # ---- resident stepping leg: full drift-kick-drift steps, no PCIe --------------
    resident = None
    if not args.no_e2e and not com:
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dt_step = 1e-3*2*np.pi*np.sqrt(float(np.hypot(w["x"][1], w["y"][1]))**3/(w["G"]*w["m"][0]))
        # one fused sweep per step; sharded, every rank advances its replica of the central body itself and
        # only the back-reaction sums (none for radiation_forces) are all-reduced
        step = lambda: shard.fused_leapfrog_step(dt_step, w["G"], group=group)
        for _ in range(3):
            step()
        barrier()
        l0 = shard.launches
        w0 = time.time()
        s0.record()
        for _ in range(args.steps):
            step()
        s1.record()
        barrier()
        clocks.window(w0, time.time())
        tr = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        ms_r = float(tr.item())/args.steps
        resident = {"value": world*n*nf/(ms_r*1e-3), "unit": "evals/s", "ms_per_step": ms_r,
                    "launches_per_step": (shard.launches - l0)//args.steps,
                    "what": ("device-resident drift-kick-drift step, state never leaves HBM: one fused sweep (rebx_b200_leapfrog_step" +
                             (")" if world == 1 else "'s two halves with the all-reduce of the back-reaction sums between them; the central body's record is replicated and advanced by every rank)"))}

    # ---- resident Wisdom-Holman leg: BASELINE config 2 is quoted "under WHFast" -- Kepler drift about the star on
    # the device + kick by the stacked effects (device.Shard.wh_step), state never leaves HBM -----------------------
    resident_wh = None
    if not args.no_e2e and not com:
        shard.wh_steps(3, dt_step, w["G"], group)
        barrier()
        l0 = shard.launches
        w0 = time.time()
        s0.record()
        shard.wh_steps(args.steps, dt_step, w["G"], group)
        s1.record()
        barrier()
        clocks.window(w0, time.time())
        tr = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        ms_w = float(tr.item())/args.steps
        resident_wh = {"value": world*n*nf/(ms_w*1e-3), "unit": "evals/s", "ms_per_step": ms_w,
                       "launches_per_step": (shard.launches - l0)//args.steps,
                       "what": "device-resident Wisdom-Holman steps (device.Shard.wh_steps): Kepler drift about the central body (rebx_b200_kepler_drift, universal variables), kick by the stacked effects; the half drifts between consecutive kicks merged into one, as WHFast does between outputs; kick + drift in one fused sweep (rebx_b200_wh_kick_drift) for a single effect without back-reaction"}

    # ---- end-to-end leg: the REBOUND-facing call on host particle arrays ---------------------------
    e2e = None
    if not args.no_e2e:
        del shard
        torch.cuda.empty_cache()
        sim, rebx, _ = scenarios.build(w, add_order, coordinates=args.coordinates)
        for _ in range(2):
            sim.additional_forces()
        drain = sim.messages if args.coordinates == "BARYCENTRIC" else sim.process_messages   # BARYCENTRIC evaluates the star too, whose
        drain()                                                                              # missing d_factor is reported as in the reference
        s0 = rebx.stats()
        steps = args.e2e_steps or min(args.steps, 10)
        barrier()
        w0 = time.time()
        t0 = time.perf_counter()
        for _ in range(steps):
            sim.additional_forces()            # synchronous: accelerations are valid on return
        dt = time.perf_counter() - t0
        clocks.window(w0, time.time())
        s1 = rebx.stats()
        drain()
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": world*n*nf*steps/dt, "unit": "evals/s",
               "h2d_bytes_per_step": (s1["h2d_bytes"] - s0["h2d_bytes"])//steps,
               "d2h_bytes_per_step": (s1["d2h_bytes"] - s0["d2h_bytes"])//steps,
               "ms_per_step": dt/steps*1e3, "steps": steps,
               "launches_per_step": (s1["launches"] - s0["launches"])//steps,
               "call": "sim->additional_forces(sim) == rebx_additional_forces, host AoS particles (cudaHostRegister-pinned)"}
        if numa:
            e2e["numa"] = numa
        rebx.detach()
        sim.free()

    clocks.__exit__(None, None, None)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_leg(builder, add_order, min(args.cpu_sample, n) if not com else min(20000, n), 5, args.coordinates)
        cpu["unit"] = "evals/s"
        cpu.pop("ms_per_dispatch")
        if not args.no_parallel_bound and not com:
            cpu["parallel_upper_bound"] = cpu_parallel_upper_bound(args.workload, min(500_000, n))

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"{args.workload}: {n} test particles per GPU, {'+'.join(add_order)}" + (f", REBX_COORDINATES_{args.coordinates}" if com else ""),
                           "particles_per_gpu": n, "forces": add_order,
                           "l2": f"inputs larger than L2 ({shard_bytes(n, exec_order)/1e6:.0f} MB streamed per step)",
                           "sharding": ("index ranges; body state broadcast per step (NCCL) when n_gpus > 1" if not com else
                                        "index ranges; per step an all-gather of 7 mass moments and of 3 back-reaction sums per rank (JACOBI: exclusive prefix / suffix over ranks; BARYCENTRIC: all-reduce)")},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "resident_step": resident, "resident_wh_step": resident_wh, "gpu_launches": gpu_launches,
                "clocks": clocks.summary()}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def shard_bytes(n, exec_order):
    from reboundx_b200 import device
    return device.algorithmic_bytes_per_particle([device.KIND[k] for k in exec_order])*n


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
