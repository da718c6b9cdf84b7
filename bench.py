#!/usr/bin/env python
"""bench.py -- particle-force evaluations per second of the rebx_additional_forces hot path.

Headline workload (BASELINE.json configs[1]): synthetic debris disk, 10^7 test particles per GPU,
radiation_forces with Poynting-Robertson drag, beta per particle.  One "step" is one force evaluation (one
dispatch) over the whole ensemble.

  python bench.py --gpus 1 --steps K --warmup W                  # this build
  torchrun ... bench.py --gpus N ...                             # one rank per GPU
  python bench.py --impl reference ...                           # the reference's own CPU loop, full size

One JSON line (rank 0).  value = evals/s with state resident in HBM (CUDA events on the launching stream, max over
ranks), weak scaling (10^7 particles per GPU).  e2e = the same through the REBOUND-facing call
sim->additional_forces(sim) on HOST particle arrays, copies inside the timed region.  roofline = the sweep kernel's
algorithmic HBM bytes (or, for the FLOP-heavy stacks, its FP64 operation slots) over its event-timed duration against
the measured peak.  cpu_baseline = the reference's own C loop (oracle/_ref) on one host core, bounded sample.
Further keys:
  strong          the SAME 10^7 particles split over the N GPUs (strong scaling), with the 1-GPU time of the same
                  ensemble measured in the same run -> efficiency_vs_1gpu
  sharded_parity  (N > 1) a subsample of every rank's shard re-evaluated as ONE shard on rank 0 and compared bit for bit
  other_math      (N = 1) the headline sweep in the other arithmetic build (tolerance when the run is exact)
  workloads       (N = 1, headline workload) the other BASELINE configurations -- ring, asteroids, planetesimals in
                  PARTICLE and JACOBI coordinates -- each with ms_per_step, roofline (bit-exact and tolerance-mode
                  arithmetic) and e2e
  resident_step / resident_wh_step   whole integrator steps with the state kept in HBM
  resident_session_wh_step           (N = 1) the same Wisdom-Holman steps driven through the C session API (no torch, no Python per step)
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
FP64_PEAK_TFLOPS = 36.5        # DFMA = 2 flops; measured with tools/fp64_probe.cu on this pool's B200 (profiles/r01_fp64_probe.txt)
WORKLOADS = {
    # name: (builder, forces in ADD order, default particles per GPU)
    "debris_disk": ("debris_disk", ["radiation_forces"], 10_000_000),
    "circumplanetary_ring": ("circumplanetary_ring", ["gravitational_harmonics", "gr_potential"], 10_000_000),
    "asteroid_ensemble": ("asteroid_ensemble", ["yarkovsky_effect", "gr_potential"], 1_000_000),
    "planetesimal_disk": ("planetesimal_disk", ["modify_orbits_forces", "gas_damping_timescale", "type_I_migration"], 1_000_000),
}
FP64_BOUND = {"asteroid_ensemble"}          # config 4 is bounded by the FP64 pipe, the others by HBM


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
    ap.add_argument("--no-extras", action="store_true", help="skip the other BASELINE configurations (`workloads`), the strong-scaling leg and the sharded-parity check")
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


def captured():
    """ncu figures of the sweeps (dram bytes, FP64 slots per particle, FP64 pipe busy), captured once per kernel
    change under a profiler and committed; bench.py only quotes them, with their provenance."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        return json.load(f)


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


# ---- the reference's CPU implementation ----------------------------------------------------------------------------
def cpu_reference_leg(workload, add_order, sample_n, steps, warmup, coordinates="PARTICLE"):
    """The reference's own C implementation (oracle/_ref, kind "reference") or, if that library was never built, the C
    restatement (kind "port"), one host thread: `steps` dispatches of rebx_additional_forces over `sample_n` test
    particles after `warmup` untimed ones, MEAN time per dispatch (the statistic of the GPU arm)."""
    from reboundx_b200 import scenarios, workloads
    from oracle import oracle
    w = getattr(workloads, workload)(sample_n)
    n = w["x"].shape[0]
    nf = len(add_order)
    if oracle.have_ref():
        sim, rebx, _ = scenarios.build(w, add_order, lib=oracle.ref_lib(), coordinates=coordinates)
        for _ in range(warmup):
            sim.additional_forces()
        t0 = time.perf_counter()
        for _ in range(steps):
            sim.additional_forces()
        t = (time.perf_counter() - t0)/steps
        sim.messages()
        rebx.detach()
        sim.free()
        kind = "reference"
    else:
        for _ in range(warmup):
            oracle.port_apply(w, add_order[::-1], None, coordinates)
        t0 = time.perf_counter()
        for _ in range(steps):
            oracle.port_apply(w, add_order[::-1], None, coordinates)
        t, kind = (time.perf_counter() - t0)/steps, "port"
    return {"value": (n - 1)*nf/t, "unit": "evals/s", "cores": 1, "kind": kind,
            "sample": f"{workload}: {n - 1} particles x {nf} force(s), mean of {steps} dispatches of rebx_additional_forces after {warmup} warm-up, 1 thread (the reference dispatch is single-threaded)",
            "ms_per_dispatch": t*1e3}


def cpu_parallel_upper_bound(workload, per_worker):
    """What the box's host cores could do if the reference's dispatch were run as `nproc` independent shards at
    once (it is not: REBOUNDx's force path is single-threaded).  Reported next to the 1-core
    reference figure as an upper bound for a hypothetical embarrassingly-parallel CPU version -- the number the
    PCIe-bound end-to-end figure has to be read against; the resident figures are three orders above both."""
    n = len(os.sched_getaffinity(0))
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload, "--steps", "5", "--warmup", "1",
           "--n", str(per_worker), "--no-parallel-bound"]
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


def workload_label(workload, n, add_order, coordinates):
    return (f"{workload}: {n} test particles per GPU, {'+'.join(add_order)}" +
            (f", REBX_COORDINATES_{coordinates}" if coordinates != "PARTICLE" else ""))


def run_reference(args):
    """--impl reference: the reference's own CPU loop on the GPU arm's config (full size in PARTICLE coordinates; the
    centre-of-mass frames are O(N^2) there and get a 2*10^4 sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    builder, add_order, n_default = WORKLOADS[args.workload]
    n = args.n or n_default
    sample = n if args.coordinates == "PARTICLE" else min(20000, n)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    base = cpu_reference_leg(builder, add_order, sample, steps, warmup, args.coordinates)
    full = sample == n
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": "evals/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_dispatch"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_label(args.workload, n, add_order, args.coordinates),
                       "particles_per_gpu": n, "forces": add_order,
                       "sample": (f"each step = one dispatch of the reference's rebx_additional_forces over all {sample} test particles of this workload" if full else
                                  f"each step = one dispatch over a bounded sample of {sample} test particles (the reference's centre-of-mass frames are O(N^2))") +
                                 ", 1 host thread (the reference is single-threaded)"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": base["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
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


# ---- GPU arm ---------------------------------------------------------------------------------------------------------
class Ctx:
    """What every leg needs: torch, the process group, this rank's device."""
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dev = f"cuda:{self.local}"
        self.clocks = ClockSampler(self.local)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, fn, steps, warmup=3):
        """`steps` calls of fn() bracketed by barrier + synchronize, CUDA events on the launching (current) stream;
        returns ms per call, max over ranks."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        w0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        self.clocks.window(w0, time.time())
        return self.max_over_ranks(e0.elapsed_time(e1))/steps


def roofline_for(workload, math, coordinates, shard, kernel_ms, exec_order):
    """Roofline of the dominant kernel: HBM (algorithmic bytes / event-timed duration vs the measured copy bandwidth) or,
    for the FLOP-heavy stacks, the FP64 pipe (FP64 operation slots per particle, counted by ncu for this kernel, x 2 flops
    vs the measured DFMA peak)."""
    peak, peak_src = peaks()
    com = coordinates != "PARTICLE"
    achieved = shard.algorithmic_bytes/(kernel_ms*1e-3)/1e9
    kname = ("pass_kernel_tol(%s)" if math == "tolerance" else "pass_kernel(%s)") % ",".join(exec_order)
    if com:
        kname = (("jacobi_fused_tol_kernel(%s)" if math == "tolerance" else "jacobi_fused_kernel(%s)") % ",".join(exec_order) + ": moments, forces, back-reaction suffix in one cooperative launch"
                 if coordinates == "JACOBI" and shard.standalone else
                 coordinates.lower() + " pipeline: moments, scans, sweep(%s), apply" % ",".join(exec_order))
    hbm = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved/peak, "traffic": None,
           "kernel": kname, "algorithmic_bytes_per_particle": shard.algorithmic_bytes//shard.n, "kernel_ms": kernel_ms,
           "peak_source": peak_src, "math": math}
    if hbm["frac"] > 1.0:
        hbm["note"] = ("above 1: the denominator is the measured COPY bandwidth (half reads, half writes); this sweep reads more than it writes "
                       "(radiation_forces: 10 of its 13 doubles per particle are reads) and HBM3e streams reads faster than a copy")
    cap = captured().get(f"{workload}/{math}") if not com else (captured().get(f"{workload}_{coordinates}/{math}") if coordinates == "JACOBI" and shard.standalone else None)
    if cap and shard.n - 1 == WORKLOADS[workload][2]:
        hbm["traffic"] = cap.get("dram_bytes")
        hbm["traffic_from"] = cap.get("from")
        hbm["fp64_pipe_busy_pct"] = cap.get("fp64_pipe_busy_pct")
    if workload in FP64_BOUND and cap and cap.get("fp64_slots_per_particle"):
        slots = cap["fp64_slots_per_particle"]
        tf = 2.0*slots*shard.n/(kernel_ms*1e-3)/1e12
        return {"bound": "fp64", "achieved": tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": tf/FP64_PEAK_TFLOPS,
                "traffic": hbm["traffic"], "kernel": kname, "kernel_ms": kernel_ms, "math": math,
                "fp64_slots_per_particle": slots, "fp64_slots_from": cap.get("from"),
                "what": "FP64 pipe slots the sweep executes per particle (ncu sm__pipe_fp64_cycles_active x 64 lanes / particles, each slot counted as one DFMA = 2 flops) over the event-timed duration, against the measured DFMA peak (tools/fp64_probe.cu)",
                "hbm": {k: hbm[k] for k in ("achieved", "peak", "frac", "algorithmic_bytes_per_particle")}}
    return hbm


def session_leg(cx, sim, rebx, w, steps, nf, n):
    """Resident stepping driven through the C session API (include/rebx_b200.h: rebx_b200_session_*; no torch, no Python
    between the steps): sim->particles uploaded once, `steps` Wisdom-Holman steps in ONE C call, timed on the host around
    that call + a device synchronisation."""
    import ctypes
    from reboundx_b200 import _lib
    lib = _lib.load()
    lib.rebx_b200_session_create.restype = ctypes.c_void_p
    lib.rebx_b200_session_create.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.rebx_b200_session_step_wh.restype = ctypes.c_int
    lib.rebx_b200_session_step_wh.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
    lib.rebx_b200_session_launches.restype = ctypes.c_uint64
    lib.rebx_b200_session_launches.argtypes = [ctypes.c_void_p]
    lib.rebx_b200_session_free.argtypes = [ctypes.c_void_p]
    s = lib.rebx_b200_session_create(ctypes.cast(sim.ptr, ctypes.c_void_p), ctypes.cast(rebx._ptr, ctypes.c_void_p))
    if not s:
        return None
    dt = 1e-3*2*np.pi*np.sqrt(float(np.hypot(w["x"][1], w["y"][1]))**3/(w["G"]*w["m"][0]))
    try:
        if lib.rebx_b200_session_step_wh(s, 3, dt) != 0:
            return None
        cx.torch.cuda.synchronize()
        l0 = lib.rebx_b200_session_launches(s)
        w0 = time.time()
        t0 = time.perf_counter()
        rc = lib.rebx_b200_session_step_wh(s, steps, dt)
        cx.torch.cuda.synchronize()
        ms = (time.perf_counter() - t0)*1e3/steps
        cx.clocks.window(w0, time.time())
        if rc != 0:
            return None
        return {"value": n*nf/(ms*1e-3), "unit": "evals/s", "ms_per_step": ms, "steps": steps,
                "launches_per_step": (lib.rebx_b200_session_launches(s) - l0)/steps,
                "what": "resident Wisdom-Holman steps driven from C: rebx_b200_session_create(sim, rebx) uploads sim->particles once, "
                        "rebx_b200_session_step_wh(session, steps, dt) issues every step (host timer around the one C call + device synchronisation)"}
    finally:
        lib.rebx_b200_session_free(s)


def e2e_leg(cx, w, add_order, coordinates, steps, nf, n, n_active=None, session_steps=0):
    """The REBOUND-facing call on host particle arrays, copies inside the timed region."""
    from reboundx_b200 import scenarios
    sim, rebx, _ = scenarios.build(w, add_order, coordinates=coordinates)
    if n_active is not None:
        sim.N_active = n_active
    t_first = time.perf_counter()
    sim.additional_forces()                # first call: pins the caller's array (cudaHostRegister), builds and uploads the tables
    first_ms = (time.perf_counter() - t_first)*1e3
    sim.additional_forces()
    drain = sim.messages if coordinates == "BARYCENTRIC" else sim.process_messages   # BARYCENTRIC evaluates the star too, whose
    drain()                                                                          # missing d_factor is reported as in the reference
    s0 = rebx.stats()
    cx.barrier()
    w0 = time.time()
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.additional_forces()            # synchronous: accelerations are valid on return
    dt = time.perf_counter() - t0
    cx.clocks.window(w0, time.time())
    s1 = rebx.stats()
    drain()
    dt = cx.max_over_ranks(dt)
    out = {"value": cx.world*n*nf*steps/dt, "unit": "evals/s",
           "h2d_bytes_per_step": (s1["h2d_bytes"] - s0["h2d_bytes"])//steps,
           "d2h_bytes_per_step": (s1["d2h_bytes"] - s0["d2h_bytes"])//steps,
           "ms_per_step": dt/steps*1e3, "steps": steps,
           "launches_per_step": (s1["launches"] - s0["launches"])//steps,
           "first_call_ms": first_ms,
           "first_call_note": "outside the timed region (two warm-up calls): the first call pins the caller's particle array with cudaHostRegister (~1.1 s for 1.28 GB) and builds + uploads the parameter tables",
           "call": "sim->additional_forces(sim) == rebx_additional_forces, host AoS particles (cudaHostRegister-pinned)"}
    if session_steps > 0:
        out["_session"] = session_leg(cx, sim, rebx, w, session_steps, nf, n)
    rebx.detach()
    sim.free()
    return out


def extra_workload(cx, name, coordinates, steps):
    """One of the other BASELINE configurations on this GPU: resident sweep in both arithmetic modes + e2e."""
    from reboundx_b200 import device, workloads
    torch = cx.torch
    builder, add_order, n = WORKLOADS[name]
    exec_order, nf = add_order[::-1], len(add_order)
    w = getattr(workloads, builder)(n)
    out = {"config": workload_label(name, n, add_order, coordinates), "evals_per_step": n*nf}
    for math in (("exact", "tolerance") if coordinates in ("PARTICLE", "JACOBI") else ("exact",)):
        shard = device.Shard(w, exec_order, cx.dev, coordinates=coordinates, math=math, standalone=True)
        l0 = shard.launches
        ms = cx.timed(shard.evaluate, steps)
        launches = (shard.launches - l0)//(steps + 3)
        kernel_ms = cx.timed(shard.launch_local, steps)
        leg = {"ms_per_step": ms, "value": n*nf/(ms*1e-3), "unit": "evals/s", "launches_per_step": launches,
               "roofline": roofline_for(name, math, coordinates, shard, kernel_ms, exec_order)}
        if math == "exact":
            out.update(leg)
        else:
            out["tolerance_mode"] = leg
        del shard
        torch.cuda.empty_cache()
    if not cx.args.no_e2e:
        out["e2e"] = e2e_leg(cx, w, add_order, coordinates, 5, nf, n)
        if name == "asteroid_ensemble" and coordinates == "PARTICLE":
            # BASELINE config 4 names `gr` (the 1PN effect of src/gr.c:65-164, N_active = 1: the star is active, the
            # asteroids are test particles) next to gr_potential: it exists in the host-facing dispatcher only
            # (device reduction of the test particles' pull -> the active bodies on the host -> device sweep)
            gr = e2e_leg(cx, w, ["yarkovsky_effect", "gr"], coordinates, 5, 2, n, n_active=1)
            gr["config"] = workload_label(name, n, ["yarkovsky_effect", "gr"], coordinates) + ", N_active = 1"
            out["gr_variant_e2e"] = gr
    return out


def graph_timed(cx, shard, steps, barrier=True):
    """`steps` sweeps of `shard` captured once in a CUDA graph and replayed: the launch gaps of a Python / ctypes issue
    loop (3-4 us per launch, a fifth of a 22 us sweep) are taken out, what remains is the device's own time.  Only for
    stacks without back-reaction (no collective between the sweeps).  ms per sweep, max over ranks."""
    torch = cx.torch
    shard.launch()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(shard.device)
    side.wait_stream(torch.cuda.current_stream(shard.device))
    with torch.cuda.graph(graph, stream=side):
        for _ in range(steps):
            shard.launch()
    for _ in range(3):
        graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if barrier:
        cx.barrier()
    else:
        torch.cuda.synchronize()
    w0 = time.time()
    e0.record()
    graph.replay()
    e1.record()
    if barrier:
        cx.barrier()
    else:
        torch.cuda.synchronize()
    cx.clocks.window(w0, time.time())
    ms = e0.elapsed_time(e1)/steps
    return cx.max_over_ranks(ms) if barrier else ms


def strong_leg(cx, w_rank0, builder, add_order, n_total, steps):
    """Strong scaling: ONE ensemble of n_total test particles (the headline 10^7) split by index range over the ranks,
    every rank holding the replicated central body.  The 1-GPU time of the same ensemble is measured in the same run on
    rank 0 (its weak-scaling shard IS that ensemble), so the line carries the efficiency itself."""
    from reboundx_b200 import device, sharding, workloads
    torch = cx.torch
    exec_order, nf = add_order[::-1], len(add_order)
    world, rank = cx.world, cx.rank
    # rank r owns test particles [lo, hi) of rank 0's ensemble (every rank regenerates it from the same seed: no
    # scatter needed); global indices 1 + lo .. 1 + hi
    w = w_rank0 if rank == 0 else getattr(workloads, builder)(n_total, seed=3)
    lo, hi = sharding.shard_range(n_total, world, rank)
    sub = {k: (np.concatenate([v[:1], v[1 + lo:1 + hi]]) if isinstance(v, np.ndarray) and v.shape[0] == n_total + 1 else
               (v[lo:hi] if isinstance(v, np.ndarray) and v.shape[0] == n_total else v)) for k, v in w.items()}
    shard = device.Shard(sub, exec_order, cx.dev, lo=(0 if rank == 0 else 1), index_base=(0 if rank == 0 else 1 + lo), math=cx.args.math)
    ms = cx.timed(lambda: shard.evaluate(None), steps)
    out = {"particles_total": n_total, "particles_per_gpu": hi - lo, "ms_per_step": ms, "value": n_total*nf/(ms*1e-3), "unit": "evals/s",
           "what": "the headline ensemble split by index range over the ranks; per evaluation each rank launches one sweep over its range, "
                   "the body record is re-broadcast only when its owner changed it (never in a force-evaluation loop), back-reaction sums are all-reduced"}
    if all(k in (1, 4) for k in shard.kinds):
        out["graph_ms_per_step"] = graph_timed(cx, shard, max(steps, 50))
    del shard
    torch.cuda.empty_cache()
    return out


def sharded_parity(cx, shard, w, exec_order, per_rank=50_000):
    """Correctness of the sharded evaluation inside the scaling run: every rank contributes the last `per_rank`
    particles of its shard (inputs and the accelerations its sharded sweep produced); rank 0 evaluates all of them as
    ONE shard and compares bit for bit."""
    from reboundx_b200 import device, scenarios
    torch, dist = cx.torch, cx.dist
    if any(it is not None for it in shard.itables):
        return {"checked": False, "why": "stack with an integer table"}
    k = min(per_rank, shard.n - 1)
    sl = slice(shard.n - k, shard.n)                   # the last k elements are test particles on every rank
    keys = ["x", "y", "z", "vx", "vy", "vz", "m", "r"]
    for q in ("ax", "ay", "az"):
        shard.state[q].zero_()
    shard.evaluate(None)
    torch.cuda.synchronize()
    cols = [shard.state[q][sl] for q in keys] + [t[sl] for tabs in shard.tables for t in tabs] + [shard.state[q][sl] for q in ("ax", "ay", "az")]
    mine = torch.stack([c.to(torch.float64) for c in cols])
    allr = [torch.empty_like(mine) for _ in range(cx.world)]
    dist.all_gather(allr, mine)
    out = None
    if cx.rank == 0:
        cat = torch.cat(allr, dim=1).cpu().numpy()
        wsub = {q: np.concatenate([[w[q][0]], cat[i]]) for i, q in enumerate(keys)}
        for name, v in w.items():
            if not isinstance(v, np.ndarray):
                wsub[name] = v
        j = len(keys)
        for fname in exec_order:
            for pname in scenarios._PARTICLE_DOUBLES.get(fname, []):
                wsub[pname] = cat[j]
                j += 1
        one = device.Shard(wsub, exec_order, cx.dev, math=cx.args.math, standalone=True)
        one.evaluate()
        torch.cuda.synchronize()
        got = np.stack([a[1:] for a in one.acc()])
        want = cat[j:j + 3]
        out = {"checked": True, "bit_identical": bool(np.array_equal(got, want)), "particles": int(got.shape[1]),
               "max_abs_diff": float(np.max(np.abs(got - want))),
               "what": f"{k} test particles of every rank's shard: the sharded evaluation vs the same particles evaluated as one shard on rank 0"}
        del one
    return out


def run_ours(args):
    cx = Ctx(args)
    world, rank, local = cx.world, cx.rank, cx.local
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    torch, dist = cx.torch, cx.dist
    from reboundx_b200 import device, workloads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the force path has no CPU fallback")
    torch.cuda.set_device(local)
    os.environ["REBX_B200_DEVICE"] = str(local)
    if args.math == "tolerance":
        os.environ["REBX_B200_MATH"] = "tolerance"          # the host-facing leg reads it when its device context is created
    # stdout carries ONE JSON line: anything a library prints to fd 1 meanwhile (NCCL's version banner under
    # NCCL_DEBUG=VERSION, ...) is sent to stderr, and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = cx.dev
    builder, add_order, n_default = WORKLOADS[args.workload]
    n = args.n or n_default
    exec_order = add_order[::-1]
    nf = len(add_order)
    com = args.coordinates != "PARTICLE"
    if com and args.workload != "planetesimal_disk":
        raise SystemExit("--coordinates applies to the rebx_com_force effects of --workload planetesimal_disk")
    clocks = cx.clocks
    clocks.__enter__()          # sampled from here to the end of the e2e leg; the timed regions are marked as windows
    steps = max(1, args.steps)

    # Weak scaling: every rank owns n test particles of one global ensemble of world*n; the central body (global
    # index 0) is replicated; its record is broadcast from rank 0 when the owner changed it.
    w = getattr(workloads, builder)(n, seed=3 + rank)
    lo = 0 if rank == 0 else 1
    shard = device.Shard(w, exec_order, dev, lo=lo, index_base=(0 if rank == 0 else 1 + rank*n), coordinates=args.coordinates, math=args.math,
                         standalone=(world == 1))

    group = None
    for _ in range(max(3, args.warmup)):
        shard.evaluate(group)
    # ---- resident leg: K evaluations, CUDA events on the launching stream, max over ranks -------
    launches0 = shard.launches
    ms_step = cx.timed(lambda: shard.evaluate(group), steps, warmup=0)
    gpu_launches = shard.launches - launches0                    # launches of the timed `value` region only
    value = world*n*nf/(ms_step*1e-3)
    # ---- roofline of the dominant kernel (the fused sweep), timed alone per launch -----------
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    w0 = time.time()
    k0.record()
    for _ in range(steps):
        shard.launch_local()
    k1.record()
    torch.cuda.synchronize()
    clocks.window(w0, time.time())
    kernel_ms = k0.elapsed_time(k1)/steps
    roofline = roofline_for(args.workload, args.math, args.coordinates, shard, kernel_ms, exec_order)

    parity = None
    if world > 1 and not com and not args.no_extras:
        parity = sharded_parity(cx, shard, w, exec_order)

    # ---- the same sweep in the other arithmetic build (the headline stays the default, bit-exact one) ---------------
    other_math = None
    if world == 1 and not com and not args.no_extras:
        om = "tolerance" if args.math == "exact" else "exact"
        osh = device.Shard(w, exec_order, dev, math=om, standalone=True)
        oms = cx.timed(osh.evaluate, steps)
        okm = cx.timed(osh.launch_local, steps)
        other_math = {"math": om, "ms_per_step": oms, "value": n*nf/(oms*1e-3), "unit": "evals/s",
                      "roofline": roofline_for(args.workload, om, args.coordinates, osh, okm, exec_order)}
        if not args.no_e2e and nf == 1:
            # resident Wisdom-Holman steps in that build as well (fused kick + Kepler drift sweep)
            dt_o = 1e-3*2*np.pi*np.sqrt(float(np.hypot(w["x"][1], w["y"][1]))**3/(w["G"]*w["m"][0]))
            osh.wh_steps(3, dt_o, w["G"], None)
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            wo = time.time()
            o0.record()
            osh.wh_steps(steps, dt_o, w["G"], None)
            o1.record()
            torch.cuda.synchronize()
            clocks.window(wo, time.time())
            other_math["resident_wh_step_ms"] = o0.elapsed_time(o1)/steps
            other_math["resident_step_ms"] = cx.timed(lambda: osh.fused_leapfrog_step(dt_o, w["G"]), steps)
        del osh
        torch.cuda.empty_cache()

    # ---- resident stepping leg: full drift-kick-drift steps, no PCIe --------------
    resident = resident_wh = None
    if not args.no_e2e and not com:
        dt_step = 1e-3*2*np.pi*np.sqrt(float(np.hypot(w["x"][1], w["y"][1]))**3/(w["G"]*w["m"][0]))
        # one fused sweep per step; sharded, every rank advances its replica of the central body itself and
        # only the back-reaction sums (none for radiation_forces) are all-reduced
        l0 = shard.launches
        ms_r = cx.timed(lambda: shard.fused_leapfrog_step(dt_step, w["G"], group=group), steps)
        resident = {"value": world*n*nf/(ms_r*1e-3), "unit": "evals/s", "ms_per_step": ms_r,
                    "launches_per_step": (shard.launches - l0)//(steps + 3),
                    "what": ("device-resident drift-kick-drift step, state never leaves HBM: one fused sweep (rebx_b200_leapfrog_step" +
                             (")" if world == 1 else "'s two halves with the all-reduce of the back-reaction sums between them; the central body's record is replicated and advanced by every rank)"))}
        # BASELINE config 2 is quoted "under WHFast": Kepler drift about the star on the device + kick by the stacked effects
        shard.wh_steps(3, dt_step, w["G"], group)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cx.barrier()
        l0 = shard.launches
        w0 = time.time()
        s0.record()
        shard.wh_steps(steps, dt_step, w["G"], group)
        s1.record()
        cx.barrier()
        clocks.window(w0, time.time())
        ms_w = cx.max_over_ranks(s0.elapsed_time(s1))/steps
        resident_wh = {"value": world*n*nf/(ms_w*1e-3), "unit": "evals/s", "ms_per_step": ms_w,
                       "launches_per_step": (shard.launches - l0)//steps,
                       "what": "device-resident Wisdom-Holman steps (device.Shard.wh_steps): Kepler drift about the central body (rebx_b200_kepler_drift, universal variables), kick by the stacked effects; the half drifts between consecutive kicks merged into one, as WHFast does between outputs; kick + drift in one fused sweep (rebx_b200_wh_kick_drift) for a single effect without back-reaction"}

    l2_note = f"inputs larger than L2 ({shard.algorithmic_bytes/1e6:.0f} MB streamed per step)"
    del shard
    torch.cuda.empty_cache()

    # ---- strong scaling: the headline ensemble (10^7 in all) over the N GPUs ---------------------------------------
    strong = None
    if not com and not args.no_extras:
        n_total = n
        if world == 1:
            strong = {"particles_total": n_total, "particles_per_gpu": n_total, "ms_per_step": ms_step, "value": value, "unit": "evals/s",
                      "efficiency_vs_1gpu": 1.0, "what": "one GPU: the weak and the strong ensemble coincide"}
        else:
            strong = strong_leg(cx, w, builder, add_order, n_total, steps)
            # the 1-GPU time of the same ensemble: rank 0's weak shard is that ensemble; measured alone, the other ranks wait
            t1 = None
            if rank == 0:
                one = device.Shard(w, exec_order, dev, math=args.math, standalone=True)
                for _ in range(3):
                    one.evaluate()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a0.record()
                for _ in range(steps):
                    one.evaluate()
                a1.record()
                torch.cuda.synchronize()
                t1 = a0.elapsed_time(a1)/steps
                t1g = graph_timed(cx, one, max(steps, 50), barrier=False) if "graph_ms_per_step" in strong else None
                del one
                torch.cuda.empty_cache()
            cx.barrier()
            if rank == 0:
                strong["ms_per_step_1gpu_same_run"] = t1
                strong["efficiency_vs_1gpu"] = t1/(world*strong["ms_per_step"])
                if t1g is not None:
                    g = strong.pop("graph_ms_per_step")
                    strong["cuda_graph"] = {"ms_per_step": g, "value": n_total*nf/(g*1e-3), "ms_per_step_1gpu_same_run": t1g,
                                            "efficiency_vs_1gpu": t1g/(world*g),
                                            "what": "the same sweeps captured once in a CUDA graph and replayed (no Python / launch gaps between them), both the sharded and the 1-GPU run"}

    # ---- end-to-end leg: the REBOUND-facing call on host particle arrays ---------------------------
    e2e = resident_session = None
    if not args.no_e2e:
        e2e = e2e_leg(cx, w, add_order, args.coordinates, args.e2e_steps or min(steps, 10), nf, n,
                      session_steps=(steps if world == 1 and not com and nf == 1 else 0))
        resident_session = e2e.pop("_session", None)
        if numa:
            e2e["numa"] = numa

    # ---- ONE simulation of the headline size spread over all N GPUs by the C dispatcher itself (REBX_B200_DEVICES): the
    # unchanged drop-in call of a single-process C host; rank 0 alone runs it, the other ranks wait ----------------------
    e2e_one = None
    if world > 1 and not args.no_e2e and not com and not args.no_extras:
        cx.barrier()
        if rank == 0:
            os.environ["REBX_B200_DEVICES"] = ",".join(str(d) for d in range(world))
            solo = Ctx(args)
            solo.world, solo.clocks = 1, clocks          # no collectives inside: one process drives every device
            e2e_one = e2e_leg(solo, w, add_order, args.coordinates, args.e2e_steps or min(steps, 10), nf, n)
            e2e_one["devices"] = world
            e2e_one["what"] = ("ONE simulation of %d test particles, one process, one host thread: sim->additional_forces(sim) with REBX_B200_DEVICES "
                               "naming all %d GPUs -- the dispatcher splits the particle array by index range over per-device copy/compute pipelines" % (n, world))
            del os.environ["REBX_B200_DEVICES"]
        cx.barrier()

    # ---- the other BASELINE configurations (one GPU, default invocation) -----------------------------------------
    extras = None
    if world == 1 and args.workload == "debris_disk" and not com and not args.no_extras and args.n == 0:
        del w
        extras = {}
        for name, coords in (("circumplanetary_ring", "PARTICLE"), ("asteroid_ensemble", "PARTICLE"),
                             ("planetesimal_disk", "PARTICLE"), ("planetesimal_disk", "JACOBI")):
            key = name if coords == "PARTICLE" else f"{name}_{coords}"
            extras[key] = extra_workload(cx, name, coords, steps)

    clocks.__exit__(None, None, None)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_leg(builder, add_order, min(args.cpu_sample, n) if not com else min(20000, n), 5, 1, args.coordinates)
        cpu.pop("ms_per_dispatch")
        if not args.no_parallel_bound and not com:
            cpu["parallel_upper_bound"] = cpu_parallel_upper_bound(args.workload, min(500_000, n))

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_label(args.workload, n, add_order, args.coordinates),
                           "particles_per_gpu": n, "forces": add_order, "math": args.math,
                           "l2": l2_note,
                           "sharding": ("index ranges; the central body's record is replicated and re-broadcast (NCCL) only when its owner changed it" if not com else
                                        "index ranges; per step an all-gather of 7 mass moments and of 3 back-reaction sums per rank (JACOBI: exclusive prefix / suffix over ranks; BARYCENTRIC: all-reduce)")},
                "roofline": roofline, "other_math": other_math, "cpu_baseline": cpu, "e2e": e2e, "e2e_one_simulation_all_gpus": e2e_one, "strong": strong, "sharded_parity": parity,
                "workloads": extras, "resident_step": resident, "resident_wh_step": resident_wh, "resident_session_wh_step": resident_session, "gpu_launches": gpu_launches,
                "clocks": clocks.summary()}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
