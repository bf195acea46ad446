#!/usr/bin/env python
"""bench.py -- molecules/s through the CeNTREX beamline (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A step is one pass of the hot path over one batch of synthetic molecules:
BASELINE.json configs[1], the full CeNTREX beamline of
examples/lens_simulation_beamline.py with the ElectrostaticLens (J=2 Stark
curve), 1e7 molecules per GPU per step (weak scaling: the molecules of a run
are independent and are sharded over ranks by global index).

  value      device-resident initial conditions (SoA FP64 in HBM, 480 MB per
             step > 126 MB L2, so no flush is needed) -> fates + Counter;
             CUDA events around the K steps, max over ranks.
  e2e        the same batch through the C ABI's host-buffer entry point
             (cmt_run_host_ic): pinned host ICs -> H2D -> kernels -> fates and
             Counter D2H, all inside the timed region.
  e2e_philox the run_simulation default path: Philox source on the device, only
             the Counter comes back (reported beside e2e, not instead of it).
  roofline   the dominant kernel (lens integrator) against the FP64 pipe, and
             roofline_walk for the ballistic/aperture kernel against HBM.
  cpu_baseline / --impl reference: the CPU oracle port (oracle/cmt_oracle.c,
             OpenMP over all host cores) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG = ROOT / "centrex-molecule-trajectories_b200"
for p in (str(ROOT), str(PKG)):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "molecules/s through CeNTREX beamline"
UNIT = "molecules/s"
WORKLOAD = ("configs[1]: examples/lens_simulation_beamline.py full CeNTREX beamline with ElectrostaticLens, "
            "J=2 mJ=0 Stark curve at 27.6 kV, CeNTREX velocity/position distributions")
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of this workload at
# 1e7 molecules (profiles/r02_full_step.txt); scaled linearly when --molecules differs
NCU_TRAFFIC_LENS_1E7 = 13.2e6            # four segment launches + tail: 3.52 + 2.77 + 2.41 + 2.25 + 2.21 MB read
NCU_TRAFFIC_WALK_1E7 = 517.0e6 + 13.4e6
# smsp__inst_executed_pipe_fp64.sum over the launches of one 1e7-molecule step (walk + 4 segments + tail), ncu
NCU_FP64_WARP_INST_STEP_1E7 = 1.2646e8     # profiles/r02_fp64_inst.csv: 0.64 (walk) + 40.47 + 31.83 + 27.62 + 25.77 (segments) + 0.13 (tail) million
# algorithmic work per unit
FLOP_PER_ROW = 30      # one ballistic step + hit test
FLOP_PER_STEP = 162    # one lens RK step (4 force evaluations)
BYTES_PER_MOLECULE = 49  # IC replay: 6 x 8 B read + 1 B fate written


N_PER_STEP = 10_000_000   # molecules per GPU per step of the GPU arm (configs[1])


def shared_config(n=N_PER_STEP):
    """The same dict from both arms (the driver compares them)."""
    return {"workload": WORKLOAD, "molecules_per_step_per_gpu": int(n),
            "inputs": "SoA FP64 initial conditions resident in HBM (480 MB per 1e7 molecules, larger than the 126 MB L2: no flush needed)",
            "outputs": "fate byte per molecule + per-fate Counter (+ NCCL all-reduce of the Counter when n_gpus > 1)",
            "sharding": "independent molecules, contiguous global-index block per rank"}


def reference_python(cores: int):
    """The UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.sh) timed on the host cores:
    run_simulation(n_jobs=cores) and run_simulation(n_jobs=1) on the lens beamline with the same injected table
    (BASELINE.md section 3).  Runs in a subprocess: its package is also called `trajectories`."""
    import subprocess
    import tempfile

    ref = ROOT / "baseline" / "_ref"
    if not (ref / "trajectories" / "trajectory_simulator.py").exists():
        return {"unavailable": "baseline/_ref is empty (run baseline/install_ref.sh where /root/reference exists)"}
    from trajectories.centrex import lens_table

    r, a = lens_table()
    with tempfile.TemporaryDirectory() as td:
        np.savez(os.path.join(td, "table.npz"), r=r, a=a)
        env = dict(os.environ, PYTHONPATH=f"{ref}:{ROOT / 'oracle' / 'stubs'}")
        for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
            env[k] = "1"                      # one process per core through joblib, like the reference's users
        cmd = [sys.executable, str(ROOT / "baseline" / "run_reference.py"), os.path.join(td, "table.npz"),
               str(8000 * cores), str(cores), "20000", "1"]
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=td, timeout=600)
        except subprocess.TimeoutExpired:
            return {"unavailable": "reference run exceeded 600 s"}
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("REFERENCE_JSON ")]
    if out.returncode != 0 or not lines:
        return {"unavailable": "reference run failed: " + (out.stderr.strip().splitlines() or ["?"])[-1][:200]}
    runs = json.loads(lines[-1][len("REFERENCE_JSON "):])
    return {"what": "otimgren/centrex-molecule-trajectories, unmodified, TrajectorySimulator.run_simulation on the lens beamline "
                    "(wall clock incl. sampling and the joblib pool); matplotlib/h5py/hexalattice/centrex_TlF stubbed, lens table injected",
            "all_cores": {"value": runs[0]["molecules_per_s"], "unit": UNIT, "cores": runs[0]["n_jobs"], "n": runs[0]["molecules"],
                          "seconds": runs[0]["seconds"]},
            "one_core": {"value": runs[1]["molecules_per_s"], "unit": UNIT, "cores": 1, "n": runs[1]["molecules"],
                         "seconds": runs[1]["seconds"]}}


def build_workload():
    from trajectories.centrex import lens_beamline, lens_table
    from trajectories.distributions import CeNTREXPositionDistribution, CeNTREXVelocityDistribution

    bl = lens_beamline(lens_table())
    return bl, CeNTREXVelocityDistribution(), CeNTREXPositionDistribution()


# ---------------------------------------------------------------------------
# CPU arm (oracle port)
# ---------------------------------------------------------------------------
def cpu_rate(bl, vdist, xdist, target_s: float, seed: int = 1):
    """molecules/s of the oracle port on all host cores over ~target_s seconds of work."""
    from oracle import oracle

    src = oracle.make_source(vdist, xdist)
    flat = oracle.flatten(bl.elements)
    threads = oracle.host_cores()                 # all host cores (torchrun would pin OpenMP to 1)
    n = 200_000
    oracle.run(flat, src, seed, 0, n, n_threads=threads)
    t0 = time.perf_counter()
    oracle.run(flat, src, seed, 0, n, n_threads=threads)
    probe = time.perf_counter() - t0
    n = int(max(n, min(2e9, n * target_s / max(probe, 1e-3))))
    t0 = time.perf_counter()
    res = oracle.run(flat, src, seed, 10_000_000_000, n, n_threads=threads)
    dt = time.perf_counter() - t0
    return dict(value=n / dt, n=n, seconds=dt, cores=threads, counters=res["counters"].tolist())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle

    bl, vdist, xdist = build_workload()
    src = oracle.make_source(vdist, xdist)
    flat = oracle.flatten(bl.elements)
    threads = oracle.host_cores()                 # all host cores (torchrun would pin OpenMP to 1)
    oracle.run(flat, src, 1, 0, 200_000, n_threads=threads)
    t0 = time.perf_counter()
    oracle.run(flat, src, 1, 0, 200_000, n_threads=threads)
    probe = time.perf_counter() - t0
    # each step a bounded sample: ~2.5 s of host work, so K+W steps end within a few minutes
    per_step = int(min(1e7, max(200_000, 200_000 * 2.5 / max(probe, 1e-3))))
    for w in range(args.warmup):
        oracle.run(flat, src, 1, (w + 1) * per_step, per_step, n_threads=threads)
    t0 = time.perf_counter()
    for k in range(args.steps):
        oracle.run(flat, src, 1, (args.warmup + k + 1) * per_step, per_step, n_threads=threads)
    dt = time.perf_counter() - t0
    value = args.steps * per_step / dt
    sample = f"{per_step} molecules per step (bounded sample of the 1e7-molecule step), Philox source + propagation + Counter"
    # the reference itself (pure Python) on the same cores, reported beside the port; the port stays the arm's
    # value: it is ~2000x faster than the reference and therefore the harder baseline
    ref_py = None if args.no_reference_python else reference_python(threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": shared_config(), "sample_molecules_per_step": per_step,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "reference_python": ref_py},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while a region runs."""

    REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        if not self.nv:
            return
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            try:
                bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                bits = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if bits & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def start(self):
        def loop():
            while not self._stop.is_set():
                self.sample()
                self._stop.wait(0.001)

        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def bind_near_gpu(index: int):
    """Restrict this process to the CPUs of the GPU's own NUMA node, so that the pinned host buffers of the e2e leg
    are allocated in the memory next to the PCIe root the GPU hangs on (with several ranks per box the H2D streams
    otherwise share one socket's memory and the inter-socket link).  Returns (previous affinity, description)."""
    try:
        before = os.sched_getaffinity(0)
        try:        # the CUDA device's own PCI address (NVML's enumeration order need not be CUDA's)
            import torch

            pr = torch.cuda.get_device_properties(index)
            bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        except Exception:
            import pynvml

            pynvml.nvmlInit()
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
            bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()[-12:]
        dev = "/sys/bus/pci/devices/" + bus
        with open(dev + "/local_cpulist") as f:
            spec = f.read().strip()
        node = open(dev + "/numa_node").read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= before
        if cpus and cpus != before:
            os.sched_setaffinity(0, cpus)
        return before, {"numa_node": node, "cpus_bound": len(cpus) or len(before), "cpus_before": len(before)}
    except Exception as exc:      # no NVML, no sysfs entry, not Linux: run unbound
        return None, {"unbound": type(exc).__name__}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_ours(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from trajectories import _engine as eng
    from trajectories import _native as nat

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the propagation path has no CPU fallback")
    affinity_before, numa = bind_near_gpu(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = nat.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    bl, vdist, xdist = build_workload()
    prop = eng.Propagator(bl.elements, local, n_slots=args.slots)
    source = eng.make_source(vdist, xdist)
    n = int(args.molecules)
    seed = 2026
    first = rank * n                               # this rank's block of the global index range
    ic = prop.draw(source, seed, first, n)         # synthetic CeNTREX-shaped ICs, resident in HBM
    torch.cuda.synchronize()

    def step(slot=None):
        return prop.propagate_ic(ic, first_index=first, want_fate=True, slot=slot)

    def merge_counter():
        prop.join()
        if world > 1:
            dist.all_reduce(prop.counters)         # the Counter merge (tiny, NCCL over NVLink), once per run
            dist.all_reduce(prop.work)

    warm = max(args.warmup, 3)
    for _ in range(warm):
        prop.reset()
        res = step()
    torch.cuda.synchronize()
    counters1 = res.counters.cpu().numpy().copy()  # one step of this rank
    work1 = res.work.cpu().numpy().copy()
    work = work1

    # ---- pass A: K steps back to back on one stream, every kernel timed with CUDA events ----
    sampler = ClockSampler(local)          # NVML start-up takes a rank-dependent time: keep it ahead of the barriers
    sampler.start()                        # SM clock and throttle reasons are sampled through passes A and B
    barrier()
    lib.cmt_timing_enable(1)
    lib.cmt_timing_read(None, None, 1)
    lib.cmt_launch_count(1)
    prop.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step()
    launches = int(lib.cmt_launch_count(0))       # walk + lens segments + tail, per step; pass B replays the same launches
    merge_counter()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms_seq = max_over_ranks(e0.elapsed_time(e1))
    ms_k = (C.c_double * 4)()
    n_k = (C.c_int64 * 4)()
    lib.cmt_timing_read(ms_k, n_k, 1)
    lib.cmt_timing_enable(0)

    # ---- pass B (the headline value): the same K steps, consecutive steps on alternating streams so
    # that the lens integrator of one batch overlaps the walk kernel of the next ----
    slots = [None] * args.steps if args.no_overlap else [k % prop.n_slots for k in range(args.steps)]
    graphs = None
    if not args.no_overlap and not args.no_graphs:
        # one CUDA graph per stream (header memset + walk + lens segments + tail): a replay is a single host-side launch
        graphs = [prop.capture_ic(ic, first_index=first, want_fate=True, slot=s) for s in range(prop.n_slots)]

        def step(slot=None, _plain=step):                      # noqa: F811
            if slot is None:
                return _plain(None)
            graphs[slot].replay()
            return res
    for k in range(warm):
        step(slots[k % len(slots)])
    prop.join()
    prop.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        res = step(slots[k])
    merge_counter()
    e1.record()
    while not e1.query():
        time.sleep(0.002)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    counters = res.counters.cpu().numpy()
    total_expected = counters1 * args.steps
    if world > 1:
        t = torch.from_numpy(total_expected).cuda()
        dist.all_reduce(t)
        total_expected = t.cpu().numpy()
    assert (counters == total_expected).all(), "Counter of the timed run differs from K x one step"
    counters = counters1
    value = world * n * args.steps / (ms * 1e-3)
    value_seq = world * n * args.steps / (ms_seq * 1e-3)

    # ---- the opt-in contracted arithmetic (fused multiply-adds, reciprocal multiplications): same
    # workload, same two passes, reported beside the default; not the headline ----
    contracted = None
    if not args.no_contracted:
        propc = eng.Propagator(bl.elements, local, n_slots=args.slots, math="contracted")
        for _ in range(warm):
            propc.reset()
            resc = propc.propagate_ic(ic, first_index=first, want_fate=True)
        torch.cuda.synchronize()
        cnt_c1 = resc.counters.cpu().numpy().copy()
        lib.cmt_timing_enable(1)
        lib.cmt_timing_read(None, None, 1)
        propc.reset()
        for _ in range(args.steps):
            propc.propagate_ic(ic, first_index=first, want_fate=True)
        torch.cuda.synchronize()
        msc_k = (C.c_double * 4)()
        nc_k = (C.c_int64 * 4)()
        lib.cmt_timing_read(msc_k, nc_k, 1)
        lib.cmt_timing_enable(0)
        gc = [propc.capture_ic(ic, first_index=first, want_fate=True, slot=s_) for s_ in range(propc.n_slots)]
        for k in range(warm):
            gc[k % len(gc)].replay()
        propc.join()
        barrier()
        propc.reset()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for k in range(args.steps):
            gc[k % len(gc)].replay()
        propc.join()
        c1.record()
        torch.cuda.synchronize()
        barrier()
        ms_c = max_over_ranks(c0.elapsed_time(c1))
        lens_c = msc_k[1] / max(nc_k[1], 1)
        contracted = {
            "value": world * n * args.steps / (ms_c * 1e-3), "unit": UNIT, "ms_per_step": ms_c / args.steps,
            "kernel_ms_per_step": {"walk": msc_k[0] / max(nc_k[0], 1), "lens": lens_c},
            "fates_differing_from_default": int(np.abs(cnt_c1 - counters1).sum() // 2),
            "note": "TrajectorySimulator(math='contracted'): same algorithm, relaxed roundings; final rows agree with "
                    "the reference to ~1e-13 relative (tests assert 1e-9), fates can differ only within that distance of an edge",
        }

    # ---- roofline ----
    dfma, dadd = C.c_double(), C.c_double()
    nat.check(lib.cmt_fp64_peak(local, C.byref(dfma), C.byref(dadd)))
    peaks = {}
    mp = ROOT / "MEASURED_PEAKS.json"
    if mp.exists():
        peaks = json.loads(mp.read_text())
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback")
    walk_ms = ms_k[0] / max(n_k[0], 1)
    lens_ms = ms_k[1] / max(n_k[1], 1)
    # rows committed by the walk kernel = all rows of molecules retired there + entrance rows of survivors;
    # the split is not needed for the totals: report the whole-step algorithmic flop and each kernel's share
    flop_lens = FLOP_PER_STEP * float(work[1])
    flop_rows = FLOP_PER_ROW * float(work[0])
    lens_tflops = flop_lens / (lens_ms * 1e-3) / 1e12 if lens_ms > 0 else None
    fp64_peak_tflops = 2 * dfma.value / 1e12       # FMA = 2 flop
    step_ms = ms / args.steps
    lens_tflops_overlapped = flop_lens / (step_ms * 1e-3) / 1e12
    # FP64 pipe time of one step from the instruction counts ncu reports for the same launches (they do not depend
    # on how launches overlap): warp instructions on the FP64 pipe x 2 issue cycles / (SMs x 4 sub-partitions)
    sm_mhz = (torch.cuda.get_device_properties(local).clock_rate / 1e3)
    fp64_cycles = NCU_FP64_WARP_INST_STEP_1E7 * (n / 1e7) * 2 / (prop.dev_sm_count() * 4)
    roofline = {
        "kernel": "lens_seg_kernel (the RK integrator of the ElectrostaticLens: 4 segment launches of 150 steps per step)",
        "regime": "overlapped: pass B, the headline -- consecutive steps on alternating streams, so the lens stage of one step runs "
                  "beside the walk kernel and the lens stages of its neighbours; a step then takes ms_per_step of device time",
        "bound": "fp64", "achieved": lens_tflops_overlapped, "peak": fp64_peak_tflops, "unit": "TFLOP/s",
        "frac": lens_tflops_overlapped / fp64_peak_tflops,
        "algorithmic_flop_per_launch": flop_lens, "avg_launch_ms": step_ms,
        "avg_launch_ms_note": "device time per step in the overlapped pass (CUDA events around the K steps / K): kernels of different steps "
                              "overlap, so a per-kernel duration does not exist in this regime; the lone durations are in roofline_one_stream",
        "fp64_pipe_busy": fp64_cycles / (step_ms * 1e-3 * sm_mhz * 1e6),
        "fp64_pipe_busy_source": "derived: smsp__inst_executed_pipe_fp64.sum of one step's launches (ncu, profiles/r02_*) x 2 cycles per warp "
                                 "instruction / (592 sub-partitions x SM clock x ms_per_step)",
        "traffic": NCU_TRAFFIC_LENS_1E7 * n / 1e7, "traffic_source": "profiles/r02_full_step.txt (ncu --set full, summed over the stage's launches)",
        "peak_source": "measured live: cmt_fp64_peak DFMA stream (no FP64 figure in MEASURED_PEAKS.json)",
        "dadd_peak_tops": dadd.value / 1e12,
    }
    roofline_one_stream = {
        "kernel": "lens_seg_kernel x 4 + tail_kernel (the lens stage of one step)", "regime": "one stream: pass A, nothing overlapped",
        "bound": "fp64", "achieved": lens_tflops, "peak": fp64_peak_tflops,
        "unit": "TFLOP/s", "frac": (lens_tflops / fp64_peak_tflops) if lens_tflops else None,
        "algorithmic_flop_per_launch": flop_lens, "avg_launch_ms": lens_ms,
        "share_of_step": lens_ms / (ms_seq / args.steps) if ms_seq > 0 else None,
        "timed_in": "pass A: the same K steps back to back on one stream (kernels not overlapped), CUDA events around the lens stage of every step",
    }
    walk_gbs = BYTES_PER_MOLECULE * n / (walk_ms * 1e-3) / 1e9 if walk_ms > 0 else None
    roofline_walk = {
        "kernel": "walk_kernel<ic>", "bound": "hbm", "achieved": walk_gbs, "peak": hbm_peak, "unit": "GB/s",
        "frac": (walk_gbs / hbm_peak) if walk_gbs else None, "traffic": NCU_TRAFFIC_WALK_1E7 * n / 1e7,
        "traffic_source": "profiles/r02_full_step.txt (ncu --set full)", "peak_source": hbm_src,
        "algorithmic_bytes_per_launch": BYTES_PER_MOLECULE * n, "avg_launch_ms": walk_ms,
        "share_of_step": walk_ms / (ms_seq / args.steps) if ms_seq > 0 else None,
        "algorithmic_flop_per_launch": flop_rows,
    }

    # ---- e2e: host buffers through the C ABI ----
    ic_host = torch.empty((6, n), dtype=torch.float64, pin_memory=True)
    ic_host.copy_(ic)
    fate_host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    cnt_host = np.zeros(len(prop.flat.fate_names), dtype=np.int64)
    work_host = np.zeros(8, dtype=np.int64)
    torch.cuda.synchronize()

    def e2e_step():
        cnt_host[:] = 0
        nat.check(lib.cmt_run_host_ic(prop.dev.handle, n, ic_host.data_ptr(), fate_host.data_ptr(), None,
                                      cnt_host.ctypes.data, work_host.ctypes.data))

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = world * n * args.steps / e2e_s
    e2e_ok = bool((cnt_host == counters).all()) if world == 1 else None
    # what the host link delivers for the same pinned buffer (one cudaMemcpyAsync of all six components)
    scratch = torch.empty_like(ic)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best_h2d = 0.0
    for _ in range(4):
        h0.record()
        scratch.copy_(ic_host, non_blocking=True)
        h1.record()
        torch.cuda.synchronize()
        best_h2d = max(best_h2d, ic_host.numel() * 8 / (h0.elapsed_time(h1) * 1e-3) / 1e9)
    del scratch
    e2e_gbs = (48 * n + n) * args.steps / e2e_s / 1e9

    # ---- e2e_philox: the run_simulation default path (device source, Counter back) ----
    def philox_step():
        cnt_host[:] = 0
        nat.check(lib.cmt_run_host_philox(prop.dev.handle, C.byref(source), seed, first, n, cnt_host.ctypes.data,
                                          work_host.ctypes.data))

    for _ in range(3):
        philox_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        philox_step()
    torch.cuda.synchronize()
    ph_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    philox_value = world * n * args.steps / ph_s
    philox_same = bool((cnt_host == counters).all()) if world == 1 else None

    # ---- e2e_api: the call a user of the reference makes, examples/lens_simulation_beamline.py:87-89 ----
    from trajectories.trajectory_simulator import TrajectorySimulator

    sim = TrajectorySimulator(device=local, seed=seed)
    api_steps = max(1, min(args.steps, 5))
    for _ in range(6):   # reach the steady state of the pinned result blocks (the first runs allocate them: ~45 ms each)
        sim.run_simulation(bl, "bench", N_traj=n, apertures_of_interest=["Detected"], n_jobs=10)
    # Every call builds ~5000 Python objects (2385 Molecule + Trajectory shells), so the interpreter's cyclic
    # collector runs a FULL collection every ~14 calls, and with torch imported a full collection walks ~1e6
    # long-lived objects: 35-40 ms, six calls' worth (profiles/diag_api.py).  The objects alive now are moved to the
    # permanent generation, as a long-running service would do after start-up; the calls themselves are untouched.
    import gc
    gc.collect()
    gc.freeze()
    barrier()
    api_calls = []
    t0 = time.perf_counter()
    for _ in range(api_steps):
        t1 = time.perf_counter()
        sim.run_simulation(bl, "bench", N_traj=n, apertures_of_interest=["Detected"], n_jobs=10)
        api_calls.append(round(1e3 * (time.perf_counter() - t1), 3))
    torch.cuda.synchronize()
    api_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    n_saved = len(sim.result.molecules)
    rows_bytes = sum(m.trajectory.n for m in sim.result.molecules) * 80
    api_value = n * api_steps / api_s      # under torch.distributed run_simulation shards N_traj itself

    # ---- CPU baseline (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        if affinity_before:
            os.sched_setaffinity(0, affinity_before)      # the CPU baseline gets every host core back
        r = cpu_rate(bl, vdist, xdist, target_s=12.0)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"{r['n']} molecules of the same workload ({r['seconds']:.1f} s): Philox source + propagation + Counter, oracle/cmt_oracle.c with OpenMP",
               "reference_python": None if args.no_reference_python else reference_python(r["cores"])}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": shared_config(n),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 48 * n, "d2h_bytes_per_step": n + 8 * (len(cnt_host) + 4),
                    "path": "cmt_run_host_ic: pinned host ICs -> H2D (2 streams, 2^21-molecule chunks) -> kernels -> fates + Counter D2H",
                    "counters_match_device_run": e2e_ok,
                    "roofline": {"bound": "pcie", "achieved": e2e_gbs, "peak": best_h2d, "unit": "GB/s", "frac": e2e_gbs / best_h2d if best_h2d else None,
                                 "peak_source": "measured live: one pinned-host to device copy of the step's 480 MB, best of 4 (per rank)"}},
            "e2e_philox": {"value": philox_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8 * (len(cnt_host) + 4),
                           "path": "cmt_run_host_philox (run_simulation default): Philox4x32-10 source on device, Counter D2H",
                           "counters_match_device_run": philox_same},
            "e2e_api": {"value": api_value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": rows_bytes + 8 * 16,
                        "path": "TrajectorySimulator.run_simulation(beamline, N_traj=n, apertures_of_interest=['Detected'], n_jobs=10): "
                                "Philox source, Counter, and the detected molecules' full trajectories back on the host (result.molecules: Molecule views of the row block, made on access)",
                        "saved_molecules_per_step": n_saved, "steps": api_steps, "ms_per_call": api_calls,
                        "gc": "gc.freeze() after the warm-up calls (a full collection of the interpreter's import-time objects costs 35-40 ms every ~14 calls otherwise)"},
            "contracted_math": contracted,
            "gpu_launches": launches,
            "value_one_stream": value_seq, "ms_per_step_one_stream": ms_seq / args.steps,
            "overlap": "none" if args.no_overlap else f"{prop.n_slots} streams: consecutive steps alternate streams (independent batches)"
                       + ("" if graphs is None else "; each step replays a CUDA graph (memset + walk + lens segments + tail)"),
            "roofline": roofline, "roofline_one_stream": roofline_one_stream, "roofline_walk": roofline_walk,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "host_binding": numa,
            "counters": dict(zip(prop.flat.fate_names, counters.tolist())),
            "work_per_step": {"ballistic_rows": int(work[0]), "lens_rk_steps": int(work[1]),
                              "table_out_of_range": int(work[2]), "lens_entries": int(work[3]),
                              "rk_steps_on_reference_path": int(work[4])},
            "kernel_ms_per_step": {"walk": walk_ms, "lens": lens_ms},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--molecules", type=float, default=N_PER_STEP, help="molecules per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-reference-python", action="store_true", help="skip timing the unmodified Python reference (baseline/_ref)")
    ap.add_argument("--no-overlap", action="store_true", help="issue every step on one stream")
    ap.add_argument("--no-contracted", action="store_true", help="skip the contracted-arithmetic pass")
    ap.add_argument("--slots", type=int, default=6, help="streams the overlapped steps alternate over")
    ap.add_argument("--no-graphs", action="store_true", help="launch the overlapped steps individually instead of replaying CUDA graphs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
