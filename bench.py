#!/usr/bin/env python
"""Benchmark of the SPAI reward hot path (BASELINE.json metric: candidate
patterns scored / s, row least-squares solves / s, % of the HBM roofline).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU port of the reference

One "step" scores one batch of B sampled patterns (trajectories) of the named
config. Default workload (N=1): cfg2 = 2-D 5-point Poisson 256x256
(n=65 536), candidate superset <= 8 per row (E = 524 280), B = 4096
trajectories per GPU, copy mode / fp32 — the semantics and precision of the
reference's own `PreconditionerEnv.update` (the reference arm computes exactly
this). `--mode ls` switches the headline to the north-star least-squares
re-solve, which the reference cannot run.

Prints ONE JSON line on rank 0 (see the contract in the task statement).
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "spai_patterns_scored_per_s"
UNIT = "patterns/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the grid (testing only)")
    ap.add_argument("--mode", default="copy", choices=["copy", "ls", "ls_gram"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--batch", type=int, default=0, help="trajectories per GPU (0 = config default)")
    ap.add_argument("--cpu-sample", type=int, default=96, help="trajectories in the CPU-baseline sample")
    ap.add_argument("--no-extras", action="store_true", help="skip the ls-mode side measurements")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sampler", action="store_true", help="skip the sampler (K4 / K4g) side measurements")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg3/cfg4/cfg5 block of the JSON line")
    ap.add_argument("--configs", default="cfg3,cfg4,cfg5", help="side configs measured after the headline")
    ap.add_argument("--cfg5-batch", type=int, default=16384, help="cfg5 global batch (strong-scaled over the ranks)")
    ap.add_argument("--no-lengths", action="store_true",
                    help="headline without caller-supplied row lengths (the kernels then stream the -1 padding too)")
    ap.add_argument("--max-frac", type=float, default=0.0,
                    help="trajectory b deletes floor(u_b*E*max_frac) edges (0 = 0.5, SURVEY 8d)")
    return ap.parse_args()


CFG_DESC = {"cfg1": "2-D Poisson 10x10", "cfg2": "2-D 5-pt Poisson 256x256", "cfg3": "3-D 7-pt Poisson 64^3",
            "cfg4": "2-D convection-diffusion 512x512", "cfg5": "banded+power-law n=1M"}


def workload_name(cfg, p, batch, mode, dtype):
    return f"{cfg}: {CFG_DESC[cfg]} n={p.n} E={p.num_edges} (<= {p.k}/row), B={batch}/GPU, {mode}/{dtype}"


def config_dict(args, p, batch, world, max_frac):
    """The SAME dict on both arms (b200 and reference): static facts of the workload only."""
    return {"workload": workload_name(args.config, p, batch, args.mode, args.dtype), "mode": args.mode, "n": p.n,
            "num_edges": p.num_edges, "batch_per_gpu": batch, "alpha": 0.5, "max_deleted_fraction": max_frac,
            "trajectories": "SURVEY 8d: trajectory b deletes floor(u_b*E*max_deleted_fraction) distinct edges "
                            "(torch.randperm, seed 1000+b), then the terminal id; int64 [B, T], -1 padded",
            "parallelism": f"trajectory-sharded x{world}",
            "l2": "256 MiB flush between timed iterations; actions exceed L2",
            "timing": "CUDA events per step on the launch stream, max over ranks"}


# ---------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------
class ClockSampler:
    """SM clock + throttle reasons sampled every ~5 ms DURING the timed region
    (NVML in a thread; falls back to one nvidia-smi query)."""

    def __init__(self, index: int):
        self.index = index
        self.sm, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        self.gate = True            # samples are recorded only while the gate is open (timed regions)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None
            return
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        nv = self._nvml
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            if not self.gate:
                time.sleep(0.001)
                continue
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, nm in names.items():
                    if bits & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join(timeout=1)
        if not self.sm:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=10).stdout.split(",")
                self.sm = [float(out[0])]
                self.max_mhz = float(out[1])
            except Exception:
                pass
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ---------------------------------------------------------------------------
# synthetic trajectories on the device (seeded per global trajectory index)
# ---------------------------------------------------------------------------
def device_trajectories(num_edges, batch, first, device, max_frac=0.5):
    """(actions int64[batch, T], lengths int32[batch]) on the device; trajectory `first + b` is seeded
    1000 + first + b (a shard of a batch equals the same rows of the full batch). lengths[b] counts the
    deleted edges and the terminal id; the rest of the row is -1 padding."""
    import torch
    lens = []
    gens = []
    for b in range(batch):
        g = torch.Generator(device=device)
        g.manual_seed(1000 + first + b)
        u = float(torch.rand(1, generator=g, device=device))
        lens.append(int(np.floor(u * num_edges * max_frac)))
        gens.append(g)
    tmax = max(lens) + 1
    acts = torch.full((batch, tmax), -1, dtype=torch.int64, device=device)
    for b in range(batch):
        t = lens[b]
        if t:
            acts[b, :t] = torch.randperm(num_edges, generator=gens[b], device=device)[:t]
        acts[b, t] = num_edges
    return acts, torch.tensor([t + 1 for t in lens], dtype=torch.int32, device=device)


def pin_to_gpu_numa_node(index):
    """Pin this process (and so its pinned allocations and the library's host threads) to the CPUs of the
    GPU's NUMA node: at N > 1 every rank then streams its actions out of socket-local memory."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"numa_node": node, "note": "the platform reports no NUMA affinity for this GPU"}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"numa_node": node, "cpus": len(cpus)}
        return {"numa_node": node, "cpus": 0}
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _ncu_side_facts(kernel_key):
    """Issue-slot / DRAM utilisation of the kernel from the committed ncu capture (profiles/traffic.json)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return {"issue_active_pct": d.get("_issue_active_pct", {}).get(kernel_key),
                "dram_frac_of_measured_peak": d.get("_dram_frac_of_measured_peak", {}).get(kernel_key),
                "source": d.get("_source")}
    except Exception:
        return None


def ncu_traffic(kernel_key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(kernel_key)
    except Exception:
        return None


# ---------------------------------------------------------------------------
# CPU arms (oracle port of the reference; bench.py may execute oracle/ only here)
# ---------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(cfg, scale):
    from gflownet_spai_b200 import synth
    p = synth.make_problem(cfg, scale)
    _W["p"] = p
    _W["a32"] = p.a.astype(np.float32)
    _W["v32"] = p.edge_val.astype(np.float32)


def _cpu_worker(args):
    from gflownet_spai_b200 import synth
    from oracle import spai_oracle as orc
    first, count = args
    p = _W["p"]
    acts = synth.make_trajectories(p.num_edges, count, first=first)
    res0, flops0 = _W.get("base") or orc.baseline_constants(_W["a32"], None, np.float32)
    _W["base"] = (res0, flops0)
    t0 = time.perf_counter()
    out = []
    for b in range(count):
        kept = orc.kept_edge_mask(p.num_edges, acts[b])
        m = orc.build_pattern_matrix(p.n, p.edge_row, p.edge_col, _W["v32"], kept, dtype=np.float32)
        res = orc.residual_copy(m, _W["a32"], dtype=np.float32)
        out.append(orc.reward_from_residual(res, m.nnz, p.n, res0, flops0, 0.5))
    return time.perf_counter() - t0, out


def cpu_port_throughput(cfg, scale, sample, procs):
    """patterns/s of the numpy/scipy restatement of PreconditionerEnv.update
    (preconditioner.py:32-52, copy mode, fp32) on `procs` host processes."""
    import multiprocessing as mp
    if procs <= 1:
        _cpu_worker_init(cfg, scale)
        _cpu_worker((0, 1))                      # baseline constants + warm caches
        dt, _ = _cpu_worker((0, sample))
        return sample / dt, dt
    ctx = mp.get_context("fork")
    per = max(1, sample // procs)
    with ctx.Pool(procs, initializer=_cpu_worker_init, initargs=(cfg, scale)) as pool:
        pool.map(_cpu_worker, [(0, 1)] * procs)  # warm every worker
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(i * per, per) for i in range(procs)])
        dt = time.perf_counter() - t0
    return per * procs / dt, dt


def reference_itself(p, sample):
    """patterns/s of the UNMODIFIED reference (`PreconditionerEnv.update`, preconditioner.py:32-52, staged
    under baseline/_ref by __graft_entry__.build()): as is (gc.collect() after every trajectory, :51) and
    with gc.collect patched out (SURVEY 8d items i / ii). None when the staged copy is absent."""
    import gc
    try:
        import torch
        from oracle import ref_shim
        if not ref_shim.reference_available():
            return None
        from gflownet_spai_b200 import synth
        coo = p.a.tocoo()
        acts = synth.make_trajectories(p.num_edges, sample)
        torch.set_num_threads(os.cpu_count() or 1)
        out = {"threads": torch.get_num_threads(), "sample": f"{sample} trajectories of the workload",
               "source": ref_shim.REFERENCE_ROOT}
        t0 = time.perf_counter()
        ref_shim.reference_update(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data, acts, 0.5)
        out["as_is_patterns_per_s"] = sample / (time.perf_counter() - t0)
        real = gc.collect
        mods = [m for m in ref_shim.load_reference().modules if hasattr(m, "gc")]
        try:
            for m in mods:
                m.gc = type("NoGc", (), {"collect": staticmethod(lambda *a, **k: 0)})
            t0 = time.perf_counter()
            ref_shim.reference_update(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data, acts, 0.5)
            out["minus_gc_patterns_per_s"] = sample / (time.perf_counter() - t0)
        finally:
            for m in mods:
                m.gc = gc
            gc.collect = real
        return out
    except Exception as exc:            # report, never hide
        return {"error": f"{type(exc).__name__}: {exc}"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    from gflownet_spai_b200 import synth
    p = synth.make_problem(args.config, args.scale)
    batch = args.batch or p.batch
    max_frac = args.max_frac or 0.5
    per_step = max(procs, min(args.cpu_sample, 4 * procs))
    vals = []
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    per = max(1, per_step // procs)
    with ctx.Pool(procs, initializer=_cpu_worker_init, initargs=(args.config, args.scale)) as pool:
        for _ in range(max(1, args.warmup)):
            pool.map(_cpu_worker, [(i * per, per) for i in range(procs)])
        t_all = time.perf_counter()
        for s in range(args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [((s * procs + i) * per, per) for i in range(procs)])
            vals.append(time.perf_counter() - t0)
        total = time.perf_counter() - t_all
    value = per * procs * args.steps / total
    sample = (f"each step = {per * procs} trajectories of the workload (of B={batch}) over {procs} processes; "
              "numpy/scipy restatement of the reference's update() (no gc.collect, no Python set loop)")
    itself = reference_itself(p, 3 if p.n > 10000 else 16) if args.mode == "copy" else None
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic", "config": config_dict(args, p, batch, world, max_frac),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample},
        "reference_itself": itself,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------
KNAMES = {"copy": "k3_copy_kernel", "ls": "k2_ls_kernel", "ls_gram": "k2g_solve_kernel"}


def algorithmic_bytes(ctx, p, acts_dev, mode, wbytes):
    """SURVEY.md §8d gather-inclusive G summed over the batch, from the inputs
    alone: for every kept candidate (i, c): 4 (pattern col id) + w (M value, copy
    mode) + 2*4 (rowptr of A[c,:]) + nnz(A[c,:]) * (4 + w); + 8 per pattern."""
    import torch
    rlen = np.diff(p.a.indptr)[p.edge_col].astype(np.float64)
    cost = 4.0 + (wbytes if mode == "copy" else 0.0) + 8.0 + rlen * (4.0 + wbytes)
    cost_t = torch.from_numpy(cost).to(acts_dev.device)
    total = 0.0
    step = max(1, min(256, int(2e9 // (8 * max(1, p.num_edges)))))
    for b0 in range(0, acts_dev.shape[0], step):
        kept = ctx.kept_mask(acts_dev[b0:b0 + step])
        total += float((kept.to(torch.float64) @ cost_t).sum())
        del kept
    return total + 8.0 * acts_dev.shape[0]


def reward_kernel_name(ctx, p, mode, B, tmax, dtype=None):
    """Which reward kernel the library picks for this call (mirrors eval_masks in spai_b200.cu)."""
    info = ctx.info()
    k = info.max_row_slots
    if mode == "copy":
        if tmax * 40 <= p.num_edges:
            return "k3s_sparse_kernel"
        if k <= 8 and B >= 64:
            return "k3t_lookup_kernel"
        if (dtype in ("f32", None) and 8 < k <= 32 and B >= 64 and not info.has_duplicates
                and os.environ.get("SPAI_K3_MMA", "1") != "0"):
            return "k3m_kernel(tcgen05)"
        return "k3_copy_kernel"
    if k <= 8 and B >= 64:
        return "k3t_lookup_kernel(ls table)"
    return KNAMES[mode]


def mask_kernel_name(p):
    return "k0_mask_build_smem_kernel" if (p.num_edges + 31) // 32 * 4 <= 100 * 1024 else "k0b_sort_kernel+k0b_build2_kernel"


def phase_times(ctx, call, flush, nrep):
    """Per-phase CUDA-event times of the library (masks / transpose+popcount / reward kernel / finalize), ms."""
    ctx.enable_timing(True)
    ph = {"masks": 0.0, "transpose": 0.0, "reward": 0.0, "finalize": 0.0}
    for _ in range(nrep):
        flush.zero_()
        call()
        tm = ctx.last_timing()
        ph["masks"] += tm.ms_masks / nrep
        ph["transpose"] += tm.ms_transpose / nrep
        ph["reward"] += tm.ms_reward / nrep
        ph["finalize"] += tm.ms_finalize / nrep
    ctx.enable_timing(False)
    return ph


def oracle_check(p, acts_rows, got_reward, mode, dtype):
    """In-run parity: timed trajectories re-scored by the CPU oracle (copy mode; ls on small n)."""
    from oracle import spai_oracle as orc
    npdt = np.float32 if dtype == "f32" else np.float64
    if mode == "copy":
        want = orc.reward_batch_copy(p.n, p.edge_row, p.edge_col, p.edge_val.astype(npdt), p.a.astype(npdt),
                                     acts_rows, 0.5, dtype=npdt)["reward"]
    elif p.n <= 70000:
        want = orc.reward_batch_ls(p.n, p.edge_row, p.edge_col, p.a, acts_rows, 0.5, dtype=npdt,
                                   baseline_dtype=npdt)["reward"]
    else:
        return None
    err = np.abs(got_reward - want) / np.maximum(np.abs(want), 1e-12)
    return {"trajectories": int(acts_rows.shape[0]), "max_rel_err_vs_oracle": float(np.max(err)),
            "tolerance": 1e-4 if dtype == "f32" else 1e-10}


def side_config(name, args, rank, world, dev, local, flush, peak):
    """One entry of the `configs` block: a BASELINE config other than the headline, measured on this
    rank's shard with the same rules (warm-up, CUDA events, inputs larger than L2 / L2 flush, max over
    ranks), per-kernel shares, compulsory-byte fraction and an in-run oracle check.
      cfg3  B = 4096 per GPU (weak, trajectory-sharded), copy/fp32
      cfg4  B = 1024 per GPU (weak), copy fp32 + fp64 and ls_gram fp64
      cfg5  B = 16 384 GLOBAL, strong-scaled (16 384 / N per rank), copy/fp32, SURVEY 8d trajectories"""
    import torch
    import torch.distributed as dist
    from gflownet_spai_b200 import synth
    from gflownet_spai_b200.env import SpaiContext

    t_host = time.perf_counter()
    p = synth.make_problem(name, args.scale)
    coo = p.a.tocoo()
    ctx = SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data, device=local)
    setup_s = time.perf_counter() - t_host
    e = p.num_edges
    W = (e + 31) // 32
    strong = name == "cfg5"
    if strong:
        total = max(world, int(args.cfg5_batch * min(1.0, args.scale * args.scale)) if args.scale < 1 else args.cfg5_batch)
        lo, hi = rank * total // world, (rank + 1) * total // world
    else:
        per = args.batch or {"cfg3": 4096, "cfg4": 1024}.get(name, p.batch)
        if args.scale < 1:
            per = max(64, int(per * args.scale))
        total, lo, hi = per * world, rank * per, (rank + 1) * per
    variants = [("copy", "f32")]
    if name == "cfg4":
        variants += [("copy", "f64"), ("ls_gram", "f64")]
    # trajectory chunks that fit next to the workspace (int64 [chunk, T], T up to E/2 + 1)
    tmax_bound = e // 2 + 2
    chunk = max(32, min(hi - lo, int((40 << 30) // (8 * tmax_bound)) // 32 * 32))
    res = {"n": p.n, "num_edges": e, "global_batch": total, "batch_this_rank": hi - lo,
           "scaling": "strong" if strong else "weak", "chunk_trajectories": chunk, "setup_s": setup_s,
           "mask_kernel": mask_kernel_name(p), "variants": {}}
    tot_ms = {v: 0.0 for v in variants}
    ph_acc = {v: None for v in variants}
    valid_ids = 0
    tmax_seen = 0
    parity = {}
    first = True
    gen_s = 0.0
    chunk_log = []
    sampler = ClockSampler(local)
    sampler.gate = False
    sampler.start()
    for c0 in range(lo, hi, chunk):
        c1 = min(hi, c0 + chunk)
        tg = time.perf_counter()
        acts, lens = device_trajectories(e, c1 - c0, c0, dev, 0.5)
        torch.cuda.synchronize()
        gen_s += time.perf_counter() - tg
        valid_ids += int(lens.sum())
        tmax_seen = max(tmax_seen, int(acts.shape[1]))
        for (mode, dt) in variants:
            tdt = torch.float32 if dt == "f32" else torch.float64
            sub, sl = acts, lens
            if mode == "ls_gram" and name == "cfg4":          # ~0.4 s per 1024 patterns: time a 256-pattern slice
                sub, sl = acts[:256], lens[:256]
            call = lambda: ctx.reward_batch(sub, 0.5, mode, tdt, want=("reward",), lengths=sl)
            if first:                                          # warm-up: plans, tables, workspace, clocks
                for _ in range(2):
                    out = call()
                torch.cuda.synchronize()
                if rank == 0 and mode == "copy":
                    try:
                        parity[f"{mode}/{dt}"] = oracle_check(p, sub[:1].cpu().numpy(), out["reward"][:1].cpu().numpy(), mode, dt)
                    except Exception as exc:
                        parity[f"{mode}/{dt}"] = {"error": f"{type(exc).__name__}: {exc}"}
                ph_acc[(mode, dt)] = phase_times(ctx, call, flush, 2)
            # the host-side trajectory generation leaves the GPU idle for seconds and its clocks drop to
            # idle (observed: one 32 ms call took 637 ms at 120 MHz): bring them back before timing
            # (three 96-trajectory calls were not enough: at N = 2 two of one rank's nine cfg5 chunks measured 1.4x and 2.4x the
            # others) -> one full untimed pass over the chunk, then the timed one
            warm = lambda: ctx.reward_batch(sub[:96], 0.5, mode, tdt, want=("reward",), lengths=sl[:96])
            for _ in range(2):
                warm()
            call()
            flush.zero_()
            torch.cuda.synchronize()
            reps = 3 if not strong else 1
            ms = 0.0
            for _ in range(reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                sampler.gate = True
                e0.record()
                call()
                e1.record()
                torch.cuda.synchronize()
                sampler.gate = False
                ms += e0.elapsed_time(e1) / reps
            scale_up = (acts.shape[0] / sub.shape[0])
            tot_ms[(mode, dt)] += ms * scale_up
            chunk_log.append((f"{mode}/{dt}", int(acts.shape[0]), int(acts.shape[1]), round(ms * scale_up, 3)))
        first = False
        del acts, lens
    for (mode, dt) in variants:
        t = torch.tensor([tot_ms[(mode, dt)]], dtype=torch.float64, device=dev)
        ids = torch.tensor([float(valid_ids)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(ids, op=dist.ReduceOp.SUM)
        ms = float(t)
        per_rank = [tot_ms[(mode, dt)]]
        if world > 1:
            allms = [None] * world
            dist.all_gather_object(allms, tot_ms[(mode, dt)])
            per_rank = [float(x) for x in allms]
        ph = ph_acc[(mode, dt)] or {}
        comp = float(ids) * 8.0 + total * (W * 4.0 + 8.0)
        k0_bytes_rank = valid_ids * 8.0 + (hi - lo) * W * 4.0
        share_tot = max(sum(ph.values()), 1e-9) if ph else 1.0
        k0_ms_rank = tot_ms[(mode, dt)] * (ph.get("masks", 0.0) / share_tot) if ph else None
        res["variants"][f"{mode}/{dt}"] = {
            "ms_per_step": ms, "patterns_per_s": total / (ms / 1e3), "row_solves_per_s": total * p.n / (ms / 1e3),
            "per_rank_ms": per_rank,
            "kernel_share": {k: v / share_tot for k, v in ph.items()} if ph else None,
            "reward_kernel": reward_kernel_name(ctx, p, mode, min(chunk, hi - lo), tmax_seen, dt),
            "compulsory_bytes_per_step": comp, "compulsory_frac_of_hbm_peak": comp / (ms / 1e3) / 1e9 / peak,
            "k0_valid_bytes_gbps": (k0_bytes_rank / (k0_ms_rank / 1e3) / 1e9) if k0_ms_rank else None,
            "k0_frac_of_hbm_peak_on_valid_bytes": (k0_bytes_rank / (k0_ms_rank / 1e3) / 1e9 / peak) if k0_ms_rank else None,
            "note": ("ls_gram timed on 256 of every 1024 patterns and scaled" if (mode == "ls_gram" and name == "cfg4") else None),
        }
    if world > 1:
        logs = [None] * world
        dist.all_gather_object(logs, chunk_log[:24])
        res["chunks_per_rank"] = logs
    else:
        res["chunks_per_rank"] = [chunk_log[:24]]
    # the path without id lists: K4g taken-bitmask -> spai_reward_from_taken_dev (neither K0 nor K0b runs); bounded sample
    if rank == 0:
        try:
            nb = int(min(1024, hi - lo))
            g = torch.Generator(device=dev)
            g.manual_seed(5)
            lg = (torch.randn(e + 1, generator=g, device=dev) * 0.25).contiguous()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ctx.sample_taken(lg, 64, 1, 0)
            tk, ln = ctx.sample_taken(lg, nb, 99, 0)
            ctx.reward_from_taken(tk, 0.5, "copy", torch.float32)
            torch.cuda.synchronize()
            flush.zero_()
            ev[0].record()
            tk, ln = ctx.sample_taken(lg, nb, 99, 0)
            ev[1].record()
            ctx.reward_from_taken(tk, 0.5, "copy", torch.float32)
            ev[2].record()
            torch.cuda.synchronize()
            t_s, t_r = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
            res["sample_to_reward(K4g taken-bitmask, no id lists)"] = {
                "batch": nb, "sample_ms": t_s, "reward_ms": t_r, "patterns_per_s": nb / ((t_s + t_r) / 1e3),
                "mean_deleted_fraction": float(ln.float().mean()) / (e + 1)}
            del tk, ln, lg
        except Exception as exc:
            res["sample_to_reward(K4g taken-bitmask, no id lists)"] = {"error": f"{type(exc).__name__}: {exc}"}
    clk = sampler.stop()
    if world > 1:
        allc = [None] * world
        dist.all_gather_object(allc, clk)
        sm = [c["sm_mhz"] for c in allc if c and c["sm_mhz"]]
        clk = {"sm_mhz": min(sm) if sm else None, "sm_max_mhz": clk["sm_max_mhz"],
               "reasons": sorted({r for c in allc if c for r in c["reasons"]}), "samples": sum(c["samples"] for c in allc if c)}
    res["clocks"] = clk
    res["max_trajectory_len"] = tmax_seen
    res["trajectory_generation_s_untimed"] = gen_s
    res["parity_check"] = parity or None
    res["l2"] = "inputs per call (valid ids * 8 B) exceed L2; 256 MiB flush between timed calls"
    ctx.close()
    del ctx
    torch.cuda.empty_cache()
    return res


def sampler_bench(ctx, p, dev, batch, flush):
    """North-star kernel (4) and SURVEY 8f-1 at the headline config's size: trajectories drawn per second
    by the whole-trajectory kernels (K4g: taken-bitmask + length, then the ordered ids), the masked
    categorical single step (K4) per launch, and sample -> reward with everything on the device."""
    import torch
    a = p.num_edges + 1
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    logits = (torch.randn(a, generator=g, device=dev) * 0.25).contiguous()     # near-uniform policy (early training)
    out = {"A": a, "batch": batch, "policy": "logits ~ N(0, 0.25^2), one vector per epoch (graph and weights are constant "
                                             "inside sample_states, gflownet.py:164-172)"}

    def ev_time(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.median(ms)), r

    ms_t, (taken, length) = ev_time(lambda: ctx.sample_taken(logits, batch, 12345, 0))
    ids = int(length.sum())
    ld = int(length.max())
    ms_o, acts = ev_time(lambda: ctx.sample_order(logits, length, 12345, 0, dtype=torch.int32, ld=ld))
    ms_r, _ = ev_time(lambda: ctx.reward_from_taken(taken, 0.5, "copy", torch.float32))
    # spot check inside the run: one row's order against a device-side sort of the exported keys
    _, l1, k1 = ctx.sample_taken(logits, 1, 12345, 0, export_keys=True)
    sel = torch.nonzero(k1[0, : a - 1] < 0).squeeze(1)
    want = sel[torch.sort(k1[0, sel], stable=True).indices]
    ok = bool(torch.equal(acts[0, : int(l1[0]) - 1].long(), want)) and int(acts[0, int(l1[0]) - 1]) == a - 1
    out["whole_trajectory(K4g)"] = {
        "taken_ms": ms_t, "order_ms": ms_o, "trajectories_per_s": batch / ((ms_t + ms_o) / 1e3),
        "trajectories_per_s_mask_only": batch / (ms_t / 1e3), "ids_drawn_per_step": ids, "max_length": ld,
        "ids_ordered_per_s": ids / (ms_o / 1e3), "keys_generated_per_s": 1.0 * batch * a / (ms_t / 1e3),
        "actions_bytes": int(acts.numel()) * 4, "order_check_row0": ok,
        "sample_plus_reward_ms": ms_t + ms_r, "reward_from_taken_ms": ms_r,
        "sample_plus_reward_patterns_per_s": batch / ((ms_t + ms_r) / 1e3)}
    del acts
    # K4: one masked categorical step for every sample (three passes over the A logits per sample)
    words = (a + 31) // 32
    tk = torch.zeros((batch, words), dtype=torch.int32, device=dev)
    done = torch.zeros(batch, dtype=torch.uint8, device=dev)
    act = torch.empty(batch, dtype=torch.int64, device=dev)
    pr = torch.empty(batch, dtype=torch.float32, device=dev)
    u = torch.rand(batch, generator=g, device=dev)
    ms_s, _ = ev_time(lambda: ctx.sample_step(logits, tk, u, done, act, pr), reps=5)
    out["single_step(K4)"] = {"ms_per_step": ms_s, "sample_steps_per_s": batch / (ms_s / 1e3),
                              "logit_bytes_read_per_step": 3.0 * batch * a * 4,
                              "gbps_on_logit_reads(L2)": 3.0 * batch * a * 4 / ms_s / 1e6,
                              "note": "a whole trajectory needs T ~ A/2 of these: use the whole-trajectory kernels"}
    # K4p: the same step with running block sums, whole trajectories in ONE launch (a warp per sample)
    try:
        bp = min(batch, 1024)
        tk2 = torch.zeros((bp, words), dtype=torch.int32, device=dev)
        dn2 = torch.zeros(bp, dtype=torch.uint8, device=dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.sample_steps(logits, torch.zeros((8, words), dtype=torch.int32, device=dev),
                         torch.zeros(8, dtype=torch.uint8, device=dev), 64, seed=1)          # warm-up (module load)
        torch.cuda.synchronize()
        e0.record()
        acts2, _, steps2 = ctx.sample_steps(logits, tk2, dn2, a, seed=777, want_probs=True)
        e1.record()
        torch.cuda.synchronize()
        ms_p = e0.elapsed_time(e1)
        nst = int(steps2.sum())
        out["multi_step(K4p)"] = {"batch": bp, "ms": ms_p, "steps_drawn": nst, "sample_steps_per_s": nst / (ms_p / 1e3),
                                  "trajectories_per_s": bp / (ms_p / 1e3), "all_done": bool(dn2.all()),
                                  "longest_trajectory": int(steps2.max()),
                                  "speedup_per_step_vs_K4": (nst / (ms_p / 1e3)) / (batch / (ms_s / 1e3))}
        del acts2
    except Exception as exc:
        out["multi_step(K4p)"] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from gflownet_spai_b200 import synth
    from gflownet_spai_b200.env import SpaiContext

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    # cpu_baseline (rank 0, N = 1), same convention as the reference arm: the port on ALL host cores, bounded sample, with the
    # single-core figure next to it. Measured BEFORE this process creates its CUDA context (the pool forks).
    cpu_pre = None
    if rank == 0 and world == 1:
        procs = max(1, min(os.cpu_count() or 1, 64))
        n1 = max(2, min(args.cpu_sample, 32))
        v1, dt1 = cpu_port_throughput(args.config, args.scale, n1, 1)
        if procs > 1 and args.cpu_sample >= 8:
            nall = max(procs, min(4 * procs, 4 * args.cpu_sample))
            v, dt = cpu_port_throughput(args.config, args.scale, nall, procs)
        else:
            v, dt, procs, nall = v1, dt1, 1, n1
        cpu_pre = {"value": v, "unit": UNIT, "cores": procs, "kind": "port", "single_core_value": v1,
                   "sample": f"{nall} trajectories of the same workload, copy/fp32, numpy/scipy oracle port over {procs} processes "
                             f"({dt:.1f} s); single process: {v1:.1f} patterns/s ({n1} trajectories, {dt1:.1f} s); "
                             "`--impl reference` adds the staged reference itself"}
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        # NCCL prints its version banner on stdout when NCCL_DEBUG is set in the environment: keep stdout for the ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    tdtype = torch.float32 if args.dtype == "f32" else torch.float64
    wbytes = 4.0 if args.dtype == "f32" else 8.0

    p = synth.make_problem(args.config, args.scale)
    batch = args.batch or p.batch
    coo = p.a.tocoo()
    t0 = time.perf_counter()
    ctx = SpaiContext(p.n, p.edge_row, p.edge_col, p.edge_val, coo.row, coo.col, coo.data, device=local)
    setup_s = time.perf_counter() - t0
    max_frac = args.max_frac or 0.5
    acts, lens = device_trajectories(p.num_edges, batch, rank * batch, dev, max_frac)     # weak scaling: B per GPU
    B, T = acts.shape
    use_len = None if args.no_lengths else lens
    valid_ids = int(lens.sum())
    gathered = torch.empty(world * B, dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)         # > 126 MB L2

    def step_dev():
        out = ctx.reward_batch(acts, 0.5, args.mode, tdtype, lengths=use_len)
        if world > 1 and not os.environ.get("SPAI_BENCH_NO_GATHER"):
            dist.all_gather_into_tensor(gathered, out["reward"])
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step_dev()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()                                   # every rank samples its own GPU
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.zero_()                                  # evict L2 between timed iterations
        ev[s][0].record()
        out = step_dev()
        ev[s][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    # the timed region lasts ~20 ms and the NVML thread catches 1-10 samples in it: the same step keeps running untimed
    # (a fixed count, so every rank issues the same collectives) while the sampler goes on; both counts are reported
    n_in_region = len(sampler.sm)
    for _ in range(2 * args.steps):
        flush.zero_()
        step_dev()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["samples_in_timed_region"] = n_in_region
    if world > 1:
        allc = [None] * world
        dist.all_gather_object(allc, clocks)
        sm = [c["sm_mhz"] for c in allc if c and c["sm_mhz"]]
        clocks = {"sm_mhz": min(sm) if sm else None, "sm_max_mhz": clocks["sm_max_mhz"],
                  "reasons": sorted({r for c in allc if c for r in c["reasons"]}),
                  "samples": sum(c["samples"] for c in allc if c),
                  "samples_in_timed_region": sum(c.get("samples_in_timed_region", 0) for c in allc if c),
                  "per_rank_sm_mhz": [c["sm_mhz"] if c else None for c in allc]}
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    if os.environ.get("SPAI_BENCH_DEBUG"):
        print(f"[rank {rank}] T={T} valid={valid_ids} clocks={clocks}", file=sys.stderr, flush=True)
        print(f"[rank {rank}] per-step ms: {[round(a.elapsed_time(b), 3) for a, b in ev]}", file=sys.stderr, flush=True)
    launches = ctx.last_timing().launches * args.steps
    t_ms = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    dev_ms = float(t_ms)
    value = world * B * args.steps / (dev_ms / 1e3)

    # ---- per-kernel share + roofline (timing pass, device-resident inputs)
    nrep = max(3, min(args.steps, 10))
    ph = phase_times(ctx, lambda: ctx.reward_batch(acts, 0.5, args.mode, tdtype, lengths=use_len), flush, nrep)
    W = (p.num_edges + 31) // 32
    read_ids = float(valid_ids) if use_len is not None else float(B) * T
    k0_bytes = read_ids * 8 + float(B) * W * 4            # ids streamed once + the bitmask written once
    k0_bytes_padded = float(B) * T * 8 + float(B) * W * 4  # what a caller without row lengths makes K0 read
    peak, peak_src = measured_peak()
    reward_kernel = reward_kernel_name(ctx, p, args.mode, B, T, args.dtype)
    mask_kernel = mask_kernel_name(p)
    info0 = ctx.info()
    rec_bytes = 16.0 * info0.contributions + 16.0 * p.n
    table = reward_kernel.startswith("k3t") or reward_kernel.startswith("k3s")
    reward_comp = float(B) * W * 4 + 8.0 * B + (p.n * 256.0 * wbytes if reward_kernel.startswith("k3t") else rec_bytes)
    kern = {
        f"k0_masks({mask_kernel}: actions->kept bitmask)": {"ms": ph["masks"], "algorithmic_bytes": k0_bytes,
                                                             "padded_bytes": k0_bytes_padded},
        "k0_transpose+popcount": {"ms": ph["transpose"], "algorithmic_bytes": 2.0 * B * W * 4},
        reward_kernel: {"ms": ph["reward"], "algorithmic_bytes": reward_comp,
                        "bytes_kind": "compulsory HBM bytes (kept-mask words once + table/plan once + sums)"},
        "k3_finalize": {"ms": ph["finalize"], "algorithmic_bytes": 16.0 * B},
    }
    if not table and rank == 0:
        # row-sweep / solve kernels: SURVEY 8d gather-inclusive G (exceeds DRAM traffic by the batch re-use)
        kern[reward_kernel]["gather_inclusive_G_bytes"] = algorithmic_bytes(ctx, p, acts, args.mode, wbytes)
        kern[reward_kernel]["G_gbps"] = kern[reward_kernel]["gather_inclusive_G_bytes"] / max(ph["reward"], 1e-9) / 1e6
    for v in kern.values():
        v["gbps"] = v["algorithmic_bytes"] / max(v["ms"], 1e-9) / 1e6
        v["frac_of_hbm_peak"] = v["gbps"] / peak
        v["share"] = v["ms"] / max(sum(ph.values()), 1e-9)
    dom = max(kern, key=lambda k: kern[k]["ms"])
    step_bytes = k0_bytes + 8.0 * B
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbps"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["gbps"] / peak, "traffic": ncu_traffic(dom.split("(")[0] if dom.startswith("k0_masks") else dom),
                "peak_source": peak_src,
                "launch_ms": kern[dom]["ms"], "algorithmic_bytes_per_launch": kern[dom]["algorithmic_bytes"],
                "padded_bytes_per_launch": kern[dom].get("padded_bytes"),
                "whole_step_compulsory_frac": step_bytes / (dev_ms / args.steps / 1e3) / 1e9 / peak,
                "ncu": _ncu_side_facts(dom),
                "note": "k0_masks: algorithmic bytes = the VALID action ids (sum_b T_b * 8 B; row lengths are passed, so "
                        "the -1 padding is never read) + the bitmask written once (B*W*4). Reward kernels: compulsory "
                        "HBM bytes (masks + table/plan once); their gather-inclusive G (SURVEY 8d) is printed "
                        "separately for the row-sweep kernels only, because the batch re-uses every gathered row from "
                        "shared memory and G/time exceeds the HBM peak by design."}

    # ---- the other collective of the training step: the policy-gradient all-reduce (N > 1)
    dp_training = None
    if world > 1:
        from gflownet_spai_b200.dist import allreduce_gradients
        hid, a_n = 4, p.num_edges + 1                       # GFlowNet100.py:180 hidden_dim; policy.py:30 fc = hid x A
        shapes = [(a_n, hid), (a_n,), (a_n, hid), (a_n,), (4 * hid, hid + 1), (4 * hid,)]   # fwd fc, bwd fc, LSTM
        params = [torch.nn.Parameter(torch.zeros(sh, device=dev)) for sh in shapes]
        for q in params:
            q.grad = torch.ones_like(q)
        for _ in range(3):
            nbytes = allreduce_gradients(params)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(10):
            allreduce_gradients(params)
        g1.record()
        barrier()
        t_g = torch.tensor([g0.elapsed_time(g1) / 10], dtype=torch.float64, device=dev)
        dist.all_reduce(t_g, op=dist.ReduceOp.MAX)
        dp_training = {"grad_bytes": int(nbytes), "allreduce_ms": float(t_g),
                       "algbw_gbps": nbytes / float(t_g) / 1e6, "reward_step_ms": dev_ms / args.steps,
                       "shapes": "forward + backward policy fc = hid x (E+1) with hid = 4 (GFlowNet100.py:180, policy.py:30), "
                                 "one flat bucket, NCCL all-reduce (dist.allreduce_gradients)"}
        del params

    # ---- end to end through the host entry point (pinned host actions in, rewards out)
    e2e = None
    if not args.no_e2e:
        ok, host = 1, None
        try:
            host = torch.empty((B, T), dtype=torch.int64, pin_memory=True)
            host.copy_(acts)
            host_len = lens.cpu()
            torch.cuda.synchronize()
        except Exception as exc:        # pinned-memory pressure with many ranks: report, never hide
            ok, e2e = 0, {"error": f"{type(exc).__name__}: {exc}"}
        if world > 1:                   # all ranks take the same branch (no collective mismatch)
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = int(flag)
        if ok:
            e2e = {}
            for key, hl in (("with_lengths", host_len), ("scan", None)):
                if key == "with_lengths" and args.no_lengths:
                    continue
                for _ in range(2):
                    r = ctx.reward_batch(host, 0.5, args.mode, tdtype, want=("reward",), lengths=hl)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                w0 = time.perf_counter()
                e0.record()
                nst = args.steps if key == "with_lengths" or args.no_lengths else max(2, args.steps // 4)
                for _ in range(nst):
                    r = ctx.reward_batch(host, 0.5, args.mode, tdtype, want=("reward",), lengths=hl)   # syncs: result on the host
                    if world > 1:
                        dist.all_gather_into_tensor(gathered, r["reward"].to(dev))
                e1.record()
                barrier()
                e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - w0))
                t_e = torch.tensor([e_ms], dtype=torch.float64, device=dev)
                if world > 1:
                    dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
                e2e[key] = {"value": world * B * nst / (float(t_e) / 1e3), "ms_per_step": float(t_e) / nst,
                            "h2d_bytes_per_step": int(ctx.last_timing().h2d_bytes), "steps": nst}
            # the platform's own ceiling in the same run: plain cudaMemcpyAsync of 2 GiB of the same pinned buffer on every
            # rank at once (no kernel of this repository). N ranks share the box's host links; the e2e rate is read against it.
            local_gbps, n_el = -1.0, 0
            try:
                flat = host.view(-1)
                n_el = min(flat.numel(), (2 << 30) // 8)
                scratch = torch.empty(n_el, dtype=torch.int64, device=dev)
                scratch.copy_(flat[:n_el], non_blocking=True)
                torch.cuda.synchronize()
            except Exception:            # no collective inside: every rank reaches the barriers below
                scratch = None
            barrier()
            if scratch is not None:
                p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p0.record()
                for _ in range(2):
                    scratch.copy_(flat[:n_el], non_blocking=True)
                p1.record()
                torch.cuda.synchronize()
                local_gbps = 2 * n_el * 8 / (p0.elapsed_time(p1) * 1e-3) / 1e9
            barrier()
            gbps = torch.tensor([local_gbps], dtype=torch.float64, device=dev)
            g_min, g_sum = gbps.clone(), gbps.clone()
            if world > 1:
                dist.all_reduce(g_min, op=dist.ReduceOp.MIN)
                dist.all_reduce(g_sum, op=dist.ReduceOp.SUM)
            probe = ({"min_gbps_per_gpu": float(g_min), "aggregate_gbps": float(g_sum), "bytes_per_rank": int(2 * n_el * 8),
                      "what": "plain cudaMemcpyAsync from the same pinned buffer, all ranks at once"}
                     if float(g_min) > 0 else {"error": "scratch allocation failed on a rank"})
            del scratch
            head = e2e.get("with_lengths") or e2e["scan"]
            e2e = {"value": head["value"], "unit": UNIT, "h2d_bytes_per_step": head["h2d_bytes_per_step"],
                   "d2h_bytes_per_step": int(B * 8), "input_bytes_per_step": int(B * T * 8),
                   "ms_per_step": head["ms_per_step"], "per_gpu_value": head["value"] / world,
                   "h2d_gbps_per_gpu": head["h2d_bytes_per_step"] / head["ms_per_step"] / 1e6,
                   "h2d_aggregate_gbps": world * head["h2d_bytes_per_step"] / head["ms_per_step"] / 1e6,
                   "platform_h2d_probe": probe,
                   "variants": e2e, "numa": numa,
                   "api": "SpaiContext.reward_batch(pinned host int64 actions[B,T], lengths=int32[B]) -> "
                          "spai_reward_batch_host_len: the valid prefix of every row streams over PCIe straight into "
                          "the mask kernel (zero-copy), nothing on the host touches the data; variant 'scan' = the "
                          "plain reference call without lengths (worker threads find the -1 padding)"}
        elif e2e is None:
            e2e = {"error": "pinned host allocation failed on another rank"}
        del host

    # ---- side measurements on the headline config (not the headline): other modes, padded input, int32 ids
    extras = {}

    def timed(call, reps=3):
        for _ in range(2):
            call()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            call()
        a1.record()
        torch.cuda.synchronize()
        return a0.elapsed_time(a1) / reps

    if rank == 0 and world == 1 and not args.no_extras and args.mode == "copy":
        for md, dt in (("ls", torch.float32), ("ls", torch.float64), ("ls_gram", torch.float32),
                       ("ls_gram", torch.float64), ("copy", torch.float64)):
            small = args.config in ("cfg1", "cfg2")
            sub, sl = (acts, use_len) if small else (acts[: min(B, 64)], None if use_len is None else use_len[: min(B, 64)])
            try:
                ms = timed(lambda: ctx.reward_batch(sub, 0.5, md, dt, lengths=sl))
                extras[f"{md}/{'f32' if dt == torch.float32 else 'f64'}"] = {
                    "patterns_per_s": sub.shape[0] / (ms / 1e3), "row_solves_per_s": sub.shape[0] * p.n / (ms / 1e3),
                    "batch": int(sub.shape[0]), "ms": ms}
            except Exception as exc:        # report, never hide
                extras[f"{md}/{dt}"] = {"error": str(exc)}
        try:
            ms = timed(lambda: ctx.reward_batch(acts, 0.5, args.mode, tdtype))
            extras["no row lengths (the -1 padding is streamed and ignored)"] = {"patterns_per_s": B / (ms / 1e3), "ms": ms}
            a32 = acts.to(torch.int32)
            ms = timed(lambda: ctx.reward_batch(a32, 0.5, args.mode, tdtype, lengths=lens))
            extras["int32 ids + row lengths (device-resident sampler format)"] = {"patterns_per_s": B / (ms / 1e3), "ms": ms}
            del a32
        except Exception as exc:
            extras["input formats"] = {"error": str(exc)}
        # early-training regime: short trajectories (64 deletions) -> most rows untouched -> incremental path
        try:
            g = torch.Generator(device=dev)
            g.manual_seed(7)
            short = torch.randint(0, p.num_edges, (B, 65), generator=g, device=dev, dtype=torch.int64)
            short[:, -1] = p.num_edges
            for md in ("copy", "ls", "ls_gram"):
                ms = timed(lambda: ctx.reward_batch(short, 0.5, md, tdtype), reps=5)
                extras[f"{md}/{args.dtype} short trajectories (64 deletions, untouched rows skipped)"] = {
                    "patterns_per_s": B / (ms / 1e3), "batch": int(B), "ms": ms}
        except Exception as exc:
            extras["short trajectories"] = {"error": str(exc)}

    cpu = None
    parity = None
    if rank == 0 and world == 1:
        # the CPU leg doubles as an in-run check that the timed kernels do the work:
        # two of the timed trajectories re-scored by the oracle
        from oracle import spai_oracle as orc
        sel = acts[:2].cpu().numpy()
        nsel = 2 if args.mode == "copy" else 1
        try:
            got = ctx.reward_batch(acts[:nsel], 0.5, args.mode, tdtype, lengths=None if use_len is None else use_len[:nsel])
            parity = oracle_check(p, sel[:nsel], got["reward"].cpu().numpy(), args.mode, args.dtype)
        except Exception as exc:
            parity = {"error": f"{type(exc).__name__}: {exc}"}
        # ls-mode CPU restatement (LAPACK lstsq per row), bounded sample: 2048 rows of one pattern
        try:
            kept = orc.kept_edge_mask(p.num_edges, sel[0])
            pat = orc.build_pattern_matrix(p.n, p.edge_row, p.edge_col, np.ones(p.num_edges), kept, np.float64)
            rows_s = np.linspace(0, p.n - 1, num=min(p.n, 2048), dtype=np.int64)
            t0 = time.perf_counter()
            for i in rows_s:
                orc.ls_row_residual2(p.a, int(i), np.unique(pat.indices[pat.indptr[i]:pat.indptr[i + 1]]))
            dt_ls = time.perf_counter() - t0
            extras["ls_cpu_port/f64"] = {"row_solves_per_s": rows_s.size / dt_ls,
                                         "patterns_per_s": rows_s.size / dt_ls / p.n, "cores": 1,
                                         "sample": f"{rows_s.size} rows of one pattern, numpy.linalg.lstsq"}
        except Exception as exc:
            extras["ls_cpu_port/f64"] = {"error": str(exc)}
        cpu = cpu_pre

    sampler_res = None
    if rank == 0 and world == 1 and not args.no_extras and not args.no_sampler:
        try:
            sampler_res = sampler_bench(ctx, p, dev, B, flush)
        except Exception as exc:        # report, never hide
            sampler_res = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()

    info = ctx.info()
    ctx_bytes = int(info.device_bytes)
    contributions, max_union = int(info.contributions), int(info.max_row_union)
    class_rows = [int(x) for x in info.ls_class_rows]
    ctx.close()
    del ctx, acts, lens
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs (each rank scores its shard; rank 0 reports)
    configs = {}
    if not args.no_configs:
        for name in [c for c in args.configs.split(",") if c and c != args.config]:
            try:
                configs[name] = side_config(name, args, rank, world, dev, local, flush, peak)
            except Exception as exc:        # report, never hide; keep the ranks in step
                configs[name] = {"error": f"{type(exc).__name__}: {exc}"}
                if world > 1:
                    raise

    if rank == 0:
        cfg = config_dict(args, p, B, world, max_frac)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": cfg,
            "workload_detail": {"max_trajectory_len": T, "valid_ids_per_step": valid_ids, "padded_ids_per_step": B * T,
                                "row_lengths_passed": use_len is not None, "id_bytes": 8},
            "row_solves_per_s": value * p.n,
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "kernels": kern,
            "e2e": e2e,
            "dp_training": dp_training,
            "cpu_baseline": cpu,
            "parity_check": parity,
            "extras": extras,
            "sampler": sampler_res,
            "configs": configs,
            "context": {"setup_s": setup_s, "plan_contributions": contributions,
                        "device_bytes": ctx_bytes, "max_row_union": max_union,
                        "ls_class_rows": class_rows},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    sys.setswitchinterval(2e-4)         # let the clock-sampling thread run between the timed launches
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
