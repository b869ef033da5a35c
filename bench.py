#!/usr/bin/env python
"""bench.py -- G-Shell inverse-rendering hot path on B200 (driver contract: see task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid 256]

One "step" = one pass of the hot path over one batch of synthetic input at BASELINE.json's
configs[3] shape: "256" tet grid (BCC N=103: 2,217,591 verts / 12,985,416 tets), 8 views @ 1024^2,
n_samples=16.  The stages a step currently executes are listed in config.stages (the list grows as
rows of SURVEY.md section 8 land; a stage that is not listed is NOT in the timed region).

N>1: launched under torchrun, one rank per GPU; views shard across ranks, extraction is replicated,
one NCCL all-reduce over the flat (sdf|msdf|pos) gradient bucket per step ("weak": views per rank
fixed at 8).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID_N = {64: 26, 128: 52, 256: 103}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=256, choices=sorted(GRID_N))
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--n-samples", type=int, default=16)
    ap.add_argument("--cpu-sample-grid", type=int, default=128,
                    help="grid the CPU baseline is timed on (scaled to --grid by tet count)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def synth_grid(name):
    import numpy as np
    import torch
    from gshell_b200.grids import bcc_tet_grid
    n = GRID_N[name]
    v, t = bcc_tet_grid(n)
    g = torch.Generator().manual_seed(0)
    pos = torch.tensor(v) - 0.5
    nv = v.shape[0]
    # the reference's own random init (gshell_tets_geometry.py:110,139)
    sdf = torch.rand(nv, generator=g) - 0.1
    msdf = (torch.rand(nv, generator=g) - 0.01).clamp(-1, 1)
    return pos, sdf, msdf, torch.tensor(t), n


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); out["sm_max_mhz"] = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if "Active" in val and "Not" not in val:
                    reasons.add(name)
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------------
def cpu_extraction_seconds(sample_grid, threads):
    """Oracle (port of the reference's PyTorch extraction, row-wise unique as at gshell_tets.py:268)
    forward+backward on the host cores, on a bounded sample grid."""
    import torch
    from oracle.mt_oracle import gshell_marching_tets
    torch.set_num_threads(threads)
    pos, sdf, msdf, tets, n = synth_grid(sample_grid)
    leaves = [x.clone().requires_grad_() for x in (pos, sdf, msdf)]
    t0 = time.perf_counter()
    va, fa, _, _, _, ex = gshell_marching_tets(*leaves, tets, unique_mode="rows", with_tangents=False)
    (va.sum() + ex["msdf"].sum()).backward()
    dt = time.perf_counter() - t0
    return dt, int(tets.shape[0]), n


def run_reference(args):
    """--impl reference: the reference's CPU path for the same stages (oracle port), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    _, _, _, tets_full, n_full = synth_grid(args.grid)
    full_tets = int(tets_full.shape[0])
    times = []
    for i in range(args.warmup + args.steps):
        dt, sample_tets, n_s = cpu_extraction_seconds(args.cpu_sample_grid, cores)
        if i >= args.warmup:
            times.append(dt)
        if sum(times) > 120:      # keep the arm within a few minutes
            break
    per_step = sorted(times)[len(times) // 2] * full_tets / sample_tets
    value = 1.0 / per_step
    sample = (f"oracle/mt_oracle.py fwd+bwd on BCC N={n_s} ({sample_tets} tets), median of {len(times)}, "
              f"scaled x{full_tets / sample_tets:.2f} by tet count to N={n_full}; rendering stages have no CPU implementation")
    line = {"impl": "reference", "metric": "train_iters_per_sec", "value": value, "unit": "iters/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, n_full, full_tets, None),
            "cpu_baseline": {"value": value, "unit": "iters/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, n, n_tets, stages):
    return {"workload": f"gshell_tets '{args.grid}' grid = BCC N={n} ({n_tets} tets), {args.views} views @ {args.res}^2 "
                        f"per GPU, n_samples={args.n_samples} ({2 * args.n_samples ** 2} BSDF evals/px)",
            "stages": stages, "l2": "inputs larger than L2 (tet tables 0.5 GB)",
            "parallelism": f"view-sharded dp{args.gpus}"}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from gshell_b200 import _lib  # noqa: F401  (fails loudly if the CUDA library is missing)
    from gshell_b200.geometry.gshell_tets import GShell_Tets

    pos_h, sdf_h, msdf_h, tets_h, n = synth_grid(args.grid)
    n_tets, nv = int(tets_h.shape[0]), int(pos_h.shape[0])
    tets = tets_h.to(dev)
    pos = pos_h.to(dev).requires_grad_()
    sdf = sdf_h.to(dev).requires_grad_()
    msdf = msdf_h.to(dev).requires_grad_()
    mt = GShell_Tets(index_dtype=torch.int32)
    # pinned host staging for the e2e leg (per-step inputs of this stage set: the field values)
    host_in = [x.clone().pin_memory() for x in (pos_h, sdf_h, msdf_h)]
    host_out = torch.zeros(1).pin_memory()
    stages = ["mt_extract_fwd", "mt_extract_bwd"] + (["nccl_allreduce_grads"] if world > 1 else [])
    info = {}

    def step(e2e=False):
        if e2e:
            for d, h in zip((pos, sdf, msdf), host_in):
                d.data.copy_(h, non_blocking=True)
        for p in (pos, sdf, msdf):
            p.grad = None
        va, fa, _, _, _, ex = mt(pos, sdf, msdf, tets)
        loss = va.sum() + ex["msdf"].sum()
        loss.backward()
        if world > 1:
            flat = torch.cat([pos.grad.reshape(-1), sdf.grad, msdf.grad])
            dist.all_reduce(flat)
        info.update(Vw=ex["n_verts_watertight"], Va=int(va.shape[0]), Fa=int(fa.shape[0]),
                    Fw=int(ex["faces_watertight"].shape[0]))
        if e2e:
            host_out.copy_(loss.detach().reshape(1), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return loss

    def timed(nsteps, e2e):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(nsteps):
            step(e2e)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 3)):
        step()
    sampler = ClockSampler(local) if rank == 0 else None
    ms_total = timed(args.steps, e2e=False)
    ms_e2e = timed(args.steps, e2e=True)
    clocks = sampler.stop() if sampler else None

    # roofline of the dominant kernel group: extraction forward (7 kernels), timed with CUDA events on
    # the launching stream; algorithmic bytes per SURVEY.md 8(d):
    #   16T (tet ids) + 20Nv (sdf,msdf,pos) + 16Va (verts_aug+msdf_aug) + 12Fa + 12Fw
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    fwd_ms = []
    for _ in range(5):
        torch.cuda.synchronize()
        ev[0].record()
        with torch.no_grad():
            mt(pos, sdf, msdf, tets)
        ev[1].record()
        torch.cuda.synchronize()
        fwd_ms.append(ev[0].elapsed_time(ev[1]))
    fwd = sorted(fwd_ms)[len(fwd_ms) // 2]
    alg_bytes = 16 * n_tets + 20 * nv + 16 * info["Va"] + 12 * info["Fa"] + 12 * info["Fw"]
    peak, how = measured_peaks()
    achieved = alg_bytes / (fwd * 1e-3) / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = ms_total / args.steps
    line = {"metric": "train_iters_per_sec", "value": 1e3 / ms_step, "unit": "iters/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, n, n_tets, stages),
            "mesh": info,
            "e2e": {"value": 1e3 / (ms_e2e / args.steps), "unit": "iters/s",
                    "h2d_bytes_per_step": sum(h.numel() * 4 for h in host_in), "d2h_bytes_per_step": 4},
            "gpu_launches": 9 * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "mt_extract forward (gsb_mt_count+gsb_mt_emit, 7 kernels, incl. 1 host sync)",
                         "achieved": achieved, "peak": peak, "peak_source": how, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "algorithmic_bytes": alg_bytes, "ms": fwd}}
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        dt, sample_tets, n_s = cpu_extraction_seconds(args.cpu_sample_grid, cores)
        scaled = dt * n_tets / sample_tets
        line["cpu_baseline"] = {"value": 1.0 / scaled, "unit": "iters/s", "cores": cores, "kind": "port",
                                "sample": f"oracle/mt_oracle.py (reference algorithm, torch CPU) fwd+bwd once on BCC N={n_s} "
                                          f"({sample_tets} tets): {dt:.2f} s, scaled x{n_tets / sample_tets:.2f} by tet count"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
