#!/usr/bin/env python
"""bench.py -- G-Shell inverse-rendering hot path on B200 (driver contract: see task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--grid 256]

One "step" = one pass of the hot path over one batch of synthetic input at BASELINE.json's
configs[3] shape: "256" tet grid (BCC N=103: 2,217,591 verts / 12,985,416 tets), 8 views @ 1024^2,
n_samples=16.  The stages a step currently executes are listed in config.stages (the list grows as
rows of SURVEY.md section 8 land; a stage that is not listed is NOT in the timed region).

N>1: launched under torchrun, one rank per GPU; the 8 views of the batch shard across the ranks
(8/N each), extraction is replicated, one NCCL all-reduce over the flat (sdf|msdf|deform|light)
gradient bucket per step.  Total work is fixed ("strong" scaling), `value` = optimiser steps per second
of the whole job, `rendered_mpix_per_s` = pixels of all 8 views per second.
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
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=256, choices=sorted(GRID_N))
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--n-samples", type=int, default=16)
    ap.add_argument("--cpu-sample-grid", type=int, default=128,
                    help="grid the CPU baseline is timed on (scaled to --grid by tet count)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true", help="skip the secondary (non-headline) measurements")
    ap.add_argument("--sdf-init", default="random", choices=["random", "sphere"],
                    help="profiling aid: 'sphere' makes the reference's sphere_init field the main workload (the headline "
                         "metric is quoted on 'random', BASELINE.json)")
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


COVERAGE_FOR_CPU_SCALING = 0.5      # fraction of the pixels that carry geometry in the synthetic views (measured ~0.55 on the GPU arm)


def cpu_shading_seconds(sample_res, n_samples):
    """The reference's OWN env-light integrator (envsampling/kernel.cu, unmodified, compiled for the CPU by oracle/build_ref.py
    with OpenMP over pixels) forward + backward on one fully covered sample_res^2 view, no occluders.  Returns (seconds, pixels)
    or None when the library is not available."""
    try:
        import torch
        from oracle import build_ref
        if build_ref.build() is None:
            return None
        from oracle import ref_env_shade as ref
        from oracle import shade_oracle as so
        g = torch.Generator().manual_seed(0)
        B, H, W, n = 1, sample_res, sample_res, n_samples
        nrm = torch.nn.functional.normalize(torch.randn(B, H, W, 3, generator=g), dim=-1)
        nrm[..., 2] = nrm[..., 2].abs() + 0.1
        nrm = torch.nn.functional.normalize(nrm, dim=-1)
        pos = torch.rand(B, H, W, 3, generator=g) - 0.5
        view = torch.tensor([0.0, 0.0, 3.0]).view(1, 1, 1, 3)
        kd = torch.rand(B, H, W, 3, generator=g)
        ks = torch.stack([torch.zeros(B, H, W), 0.08 + 0.9 * torch.rand(B, H, W, generator=g), torch.rand(B, H, W, generator=g)], -1)
        light = torch.rand(256, 256, 3, generator=g) * 0.5 + 0.25
        pdf, rows, cols = so.light_pdf_tables(light)
        perms = torch.argsort(torch.rand(32768, n * n, generator=g), dim=-1).int()
        a = (torch.ones(B, H, W), pos, pos, nrm, view, kd, ks, light, pdf, rows, cols, perms)
        t0 = time.perf_counter()
        d, sp = ref.env_shade_fwd(*a, bsdf=0, n_samples_x=n, rnd_seed=1)
        ref.env_shade_bwd(*a, torch.ones_like(d), torch.ones_like(sp), bsdf=0, n_samples_x=n, rnd_seed=1)
        return time.perf_counter() - t0, B * H * W
    except Exception as e:                                     # the baseline must never take the bench line down
        print(f"[bench] cpu shading baseline unavailable: {e!r}", file=sys.stderr)
        return None


def cpu_step_seconds(args, full_tets, dt_extract, sample_tets):
    """Host-core time of one step of the reference: extraction (oracle port, scaled by tet count) + env_shade fwd+bwd (the
    reference's own kernel compiled for the CPU, scaled by covered pixels).  Shadow rays, rasterisation, denoiser and the
    G-buffer passes have no CPU implementation in the reference and are NOT included, which favours the baseline."""
    t_ext = dt_extract * full_tets / sample_tets
    sh = cpu_shading_seconds(256, args.n_samples)
    if sh is None:
        return t_ext, t_ext, None, "rendering stages have no CPU implementation"
    dt_sh, px = sh
    covered = COVERAGE_FOR_CPU_SCALING * args.views * args.res * args.res
    t_sh = dt_sh * covered / px
    note = (f"+ env_shade fwd+bwd by the reference's own envsampling/kernel.cu compiled for the CPU (oracle/_ref, OpenMP): {dt_sh:.2f} s "
            f"on {px} px, scaled to {int(covered)} covered px ({COVERAGE_FOR_CPU_SCALING:.0%} of {args.views}x{args.res}^2) = {t_sh:.1f} s; "
            f"no shadow rays / raster / denoiser on the CPU side")
    return t_ext + t_sh, t_ext, t_sh, note


def reference_shading_full(args, coverage=COVERAGE_FOR_CPU_SCALING):
    """The reference's own env-light integrator (oracle/_ref: envsampling/kernel.cu compiled unmodified for the CPU, OpenMP over
    pixels), forward + backward over EVERY view of the configured batch at the configured resolution and sample count; a centred
    disc covers `coverage` of each frame (the reference skips masked pixels, kernel.cu:478).  Returns (seconds, covered pixels)."""
    import torch
    from oracle import build_ref
    if build_ref.build() is None:
        return None
    from oracle import ref_env_shade as ref
    from oracle import shade_oracle as so
    g = torch.Generator().manual_seed(0)
    H = W = args.res
    n = args.n_samples
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    mask = ((xx * xx + yy * yy) < coverage * 4.0 / 3.14159265).float()[None]
    light = torch.rand(256, 256, 3, generator=g) * 0.5 + 0.25
    pdf, rows, cols = so.light_pdf_tables(light)
    perms = torch.argsort(torch.rand(32768, n * n, generator=g), dim=-1).int()
    view = torch.tensor([0.0, 0.0, 3.0]).view(1, 1, 1, 3)
    total, covered = 0.0, 0
    for _ in range(args.views):
        nrm = torch.nn.functional.normalize(torch.randn(1, H, W, 3, generator=g), dim=-1)
        nrm[..., 2] = nrm[..., 2].abs() + 0.1
        nrm = torch.nn.functional.normalize(nrm, dim=-1)
        pos = torch.rand(1, H, W, 3, generator=g) - 0.5
        kd = torch.rand(1, H, W, 3, generator=g)
        ks = torch.stack([torch.zeros(1, H, W), 0.08 + 0.9 * torch.rand(1, H, W, generator=g), torch.rand(1, H, W, generator=g)], -1)
        a = (mask, pos, pos, nrm, view, kd, ks, light, pdf, rows, cols, perms)
        t0 = time.perf_counter()
        d, sp = ref.env_shade_fwd(*a, bsdf=0, n_samples_x=n, rnd_seed=1)
        ref.env_shade_bwd(*a, torch.ones_like(d), torch.ones_like(sp), bsdf=0, n_samples_x=n, rnd_seed=1)
        total += time.perf_counter() - t0
        covered += int(mask.sum())
    return total, covered


def run_reference(args):
    """--impl reference: ONE step of the reference's own CPU-runnable code for this path at the FULL configuration, measured,
    not extrapolated: the PyTorch extraction algorithm (row-wise `unique`, gshell_tets.py:268; oracle port on torch CPU) forward
    + backward on the full grid, plus env_shade forward + backward through the reference's own integrator compiled for the CPU
    over all views.  Rasterisation, shadow rays (OptiX), the denoiser and the G-buffer passes have no CPU implementation in the
    reference and are absent -- which favours the reference.  One step takes minutes on the host cores, so exactly one step is
    timed whatever --steps says (BASELINE.md section 3.1 allows a single repetition at this scale); `steps` reports that."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    t_ext, full_tets, n_full = cpu_extraction_seconds(args.grid, cores)
    sh = None
    try:
        sh = reference_shading_full(args)
    except Exception as e:                                     # the arm must still print its line
        print(f"[bench] reference shading unavailable: {e!r}", file=sys.stderr)
    t_sh, covered = sh if sh is not None else (0.0, 0)
    per_step = t_ext + t_sh
    wall = time.perf_counter() - t0
    value = 1.0 / per_step
    kind = "reference" if sh is not None else "port"
    sample = (f"one full step, measured: extraction (reference algorithm on torch CPU, oracle/mt_oracle.py, fwd+bwd) on BCC N={n_full} "
              f"({full_tets} tets): {t_ext:.1f} s; "
              + (f"env_shade fwd+bwd by the reference's own envsampling/kernel.cu compiled for the CPU (oracle/_ref, OpenMP) on all "
                 f"{args.views} views @ {args.res}^2, n_samples={args.n_samples}, {covered} covered px: {t_sh:.1f} s; "
                 if sh is not None else "env_shade: oracle/_ref not available; ")
              + f"no raster / shadow rays / denoiser on the CPU side (the reference has no CPU code for them); wall {wall:.1f} s")
    line = {"impl": "reference", "metric": "train_iters_per_sec", "value": value, "unit": "iters/s",
            "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": per_step * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, n_full, full_tets, None),
            "parts_s": {"extraction": t_ext, "env_shade": t_sh},
            "cpu_baseline": {"value": value, "unit": "iters/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args, n, n_tets, stages):
    return {"workload": f"gshell_tets '{args.grid}' grid = BCC N={n} ({n_tets} tets), {args.views} views @ {args.res}^2 "
                        f"in total ({args.views // max(args.gpus, 1)} per GPU), n_samples={args.n_samples} ({2 * args.n_samples ** 2} BSDF evals/px), "
                        f"{'random SDF/mSDF' if args.sdf_init == 'random' else 'sphere_init SDF'}",
            "stages": stages, "l2": "inputs larger than L2 (tet tables 0.5 GB)",
            "parallelism": f"view-sharded dp{args.gpus}" + ("; forward shading + shadow rays dealt out over the ranks in 8-row blocks (2 all-to-alls), "
                                                            "1 gradient all-reduce" if args.gpus > 1 else "")}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
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
    from gshell_b200 import _lib   # fails loudly if the CUDA library is missing
    from gshell_b200 import synthetic, timing
    from gshell_b200.denoiser.denoiser import BilateralDenoiser
    from gshell_b200.distributed import allreduce_mean_grads_
    from gshell_b200.geometry.gshell_tets_geometry import GShellTetsGeometry, default_flags
    from gshell_b200.grids import save_tets_npz
    from gshell_b200.render import light
    from gshell_b200.render import render as _render
    from gshell_b200.render import renderutils as ru

    n = GRID_N[args.grid]
    # strong scaling: BASELINE.json's metric names ONE batch ("8 x 1024^2 views") on 1/2/4/8 GPUs, so the views of that batch
    # shard across the ranks and `value` = optimiser steps per second of the whole job
    from gshell_b200.distributed import shard_views
    if args.views % world != 0:
        raise SystemExit(f"--views {args.views} must be divisible by the number of GPUs ({world}): equal view counts per rank "
                         "keep the mean all-reduce of the gradients exact")
    B, res = len(shard_views(args.views, rank, world)), [args.res, args.res]
    npz = os.path.join(tempfile.gettempdir(), f"gsb_bcc_{n}_{rank}.npz")
    save_tets_npz(npz, n)
    loss_fn = lambda img, ref: ru.image_loss(img, ref, loss="l1", tonemapper="log_srgb")   # 'logl1', the reference default

    class Workload:
        """One optimisation problem: geometry parameters, material leaf, light, targets; step() = one training iteration."""

        def __init__(self, sphere_init, kind="tets", grid=None, B=B, res=res, n_samples=None):
            torch.manual_seed(0)
            self.FLAGS = default_flags(n_samples=args.n_samples if n_samples is None else n_samples, sphere_init=sphere_init)
            if kind == "tets":
                f = npz
                if grid is not None and grid != args.grid:
                    f = os.path.join(tempfile.gettempdir(), f"gsb_bcc_{GRID_N[grid]}_{rank}.npz")
                    save_tets_npz(f, GRID_N[grid])
                self.geometry = GShellTetsGeometry(args.grid if grid is None else grid, 2.0, self.FLAGS, tet_init_file=f, device=dev)
                if f != npz:
                    os.unlink(f)
            else:
                from gshell_b200.geometry.gshell_flexicubes_geometry import GShellFlexiCubesGeometry
                self.geometry = GShellFlexiCubesGeometry(grid, 2.0, self.FLAGS, device=dev)
            gen = torch.Generator().manual_seed(1000 + rank)         # each rank shades its own views
            rng = np.random.RandomState(1000 + rank)
            self.res = res
            self.mat = synthetic.LeafMaterialField(B, res[0], res[1], dev, gen)
            self.material = {"kd_ks": self.mat, "bsdf": "pbr"}
            self.lgt = light.create_trainable_env_rnd(256, scale=0.5, bias=0.25, device=dev)
            self.denoiser = BilateralDenoiser().to(dev)
            g = self.geometry
            geo_params = [g.sdf, g.msdf, g.deform] + ([g.per_cube_weights] if kind != "tets" else [])
            self.shared = geo_params + [self.lgt.base]               # replicated across ranks -> all-reduced
            self.optim = torch.optim.Adam([{"params": geo_params, "lr": 1e-3},
                                           {"params": [self.mat.tex], "lr": 1e-2}, {"params": [self.lgt.base], "lr": 1e-2}],
                                          fused=True)
            mvp, campos = synthetic.random_cameras(B, res, "cpu", rng)
            img, bg = synthetic.random_target(B, res, "cpu", gen)
            # per-step inputs: pinned host memory for the e2e leg, resident on the device for `value`
            self.host = {k: v.pin_memory() for k, v in dict(mvp=mvp, campos=campos, img=img, background=bg).items()}
            self.resident = {k: v.to(dev) for k, v in self.host.items()}
            self.staging = {k: torch.empty_like(v, device=dev) for k, v in self.host.items()}
            self.host_out = torch.zeros(1).pin_memory()
            self.it = 0
            # replicated parameters were initialised under the common seed; per-rank streams from here on
            # (MC sample seeds and the jitter noise must differ between ranks)
            torch.manual_seed(1 + rank)
            _render.rnd_seed = rank * 1000003

        def step(self, e2e=False, it_base=1000, fixed_it=None):
            if e2e:
                for k in self.host:
                    self.staging[k].copy_(self.host[k], non_blocking=True)
                t = self.staging
            else:
                t = self.resident
            target = {"mvp": t["mvp"], "campos": t["campos"], "img": t["img"], "background": t["background"],
                      "resolution": self.res, "spp": 1}
            with timing.stage("light_tables"):
                self.lgt.update_pdf()
            self.optim.zero_grad(set_to_none=True)
            # it_base = 1000: full shadow ramp and full-radius denoiser, the steady state of training
            with timing.stage("forward_total"):
                img_loss, depth_loss, reg_loss = self.geometry.tick(None, target, self.lgt, self.material, loss_fn,
                                                                    (it_base + self.it) if fixed_it is None else fixed_it, self.denoiser)
                total = img_loss + depth_loss + reg_loss
            with timing.stage("backward_total"):
                total.backward()
            with timing.stage("allreduce"):
                allreduce_mean_grads_(self.shared)            # one flat NCCL all-reduce (no-op at world size 1)
            with timing.stage("adam+clamp"):
                self.optim.step()
                with torch.no_grad():
                    self.geometry.clamp_deform()
                    self.lgt.clamp_(min=0.0)
            self.it += 1
            if e2e:
                self.host_out.copy_(total.detach().reshape(1), non_blocking=True)
                torch.cuda.current_stream().synchronize()
            return total

    def timed(wl, nsteps, e2e, it_base=1000, fixed_it=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(nsteps):
            wl.step(e2e, it_base, fixed_it)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    stages = ["light.update_pdf", "mt_extract", "occluder grid build", "vertex_normals", "xfm_points", "rasterize",
              "gbuffer (fused: 5 interpolations, face normals, depth)", "prepare_shading_normal",
              f"env_shade n={args.n_samples} incl. shadow rays (wavefront trace)", "bilateral_denoiser (fused pair, radius 11)",
              "compose (fused: regulariser taps, combine, 12 buffers, composite)", "antialias (every composited buffer)",
              "image_loss + mask/msdf/regulariser losses", "backward (all of the above)", "fused adam step"] + \
             (["nccl_allreduce_grads"] if world > 1 else [])
    wl = Workload(sphere_init=args.sdf_init == "sphere")
    n_tets = int(wl.geometry.indices.shape[0])
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        wl.step()
    launches0 = _lib.launch_count
    _lib.lib.gsb_trace_ray_count(1)
    _lib.lib.gsb_trace_timing(1)                      # CUDA events around the trace launches, on the launching stream
    ms_total = timed(wl, args.steps, e2e=False)
    trace_ms = float(_lib.lib.gsb_trace_timing(0))
    trace_rays = int(_lib.lib.gsb_trace_ray_count(1))
    trace_launches = int(_lib.lib.gsb_trace_launches())
    launches = _lib.launch_count - launches0
    occ = wl.geometry.optix_ctx
    occ_info = {"grid_res": getattr(occ, "grid_res", None), "entries": getattr(occ, "n_entries", None)}
    ms_e2e = timed(wl, args.steps, e2e=True)
    # per-rank stage breakdown (2 extra steps, outside every timed region): device ms per step of each stage, MAX over ranks,
    # plus each rank's own shading time (the views differ in covered pixels -> the slowest rank sets the step time)
    timing.start()
    for _ in range(2):
        wl.step()
    st = {k: v / 2 for k, v in timing.stop().items()}
    names = sorted(st)
    tvec = torch.tensor([st[k] for k in names], device=dev)
    per_rank_shade = torch.zeros(world, device=dev)
    per_rank_shade[rank] = st.get("env_shade_fwd(gen+trace+shade)", 0.0)
    if world > 1:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
        dist.all_reduce(per_rank_shade)
    stages_ms = {k: round(float(v), 3) for k, v in zip(names, tvec)}
    stages_ms["_env_shade_fwd_per_rank"] = [round(float(x), 1) for x in per_rank_shade]
    clocks = sampler.stop() if sampler else None
    with torch.no_grad():
        d = wl.geometry.getMesh(wl.material)
    mesh_info = {"Va": int(d["imesh"].v_pos.shape[0]), "Fa": int(d["imesh"].t_pos_idx.shape[0]),
                 "Vw": int(d["n_verts_watertight"])}

    # ---- secondary measurements (same code path, 3 steps each; NOT the headline) ------------------------------------------
    variants = {}
    if not args.no_variants:
        wl.step(False, 0, 0)
        ms0 = timed(wl, 3, e2e=False, fixed_it=0) / 3
        variants["random_sdf_iteration0"] = {"ms_per_step": ms0, "note": "shadow_scale=0 (no rays needed), denoiser radius 3: "
                                             "the first iteration of the reference's ramp (gshell_tets_geometry.py:264)"}
        del wl
        torch.cuda.empty_cache()
        ws = Workload(sphere_init=True)
        for _ in range(3):
            ws.step()
        mss = timed(ws, 3, e2e=False) / 3
        with torch.no_grad():
            ds = ws.geometry.getMesh(ws.material)
        variants["sphere_init"] = {"ms_per_step": mss, "faces": int(ds["imesh"].t_pos_idx.shape[0]),
                                   "note": "reference's sphere_init SDF (gshell_tets_geometry.py:112-113): closed surface, "
                                           "same grid / views / samples / shadows"}
        lgt_for_roof = ws.lgt
        # BASELINE.json configs[1] / configs[2] at their own shapes (world size 1 only: they name 4 views)
        if world == 1:
            for name, kind, grid, note in (("C2_polycam_mc_128", "tets", 128, "'128' tet grid (BCC N=52), 4 views @ 512^2, n_samples=8"),
                                           ("C3_deepfashion_mc_80", "flex", 80, "80^3 G-FlexiCubes grid, 4 views @ 512^2, n_samples=8")):
                try:
                    wv = Workload(sphere_init=True, kind=kind, grid=grid, B=4, res=[512, 512], n_samples=8)
                    for _ in range(3):
                        wv.step()
                    variants[name] = {"ms_per_step": timed(wv, 5, e2e=False) / 5, "note": note + ", sphere_init, full shadow ramp"}
                    with torch.no_grad():
                        dm = wv.geometry.getMesh(wv.material)
                    variants[name]["faces"] = int(dm["imesh"].t_pos_idx.shape[0])
                    del wv, dm
                    torch.cuda.empty_cache()
                except Exception as e:          # pragma: no cover
                    variants[name] = {"error": repr(e)[:300]}
    else:
        lgt_for_roof = wl.lgt
    os.unlink(npz)

    # ---- roofline ------------------------------------------------------------------------------------------------------
    # dominant kernel of the step = k_trace_pool (shadow rays), timed live in the timed region above.  Algorithmic bytes per
    # launch: 33 B per ray (32 B list entry in, 1 B visibility out) + the occluder tables read once (4 B/cell + 48 B/entry).
    peak, how = measured_peaks()
    n_chunks = min(1024, trace_launches)                              # the library times at most 1024 trace launches
    roof = {"bound": "hbm", "kernel": "k_trace_pool (any-hit shadow rays through the brick / cell / sub-voxel bit hierarchy; issue / ALU / L1 / latency "
            "bound, not HBM -- see profiles/r2l_trace_kernel_ncu.md -- HBM fraction reported as required)",
            "peak": peak, "peak_source": how, "unit": "GB/s",
            "traffic": None, "occluder": occ_info}
    if trace_ms > 0 and trace_rays > 0 and occ_info["grid_res"]:
        # occupancy bits (1 bit per cell) + 16-B cell records + 48-B triangle records, each read once
        tables = (16 + 0.125) * occ_info["grid_res"] ** 3 + 48 * occ_info["entries"]
        launches_tr = max(1, n_chunks)
        if trace_launches > 1024:                                     # rays counted over all launches, time over the first 1024
            trace_rays = int(trace_rays * 1024 / trace_launches)
        alg = int(33 * trace_rays + tables * launches_tr)
        ach = alg / (trace_ms * 1e-3) / 1e9
        roof.update(achieved=ach, frac=ach / peak, algorithmic_bytes_total=alg, ms_total=trace_ms, launches=launches_tr,
                    rays=trace_rays, rays_per_s=trace_rays / (trace_ms * 1e-3), share_of_step=trace_ms / ms_total)
        roof["traffic"], roof["traffic_note"] = committed_traffic(trace_rays / launches_tr)
        roof["algorithmic_bytes_per_launch"] = alg / launches_tr
    else:
        roof.update(achieved=None, frac=None)
    try:
        roof["other_kernels"] = {"env_shade": env_shade_roofline(args, dev, lgt_for_roof, peak, how, B, res)}
    except Exception as e:          # pragma: no cover
        roof["other_kernels"] = {"error": repr(e)}
    try:
        roof["other_kernels"]["mt_extract_fwd"] = mt_roofline(args, dev, peak, how)
    except Exception as e:          # pragma: no cover
        roof["other_kernels"]["mt_extract_fwd"] = {"error": repr(e)}
    if not args.no_variants:
        try:
            roof["other_kernels"]["flexicubes_extract_fwd"] = flex_roofline(dev, peak, how)
        except Exception as e:          # pragma: no cover
            roof["other_kernels"]["flexicubes_extract_fwd"] = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_step = ms_total / args.steps
    mpix = args.views * res[0] * res[1] / 1e6
    h2d = world * sum(v.numel() * v.element_size() for v in (wl.host if args.no_variants else ws.host).values())
    line = {"metric": "train_iters_per_sec", "value": 1e3 / ms_step, "unit": "iters/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, n, n_tets, stages),
            "rendered_mpix_per_s": mpix * 1e3 / ms_step, "mesh": mesh_info,
            "e2e": {"value": 1e3 / (ms_e2e / args.steps), "unit": "iters/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": launches, "stages_ms_max_over_ranks": stages_ms, "variants": variants,
            "clocks": clocks, "roofline": roof}
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        # the extraction does not scale linearly with the tet count on the host (measured: N=52 13-28 s, N=103 70 s), so the
        # sample IS the full grid, once (about a minute of CPU work); env_shade is sampled on one 256^2 view and scaled by pixels
        dt, sample_tets, n_s = cpu_extraction_seconds(args.grid if args.cpu_sample_grid == 128 else args.cpu_sample_grid, cores)
        scaled, t_ext, t_sh, note = cpu_step_seconds(args, n_tets, dt, sample_tets)
        line["cpu_baseline"] = {"value": 1.0 / scaled, "unit": "iters/s", "cores": cores, "kind": "port",
                                "parts_s": {"extraction_port": t_ext, "env_shade_reference_compiled": t_sh},
                                "sample": f"extraction: oracle/mt_oracle.py (reference algorithm, torch CPU) fwd+bwd once on BCC "
                                          f"N={n_s} ({sample_tets} tets): {dt:.2f} s" + (f", scaled x{n_tets / sample_tets:.2f} by tet count = "
                                          f"{t_ext:.1f} s" if sample_tets != n_tets else " (the full grid, not scaled)") + f"; {note}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def committed_traffic(rays_per_launch):
    """`roofline.traffic` cannot be measured inside a timed run (it needs ncu's replay): it is the DRAM bytes per ray of the
    committed `ncu --set full` capture of the same kernel (profiles/r2l_trace_traffic.json) times this run's rays per launch, in GB
    per launch like `achieved`'s numerator, and the note says so."""
    path = os.path.join(ROOT, "profiles", "r2l_trace_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
    except (OSError, ValueError):
        return None, "no committed capture"
    gb = t["dram_bytes_per_ray"] * rays_per_launch / 1e9
    return gb, (f"source: committed capture, not this run -- {t['dram_bytes_per_ray']:.0f} B of DRAM traffic per ray "
                f"(dram__bytes_read+write of {t['capture']}) x {rays_per_launch / 1e6:.0f} M rays per launch")


def mt_roofline(args, dev, peak, how):
    """HBM roofline of the extraction forward (gsb_mt_count + gsb_mt_emit = 7 kernels incl. the one host read of the counts).
    Algorithmic bytes: 16T (tet ids) + 20Nv (sdf, msdf, pos) + 16Va (verts_aug + msdf_aug) + 12Fa + 12Fw."""
    import torch
    from gshell_b200.geometry.gshell_tets import GShell_Tets
    pos, sdf, msdf, tets, n = synth_grid(args.grid)
    pos, sdf, msdf, tets = pos.to(dev), sdf.to(dev), msdf.to(dev), tets.to(dev)
    mt = GShell_Tets(index_dtype=torch.int32, with_tangents=False)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times = []
    for it in range(6):
        torch.cuda.synchronize()
        ev[0].record()
        with torch.no_grad():
            va, fa, _, _, _, ex = mt(pos, sdf, msdf, tets)
        ev[1].record()
        torch.cuda.synchronize()
        if it:
            times.append(ev[0].elapsed_time(ev[1]))
    ms = sorted(times)[len(times) // 2]
    T, nv = int(tets.shape[0]), int(pos.shape[0])
    alg = 16 * T + 20 * nv + 16 * int(va.shape[0]) + 12 * int(fa.shape[0]) + 12 * int(ex["faces_watertight"].shape[0])
    ach = alg / (ms * 1e-3) / 1e9
    out = {"bound": "hbm", "kernel": "marching-tets forward (7 kernels + 1 host read of the counts)", "achieved": ach, "peak": peak,
           "peak_source": how, "unit": "GB/s", "frac": ach / peak, "algorithmic_bytes": alg, "ms": ms,
           "active_tets_per_s": (int(ex["faces_watertight"].shape[0])) / (ms * 1e-3)}
    # the reference's PyTorch extraction as the train scripts execute it -- ON the B200 (same algorithm
    # incl. the row-wise unique of gshell_tets.py:268; the oracle port run under a CUDA default device), forward only
    try:
        from oracle.mt_oracle import gshell_marching_tets
        del va, fa, ex
        tl = tets.long()
        rt = []
        with torch.device(dev), torch.no_grad():
            for it in range(3):
                torch.cuda.synchronize()
                ev[0].record()
                gshell_marching_tets(pos, sdf, msdf, tl, unique_mode="rows", with_tangents=False)
                ev[1].record()
                torch.cuda.synchronize()
                if it:
                    rt.append(ev[0].elapsed_time(ev[1]))
        out["reference_torch_on_this_gpu_ms"] = min(rt)
        out["speedup_vs_reference_torch_on_this_gpu"] = min(rt) / ms
    except Exception as e:          # pragma: no cover
        out["reference_torch_on_this_gpu_ms"] = None
        out["reference_torch_note"] = repr(e)[:200]
    return out


def flex_roofline(dev, peak, how, res=80):
    """HBM roofline of the G-FlexiCubes extraction forward at BASELINE.json configs[2] (80^3).  Algorithmic bytes:
    32T (cube indices) + 84T (weights) + 20Nv (x, s, nu) + 16 Vopen + 12 Fopen + 4 |L_dev|."""
    import torch
    from gshell_b200.geometry.gshell_flexicubes import GShellFlexiCubes
    fc = GShellFlexiCubes(device=dev, index_dtype=torch.int32)
    verts, cubes = fc.construct_voxel_grid(res)
    g = torch.Generator().manual_seed(0)
    nv, nc = verts.shape[0], cubes.shape[0]
    x = verts + (0.2 / res * (torch.rand(nv, 3, generator=g) - 0.5)).to(dev)
    s = verts.norm(dim=1) - 0.35
    nu = verts[:, 1] + 0.15
    w = (torch.randn(nc, 21, generator=g) * 0.5).to(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times = []
    for it in range(6):
        torch.cuda.synchronize()
        ev[0].record()
        with torch.no_grad():
            vo, fa, L, ex = fc(x, s, nu, cubes, res, w[:, :12], w[:, 12:20], w[:, 20])
        ev[1].record()
        torch.cuda.synchronize()
        if it:
            times.append(ev[0].elapsed_time(ev[1]))
    ms = sorted(times)[len(times) // 2]
    alg = 32 * nc + 84 * nc + 20 * nv + 16 * int(vo.shape[0]) + 12 * int(fa.shape[0]) + 4 * int(L.numel())
    ach = alg / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "G-FlexiCubes forward at 80^3 (topology + float kernels, host reads of the counts included)",
            "achieved": ach, "peak": peak, "peak_source": how, "unit": "GB/s", "frac": ach / peak, "algorithmic_bytes": alg, "ms": ms,
            "cubes": nc, "faces": int(fa.shape[0])}


def env_shade_roofline(args, dev, lgt, peak, how, B, res):
    """Dominant kernel = k_env_shade<BWD>.  Algorithmic bytes per launch: mask 4 B/px + covered px x
    (60 B G-buffer + 24 B d/d(diff,spec) in, 48 B gradients out) + light/pdf/cdf tables + perms table + 786 KB light grad."""
    import torch
    from gshell_b200.render import optixutils as ou
    H, W = res
    g = torch.Generator().manual_seed(5)
    nrm = torch.nn.functional.normalize(torch.randn(B, H, W, 3, generator=g), dim=-1)
    nrm[..., 2] = nrm[..., 2].abs()
    pos = (torch.rand(B, H, W, 3, generator=g) - 0.5).to(dev).requires_grad_()
    nrm = nrm.to(dev).requires_grad_()
    view = torch.tensor([0.0, 0.0, 3.0], device=dev).view(1, 1, 1, 3).expand(B, 1, 1, 3)
    kd = torch.rand(B, H, W, 3, generator=g).to(dev).requires_grad_()
    ks = torch.stack([torch.zeros(B, H, W), 0.08 + 0.9 * torch.rand(B, H, W, generator=g), torch.rand(B, H, W, generator=g)], -1).to(dev).requires_grad_()
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    mask = ((xx * xx + yy * yy) < 0.45).float()[None].expand(B, H, W).contiguous().to(dev)
    covered = int(mask.sum())
    n = args.n_samples
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    base = lgt.base.detach().clone().requires_grad_()
    fwd_ms, bwd_ms = [], []
    for it in range(4):
        diff, spec = ou.optix_env_shade(None, mask, pos.detach(), pos, nrm, view, kd, ks, base, lgt._pdf, lgt.rows[:, 0], lgt.cols,
                                        BSDF="pbr", n_samples_x=n, rnd_seed=it, shadow_scale=1.0)
        gd, gs = torch.ones_like(diff), torch.ones_like(spec)
        torch.cuda.synchronize()
        ev[0].record()
        ou.optix_env_shade(None, mask, pos.detach(), pos.detach(), nrm.detach(), view, kd.detach(), ks.detach(), base.detach(),
                           lgt._pdf, lgt.rows[:, 0], lgt.cols, BSDF="pbr", n_samples_x=n, rnd_seed=it, shadow_scale=1.0)
        ev[1].record()
        torch.cuda.synchronize()
        ev[2].record()
        torch.autograd.backward([diff, spec], [gd, gs])
        ev[3].record()
        torch.cuda.synchronize()
        if it:
            fwd_ms.append(ev[0].elapsed_time(ev[1]))
            bwd_ms.append(ev[2].elapsed_time(ev[3]))
    fwd, bwd = sorted(fwd_ms)[len(fwd_ms) // 2], sorted(bwd_ms)[len(bwd_ms) // 2]
    tables = 256 * 256 * 4 * 5 + 256 * 4 + 32768 * n * n * 4
    alg_bwd = 4 * B * H * W + covered * (60 + 24 + 48) + tables + 256 * 256 * 3 * 4
    alg_fwd = 4 * B * H * W + covered * (60 + 24) + tables
    achieved = alg_bwd / (bwd * 1e-3) / 1e9
    evals = covered * 2 * n * n
    return {"bound": "hbm", "kernel": "k_env_shade<BWD> (MC integrator backward; FP32-ALU/SFU bound by construction, SURVEY 8d)",
            "achieved": achieved, "peak": peak, "peak_source": how, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
            "algorithmic_bytes": alg_bwd, "ms": bwd, "covered_px": covered,
            "bsdf_evals_per_s": evals / (bwd * 1e-3),
            "fwd": {"ms": fwd, "achieved": alg_fwd / (fwd * 1e-3) / 1e9, "bsdf_evals_per_s": evals / (fwd * 1e-3)},
            "note": "timed alone (includes the Python wrapper's allocations of the 5 gradient tensors)"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
