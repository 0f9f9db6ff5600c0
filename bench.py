#!/usr/bin/env python
"""Benchmark of the MS-UNet hot path (BASELINE.json): train img/s, fwd + DynamicLoss + bwd, T96 @ 512x512.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU), 16 images / GPU
    python bench.py --gpus N --global-batch 128               # SURVEY config 3: equal GLOBAL batch (128 / N per GPU, micro-batches)
    python bench.py --img 1024 --batch 4                      # the 1024x1024 rows of the north star
    python bench.py --impl reference --steps K --warmup W     # the reference's own modules on the host CPU cores
    python bench.py --impl reference-gpu                      # the reference's own modules, PyTorch eager fp16 autocast, on cuda:0

One JSON line on stdout (rank 0).  Keys follow the driver contract: `value` = whole-job img/s with inputs
resident in HBM (device-timed, max over ranks); `e2e` = the same through the public API with pinned-host
inputs copied H2D and the loss read back D2H every step; `roofline` = the dominant kernel (tcgen05 implicit
3x3 conv) timed live with CUDA events; `cpu_baseline` = the oracle port on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train_img_per_s_512x512_msunet_fwd_bwd"
UNIT = "img/s"


def metric_name(S):
    return METRIC if S == 512 else f"train_img_per_s_{S}x{S}_msunet_fwd_bwd"

IMG, PER_GPU_BATCH = 512, 16
T96 = dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24])
GFLOP_PER_IMG_FWD_BWD = 594.2  # BASELINE.md, T96 @ 512^2, live graph
GFLOP_TABLE = {("T96", 224): 109.8, ("T96", 512): 594.2, ("T96", 1024): 2345.5,
               ("B128", 224): 284.5, ("B128", 512): 1554.3, ("B128", 1024): 6162.8}


def synth_batch(B, S, seed):
    """Synthetic face-shaped batch: low-frequency images in [0,1], 60 % of masks carry 1-3 elliptical blobs,
    40 % are all-zero 'real' images."""
    import torch
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(B, 3, S // 16, S // 16, generator=g)
    x = torch.nn.functional.interpolate(base, size=(S, S), mode="bilinear", align_corners=False).clamp_(0, 1)
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    face = (((yy - S / 2) / (0.45 * S)) ** 2 + ((xx - S / 2) / (0.35 * S)) ** 2 < 1).float()
    x = (0.6 * x + 0.4 * face[None, None]).contiguous()
    y = torch.zeros(B, S, S)
    for b in range(B):
        if torch.rand((), generator=g) < 0.4:
            continue
        for _ in range(int(torch.randint(1, 4, (), generator=g))):
            cy, cx = (torch.rand(2, generator=g) * S).tolist()
            ry, rx = (torch.rand(2, generator=g) * 0.12 * S + 0.03 * S).tolist()
            y[b] = torch.maximum(y[b], (((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 < 1).float())
    return x, y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:                                # let the nvidia-smi process leave the driver before anything else is timed: its exit
            self.proc.wait(timeout=5)       # was seen to stall the launching thread for 40-50 ms (a 10 % dent in a 10-step e2e loop)
        except Exception:
            self.proc.kill()
        time.sleep(0.2)
        sm, mx, reasons = [], None, set()
        for l in self.lines:
            f = [t.strip() for t in l.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_reference_modules():
    """(MSUNetSys, DynamicLoss) of the UNMODIFIED reference from baseline/_ref (tools/install_reference.py copies it there; it
    travels to the GPU box), with the timm.layers stand-in of oracle/_shims; None when the copy is absent."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(ref, "network", "model_parts.py")):
        return None
    for p in (os.path.join(ROOT, "oracle", "_shims"), ref):
        if p not in sys.path:
            sys.path.append(p)
    try:
        from network.model_parts import MSUNetSys as RefNet      # noqa: the reference's own module
        from loss.DynamicLoss import DynamicLoss as RefLoss
    except Exception as e:  # pragma: no cover
        print(f"[bench] reference modules not importable ({type(e).__name__}: {e}); using the oracle port", file=sys.stderr)
        return None
    return RefNet, RefLoss


def cpu_reference_step_factory(img, batch, device="cpu", autocast=None):
    """One fwd + DynamicLoss + bwd of the reference on `device`: the reference's own modules (kind "reference") when
    baseline/_ref is present, else the oracle port (kind "port").  Returns (step, threads, kind)."""
    import contextlib
    import torch
    from oracle import msunet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.Cfg(img_size=img, embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24))
    sd = O.make_weights(cfg)
    x, y = synth_batch(batch, img, 4321)
    x, y = x.to(device), y.to(device)
    ctx = (lambda: torch.amp.autocast("cuda", dtype=autocast)) if autocast is not None else contextlib.nullcontext
    mods = load_reference_modules()
    if mods is not None:
        RefNet, RefLoss = mods
        with contextlib.redirect_stdout(sys.stderr):
            m = RefNet(img_size=img, embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=7,
                       drop_path_rate=0.1, attn_drop_rate=0.0, drop_rate=0.0)
        m.load_state_dict(sd, strict=True)
        m.to(device).train()
        crit = RefLoss(alpha=0.2, beta=0.8, tversky_bce_mix=0.45)
        scaler = torch.amp.GradScaler("cuda") if autocast == torch.float16 else None

        def step():
            for p in m.parameters():
                p.grad = None
            with ctx():
                loss = crit(m(x), y)
            (scaler.scale(loss) if scaler is not None else loss).backward()
            return float(loss)
        return step, torch.get_num_threads(), "reference"
    if device != "cpu":
        sd = {k: v.to(device) for k, v in sd.items()}

    def step():
        with ctx():
            _, loss, _ = O.train_step(sd, x, y, cfg)
        return float(loss)
    return step, torch.get_num_threads(), "port"


def run_reference(args, emit):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    S = args.img
    if args.impl == "reference-gpu":
        return run_reference_gpu(args, emit)
    sample_b = 1
    step, cores, kind = cpu_reference_step_factory(S, sample_b)
    for _ in range(max(1, min(args.warmup, 1))):
        step()
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    v = sample_b / dt
    sample = (f"T96 {S}x{S} fwd+DynamicLoss+bwd on {sample_b} image/step, {steps} timed steps (bounded sample of batch "
              f"{PER_GPU_BATCH}); {'the reference modules from baseline/_ref' if kind == 'reference' else 'oracle port'}, fp32")
    emit(({
        "impl": "reference", "metric": metric_name(S), "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"MS-UNet T96 training step {S}x{S} (reference, CPU)", "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_reference_gpu(args, emit):
    """BASELINE.md plan item 5: the reference's own modules in PyTorch eager on ONE B200, fp16 autocast + GradScaler exactly as
    trainer.py:182, 308-314 runs them (cuDNN / cuBLAS kernels; none of this repo's kernels).  The eager graph keeps every
    intermediate (~1.5 GiB / image under autocast at 512x512), so the batch is the largest of 16 / 8 / 4 / 2
    that fits."""
    import torch
    S = args.img
    dev = "cuda:0"
    torch.backends.cuda.matmul.allow_tf32 = True          # train.py:20-21
    torch.backends.cudnn.allow_tf32 = True
    last = None
    for B in (args.batch, 8, 4, 2, 1):
        if B > args.batch:
            continue
        try:
            step, _, kind = cpu_reference_step_factory(S, B, device=dev, autocast=torch.float16)
            for _ in range(max(2, min(args.warmup, 3))):
                step()
            torch.cuda.synchronize()
            steps = max(1, min(args.steps, 10))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            v = B / (ms / 1e3)
            sample = (f"T96 {S}x{S} fwd+DynamicLoss+bwd, batch {B}, PyTorch eager fp16 autocast + GradScaler on one B200 "
                      f"({'reference modules from baseline/_ref' if kind == 'reference' else 'oracle port'}), {steps} timed steps; "
                      f"peak memory {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB")
            emit({"impl": "reference-gpu", "metric": metric_name(S), "value": v, "unit": UNIT, "n_gpus": 1, "steps": steps,
                  "warmup": 2, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                  "dtype": "f16", "data": "synthetic",
                  "config": {"workload": f"MS-UNet T96 training step {S}x{S}, batch {B} (reference, GPU eager)", "sample": sample},
                  "gpu_eager_baseline": {"value": v, "unit": UNIT, "kind": kind, "batch": B, "sample": sample}})
            return
        except torch.OutOfMemoryError as e:
            last = e
            torch.cuda.empty_cache()
    emit({"impl": "reference-gpu", "unavailable": f"out of memory at every batch size: {last}"})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=PER_GPU_BATCH)
    ap.add_argument("--img", type=int, default=IMG)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default="T96", choices=["T96", "B128"], help="B128 = config.yaml default (embed 128, depths 2-2-18-2)")
    ap.add_argument("--drop-path", type=float, default=0.1)
    ap.add_argument("--attn-drop", type=float, default=0.0, help="attention dropout (config.yaml of the reference: 0.05)")
    ap.add_argument("--global-batch", type=int, default=0, help="equal GLOBAL batch (SURVEY config 3): each rank takes global/N images "
                    "per step in micro-batches of --micro-batch, gradients accumulated, one exchange per step; scaling = strong")
    ap.add_argument("--micro-batch", type=int, default=16)
    ap.add_argument("--optimizer", default="none", choices=["none", "fused", "sharded"],
                    help="include the AdamW step in the timed step: fused = all-reduce + replicated one-launch FusedAdamW; "
                         "sharded = reduce-scatter -> AdamW on the shard -> all-gather, overlapped with backward (N > 1)")
    ap.add_argument("--no-dp-check", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: everything else written to file descriptor 1 during the run (NCCL's version banner,
    # library chatter) is sent to stderr, and the result line is written to the saved descriptor at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())
    if args.impl in ("reference", "reference-gpu"):
        return run_reference(args, emit)

    import torch
    import torch.distributed as dist
    import semantic_segmentation_of_stylegan2_artifacts_b200 as pkg
    from semantic_segmentation_of_stylegan2_artifacts_b200 import ops
    from semantic_segmentation_of_stylegan2_artifacts_b200.loss.DynamicLoss import DynamicLoss
    from semantic_segmentation_of_stylegan2_artifacts_b200.network.model_parts import MSUNetSys

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        # stdout carries exactly one JSON line: NCCL's own log lines (a version banner when NCCL_DEBUG is set) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    W = max(3, args.warmup)
    B, S = args.batch, args.img
    strong = args.global_batch > 0
    if strong:
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
        B = args.global_batch // world                      # images per rank and step
    MB = min(args.micro_batch, B) if strong else B          # images per forward / backward
    if B % MB:
        raise SystemExit(f"per-rank batch {B} is not a multiple of the micro-batch {MB}")
    n_micro = B // MB

    torch.manual_seed(1234)
    arch = T96 if args.model == "T96" else dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32])
    gflop_img = GFLOP_TABLE.get((args.model, S), GFLOP_PER_IMG_FWD_BWD * (S / 512.0) ** 2)
    model = MSUNetSys(img_size=S, drop_path_rate=args.drop_path, attn_drop_rate=args.attn_drop, **arch).set_precision(args.precision).to(dev).train()
    if world > 1:
        from semantic_segmentation_of_stylegan2_artifacts_b200.dp import DataParallelB200
        model = DataParallelB200(model)
    crit = DynamicLoss(alpha=0.2, beta=0.8, tversky_bce_mix=0.45)
    x_h, y_h = synth_batch(B, S, 4321 + rank)
    x_h, y_h = x_h.pin_memory(), y_h.pin_memory()
    x_d, y_d = x_h.to(dev), y_h.to(dev)
    params = [p for p in model.parameters()]
    import contextlib
    opt = None                       # set after the first warm-up step (the sharded optimizer needs the learnt bucket layout)

    fused = None                     # replicated one-launch AdamW after the exchange (--optimizer fused)

    def step_core(x, y):             # the capturable part of a step
        for p in params:
            p.grad = None
        for i in range(n_micro):     # n_micro == 1 unless --global-batch: micro-batches accumulate, the last one exchanges
            last = i == n_micro - 1
            with (contextlib.nullcontext() if (last or world == 1) else model.no_sync()):
                loss = crit(model(x[i * MB:(i + 1) * MB]), y[i * MB:(i + 1) * MB])
                (loss if n_micro == 1 else loss / n_micro).backward()
        if opt is not None:
            opt.step()               # sharded: joins the per-bucket reduce-scatter / update / all-gather launched during backward
        elif world > 1:
            model.finish_gradient_sync()
        return loss

    def step(x, y):
        loss = step_core(x, y)
        if fused is not None:
            fused.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- warm-up (also builds weight shadows, workspaces, func attributes)
    step(x_d, y_d)
    if args.optimizer != "none":
        named = [(k, p) for k, p in model.named_parameters() if p.requires_grad]
        nd = [p for k, p in named if p.ndim == 1 or k.endswith(".bias") or "norm" in k.lower()]      # trainer.py:133-140
        dc = [p for k, p in named if not (p.ndim == 1 or k.endswith(".bias") or "norm" in k.lower())]
        groups = [{"params": dc, "weight_decay": 1e-3}, {"params": nd, "weight_decay": 0.0}]
        if args.optimizer == "sharded":
            if world == 1:
                raise SystemExit("--optimizer sharded needs N > 1")
            from semantic_segmentation_of_stylegan2_artifacts_b200.dp import ShardedAdamW
            opt = ShardedAdamW(model, groups, lr=1e-5, betas=(0.9, 0.999), eps=1e-8)
        else:
            from semantic_segmentation_of_stylegan2_artifacts_b200.optim import FusedAdamW
            fused = FusedAdamW(groups, lr=1e-5, betas=(0.9, 0.999), eps=1e-8)
    for _ in range(W - 1):
        step(x_d, y_d)
    barrier()

    # ---------------- optional whole-step CUDA graph (fwd + loss + bwd [+ allreduce])
    graph, g_loss = None, None
    use_graph = not args.no_graph
    if use_graph:
        try:
            if args.optimizer == "sharded":
                opt.prepare_step()       # per-step coefficients are uploaded from the host OUTSIDE the graph, before each replay
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                g_loss = step_core(x_d, y_d)
            graph.replay()
            if args.optimizer == "sharded":
                opt.after_replay()
            torch.cuda.synchronize()
        except Exception as e:  # capture is an optimisation, never a requirement
            print(f"[bench] CUDA graph capture unavailable ({type(e).__name__}: {e}); running eagerly", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    def run_step():
        if graph is not None:
            if args.optimizer == "sharded":
                opt.prepare_step()
            graph.replay()
            if args.optimizer == "sharded":
                opt.after_replay()
            if fused is not None:
                fused.step()             # eager, after the replayed fwd + bwd (+ exchange), as graphs.GraphedStep users do
            return g_loss
        return step(x_d, y_d)

    for _ in range(2):
        run_step()
    barrier()

    # ---------------- timed region: K steps, device-timed, max over ranks
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = pkg.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        run_step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = pkg.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)
    if graph is not None:
        # kernels replayed from the graph are the ones captured once: count them per replay
        n1 = pkg.launch_count()
        step(x_d, y_d)
        torch.cuda.synchronize()
        launches = (pkg.launch_count() - n1) * args.steps

    # ---------------- e2e: public API, pinned host -> device every step, loss read back every step.
    # The batches come through the repo's input stager (data.CudaPrefetcher): batch i+1 is copied from pinned host memory
    # on a copy stream while step i runs, exactly one H2D copy of images + masks and one D2H loss read per step.
    from semantic_segmentation_of_stylegan2_artifacts_b200.data import CudaPrefetcher
    # The loss of every step is read on the host (4 B, pinned ring buffer): the copy of step i is issued behind the step and waited
    # for after step i + E2E_DEPTH has been launched, so that (1) the CPU-side launch of a step (a ~0.7 ms graph launch with ~830
    # kernel nodes) overlaps the previous step instead of sitting between two steps and (2) the device has E2E_DEPTH steps queued when
    # the host hiccups (a 20-50 ms scheduling stall in a 10-step, 230 ms loop was a 10-20 % outlier with a queue of one).  Every loss
    # still reaches the host inside the timed region, the last ones before the clock stops.
    E2E_DEPTH = 3
    loss_host = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(E2E_DEPTH + 1)]
    loss_ev = [torch.cuda.Event() for _ in range(E2E_DEPTH + 1)]
    losses_read = []
    e2e_trace = os.environ.get("MSU_BENCH_E2E_TRACE") == "1"
    gpu_marks = []

    def e2e_loop(loader_):
        pending = []                     # ring slots of launched steps whose loss has not been read yet, oldest first
        t_prev = time.perf_counter()
        for i, batch in enumerate(loader_):
            if e2e_trace:
                t_now = time.perf_counter()
                sys.stderr.write(f"[e2e] step {i}: host {1e3 * (t_now - t_prev):.2f} ms since the previous\n")
                t_prev = t_now
            xd, yd = batch["image"], batch["label"]
            if e2e_trace:
                ga = torch.cuda.Event(enable_timing=True); ga.record()
            if graph is not None:
                x_d.copy_(xd, non_blocking=True)
                y_d.copy_(yd, non_blocking=True)
                lt = run_step()
            else:
                lt = step(xd, yd)
            if e2e_trace:
                gb = torch.cuda.Event(enable_timing=True); gb.record()
                gpu_marks.append((ga, gb))
            k = i % (E2E_DEPTH + 1)
            loss_host[k].copy_(lt.detach().reshape(1).float(), non_blocking=True)
            loss_ev[k].record()
            pending.append(k)
            if len(pending) > E2E_DEPTH:
                j = pending.pop(0)
                loss_ev[j].synchronize()
                losses_read.append(float(loss_host[j]))
        for j in pending:
            loss_ev[j].synchronize()
            losses_read.append(float(loss_host[j]))
        if e2e_trace and gpu_marks:
            torch.cuda.synchronize()
            sys.stderr.write("[e2e-gpu] step ms: " + " ".join(f"{a.elapsed_time(b):.2f}" for a, b in gpu_marks) + "\n")
            sys.stderr.write("[e2e-gpu] gap ms:  " + " ".join(f"{gpu_marks[i][1].elapsed_time(gpu_marks[i + 1][0]):.2f}" for i in range(len(gpu_marks) - 1)) + "\n")
            gpu_marks.clear()

    # untimed warm-up of the loop itself (max(3, W) steps): the stager's device buffers come out of the caching allocator's
    # pool afterwards instead of cudaMalloc (a device-wide stall of ~1 ms each, up to 25 % of a 10-step run on some boxes)
    # (ONE stager for warm-up and timed loop, as a training run has one: a second instance would bring a new copy stream and new
    # device buffers, i.e. six cudaMalloc calls of 16-50 MB inside the timed region, 20-80 ms on some boxes)
    loader = CudaPrefetcher([{"image": x_h, "label": y_h} for _ in range(max(3, args.warmup))], dev)
    e2e_loop(loader)
    losses_read.clear()
    loader.loader = [{"image": x_h, "label": y_h} for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import gc
    gc.collect()
    gc.disable()              # a generation-2 collection inside a 10-step host-paced loop is a 100 ms outlier
    barrier()
    e0.record()
    e2e_loop(loader)
    e1.record()
    barrier()
    gc.enable()
    assert len(losses_read) == args.steps and all(v == v for v in losses_read), "every step's loss must reach the host"
    te = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(te.item()) / 1e3)
    h2d = x_h.numel() * 4 + y_h.numel() * 4
    d2h = 4

    # the same loop fed with the RAW uint8 arrays of dataset/dataset.py:41-42; `/255`, CHW, flip and `>127` run on the device
    # (ops.stage_u8): 4 B per pixel cross PCIe instead of 16.  Reported beside `e2e`, not instead of it.
    xu_h = (x_h.permute(0, 2, 3, 1) * 255.0).round().clamp(0, 255).to(torch.uint8).contiguous().pin_memory()
    yu_h = (y_h * 255.0).to(torch.uint8).contiguous().pin_memory()
    fl_h = (torch.arange(B) % 2).to(torch.uint8).pin_memory()
    loader = CudaPrefetcher([{"image": xu_h, "label": yu_h, "flip": fl_h} for _ in range(max(3, args.warmup))], dev, stage_uint8=True)
    e2e_loop(loader)
    losses_read.clear()       # untimed warm-up, as above (also the first launch of the staging kernel)
    loader.loader = [{"image": xu_h, "label": yu_h, "flip": fl_h} for _ in range(args.steps)]
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    gc.collect()
    gc.disable()
    barrier()
    u0.record()
    e2e_loop(loader)
    u1.record()
    gc.enable()
    barrier()
    tu = torch.tensor([u0.elapsed_time(u1)], device=dev)
    if world > 1:
        dist.all_reduce(tu, op=dist.ReduceOp.MAX)
    e2e_u8 = {"value": world * B * args.steps / (float(tu.item()) / 1e3), "unit": UNIT,
              "h2d_bytes_per_step": xu_h.numel() + yu_h.numel() + fl_h.numel(), "d2h_bytes_per_step": 4}

    # ---------------- roofline of the dominant kernel, timed live with CUDA events on the launch stream
    roof = None
    ops.PROF = [] if rank == 0 else None
    step(x_d, y_d)          # every rank runs it (the step contains collectives); only rank 0 records events
    barrier()
    if rank == 0:
        agg, fam = {}, {}
        for tag, a, b, _nb, _fl in ops.PROF:
            t_ms = a.elapsed_time(b)
            f = fam.setdefault(tag[3], [0.0, 0, 0, 0])       # per kernel family: ms, launches, algorithmic bytes, flops
            f[0] += t_ms; f[1] += 1; f[2] += _nb; f[3] += _fl
            if not (tag[3].endswith("_tc") or tag[3].endswith("_simt")):
                continue          # GEMM-shaped launches only for the dominant kernel; tools/op_table.py prints the full table
            d = agg.setdefault(tag, [0.0, 0])
            d[0] += t_ms
            d[1] += 1
        ops.PROF = None
        fam_total = sum(v[0] for v in fam.values())
        top = max(agg.items(), key=lambda kv: kv[1][0])
        (M, N, K, kind), (tot_ms, cnt) = top
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        ach = 2.0 * M * N * K / (tot_ms / cnt * 1e-3) / 1e12
        traffic = None         # DRAM bytes of ONE ncu --set full capture: only reported for the kernel + shape it was taken on
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
            if tj.get("kind") == kind and [tj.get("M"), tj.get("N"), tj.get("K")] == [M, N, K]:
                traffic = tj.get("bytes_per_launch")
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6451.0)) if peaks else 6451.0
        roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": traffic, "kernel": f"{kind} M={M} N={N} K={K}", "launches_per_step": cnt,
                "share_of_step": tot_ms / ms_per_step if graph is None else None,
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s",
                "gemm_ms_by_kind": {f"{k[3]}:{k[0]}x{k[1]}x{k[2]}": round(v[0], 3) for k, v in
                                    sorted(agg.items(), key=lambda kv: -kv[1][0])[:8]},
                # every kernel family of one eager step, bracketed by CUDA events on the launch stream: share of the summed op
                # time, achieved TFLOP/s (contractions) and GB/s of algorithmic bytes (memory-bound families) against the peaks
                "families": {k: {"ms": round(v[0], 3), "share": round(v[0] / fam_total, 4), "launches": v[1],
                                 **({"tflops": round(v[3] / v[0] / 1e9, 1), "frac_tensor": round(v[3] / v[0] / 1e9 / peak, 3)} if v[3] else {}),
                                 **({"gbps": round(v[2] / v[0] / 1e6, 0), "frac_hbm": round(v[2] / v[0] / 1e6 / hbm_peak, 3)} if v[2] else {})}
                             for k, v in sorted(fam.items(), key=lambda kv: -kv[1][0])},
                "families_note": "eager pass, per-launch CUDA events (no side-stream overlap); shares are of the summed op time"}

    # ---------------- numerical check of the multi-GPU path (untimed): the exchanged gradients of named tensors and the loss
    # against a single-process pass over the GLOBAL batch (every rank's shard replayed on this rank with the same
    # stochastic-depth noise, gradients accumulated under no_sync)
    dp_check = None
    if world > 1 and not args.no_dp_check and args.optimizer == "none":
        from oracle import msunet_oracle as O            # noise generator only (test infrastructure used as the checker)
        core = model.module
        ocfg = O.Cfg(img_size=S, embed_dim=arch["embed_dim"], depths=tuple(arch["depths"]), num_heads=tuple(arch["num_heads"]))
        names = ["patch_embed.proj.weight", "layers.0.blocks.1.attn.qkv.weight", "layers.2.blocks.3.mlp.0.weight",
                 "concat_back_dim.3.weight", "layers_up.3.blocks.1.attn.relative_position_bias_table", "up.refine2.weight", "output.weight"]
        pd = dict(core.named_parameters())
        shards = [synth_batch(B, S, 4321 + r) for r in range(world)]
        noise = [O.draw_sd_noise(ocfg, MB, args.drop_path, seed=1000 + r) for r in range(world)]
        # (1) the data-parallel step on this rank's shard
        core.inject_drop_path_noise(noise[rank] if args.drop_path > 0 else None)
        loss_dp = step(x_d, y_d).detach().float().clone()
        dist.all_reduce(loss_dp, op=dist.ReduceOp.SUM)
        got = {k: pd[k].grad.detach().clone() for k in names}
        # (2) single process over the global batch: shard by shard, gradients summed, divided by N
        for p_ in params:
            p_.grad = None
        loss_sp = torch.zeros((), device=dev)
        with model.no_sync():
            for r in range(world):
                core.inject_drop_path_noise(noise[r] if args.drop_path > 0 else None)
                xr, yr = shards[r][0].to(dev), shards[r][1].to(dev)
                for i in range(n_micro):
                    l_ = crit(core(xr[i * MB:(i + 1) * MB]), yr[i * MB:(i + 1) * MB])
                    (l_ / (n_micro * world)).backward()
                loss_sp += l_.detach().float() if n_micro == 1 else 0.0
        core.inject_drop_path_noise(None)
        ops.next_grad_pass()
        worst = 0.0
        for k in names:
            ref_g = pd[k].grad
            worst = max(worst, float((got[k] - ref_g).abs().max() / ref_g.abs().max().clamp_min(1e-30)))
        for p_ in params:
            p_.grad = None
        w_t = torch.tensor([worst], device=dev)
        dist.all_reduce(w_t, op=dist.ReduceOp.MAX)
        dp_check = {"max_rel": float(w_t.item()), "tensors": names,
                    "loss_rel": (abs(float(loss_dp.item()) - float(loss_sp.item())) / abs(float(loss_sp.item()))) if n_micro == 1 else None,
                    "bucket_writes": dict(model.stats),
                    "how": "rank-averaged gradients after the NCCL exchange vs one process replaying every rank's shard (same "
                           "stochastic-depth noise), max |a-b| / max |b| over the listed tensors, max over ranks"}
        barrier()

    # ---------------- CPU baseline (rank 0, N=1 only): the oracle port on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cstep, cores, ckind = cpu_reference_step_factory(S, 1)
        cstep()
        t0 = time.perf_counter()
        n = 2
        for _ in range(n):
            cstep()
        cdt = (time.perf_counter() - t0) / n
        cpu = {"value": 1.0 / cdt, "unit": UNIT, "cores": cores, "kind": ckind,
               "sample": f"T96 {S}x{S} fwd+DynamicLoss+bwd on 1 image/step, {n} timed steps after 1 warm-up "
                         f"({'the reference modules from baseline/_ref' if ckind == 'reference' else 'oracle port'}, fp32)"}

    if rank == 0:
        emit(({
            "metric": metric_name(S), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"MS-UNet {args.model} ({'embed 96, depths 2-2-6-2' if args.model == 'T96' else 'embed 128, depths 2-2-18-2'}, window 7) training step fwd+DynamicLoss+bwd, "
                                   f"{S}x{S}, batch {B}/GPU (global {B * world})"
                                   + (f" in {n_micro} micro-batches of {MB} with gradient accumulation" if n_micro > 1 else "")
                                   + f", drop_path {args.drop_path}, attn_drop {args.attn_drop}",
                       "parallelism": f"dp{world}", "cuda_graph": graph is not None,
                       "l2": "working set >> L2: ~10 GB of activations are written and re-read every step",
                       "e2e_loop": "per step: H2D of that step's images + masks from pinned host memory (copy stream, one step ahead), "
                                   "the step through the public API, D2H of its loss into a pinned ring buffer read on the host up to "
                                   "3 steps later (the last ones before the clock stops)",
                       "optimizer": {"none": "excluded (metric is fwd+bwd)",
                                     "fused": "included: gradient all-reduce + replicated one-launch FusedAdamW",
                                     "sharded": "included: reduce-scatter -> AdamW on the rank's shard -> all-gather per bucket, "
                                                "overlapped with backward (dp.ShardedAdamW)"}[args.optimizer]},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "e2e_uint8_staging": e2e_u8,
            "gpu_launches": int(launches),
            "model_tflops": value * gflop_img / 1e3 / world,
            "roofline": roof, "cpu_baseline": cpu, "dp_check": dp_check,
        }))
    if world > 1:
        # tear down without ever hanging the launcher: captured graphs hold NCCL work, so give the orderly
        # shutdown a few seconds and then leave
        sys.stdout.flush()
        threading.Timer(8.0, lambda: os._exit(0)).start()
        try:
            barrier()
            del graph
            dist.destroy_process_group()
        finally:
            os._exit(0)


if __name__ == "__main__":
    main()
