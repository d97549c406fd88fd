#!/usr/bin/env python
"""Benchmark of the favit_b200 hot path: images/sec, fwd+bwd, ViT-MHLA 224px on N x B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl favit|reference]

N > 1 is launched by `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...` (one rank per GPU,
NCCL).  Rank 0 prints ONE JSON line.  A step = zero_grad -> forward (bf16 autocast) -> cross-entropy -> backward with
the bucketed gradient all-reduce overlapped -> AdamW step, on a fixed per-GPU batch of synthetic images (weak scaling).

Workloads:
  vitb16_mhla_224    VisionTransformerMHLA ViT-B/16, 224px, window 7, per-GPU batch 256       (BASELINE configs[3]; default)
  sppp_vits_mhla_224 SPPPViTMHLA ViT-S/16, 224px, 16 superpixels, window 7, per-GPU batch 256  (BASELINE configs[1])
  vit_tiny_cifar_32  VisionTransformerMHLA tiny, 32px / patch 4, D 192, dropout 0.1, batch 64  (BASELINE configs[0])
The default run also reports, under "also": configs[1], configs[0], configs[4] (512 px SPPP inference, every rank runs its
32-image share of the 256-image batch, no communication) and configs[2] (the attention / SPPP microbenchmark sweeps).

`--impl reference` times the CPU restatement of the reference (oracle/, the reference itself is Python and does not
travel to the GPU box) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    "vitb16_mhla_224": dict(kind="vit", img=224, ps=16, D=768, depth=12, H=12, W=7, classes=1000, B=256, cpu_B=8),
    "sppp_vits_mhla_224": dict(kind="sppp", img=224, ps=16, D=384, depth=12, H=6, W=7, K=16, classes=1000, B=256,
                               cpu_B=16),
    # BASELINE configs[0]: the reference's own CPU-runnable case, main.py-style (dropout 0.1 in the MLP, main.py:106;
    # batch 64, main.py:90; AdamW lr 1e-4 wd 0.05, main.py:129-132); the CPU arm runs it at the full batch
    "vit_tiny_cifar_32": dict(kind="vit", img=32, ps=4, D=192, depth=12, H=3, W=7, classes=10, B=64, cpu_B=64,
                              dropout=0.1),
}
METRIC = "images/sec fwd+bwd ViT-MHLA 224px"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback")


def build_model(wl, device):
    import favit_b200  # noqa: F401
    from favit_b200.models import SPPPViTMHLA, VisionTransformerMHLA
    torch.manual_seed(1234)
    if wl["kind"] == "vit":
        m = VisionTransformerMHLA(img_size=wl["img"], patch_size=wl["ps"], num_classes=wl["classes"], embed_dim=wl["D"],
                                  depth=wl["depth"], num_heads=wl["H"], window_size=wl["W"], use_mhla=True,
                                  dropout=wl.get("dropout", 0.0))
    else:
        m = SPPPViTMHLA(img_size=wl["img"], patch_size=wl["ps"], num_classes=wl["classes"], embed_dim=wl["D"],
                        depth=wl["depth"], num_heads=wl["H"], num_superpixels=wl["K"], window_size=wl["W"],
                        use_mhla=True, pooling_type="mean", dropout=wl.get("dropout", 0.0))
    return m.to(device)


def make_batch(wl, B, seed, device):
    from favit_b200 import synth
    x = synth.images(B, wl["img"], seed=seed, device=device)
    y = synth.class_labels(B, wl["classes"], seed=seed, device=device)
    maps = None
    if wl["kind"] == "sppp":
        maps = synth.voronoi_label_maps(B, wl["img"], wl["K"], seed=seed, device=device, exact_k=True,
                                        patch_size=wl["ps"])
    return x, y, maps


class ClockSampler:
    """SM clock and throttle reasons during the timed region (pynvml; B200_PROFILING.md 'clocks line')."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        if not self.samples:
            return None
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_oracle_step_fn(wl, B):
    """Returns (step, n_images): one fwd + CE + bwd + AdamW step of the oracle model on B images (fp32, CPU)."""
    import oracle
    from favit_b200.models import SPPPViTMHLA, VisionTransformerMHLA  # constructors only (parameter shapes / init)
    torch.manual_seed(1234)
    if wl["kind"] == "vit":
        m = VisionTransformerMHLA(img_size=wl["img"], patch_size=wl["ps"], num_classes=wl["classes"], embed_dim=wl["D"],
                                  depth=wl["depth"], num_heads=wl["H"], window_size=wl["W"], use_mhla=True)
    else:
        m = SPPPViTMHLA(img_size=wl["img"], patch_size=wl["ps"], num_classes=wl["classes"], embed_dim=wl["D"],
                        depth=wl["depth"], num_heads=wl["H"], num_superpixels=wl["K"], window_size=wl["W"],
                        use_mhla=True)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    del m
    opt = torch.optim.AdamW(list(sd.values()), lr=1e-4, weight_decay=0.05)
    x, y, maps = make_batch(wl, B, seed=1234, device="cpu")

    def step():
        opt.zero_grad(set_to_none=True)
        if wl["kind"] == "vit":
            logits = oracle.vit_mhla_forward(x, sd, wl["ps"], wl["H"], wl["W"], mlp_dropout=wl.get("dropout", 0.0))
        else:
            logits = oracle.sppp_vit_mhla_forward(x, maps, sd, wl["ps"], wl["H"], wl["W"], wl["K"])
        loss = torch.nn.functional.cross_entropy(logits, y)
        loss.backward()
        opt.step()
        return float(loss.detach())

    return step, B


def time_cpu_oracle(wl, B, warmup, steps):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, n = cpu_oracle_step_fn(wl, B)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return n / dt, dt * 1e3, cores


def run_reference_arm(args, wl, rank):
    if rank != 0:
        return
    B = wl["cpu_B"]
    ips, ms, cores = time_cpu_oracle(wl, B, args.warmup, args.steps)
    sample = f"{B} images/step of {args.workload}, fp32, oracle port of the reference modules, AdamW step included"
    out = {
        "impl": "reference", "metric": METRIC, "value": round(ips, 3), "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the favit arm's config; each timed step is a bounded sample of it (B images instead of the per-GPU batch)
        "config": {"workload": args.workload, "global_batch": wl["B"] * max(args.gpus, 1), "per_gpu_batch": wl["B"],
                   "img": wl["img"], "patch": wl["ps"], "embed_dim": wl["D"], "depth": wl["depth"], "heads": wl["H"],
                   "window": wl["W"], "superpixels": wl.get("K"), "sample_batch_per_step": B, "device": "cpu",
                   "optimizer": "adamw in step"},
        "cpu_baseline": {"value": round(ips, 3), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(ips, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------------------
# favit arm
# --------------------------------------------------------------------------------------------------
def timed_steps(fn, steps, dist_on, device):
    """K steps bracketed by barrier + synchronize, timed with CUDA events; returns the max over ranks (ms total)."""
    import torch.distributed as dist
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize(device)
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    timed_steps.by_rank = [float(ms.item())]
    if dist_on:
        dist.barrier()
        every = [torch.zeros_like(ms) for _ in range(dist.get_world_size())]
        dist.all_gather(every, ms)
        timed_steps.by_rank = [float(t.item()) for t in every]      # each rank's own elapsed time (ms, total)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def graph_time_us(fn, iters, nbuf):
    """Average duration of `fn(i % nbuf)` with `iters` launches captured in one CUDA graph (best of 5 replays): what a
    ~10 us kernel costs inside the captured training step, without the host launch rate in the way."""
    for i in range(2):
        fn(i % nbuf)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    keep = []
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for i in range(iters):
                keep.append(fn(i % nbuf))
    best = None
    for _ in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / iters * 1e3
        best = t if best is None else min(best, t)
    return best


def sppp_kernel_rooflines(wl, B, device, peaks):
    """The three SPPP launches of a step, each timed alone through a CUDA graph of 16 launches over rotating input
    buffers (8 x 38.5 MB of embeddings at the workload's batch: nothing survives in the 126 MB L2), against the measured
    copy bandwidth.  At the workload's batch a launch moves 45-103 MB, i.e. 7-16 us at peak, so ~5 us of launch, ramp
    and tail weigh in; `at_4x_batch` repeats the measurement at four times the batch (3 buffers), where they do not."""
    from favit_b200 import ops, synth
    S, ps, K, D = wl["img"], wl["ps"], wl["K"], wl["D"]
    P = (S // ps) ** 2

    def measure(Bm, nbuf, iters):
        lms = [synth.voronoi_label_maps(Bm, S, K, seed=77 + i, device=device, exact_k=True, patch_size=ps)
               for i in range(nbuf)]
        asg = [ops.sppp_assign(lm, ps, S, K) for lm in lms]
        x = [torch.randn(Bm, P, D, device=device).to(torch.bfloat16) for _ in range(nbuf)]
        g = [torch.randn(Bm, K, D, device=device) for _ in range(nbuf)]
        res = {}

        def rec(name, us, nbytes):
            gbps = nbytes / us / 1e3
            res[name] = {"us_per_launch": round(us, 2), "algorithmic_bytes": int(nbytes), "gbps": round(gbps, 1),
                         "frac_hbm": round(gbps / peaks["hbm"], 4)}

        rec("sppp_assign", graph_time_us(lambda i: ops.sppp_assign(lms[i], ps, S, K), iters, nbuf),
            Bm * S * S * 8.0 + 2.0 * Bm * P * 4 + Bm * K * 4)
        rec("sppp_pool_fwd",
            graph_time_us(lambda i: ops.sppp_pool_fwd(x[i], asg[i][6], asg[i][5], asg[i][2], K, torch.float32), iters, nbuf),
            Bm * P * D * 2.0 + Bm * P * 4 + Bm * K * D * 4.0 + Bm * K * 4)
        rec("sppp_pool_bwd",
            graph_time_us(lambda i: ops.sppp_pool_bwd(g[i], asg[i][1], asg[i][3], torch.bfloat16), iters, nbuf),
            Bm * K * D * 4.0 + Bm * P * 4 + Bm * P * D * 2.0)
        return res

    out = measure(B, 8, 16)
    big = measure(4 * B, 3, 9)
    for k, v in out.items():
        v["timing"] = "CUDA graph of 16 launches over 8 rotating buffers (L2-cold), CUDA events"
        v["at_4x_batch"] = {"batch": 4 * B, **{kk: big[k][kk] for kk in ("us_per_launch", "gbps", "frac_hbm")}}
    return out


def attn_sweep(device, peaks, iters=8):
    """BASELINE configs[2], attention part: favit_mhla_attn_fwd / bwd over sequence lengths 17..4097 (latent tokens
    16-256 + cls, patch tokens 64-4096 + cls), 3 and 12 heads of 64, windows 7 (reference default) / 15 / 31 / 63, bf16,
    B*N ~ 64k tokens; each kernel in a CUDA graph of `iters` launches over 3 rotating inputs.  Algorithmic bytes:
    forward 4 B N D e, backward 8 B N D e; `frac` is of the measured copy bandwidth."""
    from favit_b200 import _lib as L, raw
    bf = torch.bfloat16
    rows = []
    for N in (17, 65, 197, 257, 1025, 4097):
        for H in (3, 12):
            for W in (7, 15, 31, 63):
                hd = 64
                D = H * hd
                B = max(1, 65536 // N)
                M = B * N
                qkv = [torch.randn(M, 3 * D, device=device).to(bf) for _ in range(3)]
                do = [torch.randn(M, D, device=device).to(bf) for _ in range(3)]
                fwd = graph_time_us(lambda i: raw.attn_fwd(qkv[i], B, N, H, hd, W), iters, 3)
                kern = L.last_kernel().split(" ")[0]
                o, lse = raw.attn_fwd(qkv[0], B, N, H, hd, W)
                bwd = graph_time_us(lambda i: raw.attn_bwd(qkv[0], o, lse, do[i], B, N, H, hd, W)[0], iters, 3)
                fb, bb = 4.0 * M * D * 2, 8.0 * M * D * 2
                rows.append({"N": N, "H": H, "W": W, "B": B, "kernel": kern, "fwd_us": round(fwd, 1),
                             "fwd_frac_hbm": round(fb / fwd / 1e3 / peaks["hbm"], 3), "bwd_us": round(bwd, 1),
                             "bwd_frac_hbm": round(bb / bwd / 1e3 / peaks["hbm"], 3),
                             "gflops_fwd": round(4.0 * M * W * D / fwd / 1e3, 1)})
                del qkv, do, o, lse
    return {"what": "MHLA window-attention microbenchmark (BASELINE configs[2]); bf16, head_dim 64, B*N ~ 64k tokens",
            "timing": f"CUDA graph of {iters} launches over 3 rotating inputs, best of 5 replays, CUDA events",
            "rows": rows}


def sppp_sweep(device, peaks, iters=8):
    """BASELINE configs[2], SPPP part: favit_sppp_assign / pool_fwd / pool_bwd over patch tokens P in {64..4096},
    superpixel (latent) tokens R in {16, 64, 256}, D in {192, 384, 768}, bf16 embeddings, B*P ~ 50k patches."""
    from favit_b200 import ops, synth
    shapes = {64: (32, 4), 196: (224, 16), 256: (128, 8), 1024: (256, 8), 4096: (512, 8)}
    rows = []
    for P, (S, ps) in shapes.items():
        for R in (16, 64, 256):
            if R * 4 > P:
                continue            # fewer than 4 patches per superpixel: not a configuration the models can reach
            B = max(2, 50176 // P)
            lms = [synth.voronoi_label_maps(B, S, R, seed=5 + i, device=device, patch_size=ps) for i in range(3)]
            asg = [ops.sppp_assign(lm, ps, S, R) for lm in lms]
            t_as = graph_time_us(lambda i: ops.sppp_assign(lms[i], ps, S, R), iters, 3)
            for D in (192, 384, 768):
                x = [torch.randn(B, P, D, device=device).to(torch.bfloat16) for _ in range(3)]
                g = [torch.randn(B, R, D, device=device) for _ in range(3)]
                tf = graph_time_us(lambda i: ops.sppp_pool_fwd(x[i], asg[i][6], asg[i][5], asg[i][2], R, torch.float32), iters, 3)
                tb = graph_time_us(lambda i: ops.sppp_pool_bwd(g[i], asg[i][1], asg[i][3], torch.bfloat16), iters, 3)
                bf_ = B * P * D * 2.0 + B * P * 4 + B * R * D * 4.0 + B * R * 4
                bb_ = B * R * D * 4.0 + B * P * 4 + B * P * D * 2.0
                ba_ = B * S * S * 8.0 + 2.0 * B * P * 4 + B * R * 4
                rows.append({"P": P, "R": R, "D": D, "B": B, "assign_us": round(t_as, 1),
                             "assign_frac_hbm": round(ba_ / t_as / 1e3 / peaks["hbm"], 3), "pool_fwd_us": round(tf, 1),
                             "pool_fwd_frac_hbm": round(bf_ / tf / 1e3 / peaks["hbm"], 3), "pool_bwd_us": round(tb, 1),
                             "pool_bwd_frac_hbm": round(bb_ / tb / 1e3 / peaks["hbm"], 3)})
                del x, g
    return {"what": "SPPP microbenchmark (BASELINE configs[2]): patch tokens x superpixel tokens x width",
            "timing": f"CUDA graph of {iters} launches over 3 rotating inputs, best of 5 replays, CUDA events",
            "rows": rows}


def load_traffic(workload):
    """Per-launch DRAM bytes of the roofline kernel from the committed ncu capture (profiles/r2_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        return json.load(open(p)).get(workload)
    except Exception:
        return None


def run_infer(device, steps, warmup, rank=0, world=1, dist_on=False):
    """BASELINE configs[4]: high-resolution SPPP + MHLA inference (512 px, patch 8 -> 4096 patch tokens pooled to 64
    superpixel tokens), 32 images per GPU: the batch of 256 sharded over 8 GPUs, no communication (every rank runs its
    own images; with fewer ranks the global batch is world x 32).  Forward only, eval mode, bf16 autocast, captured in a
    CUDA graph; inputs resident in HBM, two alternating batches; time = max over ranks."""
    from favit_b200 import synth
    from favit_b200.models import SPPPViTMHLA
    cfg = dict(img=512, ps=8, D=384, depth=12, H=6, K=64, W=7, B=32)
    torch.manual_seed(1234)
    m = SPPPViTMHLA(img_size=cfg["img"], patch_size=cfg["ps"], num_classes=1000, embed_dim=cfg["D"], depth=cfg["depth"],
                    num_heads=cfg["H"], num_superpixels=cfg["K"], window_size=cfg["W"], use_mhla=True,
                    pooling_type="mean").to(device).eval()
    m.validate_slots = False
    batches = []
    for i in range(2):
        x = synth.images(cfg["B"], cfg["img"], seed=4321 + i + 10 * rank, device=device)
        maps = synth.voronoi_label_maps(cfg["B"], cfg["img"], cfg["K"], seed=4321 + i + 10 * rank, device=device,
                                        exact_k=True, patch_size=cfg["ps"])
        batches.append((x, maps))

    def fwd(x, maps):
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return m(x, maps)

    for i in range(3):
        fwd(*batches[i % 2])
    torch.cuda.synchronize(device)
    sx, smaps = batches[0][0].clone(), batches[0][1].clone()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fwd(sx, smaps)

    def step(i):
        x, maps = batches[i % 2]
        sx.copy_(x, non_blocking=True)
        smaps.copy_(maps, non_blocking=True)
        g.replay()
        return out

    for i in range(max(warmup, 3)):
        step(i)
    ms = timed_steps(step, steps, dist_on, device) / steps
    return {"value": round(world * cfg["B"] / (ms / 1e3), 1), "unit": "images/s, forward only (eval, no_grad)",
            "ms_per_step": round(ms, 3), "steps": steps, "n_gpus": world, "scaling": "weak (batch sharded, no collective)",
            "config": {"workload": "sppp_vits_mhla_512_infer", "global_batch": world * cfg["B"],
                       "per_gpu_batch": cfg["B"], "img": cfg["img"],
                       "patch": cfg["ps"], "patch_tokens": (cfg["img"] // cfg["ps"]) ** 2, "superpixels": cfg["K"],
                       "embed_dim": cfg["D"], "depth": cfg["depth"], "heads": cfg["H"], "window": cfg["W"],
                       "cuda_graph": True, "dtype": "bf16"}}


def run_favit(args, wl, rank, world, local_rank, keep_pg=False):
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=favit) needs a B200: the favit kernels have no CPU fallback")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    dist_on = world > 1
    if dist_on:
        dist.init_process_group("nccl", device_id=device)
    import favit_b200
    from favit_b200 import _lib as L
    from favit_b200.engine import HostBatchFeeder, TrainStep, bind_host_to_gpu
    cc = L.lib().favit_device_cc()
    if cc != 100:
        raise RuntimeError(f"favit_b200 is built for sm_100a only; this device reports compute capability {cc}")

    B = args.batch or wl["B"]
    model = build_model(wl, device)
    if wl["kind"] == "sppp":
        model.validate_slots = False          # synthetic maps are validated once, below, not once per step
    step = TrainStep(model, process_group=None, cuda_graph=args.cuda_graph, dp_mode=args.dp_mode,
                     dp_grad_dtype=torch.bfloat16 if args.dp_grad_dtype == "bf16" else torch.float32,
                     backward_gemm_tiles=None if args.backward_gemm_tiles == "auto" else args.backward_gemm_tiles)
    # distinct batches so that no step can reuse a cached input; seed differs per rank
    nb = 2
    batches = [make_batch(wl, B, seed=1234 + rank * 100 + i, device=device) for i in range(nb)]
    if wl["kind"] == "sppp":
        for (_, _, maps) in batches:
            a = model.patch_mapper.assign_batch(maps, wl["img"], r_cap=wl["K"])
            assert int(a.num_slots.min()) == wl["K"] == int(a.num_slots.max()), "synthetic label maps must give R == K"

    def dev_step(i):
        x, y, maps = batches[i % nb]
        return step(x, y, maps)

    loss_first = None
    for i in range(max(args.warmup, 5 if args.cuda_graph else 0)):   # graph mode: 3 eager steps, capture, 1 replay
        l = dev_step(i)
        if loss_first is None:
            loss_first = float(l)
    torch.cuda.synchronize(device)

    # ---- device-resident throughput (`value`) ----
    sampler = ClockSampler(local_rank)
    L.PROFILE = None if args.cuda_graph else []   # eager mode: per-launch events straight in the timed region
    launches0 = L.launch_count()
    sampler.start()
    ms_total = timed_steps(dev_step, args.steps, dist_on, device)
    ms_by_rank = [round(t / args.steps, 3) for t in timed_steps.by_rank]
    clocks = sampler.stop()
    launches = L.launch_count() - launches0
    prof, L.PROFILE = (L.PROFILE or []), None
    ms_step = ms_total / args.steps
    value = world * B / (ms_step / 1e3)
    fam_acc = {}            # (family, role) -> [work, ms, launches] over the profiled steps

    def add(rec_list):
        for name, role, work, e0, e1 in rec_list:
            f = fam_acc.setdefault((name, role), [0.0, 0.0, 0])
            f[0] += work
            f[1] += e0.elapsed_time(e1)
            f[2] += 1

    add(prof)
    prof_ms_step, prof_mode = ms_step, "events around every favit launch of the timed region (eager launches)"

    # ---- end to end from pinned host memory (`e2e`) ----
    # Label maps cross PCIe in the narrowest integer type that holds their labels (K <= 255 superpixels: uint8) and are
    # widened to the int64 the reference's maps have (sppp.py:64-66) on the device, by the copy into the step's static
    # input: 12.8 MB per step instead of 102.8 MB at C2.  The images stay fp32, the model's input type.
    def host_of(t, is_map):
        if t is None:
            return None
        if is_map and args.host_label_dtype != "int64":
            t = t.to(torch.uint8 if int(t.max()) < 256 else torch.int32)
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        return h

    # the pinned buffers are allocated with this thread bound to the GPU's own socket (NUMA-local pages: the copy does
    # not cross the inter-socket link); the previous affinity comes back right after, before any host thread pool starts
    affinity_before = bind_host_to_gpu(local_rank)
    host = [tuple(host_of(t, j == 2) for j, t in enumerate(b)) for b in batches]
    h2d = sum(t.numel() * t.element_size() for t in host[0] if t is not None)
    feeder = HostBatchFeeder(device, 3)
    loss_ring = torch.zeros(2, pin_memory=True)
    if affinity_before is not None:
        os.sched_setaffinity(0, affinity_before)
    loss_evs = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_losses = []

    def e2e_step(i):
        """Host buffers -> device (H2D of step i+1 overlaps step i) -> step -> the loss comes back to pinned host memory
        EVERY step.  The host reads the loss of step i-1 while step i runs (and the last one before the clock stops): a
        blocking read right after each launch would expose the host's launch latency — hundreds of microseconds to
        milliseconds on a shared host — in every step, which the reference's `loss.item()` loop does and a pipelined
        training loop does not."""
        if i == 0:
            feeder.prefetch(host[0])
        x, y, maps = feeder.get(i)
        if i + 1 < args.steps:
            feeder.prefetch(host[(i + 1) % nb])
        loss = step(x, y, maps)
        loss_ring[i % 2:i % 2 + 1].copy_(loss.reshape(1), non_blocking=True)
        loss_evs[i % 2].record()
        if args.e2e_loss_read == "blocking":      # A/B: the reference's `loss.item()` right after the launch
            loss_evs[i % 2].synchronize()
            e2e_losses.append(float(loss_ring[i % 2]))
            return
        if i >= 1:
            loss_evs[(i - 1) % 2].synchronize()
            e2e_losses.append(float(loss_ring[(i - 1) % 2]))
        if i == args.steps - 1:
            loss_evs[i % 2].synchronize()
            e2e_losses.append(float(loss_ring[i % 2]))

    feeder.i = 0
    e2e_ms = timed_steps(e2e_step, args.steps, dist_on, device) / args.steps
    e2e_value = world * B / (e2e_ms / 1e3)
    # what the host -> device copy of one step's inputs costs alone (the floor of an input-bound e2e step)
    torch.cuda.synchronize(device)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(feeder.stream):
        c0.record()
        for dst, src in zip(feeder.slots[0], host[0]):
            if dst is not None:
                dst.copy_(src, non_blocking=True)
        c1.record()
    torch.cuda.synchronize(device)
    h2d_alone_ms = c0.elapsed_time(c1)
    # the step must be doing real training: every loss read back is finite and the optimizer has moved it
    import math
    if not (math.isfinite(loss_first) and all(math.isfinite(v) for v in e2e_losses)):
        raise RuntimeError(f"non-finite loss in the benchmark: first {loss_first}, e2e {e2e_losses}")

    # ---- per-kernel timing inside REPLAYED steps: the step is captured once more with an external event node before
    # and after every favit launch; after each replay the events hold that launch's duration inside the running step ----
    if args.cuda_graph:
        L.PROFILE = []
        step.recapture()
        l0 = L.launch_count()
        dev_step(0)                                   # capture (one step's launches are recorded) + first replay
        per_step_launches = L.launch_count() - l0
        recs, L.PROFILE = L.PROFILE, None
        launches = per_step_launches * args.steps
        torch.cuda.synchronize(device)
        tot = 0.0
        for i in range(args.steps):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            dev_step(i)
            s1.record()
            torch.cuda.synchronize(device)
            tot += s0.elapsed_time(s1)
            add(recs)
        prof_ms_step = tot / args.steps
        prof_mode = (f"external CUDA event nodes around every favit launch inside the captured step graph, read after each of "
                     f"{args.steps} replays (instrumented step {prof_ms_step:.3f} ms vs {ms_step:.3f} ms uninstrumented)")
    prof_total_ms = prof_ms_step * args.steps

    fam = {}
    for (name, role), v in fam_acc.items():
        f = fam.setdefault(name, [0.0, 0.0, 0])
        for j in range(3):
            f[j] += v[j]
    peaks = load_peaks()
    gemm = [fam[k] for k in ("gemm_fwd", "gemm_dgrad", "gemm_wgrad") if k in fam]
    g_work, g_ms, g_n = (sum(v[0] for v in gemm), sum(v[1] for v in gemm), sum(v[2] for v in gemm)) if gemm else (0, 1, 0)
    achieved = g_work / (g_ms * 1e-3) / 1e12
    roofline = {
        "kernel": "gemm_bf16_tcgen05_2cta_kernel (+ the single-CTA gemm_bf16_tcgen05_kernel for the classification head): "
                  "every GEMM launch of the profiled steps (patch embedding, MHLA qkv / proj, the block's fc1 / fc2, head; "
                  "forward, dgrad, wgrad)",
        "bound": "tensor", "achieved": round(achieved, 2), "peak": peaks["tf_sust"], "unit": "TFLOP/s",
        "frac": round(achieved / peaks["tf_sust"], 4), "traffic": None, "peak_source": peaks["source"] + " (sustained)",
        "launches": g_n, "avg_us_per_launch": round(g_ms * 1e3 / max(g_n, 1), 2), "share_of_step": round(g_ms / prof_total_ms, 4),
        "timing": prof_mode,
    }
    # the metric's "MHLA kernel % of roofline": the MHLA module alone (mhla.py:85-161 and its autograd) = qkv GEMMs +
    # window attention + proj GEMMs + latent fold, algorithmic FLOPs F_min = B N (8 D^2 + 4 W D) per layer forward
    # (the latent projection is folded away), x3 for forward + backward
    n_tok = (wl["img"] // wl["ps"]) ** 2 + 1 if wl["kind"] == "vit" else wl["K"] + 1
    f_min = 3.0 * wl["depth"] * B * n_tok * (8.0 * wl["D"] ** 2 + 4.0 * wl["W"] * wl["D"]) * args.steps
    parts = {}
    for (name, role), v in fam_acc.items():
        key = role if role in ("qkv", "attn", "proj") else ("fold" if name == "fold" else None)
        if key:
            parts[key] = parts.get(key, 0.0) + v[1]
    mhla_ms = sum(parts.values())
    roofline_mhla = None
    if mhla_ms > 0:
        a_m = f_min / (mhla_ms * 1e-3) / 1e12
        roofline_mhla = {
            "kernel": "MHLA module: qkv GEMM + window attention + proj GEMM + latent fold, forward and backward",
            "bound": "tensor", "achieved": round(a_m, 2), "peak": peaks["tf_sust"], "unit": "TFLOP/s",
            "frac": round(a_m / peaks["tf_sust"], 4), "algorithmic_flops_per_step": f_min / args.steps,
            "ms_per_step": {k: round(v / args.steps, 4) for k, v in parts.items()},
            "share_of_step": round(mhla_ms / prof_total_ms, 4)}
    families = {k: {"launches": v[2], "ms_per_step": round(v[1] / args.steps, 4),
                    ("gbps" if k.startswith(("sppp", "attn", "ln")) else "tflops"):
                        round(v[0] / (v[1] * 1e-3) / (1e9 if k.startswith(("sppp", "attn", "ln")) else 1e12), 2)}
                for k, v in fam.items() if v[1] > 0}
    for k, v in families.items():
        if "gbps" in v:
            v["frac_hbm"] = round(v["gbps"] / peaks["hbm"], 4)
    if wl["kind"] == "sppp" and rank == 0:
        # the three 10-20 us SPPP kernels are also timed alone, L2-cold, in a graph over rotating buffers
        for k, v in sppp_kernel_rooflines(wl, B, device, peaks).items():
            v["launches"] = families.get(k, {}).get("launches", args.steps)
            v["in_step_us_per_launch"] = round(families[k]["ms_per_step"] * 1e3 * args.steps / max(families[k]["launches"], 1), 2) if k in families else None
            families[k] = v
    tr = load_traffic(args.workload)
    if tr:
        roofline["traffic"] = tr["dram_bytes_per_launch"]
        roofline["traffic_source"] = tr["source"]
        roofline["algorithmic_bytes_per_launch"] = tr.get("algorithmic_bytes_per_launch")

    out = {
        "metric": METRIC, "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "ms_per_step_by_rank": ms_by_rank,
        "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": args.workload, "global_batch": world * B, "per_gpu_batch": B, "img": wl["img"],
                   "patch": wl["ps"], "embed_dim": wl["D"], "depth": wl["depth"], "heads": wl["H"], "window": wl["W"],
                   "superpixels": wl.get("K"), "parallelism": f"dp{world}", "optimizer": "favit multi-tensor AdamW kernel, in step", "cuda_graph": bool(args.cuda_graph),
                   "dp_mode": (args.dp_mode + (" (NCCL all-reduce nodes inside the step graph)" if args.cuda_graph and
                                               args.dp_mode != "split" else "")) if dist_on else None,
                   "dp_grad_dtype": args.dp_grad_dtype if dist_on else None,
                   "backward_gemm_tiles": step.backward_gemm_tiles,
                   "grad_allreduce_calls_per_step": (len(step.reducer.buckets) if args.dp_mode == "overlap" else 1)
                   if dist_on else 0,
                   "l2": "working set >> L2 every step (inputs %.0f MB, activations several GB); no flush needed"
                         % (h2d / 1e6)},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 2), "unit": "images/s", "ms_per_step": round(e2e_ms, 3),
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "loss_read": "every step, one step behind the launch" if args.e2e_loss_read == "lagged" else
                "every step, blocking right after the launch",
                "h2d_alone_ms": round(h2d_alone_ms, 3), "h2d_alone_gbps": round(h2d / h2d_alone_ms / 1e6, 1),
                "host_numa_bound": affinity_before is not None,
                "host_dtypes": [str(t.dtype).replace("torch.", "") for t in host[0] if t is not None]},
        "gpu_launches": int(launches),
        "loss": {"first_step": round(loss_first, 4), "last_step": round(e2e_losses[-1], 4),
                 "note": "random labels, two alternating batches: the loss falls as AdamW memorises them"},
        "roofline": roofline,
        "roofline_mhla": roofline_mhla,
        "kernel_families": families,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ips, ms, cores = time_cpu_oracle(wl, wl["cpu_B"], 1, 2)
        out["cpu_baseline"] = {"value": round(ips, 3), "unit": "images/s", "cores": cores, "kind": "port",
                               "sample": f"{wl['cpu_B']} images/step of {args.workload} (fp32 oracle port, 1 warm-up + 2 "
                                         f"timed steps, AdamW included)"}
    if dist_on and not keep_pg:
        dist.barrier()
        dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="favit", choices=["favit", "reference"])
    ap.add_argument("--workload", default="vitb16_mhla_224", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", dest="also", action="store_false",
                    help="skip the secondary workload (sppp_vits_mhla_224) that the default run reports under 'also'")
    ap.add_argument("--cuda-graph", dest="cuda_graph", action="store_true",
                    help="capture the training step in a CUDA graph (default)")
    ap.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false")
    ap.add_argument("--host-label-dtype", default="narrow", choices=["narrow", "int64"],
                    help="e2e: dtype of the label maps in pinned host memory (narrow = uint8 / int32, widened on the device)")
    ap.add_argument("--dp-grad-dtype", default="bf16", choices=["fp32", "bf16"],
                    help="dtype of the gradients on the wire under data parallelism (default bf16, the benchmark's compute "
                         "dtype: one flat bf16 buffer per all-reduce, half the NVLink bytes, fp32 accumulation inside NCCL; "
                         "fp32 = the gradients travel as they are)")
    ap.add_argument("--e2e-loss-read", default="lagged", choices=["lagged", "blocking"],
                    help="e2e: read each step's loss one step behind the launch (default) or block on it right away")
    ap.add_argument("--backward-gemm-tiles", default="auto", choices=["auto", "static", "steal"],
                    help="tile scheduler of the persistent GEMM during backward: 'static' = plain striding (auto), 'steal' = "
                         "work stealing against SM contention from the overlapped NCCL all-reduce (A/B)")
    ap.add_argument("--dp-mode", default="overlap", choices=["overlap", "deferred", "split", "none"],
                    help="gradient all-reduce under data parallelism: per-bucket collectives overlapped with backward "
                         "(default), one collective after backward, or one collective outside the step graph")
    ap.set_defaults(cuda_graph=None)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "favit":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    if args.cuda_graph is None:
        args.cuda_graph = True
    if args.impl == "reference":
        run_reference_arm(args, wl, rank)
        return
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run "
                         f"--nproc-per-node {args.gpus}")
    also_on = args.also and args.workload == "vitb16_mhla_224"
    out = run_favit(args, wl, rank, world, local_rank, keep_pg=also_on)
    if also_on:
        device = torch.device("cuda", local_rank)
        also = {}
        if world == 1:
            keys = ("value", "unit", "ms_per_step", "steps", "e2e", "gpu_launches", "loss", "config", "roofline_mhla",
                    "kernel_families")
            # BASELINE configs[1] (SPPP + MHLA ViT-S, batch 256) and configs[0] (tiny CIFAR-shaped ViT + MHLA, batch 64,
            # dropout 0.1) measured in the same run
            for name, steps2, cpu in (("sppp_vits_mhla_224", max(args.steps, 20), False),
                                      ("vit_tiny_cifar_32", max(args.steps, 20), True)):
                args2 = argparse.Namespace(**vars(args))
                args2.workload, args2.no_cpu_baseline, args2.steps = name, not cpu, steps2
                o2 = run_favit(args2, WORKLOADS[name], rank, world, local_rank)
                also[name] = {k: o2[k] for k in keys + (("cpu_baseline",) if cpu and "cpu_baseline" in o2 else ())}
            peaks = load_peaks()
            also["attn_sweep"] = attn_sweep(device, peaks)          # BASELINE configs[2]
            also["sppp_sweep"] = sppp_sweep(device, peaks)
        # BASELINE configs[4]: every rank runs its 32-image share of the high-resolution inference batch
        import torch.distributed as dist
        also["sppp_vits_mhla_512_infer"] = run_infer(device, 20, 3, rank, world, world > 1)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        out["also"] = also
    if rank == 0:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
