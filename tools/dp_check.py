"""Multi-GPU check of favit_b200.dp / engine.TrainStep (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py

Every rank trains on its shard of a fixed global batch with lr = 0; the all-reduced gradients of every dp_mode
(overlap / deferred / split) x (eager / CUDA graph) must equal the gradient of the WHOLE batch computed by the same rank
without data parallelism (mean loss; every op is per image).  Prints one line per mode and exits non-zero on mismatch."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    import favit_b200  # noqa: F401
    from favit_b200.engine import TrainStep
    from favit_b200.models import VisionTransformerMHLA
    Bl = 8
    g = torch.Generator(device="cpu").manual_seed(11)
    X = torch.randn(world * Bl, 3, 64, 64, generator=g).to(dev)
    Y = torch.randint(0, 10, (world * Bl,), generator=g).to(dev)
    xs, ys = X[rank * Bl:(rank + 1) * Bl], Y[rank * Bl:(rank + 1) * Bl]

    def model():
        torch.manual_seed(3)
        return VisionTransformerMHLA(img_size=64, patch_size=8, num_classes=10, embed_dim=128, depth=3, num_heads=2,
                                     window_size=7, use_mhla=True).to(dev)

    ref_m = model()
    ref = TrainStep(ref_m, lr=0.0, weight_decay=0.0, dp_mode="none")
    ref(X, Y)
    ref_g = {k: p.grad.clone() for k, p in ref_m.named_parameters()}
    scale = max(float(v.abs().max()) for v in ref_g.values())
    bad = 0
    for graph, mode, gd in [(False, "overlap", torch.float32), (False, "deferred", torch.float32), (True, "overlap", torch.float32),
                            (True, "deferred", torch.float32), (True, "split", torch.float32), (True, "deferred", torch.bfloat16)]:
        if True:
            m = model()
            step = TrainStep(m, lr=0.0, weight_decay=0.0, cuda_graph=graph, dp_mode=mode, bucket_mb=0.5, dp_grad_dtype=gd)
            for _ in range(6 if graph else 2):
                loss = step(xs, ys)
            torch.cuda.synchronize()
            worst = 0.0
            for k, p in m.named_parameters():
                e = float((p.grad - ref_g[k]).abs().max()) / max(float(ref_g[k].abs().max()), 1e-2 * scale)
                worst = max(worst, e)
            ok = worst < 3e-2          # bf16 compute: shards vs whole batch differ by summation order only
            bad += 0 if ok else 1
            if rank == 0:
                print(f"graph={graph!s:5} dp_mode={mode:8} wire={str(gd)[6:]:8} buckets={len(step.reducer.buckets):2d} collectives={step.reducer.collectives:3d} "
                      f"worst rel grad err vs whole batch {worst:.2e} {'ok' if ok else 'MISMATCH'}", flush=True)
            step.recapture()             # a captured graph keeps its NCCL kernels alive: drop it before the communicator
            torch.cuda.synchronize()
            step.reducer.remove()
    t = torch.tensor([bad], device=dev)
    dist.all_reduce(t)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(1 if int(t.item()) else 0)


if __name__ == "__main__":
    main()
