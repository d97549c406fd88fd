"""What the gradient all-reduce of a ViT-B (86.6 M fp32 gradients) costs on this box, outside the training step:
one flat buffer vs a NCCL group over the ~150 individual gradient tensors vs bf16.  Run under torchrun."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    import favit_b200  # noqa: F401
    from favit_b200.dp import NcclComm
    from favit_b200.models import VisionTransformerMHLA
    comm = NcclComm()
    m = VisionTransformerMHLA(img_size=224, patch_size=16, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                              window_size=7, use_mhla=True)
    shapes = [p.shape for p in m.parameters()]
    del m
    grads = [torch.randn(s, device=dev) for s in shapes]
    n = sum(g.numel() for g in grads)
    flat = torch.randn(n, device=dev)
    big = [g for g in grads if g.numel() >= 65536]
    small_flat = torch.randn(sum(g.numel() for g in grads if g.numel() < 65536), device=dev)

    def timeit(fn, iters=10):
        for _ in range(3):
            torch.cuda.current_stream().wait_event(fn())
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            torch.cuda.current_stream().wait_event(fn())
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    res = {
        "flat fp32 (1 tensor)": timeit(lambda: comm.all_reduce_avg([flat])),
        f"group of {len(grads)} tensors": timeit(lambda: comm.all_reduce_avg(grads)),
        f"group of {len(big)} large + 1 flat of the small ones": timeit(lambda: comm.all_reduce_avg(big + [small_flat])),
        "torch.distributed all_reduce flat fp32": timeit(lambda: (dist.all_reduce(flat, op=dist.ReduceOp.AVG), torch.cuda.current_stream().record_event())[1]),
    }
    flat_b = flat.to(torch.bfloat16)
    res_b = {
        "flat bf16 (1 tensor, 173 MB)": timeit(lambda: comm.all_reduce_avg([flat_b])),
        "one 32 MB-bucket's worth in bf16 (16 MB)": timeit(lambda: comm.all_reduce_avg([flat_b[:8 << 20]])),
        "12 x 14 MB bf16 back to back": timeit(lambda: [comm.all_reduce_avg([flat_b[i * (7 << 20):(i + 1) * (7 << 20)]])
                                                        for i in range(12)][-1]),
    }
    if rank == 0:
        print(f"world size {dist.get_world_size()}", flush=True)
        for k, v in res_b.items():
            print(f"{k:55s} {v:7.3f} ms", flush=True)
        for k, v in res.items():
            print(f"{k:55s} {v:7.3f} ms  {n * 4 / v / 1e6:7.1f} GB/s algbw", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
