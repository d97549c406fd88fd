"""CPU, world_size 2, gloo: the bucketed gradient all-reduce of favit_b200.dp gives every rank the gradient of the
concatenated batch (mean loss), bucket by bucket, and overlapping hooks fire for every bucket."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(12, 32), torch.nn.GELU(), torch.nn.LayerNorm(32),
                               torch.nn.Linear(32, 20), torch.nn.Tanh(), torch.nn.Linear(20, 5))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import favit_b200  # noqa: F401
    from favit_b200.dp import GradAllReducer
    m = _model()
    red = GradAllReducer(m.parameters(), bucket_mb=0.002)      # ~500 floats per bucket -> several buckets
    assert red.enabled and len(red.buckets) >= 3
    torch.manual_seed(42)
    x, y = torch.randn(8, 12), torch.randint(0, 5, (8,))
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    for _ in range(2):                                         # two steps: zero_grad must reset the views in place
        red.zero_grad()
        torch.nn.functional.cross_entropy(m(xs), ys).backward()
        red.finish()
    grads = [p.grad.clone() for p in m.parameters()]
    # deferred mode (what a CUDA-graph replay of backward uses): hooks are silent, finish() reduces the one flat buffer
    red.overlap = False
    red.zero_grad()
    torch.nn.functional.cross_entropy(m(xs), ys).backward()
    red.finish()
    for g, p in zip(grads, m.parameters()):
        assert torch.allclose(g, p.grad, rtol=1e-6, atol=1e-7)
    q.put((rank, [g.numpy() for g in grads]))      # by value: the worker may exit before the parent reads
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process_large_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    import socket
    with socket.socket() as s:            # a free port on the loopback interface
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = _model()
    torch.manual_seed(42)
    x, y = torch.randn(8, 12), torch.randint(0, 5, (8,))
    torch.nn.functional.cross_entropy(m(x), y).backward()
    for r in range(2):
        for g, p in zip(res[r], m.parameters()):
            assert torch.allclose(torch.from_numpy(g), p.grad, rtol=1e-5, atol=1e-6)


def test_single_process_is_a_noop_reducer():
    sys.path.insert(0, ROOT)
    from favit_b200.dp import GradAllReducer
    m = _model()
    red = GradAllReducer(m.parameters(), bucket_mb=0.002)
    assert not red.enabled
    red.zero_grad()
    m(torch.randn(3, 12)).sum().backward()
    red.finish()
    assert all(p.grad is not None and float(p.grad.abs().sum()) > 0 for p in m.parameters())
    assert red.grad_bytes == 4 * sum(p.numel() for p in m.parameters())
    red.zero_grad()
    assert all(p.grad is None for p in m.parameters())
