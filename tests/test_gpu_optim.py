"""GPU parity of favit's multi-tensor AdamW (csrc/adamw.cu) against torch.optim.AdamW on the same parameters and
gradients: several steps, the reference's three parameter groups (experiments/mhla_pretrained.py:320-327), odd sizes
and unaligned views, a gradient scale, and capture in a CUDA graph."""
import pytest
import torch

from util import rel_err

pytestmark = pytest.mark.gpu


def _params(seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    shapes = [(768, 768), (2304,), (64, 64), (64,), (1000, 768), (3,), (1, 197, 768), (5, 7, 11)]
    return [torch.nn.Parameter(torch.randn(*s, device="cuda", generator=g) * 0.1) for s in shapes]


def test_adamw_matches_torch_over_steps_with_three_groups():
    from favit_b200.optim import FusedAdamW
    a, b = _params(1), _params(1)
    mk = lambda ps: [{"params": ps[:3], "lr": 1e-3}, {"params": ps[3:5], "lr": 5e-3},
                     {"params": ps[5:], "lr": 2e-3, "weight_decay": 0.0, "betas": (0.8, 0.95)}]
    ours = FusedAdamW(mk(a), lr=1e-3, weight_decay=0.05)
    ref = torch.optim.AdamW(mk(b), lr=1e-3, weight_decay=0.05)
    g = torch.Generator(device="cuda").manual_seed(2)
    for step in range(6):
        for pa, pb in zip(a, b):
            grad = torch.randn(pa.shape, device="cuda", generator=g) * (0.5 if step % 2 else 2.0)
            pa.grad, pb.grad = grad.clone(), grad.clone()
        ours.step()
        ref.step()
        for i, (pa, pb) in enumerate(zip(a, b)):
            assert rel_err(pa, pb) < 2e-6, (step, i)
    for pa, pb in zip(a, b):
        assert rel_err(ours.state[pa]["exp_avg"], ref.state[pb]["exp_avg"]) < 2e-6
        assert rel_err(ours.state[pa]["exp_avg_sq"], ref.state[pb]["exp_avg_sq"]) < 2e-6


def test_adamw_grad_scale_and_missing_grads():
    from favit_b200.optim import FusedAdamW
    a, b = _params(3), _params(3)
    ours = FusedAdamW(a, lr=1e-2, weight_decay=0.01, grad_scale=0.25)       # e.g. 1 / world after a summing all-reduce
    ref = torch.optim.AdamW(b, lr=1e-2, weight_decay=0.01)
    for pa, pb in zip(a[:-1], b[:-1]):                                       # the last parameter gets no gradient
        grad = torch.randn_like(pa)
        pa.grad, pb.grad = grad.clone(), grad * 0.25
    before = a[-1].detach().clone()
    ours.step()
    ref.step()
    for pa, pb in zip(a, b):
        assert rel_err(pa, pb) < 2e-6
    assert torch.equal(a[-1], before)


def test_adamw_inside_a_cuda_graph_advances_its_step_count():
    from favit_b200.optim import FusedAdamW
    a, b = _params(4), _params(4)
    ours = FusedAdamW(a, lr=1e-3, weight_decay=0.05)
    ref = torch.optim.AdamW(b, lr=1e-3, weight_decay=0.05)
    grads = [torch.randn_like(p) for p in a]
    for pa, pb, g in zip(a, b, grads):
        pa.grad, pb.grad = g.clone(), g.clone()
    ours.step()                       # eager step: allocates the state
    ref.step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ours.step()
    for _ in range(3):                # capture does not execute: three replays = steps 2..4
        graph.replay()
        ref.step()
    for pa, pb in zip(a, b):
        assert rel_err(pa, pb) < 5e-6


def test_train_step_with_reference_param_groups():
    """engine.TrainStep with optim.reference_param_groups == the same model stepped by torch.optim.AdamW on the same
    three groups (gradients come from the same favit kernels in both)."""
    import copy
    from favit_b200.engine import TrainStep
    from favit_b200.models import VisionTransformerMHLA
    from favit_b200.optim import reference_param_groups
    torch.manual_seed(5)
    m1 = VisionTransformerMHLA(img_size=32, patch_size=4, num_classes=10, embed_dim=128, depth=2, num_heads=2,
                               window_size=7, use_mhla=True).cuda()
    m2 = copy.deepcopy(m1)
    x = torch.randn(8, 3, 32, 32, device="cuda")
    y = torch.randint(0, 10, (8,), device="cuda")
    s1 = TrainStep(m1, lr=1e-3, weight_decay=0.05, autocast_dtype=None,
                   param_groups=reference_param_groups(m1, 1e-3, 4e-3))
    s2 = TrainStep(m2, lr=1e-3, weight_decay=0.05, autocast_dtype=None, optimizer="torch",
                   param_groups=reference_param_groups(m2, 1e-3, 4e-3))
    for _ in range(3):
        l1, l2 = s1(x, y), s2(x, y)
    assert abs(float(l1) - float(l2)) < 1e-4
    D = 128
    for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        if k.endswith("attn.qkv.bias"):
            # the K third of the qkv bias has an analytically zero gradient (a shift of all keys cancels in the softmax):
            # Adam normalises the rounding noise there into +-lr steps, which no two implementations reproduce
            p1, p2 = torch.cat([p1[:D], p1[2 * D:]]), torch.cat([p2[:D], p2[2 * D:]])
        # three AdamW steps of ~lr each; gradients of the two runs differ in their last bits (split-K reduce-add order),
        # which Adam's m / sqrt(v) turns into differences of a fraction of a step for the smallest gradients
        assert rel_err(p1, p2, floor=0.02) < 2e-3, k
