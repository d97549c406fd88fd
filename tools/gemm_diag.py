"""GPU diagnostic for the tcgen05 GEMM: one (a_mn, b_mn, bn, splits, M, N, K) case per process.
usage: python tools/gemm_diag.py a_mn b_mn bn splits M N K
"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import favit_b200
from favit_b200 import _lib as L

a_mn, b_mn, bn, splits, M, N, K = [int(v) for v in sys.argv[1:8]]
torch.manual_seed(0)
dev = "cuda"
A = torch.randn(M, K, device=dev).to(torch.bfloat16)       # logical A[m,k]
B = torch.randn(N, K, device=dev).to(torch.bfloat16)       # logical B[n,k]
ref = A.double() @ B.double().t()
As = A.t().contiguous() if a_mn else A                      # stored [K,M] when MN-major
Bs = B.t().contiguous() if b_mn else B
C = torch.zeros(M, N, device=dev, dtype=torch.float32)
st = torch.cuda.current_stream().cuda_stream
rc = L.lib().favit_gemm_bf16_raw(As.data_ptr(), a_mn, As.stride(0), Bs.data_ptr(), b_mn, Bs.stride(0), C.data_ptr(), N,
                                 L.F32, M, N, K, bn, splits, st)
if rc:
    print("rc", rc, L.lib().favit_last_error().decode())
    sys.exit(2)
torch.cuda.synchronize()
err = (C.double() - ref).abs()
scale = ref.abs().max().item()
print(f"case a_mn={a_mn} b_mn={b_mn} bn={bn} splits={splits} M={M} N={N} K={K}: max_abs_err={err.max().item():.4g} "
      f"ref_max={scale:.4g} rel={err.max().item()/scale:.3g}")
ok = err.max().item() <= 2e-3 * scale
if not ok:
    # coarse error map: which 32x32 blocks are wrong
    bad = (err > 2e-3 * scale)
    mb, nb = (M + 31) // 32, (N + 31) // 32
    pad = torch.zeros(mb * 32, nb * 32, dtype=torch.bool, device=dev)
    pad[:M, :N] = bad
    blocks = pad.view(mb, 32, nb, 32).any(dim=3).any(dim=1).cpu()
    for r in range(min(mb, 16)):
        print("".join("X" if blocks[r, c] else "." for c in range(min(nb, 64))))
    print("C[0,:8]  ", C[0, :8].tolist())
    print("ref[0,:8]", ref[0, :8].float().tolist())
    # does C match a K-permuted / partial product? ratio check
    print("C row0 / ref row0 mean ratio", (C[0].double() / ref[0]).median().item())
print("OK" if ok else "FAIL")
sys.exit(0 if ok else 1)
