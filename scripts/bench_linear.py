"""Times the conv layers' dense transform at cfg5 size (10M x 128 @ 128 x 128, + bias, ReLU):
pglb_linear_tf32x3_f32 (3xTF32 tensor cores, fused epilogue) vs torch addmm + relu (fp32 cuBLAS)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgl_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
K = N = 128
x = torch.randn(M, K, device="cuda")
w = torch.randn(K, N, device="cuda") * 0.1
b = torch.randn(N, device="cuda")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) / n


t_tc = timeit(lambda: ops._linear_tc_raw(x, w, b, "relu"))
t_torch = timeit(lambda: torch.relu_(torch.addmm(b, x, w)))
out = ops._linear_tc_raw(x[:200000], w, b, "relu")
ref = torch.relu(x[:200000].double() @ w.double() + b.double())
err = float((out.double() - ref).abs().max() / ref.abs().max())
err_t = float((torch.relu(torch.addmm(b, x[:200000], w)).double() - ref).abs().max() / ref.abs().max())
flops = 2.0 * M * K * N
print(json.dumps({"M": M, "K": K, "N": N, "linear_tf32x3_ms": round(t_tc, 3), "torch_addmm_relu_ms": round(t_torch, 3),
                  "tf32x3_fp32_equiv_TFLOPs": round(flops / t_tc / 1e9, 1),
                  "tf32x3_hbm_GBs": round((M * K * 4 + M * N * 4) / t_tc / 1e6, 1),
                  "rel_err_vs_fp64": err, "torch_rel_err_vs_fp64": err_t}))
