"""Times pglb_csr_build (hand-written scan + stable radix sort) and checks it against torch's
stable sort on the same keys.  usage: python scripts/time_csr_build.py [E] [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pgl_b200 import ops

E = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
g = torch.Generator(device="cuda").manual_seed(1)
u = torch.randint(0, N, (E,), device="cuda", generator=g)
v = torch.randint(0, N, (E,), device="cuda", generator=g)
for _ in range(2):
    out = ops.csr_build(u, v, N)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    out = ops.csr_build(u, v, N)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
degree, sv, su, se, indptr = out
want_u, want_e = torch.sort(u, stable=True)
ok = bool((su == want_u).all() and (se == want_e).all() and (sv == v[want_e]).all()
          and (indptr[1:] - indptr[:-1] == torch.bincount(u, minlength=N)).all())
print({"E": E, "N": N, "csr_build_ms": round(ms, 2), "bit_exact_vs_stable_sort": ok})
