"""Build and check experimental/linear_tcgen05.cu (NOT part of the product library).

    timeout 120 python experimental/check_linear_tcgen05.py [M]

Builds experimental/_build/liblinear_tcgen05.so with nvcc (sm_100a), runs the kernel on cuda:0 against an
fp64 matmul (small M first, so a wrong result is seen before the long run) and times it next to the
mma.sync kernel of the product library (pglb_linear_tf32x3_f32).  A wrong mbarrier protocol hangs: keep the
`timeout`.  Never run so far (written after round 1's GPU budget was spent)."""
import ctypes
import json
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def build():
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "liblinear_tcgen05.so")
    src = os.path.join(HERE, "linear_tcgen05.cu")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-shared",
                               "-Xcompiler", "-fPIC", "-o", so, src])
    return so


def main():
    lib = ctypes.CDLL(build())
    fn = lib.exp_linear_tcgen05_f32
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                   ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    dev = torch.device("cuda", 0)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count

    def run(x, w, b, act):
        out = torch.empty(x.shape[0], 128, device=dev)
        rc = fn(x.data_ptr(), x.stride(0), w.data_ptr(), b.data_ptr() if b is not None else None, out.data_ptr(),
                128, x.shape[0], act, sms, torch.cuda.current_stream(dev).cuda_stream)
        assert rc == 0, rc
        return out

    res = {}
    gen = torch.Generator(device=dev).manual_seed(0)
    w = torch.randn(128, 128, device=dev, generator=gen) * 0.1
    b = torch.randn(128, device=dev, generator=gen)
    for m in (128, 1000, 128 * 148 * 3 + 17):
        x = torch.randn(m, 128, device=dev, generator=gen)
        out = run(x, w, b, 1)
        torch.cuda.synchronize()
        ref = torch.relu(x.double() @ w.double() + b.double())
        res["rel_err_M%d" % m] = float((out.double() - ref).abs().max() / ref.abs().max())
    print(json.dumps(res), flush=True)
    if max(res.values()) > 2e-5:
        print("WRONG RESULT -- not timing")
        return 1
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    x = torch.randn(M, 128, device=dev, generator=gen)

    def timeit(f, n=10):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            f()
        e.record()
        torch.cuda.synchronize()
        return a.elapsed_time(e) / n

    from pgl_b200 import ops
    res.update({"M": M, "tcgen05_ms": timeit(lambda: run(x, w, b, 1)),
                "mma_sync_ms": timeit(lambda: ops._linear_tc_raw(x, w, b, "relu")),
                "hbm_bound_ms": 2 * M * 128 * 4 / 6582.5e6})
    print(json.dumps(res))
    return 0


if __name__ == "__main__":
    sys.exit(main())
