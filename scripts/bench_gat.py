"""cfg3 timing: GATConv(128 -> 8 heads x 16) forward on an RMAT scale-20 graph with 10M edges
(BASELINE.json configs[2]).  Prints per-op device times and the layer's fraction of the HBM
roofline with BASELINE.md's algorithmic bytes for a fused layer (6.42 GB)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rmat(scale, e, a=0.57, b=0.19, c=0.19, seed=1, device="cuda"):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    src = torch.zeros(e, dtype=torch.int64, device=device)
    dst = torch.zeros(e, dtype=torch.int64, device=device)
    for lvl in range(scale):
        r = torch.rand(e, generator=g, device=device)
        sb = (r >= a + b).to(torch.int64)                       # quadrants c, d -> lower half
        db = (((r >= a) & (r < a + b)) | (r >= a + b + c)).to(torch.int64)  # b, d -> right half
        src = src * 2 + sb
        dst = dst * 2 + db
    return torch.stack([src, dst], 1)


def main():
    import pgl_b200 as pgl
    import pgl_b200.nn.functional as GF
    dev = torch.device("cuda", 0)
    n, e, H, Dh = 1 << 20, 10_000_000, 8, 16
    edges = rmat(20, e)
    g = pgl.Graph(edges=edges, num_nodes=n)
    torch.manual_seed(2)
    x = torch.randn(n, 128, device=dev)
    conv = pgl.nn.GATConv(128, Dh, feat_drop=0, attn_drop=0, num_heads=H).to(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def timed(fn, iters=20):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(iters):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters, out

    with torch.no_grad():
        f = (x @ conv.linear.weight + conv.linear.bias).reshape(-1, H, Dh)
        a_s = (f * conv.weight_src).sum(-1)
        a_d = (f * conv.weight_dst).sum(-1)
        t_uv, al = timed(lambda: g.send_uv(a_s, a_d, "add"))
        t_lr, al2 = timed(lambda: torch.nn.functional.leaky_relu(al, 0.2))
        t_sm, alpha = timed(lambda: GF.edge_softmax(g, al2))
        t_ue, out = timed(lambda: g.send_ue_recv(f, alpha.reshape(-1, H, 1), "mul", "sum"))
        t_layer, _ = timed(lambda: conv(g, x))
        csr = g._fwd_csr()
        t_att, alpha_s = timed(lambda: pgl.ops.gat_attention_csr(csr, a_s, a_d, 0.2))
        t_ues, _ = timed(lambda: pgl.ops.aggregate_ue_slots(f, alpha_s.reshape(-1, H, 1), csr, n, "mul", "sum"))
        t_one, _ = timed(lambda: pgl.ops.gat_fused(csr, f, a_s, a_d, 0.2))
        t_gemm, _ = timed(lambda: x @ conv.linear.weight + conv.linear.bias)
    peak = 6582.5
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    b_alg = 6.42e9
    agg_ms = t_uv + t_lr + t_sm + t_ue
    print(json.dumps({
        "workload": "cfg3 RMAT scale 20, 10M edges, GATConv 128 -> 8x16, eval", "max_in_degree": g.adj_dst_index.max_degree,
        "ms": {"send_uv": t_uv, "leaky_relu(torch)": t_lr, "edge_softmax": t_sm, "send_ue_recv": t_ue,
               "attention+aggregation": agg_ms, "linear(fp32 gemm)": t_gemm, "GATConv.forward(fused inference)": t_layer,
               "fused attention": t_att, "slot-ordered aggregation": t_ues, "single-pass fused GAT aggregation": t_one},
        "roofline_frac_single_pass": b_alg / (t_one * 1e-3) / 1e9 / peak,
        "roofline_frac_fused_pair": b_alg / ((t_att + t_ues) * 1e-3) / 1e9 / peak,
        "edges_per_s_attention_aggregation": e / (agg_ms * 1e-3),
        "roofline_frac_fused_model": b_alg / (agg_ms * 1e-3) / 1e9 / peak,
        "send_ue_recv_alg_GBs": (e * (512 + 8 + 8 + 32) + n * 512) / (t_ue * 1e-3) / 1e9}))


if __name__ == "__main__":
    main()
