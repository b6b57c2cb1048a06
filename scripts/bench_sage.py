"""cfg4 timing (BASELINE.json configs[3], SURVEY.md 8d "Config 4"): three GraphSageConv(mean) layers
100 -> 128 -> 128 -> 47 over a synthetic stand-in with the published ogbn-products shape
(N = 2 449 029, 61 859 140 undirected = 123 718 280 directed edges, X [N, 100] f32; the dataset itself
is not on the box: Chung-Lu weights, exponent 0.6, seed 5, symmetrised).  Prints per-layer device times,
the mean-aggregation's fraction of the HBM roofline (B_alg = E(4D+8) + N*4D + (N+1)*8 per SURVEY 8d),
and the sector efficiency note for 400-byte rows.  Not run in round 1 (written after the GPU budget
was spent); usage: python scripts/bench_sage.py [--nodes N --und-edges M]."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (gen_edges, measured peak)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=2_449_029)
    ap.add_argument("--und-edges", type=int, default=61_859_140)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    import pgl_b200 as pgl
    dev = torch.device("cuda", 0)
    n = args.nodes
    half = bench.gen_edges(torch, n, args.und_edges, 0.6, 5, dev)
    edges = torch.cat([half, half.flip(1)], 0)
    del half
    e = int(edges.shape[0])
    g = pgl.Graph(edges=edges, num_nodes=n)
    torch.manual_seed(4)
    x = torch.randn(n, 100, device=dev)
    dims = [100, 128, 128, 47]
    layers = [pgl.nn.GraphSageConv(dims[i], dims[i + 1], aggr_func="mean").to(dev) for i in range(3)]
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def timed(fn):
        for _ in range(3):
            out = fn()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(args.iters):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / args.iters, out

    peak = bench.peaks()[0]
    res = {"workload": "cfg4 products-shape stand-in, %d nodes / %d directed edges" % (n, e),
           "max_in_degree": int(g.indegree().max()), "layers": []}
    h = x
    with torch.no_grad():
        for i, conv in enumerate(layers):
            d = dims[i]
            t_agg, _ = timed(lambda: g.send_recv(h, "mean"))
            t_layer, out = timed(lambda: conv(g, h, act="relu" if i < 2 else None))
            b_alg = e * (4 * d + 8) + n * 4 * d + (n + 1) * 8
            res["layers"].append({"in": d, "out": dims[i + 1], "aggregation_ms": t_agg, "layer_ms": t_layer,
                                  "agg_edges_per_s": e / (t_agg * 1e-3),
                                  "agg_alg_GBs": b_alg / t_agg / 1e6,
                                  "agg_roofline_frac": b_alg / t_agg / 1e6 / peak})
            h = out
    res["total_forward_ms"] = sum(l["layer_ms"] for l in res["layers"])
    print(json.dumps(res))


if __name__ == "__main__":
    main()
