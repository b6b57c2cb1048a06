"""Diagnostic: every intermediate of the fused GAT forward / backward (out, lse, alpha_e, dz_e, the three gradients)
against a float64 torch restatement on the device, for several head layouts, on a small Chung-Lu graph and on cfg3."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pgl_b200 as pgl  # noqa: E402
from oracle import oracle as O  # noqa: E402
from pgl_b200 import ops  # noqa: E402
from pgl_b200._lib import check, lib  # noqa: E402


def truth64(f, a_s, a_d, src, dst, n, slope, go):
    f, a_s, a_d, go = [t.double() for t in (f, a_s, a_d, go)]
    f.requires_grad_(True); a_s.requires_grad_(True); a_d.requires_grad_(True)
    z = a_s[src] + a_d[dst]
    lz = torch.nn.functional.leaky_relu(z, slope)
    H = z.shape[1]
    m = torch.full((n, H), -float("inf"), device=z.device, dtype=torch.float64).scatter_reduce(
        0, dst[:, None].expand(-1, H), lz, "amax")
    p = torch.exp(lz - m[dst])
    s = torch.zeros((n, H), device=z.device, dtype=torch.float64).index_add(0, dst, p)
    al = p / s[dst]
    out = torch.zeros_like(f).index_add(0, dst, f[src] * al.unsqueeze(-1))
    out.backward(go)
    with torch.no_grad():
        lse = m + torch.log(s)
        da = (go[dst] * f[src]).sum(-1)
        delta = (go * out).sum(-1)
        dl = al * (da - delta[dst])
        dz = torch.where(z > 0, dl, dl * slope)
    return out.detach(), lse, al.detach(), dz, f.grad, a_s.grad, a_d.grad


def err(a, b, mask=None):
    a, b = a.double(), b.double()
    if mask is not None:
        a, b = a[mask], b[mask]
    d = (a - b).abs()
    return {"max_abs": float(d.max()), "ref_max": float(b.abs().max()), "argmax": int(d.reshape(-1).argmax())}


def case(name, edges, n, H, Dh, slope):
    dev = edges.device
    g = pgl.Graph(edges=edges, num_nodes=n)
    gen = torch.Generator(device=dev).manual_seed(7)
    f = torch.randn(n, H, Dh, device=dev, generator=gen)
    a_s = torch.randn(n, H, device=dev, generator=gen)
    a_d = torch.randn(n, H, device=dev, generator=gen)
    go = torch.randn(n, H, Dh, device=dev, generator=gen)
    src, dst = edges[:, 0], edges[:, 1]
    t_out, t_lse, t_al, t_dz, t_gf, t_gs, t_gd = truth64(f, a_s, a_d, src, dst, n, slope, go)
    fwd, bwd = g._fwd_csr(), g._bwd_csr()
    E = int(edges.shape[0])
    f2 = f.reshape(n, H * Dh).contiguous()
    out = torch.empty(n, H * Dh, device=dev)
    lse = torch.zeros(n, H, device=dev)
    need = ctypes.c_size_t(0)
    check(lib.pglb_spmm_csr_ws(n, E, H * Dh, ctypes.byref(need)))
    ws = ops.workspace(dev, need.value)
    P = ops._ptr
    check(lib.pglb_gat_fused_train_csr_f32(P(fwd["indptr"]), P(fwd["cols"]), P(f2), f2.stride(0), P(a_s), P(a_d), slope,
                                           P(out), out.stride(0), P(lse), n, n, E, H, Dh, P(ws), ws.numel(), ops._stream()))
    alpha = torch.empty(E, H, device=dev)
    dz = torch.empty(E, H, device=dev)
    g2 = go.reshape(n, H * Dh).contiguous()
    check(lib.pglb_gat_bwd_edge_f32(P(fwd["rows"]), P(fwd["cols"]), P(fwd["eid"]), P(f2), f2.stride(0), P(g2), g2.stride(0),
                                    P(out), out.stride(0), P(a_s), P(a_d), P(lse), slope, E, H, Dh, P(alpha), P(dz),
                                    ops._stream()))
    has = (fwd["degree"][:n] > 0)
    fa, sa, da_ = [t.detach().clone().requires_grad_(True) for t in (f, a_s, a_d)]
    o2 = ops.gat_fused_train(fwd, g._bwd_csr, fa, sa, da_, slope)
    res = {"case": name, "H": H, "Dh": Dh, "slope": slope, "E": E}
    if o2 is None:
        res["unsupported"] = True
        print(json.dumps(res), flush=True)
        return
    o2.backward(go)
    res.update({
        "out": err(out.reshape(n, H, Dh), t_out), "lse": err(lse, t_lse, has), "alpha_e": err(alpha, t_al),
        "dz_e": err(dz, t_dz), "grad_f": err(fa.grad, t_gf), "grad_attn_src": err(sa.grad, t_gs),
        "grad_attn_dst": err(da_.grad, t_gd)})
    # the op-by-op path against the same truth
    fb, sb, db = [t.detach().clone().requires_grad_(True) for t in (f, a_s, a_d)]
    import pgl_b200.nn.functional as GF
    al = g.send_uv(sb, db, "add")
    al = torch.nn.functional.leaky_relu(al, slope)
    al = GF.edge_softmax(g, al)
    o3 = g.send_ue_recv(fb, al.reshape(-1, H, 1), "mul", "sum")
    o3.backward(go)
    res["op_by_op"] = {"grad_f": err(fb.grad, t_gf), "grad_attn_src": err(sb.grad, t_gs), "grad_attn_dst": err(db.grad, t_gd)}
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    small = torch.from_numpy(O.chung_lu_edges(3000, 80000, exponent=0.9, seed=301)).to(dev)
    for H, Dh, slope in ((8, 16, 0.2), (32, 4, 0.2), (12, 8, 0.2), (8, 16, 1.0), (16, 8, 0.0)):
        try:
            case("chung-lu 3000 / 80000", small, 3000, H, Dh, slope)
        except Exception as ex:
            print(json.dumps({"case": "small", "H": H, "Dh": Dh, "slope": slope, "error": repr(ex)[:400]}), flush=True)
    if "big" in sys.argv[1:]:
        edges = bench.rmat_edges(torch, 20, 10_000_000, seed=1, device=dev)
        case("cfg3 RMAT", edges, 1 << 20, 8, 16, 0.2)
