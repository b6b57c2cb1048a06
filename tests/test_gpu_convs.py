"""GPU parity of the composed conv layers against oracle/oracle_conv.py -- the
literal numpy restatement of each reference ``forward`` (pgl/nn/conv.py) -- with shared weights.
fp32 tolerance 1e-4 relative (north_star), eval mode (dropout = identity)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from oracle import oracle_conv as OC

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12) if a.size else 0.0


@pytest.fixture(scope="module")
def pgl():
    import pgl_b200
    return pgl_b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make_graph(pgl, edges, n):
    g = pgl.Graph(edges=np.asarray(edges, np.int64), num_nodes=n)
    g.tensor()
    return g


def randomize(module, seed):
    """Seeded N(0, 0.3) for every parameter (biases and LayerNorm affine included, so that nothing is
    hidden behind a zero / one initialisation); returns the numpy copies."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    with torch.no_grad():
        for name, p in module.named_parameters():
            v = torch.randn(p.shape, generator=gen) * 0.3
            if name.endswith("layer_norm.weight"):
                v = v + 1.0
            p.copy_(v.to(p.device))
            out[name] = v.numpy().astype(np.float32)
    return out


def setup(pgl, n=700, e=9000, d=24, seed=1, exponent=0.8):
    edges = O.chung_lu_edges(n, e, exponent=exponent, seed=seed)
    x = np.random.default_rng(seed + 1).standard_normal((n, d)).astype(np.float32)
    return edges, x, make_graph(pgl, edges, n)


def run(conv, *args):
    conv = conv.eval()
    with torch.no_grad():
        out = conv(*args)
    # the same layer with autograd on must agree with the no-grad path and give finite gradients
    params = [p for p in conv.parameters()]
    if params:
        out_g = conv(*args)
        out_g.square().sum().backward()
        assert rel_err(out_g.detach().cpu().numpy(), out.cpu().numpy()) <= RTOL
        for p in params:
            assert p.grad is not None and bool(torch.isfinite(p.grad).all())
    return out.cpu().numpy()


def test_pinsage_conv(pgl):
    n, d, hid = 700, 24, 16
    edges, x, g = setup(pgl, n=n, d=d, seed=301)
    w_e = np.random.default_rng(303).random((edges.shape[0], 1)).astype(np.float32)
    for aggr in ("sum", "mean", "max", "min"):
        conv = pgl.nn.PinSageConv(d, hid, aggr_func=aggr).cuda()
        p = randomize(conv, 304)
        conv = conv.eval()
        with torch.no_grad():
            out = conv(g, dev(x), dev(w_e), act="relu").cpu().numpy()
        want = OC.pinsage_conv(edges, n, x, w_e, p["self_linear.weight"], p["self_linear.bias"],
                               p["neigh_linear.weight"], p["neigh_linear.bias"], aggr, act="relu")
        assert out.shape == want.shape and rel_err(out, want) <= RTOL, aggr


def test_gatv2_conv(pgl):
    n, d, H, Dh = 700, 24, 4, 8
    edges, x, g = setup(pgl, n=n, d=d, seed=311)
    for concat in (True, False):
        conv = pgl.nn.GATv2Conv(d, Dh, feat_drop=0, attn_drop=0, num_heads=H, concat=concat,
                                activation="relu").cuda()
        p = randomize(conv, 312)
        out = run(conv, g, dev(x))
        want = OC.gatv2_conv(edges, n, x, p["linear.weight"], p["linear.bias"], p["attn"], H, Dh,
                             concat=concat, activation="relu")
        assert out.shape == want.shape and rel_err(out, want) <= RTOL


def test_appnp_and_lightgcn(pgl):
    n, d = 700, 24
    edges, x, g = setup(pgl, n=n, d=d, seed=321)
    out = run(pgl.nn.LightGCNConv().cuda(), g, dev(x))
    assert rel_err(out, OC.lightgcn_conv(edges, n, x)) <= RTOL
    for self_loop in (False, True):
        conv = pgl.nn.APPNP(alpha=0.2, k_hop=6, self_loop=self_loop).cuda()
        out = run(conv, g, dev(x))
        want = OC.appnp(edges, n, x, alpha=0.2, k_hop=6, self_loop=self_loop)
        assert rel_err(out, want) <= RTOL, self_loop
    # a caller-supplied norm goes through the same fused path
    norm = np.random.default_rng(322).random((n, 1)).astype(np.float32) + 0.5
    out = run(pgl.nn.APPNP(alpha=0.1, k_hop=3).cuda(), g, dev(x), dev(norm))
    assert rel_err(out, OC.appnp(edges, n, x, alpha=0.1, k_hop=3, norm=norm)) <= RTOL


def test_gpr_conv(pgl):
    n, d, hid, od = 700, 24, 32, 7
    edges, x, g = setup(pgl, n=n, d=d, seed=331)
    for self_loop, init in ((False, "PPR"), (True, "NPPR"), (False, "SGC")):
        alpha = 2 if init == "SGC" else 0.1
        conv = pgl.nn.GPRConv(d, hid, od, drop=0.0, dprate=0.0, self_loop=self_loop, alpha=alpha,
                              k_hop=5, init_method=init).cuda()
        temp0 = conv.temp.detach().cpu().numpy().copy()
        if init == "PPR":
            ref = 0.1 * 0.9 ** np.arange(6)
            ref[-1] = 0.9 ** 5
            np.testing.assert_allclose(temp0, ref.astype(np.float32), rtol=1e-6)
        p = randomize(conv, 332)
        out = run(conv, g, dev(x))
        want = OC.gpr_conv(edges, n, x, p["linear_1.weight"], p["linear_1.bias"],
                           p["linear_2.weight"], p["linear_2.bias"], p["temp"], k_hop=5,
                           self_loop=self_loop)
        assert out.shape == (n, od) and rel_err(out, want) <= RTOL, init


def test_gcnii(pgl):
    n, d = 700, 24
    edges, x, g = setup(pgl, n=n, d=d, seed=341)
    conv = pgl.nn.GCNII(d, activation="relu", lambda_l=0.5, alpha=0.2, k_hop=4, dropout=0.0).cuda()
    p = randomize(conv, 342)
    out = run(conv, g, dev(x))
    ws = [p["mlps.%d.weight" % i] for i in range(4)]
    bs = [p["mlps.%d.bias" % i] for i in range(4)]
    want = OC.gcnii(edges, n, x, ws, bs, activation="relu", lambda_l=0.5, alpha=0.2)
    assert rel_err(out, want) <= RTOL


def _transformer_params(p, skip, gate, ln):
    q = {k: (p[k + ".weight"], p[k + ".bias"]) for k in ("q", "k", "v")}
    if skip:
        q["skip"] = (p["skip_feat.weight"], p["skip_feat.bias"])
    if gate:
        q["gate"] = (p["gate.weight"], p["gate.bias"])
    if ln:
        q["ln"] = (p["layer_norm.weight"], p["layer_norm.bias"])
    return q


def test_transformer_conv(pgl):
    n, d, H, Dh = 600, 20, 4, 8
    edges, x, g = setup(pgl, n=n, d=d, seed=351)
    ef = np.random.default_rng(353).standard_normal((edges.shape[0], H * Dh)).astype(np.float32)
    cases = [(True, True, False, True, None), (True, True, True, True, None),
             (False, True, True, False, None), (True, False, False, True, None),
             (True, True, False, True, ef), (False, True, True, True, ef)]
    for concat, skip, gate, ln, efeat in cases:
        conv = pgl.nn.TransformerConv(d, Dh, num_heads=H, feat_drop=0, attn_drop=0, concat=concat,
                                      skip_feat=skip, gate=gate and skip, layer_norm=ln).cuda()
        p = randomize(conv, 352)
        args = (g, dev(x)) if efeat is None else (g, dev(x), dev(efeat))
        out = run(conv, *args)
        want = OC.transformer_conv(edges, n, x, _transformer_params(p, skip, gate and skip, ln), H,
                                   Dh, concat=concat, edge_feat=efeat)
        assert out.shape == want.shape
        assert rel_err(out, want) <= RTOL, (concat, skip, gate, ln, efeat is not None)


def test_gin_conv(pgl):
    n, d, od = 700, 24, 16
    edges, x, g = setup(pgl, n=n, d=d, seed=361)
    for train_eps in (False, True):
        conv = pgl.nn.GINConv(d, od, activation="relu", init_eps=0.25, train_eps=train_eps).cuda()
        p = randomize(conv, 362)
        eps = float(p["epsilon"].reshape(-1)[0]) if train_eps else 0.25
        out = run(conv, g, dev(x))
        want = OC.gin_conv(edges, n, x, p["linear1.weight"], p["linear1.bias"], p["linear2.weight"],
                           p["linear2.bias"], p["layer_norm.weight"], p["layer_norm.bias"],
                           epsilon=eps, activation="relu")
        assert rel_err(out, want) <= RTOL


def test_rgcn_conv(pgl):
    n, d, od = 500, 16, 12
    etypes = ["cites", "writes", "likes"]
    rng = np.random.default_rng(371)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ebt = [(t, O.chung_lu_edges(n, 4000 + 500 * i, exponent=0.7, seed=372 + i))
           for i, t in enumerate(etypes)]
    graphs = {t: make_graph(pgl, e, n) for t, e in ebt}
    for num_bases in (0, 2):
        conv = pgl.nn.RGCNConv(d, od, etypes, num_bases=num_bases).cuda()
        p = randomize(conv, 375)
        out = run(conv, graphs, dev(x))
        want = OC.rgcn_conv(ebt, n, x, p["weight"], p.get("w_comp"))
        assert out.shape == (n, od) and rel_err(out, want) <= RTOL, num_bases


def test_sgc_ssgc_conv(pgl):
    n, d, od = 700, 24, 10
    edges, x, g = setup(pgl, n=n, d=d, seed=381)
    conv = pgl.nn.SGCConv(d, od, k_hop=3, cached=True, activation="relu", bias=True).cuda()
    p = randomize(conv, 382)
    out = run(conv, g, dev(x))
    want = OC.sgc_conv(edges, n, x, p["linear.weight"], k_hop=3, bias=p["bias"], activation="relu")
    assert rel_err(out, want) <= RTOL
    # cached=True: the smoothed features are reused, a different input no longer matters
    with torch.no_grad():
        again = conv(g, dev(x) * 0).cpu().numpy()
    assert rel_err(again, want) <= RTOL
    x_in = dev(x)
    conv = pgl.nn.SSGCConv(d, od, k_hop=5, alpha=0.05, cached=False, bias=False).cuda()
    p = randomize(conv, 383)
    out = run(conv, g, x_in)
    want = OC.ssgc_conv(edges, n, x, p["linear.weight"], k_hop=5, alpha=0.05)
    assert rel_err(out, want) <= RTOL
    assert (x_in.cpu().numpy() == x).all()  # the caller's tensor is never updated in place


def test_ngcf_conv(pgl):
    n, d, od = 700, 24, 16
    edges, x, g = setup(pgl, n=n, d=d, seed=391)
    conv = pgl.nn.NGCFConv(d, od).cuda()
    p = randomize(conv, 392)
    out = run(conv, g, dev(x))
    want = OC.ngcf_conv(edges, n, x, p["linear.weight"], p["linear.bias"], p["linear2.weight"],
                        p["linear2.bias"])
    assert rel_err(out, want) <= RTOL


def test_fa_conv(pgl):
    n, d = 700, 24
    edges, x, g = setup(pgl, n=n, d=d, seed=401)
    conv = pgl.nn.FAConv(d, drop=0.0).cuda()
    p = randomize(conv, 402)
    out = run(conv, g, dev(x))
    want = OC.fa_conv(edges, n, x, p["gate.weight"], p["gate.bias"])
    assert out.shape == (n, d) and rel_err(out, want) <= RTOL


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_reference_conv_list_in_both_dtypes(pgl, dtype):
    """reference tests/test_conv.py:43-71: the same conv list on the 5-node graph, under default dtype float32 and
    float64.  float64 features go through the fp32 kernels (ops.f64_through) and come back as float64; the values
    must agree with the float32 run to the path's tolerance."""
    td = getattr(torch, dtype)
    old = torch.get_default_dtype()
    rng = np.random.default_rng(5)
    D = 8
    edges = [(0, 1), (1, 2), (3, 4)]
    nfeat = rng.standard_normal((5, D))
    outs = {}
    try:
        for which in (torch.float32, td):
            torch.set_default_dtype(which)
            torch.manual_seed(0)
            prev = outs.get("convs")
            convs = [pgl.nn.GCNConv(input_size=D, output_size=D), pgl.nn.GraphSageConv(input_size=D, hidden_size=D),
                     pgl.nn.GATConv(input_size=D, hidden_size=D), pgl.nn.GCNII(hidden_size=D), pgl.nn.APPNP(),
                     pgl.nn.SGCConv(input_size=D, output_size=D), pgl.nn.SSGCConv(input_size=D, output_size=D)]
            g = pgl.Graph(edges=edges, num_nodes=5, node_feat={"nfeat": nfeat.astype(str(which).split(".")[1])})
            pg = g.tensor()
            feat = pg.node_feat["nfeat"]
            assert feat.dtype == which
            res = []
            if prev is not None:   # same parameters as the float32 run (random init differs per dtype)
                for c_new, c_old in zip(convs, prev):
                    c_new.load_state_dict({k: v.to(which) for k, v in c_old.state_dict().items()})
            outs["convs"] = convs
            for conv in convs:
                conv = conv.to("cuda").eval()
                with torch.no_grad():
                    out = conv(pg, feat)
                assert isinstance(out, torch.Tensor) and out.dtype == which, (type(conv).__name__, out.dtype)
                res.append(out.double().cpu().numpy())
            outs[which] = res
    finally:
        torch.set_default_dtype(old)
    for a, b in zip(outs[torch.float32], outs[td]):
        assert np.abs(a - b).max() <= 1e-4 * max(1.0, np.abs(a).max())
