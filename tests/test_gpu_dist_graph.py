"""DistGPUGraph on CUDA with world size 2 (reference pgl/graph.py:1410-1553, tests/test_dist_graph.py:26-137, which
needs a manual multi-process launch in the reference): edges sharded by dst % world, features replicated, every
recv / degree result all-reduce-summed.  Two ranks are spawned here; with two visible GPUs they use NCCL on one
device each, with a single GPU both ranks share cuda:0 and the (differentiable) all-reduce runs over gloo -- the
aggregation kernels are the same either way.  Expected values are the reference's own KATs (tests/golden/kat.json)
plus the oracle on a random power-law graph."""
import json
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ngpu, ret):
    try:
        import sys
        sys.path.insert(0, ROOT)
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dev = torch.device("cuda", rank if ngpu >= world else 0)
        torch.cuda.set_device(dev)
        backend = "nccl" if ngpu >= world else "gloo"
        dist.init_process_group(backend, rank=rank, world_size=world)
        import warnings
        import pgl_b200 as pgl
        from oracle import oracle as O
        kat = json.load(open(os.path.join(ROOT, "tests", "golden", "kat.json")))

        def dgraph(edges, n, nfeat=None, efeat=None):
            g = pgl.Graph(edges=np.asarray(edges, np.int64), num_nodes=n,
                          node_feat=None if nfeat is None else {"nfeat": np.asarray(nfeat, np.float32)},
                          edge_feat=None if efeat is None else {"efeat": np.asarray(efeat, np.float32)})
            g.tensor()
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                return pgl.graph.DistGPUGraph(g)

        # degree (tests/test_dist_graph.py:26-51)
        k = kat["send_recv_sum"]
        g = dgraph(k["edges"], k["num_nodes"], k["nfeat"])
        indeg = np.bincount(np.asarray(k["edges"])[:, 1], minlength=k["num_nodes"])
        outdeg = np.bincount(np.asarray(k["edges"])[:, 0], minlength=k["num_nodes"])
        assert g.indegree().cpu().numpy().tolist() == indeg.tolist()
        assert g.outdegree().cpu().numpy().tolist() == outdeg.tolist()
        # every rank holds only its dst shard
        assert int(g.num_edges) == int((np.asarray(k["edges"])[:, 1] % world == rank).sum())
        # send_recv (:53-76) and send -> recv (:78-113): exact
        out = g.send_recv(g.node_feat["nfeat"], "sum")
        assert out.cpu().numpy().tolist() == k["ground"]
        msg = g.send(lambda s, d, e: {"h": s["h"]}, src_feat={"h": g.node_feat["nfeat"]})
        out2 = g.recv(lambda m: m.reduce_sum(m["h"]), msg)
        assert out2.cpu().numpy().tolist() == k["ground"]
        # send_ue_recv add / sum with an [E] edge feature (:115-137)
        k2 = kat["send_ue_recv_add_sum"]
        g2 = dgraph(k2["edges"], k2["num_nodes"], k2["nfeat"], k2["efeat"])
        out3 = g2.send_ue_recv(g2.node_feat["nfeat"], g2.edge_feat["efeat"], "add", "sum")
        assert out3.cpu().numpy().tolist() == k2["ground"]

        # a real graph: every reduce op through the sharded path == the oracle on the whole graph, and the
        # all-reduce is differentiable (gradient of the sum reaches every rank's replica)
        n, e, d = 3000, 40000, 32
        edges = O.chung_lu_edges(n, e, exponent=0.9, seed=77)
        x = np.random.default_rng(78).standard_normal((n, d)).astype(np.float32)
        g3 = dgraph(edges, n)
        xd = torch.from_numpy(x).to(dev)
        for op in ("sum", "mean", "max", "min"):
            got = g3.send_recv(xd, op).cpu().numpy()
            want = O.send_u_recv(x, edges[:, 0], edges[:, 1], op)
            err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-12)
            assert err <= 1e-4, (op, err)
        xg = xd.clone().requires_grad_(True)
        g3.send_recv(xg, "sum").sum().backward()
        # d/dx of sum over all outputs = out-degree of the node, summed over the ranks' shards by the all-reduce's
        # own backward (an all-reduce of the upstream gradient: every rank sees `world` x its shard's share)
        sh_out = np.bincount(edges[edges[:, 1] % world == rank, 0], minlength=n).astype(np.float32)
        assert np.allclose(xg.grad[:, 0].cpu().numpy(), world * sh_out)
        ret[rank] = "ok"
    except Exception:  # pragma: no cover
        import traceback
        ret[rank] = "FAIL: " + traceback.format_exc()
    finally:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass


def test_dist_gpu_graph_world2_cuda():
    world = 2
    ngpu = torch.cuda.device_count()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ngpu, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret.get(r) == "ok", ret.get(r)
