"""World-size-2 gloo tests (CPU) of the N>1 host logic: partition relabelling, HaloPlan
construction, the halo all-to-all, and DistGPUGraph's edge sharding.  The local aggregation is
done by the ORACLE here (tests only); the CUDA kernels are covered by the -m gpu tests."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, method, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from pgl_b200.distributed.halo import HaloPlan, block_offsets, relabel_by_partition
        n, e, d = 400, 5000, 12
        edges = O.chung_lu_edges(n, e, exponent=0.7, seed=5)
        rng = np.random.default_rng(6)
        x = rng.standard_normal((n, d)).astype(np.float32)
        want = O.send_u_recv(x, edges[:, 0], edges[:, 1], "sum")
        if method == "block":
            offsets = block_offsets(n, world)
            new_id = np.arange(n)
        else:
            part = (np.arange(n) * 7919 % 3 == 0).astype(np.int64)  # an arbitrary 2-way partition
            new_id, offsets = relabel_by_partition(part, world)
            assert offsets[-1] == n and len(offsets) == world + 1
            # parts are contiguous and stable inside a part
            inv = np.argsort(new_id)
            assert (np.diff(part[inv]) >= 0).all()
        e2 = new_id[edges]
        x2 = np.empty_like(x)
        x2[new_id] = x
        plan = HaloPlan.build(torch.from_numpy(e2), n, offsets, rank, world)
        lo, hi = offsets[rank], offsets[rank + 1]
        # every in-edge of an owned node is local, ids are ascending global edge ids
        mine = (e2[:, 1] >= lo) & (e2[:, 1] < hi)
        assert plan.eid.tolist() == np.nonzero(mine)[0].tolist()
        assert plan.n_local == hi - lo
        halo = plan.halo_ids.numpy()
        assert (np.diff(halo) > 0).all() and ((halo < lo) | (halo >= hi)).all()
        assert sum(plan.recv_counts) == plan.n_halo and plan.recv_counts[rank] == 0
        # exchange: test-side packing (plain indexing); the product passes the CUDA gather kernel
        x_local = torch.from_numpy(x2[lo:hi].copy())
        x_ext = plan.exchange(x_local, pack=lambda t, idx: t[idx])
        assert np.array_equal(x_ext[: plan.n_local].numpy(), x2[lo:hi])
        assert np.array_equal(x_ext[plan.n_local:].numpy(), x2[halo])
        # local aggregation (oracle) over [own | halo] == the owned rows of the global result
        got = O.send_u_recv(x_ext.numpy(), plan.col_local.numpy(), plan.dst_local.numpy(), "sum",
                            out_size=plan.n_local)
        want2 = np.empty_like(want)
        want2[new_id] = want
        np.testing.assert_array_equal(got, want2[lo:hi])
        # second exchange with a different width reuses the plan
        v = torch.from_numpy(x2[lo:hi, :1].copy())
        v_ext = plan.exchange(v, pack=lambda t, idx: t[idx])
        assert np.array_equal(v_ext[plan.n_local:, 0].numpy(), x2[halo, 0])

        # reference DistGPUGraph semantics: shard by dst % world, all-reduce of the partial outputs
        sh, eid = O.shard_edges_by_dst(edges, world, rank)
        part_out = torch.from_numpy(O.send_u_recv(x, sh[:, 0], sh[:, 1], "sum", out_size=n))
        from pgl_b200.utils.op import all_reduce_sum_with_grad
        full = all_reduce_sum_with_grad(part_out)
        np.testing.assert_allclose(full.numpy(), want, rtol=1e-6, atol=1e-6)
        # differentiable: grad of all-reduce-sum is an all-reduce-sum of the upstream gradient
        t = part_out.clone().requires_grad_(True)
        all_reduce_sum_with_grad(t).sum().backward()
        assert torch.all(t.grad == world)
        ret[rank] = "ok"
    except Exception as ex:  # pragma: no cover
        import traceback
        ret[rank] = "FAIL: " + traceback.format_exc()
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("method", ["block", "relabel"])
def test_halo_plan_world2_gloo(method):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, method, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret.get(r) == "ok", ret.get(r)


def test_block_offsets_and_relabel():
    from pgl_b200.distributed.halo import block_offsets, relabel_by_partition
    assert block_offsets(10, 4) == [0, 3, 6, 9, 10]
    assert block_offsets(8, 8) == list(range(9))
    part = np.array([1, 0, 1, 0, 2, 2, 0])
    new_id, off = relabel_by_partition(part, 3)
    assert off == [0, 3, 5, 7]
    assert new_id.tolist() == [3, 0, 4, 1, 5, 6, 2]


def _colshard_worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from pgl_b200.distributed import ColumnShardedGraph, block_offsets
        d = 8
        rng = np.random.default_rng(16)
        x = rng.standard_normal((n, d)).astype(np.float32)
        cs = ColumnShardedGraph(None, d, world, rank)
        lo, hi = cs.column_range()
        assert (lo, hi) == (rank * d // world, (rank + 1) * d // world)
        x_cols = cs.slice_columns(torch.from_numpy(x))
        assert np.array_equal(x_cols.numpy(), x[:, lo:hi])
        off = block_offsets(n, world)
        rows = cs.to_rows(x_cols)                     # my row block, every column
        assert np.array_equal(rows.numpy(), x[off[rank]:off[rank + 1]])
        back = cs.to_cols(rows * 2.0, n)              # and back (after a row-wise op)
        assert np.array_equal(back.numpy(), 2.0 * x[:, lo:hi])
        # the aggregation itself needs no exchange: a column slice of the result depends only on the
        # same column slice of the input
        edges = O.chung_lu_edges(n, 6 * n, exponent=0.7, seed=17)
        full = O.send_u_recv(x, edges[:, 0], edges[:, 1], "sum")
        mine = O.send_u_recv(x[:, lo:hi], edges[:, 0], edges[:, 1], "sum")
        np.testing.assert_array_equal(mine, full[:, lo:hi])
        # a whole GCN layer on the column-sharded layout == the oracle's layer, column slice by slice
        w = (rng.standard_normal((d, 6)) * 0.3).astype(np.float32)
        bias = rng.standard_normal(6).astype(np.float32)
        norm = O.degree_norm(O.adj_dst_index(edges, n)[0], np.float32)

        def oracle_agg(xc, nrm):
            xs = xc.numpy() * nrm
            return torch.from_numpy(O.send_u_recv(xs, edges[:, 0], edges[:, 1], "sum") * nrm)

        got = cs.gcn_layer(x_cols, norm, torch.from_numpy(w), torch.from_numpy(bias), torch.relu,
                           aggregate=oracle_agg)
        want = O.gcn_conv(edges, n, x, w, bias, activation="relu")
        olo, ohi = rank * 6 // world, (rank + 1) * 6 // world
        np.testing.assert_allclose(got.numpy(), want[:, olo:ohi], rtol=1e-5, atol=1e-5)
        ret[rank] = "ok"
    except Exception:  # pragma: no cover
        import traceback
        ret[rank] = "FAIL: " + traceback.format_exc()
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [400, 401])
def test_column_shard_reshard_world2_gloo(n):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_colshard_worker, args=(world, port, n, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret.get(r) == "ok", ret.get(r)
