"""BASELINE config 2 on the real Cora graph (fixture tests/golden/cora.npz): the reference's
citation benchmark model -- GCNConv(1433 -> 16, relu) + Dropout(0.5) + GCNConv(16 -> 7)
(examples/citation_benchmark/model.py:36-52, config/gcn.yaml) -- forward parity against the oracle
with shared weights, and the training sanity SURVEY.md section 8d asks for: Adam lr 0.01, weight decay
5e-4, 200 epochs (train.py:66-75) reaches the published test accuracy 0.807 +- 0.010
(citation_benchmark/README.md:16)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def pgl():
    import pgl_b200
    return pgl_b200


@pytest.fixture(scope="module")
def cora():
    return O.load_cora(os.path.join(GOLDEN, "cora.npz"))


class GCN(torch.nn.Module):
    """examples/citation_benchmark/model.py:23-60 with num_layers=1, hidden 16."""

    def __init__(self, pgl, input_size, num_class, hidden_size=16, dropout=0.5):
        super().__init__()
        self.conv1 = pgl.nn.GCNConv(input_size, hidden_size, activation="relu", norm=True)
        self.drop = torch.nn.Dropout(dropout)
        self.conv2 = pgl.nn.GCNConv(hidden_size, num_class)

    def forward(self, graph, feature):
        return self.conv2(graph, self.drop(self.conv1(graph, feature)))


def test_cora_forward_parity(pgl, cora):
    n, edges, x = cora["num_nodes"], cora["edges"], cora["x"]
    g = pgl.Graph(edges=edges, num_nodes=n)
    g.tensor()
    assert int(g.indegree().max()) == 169
    torch.manual_seed(0)
    model = GCN(pgl, 1433, 7).cuda().eval()
    with torch.no_grad():
        model.conv1.bias.normal_(0, 0.1)
        model.conv2.bias.normal_(0, 0.1)
        out = model(g, torch.from_numpy(x).cuda()).cpu().numpy()
    p = {k: v.detach().cpu().numpy() for k, v in model.named_parameters()}
    h = O.gcn_conv(edges, n, x, p["conv1.linear.weight"], p["conv1.bias"], activation="relu")
    want = O.gcn_conv(edges, n, h, p["conv2.linear.weight"], p["conv2.bias"])
    err = np.abs(out.astype(np.float64) - want).max() / np.abs(want).max()
    assert out.shape == (n, 7) and err <= 1e-4


def test_cora_training_reaches_published_accuracy(pgl, cora):
    n = cora["num_nodes"]
    g = pgl.Graph(edges=cora["edges"], num_nodes=n)
    g.tensor()
    x = torch.from_numpy(cora["x"]).cuda()
    y = torch.from_numpy(cora["y"]).cuda()
    idx = {k: torch.from_numpy(cora[k + "_index"]).cuda() for k in ("train", "val", "test")}
    torch.manual_seed(0)
    model = GCN(pgl, 1433, cora["num_classes"]).cuda()
    optim = torch.optim.Adam(model.parameters(), lr=0.01, weight_decay=5e-4)
    loss_fn = torch.nn.CrossEntropyLoss()
    l0 = pgl.ops.launch_count()
    best_val, test_at_best, first_loss = -1.0, 0.0, None
    for epoch in range(200):
        model.train()
        pred = model(g, x)
        loss = loss_fn(pred[idx["train"]], y[idx["train"]])
        optim.zero_grad()
        loss.backward()
        optim.step()
        if first_loss is None:
            first_loss = float(loss.detach())
        model.eval()
        with torch.no_grad():
            pred = model(g, x).argmax(1)
            val = float((pred[idx["val"]] == y[idx["val"]]).float().mean())
            test = float((pred[idx["test"]] == y[idx["test"]]).float().mean())
        if val > best_val:
            best_val, test_at_best = val, test
    assert pgl.ops.launch_count() - l0 >= 200 * 6  # every aggregation, forward and backward, is ours
    assert float(loss.detach()) < 0.5 * first_loss
    # published: 0.807 +- 0.010 over 10 runs; one seeded run must land in a generous band around it
    assert 0.77 <= test_at_best <= 0.85, (best_val, test_at_best)
