"""Generate tests/golden/cora.npz -- BASELINE config 2 (SURVEY.md 8d "Config 2 (Cora)").

Run in the build container (where /root/reference exists):  python tests/golden/make_cora.py

Input: the dataset files shipped inside the reference tree, pgl/data/cora/{cora.content,cora.cites}
(the public Cora citation corpus; data, not code).  The loading logic is restated from
pgl/dataset.py:195-246 (CoraDataset): paper ids -> row numbers in file order, class names -> ids in
first-seen order, every citation both ways, one self loop per node, duplicates dropped, the fixed
140 / 300 / 1000 train / val / test split.  The reference de-duplicates through a Python ``set`` whose
iteration order is an implementation detail; the fixture stores the edges sorted by (src, dst), which
only fixes the floating-point summation order.

Stored compactly: edges int16 [E, 2]; the 0/1 bag-of-words matrix as the (row, col) coordinates of
its ones; labels int8.  tests/conftest-free loader: ``load_cora`` in oracle/oracle.py.
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/pgl/data/cora"


def main():
    paper_ids, labels, rows, cols = [], [], [], []
    classes = {}
    with open(os.path.join(SRC, "cora.content")) as f:
        for r, line in enumerate(f):
            parts = line.strip().split()
            paper_ids.append(int(parts[0]))
            labels.append(classes.setdefault(parts[-1], len(classes)))
            bits = np.array(parts[1:-1], dtype=np.int8)
            nz = np.nonzero(bits)[0]
            rows.extend([r] * len(nz))
            cols.extend(nz.tolist())
    vid = {p: i for i, p in enumerate(paper_ids)}
    n = len(paper_ids)
    pairs = set()
    with open(os.path.join(SRC, "cora.cites")) as f:
        for line in f:
            u, v = line.split()
            u, v = vid[int(u)], vid[int(v)]
            pairs.add((u, v))
            pairs.add((v, u))
    for i in range(n):
        pairs.add((i, i))
    edges = np.array(sorted(pairs), dtype=np.int16)
    out = os.path.join(HERE, "cora.npz")
    np.savez_compressed(out, num_nodes=np.int32(n), num_words=np.int32(1433), edges=edges,
                        feat_row=np.array(rows, np.int16), feat_col=np.array(cols, np.int16),
                        labels=np.array(labels, np.int8))
    deg = np.bincount(edges[:, 1].astype(np.int64), minlength=n)
    print("nodes", n, "edges", len(edges), "max in-degree", deg.max(), "mean", round(deg.mean(), 2),
          "ones", len(rows), "classes", len(classes), "bytes", os.path.getsize(out))


if __name__ == "__main__":
    main()
