"""Property-based checks (hypothesis) of the integer half on the CPU side of the C-ABI: for arbitrary small
edge lists -- empty, single node, duplicates, self loops, gaps in the id range, num_nodes larger than the
largest id -- pglb_build_index_host == the numpy oracle == the loop restatement == the reference's own
compiled graph_kernel.build_index (when oracle/_ref is present, i.e. in the build container), plus the
structural invariants a CSR must satisfy.  Also the reindex oracle against a dictionary restatement."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import build as obuild
from oracle import oracle as O

REF = obuild.load_ref_graph_kernel()


@st.composite
def edge_lists(draw):
    n = draw(st.integers(1, 40))
    e = draw(st.integers(0, 120))
    u = draw(st.lists(st.integers(0, n - 1), min_size=e, max_size=e))
    v = draw(st.lists(st.integers(0, n - 1), min_size=e, max_size=e))
    extra = draw(st.integers(0, 5))
    return n + extra, np.array(u, np.int64), np.array(v, np.int64)


@settings(max_examples=150, deadline=None)
@given(edge_lists())
def test_build_index_host_matches_oracle_and_reference(case):
    from pgl_b200.utils.edge_index import build_index_host
    n, u, v = case
    got = build_index_host(u, v, n)
    want = O.build_index(u, v, n)
    loops = O.build_index_loops(u, v, n)
    for a, b, c in zip(got, want, loops):
        assert (np.asarray(a) == np.asarray(b)).all() and (np.asarray(b) == np.asarray(c)).all()
    if REF is not None:
        ref = REF.build_index(u, v, n)
        for a, b in zip(got, ref):
            assert (np.asarray(a) == np.asarray(b)).all()
    degree, sorted_v, sorted_u, sorted_eid, indptr = got
    assert indptr[0] == 0 and indptr[-1] == len(u) and (np.diff(indptr) == degree).all()
    assert (np.diff(sorted_u) >= 0).all()                                   # keyed by u
    assert (u[sorted_eid] == sorted_u).all() and (v[sorted_eid] == sorted_v).all()
    for r in range(n):                                                      # stable: ascending edge id per row
        assert (np.diff(sorted_eid[indptr[r]:indptr[r + 1]]) > 0).all()
    assert sorted(sorted_eid.tolist()) == list(range(len(u)))


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 12).flatmap(lambda n: st.tuples(
    st.permutations(list(range(30))).map(lambda p: p[:n]),
    st.lists(st.integers(0, 5), min_size=n, max_size=n))), st.randoms(use_true_random=False))
def test_reindex_oracle_first_appearance(xc, rnd):
    x, count = xc
    nb = [rnd.randrange(30) for _ in range(sum(count))]
    src, dst, out = O.reindex_graph(x, nb, count)
    assert out[:len(x)].tolist() == list(x) and len(set(out.tolist())) == len(out)
    assert (out[src] == np.asarray(nb, np.int64)).all()                     # the mapping inverts
    assert dst.tolist() == [i for i, c in enumerate(count) for _ in range(c)]
    seen, order = set(x), []
    for v in nb:
        if v not in seen:
            seen.add(v)
            order.append(v)
    assert out[len(x):].tolist() == order
