"""GPU parity of the device-side h-hop extraction (csrc/khop.cu) against the host extractor
(gmeta_b200/subgraphs.py, itself pinned to the reference's generate_subgraph[_link_pred] in
tests/test_host_logic.py).  Integer work: bit-exact node sets, centre indices and induced CSR whenever
the closure fits sample_nodes; for capped subgraphs the reference's np.random.choice over a python-set
order cannot be reproduced, so size / membership / uniformity properties are checked instead."""
import numpy as np
import pytest
import torch

from gmeta_b200.subgraphs import DeviceExtractor, ParentGraph, extract_subgraph, extract_subgraph_link_pred

pytestmark = pytest.mark.gpu


def _graphs(rng, sizes, avg_deg, directed=False):
    gs = []
    for n in sizes:
        m = int(n * avg_deg / 2)
        src = rng.integers(0, n, m)
        dst = (src + 1 + rng.integers(0, n - 1, m)) % n
        if directed:
            gs.append(ParentGraph.from_edges(src, dst, n))
        else:
            gs.append(ParentGraph.from_edges(np.concatenate([src, dst]), np.concatenate([dst, src]), n))
    return gs


def _unpack(out, r, pairs=False):
    np_, ep = out["node_ptr"].cpu().numpy(), out["edge_ptr"].cpu().numpy()
    a, b = int(np_[r]), int(np_[r + 1])
    indptr = out["indptr"].cpu().numpy()[a:b + 1] - int(ep[r])
    indices = out["indices"].cpu().numpy()[int(ep[r]):int(ep[r + 1])] - a
    parent = out["parent"].cpu().numpy()[a:b]
    c = out["centre_row"].cpu().numpy()
    centre = [int(c[2 * r]) - a, int(c[2 * r + 1]) - a] if pairs else int(c[r]) - a
    return indptr, indices, parent, centre


@pytest.mark.parametrize("h", [1, 2, 3])
def test_node_extraction_bit_exact_when_uncapped(h):
    rng = np.random.default_rng(3 + h)
    gs = _graphs(rng, [700, 1500, 64, 300], 3.0)
    ex = DeviceExtractor(gs)
    gi = rng.integers(0, len(gs), 60)
    ca = np.array([rng.integers(0, gs[g].n) for g in gi])
    out = ex.extract(gi, ca, h=h, sample_nodes=2000)
    for r in range(len(gi)):
        want = extract_subgraph(gs[gi[r]], int(ca[r]), h, 10 ** 9)
        indptr, indices, parent, centre = _unpack(out, r)
        assert np.array_equal(parent, want.parent_nid), (h, r)
        assert np.array_equal(indptr, want.indptr) and np.array_equal(indices, want.indices), (h, r)
        assert centre == want.centre
    assert int(out["closure_size"].cpu()[0]) == extract_subgraph(gs[gi[0]], int(ca[0]), h, 10 ** 9).parent_nid.shape[0]


def test_link_pred_extraction_bit_exact_and_directed_multigraph():
    rng = np.random.default_rng(11)
    gs = _graphs(rng, [400, 900], 2.5, directed=True)      # link_process.py:45-47 stores each edge one way
    # multi-edges are kept with multiplicity
    g0 = gs[0]
    gs[0] = ParentGraph.from_edges(np.concatenate([np.repeat(np.arange(5), 2), [7, 7, 7]]),
                                   np.concatenate([np.tile([9, 10], 5), [9, 9, 3]]), g0.n)
    ex = DeviceExtractor(gs)
    gi = np.array([0, 0, 1, 1, 1, 0])
    ca = np.array([9, 3, 5, 100, 17, 10])
    cb = np.array([10, 9, 6, 3, 17, 7])
    out = ex.extract(gi, ca, cb, sample_nodes=1000)
    for r in range(len(gi)):
        want = extract_subgraph_link_pred(gs[gi[r]], int(ca[r]), int(cb[r]), 10 ** 9)
        indptr, indices, parent, centre = _unpack(out, r, pairs=True)
        assert np.array_equal(parent, want.parent_nid), r
        assert np.array_equal(indptr, want.indptr) and np.array_equal(indices, want.indices), r
        assert centre == list(want.centre)


def test_capped_extraction_properties():
    """|closure| > sample_nodes: exactly sample_nodes sampled nodes (+ the centre if it was not drawn), all from
    the closure, sorted, induced edges exactly those of the parent graph, and every closure node is drawn
    about equally often over many seeds."""
    rng = np.random.default_rng(5)
    n = 3000
    hub = np.zeros(2 * n, dtype=np.int64)                    # a star plus a sparse random graph: big closures
    leaves = np.arange(2 * n) % n
    g = ParentGraph.from_edges(np.concatenate([hub, leaves, rng.integers(0, n, 4000)]),
                               np.concatenate([leaves, hub, rng.integers(0, n, 4000)]), n)
    ex = DeviceExtractor([g])
    cap = 200
    centre = 17
    closure = g.khop_in_closure([centre], 2)
    assert closure.shape[0] > cap
    counts = np.zeros(n)
    trials = 60
    for seed in range(trials):
        out = ex.extract(np.array([0, 0]), np.array([centre, 0]), h=2, sample_nodes=cap, seed=1000 + seed)
        indptr, indices, parent, c = _unpack(out, 0)
        assert int(out["closure_size"].cpu()[0]) == closure.shape[0]
        assert parent.shape[0] in (cap, cap + 1) and np.all(np.diff(parent) > 0)
        assert np.isin(parent, closure).all() and parent[c] == centre
        want_indptr, want_indices = g.induced(parent)
        assert np.array_equal(indptr, want_indptr) and np.array_equal(indices, want_indices)
        counts[parent] += 1
    others = np.setdiff1d(closure, [centre])
    freq = counts[others] / trials
    expect = cap / closure.shape[0]
    assert abs(freq.mean() - expect) < 0.02 and freq.max() < expect + 0.35 and counts[np.setdiff1d(np.arange(n), closure)].sum() == 0


def test_extracted_batch_feeds_the_layer_kernel():
    """The packed CSR the device extractor emits is directly consumable by gmeta_gcn_layer_fwd and gives the
    same layer output as the host-extracted, host-packed batch."""
    from gmeta_b200 import _lib
    from tests import gpu_util as U
    rng = np.random.default_rng(8)
    gs = _graphs(rng, [500, 800], 3.0)
    ex = DeviceExtractor(gs)
    gi = rng.integers(0, 2, 12)
    ca = np.array([rng.integers(0, gs[g].n) for g in gi])
    out = ex.extract(gi, ca, h=2, sample_nodes=1000)
    N, E = out["N"], out["E"]
    # host reference of the same batch
    subs = [extract_subgraph(gs[gi[r]], int(ca[r]), 2, 1000) for r in range(len(gi))]
    src, dst, off = [], [], 0
    for sg in subs:
        k = sg.indptr.shape[0] - 1
        d = np.repeat(np.arange(k), np.diff(sg.indptr))
        src.append(sg.indices + off); dst.append(d + off); off += k
    g = U.DevGraph(np.concatenate(src), np.concatenate(dst), off)
    assert off == N and np.concatenate(src).shape[0] == E
    assert torch.equal(out["indptr"].cpu(), g.indptr.cpu()) and torch.equal(out["indices"][:E].cpu(), g.indices.cpu())
    x = U.f32(rng.standard_normal((N, 64)))
    W, b = U.f32(rng.standard_normal((64, 32)) * 0.1), U.f32(rng.standard_normal(32))
    want = U.layer_fwd(g, x, W, b, 64, 32)
    g.indptr, g.indices = out["indptr"], out["indices"]
    got = U.layer_fwd(g, x, W, b, 64, 32)
    assert torch.equal(got, want)
