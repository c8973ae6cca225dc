"""GPU: a meta-batch assembled entirely on the device (extraction + packed-set layout, device_batch.py) against the
host path (host extractor + packing.pack_meta_batch) on the same centres: every segment of the int32 buffer
bit-exact when no subgraph hits the sampling cap, and the same training step out of both."""
import numpy as np
import pytest
import torch

from gmeta_b200 import _lib, device_batch, packing
from gmeta_b200.meta import Meta
from gmeta_b200.subgraphs import DeviceExtractor
from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ['disjoint', 'shared', 'link', 'deep'])
def test_device_built_batch_equals_host_packed_batch(kind):
    ds = H.tiny_dataset(kind)
    ds.sample_nodes = 2000                       # no cap: sampled node sets depend on the sampler
    mb = ds.sample_meta_batch(np.random.default_rng(12), 3)
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
    L = ds.h
    req_s = device_batch.CentreRequests.from_host_batch(xs, cs, ns, gs, ys)
    req_q = device_batch.CentreRequests.from_host_batch(xq, cq, nq, gq, yq)
    ex = DeviceExtractor(ds.graphs)
    ps_s, ps_q, ints = device_batch.build(ex, req_s, req_q, ds.h, ds.sample_nodes, L)
    got = ints.cpu().numpy()
    goff = np.concatenate([[0], np.cumsum([f.shape[0] for f in ds.feats])])[:-1]
    st = packing.Staging(torch.device("cpu"))
    hs, hq, _ = packing.pack_meta_batch(st, mb, goff, L, _lib.lib())
    want = st.host.numpy()
    for d, h in ((ps_s, hs), (ps_q, hq)):
        assert (d.N, d.E, d.S, d.T, d.n_tiles, d.cps, d.max_rows_per_task) == (h.N, h.E, h.S, h.T, h.n_tiles, h.cps, h.max_rows_per_task)
        assert set(d.sizes) == set(h.sizes)
        for k, n in h.sizes.items():
            assert d.sizes[k] == n, k
            assert np.array_equal(got[d.off[k]:d.off[k] + n], want[h.off[k]:h.off[k] + n]), k
        for l in range(L):
            assert d.act[l]["n"] == h.act[l]["n"] and d.act[l]["n_tiles"] == h.act[l]["n_tiles"]


@pytest.mark.parametrize("kind", ['disjoint', 'shared', 'link', 'deep', 'wide'])
@pytest.mark.parametrize("own", [False, True])
def test_slim_pack_equals_full_pack(kind, own):
    """Meta.upload_batch (host packs the CSR by destination only, the rest is derived on the device) holds, segment by
    segment, what the all-host packer writes."""
    ds = H.tiny_dataset(kind)
    mb = ds.sample_meta_batch(np.random.default_rng(21), 3)
    torch.manual_seed(222)
    m = Meta(ds.args(), ds.config()).to('cuda')
    m.device_finish = True
    db = m.upload_batch(mb, ds.feats, own_buffer=own)
    torch.cuda.synchronize()
    got = db.ints.cpu().numpy()
    L = len(m.spec.conv)
    goff = np.concatenate([[0], np.cumsum([f.shape[0] for f in ds.feats])])[:-1]
    st = packing.Staging(torch.device("cpu"))
    hs, hq, _ = packing.pack_meta_batch(st, mb, goff, L, _lib.lib())
    want = st.host.numpy()
    assert db.h2d_bytes < 4 * (hs.N + hs.E + hq.N + hq.E) * 2          # the by-source half never crosses the bus
    for d, h in ((db.ps_s, hs), (db.ps_q, hq)):
        assert (d.N, d.E, d.S, d.T, d.n_tiles, d.cps, d.max_rows_per_task) == (h.N, h.E, h.S, h.T, h.n_tiles, h.cps, h.max_rows_per_task)
        for k, n in h.sizes.items():
            assert d.sizes[k] == n, k
            assert np.array_equal(got[d.off[k]:d.off[k] + n], want[h.off[k]:h.off[k] + n]), k
        for l in range(L):
            assert d.act[l]["n"] == h.act[l]["n"] and d.act[l]["n_tiles"] == h.act[l]["n_tiles"]


@pytest.mark.parametrize("kind", ['disjoint', 'link'])
def test_training_step_from_centres_equals_step_from_host_batch(kind):
    ds = H.tiny_dataset(kind)
    ds.sample_nodes = 2000
    mb = ds.sample_meta_batch(np.random.default_rng(13), 3)
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
    req_s = device_batch.CentreRequests.from_host_batch(xs, cs, ns, gs, ys)
    req_q = device_batch.CentreRequests.from_host_batch(xq, cq, nq, gq, yq)
    outs = []
    for mode in ("host", "device"):
        torch.manual_seed(222)
        m = Meta(ds.args(), ds.config()).to('cuda')
        m.return_meta_grad = True
        if mode == "host":
            accs = m(*mb, ds.feats)
        else:
            accs = m.forward_device(ds.graphs, req_s, req_q, ds.feats, ds.h, ds.sample_nodes)
            assert m.last["h2d_bytes"] < 16384                       # centre ids + labels only
        outs.append((accs, m.last["loss_q"], [g.clone() for g in m.last["meta_grad"]],
                     [p.detach().clone() for p in m.net.parameters()]))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert outs[0][1] == outs[1][1]
    for a, b in zip(outs[0][2] + outs[0][3], outs[1][2] + outs[1][3]):
        assert torch.equal(a, b)                                     # identical buffers -> identical arithmetic


def test_capped_subgraphs_run_and_respect_the_cap():
    ds = H.tiny_dataset('wide')                                      # skewed degrees, sample_nodes = 30
    mb = ds.sample_meta_batch(np.random.default_rng(14), 2)
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
    req_s = device_batch.CentreRequests.from_host_batch(xs, cs, ns, gs, ys)
    req_q = device_batch.CentreRequests.from_host_batch(xq, cq, nq, gq, yq)
    torch.manual_seed(222)
    m = Meta(ds.args(), ds.config()).to('cuda')
    db = m.build_batch_on_device(ds.graphs, req_s, req_q, ds.feats, ds.h, ds.sample_nodes)
    ints = db.ints.cpu().numpy()
    for ps in (db.ps_s, db.ps_q):
        trp = ints[ps.off["task_row_ptr"]:ps.off["task_row_ptr"] + ps.T + 1]
        assert trp[-1] == ps.N and ps.N <= ps.S * (ds.sample_nodes + 1)
    accs = m.forward_device(ds.graphs, req_s, req_q, ds.feats, ds.h, ds.sample_nodes)
    assert accs.shape == (ds.update_step + 1,) and np.all(np.isfinite(accs))
