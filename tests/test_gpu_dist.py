"""Task-sharded meta-step over 2 GPUs (NCCL): N ranks == 1 rank.  Skipped on a single-GPU box."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank)})
    import torch.distributed as td
    from gmeta_b200 import dist
    from gmeta_b200.meta import Meta
    from tests import helpers as H
    torch.cuda.set_device(rank)
    dist.init_from_env("nccl")
    ds = H.tiny_dataset('disjoint')
    mb = ds.sample_meta_batch(np.random.default_rng(7), 4)
    torch.manual_seed(222)
    m = Meta(ds.args(), ds.config()).to(torch.device('cuda', rank))
    accs = m(*dist.shard_meta_batch(mb), ds.feats)       # 2 tasks per rank, ONE all-reduce inside
    accs2 = m(*dist.shard_meta_batch(mb), ds.feats)
    torch.save({"accs": accs, "accs2": accs2, "loss": m.last["loss_q"],
                "params": [p.detach().cpu() for p in m.net.parameters()]}, os.path.join(out_dir, "rank%d.pt" % rank))
    td.barrier()
    td.destroy_process_group()


def test_two_ranks_equal_one_rank(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from gmeta_b200.meta import Meta
    from tests import gpu_util as U
    from tests import helpers as H
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt", weights_only=False)
    r1 = torch.load(tmp_path / "rank1.pt", weights_only=False)
    ds = H.tiny_dataset('disjoint')
    mb = ds.sample_meta_batch(np.random.default_rng(7), 4)
    torch.manual_seed(222)
    m = Meta(ds.args(), ds.config()).to(U.dev())
    accs = m(*mb, ds.feats)
    accs2 = m(*mb, ds.feats)
    for r in (r0, r1):
        np.testing.assert_allclose(r["accs"], accs, atol=1e-6)
        np.testing.assert_allclose(r["accs2"], accs2, atol=1e-6)
        ps = list(m.net.parameters())
        for k, (a, b) in enumerate(zip(r["params"], ps)):
            # the head bias cancels in the prototype distances: its gradient is pure rounding noise, which Adam
            # normalises to +-meta_lr per step whatever the summation order -- in the reference too
            tol = 2 * 2 * ds.args().meta_lr if k == len(ps) - 1 else 2e-6
            U.report("params after 2 steps, 2 ranks vs 1 [%d]" % k, a, b.detach().cpu(), tol, 1e-5)
    for a, b in zip(r0["params"], r1["params"]):
        assert torch.equal(a, b)                           # replicated Adam: bit-identical on every rank
