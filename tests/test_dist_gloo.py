"""N > 1 path on CPU: world_size-2 gloo processes exercising gmeta_b200.dist -- task sharding, the
single all-reduce of the flat [meta-grad | loss | accuracies] buffer, the global task count and
the replicated NaN gate.  The per-task contributions come from the oracle (CPU), so
the test also pins "sum over ranks of per-shard results == the one-process meta-step"."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "WORLD_SIZE": str(world), "LOCAL_RANK": str(rank)})
    import torch.distributed as td
    from gmeta_b200 import dist
    from oracle import gmeta_oracle as O
    from tests import helpers as H
    dist.init_from_env("gloo")
    assert dist.world_size() == world and dist.rank() == rank
    ds = H.tiny_dataset('disjoint')
    mb = ds.sample_meta_batch(np.random.default_rng(7), 4)
    mine = dist.shard_meta_batch(mb)
    assert len(mine[0]) == 2 and dist.global_task_count(len(mine[0])) == 4
    torch.manual_seed(222)
    # per-shard meta-step on the oracle (same initial parameters on every rank): it averages over the
    # shard's tasks, so rescale by the shard size before the sum and by the global count after it
    om = O.OracleMeta(ds.args(), ds.config())
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mine
    accs = om.forward([H.to_ograph(x) for x in xs], ys, [H.to_ograph(x) for x in xq], yq, cs, cq, ns, nq, gs, gq,
                      ds.feats)
    local_T = len(mine[0])
    flat = torch.cat([torch.cat([g.reshape(-1) for g in om.last_grads]) * local_T,
                      torch.tensor([om.last_loss_q * local_T], dtype=torch.float32),
                      torch.as_tensor(accs * local_T, dtype=torch.float32)])
    dist.allreduce_sum_(flat)                     # THE collective of a meta-step
    flat = flat / 4
    torch.save(flat, os.path.join(out_dir, "rank%d.pt" % rank))
    td.barrier()
    td.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path):
    from oracle import gmeta_oracle as O
    from tests import helpers as H
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    assert torch.equal(r0, r1)                     # every rank holds the identical reduced buffer
    ds = H.tiny_dataset('disjoint')
    mb = ds.sample_meta_batch(np.random.default_rng(7), 4)
    torch.manual_seed(222)
    om = O.OracleMeta(ds.args(), ds.config())
    xs, ys, xq, yq, cs, cq, ns, nq, gs, gq = mb
    accs = om.forward([H.to_ograph(x) for x in xs], ys, [H.to_ograph(x) for x in xq], yq, cs, cq, ns, nq, gs, gq,
                      ds.feats)
    want = torch.cat([torch.cat([g.reshape(-1) for g in om.last_grads]),
                      torch.tensor([om.last_loss_q], dtype=torch.float32), torch.as_tensor(accs, dtype=torch.float32)])
    assert torch.allclose(r0, want, rtol=1e-5, atol=1e-7), float((r0 - want).abs().max())


def test_shard_tasks_partitions_every_task_once():
    from gmeta_b200 import dist
    for T in (1, 4, 7, 32, 64):
        for world in (1, 2, 3, 4, 8):
            seen = sorted(t for r in range(world) for t in dist.shard_tasks(T, r, world))
            assert seen == list(range(T))
            sizes = [len(dist.shard_tasks(T, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _gather_worker(rank, world, port, q):
    import numpy as np
    import torch.distributed as td
    from gmeta_b200 import dist
    td.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rows = np.full((rank + 1, 3), float(rank), dtype=np.float32)        # unequal shares: 1 row on rank 0, 2 on rank 1
    out = dist.gather_rows(rows)
    q.put((rank, out.tolist(), dist.shard_tasks(5)))
    td.destroy_process_group()


def test_gather_rows_and_task_sharding_world2():
    """Evaluation episodes are sharded like training tasks and only accuracy rows are exchanged (train.py)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    ps = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
    want = [[0.0] * 3, [1.0] * 3, [1.0] * 3]
    assert got[0][1] == want and got[1][1] == want
    assert got[0][2] == [0, 2, 4] and got[1][2] == [1, 3]
