"""GPU: the train.py CLI end to end on tiny dataset directories (node Disjoint, Shared link prediction) and
the batched evaluation path against one-by-one `finetunning` (the reference's loop, train.py:115-146)."""
import os

import numpy as np
import pytest
import torch

from gmeta_b200 import data_io
from gmeta_b200.meta import Meta
from tests import helpers as H
from tests.test_subgraphs_dataset import _load_train

pytestmark = pytest.mark.gpu


def _cli_dataset(kind):
    """Enough classes (Disjoint: every split needs >= n_way of them) / graphs for a train/val/test split."""
    from gmeta_b200 import synthetic as S
    rng = np.random.default_rng(9)
    if kind == 'disjoint':
        g = S.er_graph(600, 2400, rng)
        return S._node_dataset('cli_disjoint', [g], 16, 12, rng, task_setup='Disjoint', n_way=3, k_spt=2, k_qry=4, h=2,
                               hidden_dim=16, update_step=3, update_lr=0.05, meta_lr=1e-3, task_num=3, sample_nodes=40,
                               update_step_test=4)
    return S._link_dataset('cli_link', 6, 120, 260, 5, rng, n_way=2, k_spt=4, k_qry=6, h=2, hidden_dim=16,
                           update_step=3, update_lr=0.05, meta_lr=5e-4, task_num=2, sample_nodes=60, update_step_test=4)


@pytest.mark.parametrize("device_extract", ['False', 'True'])
@pytest.mark.parametrize("kind", ['disjoint', 'link'])
def test_train_cli_runs_and_learns_something_finite(tmp_path, kind, device_extract, capsys):
    train = _load_train()
    ds = _cli_dataset(kind)
    root = data_io.write_synthetic_dataset(str(tmp_path / kind), ds, np.random.default_rng(5), frac=(0.5, 0.25, 0.25))
    argv = ["--data_dir", root + "/", "--task_setup", ds.task_setup, "--epoch", "2", "--batchsz", "6", "--task_num", "3",
            "--n_way", str(ds.n_way), "--k_spt", str(ds.k_spt), "--k_qry", str(ds.k_qry), "--hidden_dim", "16",
            "--update_step", "3", "--update_step_test", "4", "--update_lr", "0.05", "--h", str(ds.h),
            "--sample_nodes", str(ds.sample_nodes), "--train_result_report_steps", "1", "--eval_batch", "7"]
    if ds.link_pred:
        argv += ["--link_pred_mode", "True"]
    argv += ["--device_extract", device_extract]
    accs = train.main(train.parse(argv))
    out = capsys.readouterr().out
    for line in ("There are", "Total trainable tensors:", "------ Start Training ------", "Epoch: 1  Step: 0  training acc:",
                 "Epoch: 2  Val acc:", "Test acc:", "Early Stopped Test acc:", "Total Time:", "Max Momory:"):
        assert line in out, line                         # the reference's report lines (train.py:56-148)
    assert accs.shape == (5,) and np.all(np.isfinite(accs)) and np.all((accs >= 0) & (accs <= 1))


@pytest.mark.parametrize("kind", ['disjoint', 'shared', 'link'])
def test_batched_finetuning_equals_one_by_one(kind):
    ds = H.tiny_dataset(kind)
    rng = np.random.default_rng(2)
    eps = ds.sample_meta_batch(rng, 5)
    torch.manual_seed(222)
    m = Meta(ds.args(), ds.config()).to('cuda')
    before = [p.detach().clone() for p in m.net.parameters()]
    got = m.finetunning_batch(*eps, ds.feats)
    assert got.shape == (5, ds.update_step_test + 1)
    for e in range(5):
        one = m.finetunning(*[[lst[e]] for lst in eps], ds.feats)
        assert np.array_equal(got[e], one), (e, got[e], one)
    for a, b in zip(before, m.net.parameters()):         # fine-tuning never touches the meta-parameters
        assert torch.equal(a, b.detach())
    assert m.finetunning_batch(*[[] for _ in range(10)], ds.feats).shape == (0, ds.update_step_test + 1)
