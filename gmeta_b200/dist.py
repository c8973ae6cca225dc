"""Task-level data parallelism: tasks of a meta-batch are independent until the
query-loss sum (meta.py:118,155,161), so rank r owns tasks r, r+world, ... and a meta-step needs
exactly ONE collective -- an all-reduce (SUM) of the flat fp32 buffer
[meta-grad (P) | sum loss_q^K (1) | sum acc_q (K+1)] -- after which every rank applies the
identical Adam update (NaN gate evaluated on the reduced loss, so all ranks branch alike).
NCCL over NVLink/NVSwitch on GPUs; the same code runs on gloo for the CPU tests.
"""
import os

import torch
import torch.distributed as td


def is_dist():
    return td.is_available() and td.is_initialized()


def world_size():
    return td.get_world_size() if is_dist() else 1


def rank():
    return td.get_rank() if is_dist() else 0


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's env (RANK / WORLD_SIZE / MASTER_*)."""
    if is_dist() or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    td.init_process_group(backend=backend)


def shard_tasks(task_num, r=None, world=None):
    """Indices of the tasks rank r owns: round-robin, so a 64-task batch on 8 ranks is 8 each."""
    r = rank() if r is None else r
    world = world_size() if world is None else world
    return list(range(r, task_num, world))


def shard_meta_batch(batch, r=None, world=None):
    """Slice every per-task list of a collated meta-batch (the 10 lists of
    subgraph_data_processing.py:414-419) down to this rank's tasks."""
    idx = shard_tasks(len(batch[0]), r, world)
    return tuple([lst[i] for i in idx] for lst in batch)


def global_task_count(local_tasks):
    """task_num of the whole meta-batch (meta.py:161 divides by it).  Ranks hold equal shares
    when task_num % world == 0; otherwise the count is summed across ranks."""
    if not is_dist():
        return local_tasks
    t = torch.tensor([local_tasks], dtype=torch.int64,
                     device='cuda' if td.get_backend() == 'nccl' else 'cpu')
    td.all_reduce(t)
    return int(t.item())


def allreduce_sum_(flat):
    """In-place SUM all-reduce of the flat [grad | loss | accs] buffer (no-op on one rank)."""
    if is_dist():
        td.all_reduce(flat, op=td.ReduceOp.SUM)
    return flat


def gather_rows(rows):
    """Concatenate every rank's [n_r, C] float rows (n_r may differ) on all ranks: evaluation episodes
    are sharded like training tasks and only their accuracy rows are exchanged."""
    import numpy as np
    if not is_dist():
        return rows
    dev = 'cuda' if td.get_backend() == 'nccl' else 'cpu'
    w, r = world_size(), rank()
    cnt = torch.zeros(w, dtype=torch.int64, device=dev)
    cnt[r] = rows.shape[0]
    td.all_reduce(cnt)
    m, c = int(cnt.max().item()), rows.shape[1]
    buf = torch.zeros(w, max(m, 1), c, dtype=torch.float32, device=dev)
    if rows.shape[0]:
        buf[r, :rows.shape[0]] = torch.as_tensor(rows, dtype=torch.float32).to(dev)
    td.all_reduce(buf)
    buf, cnt = buf.cpu().numpy(), cnt.cpu().numpy()
    return np.concatenate([buf[k, :cnt[k]] for k in range(w)], axis=0)
