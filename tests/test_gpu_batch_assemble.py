"""GPU: gmeta_packed_set_finish (csrc/batch_assemble.cu) against numpy restatements of what the host packer does --
CSR by source (packed.csr_transpose), tile tables (learner.tile_table), active rows per layer and their task
pointers / tiles / centre positions (packing.active_rows + searchsorted).  Integer work: bit-exact.  Covers empty
tasks, rows without edges, by-source lists of every kind the sorter distinguishes (<= 32 entries: warp ranking,
<= 4096: shared-memory bitonic sort, longer: the quadratic path) and duplicate edges."""
import ctypes as C

import numpy as np
import pytest
import torch

from gmeta_b200 import _lib
from gmeta_b200.learner import tile_table
from gmeta_b200.packed import csr_transpose
from tests import gpu_util as G

pytestmark = pytest.mark.gpu


def _rows_concat(indptr, indices, rows):
    return np.concatenate([indices[indptr[r]:indptr[r + 1]] for r in rows]) if len(rows) else np.zeros(0, np.int64)


def _finish(indptr, indices, sub_node_ptr, task_sub_ptr, centre_row, L):
    lib = _lib.lib()
    N, E, T = indptr.shape[0] - 1, indices.shape[0], task_sub_ptr.shape[0] - 1
    cap_t = (N + 127) // 128 + T
    dev = G.dev()
    z = lambda n: torch.full((max(n, 1),), -7, dtype=torch.int32, device=dev)     # noqa: E731
    out = {"t_indptr": z(N + 1), "t_indices": z(E), "task_row_ptr": z(T + 1), "tile_row0": z(cap_t), "tile_nrows": z(cap_t),
           "tile_task": z(cap_t), "centre_pos": z(centre_row.shape[0]), "counts": z(2 + 2 * L)}
    per = {k: [z(N if k == "act_rows" else (T + 1 if k == "act_task_ptr" else cap_t)) for _ in range(L)]
           for k in ("act_rows", "act_task_ptr", "act_tile_row0", "act_tile_nrows", "act_tile_task")}
    arr = lambda k: (C.c_void_p * max(L, 1))(*[t.data_ptr() for t in per[k]])     # noqa: E731
    nb = lib.gmeta_packed_set_finish_workspace_bytes(N, E, L)
    ws = torch.empty(nb + 256, dtype=torch.uint8, device=dev)
    wp = (ws.data_ptr() + 255) // 256 * 256
    d = [G.i32(x) for x in (indptr, indices if E else np.zeros(1), sub_node_ptr, task_sub_ptr, centre_row)]
    _lib.check(lib.gmeta_packed_set_finish(G.p(d[0]), G.p(d[1]), N, E, G.p(d[2]), G.p(d[3]), T, G.p(d[4]), centre_row.shape[0], L,
                                           G.p(out["t_indptr"]), G.p(out["t_indices"]), G.p(out["task_row_ptr"]),
                                           G.p(out["tile_row0"]), G.p(out["tile_nrows"]), G.p(out["tile_task"]),
                                           arr("act_rows"), arr("act_task_ptr"), arr("act_tile_row0"), arr("act_tile_nrows"),
                                           arr("act_tile_task"), G.p(out["centre_pos"]), G.p(out["counts"]), wp, nb, G.stream()))
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}, {k: [t.cpu().numpy() for t in v] for k, v in per.items()}


def _check(indptr, indices, sub_node_ptr, task_sub_ptr, centre_row, L):
    N, E, T = indptr.shape[0] - 1, indices.shape[0], task_sub_ptr.shape[0] - 1
    got, per = _finish(indptr, indices, sub_node_ptr, task_sub_ptr, centre_row, L)
    t_indptr, t_indices = csr_transpose(indptr, indices, N)
    assert np.array_equal(got["t_indptr"][:N + 1], t_indptr)
    assert np.array_equal(got["t_indices"][:E], t_indices)
    trp = sub_node_ptr[task_sub_ptr].astype(np.int64)
    assert np.array_equal(got["task_row_ptr"][:T + 1], trp)
    row0, nrows, task = tile_table(trp)
    nt = row0.shape[0]
    assert got["counts"][0] == nt and got["counts"][1] == int(np.diff(trp).max())
    for k, want in (("tile_row0", row0), ("tile_nrows", nrows), ("tile_task", task)):
        assert np.array_equal(got[k][:nt], want), k
    rows = np.unique(centre_row.astype(np.int64))
    layers = [None] * L
    if L:
        layers[L - 1] = rows
    for l in range(L - 1, 0, -1):
        rows = np.unique(_rows_concat(indptr, indices.astype(np.int64), rows))
        layers[l - 1] = rows
    for l in range(L):
        n = layers[l].shape[0]
        assert got["counts"][2 + l] == n
        assert np.array_equal(per["act_rows"][l][:n], layers[l])
        tptr = np.searchsorted(layers[l], trp)
        assert np.array_equal(per["act_task_ptr"][l][:T + 1], tptr)
        a0, a1, a2 = tile_table(tptr.astype(np.int64))
        assert got["counts"][2 + L + l] == a0.shape[0]
        for k, want in (("act_tile_row0", a0), ("act_tile_nrows", a1), ("act_tile_task", a2)):
            assert np.array_equal(per[k][l][:a0.shape[0]], want), (k, l)
    if L:
        assert np.array_equal(got["centre_pos"], np.searchsorted(layers[L - 1], centre_row))


def _random_set(rng, sub_sizes, task_sub_ptr, deg, hubs=()):
    """Block-diagonal random CSR by destination over subgraphs of the given sizes; `hubs`: (subgraph, out_degree)
    pairs that make the first node of a subgraph the source of that many (possibly duplicate) edges."""
    sub_node_ptr = np.concatenate([[0], np.cumsum(sub_sizes)]).astype(np.int64)
    src, dst = [], []
    for k, n in enumerate(sub_sizes):
        if n == 0:
            continue
        e = int(n * deg)
        src.append(sub_node_ptr[k] + rng.integers(0, n, e))
        dst.append(sub_node_ptr[k] + rng.integers(0, n, e))
    for k, d in hubs:
        n = sub_sizes[k]
        src.append(np.full(d, sub_node_ptr[k]))
        dst.append(sub_node_ptr[k] + rng.integers(0, n, d))
    src, dst = np.concatenate(src), np.concatenate(dst)
    N = int(sub_node_ptr[-1])
    order = np.argsort(dst, kind="stable")
    indptr = np.zeros(N + 1, dtype=np.int32)
    np.cumsum(np.bincount(dst, minlength=N), out=indptr[1:])
    indices = src[order].astype(np.int32)
    centre = np.array([sub_node_ptr[k] + rng.integers(0, n) for k, n in enumerate(sub_sizes) if n > 0])
    return indptr, indices, sub_node_ptr.astype(np.int32), np.asarray(task_sub_ptr, dtype=np.int32), centre.astype(np.int32)


@pytest.mark.parametrize("L", [1, 2, 3])
def test_finish_matches_the_host_packer_rules(L):
    rng = np.random.default_rng(5 + L)
    sizes = [40, 300, 1, 129, 128, 77, 500, 256, 3]
    _check(*_random_set(rng, sizes, [0, 2, 2, 5, 9], 2.5, hubs=[(1, 31), (3, 33), (6, 450)]), L)      # task 1 is empty


def test_finish_long_by_source_lists():
    rng = np.random.default_rng(11)
    sizes = [6000, 50, 9000]
    # out-degrees 4096 (largest bitonic sort), 5000 and 20000 (quadratic path, duplicates guaranteed)
    _check(*_random_set(rng, sizes, [0, 1, 3], 1.0, hubs=[(0, 4096 - 1), (2, 5000), (2, 20000)]), 2)


def test_finish_no_edges_and_single_task():
    indptr = np.zeros(301, dtype=np.int32)
    _check(indptr, np.zeros(0, dtype=np.int32), np.array([0, 100, 300], dtype=np.int32), np.array([0, 2], dtype=np.int32),
           np.array([5, 250], dtype=np.int32), 2)


def test_finish_rejects_bad_arguments():
    lib = _lib.lib()
    assert lib.gmeta_packed_set_finish_workspace_bytes(-1, 0, 1) < 0
    rc = lib.gmeta_packed_set_finish(None, None, 10, 0, None, None, 1, None, 1, 1, None, None, None, None, None, None,
                                     None, None, None, None, None, None, None, None, 0, None)
    assert rc != 0
