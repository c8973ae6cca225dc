"""CPU: the bench.py output contract on the arm that runs without a GPU (`--impl reference`: the oracle port of the
reference's CPU path on a bounded sample) -- exactly one JSON line on stdout with the agreed keys; and invariants of
the host-side tiling / transposition helpers under random inputs (hypothesis)."""
import json
import os
import subprocess
import sys

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as hst

from gmeta_b200.learner import tile_table
from gmeta_b200.packed import csr_transpose

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


import pytest


@pytest.mark.parametrize("mounted", [True, False])
def test_reference_arm_prints_one_json_line_with_the_contract_keys(mounted):
    """mounted: the unmodified reference through oracle/ref_loader.py (build container only, kind "reference");
    not mounted (the GPU box): the oracle port (kind "port")."""
    from oracle import ref_loader
    if mounted and not ref_loader.available():
        pytest.skip("reference tree not mounted")
    env = dict(os.environ)
    if not mounted:
        env["GMETA_REFERENCE_DIR"] = "/nonexistent"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-tasks", "1", "--scale", "0.03"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[:500]
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "meta-tasks/s" and d["value"] > 0 and d["steps"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == ("reference" if mounted else "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@settings(max_examples=60, deadline=None)
@given(hst.lists(hst.integers(min_value=0, max_value=700), min_size=1, max_size=12))
def test_tile_table_covers_every_row_once_and_never_straddles_a_task(sizes):
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    row0, nrows, task = tile_table(ptr)
    assert np.all(nrows >= 1) and np.all(nrows <= 128)
    covered = np.concatenate([np.arange(r, r + n) for r, n in zip(row0, nrows)]) if len(row0) else np.zeros(0, int)
    assert np.array_equal(covered, np.arange(ptr[-1]))                       # every row exactly once, in order
    for r, n, t in zip(row0, nrows, task):
        assert ptr[t] <= r and r + n <= ptr[t + 1]                           # inside its task
    assert len(row0) == sum((s + 127) // 128 for s in sizes)


@settings(max_examples=60, deadline=None)
@given(hst.integers(min_value=1, max_value=40), hst.integers(min_value=0, max_value=200), hst.integers(min_value=0, max_value=2 ** 31 - 1))
def test_csr_transpose_is_an_involution_on_sorted_rows_and_keeps_multiplicity(n, e, seed):
    rng = np.random.default_rng(seed)
    src, dst = rng.integers(0, n, e), rng.integers(0, n, e)
    order = np.lexsort((src, dst))                                           # CSR by destination, sources ascending per row
    indptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(dst, minlength=n), out=indptr[1:])
    indices = src[order].astype(np.int32)
    tp, ti = csr_transpose(indptr, indices, n)
    assert tp[-1] == e and np.array_equal(np.diff(tp), np.bincount(src, minlength=n))
    for u in range(n):                                                       # destinations ascending per source row
        assert np.all(np.diff(ti[tp[u]:tp[u + 1]]) >= 0)
    bp, bi = csr_transpose(tp, ti, n)
    assert np.array_equal(bp, indptr) and np.array_equal(bi, indices)
