"""GPU: usage rollup kernel vs the SQLite oracle (integers and order exact, cost within 4 ulp)."""
import math
import struct
from datetime import datetime, timedelta

import numpy as np
import pytest

from llmapigateway_b200.usage import UsageTable, synth_usage_columns, _EPOCH
from oracle import rollup_oracle as ro

pytestmark = pytest.mark.gpu
NOW = datetime(2026, 9, 21, 6, 57, 17, 47518)


@pytest.fixture(scope="module")
def engine():
    import llmapigateway_b200 as L
    e = L.Engine(max_streams=64, max_step_chunks=1024, max_step_bytes=1 << 20)
    yield e
    e.close_engine()


def _ulps(a: float, b: float) -> int:
    ia, ib = struct.unpack("<q", struct.pack("<d", a))[0], struct.unpack("<q", struct.pack("<d", b))[0]
    return abs(ia - ib)


def _iso(us: int) -> str:
    return (_EPOCH + timedelta(microseconds=int(us))).isoformat()


def _load(engine, n, seed, extra_ts=()):
    ts, models, tok, cost = synth_usage_columns(n, seed=seed, end=NOW)
    if len(extra_ts):
        ts[:len(extra_ts)] = extra_ts
    t = UsageTable(engine)
    t.load_columns(ts, models, *tok, cost)
    rows = [(_iso(ts[i]), int(tok[0][i]), int(tok[1][i]), int(tok[2][i]), int(tok[3][i]), int(tok[4][i]), float(cost[i]), models[i], "P") for i in range(n)]
    return t, ro.make_db(rows)


def _compare(got, want):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        for k in ("time_period", "model", "prompt_tokens", "completion_tokens", "total_tokens", "reasoning_tokens", "cached_tokens", "count"):
            assert g[k] == w[k], (k, g, w)
        assert _ulps(g["cost"], w["cost"]) <= 4, (g, w)


def test_rollup_matches_sqlite(engine):
    from llmapigateway_b200.usage import to_us
    # include timestamps that hit SQLite's millisecond-rounding quirks (see csrc/rollup.cuh)
    quirks = [to_us(datetime(2025, 12, 31, 23, 59, 59, 999999)), to_us(datetime(2026, 8, 31, 23, 59, 59, 999600)),
              to_us(datetime(2026, 3, 1, 23, 59, 59, 999999)), to_us(datetime(2026, 2, 28, 23, 59, 59, 999999)),
              to_us(datetime(2026, 6, 30, 23, 59, 59, 999500)), to_us(datetime(2026, 6, 30, 23, 59, 59, 999499))]
    table, conn = _load(engine, 120_000, seed=5, extra_ts=quirks)
    for period in ("hour", "day", "week", "month"):
        for (s, e) in ((None, None), (NOW - timedelta(days=14), NOW), (NOW - timedelta(days=100), None), (None, NOW - timedelta(days=350))):
            _compare(table.get_aggregated_usage(period, s, e), ro.aggregated_usage(conn, period, s, e))
    assert table.get_aggregated_usage("fortnight") == []


def test_rollup_insert_usage_path(engine):
    t = UsageTable(engine)
    base = datetime(2026, 9, 20, 12, 0, 0)
    rows = []
    for i in range(500):
        d = {"prompt_tokens": i, "completion_tokens": 2 * i, "total_tokens": 3 * i, "reasoning_tokens": i % 7, "cached_tokens": i % 3, "cost": i * 1e-6}
        if i % 11:
            d["model"] = "m%d" % (i % 4)
        ts = base + timedelta(minutes=7 * i, microseconds=i)
        t.insert_usage(d, timestamp=ts)
        rows.append((ts.isoformat(), d["prompt_tokens"], d["completion_tokens"], d["total_tokens"], d["reasoning_tokens"], d["cached_tokens"], d["cost"], d.get("model"), None))
    conn = ro.make_db(rows)
    for period in ("hour", "day"):
        _compare(t.get_aggregated_usage(period), ro.aggregated_usage(conn, period))


def test_rollup_full_size_properties(engine):
    """BASELINE.json config 5 size (10 M records): conservation laws instead of an oracle run."""
    n = 10_000_000
    ts, models, tok, cost = synth_usage_columns(n, seed=5, end=NOW)
    t = UsageTable(engine)
    t.load_columns(ts, models, *tok, cost)
    for period in ("hour", "day"):
        rows = t.rollup_rows(period)
        assert int(rows["count"].sum()) == n
        for k, name in enumerate(("prompt_tokens", "completion_tokens", "total_tokens", "reasoning_tokens", "cached_tokens")):
            assert int(rows[name].sum()) == int(tok[k].astype(np.int64).sum())
        assert not rows["inexact"].any()
        assert math.isclose(float(rows["cost"].sum()), float(cost.sum()), rel_tol=1e-9)
        order = list(zip((-rows["bucket"]).tolist(), rows["model_rank"].tolist()))
        assert order == sorted(order)                                  # time_period DESC, model ASC
    ms = t.last_ms()
    assert ms["accum"] > 0


def test_privatised_and_global_paths_agree(engine):
    """The shared-memory (block-privatised) accumulate path and the global-reduction path give the same rows, incl. negative
    token counts (64-bit sign extension through the 32-bit shared halves) and costs with more than 53 bits of spread."""
    n = 150_000
    ts, models, tok, cost = synth_usage_columns(n, seed=21, end=NOW)
    rng = np.random.default_rng(5)
    tok[0][::7] = -tok[0][::7]
    tok[1][:] = rng.integers(-2**31, 2**31 - 1, n, dtype=np.int64).astype(np.int32)
    cost[::11] = -cost[::11] * 1e3
    t = UsageTable(engine)
    t.load_columns(ts, models, *tok, cost)
    for period, s, e in (("hour", NOW - timedelta(hours=24), NOW), ("day", NOW - timedelta(weeks=2), NOW), ("week", NOW - timedelta(weeks=15), NOW), ("month", NOW - timedelta(days=365), NOW)):
        engine._lib.lgw_rollup_set_path(engine._h, 0)
        a = t.rollup_rows(period, s, e)
        engine._lib.lgw_rollup_set_path(engine._h, 1)
        b = t.rollup_rows(period, s, e)
        engine._lib.lgw_rollup_set_path(engine._h, 0)
        assert len(a) == len(b) > 0 and a.tobytes() == b.tobytes(), period
        sel = (ts >= int((s - _EPOCH) / timedelta(microseconds=1))) & (ts <= int((e - _EPOCH) / timedelta(microseconds=1)))
        assert int(a["completion_tokens"].sum()) == int(tok[1][sel].astype(np.int64).sum())


def test_two_rank_rollup_merge_on_the_gpu():
    """records partitioned over ranks, per-rank dense tables from the GPU kernels, one all-reduce(sum), rank 0 emits;
    compared with SQLite over all records.  NCCL when the box has two GPUs, gloo over host copies of the device tables else."""
    import os, subprocess, sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(root / "tests" / "rollup_rank_worker.py")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "ROLLUP_MERGE_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
