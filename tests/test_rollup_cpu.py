"""CPU: rollup oracle vs the committed golden (made with the unmodified reference class), and the
host-callable bucket arithmetic of the kernel (lgw_rollup_bucket_of) vs SQLite's strftime."""
import json
import random
import sqlite3
from datetime import datetime, timedelta
from pathlib import Path

from llmapigateway_b200 import _native
from llmapigateway_b200.usage import PERIODS, period_label, to_us
from oracle import rollup_oracle as ro

GOLD = Path(__file__).resolve().parent / "golden" / "rollup_cases.json"


def test_oracle_matches_reference_golden():
    doc = json.loads(GOLD.read_text())
    rows = [tuple(r) for r in doc["rows"]]
    conn = ro.make_db(rows)
    for case in doc["queries"]:
        s = datetime.fromisoformat(case["start"]) if case["start"] else None
        e = datetime.fromisoformat(case["end"]) if case["end"] else None
        assert ro.aggregated_usage(conn, case["period"], s, e) == case["result"], case["period"]


def test_bucket_arithmetic_matches_sqlite_strftime():
    lib = _native.load()
    rng = random.Random(3)
    conn = sqlite3.connect(":memory:")
    samples = [datetime(1999, 12, 31, 23, 59, 59, 999999), datetime(2000, 1, 1), datetime(2024, 2, 29, 12), datetime(2024, 12, 30),
               datetime(2026, 1, 1), datetime(2026, 1, 4, 23, 59, 59), datetime(2026, 1, 5), datetime(2027, 1, 3), datetime(2027, 1, 4)]
    for base in (datetime(2026, 3, 1, 23, 59, 59), datetime(2026, 3, 1, 10, 59, 59), datetime(2025, 12, 31, 23, 59, 59), datetime(2026, 8, 31, 23, 58, 59),
                 datetime(2026, 8, 31, 23, 59, 59), datetime(2026, 2, 28, 23, 59, 59), datetime(2024, 2, 29, 23, 59, 59), datetime(2026, 3, 8, 23, 59, 59),
                 datetime(2026, 4, 30, 23, 59, 59), datetime(2026, 1, 29, 23, 59, 59), datetime(1999, 12, 31, 23, 59, 59)):
        samples += [base + timedelta(microseconds=u) for u in (999499, 999500, 999501, 999999, 499999, 500000, 0, 1)]
    samples += [datetime(1990, 1, 1) + timedelta(seconds=rng.randrange(0, 60 * 365 * 86400), microseconds=rng.randrange(10**6)) for _ in range(4000)]
    for dt in samples:
        text = dt.isoformat()
        for period, fmt in ro.PERIOD_FORMATS.items():
            want = conn.execute(f"SELECT strftime('{fmt}', ?)", (text,)).fetchone()[0]
            b = lib.lgw_rollup_bucket_of(to_us(dt), PERIODS[period])
            assert period_label(period, b) == want, (text, period)


def test_record_listing_matches_the_reference_sql():
    """/v1/api/usage-records (stats.py:69-87): UsageTable.get_latest_usage_records / get_total_records_count against the
    reference's statement (tokens_usage_db.py:85-103, :211) on the same rows"""
    from datetime import datetime, timedelta
    from llmapigateway_b200.usage import UsageTable
    t = UsageTable(engine=None)
    base = datetime(2026, 9, 20, 12, 0, 0)
    rows = []
    for i in range(200):
        ts = base + timedelta(seconds=(i * 7919) % 5000, microseconds=(i * 104729) % 1000000 if i % 11 else 0)
        u = {"prompt_tokens": i, "completion_tokens": 2 * i, "total_tokens": 3 * i, "reasoning_tokens": i % 5, "cached_tokens": i % 3,
             "cost": i * 1e-6, "model": None if i % 17 == 0 else "m-%d" % (i % 4), "provider": "P%d" % (i % 3)}
        t.insert_usage(u, timestamp=ts)
        rows.append((ts.isoformat(), u["prompt_tokens"], u["completion_tokens"], u["total_tokens"], u["reasoning_tokens"], u["cached_tokens"], u["cost"], u["model"], u["provider"]))
    conn = ro.make_db(rows)
    assert t.get_total_records_count() == conn.execute("SELECT COUNT(*) FROM tokens_usage").fetchone()[0] == 200
    for limit, offset in ((25, 0), (25, 25), (7, 190), (50, 180), (0, 0), (300, 0)):
        cur = conn.execute("SELECT id, timestamp, prompt_tokens, completion_tokens, total_tokens, reasoning_tokens, cached_tokens, cost, model, provider"
                           " FROM tokens_usage ORDER BY timestamp DESC LIMIT ? OFFSET ?", (limit, offset))
        cols = [d[0] for d in cur.description]
        want = [dict(zip(cols, r)) for r in cur.fetchall()]
        assert t.get_latest_usage_records(limit=limit, offset=offset) == want
