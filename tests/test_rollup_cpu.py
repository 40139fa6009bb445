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
