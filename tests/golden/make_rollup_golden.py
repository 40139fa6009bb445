"""Generate tests/golden/rollup_cases.json with the UNMODIFIED reference class
(llm_gateway_core/db/tokens_usage_db.py TokensUsageDB.get_aggregated_usage) -- dev container only.

    python tests/golden/make_rollup_golden.py
"""
import json
import logging
import pathlib
import random
import sqlite3
import sys
import tempfile
from datetime import datetime, timedelta

sys.path.insert(0, "/root/reference")
logging.disable(logging.CRITICAL)
from llm_gateway_core.db.tokens_usage_db import TokensUsageDB  # noqa: E402

rng = random.Random(5)
t0 = datetime(2026, 9, 21, 6, 57, 17, 47518)
rows = []
for i in range(3000):
    ts = (t0 - timedelta(seconds=rng.randrange(0, 400 * 86400), microseconds=rng.randrange(10**6))).isoformat()
    m = None if rng.random() < 0.02 else ["gpt-4.1", "claude", "deepseek/v3", "modèle-é", "z"][rng.randrange(5)]
    rows.append([ts, rng.randrange(2**17), rng.randrange(2**17), rng.randrange(2**17), rng.randrange(2**17), rng.randrange(2**17),
                 rng.randrange(10**6) * 1e-6, m, "P"])
db = TokensUsageDB.__new__(TokensUsageDB)          # the constructor hard-codes <reference>/db (read-only), :17-25
db.db_path = pathlib.Path(tempfile.mkdtemp()) / "t.db"
db._init_db()
c = sqlite3.connect(db.db_path)
c.executemany("INSERT INTO tokens_usage (timestamp, prompt_tokens, completion_tokens, total_tokens, reasoning_tokens, cached_tokens,"
              " cost, model, provider) VALUES (?,?,?,?,?,?,?,?,?)", rows)
c.commit(); c.close()
qs = []
for period in ("hour", "day", "week", "month", "bogus"):
    for (s, e) in ((None, None), (t0 - timedelta(days=14), t0), (t0 - timedelta(days=100), None)):
        qs.append({"period": period, "start": s.isoformat() if s else None, "end": e.isoformat() if e else None,
                   "result": db.get_aggregated_usage(period, s, e)})
out = pathlib.Path(__file__).resolve().parent / "rollup_cases.json"
json.dump({"generator": "tests/golden/make_rollup_golden.py", "sqlite": sqlite3.sqlite_version, "rows": rows, "queries": qs}, open(out, "w"))
print("wrote", out, sum(len(q["result"]) for q in qs), "result rows")
