"""Live differential of the usage-row sink against the UNMODIFIED `TokensUsageDB.insert_usage` (tokens_usage_db.py:119-162; dev
container only).  Random usage dicts -- the shapes get_token_usage can hand on from hostile upstream JSON: None, floats, numeric
strings, bools, integers beyond 64 bits, dicts and lists, missing keys -- are inserted through the reference into its SQLite file and
through `llmapigateway_b200.gateway.SqliteUsageSink` into another; the stored columns must be equal in value AND storage class, row
for row (a value SQLite cannot bind loses that row on both sides).

    python tools/fuzz_usage_sink_live.py --n 5000 --seed 1
"""
from __future__ import annotations

import argparse
import os
import random
import sqlite3
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tests" / "golden"):
    sys.path.insert(0, str(p))

VALUES = [0, 1, 12, -5, 2 ** 31, 2 ** 40, 2 ** 62, 2 ** 63 - 1, 2 ** 63, 2 ** 64, -2 ** 63, -2 ** 63 - 1, 1.5, 2.0, -0.0, 1e-9, 1e300, float("nan"), float("inf"), "7", "0.5", "x", "",
          None, True, False, {"a": 1}, [1], b"bytes"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    if not Path("/root/reference").exists():
        print("needs /root/reference (dev container)"); return 2
    import logging
    import ref_driver
    from llmapigateway_b200.gateway import SqliteUsageSink
    ref_driver.load_reference()
    logging.disable(logging.CRITICAL)
    tdb = ref_driver._loaded["tdb"]
    rng = random.Random(args.seed)
    keys = ["prompt_tokens", "completion_tokens", "total_tokens", "reasoning_tokens", "cached_tokens", "cost", "model", "provider"]
    rows = []
    for _ in range(args.n):
        u = {}
        for k in keys:
            r = rng.random()
            if r < 0.12:
                continue
            if k in ("model", "provider"):
                u[k] = rng.choice(["m/a", "P", "é", "", None, 17, 2.5, {"a": 1}, ["l"], True]) if r < 0.9 else rng.choice(VALUES)
            else:
                u[k] = rng.choice([rng.randrange(10 ** 6), rng.randrange(100), 0.000123]) if r < 0.6 else rng.choice(VALUES)
        rows.append(u)
    ref_db = tdb.TokensUsageDB()
    for u in rows:
        ref_db.insert_usage(u)
    with tempfile.TemporaryDirectory() as d:
        ours = SqliteUsageSink(os.path.join(d, "ours.db"))
        for u in rows:
            ours.insert_usage(u)
        sel = ", ".join(f"{c}, typeof({c})" for c in keys)
        a = [list(r) for r in sqlite3.connect(ref_db.db_path).execute(f"SELECT {sel} FROM tokens_usage ORDER BY id")]
        b = [list(r) for r in ours.conn.execute(f"SELECT {sel} FROM tokens_usage ORDER BY id")]
    a = a[-len(b):] if len(a) > len(b) else a            # (the reference's file may hold rows of earlier runs in this process)
    same = repr(a) == repr(b)
    print(f"{args.n} usage dicts -> {len(a)} rows stored by the reference, {len(b)} by SqliteUsageSink; equal in value and storage class: {same}")
    if not same:
        for i, (x, y) in enumerate(zip(a, b)):
            if repr(x) != repr(y):
                print("first difference at stored row", i, x, y); break
    return 0 if same else 1


if __name__ == "__main__":
    sys.exit(main())
