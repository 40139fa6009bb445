"""Goldens of the usage-row sink and the stats window, from the UNMODIFIED reference (dev container only):
  * llm_gateway_core/db/tokens_usage_db.py:118 TokensUsageDB.insert_usage -- the stored columns (value + SQLite storage class)
    for a list of usage dicts, including the shapes the tap can produce from hostile upstream JSON (null, float, text, > 2^63);
  * llm_gateway_core/api/v1/stats.py:30-58 get_aggregated_stats -- the (start_date, end_date) it asks the DB for per period.
Writes tests/golden/usage_sink_cases.json."""
from __future__ import annotations

import asyncio
import json
import sqlite3
import sys
import types
from datetime import datetime
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import ref_driver                                                     # noqa: E402

ROWS = [
    {"prompt_tokens": 12, "completion_tokens": 34, "total_tokens": 46, "reasoning_tokens": 0, "cached_tokens": 0, "cost": 0.000123, "model": "m/a", "provider": "P"},
    {"prompt_tokens": 1, "completion_tokens": 2, "total_tokens": 3},
    {"prompt_tokens": None, "completion_tokens": None, "total_tokens": None, "reasoning_tokens": None, "cached_tokens": None, "cost": None, "model": None, "provider": None},
    {"prompt_tokens": 2 ** 31, "completion_tokens": 2 ** 40, "total_tokens": 2 ** 62, "cost": 1, "model": "big"},
    {"prompt_tokens": 1.5, "completion_tokens": 2.0, "total_tokens": "7", "cost": "0.5", "model": 17, "provider": 2.5},
    {"prompt_tokens": True, "completion_tokens": False, "total_tokens": 0, "cost": 0.0, "model": "bools"},
    {"prompt_tokens": -5, "completion_tokens": -2 ** 31, "total_tokens": 0, "cost": -1e-9, "model": "neg"},
    {"prompt_tokens": 2 ** 64, "model": "overflowing: sqlite3 raises OverflowError, the reference swallows it, no row"},
    {"prompt_tokens": 5, "model": {"a": 1}},
    {},
]


def main():
    ref_driver.load_reference()
    tdb = ref_driver._loaded["tdb"]
    db = tdb.TokensUsageDB()
    for r in ROWS:
        db.insert_usage(r)
    conn = sqlite3.connect(db.db_path)
    cols = ["prompt_tokens", "completion_tokens", "total_tokens", "reasoning_tokens", "cached_tokens", "cost", "model", "provider"]
    sel = ", ".join(f"{c}, typeof({c})" for c in cols)
    stored = [list(r) for r in conn.execute(f"SELECT {sel} FROM tokens_usage ORDER BY id")]
    schema = [list(r[1:4]) for r in conn.execute("PRAGMA table_info(tokens_usage)")]

    # stats window
    stub = types.ModuleType("llm_gateway_core.api.v1.models")
    from fastapi import APIRouter
    stub.router = APIRouter()
    sys.modules.setdefault("llm_gateway_core.api.v1.models", stub)
    import llm_gateway_core.db.model_rotation_db as mdb
    import tempfile
    tmp = Path(tempfile.mkdtemp(prefix="lgw_rot_")) / "rotation.db"

    def _init(self, db_filename: str = "llmgateway_rotation.db"):
        self.db_path = tmp
        self._init_db()

    mdb.ModelRotationDB.__init__ = _init
    import importlib
    stats = None
    for name in ("llm_gateway_core.api.v1.stats", "llm_gateway_core.api.stats", "llm_gateway_core.api.v1.usage_stats"):
        try:
            stats = importlib.import_module(name)
            break
        except ModuleNotFoundError:
            continue
    if stats is None:
        import glob
        hits = [p for p in glob.glob(str(ref_driver.REF / "llm_gateway_core" / "**" / "stats.py"), recursive=True)]
        rel = Path(hits[0]).relative_to(ref_driver.REF).with_suffix("")
        stats = importlib.import_module(".".join(rel.parts))
    NOW = datetime(2025, 6, 15, 13, 45, 30, 123456)

    class FixedNow(datetime):
        @classmethod
        def now(cls, tz=None):
            return NOW

    stats.datetime = FixedNow
    windows = {}

    class DB:
        def get_aggregated_usage(self, period, start_date=None, end_date=None):
            windows[period] = [start_date.isoformat(), end_date.isoformat()]
            return []

    req = types.SimpleNamespace(app=types.SimpleNamespace(state=types.SimpleNamespace(tokens_usage_db=DB())))
    for period in ("hour", "day", "week", "month"):
        asyncio.run(stats.get_aggregated_stats(req, period))
    doc = dict(generator="tests/golden/make_usage_sink_golden.py", rows=ROWS, stored=stored, schema=schema, now=NOW.isoformat(), windows=windows)
    (HERE / "usage_sink_cases.json").write_text(json.dumps(doc, indent=1) + "\n")
    print(len(stored), "stored rows of", len(ROWS), "| windows", windows)


if __name__ == "__main__":
    main()
