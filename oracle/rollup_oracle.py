"""CPU oracle for the usage-stats rollup -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates llm_gateway_core/db/tokens_usage_db.py of the reference:
  schema            :37-50
  insert_usage      :119-162   (timestamp = datetime.now().isoformat(), :135)
  get_aggregated_usage :222-304 (strftime bucket x model GROUP BY, SUM x5 int, SUM(cost), COUNT(*),
                                 ORDER BY time_period DESC, model ASC; ISO-text window filter :255-266)

The arithmetic lives in SQLite (system library, 3.45.1 in this image): SUM over INTEGER is exact
int64, SUM over REAL is a compensated (Kahan-Babuska) double sum in scan order since 3.43.  The
oracle therefore IS SQLite: it loads the rows into an in-memory database with the reference's
schema and runs the reference's statement text.  Pinned by tests/test_rollup_cpu.py against the
unmodified reference class (TokensUsageDB.get_aggregated_usage) when /root/reference is present,
and by tests/golden/rollup_cases.json (generated from it) elsewhere.
"""
from __future__ import annotations

import sqlite3
from datetime import datetime

PERIOD_FORMATS = {"hour": "%Y-%m-%d %H:00:00", "day": "%Y-%m-%d", "week": "%Y-W%W", "month": "%Y-%m"}   # :242-250
COLUMNS = ("timestamp", "prompt_tokens", "completion_tokens", "total_tokens", "reasoning_tokens",
           "cached_tokens", "cost", "model", "provider")


def make_db(rows) -> sqlite3.Connection:
    """rows: iterable of 9-tuples in COLUMNS order (timestamp as ISO text)."""
    conn = sqlite3.connect(":memory:")
    conn.execute("""CREATE TABLE tokens_usage (id INTEGER PRIMARY KEY AUTOINCREMENT, timestamp DATETIME NOT NULL,
        prompt_tokens INTEGER DEFAULT 0, completion_tokens INTEGER DEFAULT 0, total_tokens INTEGER DEFAULT 0,
        reasoning_tokens INTEGER DEFAULT 0, cached_tokens INTEGER DEFAULT 0, cost REAL DEFAULT 0.0, model TEXT, provider TEXT)""")
    conn.executemany("INSERT INTO tokens_usage (timestamp, prompt_tokens, completion_tokens, total_tokens, reasoning_tokens,"
                     " cached_tokens, cost, model, provider) VALUES (?,?,?,?,?,?,?,?,?)", rows)
    conn.commit()
    return conn


def aggregated_usage(conn: sqlite3.Connection, period: str, start: datetime | None = None, end: datetime | None = None) -> list[dict]:
    fmt = PERIOD_FORMATS.get(period)
    if fmt is None:
        return []                                            # :251-252, :296-298
    where, params = "", []
    if start is not None:
        where += " WHERE timestamp >= ?"; params.append(start.isoformat())
    if end is not None:
        where += (" AND" if where else " WHERE") + " timestamp <= ?"; params.append(end.isoformat())
    sql = (f"SELECT strftime('{fmt}', timestamp) as time_period, model, SUM(prompt_tokens) as prompt_tokens,"
           " SUM(completion_tokens) as completion_tokens, SUM(total_tokens) as total_tokens,"
           " SUM(reasoning_tokens) as reasoning_tokens, SUM(cached_tokens) as cached_tokens, SUM(cost) as cost,"
           f" COUNT(*) as count FROM tokens_usage{where} GROUP BY time_period, model ORDER BY time_period DESC, model ASC")
    cur = conn.execute(sql, params)
    cols = [d[0] for d in cur.description]
    return [dict(zip(cols, r)) for r in cur.fetchall()]
