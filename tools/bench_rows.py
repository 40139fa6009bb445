"""Measurements for the SURVEY 8 rows that are not bench.py's headline workload:
  C2  request-body rewrite  (BASELINE.json configs[1]: 1024 non-streaming requests, 4 KiB JSON bodies)
  C5  usage rollup          (configs[4]: 10M usage records -> per-hour/day/model table)
One JSON line each (kernel time from CUDA events inside the library, e2e through the host-buffer C-ABI call,
CPU = the oracle on a bounded sample).   python tools/bench_rows.py [--records 10000000]
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time
from datetime import datetime, timedelta
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))

RULES = {"gw/chain": {"rotate_models": False, "fallback_models": [
    {"provider": "openrouter", "model": "vendor/model-large", "use_provider_order_as_fallback": False, "providers_order": ["A", "B"],
     "custom_body_params": {"reasoning_effort": "high", "temperature": 0.5}, "custom_headers": {}}]}}


def _cpu_bodies(args):
    seed, n, mode = args
    from llmapigateway_b200.synth import chat_bodies
    from oracle import body_oracle as bo
    base = chat_bodies(32, 4096, seed=seed)
    ops = bo.rule_ops(RULES["gw/chain"]["fallback_models"][0], "openrouter")
    for raw in base[:8]:
        bo.rewrite(raw, ops, mode)                                 # warm the interpreter's caches
    t0 = time.perf_counter()
    for i in range(n):
        st, out = bo.rewrite(base[i % 32], ops, mode)
        assert st == 0
    return time.perf_counter() - t0


def bench_bodies(eng, n=1024, reps=20):
    from llmapigateway_b200 import rewrite as rw
    from llmapigateway_b200.synth import chat_bodies
    from oracle import body_oracle as bo
    plans = rw.RulePlans(RULES, fallback_provider="fb")
    eng.load_rules(plans)
    base = chat_bodies(256, 4096, seed=2)
    bodies = [base[i % len(base)] for i in range(n)]
    buf, off = rw.pack_bodies(bodies)
    out_lines = []
    for stream, mode_name in ((True, "httpx028"), (False, "json5")):
        idx = np.full(n, plans.plan_index("gw/chain", 0, stream=stream), dtype=np.uint32)
        out = np.empty(int(off[-1]) + n * 512, dtype=np.uint8)
        kern, host = [], []
        for it in range(reps + 3):
            t0 = time.perf_counter()
            o, o_off, res = eng.rewrite_packed(buf, off, idx, 8192, out=out)
            t1 = time.perf_counter()
            if it >= 3:
                ms = eng.bodies_last_ms(); kern.append(ms["rewrite"] + ms["offsets"] + ms["pack"]); host.append((t1 - t0) * 1e3)
        assert int((res["status"] == 0).sum()) == n
        ops = bo.rule_ops(RULES["gw/chain"]["fallback_models"][0], "openrouter")
        for i in (0, 77, n - 1):                                   # spot parity inside the bench
            assert bytes(o[int(o_off[i]):int(o_off[i + 1])]) == bo.rewrite(bodies[i], ops, mode_name)[1]
        procs = os.cpu_count() or 1
        per = 4000 if mode_name == "httpx028" else 1200
        with mp.get_context("spawn").Pool(procs) as pool:
            times = pool.map(_cpu_bodies, [(100 + i, per, mode_name) for i in range(procs)])
        k, h = float(np.median(kern)), float(np.median(host))
        in_b, out_b = int(off[-1]), int(o_off[-1])
        out_lines.append({
            "metric": "request bodies rewritten per second", "unit": "bodies/s", "value": n / k * 1e3, "ms_per_step": k,
            "config": {"workload": f"C2: {n} request bodies x 4 KiB, one upstream attempt (model + usage + 2 body params + provider routing), render {mode_name}"},
            "dtype": "u8", "data": "synthetic",
            "e2e": {"value": n / h * 1e3, "unit": "bodies/s", "ms_per_step": h, "h2d_bytes_per_step": in_b + 8 * (n + 1) + 4 * n,
                    "d2h_bytes_per_step": out_b + 8 * (n + 1) + 16 * n, "note": "pageable numpy buffers"},
            "roofline": {"bound": "hbm", "achieved": (in_b + out_b) / k / 1e6, "unit": "GB/s", "algorithmic_bytes": in_b + out_b,
                         "note": "in + out once; the slot write and its re-read by the pack pass are not counted"},
            "cpu_baseline": {"value": procs * per / max(times), "unit": "bodies/s", "cores": procs, "kind": "port",
                             "sample": f"{procs} procs x {per} bodies (oracle/body_oracle.py: stdlib json.loads + deepcopy + dumps"
                                       + ("; json5.dumps restated in Python)" if mode_name == "json5" else ")")},
            "gpu_launches": 4})
    return out_lines


def _iso(us):
    return (datetime(1970, 1, 1) + timedelta(microseconds=int(us))).isoformat()


def bench_rollup(eng, n):
    from llmapigateway_b200.usage import UsageTable, synth_usage_columns
    from oracle import rollup_oracle as ro
    ts, models, tok, cost = synth_usage_columns(n, seed=5)
    t = UsageTable(eng)
    t.load_columns(ts, models, *tok, cost)
    lines = []
    m = min(n, 500_000)                                             # SQLite on a bounded sample of the same records
    conn = ro.make_db((_iso(ts[i]), int(tok[0][i]), int(tok[1][i]), int(tok[2][i]), int(tok[3][i]), int(tok[4][i]), float(cost[i]), models[i], "P")
                      for i in range(m))
    for period in ("hour", "day"):
        t.rollup_rows(period)                                       # warm-up (uploads the columns)
        k, h = [], []
        for _ in range(5):
            t0 = time.perf_counter(); rows = t.rollup_rows(period); t1 = time.perf_counter()
            ms = t.last_ms(); k.append(ms["accum"] + ms["emit"]); h.append((t1 - t0) * 1e3)
        km, hm = float(np.median(k)), float(np.median(h))
        t0 = time.perf_counter(); ref = ro.aggregated_usage(conn, period); cpu_s = time.perf_counter() - t0
        lines.append({
            "metric": "usage records rolled up per second", "unit": "records/s", "value": n / km * 1e3, "ms_per_step": km,
            "config": {"workload": f"C5: {n} usage records -> GROUP BY {period}, model ({len(rows)} groups)"},
            "dtype": "int64 + fixed-point cost", "data": "synthetic",
            "e2e": {"value": n / hm * 1e3, "unit": "records/s", "ms_per_step": hm, "note": "records resident in HBM (ingested as they arrive); result rows copied back"},
            "roofline": {"bound": "hbm", "achieved": n * 40 / km / 1e6, "unit": "GB/s", "algorithmic_bytes": n * 40},
            "cpu_baseline": {"value": m / cpu_s, "unit": "records/s", "cores": 1, "kind": "reference",
                             "sample": f"the reference's own SQL statement on an in-memory SQLite table of {m} of the records ({len(ref)} groups)"}})
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=10_000_000)
    ap.add_argument("--bodies", type=int, default=1024)
    args = ap.parse_args()
    import llmapigateway_b200 as L
    eng = L.Engine(max_streams=64, max_step_chunks=1024, max_step_bytes=1 << 20)
    for line in bench_bodies(eng, args.bodies) + bench_rollup(eng, args.records):
        print(json.dumps(line))


if __name__ == "__main__":
    main()
