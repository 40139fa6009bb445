"""Timing probe for the body-rewrite kernels (config 2).  python tools/exp_body.py [n] [bytes]"""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import body_cases as bc  # noqa: E402
import llmapigateway_b200 as L  # noqa: E402
from llmapigateway_b200 import rewrite as rw  # noqa: E402
from llmapigateway_b200.synth import chat_bodies  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
size = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
e = L.Engine(max_streams=64, max_step_chunks=1024, max_step_bytes=1 << 20)
plans = rw.RulePlans(bc.RULES, fallback_provider="fb")
e.load_rules(plans)
import os
if os.environ.get("LGW_BODY_MODE") == "1":
    e.set_mode(1)      # exact machine only
base = chat_bodies(min(n, 1024), size, seed=2)
bodies = [base[i % len(base)] for i in range(n)]
buf, off = rw.pack_bodies(bodies)
for att, stream in (((1, -1, False), True), ((1, -1, False), False)):
    idx = np.full(n, plans.plan_index("gw/chain", *att, stream=stream), dtype=np.uint32)
    slot = 8192 if size <= 4096 else 2 * size + 1024
    for it in range(4):
        t0 = time.perf_counter()
        out, out_off, res = e.rewrite_packed(buf, off, idx, slot)
        t1 = time.perf_counter()
        ms = e.bodies_last_ms()
    ok = int((res["status"] == 0).sum())
    print(f"n={n} size={size} stream={stream} ok={ok} in={off[-1]} out={out_off[-1]} host_call_ms={1e3 * (t1 - t0):.3f} kernels={ms}")
    tot = ms["rewrite"] + ms["offsets"] + ms["pack"]
    print(f"   bodies/s={n / tot * 1e3:.3e}  GB/s(in+out)={(int(off[-1]) + int(out_off[-1])) / tot / 1e6:.2f}")
