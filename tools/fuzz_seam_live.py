"""Live differential of the streaming seam against the UNMODIFIED reference (dev container only: needs /root/reference).

`llmapigateway_b200.gateway.make_llm_request(..., is_streaming=True)` + `StreamBatcher` over the fake engine (the host build of the
device machines: the host-side logic of the product with the exact device code underneath) against the real
`make_llm_request` + `ChunkProcessorThread` (tests/golden/ref_driver.py) on fuzzed upstream streams -- the generators of
tools/fuzz_relay2_cpu.py plus HTTP error statuses: failed / error detail, the relayed chunks WITH their boundaries, the usage rows
that reach the sink.  Streams the engine reports as holding a shape it does not model (n_exotic) are counted, their rows not compared.

    python tools/fuzz_seam_live.py --rounds 20 --streams 60 --seed 1
"""
from __future__ import annotations

import argparse
import asyncio
import json
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tests" / "golden", ROOT / "tools"):
    sys.path.insert(0, str(p))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=10)
    ap.add_argument("--streams", type=int, default=60)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--gens", default="random,template,usage,skeleton,openai,c3")
    ap.add_argument("--relay-from", default="device", choices=["device", "host"])
    args = ap.parse_args()
    if not Path("/root/reference").exists():
        print("needs /root/reference (dev container)"); return 2
    import logging
    import fuzz_relay2_cpu as F
    import ref_driver
    import test_gateway_cpu as T
    from fake_engine import FakeEngine
    from golden_io import UNPINNED_DETAIL_PREFIX, canon_rows
    from llmapigateway_b200.gateway import StreamBatcher
    ref_driver.load_reference()
    logging.disable(logging.CRITICAL)
    rng = random.Random(args.seed)
    gens = args.gens.split(",")
    tot = bad = n_exotic = n_failed = 0

    async def ours(cases):
        sink = T._Sink()
        batcher = StreamBatcher(FakeEngine(max_streams=16), window_s=0.0002, usage_sink=sink, relay_from=args.relay_from)
        return [await T._drive(c, batcher, sink) for c in cases]

    for r in range(args.rounds):
        gen, seed = rng.choice(gens), args.seed * 1000 + r
        cases = []
        for chunks in F.make_streams(gen, args.streams, seed):
            chunks = [c for c in chunks if c]
            cases.append({"chunks": chunks, "http_status": 200 if rng.random() < 0.95 else rng.choice([400, 429, 500, 503])})
        got = asyncio.run(ours(cases))
        for i, (c, g) in enumerate(zip(cases, got)):
            want = ref_driver.run_relay(c["chunks"], c["http_status"])
            rows, _ = ref_driver.run_tap(want["emitted"]) if not want["failed"] else ([], [])
            where = f"{gen}:{seed} stream {i}"
            tot += 1
            try:
                assert g["failed"] == want["failed"], "failed"
                if want["failed"]:
                    n_failed += 1
                    if isinstance(want["error_detail"], str) and want["error_detail"].startswith(UNPINNED_DETAIL_PREFIX):
                        assert g["error_detail"].startswith(UNPINNED_DETAIL_PREFIX), "error detail (unpinned tail)"
                    else:
                        assert g["error_detail"] == want["error_detail"], f"error detail {g['error_detail']!r} != {want['error_detail']!r}"
                    continue
                assert g["emitted"] == want["emitted"], "relayed chunks"
                try:
                    assert canon_rows(g["rows"]) == canon_rows(rows), f"rows {g['rows']} != {rows}"
                except TypeError:
                    n_exotic += 1               # an Unrepresentable value: a shape the engine reports instead of modelling
            except AssertionError as e:
                bad += 1
                print(f"FAIL {where}: {e}; chunks {c['chunks'][:3]!r}"[:1000], flush=True)
        print(f"round {r} {gen}:{seed}: {tot} streams so far, {n_failed} failed attempts, {n_exotic} exotic, {bad} disagreements", flush=True)
    print(f"{tot - bad}/{tot} streams: product seam == unmodified reference")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
