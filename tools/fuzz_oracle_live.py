"""Live pinning of the SSE oracle against the UNMODIFIED reference (dev container only: needs /root/reference).

The parity of the CUDA path is proven against oracle/sse_oracle.py; the oracle itself is pinned to 236 reference-generated cases
(tests/golden/sse_cases.json).  This tool widens that pin: the stream generators of tools/fuzz_relay2_cpu.py (random, template
variants, usage spellings, odd usage skeletons, OpenAI-shaped, the benchmark's deltas; with recuts) go through

  * the real make_llm_request(..., is_streaming=True) and the real ChunkProcessorThread.run (tests/golden/ref_driver.py:
    request_handler.py:8-150, chat_logging.py:69-153, `json5` = the stdlib stand-in, strict-JSON inputs only), and
  * oracle.sse_oracle.run_stream,

and the two must agree on: failed / error_detail, the relayed chunks, whether the relay generator ends in an exception, the usage
rows, and llm_response_accum at every write_log call.

    python tools/fuzz_oracle_live.py --rounds 40 --streams 100 --seed 1
"""
from __future__ import annotations

import argparse
import json
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tests" / "golden", ROOT / "tools"):
    sys.path.insert(0, str(p))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=20)
    ap.add_argument("--streams", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--gens", default="random,template,usage,skeleton,openai,c3")
    args = ap.parse_args()
    if not Path("/root/reference").exists():
        print("needs /root/reference (dev container)"); return 2
    import logging
    import fuzz_relay2_cpu as F
    import ref_driver
    from golden_io import UNPINNED_DETAIL_PREFIX, canon_rows
    from oracle.sse_oracle import run_stream
    ref_driver.load_reference()
    logging.disable(logging.CRITICAL)            # the reference logs every chunk
    rng = random.Random(args.seed)
    gens = args.gens.split(",")
    bad = n_total = n_failed = n_rows = 0
    for r in range(args.rounds):
        gen, seed = rng.choice(gens), args.seed * 1000 + r
        streams = F.make_streams(gen, args.streams, seed)
        for i, chunks in enumerate(streams):
            chunks = [c for c in chunks if c]
            try:
                b"".join(chunks).decode("utf-8", errors="surrogatepass")
            except UnicodeDecodeError:
                pass
            status = 200 if rng.random() < 0.97 else rng.choice([400, 500, 503])
            want = ref_driver.run_relay(chunks, status)
            rows, texts = ref_driver.run_tap(want["emitted"]) if not want["failed"] else ([], [])
            relay, tap = run_stream(chunks, status)
            where = f"{gen}:{seed} stream {i}"
            try:
                assert relay.failed == want["failed"], "failed"
                if want["failed"] and isinstance(want["error_detail"], str) and want["error_detail"].startswith(UNPINNED_DETAIL_PREFIX):
                    assert relay.error_detail.startswith(UNPINNED_DETAIL_PREFIX), "error_detail (unpinned tail)"
                else:
                    assert relay.error_detail == want["error_detail"], "error_detail"
                assert relay.emitted == want["emitted"], "emitted chunks"
                assert relay.end_raises == (want["end_exception"] is not None), "end exception"
                assert canon_rows(tap.rows) == json.dumps(json.loads(json.dumps(rows, sort_keys=True)), sort_keys=True) or canon_rows(tap.rows) == canon_rows(rows), "usage rows"
                assert tap.transcripts == texts, "transcripts"
            except AssertionError as e:
                bad += 1
                print(f"FAIL {where}: {e}; chunks {chunks[:4]!r}"[:900], flush=True)
            n_total += 1; n_failed += want["failed"]; n_rows += len(rows)
        print(f"round {r} {gen}:{seed}: {n_total} streams so far ({n_failed} failed attempts, {n_rows} usage rows), {bad} disagreements", flush=True)
    print(f"{n_total - bad}/{n_total} streams: oracle == unmodified reference")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
