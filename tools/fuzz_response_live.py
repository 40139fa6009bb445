"""Live pinning of the non-streaming oracles against the UNMODIFIED reference (dev container only: needs /root/reference).

Random upstream response documents (random JSON with `error` / `detail` / `usage` / `choices` / `model` keys of every shape, odd
spellings, broken text, HTTP error statuses) go through

  * the real make_llm_request(..., is_streaming=False) (request_handler.py:152-176) + what chat.py:146 and Starlette's JSONResponse do
    with its result, and the real ChunkProcessorThread in non-streaming mode over the rendered body (tests/golden/ref_driver.py), and
  * oracle.response_oracle.normalise / oracle.sse_oracle.tap_nonstream (what the CUDA path is tested against),

and the two must agree on ok / fail / raise, the failure detail (when it is not the JSON library's own message), the rendered bytes
and the usage rows of the tap.

    python tools/fuzz_response_live.py --docs 20000 --seed 1
"""
from __future__ import annotations

import argparse
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tests" / "golden"):
    sys.path.insert(0, str(p))


def rand_doc(rng):
    import body_cases as bc
    d = bc.rand_body(rng)
    r = rng.random()
    if r < 0.25:
        d["error"] = rng.choice([None, {}, {"message": "boom"}, {"message": ""}, {"message": None}, {"code": 429}, "a string", ["x"], 0, 5, 2.5, True, False,
                                 {"message": {"nested": 1}}, {"message": ["l"]}, {"message": 7}, {"message": True}, {"message": 0.5}])
    if r < 0.15 or rng.random() < 0.12:
        d["detail"] = rng.choice([None, "", "Not found", 0, 3, False, True, {}, {"a": 1}, [], ["x"], 1.5, "é \U0001F600"])
    if rng.random() < 0.4:
        d["usage"] = rng.choice([{"prompt_tokens": rng.randrange(10**6), "completion_tokens": rng.randrange(10**5), "total_tokens": rng.randrange(10**6),
                                  "cost": rng.choice([0, 1.5e-5, 0.25, 3, None, "x"]), "completion_tokens_details": rng.choice([{"reasoning_tokens": rng.randrange(50)}, None, {}, 5]),
                                  "prompt_tokens_details": rng.choice([{"cached_tokens": rng.randrange(9)}, None, "s"])}, None, [], "x", {"prompt_tokens": None}, 7])
    if rng.random() < 0.4:
        d["choices"] = rng.choice([[{"message": {"content": "hi"}}], [{"delta": None}], "str", [], [{"message": {"content": 5}}], None, [{"message": None}], [5], {"a": 1}, {}])
    if rng.random() < 0.3:
        d["model"] = rng.choice(["m", "café", 5, None, ["m"]]); d["provider"] = rng.choice(["P", None, 3])
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=5000)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    if not Path("/root/reference").exists():
        print("needs /root/reference (dev container)"); return 2
    import logging
    import body_cases as bc
    import ref_driver
    from golden_io import canon_rows
    from oracle import response_oracle as ro
    from oracle.sse_oracle import tap_nonstream
    ref_driver.load_reference()
    logging.disable(logging.CRITICAL)
    rng = random.Random(args.seed)
    bad = n_ok = n_fail = n_raise = n_rows = 0
    for it in range(args.docs):
        r = rng.random()
        if r < 0.03:
            raw = rng.choice([b"", b"not json", b'{"a":1,}', b'{"a":NaN}', b'[1,2', b'\xff\xfe', b'{"a":"\\ud800"}', b'12', b'null', b'"s"', b'[{"error":1}]', b'["error"]', b'"an error string"'])
        else:
            raw = bc.spell(rng, rand_doc(rng), plain_keys=rng.random() < 0.5).encode("utf-8")
        status = 200 if rng.random() < 0.93 else rng.choice([201, 400, 404, 500, 503])
        want = ref_driver.run_nonstream(raw, status)
        kind, val = ro.normalise(status, raw, "http://upstream.test/v1/chat/completions")
        try:
            assert kind == want["kind"], f"kind {kind} != {want['kind']}"
            if kind == "ok":
                assert val == want["body"], "rendered bytes"
                rows, _ = ref_driver.run_tap([want["body"]], is_real_streaming=False)
                assert canon_rows(tap_nonstream([want["body"]]).rows) == canon_rows(rows), "tap rows"
                n_ok += 1; n_rows += len(rows)
            elif kind == "fail":
                unpinned = (isinstance(val, str) and val.startswith("<invalid json")) or (isinstance(want["detail"], str) and want["detail"].startswith("Invalid JSON response"))
                if not unpinned:
                    assert val == want["detail"], f"detail {val!r} != {want['detail']!r}"
                n_fail += 1
            else:
                n_raise += 1
        except AssertionError as e:
            bad += 1
            print(f"FAIL doc {it} status {status}: {e}; {raw[:300]!r}"[:900], flush=True)
    print(f"{args.docs - bad}/{args.docs} documents: oracle == unmodified reference ({n_ok} ok with {n_rows} tap rows, {n_fail} failed attempts, {n_raise} render errors)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
