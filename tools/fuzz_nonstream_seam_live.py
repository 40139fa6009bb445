"""Live differential of the NON-streaming seam against the UNMODIFIED reference (dev container only).

`gateway.make_llm_request(..., is_streaming=False)` (response plan + error-detail walk of the host build of the device machines, the
kind -> text mapping of llmapigateway_b200/responses.py) against the real make_llm_request(False) + chat.py:146 + Starlette's render
(tests/golden/ref_driver.run_nonstream) on random upstream documents (tools/fuzz_response_live.rand_doc) and HTTP statuses: the
rendered response bytes, or the failure detail.  Documents the engine reports as not modelled are counted, not compared.

    python tools/fuzz_nonstream_seam_live.py --docs 3000 --seed 1
"""
from __future__ import annotations

import argparse
import asyncio
import random
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tests" / "golden", ROOT / "tools"):
    sys.path.insert(0, str(p))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    if not Path("/root/reference").exists():
        print("needs /root/reference (dev container)"); return 2
    import logging
    import httpx
    import body_cases as bc
    import fuzz_response_live as R
    import ref_driver
    from fake_engine import FakeEngine
    from llmapigateway_b200 import rewrite as rw
    from llmapigateway_b200.gateway import StreamBatcher, make_llm_request
    ref_driver.load_reference()
    logging.disable(logging.CRITICAL)
    rng = random.Random(args.seed)
    url = "http://upstream.test/v1/chat/completions"
    docs = []
    for it in range(args.docs):
        if rng.random() < 0.03:
            raw = rng.choice([b"", b"not json", b'{"a":1,}', b'[1,2', b'\xff\xfe', b'12', b'null', b'"s"', b'[{"error":1}]', b'["error"]', b'"an error string"', b'{}', b'[]', b'""'])
        else:
            raw = bc.spell(rng, R.rand_doc(rng), plain_keys=rng.random() < 0.5).encode("utf-8")
        docs.append((raw, 200 if rng.random() < 0.93 else rng.choice([201, 400, 404, 500, 503])))

    async def ours():
        batcher = StreamBatcher(FakeEngine(max_streams=4), window_s=0.0005)
        plans = rw.RulePlans(bc.RULES, fallback_provider="fb")
        batcher.load_rules(plans)
        (st, payload), = await batcher.rewrite_bodies([b'{"model":"gw/chain","messages":[]}'], [plans.plan_index("gw/chain", 1, stream=False)])
        out = []
        for raw, status in docs:
            handler = lambda request, raw=raw, status=status: httpx.Response(status, headers={"content-type": "application/json"}, content=raw)
            resp, err = await make_llm_request(url, {}, payload, False, batcher=batcher,
                                               client_factory=lambda handler=handler, **kw: httpx.AsyncClient(transport=httpx.MockTransport(handler), **kw))
            tap = (await batcher.documents_usage([bytes(resp.body)]))[0] if resp is not None else None     # what log_chat_completions stores for it
            out.append((resp, err, tap))
        return out

    got = asyncio.run(ours())
    bad = n_ok = n_fail = n_exotic = 0
    from golden_io import canon_rows
    n_rows = 0
    for it, ((raw, status), (resp, err, tap)) in enumerate(zip(docs, got)):
        want = ref_driver.run_nonstream(raw, status, url)
        try:
            if resp is None and isinstance(err, str) and "not modelled by the engine" in err:
                n_exotic += 1; continue
            if want["kind"] == "ok":
                assert resp is not None and err is None, f"reference ok, ours failed: {err!r}"
                assert bytes(resp.body) == want["body"], "rendered bytes"
                n_ok += 1
                rows, exotic = tap
                if not exotic:                                  # the middleware's tap over the served body (chat_logging.py:98-150, non-streaming mode)
                    ref_rows, _ = ref_driver.run_tap([want["body"]], is_real_streaming=False)
                    assert canon_rows(rows) == canon_rows(ref_rows), f"tap rows {rows} != {ref_rows}"
                    n_rows += len(ref_rows)
            elif want["kind"] == "raise":                      # Starlette's render raises (NaN ...): the engine must not serve it
                assert resp is None, "reference cannot render this, ours served it"
            else:
                assert resp is None, "reference failed the attempt, ours served it"
                d = want["detail"]
                unpinned = isinstance(d, str) and (d.startswith("Invalid JSON response") or ("Unexpected error during request" in d and any(t in d for t in ("Expecting", "Extra data", "Unterminated", "Invalid", "codec", "associated with a value"))))
                if not unpinned:
                    assert err == d, f"detail {err!r} != {d!r}"
                n_fail += 1
        except AssertionError as e:
            bad += 1
            print(f"FAIL doc {it} status {status}: {e}; {raw[:260]!r}"[:900], flush=True)
    print(f"{args.docs - bad}/{args.docs} documents: product non-streaming seam == unmodified reference ({n_ok} served with {n_rows} tap rows compared, {n_fail} failed attempts, {n_exotic} reported as not modelled)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
