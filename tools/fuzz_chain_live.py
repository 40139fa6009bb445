"""Live differential of the fallback-chain walker against the UNMODIFIED endpoint (dev container only: needs /root/reference).

Random rule worlds (1-4 rules per gateway model; rotation, retries, sub-provider lists walked as fallback or sent as a hint,
custom body params incl. `model`, custom headers, providers with an env key / a literal key / no key, unknown models falling back
to the fallback provider) and a seeded failing upstream (synth.ChainUpstream).  Every request goes through

  * llm_gateway_core/api/v1/chat.py:20 `chat_completions` as it is in /root/reference (tests/golden/make_chain_golden.drive), and
  * llmapigateway_b200.chat.chat_completions over the fake engine (host build of the device machines),

in the same order (rotation state carries over on both sides), and the two must agree on: stream vs HTTPException (status and
detail text), the relayed bytes, and every upstream attempt's url, headers and wire body (tests/chain_cases.check_against_golden).

    python tools/fuzz_chain_live.py --worlds 30 --requests 40 --seed 1
"""
from __future__ import annotations

import argparse
import asyncio
import base64
import os
import random
import sys
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tests" / "golden"):
    sys.path.insert(0, str(p))


def random_world(rng):
    from llmapigateway_b200.synth import Provider
    providers = {"alpha": Provider("http://alpha.test/v1", "ALPHA_KEY_ENV"), "beta": Provider("http://beta.test/api/v1/", "sk-beta-literal"),
                 "gamma": Provider("http://gamma.test/v1", ""), "openrouter": Provider("http://openrouter.test/api/v1", "sk-or-literal")}
    subs = ["Chutes", "Targon", "DeepInfra", "Lambda", "Together"]

    def rule():
        r = {"provider": rng.choice(list(providers)), "model": rng.choice(["vendor/m-%d" % rng.randrange(9), "plain-model", "org/name:free"]),
             "use_provider_order_as_fallback": False, "custom_body_params": {}, "custom_headers": {}}
        if rng.random() < 0.35:
            r["retry_count"] = rng.randrange(0, 3); r["retry_delay"] = rng.choice([0, 0.001])
        if rng.random() < 0.35:
            r["providers_order"] = rng.sample(subs, rng.randrange(1, 4)); r["use_provider_order_as_fallback"] = rng.random() < 0.5
        if rng.random() < 0.3:
            r["custom_body_params"] = dict(rng.sample([("temperature", 0.25), ("model", "ignored/by-the-walker"), ("max_tokens", 64), ("top_p", 1),
                                                       ("usage", {"include": False}), ("stream", True), ("user", "rule-user")], rng.randrange(1, 4)))
        if rng.random() < 0.3:
            r["custom_headers"] = dict(rng.sample([("X-Route", "r%d" % rng.randrange(5)), ("Authorization", "Bearer overridden"), ("X-Title", "other")], rng.randrange(1, 3)))
        return r

    rules = {"gw/m%d" % k: {"rotate_models": rng.random() < 0.4, "fallback_models": [rule() for _ in range(rng.randrange(1, 5))]} for k in range(rng.randrange(2, 6))}
    return providers, rules, rng.choice(["openrouter", "alpha", "gamma"])


def run_world(seed: int, n_requests: int):
    import httpx
    from fastapi import HTTPException
    import chain_cases as cc
    import make_chain_golden as mcg
    from fake_engine import FakeEngine
    from llmapigateway_b200 import chat as our_chat, rewrite, synth
    from llmapigateway_b200.gateway import StreamBatcher
    rng = random.Random(seed)
    providers, rules, fallback_provider = random_world(rng)
    os.environ["ALPHA_KEY_ENV"] = "sk-alpha-from-env"
    loader = types.SimpleNamespace(providers_config=providers, fallback_rules=rules)
    up = synth.ChainUpstream(n_requests, 4, seed=seed, p_fail=rng.choice([0.2, 0.5, 0.8]), max_attempts=64)
    models = list(rules) + ["some/unknown-model"]
    reqs = []
    for sid in range(n_requests):
        model = rng.choice(models)
        if run_world.rich and rng.random() < 0.7:        # rich bodies: random JSON (nested values, floats in odd spellings, escapes, odd keys)
            import body_cases as bc
            d = bc.rand_body(rng)
            d["model"] = model
            d["stream"] = True
            body = bc.spell(rng, d).encode("utf-8")
        else:
            body = synth.chain_request_bodies(1, seed=seed * 1000 + sid, model=model, pad_to=rng.choice([120, 200, 256]))[0]
        if run_world.broken and rng.random() < 0.3:      # request-side failures (chat.py:31-45): 400 texts
            m = rng.random()
            if m < 0.2:
                body = rng.choice([b"", b"[1,2]", b'"str"', b"12", b"null", b"true", b"1.5", b'{"messages":[],"stream":true}', b'{"model":"","stream":true}',
                                   b'{"model":null}', b'{"model":0}', b'{"model":false}', b'{"model":[]}', b'{"model":{}}', b'{"model":"\xff"}', b'\xff\xfe{"model":"x"}',
                                   b'{"model": "x", ', b'{"model":"x"} trailing', b'{"model":"x",}', b"{'model':'x'}", b'{"model":"x","a":NaN}'])
            elif m < 0.6:
                raw = bytearray(body); k = rng.randrange(len(raw)); op = rng.randrange(3)
                if op == 0:
                    del raw[k]
                elif op == 1:
                    raw[k] = rng.choice(b'{}[]",:\\ 0a\x80\xe2')
                else:
                    raw.insert(k, raw[k])
                body = bytes(raw)
            else:
                body = body.replace(b'"model"', rng.choice([b'"Model"', b'"model "', b'"mode"']), 1)
        key = rng.choice(["", "rot-key-0", "rot-key-1"])
        reqs.append((sid, body, {"Authorization": f"Bearer {key}"} if key else {}, key))
    # ---- the reference ----
    ref_chat = run_world.ref_chat
    ref_chat.settings.fallback_provider = fallback_provider
    import llm_gateway_core.db.model_rotation_db as mdb
    ref_chat.model_rotation_db = mdb.ModelRotationDB()                 # a fresh rotation table per world (its file lives in a temp dir, see mcg.load_chat)
    try:
        os.remove(ref_chat.model_rotation_db.db_path)
    except OSError:
        pass
    ref_chat.model_rotation_db = mdb.ModelRotationDB()
    want = []
    for sid, body, headers, key in reqs:
        r = mcg.drive(ref_chat, loader, body, headers, up, sid)
        want.append(dict(name=f"w{seed}_r{sid}", group="live", sid=sid, body=base64.b64encode(body).decode(), api_key=key, **r))
    # ---- ours ----
    got = []

    async def go():
        batcher = StreamBatcher(FakeEngine(max_streams=8), window_s=0.0002)
        batcher.load_rules(rewrite.RulePlans(rules, fallback_provider=fallback_provider, stream_mode=cc.stream_mode()))
        rotation = our_chat.ModelRotation()
        real_rewrite = batcher.rewrite_bodies
        odd = []

        async def watched(bodies, plan_idx):                       # a body the engine does not model (status != ok) at ANY attempt of the walk:
            r = await real_rewrite(bodies, plan_idx)               #   before the first attempt the request is handed back (RequestNotModelled),
            odd.extend(st for st, _ in r if st != rewrite.BODY_OK)  #   later it counts as a failed attempt -- a documented limit, not compared
            return r
        batcher.rewrite_bodies = watched
        for sid, body, headers, key in reqs:
            attempts = []
            del odd[:]

            class _Body(httpx.AsyncByteStream):
                def __init__(self, chunks):
                    self.chunks = chunks

                async def __aiter__(self):
                    for c in self.chunks:
                        yield c

            def handler(request):
                a = len(attempts)
                hdr = {k: v for k, v in request.headers.items() if k.lower() in ("authorization", "x-route", "http-referer", "x-title", "content-type")}
                attempts.append(dict(url=str(request.url), body=bytes(request.content), headers=hdr))
                ans = up.stream_chunks(sid, a)
                if isinstance(ans, tuple):
                    return httpx.Response(ans[0], content=ans[1])
                return httpx.Response(200, headers={"content-type": "text/event-stream"}, stream=_Body(ans))

            factory = lambda **kw: httpx.AsyncClient(transport=httpx.MockTransport(handler), **kw)
            try:
                resp = await our_chat.chat_completions(cc.FakeRequest(body, headers, loader), batcher=batcher, rotation=rotation, client_factory=factory)
                out = []
                async for c in resp.body_iterator:
                    out.append(bytes(c))
                g = dict(kind="stream", emitted=b"".join(out))
            except HTTPException as e:
                g = dict(kind="http_exception", status=e.status_code, detail=e.detail)
            g["attempts"] = attempts
            g["not_modelled"] = bool(odd)
            got.append(g)

    asyncio.run(go())
    n_attempts = 0
    for case, g in zip(want, got):
        if g["not_modelled"] or (g["kind"] == "http_exception" and "not modelled by the engine" in str(g.get("detail"))):
            run_world.handed_back += 1            # a documented hand-back (duplicate keys, 17-digit floats, ...): the integrator's own path
            continue
        try:                                          # non-streaming requests cross the json5.dumps boundary (request_handler.py:153): unpinned here
            import json as _json
            doc = _json.loads(base64.b64decode(case["body"]))
            if isinstance(doc, dict) and doc.get("model") and not doc.get("stream", False):
                run_world.non_streaming += 1
                continue
        except Exception:
            pass
        if case["kind"] == "http_exception" and case.get("status") == 400 and g["kind"] == "http_exception" and g.get("status") == 400:
            # the text of a JSON syntax error is the JSON library's own (json5 in production, its stdlib stand-in here): only the prefix is
            # pinned; decode errors, KeyError 'model', the TypeErrors of non-object roots and "Missing 'model'" are compared in full
            ref_d, our_d = str(case["detail"]), str(g["detail"])
            if our_d == "Error reading request body: request body is not valid JSON":
                assert ref_d.startswith("Error reading request body: ") and any(t in ref_d for t in ("Expecting", "Extra data", "Unterminated", "Invalid", "control character")), (case["name"], ref_d, our_d)
            else:
                assert our_d == ref_d, (case["name"], ref_d, our_d)
            assert len(g["attempts"]) == len(case["attempts"]) == 0
            run_world.bad_requests += 1
            continue
        cc.check_against_golden(case, g)
        n_attempts += len(case["attempts"])
    kinds = {}
    for c in want:
        k = c["kind"] + str(c.get("status", ""))
        kinds[k] = kinds.get(k, 0) + 1
    return kinds, n_attempts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worlds", type=int, default=20)
    ap.add_argument("--requests", type=int, default=30)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--rich-bodies", action="store_true", help="request bodies of random JSON instead of the padded chat shape")
    ap.add_argument("--broken", action="store_true", help="a share of broken requests (400 answers of chat.py:31-45)")
    args = ap.parse_args()
    run_world.rich, run_world.handed_back, run_world.broken, run_world.bad_requests, run_world.non_streaming = args.rich_bodies, 0, args.broken, 0, 0
    if not Path("/root/reference").exists():
        print("needs /root/reference (dev container)"); return 2
    import make_chain_golden as mcg
    run_world.ref_chat = mcg.load_chat()
    bad = 0
    for w in range(args.worlds):
        seed = args.seed * 1000 + w
        try:
            kinds, n_att = run_world(seed, args.requests)
            print(f"ok   world {seed}: {args.requests} requests, {n_att} upstream attempts, outcomes {kinds}", flush=True)
        except AssertionError as e:
            bad += 1
            print(f"FAIL world {seed}: {str(e)[:600]}", flush=True)
    print(f"{args.worlds - bad}/{args.worlds} worlds agree with the unmodified endpoint ({run_world.handed_back} requests handed back as not modelled, {run_world.bad_requests} answered 400 on both sides, {run_world.non_streaming} non-streaming skipped)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
