"""Goldens of the fallback-chain walk (config 4), made by driving the UNMODIFIED endpoint body
llm_gateway_core/api/v1/chat.py:20 `chat_completions` (dev container only; writes tests/golden/chain_cases.json).

What is real: chat_completions, make_llm_request (request_handler.py), ModelRotationDB (its file redirected to a temp
directory, model_rotation_db.py:15-22 hard-codes <root>/db which is read-only here), httpx (MockTransport upstream).
What is stubbed: `json5` (absent from this image: loads -> stdlib json on strict JSON), `api/v1/models.py` (its import loads the
config files from disk; not on this path), the Request object (body(), headers, app.state.config_loader), env vars of the
provider keys.  The upstream is llmapigateway_b200.synth.ChainUpstream (seeded failure injection).

Per case the golden holds: the request, what came back (relayed bytes / HTTPException status+detail), and every upstream
attempt in order (url, the body bytes httpx put on the wire, the headers chat.py set).
"""
from __future__ import annotations

import asyncio
import base64
import json
import os
import sys
import tempfile
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE))

import ref_driver                                                     # noqa: E402
from llmapigateway_b200 import synth                                  # noqa: E402

B64 = lambda b: base64.b64encode(bytes(b)).decode("ascii")


def load_chat():
    ref_driver.load_reference()
    import llm_gateway_core.db.model_rotation_db as mdb
    tmp = Path(tempfile.mkdtemp(prefix="lgw_rot_")) / "rotation.db"

    def _init(self, db_filename: str = "llmgateway_rotation.db"):
        self.db_path = tmp
        self._init_db()

    mdb.ModelRotationDB.__init__ = _init
    stub = types.ModuleType("llm_gateway_core.api.v1.models")
    from fastapi import APIRouter
    stub.router = APIRouter()
    sys.modules["llm_gateway_core.api.v1.models"] = stub
    import llm_gateway_core.api.v1.chat as chat
    return chat


class FakeRequest:
    def __init__(self, body: bytes, headers: dict, loader):
        self._body, self.headers = body, headers
        self.app = types.SimpleNamespace(state=types.SimpleNamespace(config_loader=loader))

    async def body(self):
        return self._body


def drive(chat, loader, body: bytes, headers: dict, upstream, sid: int):
    """One request through the real endpoint body.  `upstream.stream_chunks(sid, attempt)` answers attempt number `attempt`."""
    import httpx
    attempts = []

    class _Body(httpx.AsyncByteStream):
        def __init__(self, chunks):
            self.chunks = chunks

        async def __aiter__(self):
            for c in self.chunks:
                yield c

    def handler(request):
        a = len(attempts)
        hdr = {k: v for k, v in request.headers.items() if k.lower() in ("authorization", "x-route", "http-referer", "x-title", "content-type")}
        attempts.append(dict(url=str(request.url), body=B64(request.content), headers=hdr))
        ans = upstream.stream_chunks(sid, a)
        if isinstance(ans, tuple):
            return httpx.Response(ans[0], content=ans[1])
        return httpx.Response(200, headers={"content-type": "text/event-stream"}, stream=_Body(ans))

    real_client = httpx.AsyncClient
    rh = sys.modules["llm_gateway_core.services.request_handler"]

    def patched(**kw):
        return real_client(transport=httpx.MockTransport(handler), **kw)

    async def go():
        from fastapi import HTTPException
        rh.httpx.AsyncClient = patched
        try:
            try:
                resp = await chat.chat_completions(FakeRequest(body, headers, loader))
            except HTTPException as e:
                return dict(kind="http_exception", status=e.status_code, detail=e.detail)
            out, end_exc = [], None
            try:
                async for c in resp.body_iterator:
                    out.append(bytes(c))
            except Exception as e:                                     # request_handler.py:144 when no usage was seen
                end_exc = type(e).__name__
            return dict(kind="stream", emitted=B64(b"".join(out)), n_chunks=len(out), end_exception=end_exc)
        finally:
            rh.httpx.AsyncClient = real_client

    res = asyncio.run(go())
    res["attempts"] = attempts
    return res


def main():
    chat = load_chat()
    providers, rules, fallback_provider = synth.chain_world()
    chat.settings.fallback_provider = fallback_provider
    os.environ["ALPHA_KEY_ENV"] = "sk-alpha-from-env"
    loader = types.SimpleNamespace(providers_config=providers, fallback_rules=rules)
    cases = []
    # (a) the seeded C4 sweep at small size and a high failure rate, so that every depth of the chain and the 503 occur
    n, n_events = 96, 6
    up = synth.ChainUpstream(n, n_events, seed=4, p_fail=0.45)
    bodies = synth.chain_request_bodies(n, seed=4)
    for sid in range(n):
        r = drive(chat, loader, bodies[sid], {"Authorization": "Bearer client-key"}, up, sid)
        cases.append(dict(name=f"c4_sweep_{sid}", group="sweep", sid=sid, body=B64(bodies[sid]), api_key="client-key", **r))
    # (b) the other rule shapes: rotation (state carried from request to request), retries (log scrub), sub-providers
    up2 = synth.ChainUpstream(64, 4, seed=44, p_fail=0.5)
    k = 0
    for model, reps in (("gw/rotating", 8), ("gw/retrying", 10), ("gw/or-fallback", 12), ("gw/or-hint", 10), ("some/unknown-model", 4)):
        for rep in range(reps):
            body = synth.chain_request_bodies(1, seed=100 + k, model=model, pad_to=200)[0]
            key = "rot-key-%d" % (rep % 2)
            r = drive(chat, loader, body, {"Authorization": f"Bearer {key}"}, up2, k)
            cases.append(dict(name=f"{model}_{rep}", group="shapes", sid=k, body=B64(body), api_key=key, **r))
            k += 1
    # (c) request-side failures (chat.py:31-45)
    for name, body in (("missing_model", b'{"messages":[],"stream":true}'), ("empty_model", b'{"model":"","stream":true}'),
                       ("null_model", b'{"model":null}'), ("not_json", b'{"model": "x", '), ("array_root", b'[1,2]'), ("bad_utf8", b'{"model":"\xff"}')):
        r = drive(chat, loader, body, {}, up2, 0)
        cases.append(dict(name=name, group="request", sid=0, body=B64(body), api_key="", **r))
    doc = dict(generator="tests/golden/make_chain_golden.py", reference="llm_gateway_core/api/v1/chat.py:20 chat_completions (unmodified)",
               sweep=dict(n=n, n_events=n_events, seed=4, p_fail=0.45), shapes=dict(n=64, n_events=4, seed=44, p_fail=0.5),
               env={"ALPHA_KEY_ENV": "sk-alpha-from-env"}, cases=cases)
    (HERE / "chain_cases.json").write_text(json.dumps(doc, indent=0, sort_keys=True) + "\n")
    kinds = {}
    for c in cases:
        kk = c["kind"] + (str(c.get("status", "")))
        kinds[kk] = kinds.get(kk, 0) + 1
    print(len(cases), "cases", kinds, "attempt counts", sorted({len(c["attempts"]) for c in cases}))


if __name__ == "__main__":
    main()
