"""Shared helpers of the fallback-chain (config 4) tests: the goldens made from the unmodified `chat_completions`
(tests/golden/make_chain_golden.py) and the drivers that replay them through the oracle / the product walker."""
from __future__ import annotations

import asyncio
import base64
import json
import os
import types
from pathlib import Path

from llmapigateway_b200 import synth

GOLDEN = Path(__file__).resolve().parent / "golden" / "chain_cases.json"
D64 = base64.b64decode


def load():
    doc = json.loads(GOLDEN.read_text())
    for k, v in doc["env"].items():
        os.environ[k] = v
    ups = {"sweep": synth.ChainUpstream(doc["sweep"]["n"], doc["sweep"]["n_events"], seed=doc["sweep"]["seed"], p_fail=doc["sweep"]["p_fail"]),
           "shapes": synth.ChainUpstream(doc["shapes"]["n"], doc["shapes"]["n_events"], seed=doc["shapes"]["seed"], p_fail=doc["shapes"]["p_fail"])}
    ups["request"] = ups["shapes"]
    return doc, ups


def stream_mode() -> str:
    import httpx
    return "httpx028" if tuple(int(x) for x in httpx.__version__.split(".")[:2]) >= (0, 28) else "httpx027"


def check_against_golden(case: dict, got: dict):
    assert got["kind"] == case["kind"], (case["name"], got)
    if case["kind"] == "http_exception":
        assert got["status"] == case["status"], case["name"]
        if case["name"] == "not_json":                      # the JSON library's own message: unpinned (json5 absent)
            assert got["detail"].startswith("Error reading request body: ")
        else:
            assert got["detail"] == case["detail"], case["name"]
    else:
        assert got["emitted"] == D64(case["emitted"]), case["name"]
    assert len(got["attempts"]) == len(case["attempts"]), case["name"]
    for a, b in zip(got["attempts"], case["attempts"]):
        assert a["url"] == b["url"], case["name"]
        assert bytes(a["body"]) == D64(b["body"]), (case["name"], bytes(a["body"])[-120:], D64(b["body"])[-120:])
        assert {k.lower(): v for k, v in a["headers"].items()} == b["headers"], case["name"]


class FakeRequest:
    def __init__(self, body: bytes, headers: dict, loader):
        self._body, self.headers = body, headers
        self.app = types.SimpleNamespace(state=types.SimpleNamespace(config_loader=loader))

    async def body(self):
        return self._body


def walk_product(engine_factory, cases, ups):
    """All golden cases, in order (rotation state carries over), through llmapigateway_b200.chat.chat_completions."""
    import httpx
    from fastapi import HTTPException
    from llmapigateway_b200 import chat, rewrite
    from llmapigateway_b200.gateway import StreamBatcher
    providers, rules, fallback_provider = synth.chain_world()
    loader = types.SimpleNamespace(providers_config=providers, fallback_rules=rules)
    results = []

    async def go():
        batcher = StreamBatcher(engine_factory(), window_s=0.0005)
        batcher.load_rules(rewrite.RulePlans(rules, fallback_provider=fallback_provider, stream_mode=stream_mode()))
        rotation = chat.ModelRotation()
        for case in cases:
            attempts = []
            up, sid = ups[case["group"]], case["sid"]

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
            headers = {"Authorization": f"Bearer {case['api_key']}"} if case["api_key"] else {}
            try:
                resp = await chat.chat_completions(FakeRequest(D64(case["body"]), headers, loader), batcher=batcher, rotation=rotation, client_factory=factory)
                out = []
                async for c in resp.body_iterator:
                    out.append(bytes(c))
                got = dict(kind="stream", emitted=b"".join(out))
            except HTTPException as e:
                got = dict(kind="http_exception", status=e.status_code, detail=e.detail)
            got["attempts"] = attempts
            results.append(got)

    asyncio.run(go())
    return results
