"""Golden of BASELINE config 1 ("plumbing"): ONE non-streaming /v1/chat/completions request with a body of exactly
256 bytes, a mock upstream that returns one fixed JSON document with `choices[0].message.content` + `usage`, through the UNMODIFIED
endpoint body (llm_gateway_core/api/v1/chat.py:20 -> request_handler.py:152-176), what FastAPI renders from the dict it returns
(Starlette JSONResponse) and what the logging middleware's tap stores for it in its non-streaming mode (chat_logging.py:98-150).
Dev container only; writes tests/golden/c1_case.json.

    python tests/golden/make_c1_golden.py

The attempt's wire body comes from `json5.dumps` (request_handler.py:153); json5 is absent from this image, so that one field is
produced by the stand-in of ref_driver and is marked unpinned.  Everything else is the reference's own code.
"""
from __future__ import annotations

import asyncio
import base64
import json
import sys
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE))

import make_chain_golden as mcg                                       # noqa: E402
import ref_driver                                                     # noqa: E402

B64 = lambda b: base64.b64encode(bytes(b)).decode("ascii")

UPSTREAM_DOC = (b'{"id":"chatcmpl-c1","object":"chat.completion","created":1700000000,"model":"vendor/free-1",'
                b'"choices":[{"index":0,"message":{"role":"assistant","content":"Hello from the mock upstream."},"finish_reason":"stop"}],'
                b'"usage":{"prompt_tokens":3,"completion_tokens":2,"total_tokens":5}}')


def c1_body() -> bytes:
    head = b'{"model":"llmgateway/free-stack","messages":[{"role":"user","content":"'
    tail = b'"}]}'
    body = head + b"a" * (256 - len(head) - len(tail)) + tail
    assert len(body) == 256
    return body


def c1_world():
    """One rule for the requested model on a non-openrouter provider, and the fallback provider for everything else."""
    prov = types.SimpleNamespace
    providers = {"freeprov": prov(baseUrl="http://free.test/v1/", apikey="FREE_KEY_ENV"), "fb": prov(baseUrl="http://fb.test/v1", apikey=None)}
    rules = {"llmgateway/free-stack": {"fallback_models": [{"provider": "freeprov", "model": "vendor/free-1"}], "rotate_models": False}}
    return providers, rules, "fb"


def main():
    import httpx
    from starlette.responses import JSONResponse
    chat = mcg.load_chat()
    providers, rules, fallback_provider = c1_world()
    chat.settings.fallback_provider = fallback_provider
    loader = types.SimpleNamespace(providers_config=providers, fallback_rules=rules)
    attempts = []

    def handler(request):
        hdr = {k: v for k, v in request.headers.items() if k.lower() in ("authorization", "http-referer", "x-title", "content-type")}
        attempts.append(dict(url=str(request.url), body=B64(request.content), headers=hdr))
        return httpx.Response(200, headers={"content-type": "application/json"}, content=UPSTREAM_DOC)

    rh = sys.modules["llm_gateway_core.services.request_handler"]
    real_client = httpx.AsyncClient
    rh.httpx.AsyncClient = lambda **kw: real_client(transport=httpx.MockTransport(handler), **kw)
    try:
        resp = asyncio.run(chat.chat_completions(mcg.FakeRequest(c1_body(), {"Authorization": "Bearer client-key"}, loader)))
    finally:
        rh.httpx.AsyncClient = real_client
    assert isinstance(resp, dict)
    rendered = bytes(JSONResponse(content=resp).body)                   # what FastAPI sends for the returned dict
    rows, _ = ref_driver.run_tap([rendered], is_real_streaming=False)   # the middleware's tap over the response body (:188-190, :98-150)
    out = {"generator": "tests/golden/make_c1_golden.py", "reference": "unmodified llm_gateway_core (chat.py, request_handler.py, chat_logging.py)",
           "unpinned": ["attempts[].body (json5.dumps stand-in)"],
           "request_body": B64(c1_body()), "upstream_doc": B64(UPSTREAM_DOC), "status": 200, "response_body": B64(rendered),
           "attempts": attempts, "rows": json.dumps(rows, sort_keys=True)}
    (HERE / "c1_case.json").write_text(json.dumps(out, indent=1) + "\n")
    print("wrote c1_case.json:", len(rendered), "response bytes,", len(attempts), "attempt,", rows)


if __name__ == "__main__":
    main()
