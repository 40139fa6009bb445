"""Generate tests/golden/body_cases.json by driving the UNMODIFIED chat_completions endpoint
(llm_gateway_core/api/v1/chat.py:21) in-process (dev container only).

    python tests/golden/make_body_golden.py

make_llm_request is replaced by a recorder that fails every attempt, so the endpoint walks the whole
rule chain (including retries and sub-provider fallbacks); for every attempt we store the payload
dict (key order included) and the bytes the REAL installed httpx makes of it.
"""
import asyncio
import base64
import copy
import json
import logging
import sys
import types
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE)); sys.path.insert(0, str(HERE.parent))
import ref_driver  # noqa: E402  (json5 shim + sys.path for /root/reference)

ref_driver.load_reference()
logging.disable(logging.CRITICAL)
import httpx  # noqa: E402
from httpx._content import encode_json  # noqa: E402
import llm_gateway_core.db.model_rotation_db as mrd  # noqa: E402

mrd.ModelRotationDB.__init__ = lambda self, *a, **k: None          # the module-level instance at chat.py:16 writes under <root>/db
mrd.ModelRotationDB.get_next_model_index = lambda self, **k: 0
import llm_gateway_core.api.v1.chat as chat  # noqa: E402

RULES = {
    "gw/chain": {"rotate_models": False, "fallback_models": [
        {"provider": "plain", "model": "m-plain", "use_provider_order_as_fallback": False, "custom_body_params": {}, "custom_headers": {}},
        {"provider": "openrouter", "model": "m-or", "use_provider_order_as_fallback": False, "providers_order": ["A", "B"],
         "custom_body_params": {"reasoning_effort": "high", "temperature": 0.5}, "custom_headers": {"x-param": "demo"}},
        {"provider": "openrouter", "model": "m-sub", "use_provider_order_as_fallback": True, "providers_order": ["S1", "S2"],
         "custom_body_params": {}, "custom_headers": {}},
        {"provider": "plain", "model": "m-retry", "use_provider_order_as_fallback": False, "retry_count": 1, "retry_delay": 0,
         "custom_body_params": {"usage": {"include": False}, "top_p": 1, "stop": ["\n\n", "é"]}, "custom_headers": {}},
    ]},
}


class _Prov:
    def __init__(self, url):
        self.baseUrl = url
        self.apikey = "APIKEY_X"


PROVIDERS = {"plain": _Prov("http://plain.test/v1/"), "openrouter": _Prov("http://or.test/api/v1"), "fb": _Prov("http://fb.test")}

BODIES = [
    {"model": "gw/chain", "stream": True, "temperature": 1, "messages": [{"role": "user", "content": "hi"}]},
    {"model": "gw/chain", "messages": [{"role": "system", "content": "é 中 \U0001F600 \"q\" \\ \n\t /   \x7f"}, {"role": "user", "content": ""}],
     "max_tokens": 128, "temperature": 0.7, "top_p": 0.95, "usage": {"include": False}, "provider": {"order": ["client"]}, "n": 1},
    {"messages": [], "model": "gw/chain", "stream": False, "seed": -0, "frequency_penalty": -0.0, "presence_penalty": 1e-7, "big": 12345678901234567890123,
     "logit_bias": {"50256": -100}, "tools": [{"type": "function", "function": {"name": "f", "parameters": {"type": "object", "default": None, "enum": [1, 2.5, True]}}}]},
    {"model": "unknown-model", "messages": [{"role": "user", "content": "fallback provider path"}], "stream": True},
    {"model": "gw/chain", "allow_fallbacks": True, "messages": "<already a string>", "reasoning_effort": "low", "temperature": 1.5e3,
     "x": [1e16, 1e15, 123456.789, 0.0001, 0.00001, 1e22, 2.5, 100.0, 1.0, 0.1, 123456789012345.0, 1e-7, 6.02e23]},
    # a realistic tool-calling conversation (nested schemas, escapes, empty containers, nulls)
    {"model": "gw/chain", "stream": True, "stream_options": {"include_usage": True}, "max_completion_tokens": 2048, "parallel_tool_calls": False,
     "response_format": {"type": "json_schema", "json_schema": {"name": "answer", "strict": True, "schema": {"type": "object", "properties": {
         "city": {"type": "string", "description": "City name, e.g. \"Zürich\""}, "temp_c": {"type": "number", "minimum": -90.5, "maximum": 60},
         "tags": {"type": "array", "items": {"type": "string"}, "default": []}}, "required": ["city", "temp_c"], "additionalProperties": False}}},
     "tools": [{"type": "function", "function": {"name": "get_weather", "description": "Look up the weather.\nReturns °C.", "parameters": {
         "type": "object", "properties": {"q": {"type": "string"}, "days": {"type": "integer", "enum": [1, 3, 7]}, "units": {"type": ["string", "null"]}}, "required": ["q"]}}}],
     "tool_choice": {"type": "function", "function": {"name": "get_weather"}},
     "messages": [{"role": "system", "content": [{"type": "text", "text": "Be terse."}]},
                  {"role": "user", "content": "What's the weather in Zürich?\tUse the tool. Path: C:\\tmp\\x   end"},
                  {"role": "assistant", "content": None, "tool_calls": [{"id": "call_1", "type": "function", "function": {"name": "get_weather", "arguments": "{\"q\": \"Z\u00fcrich\", \"days\": 1}"}}]},
                  {"role": "tool", "tool_call_id": "call_1", "content": "{\"temp_c\": 21.5, \"tags\": []}"}],
     "metadata": {}, "logprobs": False, "top_logprobs": None, "seed": 1234567890123, "user": "u-🧪"},
    # the client already sends the keys the rules assign (in place replacement keeps their positions)
    {"provider": {"order": ["X"], "allow_fallbacks": True}, "usage": {"include": True}, "top_p": 0.5, "stop": None, "model": "gw/chain",
     "reasoning_effort": {"nested": [1, {"a": None}]}, "temperature": 0, "messages": [{"role": "user", "content": "q"}], "allow_fallbacks": None},
]


async def run_one(body):
    attempts = []

    async def fake_make_llm_request(target_url, headers, payload, is_streaming):
        attempts.append({"url": target_url, "headers": dict(headers), "payload": copy.deepcopy(payload), "is_streaming": bool(is_streaming)})
        return None, "forced failure"

    chat.make_llm_request = fake_make_llm_request
    chat.settings.fallback_provider = "fb"
    raw = json.dumps(body, ensure_ascii=False).encode("utf-8")
    state = types.SimpleNamespace(config_loader=types.SimpleNamespace(providers_config=PROVIDERS, fallback_rules=RULES))

    class _Req:
        app = types.SimpleNamespace(state=state)
        headers = {"Authorization": "Bearer k"}

        async def body(self):
            return raw

    status = 200
    try:
        await chat.chat_completions(_Req())
    except Exception as e:                       # HTTPException 503 after the chain is exhausted
        status = getattr(e, "status_code", 500)
    return raw, status, attempts


# chat.py:31-45: the 400 conditions, and how `stream` is read (is_streaming of the first attempt)
PARSE_CASES = [
    b'{"model":"gw/chain","stream":true,"messages":[]}', b'{"model":"gw/chain","stream":1,"messages":[]}', b'{"model":"gw/chain","stream":"yes"}',
    b'{"model":"gw/chain","stream":0}', b'{"model":"gw/chain","stream":null}', b'{"model":"gw/chain","stream":[]}', b'{"model":"gw/chain","stream":[0]}',
    b'{"model":"gw/chain"}', b'  {"stream":false,"model":"gw/chain"}  ', b'{"model":"gw\\u002fchain","stream":{}}',
    b'{"messages":[]}', b'{"model":"","messages":[]}', b'{"model":null}', b'{"model":0}', b'{"model":[]}', b'{"model":{}}', b'{"model":false}',
    b'{"model":["gw/chain"]}', b'{"model":7}', b'{"model":1.5}',
    b'[{"model":"gw/chain"}]', b'"gw/chain"', b'42', b'null', b'true', b'{}', b'', b'{"model":"gw/chain"', b'{"model":"gw/chain",}', b'\xff\xfe',
    b'{"model":"gw/chain"} trailing', b'{"x":{"model":"inner"}}', b'{"model":"a","model":""}', b'{"model":"","model":"gw/chain"}',
]


async def run_parse(raw):
    attempts = []

    async def fake_make_llm_request(target_url, headers, payload, is_streaming):
        attempts.append(bool(is_streaming))
        return None, "forced failure"

    chat.make_llm_request = fake_make_llm_request
    chat.settings.fallback_provider = "fb"
    state = types.SimpleNamespace(config_loader=types.SimpleNamespace(providers_config=PROVIDERS, fallback_rules=RULES))

    class _Req:
        app = types.SimpleNamespace(state=state)
        headers = {"Authorization": "Bearer k"}

        async def body(self):
            return raw

    try:
        await chat.chat_completions(_Req())
        return 200, None, attempts
    except Exception as e:
        return getattr(e, "status_code", 500), str(getattr(e, "detail", "")), attempts


def main():
    parse = []
    for raw in PARSE_CASES:
        status, detail, attempts = asyncio.run(run_parse(raw))
        parse.append({"body": base64.b64encode(raw).decode(), "http_status": status, "detail_head": (detail or "")[:26],
                      "is_streaming": attempts[0] if attempts else None})
    cases = []
    for body in BODIES:
        raw, status, attempts = asyncio.run(run_one(body))
        for a in attempts:
            _, stream = encode_json(a["payload"])
            a["httpx_bytes"] = base64.b64encode(b"".join(stream)).decode()
            a["payload_items"] = json.dumps(list(a["payload"].items()), ensure_ascii=True)   # key order preserved
            del a["payload"]
        cases.append({"body": base64.b64encode(raw).decode(), "status": status, "attempts": attempts})
    doc = {"generator": "tests/golden/make_body_golden.py", "httpx": httpx.__version__, "rules": RULES, "cases": cases, "parse": parse}
    out = HERE / "body_cases.json"
    out.write_text(json.dumps(doc))
    print("wrote", out, [len(c["attempts"]) for c in cases], [c["status"] for c in cases])
    print("parse:", [(p["http_status"], p["detail_head"][:14], p["is_streaming"]) for p in parse])


if __name__ == "__main__":
    main()
