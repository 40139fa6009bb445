"""TEST INFRASTRUCTURE (never imported by the product): CPU restatement of the fallback-chain walk of
llm_gateway_core/api/v1/chat.py:20-198 for streaming requests, on top of the other oracles:

  request parse / model / stream ......... chat.py:31-45          (json.loads stands in for json5.loads on strict JSON)
  rule lookup, rotation .................. chat.py:47-78          (model_rotation_db.py:56-110 restated in `Rotation`)
  per-attempt payload .................... chat.py:112-119,135-139,160-165   (oracle.body_oracle)
  attempt = make_llm_request(stream) ..... request_handler.py:8-187          (oracle.sse_oracle.run_stream)
  failure bookkeeping / 503 .............. chat.py:149-156,176-183,197-198

Pinned by tests/golden/chain_cases.json, which tests/golden/make_chain_golden.py produced by driving the UNMODIFIED
`chat_completions` (tests/test_oracle_golden.py::test_chain_oracle_matches_the_reference)."""
from __future__ import annotations

import copy
import json
import os

from . import body_oracle, sse_oracle


class Rotation:
    """model_rotation_db.py:56-110: first use of (api_key, model) -> 0, then (last + 1) % total."""

    def __init__(self):
        self.last = {}

    def next_index(self, api_key: str, model: str, total: int) -> int:
        if total <= 0:
            return 0
        key = (api_key, model)
        self.last[key] = 0 if key not in self.last else (self.last[key] + 1) % total
        return self.last[key]


def walk(body: bytes, headers: dict, providers: dict, rules: dict, fallback_provider: str, upstream, rotation: Rotation, mode: str = "httpx028"):
    """-> dict(kind="stream", emitted=bytes) | dict(kind="http_exception", status, detail), plus attempts=[dict(url, body, headers)].
    `upstream(attempt_number) -> (status>=400, text body) | list of network chunks`."""
    attempts = []
    try:                                                                   # chat.py:30-39
        doc = json.loads(body.decode("utf-8"))
        log = copy.deepcopy(doc)
        log["messages"] = "<REMOVED>"
        log["model"]
    except Exception as e:
        return dict(kind="http_exception", status=400, detail=f"Error reading request body: {str(e)}", attempts=attempts)
    requested = doc.get("model")
    if not requested:                                                      # :44-45
        return dict(kind="http_exception", status=400, detail="Missing 'model' in request body", attempts=attempts)
    cfg = rules.get(requested)
    if not cfg:                                                            # :49-54
        seq, rotate = [{"provider": fallback_provider, "model": requested}], False
    else:
        seq, rotate = cfg["fallback_models"], cfg["rotate_models"]
    api_key = headers.get("Authorization", "").replace("Bearer ", "")      # :61
    if rotate and len(seq) > 1:                                            # :63-78
        start = rotation.next_index(api_key, requested, len(seq))
        seq = seq[start:] + seq[:start]
    last = "No providers were attempted."
    payload_doc = None
    for rule in seq:                                                       # :83
        name, model = rule.get("provider"), rule.get("model")
        retry_count = rule.get("retry_count") or 0
        subs = rule.get("providers_order")
        pc = providers.get(name)
        key = os.getenv(pc.apikey) if pc.apikey else None                  # :96-101
        if not key and pc.apikey:
            key = pc.apikey
        hdr = {"Content-Type": "application/json", "HTTP-Referer": "https://github.com/fabiojbg/LLMApiGateway", "X-Title": "LLMGateway"}
        if key:
            hdr["Authorization"] = f"Bearer {key}"
        url = f"{pc.baseUrl.rstrip('/')}/chat/completions"                 # :111
        payload_doc = copy.deepcopy(doc)                                   # :112-119
        payload_doc["model"] = model
        if name == "openrouter" and "usage" not in payload_doc:
            payload_doc["usage"] = {"include": True}
        for k, v in (rule.get("custom_body_params") or {}).items():
            payload_doc[k] = v
        for k, v in (rule.get("custom_headers") or {}).items():
            hdr[k] = v

        def attempt():
            wire = body_oracle.render_httpx028(payload_doc) if mode == "httpx028" else body_oracle.render_httpx027(payload_doc)
            attempts.append(dict(url=url, body=wire, headers=dict(hdr)))
            ans = upstream(len(attempts) - 1)
            if isinstance(ans, tuple):                                     # request_handler.py:25-30
                return None, bytes(ans[1]).decode("utf-8")
            r = sse_oracle.RelayOracle(json.loads).run(list(ans), 200)
            if r.failed:
                return None, r.error_detail
            return b"".join(r.emitted), None

        while retry_count >= 0:                                            # :127
            if not subs or rule.get("use_provider_order_as_fallback", False) is False:
                payload_doc["model"] = model                               # :135
                if subs:
                    payload_doc["provider"] = {"order": subs}
                    payload_doc["allow_fallbacks"] = False
                out, err = attempt()
                if err is None:
                    return dict(kind="stream", emitted=out, attempts=attempts)
                payload_doc["messages"] = "<REMOVED>"                      # :150
                last = f"Model {model} failed with provider '{name}': {err}"
            else:
                for sp in subs:                                            # :158-183
                    payload_doc["model"] = model
                    payload_doc["provider"] = {"order": [sp]}
                    payload_doc["allow_fallbacks"] = False
                    out, err = attempt()
                    if err is None:
                        return dict(kind="stream", emitted=out, attempts=attempts)
                    last = f"Model '{model}' failed from provider '{name}' and sub-provider {sp} : {err}"
            retry_count -= 1
    return dict(kind="http_exception", status=503, detail=f"All configured providers failed for model '{requested}'. Last error: {last}", attempts=attempts)
