"""CPU oracle for the request-body rewrite path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates, for ONE upstream attempt of llm_gateway_core/api/v1/chat.py:
  parse_body          :31-45    body bytes -> dict; model / stream; the 400 conditions
  rewrite_payload     :112-119, :135-139, :150, :164-168   deepcopy + in-place/appended key assignments
  render_*            request_handler.py:23 (httpx `json=`) and :153 (json5.dumps)

Serialisers (third-party, not under /root/reference):
  httpx 0.28.1 (installed): json.dumps(obj, ensure_ascii=False, separators=(",",":"), allow_nan=False)
                            [httpx/_content.py encode_json]  -- PINNED (the golden generator calls the real httpx)
  httpx 0.27.0 (requirements.txt:3 pin): json.dumps(obj)      -- stdlib defaults, restated from memory
  json5.dumps (PyPI json5, unpinned, absent): restated from SURVEY.md Appendix B -- UNPINNED
`json5.loads` of the request body is replaced by stdlib json.loads (strict JSON inputs only).

Pinning: tests/golden/body_cases.json is produced by tests/golden/make_body_golden.py, which drives
the UNMODIFIED chat_completions endpoint (chat.py:21) in-process and records the payload dict it
hands to make_llm_request plus the bytes the real httpx encoder makes of it.
"""
from __future__ import annotations

import copy
import json
import re

ES5_RESERVED = {
    "break", "case", "catch", "continue", "debugger", "default", "delete", "do", "else", "finally", "for", "function",
    "if", "in", "instanceof", "new", "return", "switch", "this", "throw", "try", "typeof", "var", "void", "while", "with",
    "class", "const", "enum", "export", "extends", "import", "super", "null", "true", "false",
    "implements", "interface", "let", "package", "private", "protected", "public", "static", "yield",
}
_IDENT = re.compile(r"[A-Za-z_$][A-Za-z0-9_$]*\Z")


def parse_body(raw: bytes):
    """chat.py:31-45.  Returns (status, body_dict, model, is_streaming); status 0 ok, 1 parse error
    (400 "Error reading request body"), 2 missing model (400 "Missing 'model'")."""
    try:
        body = json.loads(raw.decode("utf-8"))
        probe = copy.deepcopy(body)
        probe["messages"] = "<REMOVED>"          # :34-36 (raises for non-dict bodies -> 400)
        _ = probe["model"]                        # :36 (KeyError -> 400)
    except Exception:
        return 1, None, None, None
    model = body.get("model")
    if not model:
        return 2, body, model, body.get("stream", False)
    return 0, body, model, body.get("stream", False)


def rule_ops(rule: dict, provider_name: str, sub_provider=None, retry=False):
    """The key assignments chat.py makes for one attempt, in order: list of (key, value, only_if_absent)."""
    ops = [("model", rule.get("model"), False)]                                   # :113
    if provider_name == "openrouter":
        ops.append(("usage", {"include": True}, True))                             # :114-115
    for k, v in (rule.get("custom_body_params") or {}).items():                    # :116-119
        ops.append((k, v, False))
    ops.append(("model", rule.get("model"), False))                                # :135 / :160 re-assigned inside the retry loop
    order = rule.get("providers_order")
    if sub_provider is not None:                                                   # case 2, :164-168
        ops.append(("provider", {"order": [sub_provider]}, False))
        ops.append(("allow_fallbacks", False, False))
    elif order:                                                                    # case 1, :137-139
        ops.append(("provider", {"order": list(order)}, False))
        ops.append(("allow_fallbacks", False, False))
    if retry:                                                                      # :150 leaks into the retry
        ops.append(("messages", "<REMOVED>", False))
    return ops


def rewrite_payload(body: dict, ops) -> dict:
    payload = copy.deepcopy(body)                                                  # :112
    for key, value, only_if_absent in ops:
        if only_if_absent and key in payload:
            continue
        payload[key] = value
    return payload


def render_httpx028(payload) -> bytes:
    return json.dumps(payload, ensure_ascii=False, separators=(",", ":"), allow_nan=False).encode("utf-8")


def render_httpx027(payload) -> bytes:
    return json.dumps(payload).encode("utf-8")


_JSON5_NAMED = {"\\": "\\\\", '"': '\\"', "\n": "\\n", "\r": "\\r", "\b": "\\b", "\f": "\\f", "\t": "\\t", "\v": "\\v",
                "\0": "\\0", "\u2028": "\\u2028", "\u2029": "\\u2029"}


def _json5_string(s: str) -> str:
    out = ['"']
    for ch in s:
        if ch in _JSON5_NAMED:
            out.append(_JSON5_NAMED[ch])
        elif ord(ch) < 0x20 or ord(ch) > 0x7E:
            cp = ord(ch)
            if cp > 0xFFFF:
                cp -= 0x10000
                out.append("\\u%04x\\u%04x" % (0xD800 + (cp >> 10), 0xDC00 + (cp & 0x3FF)))
            else:
                out.append("\\u%04x" % cp)
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def json5_dumps(obj) -> str:
    """SURVEY.md Appendix B restatement of json5.dumps defaults (UNPINNED: the package is absent)."""
    if obj is True:
        return "true"
    if obj is False:
        return "false"
    if obj is None:
        return "null"
    if isinstance(obj, int):
        return repr(obj)
    if isinstance(obj, float):
        if obj != obj:
            return "NaN"
        if obj in (float("inf"), float("-inf")):
            return "Infinity" if obj > 0 else "-Infinity"
        return float.__repr__(obj)
    if isinstance(obj, str):
        return _json5_string(obj)
    if isinstance(obj, (list, tuple)):
        return "[" + ", ".join(json5_dumps(v) for v in obj) + "]"
    if isinstance(obj, dict):
        parts = []
        for k, v in obj.items():
            ks = k if (_IDENT.match(k) and k not in ES5_RESERVED) else _json5_string(k)
            parts.append(ks + ": " + json5_dumps(v))
        return "{" + ", ".join(parts) + "}"
    raise TypeError(type(obj))


def render_json5(payload) -> bytes:
    return json5_dumps(payload).encode("utf-8")


RENDERERS = {"httpx028": render_httpx028, "httpx027": render_httpx027, "json5": render_json5}


def rewrite(raw: bytes, ops, mode: str):
    """Whole row a1+a3+a4 for one attempt: (status, out_bytes)."""
    status, body, _, _ = parse_body(raw)
    if status:
        return status, b""
    return 0, RENDERERS[mode](rewrite_payload(body, ops))
