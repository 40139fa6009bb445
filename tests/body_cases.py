"""Seeded request-body generators shared by the CPU and GPU body-rewrite tests."""
from __future__ import annotations

import json
import random

KEYS = ["model", "messages", "stream", "temperature", "top_p", "usage", "provider", "allow_fallbacks", "n", "stop", "tools",
        "max_tokens", "reasoning_effort", "seed", "user", "x-y", "délai", "new", "class", "a b", "$ok", "_u", "9lives", "", "k\n", "\U0001F600k",
        "metadata", "logit_bias", "response_format", "presence_penalty"]
TEXTS = ["hi", "", "é 中 \U0001F600", "line\nbreak\ttab", "quote\" back\\slash /", "\x7f del", "  ", "\x00\x01\x1f\x0b\x0c\x08", "plain ascii text " * 9,
         "ünïcödé " * 20, "<REMOVED>", "a" * 31, "b" * 32, "c" * 33, "d" * 64 + "\"" + "e" * 5]
FLOATS = [0.0, -0.0, 1.0, 2.5, 0.1, 0.7, 1e16, 1e15, 123456.789, 0.0001, 0.00001, 1e22, 100.0, 123456789012345.0, 1e-7, 6.02e23,
          -3.25, 5e-324 * 0 + 1.5e300, 9.5e-290, 0.95, 1.0e5, 12345.0, 0.000123, 1234567.0e10]
FLOAT_TEXTS = ["1.50e-4", "1E5", "1e+5", "0.0", "-0.0", "-0e0", "0E-3", "100.000", "0.10", "1.0e-7", "12.5E+1", "123456789012345e-5", "0.000100", "1e0", "10e-1"]


def rand_value(rng: random.Random, depth=0):
    t = rng.random()
    if depth < 3 and t < 0.18:
        return {rng.choice(KEYS): rand_value(rng, depth + 1) for _ in range(rng.randint(0, 4))}
    if depth < 3 and t < 0.32:
        return [rand_value(rng, depth + 1) for _ in range(rng.randint(0, 4))]
    if t < 0.55:
        return rng.choice(TEXTS)
    if t < 0.65:
        return rng.choice([0, 1, -1, -100, 128, 2 ** 31, -2 ** 63, 12345678901234567890123, 7])
    if t < 0.8:
        return rng.choice(FLOATS)
    return rng.choice([True, False, None])


def rand_body(rng: random.Random):
    body = {}
    keys = rng.sample(KEYS, rng.randint(1, 10))
    if rng.random() < 0.85 and "model" not in keys:
        keys.insert(rng.randint(0, len(keys)), "model")
    for k in keys:
        body[k] = rand_value(rng)
    if "model" in body and rng.random() < 0.8:
        body["model"] = rng.choice(["gw/chain", "unknown-model", "m", "é-model"])
    if rng.random() < 0.5:
        body["messages"] = [{"role": rng.choice(["user", "system", "assistant"]), "content": rng.choice(TEXTS)} for _ in range(rng.randint(0, 4))]
    return body


def _str_text(rng, s: str) -> str:
    """a JSON string literal for s with randomly chosen (equivalent) spellings"""
    out = ['"']
    for ch in s:
        cp = ord(ch)
        r = rng.random()
        if ch == '"' or ch == "\\" or cp < 0x20:
            short = {'"': '\\"', "\\": "\\\\", "\n": "\\n", "\r": "\\r", "\t": "\\t", "\b": "\\b", "\f": "\\f"}
            out.append(short[ch] if ch in short and r < 0.7 else "\\u%04x" % cp)
        elif ch == "/" and r < 0.5:
            out.append("\\/")
        elif cp > 0xFFFF and r < 0.4:
            v = cp - 0x10000
            out.append("\\u%04X\\u%04x" % (0xD800 + (v >> 10), 0xDC00 + (v & 0x3FF)))
        elif r < 0.1 and cp <= 0xFFFF:
            out.append("\\u%04X" % cp)
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def spell(rng: random.Random, v, float_texts=True, plain_keys=False) -> str:
    """JSON text for v with random whitespace / escape spellings (json.loads(spell(v)) == v).
    plain_keys: keys are written without optional escapes (the shape client SDKs produce)."""
    ws = lambda: rng.choice(["", "", "", " ", "\n", "\t ", "  \r\n"])
    key = (lambda k: json.dumps(k, ensure_ascii=False)) if plain_keys else (lambda k: _str_text(rng, k))
    if isinstance(v, dict):
        return "{" + ws() + ",".join(ws() + key(k) + ws() + ":" + ws() + spell(rng, x, float_texts, plain_keys) + ws() for k, x in v.items()) + "}"
    if isinstance(v, list):
        return "[" + ws() + ",".join(ws() + spell(rng, x, float_texts, plain_keys) + ws() for x in v) + "]"
    if isinstance(v, str):
        return _str_text(rng, v)
    if isinstance(v, float) and float_texts and rng.random() < 0.3:
        return rng.choice(FLOAT_TEXTS)
    return json.dumps(v)


RULES = {
    "gw/chain": {"rotate_models": False, "fallback_models": [
        {"provider": "plain", "model": "m-plain", "use_provider_order_as_fallback": False, "custom_body_params": {}, "custom_headers": {}},
        {"provider": "openrouter", "model": "m-or", "use_provider_order_as_fallback": False, "providers_order": ["A", "B"],
         "custom_body_params": {"reasoning_effort": "high", "temperature": 0.5}, "custom_headers": {"x-param": "demo"}},
        {"provider": "openrouter", "model": "m-sub", "use_provider_order_as_fallback": True, "providers_order": ["S1", "S2"],
         "custom_body_params": {}, "custom_headers": {}},
        {"provider": "plain", "model": "m-retry", "use_provider_order_as_fallback": False, "retry_count": 1, "retry_delay": 0,
         "custom_body_params": {"usage": {"include": False}, "top_p": 1, "stop": ["\n\n", "é"]}, "custom_headers": {}},
        {"provider": "openrouter", "model": "é/模型 \U0001F600", "use_provider_order_as_fallback": False,
         "custom_body_params": {"usage": {"include": False, "x": None}, "new": 1.5, "a b": [1, 2.0, "\x7f "], "délai": {"class": True, "ok_1": []}},
         "custom_headers": {}},
    ]},
}
# the attempt sequence chat.py walks for "gw/chain" when every attempt fails (first four rules = the golden generator's table)
CHAIN_ATTEMPTS = [(0, -1, False), (1, -1, False), (2, 0, False), (2, 1, False), (3, -1, False), (3, -1, True)]


def attempt_for_oracle(rule_idx, sub_idx, retry):
    rule = RULES["gw/chain"]["fallback_models"][rule_idx]
    sub = rule["providers_order"][sub_idx] if sub_idx >= 0 else None
    return rule, rule["provider"], sub, retry


def error_detail_docs(rng, n):
    """Documents whose top-level "error" / "detail" take every shape request_handler.py:167-169 can meet (shared by the CPU and GPU tests)."""
    strings = ['"m"', '""', '"caf\\u00e9 \\ud83d\\ude00 \\"q\\" \\\\ \\/ \\n\\t"', '"中文"', '"x' + "y" * 300 + '"', '"\\u0041\\u0000z"']
    scalars = ["null", "true", "false", "0", "-0", "0.0", "0e3", "-0.0E-2", "7", "-12", "1.5", "2e3", "{}", "[]", "{ }", "[ ]"]
    containers = ['{"a":1}', '[1]', '{"message":"inner"}', '[[]]']
    for it in range(n):
        members = []
        if rng.random() < 0.8:
            ev = rng.choice(strings + scalars + containers + ['{"message":%s}' % rng.choice(strings + scalars + containers),
                                                              '{"code":1,"message":%s,"type":"x"}' % rng.choice(strings + scalars),
                                                              '{"message":"first","message":%s}' % rng.choice(strings + scalars)])
            members.append((rng.choice(['"error"', '"\\u0065rror"', '"err\\u006fr"']), ev))
        if rng.random() < 0.7:
            members.append((rng.choice(['"detail"', '"deta\\u0069l"']), rng.choice(strings + scalars + containers)))
        if rng.random() < 0.2:
            members.append(('"error"', rng.choice(strings + scalars + ['{"message":"dup"}'])))
        if not members:
            continue
        members += [('"other"', '{"error":{"message":"nested, not top level"},"detail":"no"}'), ('"errors"', '"x"'), ('"id"', "3")]
        rng.shuffle(members)
        ws = lambda: rng.choice(["", " ", "\n ", "\t"])
        raw = ("{" + ",".join(ws() + k + ws() + ":" + ws() + v + ws() for k, v in members) + "}").encode("utf-8")
        yield raw
