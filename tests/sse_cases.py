"""Named SSE stream cases shared by the golden generator and the parity tests.

Each case is (name, chunks, http_status).  The pinned cases cover SURVEY.md section 8(d)'s
adversarial set and Appendix A's quirks; `fuzz_cases` adds seeded random streams built from
an event pool with random re-chunking.
"""
from __future__ import annotations

import json
import random


def ev(obj_or_text, prefix="data: ", sep="\n\n") -> bytes:
    text = obj_or_text if isinstance(obj_or_text, str) else json.dumps(obj_or_text, separators=(",", ":"), ensure_ascii=False)
    return (prefix + text + sep).encode("utf-8")


def delta(text, **extra) -> bytes:
    d = {"id": "c1", "object": "chat.completion.chunk", "choices": [{"index": 0, "delta": {"content": text}}]}
    d.update(extra)
    return ev(d)


USAGE = {"prompt_tokens": 10, "completion_tokens": 7, "total_tokens": 17, "cost": 0.00123,
         "completion_tokens_details": {"reasoning_tokens": 2}, "prompt_tokens_details": {"cached_tokens": 4}}
USAGE_EV = ev({"choices": [], "usage": USAGE, "model": "m-ok", "provider": "P"})
DONE = b"data: [DONE]\n\n"


def rechunk(blob: bytes, cuts: list[int]) -> list[bytes]:
    pts = [0] + sorted(set(c for c in cuts if 0 < c < len(blob))) + [len(blob)]
    return [blob[a:b] for a, b in zip(pts, pts[1:])]


def pinned_cases():
    C = []
    add = lambda name, chunks, status=200: C.append((name, list(chunks), status))
    plain = [delta("he"), delta("llo"), delta(" world"), USAGE_EV, DONE]
    add("plain", plain)
    add("leading_comment_chunk", [b": PROCESSING\n\n"] + plain)
    add("comment_and_first_event_one_chunk", [b": PROCESSING\n\n" + delta("a"), delta("b"), USAGE_EV, DONE])
    first = delta("split")
    add("first_event_split", [first[:23], first[23:], delta("x"), USAGE_EV, DONE])
    add("first_event_split_three", [first[:10], first[10:40], first[40:], delta("x"), USAGE_EV])
    add("first_event_split_at_lf", [first[:-1], first[-1:], delta("x"), USAGE_EV])
    add("first_event_split_then_tail_event", [first[:23], first[23:] + delta("same-chunk"), delta("x"), USAGE_EV])
    u8 = delta("caf\u00e9 \u4e2d\u6587 \U0001F600 ok")
    cut = u8.index("\u00e9".encode()) + 1
    add("utf8_split_in_priming", [u8[:cut], u8[cut:], delta("x"), USAGE_EV])
    add("utf8_split_in_relay", [delta("a"), u8[:cut], u8[cut:], delta("x"), USAGE_EV, DONE])
    add("utf8_raw_ok", [delta("a"), u8, ev({"choices": [], "usage": USAGE, "model": "mod\u00e8le-\u4e2d", "provider": "Pr\u00f6v"})])
    add("invalid_utf8_chunk_relay", [delta("a"), b"data: {\"x\":\"\xff\xfe\"}\n\n", USAGE_EV])
    add("invalid_utf8_chunk_priming", [b"\xc3\n\n", delta("a"), USAGE_EV])
    add("overlong_utf8", [delta("a"), b'data: {"x":"\xc0\xaf"}\n\n' + USAGE_EV, delta("b")])
    add("surrogate_utf8", [delta("a"), b'data: {"x":"\xed\xa0\x80"}\n\n' + USAGE_EV, delta("b")])
    add("escapes_and_braces_in_strings",
        [delta('q"uo\\te {[}] \n\t\r\b\f / \u2028'), ev(r'{"choices":[{"delta":{"content":"a\"b\\\\\"c}{\/\u0041\ud83d\ude00"}}],"usage":' + json.dumps(USAGE) + "}")])
    add("data_no_space", [b'data:{"choices":[]}\n\n', delta("a"), USAGE_EV])
    add("data_two_spaces", [b'data:  {"choices":[]}\n\n', delta("a"), USAGE_EV])
    add("crlf_delimiters", [b'data: {"choices":[]}\r\n\r\n', b'data: {"a":1}\r\n\r\n', DONE.replace(b"\n", b"\r\n")])
    add("first_event_error", [ev({"error": {"message": "boom", "code": 429}})])
    add("first_event_error_after_comment", [b": hi\n\n", ev({"error": {"message": "boom"}}), delta("never")])
    add("first_event_detail", [ev({"detail": "Not allowed"}), delta("never")])
    add("first_event_detail_null", [ev({"detail": None, "choices": []})])
    add("first_event_error_nested_only", [ev({"choices": [{"error": 1, "detail": 2}]}), USAGE_EV])
    add("first_event_error_escaped_key", [ev(r'{"\u0065rror":1}'), delta("never")])
    add("first_event_error_in_second_part_same_chunk", [delta("a") + ev({"error": "late"}), USAGE_EV])
    add("midstream_error_with_code", [delta("a"), ev({"error": {"message": "x"}, "code": 502}), delta("b"), USAGE_EV])
    add("midstream_code_with_usage", [delta("a"), ev({"code": 1, "usage": USAGE}), DONE])
    add("midstream_error_no_code_after_usage", [delta("a"), USAGE_EV, ev({"error": {"message": "late"}}), DONE])
    add("usage_null", [delta("a"), ev({"choices": [], "usage": None, "model": "m", "provider": "p"})])
    add("usage_not_dict", [delta("a"), ev({"usage": [1, 2], "model": "m"}), ev({"usage": "str"})])
    add("ctd_null", [delta("a"), ev({"usage": {"prompt_tokens": 3, "completion_tokens": 9, "total_tokens": 12, "cost": 0.5, "completion_tokens_details": None, "prompt_tokens_details": {"cached_tokens": 1}}, "model": "m", "provider": "p"})])
    add("ptd_null", [delta("a"), ev({"usage": {"prompt_tokens": 3, "completion_tokens": 9, "completion_tokens_details": {"reasoning_tokens": 4}, "prompt_tokens_details": None}, "model": "m"})])
    add("reasoning_null", [delta("a"), ev({"usage": {"completion_tokens": 9, "completion_tokens_details": {"reasoning_tokens": None}, "prompt_tokens_details": {"cached_tokens": 2}}, "model": "m"})])
    add("completion_null_with_reasoning", [delta("a"), ev({"usage": {"completion_tokens": None, "completion_tokens_details": {"reasoning_tokens": 3}}, "model": "m"})])
    add("no_usage_at_all", [delta("a"), delta("b"), DONE])
    add("duplicate_usage_events", [delta("a"), USAGE_EV, ev({"usage": {"prompt_tokens": 1, "completion_tokens": 2, "total_tokens": 3}, "model": "second"}), DONE])
    add("duplicate_keys", [delta("a"), ev('{"usage":{"prompt_tokens":1},"usage":{"prompt_tokens":5,"prompt_tokens":6,"total_tokens":8},"model":"a","model":"b"}')])
    add("two_events_one_chunk", [delta("a") + delta("b"), delta("c") + USAGE_EV + DONE])
    big = delta("x" * 300)
    add("event_over_many_chunks", [delta("a")] + rechunk(big, list(range(17, len(big), 17))) + [USAGE_EV])
    add("triple_lf_quirk", [delta("a"), b'data: {"k":1}\n\n\n', USAGE_EV, DONE])
    add("triple_lf_midchunk", [delta("a"), b'data: {"k":1}\n\n\n' + USAGE_EV, DONE])
    add("quad_lf", [delta("a"), b'data: {"k":1}\n\n\n\n', USAGE_EV])
    add("lf_runs_across_chunks", [delta("a"), b'data: {"k":1}\n', b"\n", b"\n", b"\n" + USAGE_EV, b"\n\n\n", USAGE_EV[:5], USAGE_EV[5:]])
    add("single_lf_inside_event", [delta("a"), b'data: {"choices":[],\n"usage":{"prompt_tokens":4}}\n\n', DONE])
    add("extra_line_after_json", [delta("a"), b'data: {"usage":{"prompt_tokens":4}}\nid: 7\n\n', DONE])
    add("event_field_before_data", [b'event: x\ndata: {"usage":{"prompt_tokens":4}}\n\n', delta("a"), USAGE_EV])
    add("first_real_malformed_json", [b'data: {"choices":[}\n\n', delta("a"), USAGE_EV])
    add("first_real_truncated_json", [b'data: {"choices":[1,2\n\n', delta("a")])
    add("relay_malformed_json", [delta("a"), b'data: {"usage":{"prompt_tokens":4},}\n\n', b"data: {'usage':1}\n\n", b'data: {"usage":{"prompt_tokens":5}} x\n\n', DONE])
    add("http_500_text", [b"upstream ", b"exploded"], 500)
    add("http_404_json", [b'{"error":"nope"}'], 404)
    add("brace_prefix_parts", [delta("a"), b'{"usage":{"prompt_tokens":11},"model":"bare"}\n\n', b' {"usage":{"prompt_tokens":12}}\n\n', DONE])
    add("brace_prefix_trailing_ws", [delta("a"), b'{"usage":{"prompt_tokens":11}} \n\n', b'{"usage":{"prompt_tokens":13}}\x0c\n\n'])
    add("py_whitespace_strip", [delta("a"), b'data: {"usage":{"prompt_tokens":21}}\x0c \x1f\n\n', b'data: {"usage":{"prompt_tokens":22}}\t\r\n\n'])
    add("choices_null_with_usage", [delta("a"), ev({"choices": None, "usage": USAGE, "model": "m"})])
    add("choices_number", [delta("a"), ev({"choices": 5, "usage": USAGE})])
    add("choices_string_and_dict", [delta("a"), ev({"choices": "delta", "usage": {"prompt_tokens": 1}}), ev({"choices": {}, "usage": {"prompt_tokens": 2}})])
    add("delta_null", [delta("a"), ev({"choices": [{"delta": None}], "usage": USAGE})])
    add("choice_element_scalar", [delta("a"), ev({"choices": [1], "usage": USAGE}), ev({"choices": [None], "usage": USAGE})])
    add("content_variants", [delta("a"),
                             ev({"choices": [{"delta": {"content": ""}}, {"delta": {"content": None}}, {"delta": {"content": 0}}, {"delta": {"content": False}}, {"delta": {"content": []}}, {"delta": {"content": 0.0}}], "usage": {"prompt_tokens": 1}}),
                             ev({"choices": [{"delta": {"content": 5}}], "usage": {"prompt_tokens": 2}}),
                             ev({"choices": [{"delta": {"content": True}}], "usage": {"prompt_tokens": 3}}),
                             ev({"choices": [{"delta": {"content": ["x"]}}], "usage": {"prompt_tokens": 4}}),
                             ev({"choices": [{"delta": {}, "message": {"content": "via-message"}}], "usage": {"prompt_tokens": 5}}),
                             ev({"choices": [{"delta": {"role": "assistant"}, "message": None}], "usage": {"prompt_tokens": 6}}),
                             ev({"choices": [{"message": None}], "usage": {"prompt_tokens": 7}}),
                             ev({"choices": [{"message": {"content": "m1"}}, {"delta": {"content": "d2"}}], "usage": {"prompt_tokens": 8}})])
    add("usage_value_types", [delta("a"),
                              ev({"usage": {"prompt_tokens": 1.5, "completion_tokens": 10, "total_tokens": True, "cost": 3, "completion_tokens_details": {"reasoning_tokens": 2.5}}, "model": "m"}),
                              ev('{"usage":{"prompt_tokens":-0,"completion_tokens":1e2,"total_tokens":12345678901234567890123,"cost":1E-7},"model":null,"provider":17}')])
    add("usage_big_and_float_forms", [delta("a"), ev('{"usage":{"prompt_tokens":9223372036854775807,"completion_tokens":-9223372036854775808,"total_tokens":9223372036854775808,"cost":0.1e-2,"completion_tokens_details":{"reasoning_tokens":0}},"model":"m"}')])
    add("cost_forms", [delta("a")] + [ev('{"usage":{"cost":%s},"model":"c%d"}' % (t, i)) for i, t in enumerate(
        ["0", "-0.0", "1", "0.001234", "123456.789e-3", "1e308", "1e309", "4.9e-324", "2.4703282292062327e-324", "0.1", "0.30000000000000004",
         "9007199254740993", "9007199254740993.0", "1.7976931348623157e308", "8.5e-5", "123456789012345678901234567890.5", "NaN", "Infinity", "-Infinity"])])
    add("unicode_escape_keys", [delta("a"), ev(r'{"us\u0061ge":{"prompt_tokens":31,"\u0063ompletion_tokens":5},"m\u006fdel":"esc\u00e9\ud83d\ude00"}')])
    add("lone_surrogate_model", [delta("a"), ev(r'{"usage":{"prompt_tokens":1},"model":"x\ud800y"}')])
    add("deep_nesting", [delta("a"), ev('{"x":' + "[" * 40 + "]" * 40 + ',"usage":{"prompt_tokens":77}}'), ev('{"x":' + '{"a":' * 30 + "1" + "}" * 30 + ',"usage":{"prompt_tokens":78}}')])
    add("stream_ends_mid_event", [delta("a"), USAGE_EV, delta("cut")[:30]])
    add("only_done", [DONE])
    add("empty_stream", [])
    add("only_comment", [b": ping\n\n", b": ping\n\n"])
    add("first_part_leading_lf_quirk", [b": c\n\n\n", delta("x"), delta("y"), USAGE_EV])
    add("usage_then_model_only_event", [delta("a"), USAGE_EV, ev({"usage": {}, "provider": "only"}), DONE])
    add("tool_call_event", [ev({"choices": [{"index": 0, "delta": {"role": "assistant", "content": None, "tool_calls": [{"index": 0, "id": "call_1", "type": "function", "function": {"name": "f", "arguments": "{\"a\": [1, 2, {\"b\": null}]}"}}]}}]}), delta("after"), USAGE_EV, DONE])
    add("control_char_in_string", [delta("a"), b'data: {"usage":{"prompt_tokens":1},"x":"a\x01b"}\n\n', b'data: {"usage":{"prompt_tokens":2},"x":"a\x7fb"}\n\n'])
    add("bad_escape", [delta("a"), b'data: {"usage":{"prompt_tokens":1},"x":"\\q"}\n\n', b'data: {"usage":{"prompt_tokens":2},"x":"\\u12g4"}\n\n', b'data: {"usage":{"prompt_tokens":3},"x":"\\u00e9"}\n\n'])
    add("number_forms", [delta("a")] + [ev('{"usage":{"prompt_tokens":%d},"n":%s}' % (i, t)) for i, t in enumerate(
        ["01", "-", "1.", ".5", "1e", "1e+", "-0", "0e0", "1.5E+3", "+1", "0x10", "1_0", "--1", "1.2.3", "[1,]", "[,1]", "{}", "[]", "[[],{}]", "tru", "truee", "nul", "-Infinity", "-Inf", "Nan"])])
    return C


_POOL_TEXT = ["a", "hello", " wor", "ld", "\u00e9", "\u4e2d\u6587", "\U0001F600", 'q"x', "b\\s", "{", "}", "[1]", "line\nbreak", ""]


def _random_event(rng: random.Random) -> bytes:
    k = rng.random()
    if k < 0.55:
        return delta(rng.choice(_POOL_TEXT) + rng.choice(_POOL_TEXT))
    if k < 0.65:
        u = dict(USAGE)
        u["prompt_tokens"] = rng.randrange(0, 10**6)
        if rng.random() < 0.3:
            u.pop("completion_tokens_details")
        if rng.random() < 0.1:
            u["completion_tokens_details"] = None
        return ev({"choices": [], "usage": u, "model": rng.choice(["m1", "m2", "mod\u00e8l"]), "provider": rng.choice(["P", "Q"])})
    if k < 0.70:
        return DONE
    if k < 0.75:
        return b": keepalive\n\n"
    if k < 0.79:
        return ev({"error": {"message": "e"}, "code": 500}) if rng.random() < 0.5 else ev({"error": "plain"})
    if k < 0.82:
        return ev({"detail": "d"})
    if k < 0.86:
        return b'data: {"broken": \n\n'
    if k < 0.89:
        return delta("x").replace(b"\n\n", b"\n\n\n")
    if k < 0.92:
        return b'{"usage":{"prompt_tokens":%d}}\n\n' % rng.randrange(100)
    if k < 0.94:
        return b"data: {\"x\":\"\xff\"}\n\n"
    if k < 0.96:
        return delta("crlf").replace(b"\n\n", b"\r\n\r\n")
    if k < 0.98:
        return ev({"choices": [{"delta": None}], "usage": {"prompt_tokens": 1}})
    return b"\n"


def fuzz_cases(n: int = 160, seed: int = 20260921):
    rng = random.Random(seed)
    out = []
    for i in range(n):
        n_ev = rng.randrange(1, 9)
        events = [_random_event(rng) for _ in range(n_ev)]
        if rng.random() < 0.5:
            events.insert(0, delta("lead"))      # make most streams commit
        blob = b"".join(events)
        mode = rng.random()
        if mode < 0.35:
            chunks = events
        elif mode < 0.7:
            chunks = rechunk(blob, [rng.randrange(1, max(2, len(blob))) for _ in range(rng.randrange(0, 8))])
        else:
            step = rng.randrange(1, 40)
            chunks = rechunk(blob, list(range(step, len(blob), step)))
        out.append(("fuzz_%03d" % i, chunks, 200))
    return out


def all_cases():
    return pinned_cases() + fuzz_cases()
