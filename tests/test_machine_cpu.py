"""CPU fuzz of the DEVICE code paths (host build, test aid) against the golden fixtures produced by
the unmodified reference, and of the part parser / number conversion against CPython."""
import json
import math
import random
import struct

import pytest

import host_machine as hm
from golden_io import load_sse_cases
from llmapigateway_b200 import _abi
from stream_compare import check_stream

CASES = load_sse_cases()
# shapes whose Python behaviour the device reports as "exotic" instead of modelling
EXPECTED_EXOTIC = {"usage_value_types", "usage_big_and_float_forms", "lone_surrogate_model"}


def _expect(case):
    return dict(failed=case["failed"], error_detail=case["error_detail"], emitted=case["emitted"],
                end_raises=case["end_raises"], rows=case["rows"], http_status=case["http_status"])


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
@pytest.mark.parametrize("stepping", ["one_step", "step_per_chunk", "random_steps"])
def test_stream_machine_matches_reference(case, stepping):
    chunks = case["chunks"]
    n = len(chunks)
    if stepping == "one_step" or n == 0:
        steps = [0]
    elif stepping == "step_per_chunk":
        steps = list(range(n))
    else:
        rng = random.Random(hash(case["name"]) & 0xFFFF)
        steps = [0] + sorted(set(rng.randrange(1, n) for _ in range(rng.randrange(0, 4)))) if n > 1 else [0]
    r = hm.run_stream(chunks, steps, case["http_status"])
    verdict = check_stream(_expect(case), chunks, r["step_chunk"], r["segs"], r["state"], r["detail"], r["rows"], case["name"])
    if verdict == "exotic":
        assert case["name"] in EXPECTED_EXOTIC or case["name"].startswith("fuzz_"), case["name"]


def test_exotic_share_is_small():
    n_exotic = 0
    for case in CASES:
        r = hm.run_stream(case["chunks"], [0], case["http_status"])
        n_exotic += bool(r["state"].n_exotic)
    assert n_exotic <= len(EXPECTED_EXOTIC) + 2


# ---- part parser vs CPython ---------------------------------------------------------------------

def _py_view(text: str):
    """What CPython + the reference's reads make of one event text `data: {...}`."""
    body = text[len("data: "):]
    out = {}
    try:
        doc = json.loads(body); out["valid_a"] = True
    except Exception:
        doc = None; out["valid_a"] = False
    try:
        docb = json.loads(body.strip()); out["valid_b"] = True
    except Exception:
        docb = None; out["valid_b"] = False
    out["doc"] = docb if docb is not None else doc
    return out


def _walk_choices(doc):
    """chat_logging.py:124-133; returns (type_error, appended_any)."""
    acc = ""
    try:
        if "choices" in doc:
            for ch in doc["choices"]:
                if "delta" in ch and "content" in ch["delta"]:
                    p = ch["delta"]["content"]
                    if p:
                        acc += p
                elif "message" in ch and "content" in ch["message"]:
                    p = ch["message"]["content"]
                    if p:
                        acc += p
    except Exception:
        return True, bool(acc)
    return False, bool(acc)


_ATOMS = ['0', '1', '-1', '17', '3.5', '-0.0', '1e3', '1E-2', 'true', 'false', 'null', '""', '"x"', '"\\u00e9\\n"', '"a\\"b"',
          '[]', '{}', '[1,2]', '{"k":1}', '"content"', '"delta"', 'NaN', 'Infinity', '-Infinity', '12345678901234567890', '0.1e-2']


def _rand_value(rng, depth=0):
    k = rng.random()
    if depth > 3 or k < 0.6:
        return rng.choice(_ATOMS)
    if k < 0.8:
        return "[" + ",".join(_rand_value(rng, depth + 1) for _ in range(rng.randrange(0, 4))) + "]"
    keys = ["a", "delta", "message", "content", "usage", "choices", "prompt_tokens", "x y", "\\u0061"]
    return "{" + ",".join('"%s":%s' % (rng.choice(keys), _rand_value(rng, depth + 1)) for _ in range(rng.randrange(0, 4))) + "}"


def _rand_event(rng):
    keys = ["choices", "usage", "error", "detail", "code", "model", "provider", "id", "us\\u0061ge", "choices"]
    usage_keys = ["prompt_tokens", "completion_tokens", "total_tokens", "cost", "completion_tokens_details",
                  "prompt_tokens_details", "other"]
    parts = []
    for _ in range(rng.randrange(0, 6)):
        k = rng.choice(keys)
        if k == "usage" and rng.random() < 0.7:
            inner = []
            for _ in range(rng.randrange(0, 6)):
                uk = rng.choice(usage_keys)
                if uk.endswith("_details") and rng.random() < 0.6:
                    dk = rng.choice(["reasoning_tokens", "cached_tokens", "z"])
                    inner.append('"%s":{"%s":%s}' % (uk, dk, _rand_value(rng, 3)))
                else:
                    inner.append('"%s":%s' % (uk, _rand_value(rng, 3)))
            v = "{" + ",".join(inner) + "}"
        elif k == "choices" and rng.random() < 0.7:
            els = []
            for _ in range(rng.randrange(0, 3)):
                if rng.random() < 0.75:
                    side = rng.choice(["delta", "message"])
                    if rng.random() < 0.7:
                        els.append('{"%s":{"content":%s}}' % (side, _rand_value(rng, 3)))
                    else:
                        els.append('{"%s":%s}' % (side, _rand_value(rng, 2)))
                else:
                    els.append(_rand_value(rng, 2))
            v = "[" + ",".join(els) + "]"
        else:
            v = _rand_value(rng, 1)
        parts.append('"%s":%s' % (k, v))
    text = "{" + ",".join(parts) + "}"
    m = rng.random()
    if m < 0.15:                      # mutate: delete / duplicate / replace a char
        i = rng.randrange(len(text))
        text = text[:i] + rng.choice(["", text[i] * 2, ",", "}", '"', " ", "\n", "\\"]) + text[i + 1:]
    elif m < 0.25:
        text = text + rng.choice([" ", "\n", "\t\r", "\x0c", " x", "}", "\x1f "])
    elif m < 0.30:
        text = text.replace(",", " ,\n ").replace(":", " : ")
    return text


def _unrepresentable(v):
    return isinstance(v, (str, list, dict)) or (isinstance(v, int) and not isinstance(v, bool) and not -2**63 <= v < 2**63)


def test_part_parser_matches_cpython():
    from oracle.sse_oracle import token_usage
    rng = random.Random(12345)
    n_exotic = n_checked = n_valid = n_usage = 0
    for it in range(30000):
        body = _rand_event(rng)
        text = "data: " + body
        if not body.startswith("{"):
            continue
        f, cls, rec = hm.parse_part(text.encode("utf-8"))
        assert f != 0xFFFFFFFF, text            # flags-only machine == extracting machine
        lean_mask = _abi.PF_VALID_A | _abi.PF_VALID_B | _abi.TK_ERROR | _abi.TK_DETAIL | _abi.TK_CODE | _abi.TK_USAGE | _abi.PF_TOO_DEEP
        assert hm.lean_parse(text.encode("utf-8")) == (f & lean_mask), text      # the bulk kernel's lean recogniser
        assert cls == 1
        py = _py_view(text)
        assert bool(f & _abi.PF_VALID_A) == py["valid_a"], text
        assert bool(f & _abi.PF_VALID_B) == py["valid_b"], text
        if not py["valid_b"]:
            continue
        n_valid += 1
        doc = py["doc"]
        for bit, key in ((_abi.TK_ERROR, "error"), (_abi.TK_DETAIL, "detail"), (_abi.TK_CODE, "code"),
                         (_abi.TK_USAGE, "usage"), (_abi.TK_CHOICES, "choices"), (_abi.TK_MODEL, "model"),
                         (_abi.TK_PROVIDER, "provider")):
            assert bool(f & bit) == (key in doc), (key, text)
        if f & _abi.PF_EXOTIC:
            n_exotic += 1
            continue
        terr, appended = _walk_choices(doc)
        assert bool(f & _abi.PF_TYPE_ERROR) == terr, text
        if not terr:
            assert bool(f & _abi.PF_CONTENT) == appended, text
        if "usage" in doc:
            want = token_usage(doc)
            if rec.exotic:
                n_exotic += 1
                continue
            assert not any(_unrepresentable(v) for k, v in want.items() if k not in ("model", "provider")) , text
            got = _abi.usage_rec_to_dict(rec)
            assert json.dumps(got, sort_keys=True) == json.dumps(want, sort_keys=True), text
            n_usage += 1
        n_checked += 1
    assert n_valid > 15000 and n_usage > 3000
    assert n_exotic < 0.3 * n_valid      # the generator aims at odd shapes on purpose


def test_lean_recogniser_on_golden_events():
    """Every complete event of every golden stream: lean flags == full-machine flags (subset)."""
    lean_mask = _abi.PF_VALID_A | _abi.PF_VALID_B | _abi.TK_ERROR | _abi.TK_DETAIL | _abi.TK_CODE | _abi.TK_USAGE | _abi.PF_TOO_DEEP
    n = 0
    for case in CASES:
        blob = b"".join(case["chunks"])
        for part in blob.split(b"\n\n"):
            f, cls, _ = hm.parse_part(part)
            if cls == 0:
                continue
            assert hm.lean_parse(part) == (f & lean_mask), part
            n += 1
    assert n > 500


def test_utf8_validator_matches_cpython():
    rng = random.Random(7)
    pool = [b"a", b"\xc3\xa9", b"\xe4\xb8\xad", b"\xf0\x9f\x98\x80", b"\xc0\xaf", b"\xed\xa0\x80", b"\xf4\x90\x80\x80",
            b"\xff", b"\x80", b"\xe0\x80\x80", b"\xf0\x80\x80\x80", b"\xc2", b"\xe2\x82", b"\xf0\x9f\x98", b"\xed\x9f\xbf", b"\xf4\x8f\xbf\xbf", b"\xef\xbf\xbd"]
    for _ in range(20000):
        b = b"".join(rng.choice(pool) for _ in range(rng.randrange(0, 6)))
        if rng.random() < 0.3 and b:
            b = b[:rng.randrange(len(b))]
        try:
            b.decode("utf-8"); ok = True
        except UnicodeDecodeError:
            ok = False
        assert hm.utf8_valid(b) == ok, b


def _bits(x: float) -> int:
    return struct.unpack("<Q", struct.pack("<d", x))[0]


def test_decimal_to_double_matches_cpython():
    rng = random.Random(99)
    for it in range(60000):
        nd = rng.randrange(1, 20)
        man = rng.randrange(10 ** (nd - 1), 10 ** nd)
        e = rng.randrange(-345, 310)
        ok, got = hm.dec_to_double(man, e)
        assert ok
        want = float("%de%d" % (man, e))
        assert got == _bits(want), (man, e)
    for it in range(40000):   # shortest-repr doubles incl. subnormals
        d = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64) & 0x7FEFFFFFFFFFFFFF))[0]
        if rng.random() < 0.2:
            d = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(52)))[0]
        s = repr(d)
        m, _, ex = s.partition("e")
        e = int(ex) if ex else 0
        if "." in m:
            a, b = m.split("."); e -= len(b); m = a + b
        ok, got = hm.dec_to_double(int(m), e)
        assert ok and got == _bits(d), s
    for man, e in [(0, 0), (1, 400), (1, -400), (49, -325), (24703282292062327, -340), (24703282292062328, -340),
                   (17976931348623157, 292), (17976931348623158, 292), (17976931348623159, 292), (2225073858507201, -323)]:
        ok, got = hm.dec_to_double(man, e)
        want = float("%de%d" % (man, e))
        assert ok and got == _bits(want), (man, e)
