"""Request-body rewrite (rows a1-a4) on the CPU: the device machine (host build, test aid) against the
oracle and the reference-generated goldens; the host-side plan compiler against the oracle's rule_ops."""
import base64
import json
import random

import numpy as np
import pytest

from llmapigateway_b200 import rewrite as rw
from oracle import body_oracle as bo
import body_cases as bc
import host_machine as hm
from golden_io import GOLDEN as GOLDEN_DIR

MODE_NAMES = ["httpx028", "httpx027", "json5"]


@pytest.fixture(scope="module")
def plans028():
    return rw.RulePlans(bc.RULES, fallback_provider="fb", stream_mode="httpx028")


@pytest.fixture(scope="module")
def plans027():
    return rw.RulePlans(bc.RULES, fallback_provider="fb", stream_mode="httpx027")


def test_golden_attempt_bytes(plans028):
    """every attempt the reference's chat_completions made (payload bytes by the real httpx) == the machine's output"""
    doc = json.loads((GOLDEN_DIR / "body_cases.json").read_text())
    assert doc["rules"]["gw/chain"]["fallback_models"] == bc.RULES["gw/chain"]["fallback_models"][:4]
    plans, ops, blob = plans028.packed()
    n = 0
    for case in doc["cases"]:
        raw = base64.b64decode(case["body"])
        model = json.loads(raw)["model"]
        seq = bc.CHAIN_ATTEMPTS if model == "gw/chain" else [None]
        assert len(seq) == len(case["attempts"])
        for att, a in zip(seq, case["attempts"]):
            want = base64.b64decode(a["httpx_bytes"])
            pi = plans028.plan_index(model, *att) if att else plans028.plan_index(model)
            st, out, need = hm.rewrite_body(raw, plans, ops, blob, pi)
            assert st == rw.BODY_OK and out == want, (model, att, out, want)
            n += 1
    assert n == 37


def test_plan_compiler_matches_oracle_rule_ops(plans028):
    for ri, sub, retry in bc.CHAIN_ATTEMPTS + [(4, -1, False)]:
        rule, prov, sp, rt = bc.attempt_for_oracle(ri, sub, retry)
        assert rw.attempt_assignments(rule, prov, sp, rt) == bo.rule_ops(rule, prov, sp, rt)


def _check(raw, plans_obj, packed, att, mode_name, stream):
    plans, ops, blob = packed
    ri, sub, retry = att
    rule, prov, sp, rt = bc.attempt_for_oracle(ri, sub, retry)
    pi = plans_obj.plan_index("gw/chain", ri, sub, retry, stream=stream)
    assert MODE_NAMES[int(plans[pi]["mode"])] == mode_name
    st, out, need = hm.rewrite_body(raw, plans, ops, blob, pi)
    ost, body, _, _ = bo.parse_body(raw)
    if ost == 1:
        # the rewrite entry point reports malformed bodies; a missing "model" is lgw_bodies_scan's business
        try:
            json.loads(raw.decode("utf-8"))
            is_obj = isinstance(json.loads(raw.decode("utf-8")), dict)
        except Exception:
            assert st == rw.BODY_PARSE_ERROR, raw
            return "parse_error"
        if not is_obj:
            assert st in (rw.BODY_OK, rw.BODY_EXOTIC, rw.BODY_ENCODE_ERROR)     # non-object roots are filtered by the scan
            return "non_object"
        body = json.loads(raw.decode("utf-8"))
    payload = bo.rewrite_payload(body, bo.rule_ops(rule, prov, sp, rt))
    try:
        want = bo.RENDERERS[mode_name](payload)
    except (ValueError, UnicodeEncodeError):
        assert st == rw.BODY_ENCODE_ERROR, (raw, st)
        return "encode_error"
    if st == rw.BODY_EXOTIC:
        assert _may_be_exotic(payload, mode_name), (mode_name, raw)
        return "exotic"
    assert st == rw.BODY_OK, (raw, st)
    assert out == want, (mode_name, raw, out, want)
    return "ok"


def _may_be_exotic(v, mode_name):
    """shapes the engine is allowed to hand back: json5 keys whose identifier test reaches a
    non-ASCII character, floats outside 1e-290..1e290"""
    if isinstance(v, dict):
        for k, x in v.items():
            if mode_name == "json5":
                for i, ch in enumerate(k):
                    if ch.isascii() and (ch.isalpha() or ch in "_$" or (i > 0 and ch.isdigit())):
                        continue
                    if not ch.isascii():
                        return True
                    break
            if _may_be_exotic(x, mode_name):
                return True
        return False
    if isinstance(v, list):
        return any(_may_be_exotic(x, mode_name) for x in v)
    if isinstance(v, float):
        return v != 0 and not (1e-289 < abs(v) < 1e289)
    return False


def test_fuzz_against_oracle(plans028, plans027):
    rng = random.Random(20260921)
    tally = {}
    packed = {"httpx028": plans028.packed(), "httpx027": plans027.packed()}
    atts = bc.CHAIN_ATTEMPTS + [(4, -1, False)]
    for it in range(4000):
        body = bc.rand_body(rng)
        raw = bc.spell(rng, body).encode("utf-8")
        att = atts[it % len(atts)]
        for mode_name, obj, stream in (("httpx028", plans028, True), ("httpx027", plans027, True), ("json5", plans028, False)):
            r = _check(raw, obj, packed["httpx027" if obj is plans027 else "httpx028"], att, mode_name, stream)
            tally[r] = tally.get(r, 0) + 1
    assert tally.get("ok", 0) > 9000, tally
    assert tally.get("exotic", 0) < 2500, tally


@pytest.mark.parametrize("text,status", [
    (b'{"model":"gw/chain","a":1,"a":2}', rw.BODY_EXOTIC),                     # duplicate key (dict keeps first position, last value)
    (b'{"model":"gw/chain","x":{"k":1,"k":1}}', rw.BODY_EXOTIC),
    (b'{"model":"gw/chain","t":0.1234567890123456789}', rw.BODY_EXOTIC),       # repr() needs the correctly rounded 17-digit form
    (b'{"model":"gw/chain","t":1e999}', rw.BODY_EXOTIC),
    (b'{"model":"gw/chain","t":NaN}', rw.BODY_ENCODE_ERROR),
    (b'{"model":"gw/chain","t":-Infinity}', rw.BODY_ENCODE_ERROR),
    (b'{"model":"gw/chain","t":"\\ud800"}', rw.BODY_ENCODE_ERROR),
    (b'{"model":"gw/chain","t":"\\ud800\\u0041"}', rw.BODY_ENCODE_ERROR),
    (b'{"model":"gw/chain",}', rw.BODY_PARSE_ERROR),
    (b'{"model":"gw/chain"} x', rw.BODY_PARSE_ERROR),
    (b'{"model":"gw/chain","t":01}', rw.BODY_PARSE_ERROR),
    (b'{"model":"gw/chain","t":"\xff"}', rw.BODY_PARSE_ERROR),
    (b'{"model":"gw/chain","t":"a\nb"}', rw.BODY_PARSE_ERROR),
    (b'', rw.BODY_PARSE_ERROR),
    (b'{"model":"gw/chain","t":[1,2}', rw.BODY_PARSE_ERROR),
    (b"{'model':'gw/chain'}", rw.BODY_PARSE_ERROR),                             # JSON5-only syntax: unpinned, reported as a parse error
])
def test_statuses(plans028, text, status):
    plans, ops, blob = plans028.packed()
    st, out, need = hm.rewrite_body(text, plans, ops, blob, plans028.plan_index("gw/chain", 0))
    assert st == status


def test_nan_and_lone_surrogates_in_ascii_modes(plans027, plans028):
    raw = b'{"model":"gw/chain","t":[NaN,Infinity,-Infinity,"\\ud800","\\udc00\\ud83d\\ude00\\ud83d"]}'
    for obj, stream, mode in ((plans027, True, "httpx027"), (plans028, False, "json5")):
        assert _check(raw, obj, obj.packed(), (0, -1, False), mode, stream) == "ok"


def test_overflow_reports_needed_length(plans028):
    plans, ops, blob = plans028.packed()
    raw = json.dumps({"model": "gw/chain", "messages": [{"role": "user", "content": "x" * 500}]}).encode()
    pi = plans028.plan_index("gw/chain", 1)
    st, out, need = hm.rewrite_body(raw, plans, ops, blob, pi, cap=1 << 12)
    assert st == rw.BODY_OK and need == len(out)
    st2, out2, need2 = hm.rewrite_body(raw, plans, ops, blob, pi, cap=100)
    assert st2 == rw.BODY_OVERFLOW and need2 == need and out2 == out[:100]


def test_scan_matches_parse_body():
    rng = random.Random(7)
    cases = [b'{"model":"m","stream":true}', b'{"stream":1}', b'[1]', b'"s"', b'3', b'{"model":""}', b'{"model":null,"stream":"yes"}',
             b'{"model":0.0}', b'{"model":[],"stream":[0]}', b'{"model":{"a":1}}', b'{"model":"\\u00e9\\ud83d\\ude00 x","stream":0}', b'{"model":"m"',
             b'\xff', b'{"model":"m","stream":false} ', b'  {"x":{"model":""},"model":"deep ok","stream":{}}', b'{"model":12,"stream":null}',
             b'{"model":false}', b'{"model":"a","model2":"b"}', b'null', b'{}']
    for _ in range(1500):
        cases.append(bc.spell(rng, bc.rand_body(rng)).encode("utf-8"))
    seen = set()
    for raw in cases:
        sc, model = hm.scan_body(raw)
        ost, body, omodel, ostream = bo.parse_body(raw)
        try:
            dup = len(json.loads(raw, object_pairs_hook=lambda p: p if len({k for k, _ in p}) == len(p) else (_ for _ in ()).throw(KeyError()))) < 0
        except KeyError:
            continue                                            # duplicate keys: last-wins, not modelled by the scan
        except Exception:
            pass
        assert int(sc["status"]) == ost, (raw, sc)
        seen.add(ost)
        if ost == 1:
            continue
        assert bool(sc["model_truthy"]) == bool(omodel)
        assert bool(sc["stream_truthy"]) == bool(ostream), raw
        if isinstance(omodel, str):
            assert model == omodel.encode("utf-8", "surrogatepass")[:256]
    assert seen == {0, 1, 2}


# ---- the data-parallel path (body_fast.cuh) against the sequential machine ----------------------------
FAST_IMPL = {"chunks": hm.rewrite_body_fast}


def _fast_vs_sequential(corpus, objs, atts, tally, impl="chunks"):
    fast_fn = FAST_IMPL[impl]
    for it, raw in enumerate(corpus):
        att = atts[it % len(atts)]
        for obj, stream in objs:
            plans, ops, blob = obj.packed()
            pi = obj.plan_index("gw/chain", *att, stream=stream)
            fst, fout, fneed = fast_fn(raw, plans, ops, blob, pi)
            if fst == hm.FAST_IRREGULAR:
                tally["irregular"] = tally.get("irregular", 0) + 1
                continue
            st, out, need = hm.rewrite_body(raw, plans, ops, blob, pi)
            assert (fst, fout, fneed) == (st, out, need), (raw, att, stream)
            tally["fast"] = tally.get("fast", 0) + 1


@pytest.mark.parametrize("impl", ["chunks"])
def test_fast_path_equals_sequential_machine(plans028, plans027, impl):
    """contract of body_fast.cuh: whenever it does not say "irregular", status, length and bytes are the
    sequential machine's (which the tests above pin to the oracle)"""
    from llmapigateway_b200.synth import chat_bodies
    rng = random.Random(4242)
    atts = bc.CHAIN_ATTEMPTS + [(4, -1, False)]
    objs = ((plans028, True), (plans027, True), (plans028, False))
    tally = {}
    plain = [bc.spell(rng, bc.rand_body(rng), plain_keys=True).encode("utf-8") for _ in range(2500)]
    _fast_vs_sequential(plain, objs, atts, tally, impl)
    assert tally["fast"] > 4000, tally                      # the fast path really is exercised (json5 + non-ASCII keys fall back)
    escaped = [bc.spell(rng, bc.rand_body(rng)).encode("utf-8") for _ in range(800)]
    _fast_vs_sequential(escaped, objs, atts, tally, impl)
    chat = chat_bodies(48, 4096, seed=2) + chat_bodies(48, 300, seed=3, non_ascii=0.3) + chat_bodies(16, 6000, seed=4) + chat_bodies(4, 9000, seed=5)
    t2 = {}
    _fast_vs_sequential(chat, objs, atts, t2, impl)
    assert t2["fast"] >= 3 * (48 + 48 + 16) - 10 and t2.get("irregular", 0) >= 12, t2      # > 6 KiB bodies take the sequential machine
    # small output slots: same OVERFLOW verdict and needed length
    plans, ops, blob = plans028.packed()
    for raw in chat[:20]:
        pi = plans028.plan_index("gw/chain", 1)
        assert FAST_IMPL[impl](raw, plans, ops, blob, pi, cap=1000)[::2] == hm.rewrite_body(raw, plans, ops, blob, pi, cap=1000)[::2]


@pytest.mark.parametrize("impl", ["chunks"])
def test_fast_path_never_accepts_what_the_machine_rejects(plans028, plans027, impl):
    rng = random.Random(77)
    bad = [b'{"model":"gw/chain","a":1,"a":2}', b'{"model":"gw/chain","x":{"k":1,"k":1}}', b'{"model":"gw/chain","t":0.1234567890123456789}',
           b'{"model":"gw/chain","t":NaN}', b'{"model":"gw/chain","t":"\\ud800"}', b'{"model":"gw/chain",}', b'{"model":"gw/chain"} x',
           b'{"model":"gw/chain","t":01}', b'{"model":"gw/chain","t":"\xff"}', b'{"model":"gw/chain","t":"a\nb"}', b'', b'{"model":"gw/chain","t":[1,2}',
           b'{"a":1}}', b'{"a":1}{', b'[{"a":1}]', b'{"a":tru}', b'{"a":-}', b'{"a":1.}', b'{"a":.5}', b'{"a":1e}', b'{"a" 1}', b'{"a"::1}', b'{1:2}',
           b'{"a":"\\x"}', b'{"a":"\\u12G4"}', b'{"a":"x}', b'{"a":1,,"b":2}', b'{"a":[,1]}', b'{"a":[1 2]}', b'{"a":"b" "c":1}', b'{"a":1}\\', b'\\{"a":1}',
           b'{"a":\\"b"}', b'{"a":"\xc3"}', b'{"a":"\x80"}', b'{"a":"\xed\xa0\x80"}', b'{"a":"\xf4\x90\x80\x80"}', b'{"a":"\xc0\xaf"}', b'{"a":1}\xc3\xa9']
    # mutations of valid bodies: drop / flip / duplicate one byte
    for _ in range(1500):
        raw = bytearray(bc.spell(rng, bc.rand_body(rng), plain_keys=True).encode("utf-8"))
        k = rng.randrange(len(raw))
        op = rng.randrange(3)
        if op == 0:
            del raw[k]
        elif op == 1:
            raw[k] = rng.choice(b'{}[]",:\\ 0a\x80\xe2')
        else:
            raw.insert(k, raw[k])
        bad.append(bytes(raw))
    tally = {}
    _fast_vs_sequential(bad, ((plans028, True), (plans027, True), (plans028, False)), [(1, -1, False), (3, -1, True)], tally, impl)
    assert tally.get("irregular", 0) > 500


# ---- row a12: non-streaming responses ---------------------------------------------------------------
@pytest.mark.parametrize("fast", [False, True])
def test_response_goldens(plans028, fast):
    """status + body bytes of an upstream response -> what the unmodified make_llm_request + chat.py:146 +
    Starlette made of it (tests/golden/response_cases.json); the host build stands in for the GPU"""
    from fake_body_engine import HostBodyEngine
    from llmapigateway_b200.responses import ExoticResponse, normalise_responses
    from oracle import response_oracle as ro
    doc = json.loads((GOLDEN_DIR / "response_cases.json").read_text())
    eng = HostBodyEngine(plans028, fast=fast)
    url = "http://upstream.test/v1/chat/completions"
    kinds = set()
    for c in doc["cases"]:
        content = base64.b64decode(c["content"])
        okind, oval = ro.normalise(c["status"], content, url)
        assert okind == c["kind"] and (okind != "ok" or oval == base64.b64decode(c["body"]))        # the oracle is pinned
        try:
            body, detail = normalise_responses(eng, plans028, [content], [c["status"]], url)[0]
        except ExoticResponse:
            # documented hand-backs: the reference's renderer raises; a float outside 1e-290..1e290
            try:
                root = json.loads(content)
            except Exception:
                root = None
            assert c["kind"] == "raise" or _may_be_exotic(root, "httpx028"), content
            kinds.add("raise")
            continue
        kinds.add(c["kind"])
        if c["kind"] == "ok":
            assert detail is None and body == base64.b64decode(c["body"])
        else:
            assert c["kind"] == "fail" and body is None
            if not (c["detail"] or "").startswith("Unexpected error during request") or "has no attribute" in c["detail"]:
                assert detail == c["detail"]
    assert kinds == {"ok", "fail", "raise"}


def test_response_fuzz_against_oracle(plans028):
    from fake_body_engine import HostBodyEngine
    from llmapigateway_b200.responses import normalise_responses
    from oracle import response_oracle as ro
    rng = random.Random(2718)
    eng = HostBodyEngine(plans028, fast=True)
    n_ok = 0
    for it in range(1500):
        doc = bc.rand_body(rng)
        if it % 7 == 0:
            doc[rng.choice(["error", "detail"])] = rng.choice([{"message": "m"}, "text", None, {"code": 1}, 5])
        raw = bc.spell(rng, doc, plain_keys=(it % 2 == 0)).encode("utf-8")
        status = rng.choice([200, 200, 200, 201, 404, 500])
        kind, val = ro.normalise(status, raw, "u")
        body, detail = normalise_responses(eng, plans028, [raw], [status], "u", strict=False)[0]
        if body == "exotic":
            continue
        if kind == "ok":
            assert (body, detail) == (val, None)
            n_ok += 1
        else:
            assert kind == "fail" and body is None and detail == val, (raw, detail, val)
    assert n_ok > 500
    # roots that are not objects: `"error" in x` is list membership / substring / TypeError (request_handler.py:167)
    n_ok = 0
    for it in range(600):
        t = it % 6
        if t == 0:
            doc = [rng.choice(["error", "detail", "errors", "x", 1, None, {"error": 1}, ["detail"], "d\u00e9tail"]) for _ in range(rng.randint(0, 5))]
        elif t == 1:
            doc = rng.choice(["", "an error occurred", "no problem", "detaildetail", "deta il", "err or", "xerrorx", "\u00e9rror", "detai", "e"]) * rng.randint(1, 2)
        elif t == 2:
            doc = rng.choice([0, 12, -1.5, None, True, False])
        elif t == 3:
            doc = [[bc.rand_value(rng) for _ in range(3)], "error " if it % 4 else "error"]
        else:
            doc = [bc.rand_value(rng) for _ in range(rng.randint(0, 4))]
        raw = bc.spell(rng, doc).encode("utf-8")
        kind, val = ro.normalise(200, raw, "u")
        body, detail = normalise_responses(eng, plans028, [raw], [200], "u", strict=False)[0]
        if body == "exotic":
            assert _may_be_exotic(doc, "httpx028") or kind == "raise", raw
            continue
        if kind == "ok":
            assert (body, detail) == (val, None), raw
            n_ok += 1
        else:
            assert kind == "fail" and body is None and detail == val, (raw, detail, val)
    assert n_ok > 100


def test_scan_matches_reference_parse_goldens():
    """chat.py:31-45 driven for real (tests/golden/make_body_golden.py `parse`): which bodies are answered 400 and with which
    message, and what `stream` means for the first attempt"""
    doc = json.loads((GOLDEN_DIR / "body_cases.json").read_text())
    seen = set()
    for p in doc["parse"]:
        raw = base64.b64decode(p["body"])
        sc, model = hm.scan_body(raw)
        ost = bo.parse_body(raw)[0]
        want = 1 if p["detail_head"].startswith("Error reading") else 2 if p["detail_head"].startswith("Missing 'model") else 0
        assert ost == want, raw                                   # the oracle is pinned
        assert int(sc["status"]) == want, (raw, sc)
        seen.add(want)
        if want == 0 and p["is_streaming"] is not None:
            assert bool(sc["stream_truthy"]) == p["is_streaming"], raw
    assert seen == {0, 1, 2}


def test_error_detail_walk_matches_cpython(plans028):
    """csrc/error_detail.cuh (host build) against CPython's own evaluation of request_handler.py:167-169 on documents whose
    "error" / "detail" values take every shape: escapes, surrogate pairs, escaped key names, whitespace, nested containers,
    numbers of every spelling, duplicate keys (the last one wins in a dict)."""
    from fake_body_engine import HostBodyEngine
    from llmapigateway_b200.responses import _error_detail
    from llmapigateway_b200 import rewrite as rw
    rng = random.Random(99)
    eng = HostBodyEngine(plans028, fast=True)
    n_text = n_exotic = 0
    for raw in bc.error_detail_docs(rng, 3000):
        doc = json.loads(raw)
        if not ("error" in doc or "detail" in doc):
            continue
        try:
            want = doc.get("error", {}).get("message") or doc.get("detail")
        except Exception as e:
            want = f"Unexpected error during request to u: {str(e)}"
        got, exotic = _error_detail(eng, raw, "u", rw.KIND_OBJ)
        if exotic is not None:
            assert isinstance(want, (dict, list)) or (isinstance(want, str) and any(0xD800 <= ord(ch) < 0xE000 for ch in want)), (raw, want, exotic)
            n_exotic += 1
            continue
        assert type(got) is type(want) and got == want, (raw, got, want)
        n_text += isinstance(want, str)
    assert n_text > 800 and n_exotic > 50


def test_nonstream_response_tap_host_machine():
    """The document tap (row a8, non-streaming mode) on the CPU box: the body of the GPU test (tests/test_body_gpu.py) over the host
    build of the same machine calls k_docs_usage makes -- the 20 + 9 reference-generated tap cases and 1 500 fuzzed documents
    against the oracle."""
    import test_body_gpu as G
    from fake_engine import FakeEngine
    G.test_nonstream_response_tap(FakeEngine(max_streams=2))
