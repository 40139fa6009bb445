"""GPU: request-body rewrite kernels (rows a1-a4) through the C ABI, bit-exact against the oracle and
the reference-generated goldens."""
import base64
import json
import random

import numpy as np
import pytest

import body_cases as bc
from golden_io import GOLDEN
from llmapigateway_b200 import rewrite as rw
from llmapigateway_b200.synth import chat_bodies
from oracle import body_oracle as bo

pytestmark = pytest.mark.gpu
MODE_NAMES = ["httpx028", "httpx027", "json5"]


@pytest.fixture(scope="module")
def engine():
    import llmapigateway_b200 as L
    e = L.Engine(max_streams=64, max_step_chunks=1024, max_step_bytes=1 << 20)
    yield e
    e.close_engine()


def _want(raw, att, mode_name):
    """(status, bytes) the reference path produces for one attempt; status None = engine may call it exotic"""
    ri, sub, retry = att
    rule, prov, sp, rt = bc.attempt_for_oracle(ri, sub, retry)
    try:
        body = json.loads(raw.decode("utf-8"))
    except Exception:
        return rw.BODY_PARSE_ERROR, b""
    if not isinstance(body, dict):
        return None, b""
    payload = bo.rewrite_payload(body, bo.rule_ops(rule, prov, sp, rt))
    try:
        return rw.BODY_OK, bo.RENDERERS[mode_name](payload)
    except (ValueError, UnicodeEncodeError):
        return rw.BODY_ENCODE_ERROR, b""


def test_golden_attempt_bytes(engine):
    plans = rw.RulePlans(bc.RULES, fallback_provider="fb", stream_mode="httpx028")
    engine.load_rules(plans)
    doc = json.loads((GOLDEN / "body_cases.json").read_text())
    bodies, idx, want = [], [], []
    for case in doc["cases"]:
        raw = base64.b64decode(case["body"])
        model = json.loads(raw)["model"]
        seq = bc.CHAIN_ATTEMPTS if model == "gw/chain" else [None]
        for att, a in zip(seq, case["attempts"]):
            bodies.append(raw)
            idx.append(plans.plan_index(model, *att) if att else plans.plan_index(model))
            want.append(base64.b64decode(a["httpx_bytes"]))
    got = engine.rewrite_bodies(bodies, idx)
    assert len(got) == 37
    for (st, out), w in zip(got, want):
        assert st == rw.BODY_OK and out == w


@pytest.mark.parametrize("stream_mode", ["httpx028", "httpx027"])
def test_fuzz_batch_against_oracle(engine, stream_mode):
    plans = rw.RulePlans(bc.RULES, fallback_provider="fb", stream_mode=stream_mode)
    engine.load_rules(plans)
    rng = random.Random(99 if stream_mode == "httpx028" else 100)
    atts = bc.CHAIN_ATTEMPTS + [(4, -1, False)]
    bodies, idx, meta = [], [], []
    for it in range(3000):
        raw = bc.spell(rng, bc.rand_body(rng)).encode("utf-8")
        if it % 97 == 0:
            raw = raw[:-1]                                 # truncated JSON
        if it % 101 == 0:
            raw = raw.replace(b"a", b"\xff", 1)            # invalid UTF-8
        att = atts[it % len(atts)]
        stream = it % 3 != 0
        bodies.append(raw)
        idx.append(plans.plan_index("gw/chain", *att, stream=stream))
        meta.append((att, stream_mode if stream else "json5"))
    got = engine.rewrite_bodies(bodies, idx)
    tally = {}
    for raw, (att, mode_name), (st, out) in zip(bodies, meta, got):
        wst, w = _want(raw, att, mode_name)
        if st == rw.BODY_EXOTIC or wst is None:
            tally["exotic"] = tally.get("exotic", 0) + 1
            continue
        assert st == wst, (raw, st, wst)
        assert out == w, (mode_name, raw)
        tally[st] = tally.get(st, 0) + 1
    assert tally[rw.BODY_OK] > 2000 and tally.get(rw.BODY_PARSE_ERROR, 0) > 20, tally


def test_config2_bodies_bit_exact(engine):
    """BASELINE.json configs[1]: 1024 bodies x 4 KiB, every body checked against the oracle in every mode"""
    plans = rw.RulePlans(bc.RULES, fallback_provider="fb", stream_mode="httpx028")
    engine.load_rules(plans)
    bodies = chat_bodies(1024, 4096, seed=2)
    for att, stream in (((1, -1, False), True), ((3, -1, True), True), ((4, -1, False), False), ((2, 1, False), False)):
        idx = [plans.plan_index("gw/chain", *att, stream=stream)] * len(bodies)
        got = engine.rewrite_bodies(bodies, idx)
        mode_name = "httpx028" if stream else "json5"
        for raw, (st, out) in zip(bodies, got):
            wst, w = _want(raw, att, mode_name)
            assert st == wst == rw.BODY_OK and out == w
    # unknown model -> fallback-provider plan: re-rendered, nothing assigned
    other = chat_bodies(64, 1024, seed=5, model="not-in-rules")
    got = engine.rewrite_bodies(other, [plans.plan_index("not-in-rules")] * len(other))
    for raw, (st, out) in zip(other, got):
        assert st == rw.BODY_OK and out == bo.render_httpx028(json.loads(raw))


def test_packed_offsets_and_overflow(engine):
    plans = rw.RulePlans(bc.RULES, fallback_provider="fb", stream_mode="httpx028")
    engine.load_rules(plans)
    bodies = chat_bodies(37, 600, seed=8) + [b"{bad", b'{"model":"gw/chain","a":1,"a":2}']
    buf, off = rw.pack_bodies(bodies)
    idx = np.full(len(bodies), plans.plan_index("gw/chain", 1), dtype=np.uint32)
    out, out_off, res = engine.rewrite_packed(buf, off, idx, slot_cap=2048)
    assert list(res["status"][-2:]) == [rw.BODY_PARSE_ERROR, rw.BODY_EXOTIC]
    assert out_off[0] == 0 and np.all(np.diff(out_off.astype(np.int64)) == np.where(res["status"] == 0, res["out_len"], 0))
    for i, raw in enumerate(bodies[:37]):
        assert bytes(out[int(out_off[i]):int(out_off[i + 1])]) == _want(raw, (1, -1, False), "httpx028")[1]
    # a slot too small for some bodies: those report OVERFLOW with the size they need, the rest are unaffected
    sizes = res["out_len"][:37]
    cap = int(np.sort(sizes)[18])
    out2, off2, res2 = engine.rewrite_packed(buf, off, idx, slot_cap=cap)
    for i in range(37):
        if sizes[i] > cap:
            assert res2["status"][i] == rw.BODY_OVERFLOW and res2["out_len"][i] == sizes[i] and off2[i] == off2[i + 1]
        else:
            assert res2["status"][i] == rw.BODY_OK and bytes(out2[int(off2[i]):int(off2[i + 1])]) == bytes(out[int(out_off[i]):int(out_off[i + 1])])


def test_scan_batch(engine):
    rng = random.Random(5)
    bodies = [b'{"model":"m","stream":true}', b'{"stream":1}', b'[1]', b'{"model":""}', b'{"model":"\\u00e9\\ud83d\\ude00 x","stream":0}', b'\xff', b'{}',
              b'{"model":"m"', b'{"model":12,"stream":null}']
    bodies += [bc.spell(rng, bc.rand_body(rng)).encode("utf-8") for _ in range(1000)]
    bodies += chat_bodies(64, 4096, seed=3, stream=True)
    scans, models = engine.scan_bodies(bodies)
    seen = set()
    for raw, sc, model in zip(bodies, scans, models):
        ost, body, omodel, ostream = bo.parse_body(raw)
        assert int(sc["status"]) == ost, raw
        seen.add(ost)
        if ost == 1:
            continue
        assert bool(sc["model_truthy"]) == bool(omodel) and bool(sc["stream_truthy"]) == bool(ostream)
        if isinstance(omodel, str):
            assert model == omodel.encode("utf-8", "surrogatepass")[:256]
    assert seen == {0, 1, 2}


def test_data_parallel_path_equals_exact_machine(engine):
    """mode 0 (block-per-body fast path + exact machine for what it calls irregular) == mode 1 (exact machine only)"""
    plans = rw.RulePlans(bc.RULES, fallback_provider="fb", stream_mode="httpx027")
    engine.load_rules(plans)
    rng = random.Random(31337)
    corpus = [bc.spell(rng, bc.rand_body(rng), plain_keys=True).encode("utf-8") for _ in range(1500)]
    corpus += [bc.spell(rng, bc.rand_body(rng)).encode("utf-8") for _ in range(300)]
    corpus += chat_bodies(128, 4096, seed=12) + chat_bodies(64, 300, seed=13, non_ascii=0.3) + chat_bodies(8, 9000, seed=14)
    for _ in range(600):                                            # damaged bodies
        raw = bytearray(bc.spell(rng, bc.rand_body(rng), plain_keys=True).encode("utf-8"))
        k = rng.randrange(len(raw))
        raw[k] = rng.choice(b'{}[]",:\\ 0a\x80\xe2')
        corpus.append(bytes(raw))
    atts = bc.CHAIN_ATTEMPTS + [(4, -1, False)]
    idx = [plans.plan_index("gw/chain", *atts[i % len(atts)], stream=(i % 2 == 0)) for i in range(len(corpus))]
    try:
        engine.set_mode(1)
        exact = engine.rewrite_bodies(corpus, idx)
        engine.set_mode(0)
        fast = engine.rewrite_bodies(corpus, idx)
    finally:
        engine.set_mode(0)
    assert fast == exact
    assert sum(1 for st, _ in fast if st == rw.BODY_OK) > 1500


def test_responses_normalise(engine):
    """row a12 through the engine: goldens from the unmodified reference + fuzz against the oracle"""
    from llmapigateway_b200.responses import normalise_responses
    from oracle import response_oracle as ro
    plans = rw.RulePlans(bc.RULES, fallback_provider="fb", stream_mode="httpx028")
    engine.load_rules(plans)
    doc = json.loads((GOLDEN / "response_cases.json").read_text())
    url = "http://upstream.test/v1/chat/completions"
    contents = [base64.b64decode(c["content"]) for c in doc["cases"]]
    got = normalise_responses(engine, plans, contents, [c["status"] for c in doc["cases"]], url, strict=False)
    for c, (body, detail) in zip(doc["cases"], got):
        if c["kind"] == "ok" and body != "exotic":
            assert (body, detail) == (base64.b64decode(c["body"]), None)
        elif c["kind"] == "ok":
            assert b"5e-324" in base64.b64decode(c["content"])
        elif c["kind"] == "raise":
            assert body == "exotic"
        else:
            assert body is None
    rng = random.Random(161)
    raws, sts = [], []
    for it in range(2000):
        d = bc.rand_body(rng)
        if it % 7 == 0:
            d[rng.choice(["error", "detail"])] = rng.choice([{"message": "m"}, "text", None, {"code": 1}, 5])
        raws.append(bc.spell(rng, d, plain_keys=(it % 2 == 0)).encode("utf-8"))
        sts.append(rng.choice([200, 200, 200, 201, 404, 500]))
    got = normalise_responses(engine, plans, raws, sts, "u", strict=False)
    n_ok = 0
    for raw, st, (body, detail) in zip(raws, sts, got):
        if body == "exotic":
            continue
        kind, val = ro.normalise(st, raw, "u")
        if kind == "ok":
            assert (body, detail) == (val, None)
            n_ok += 1
        else:
            assert body is None and detail == val
    assert n_ok > 600


def test_packed_output_smaller_than_needed(engine):
    """the caller's packed buffer is too small: bodies that still fit are intact, the rest report OVERFLOW"""
    plans = rw.RulePlans(bc.RULES, fallback_provider="fb", stream_mode="httpx028")
    engine.load_rules(plans)
    bodies = chat_bodies(40, 600, seed=21)
    buf, off = rw.pack_bodies(bodies)
    idx = np.full(len(bodies), plans.plan_index("gw/chain", 1), dtype=np.uint32)
    full, full_off, res = engine.rewrite_packed(buf, off, idx, slot_cap=2048)
    assert np.all(res["status"] == rw.BODY_OK)
    cut = int(full_off[25]) + 10                                  # room for 25 bodies and a bit
    small = np.zeros(cut, dtype=np.uint8)
    out2, off2, res2 = engine.rewrite_packed(buf, off, idx, slot_cap=2048, out=small)
    assert np.array_equal(off2, full_off)                          # offsets still say where everything would go
    for i in range(len(bodies)):
        if int(full_off[i + 1]) <= cut:
            assert res2["status"][i] == rw.BODY_OK and bytes(out2[int(off2[i]):int(off2[i + 1])]) == bytes(full[int(full_off[i]):int(full_off[i + 1])])
        else:
            assert res2["status"][i] == rw.BODY_OVERFLOW


def test_scan_reference_parse_goldens(engine):
    """chat.py:31-45 goldens (made by driving the unmodified endpoint) through lgw_bodies_scan"""
    doc = json.loads((GOLDEN / "body_cases.json").read_text())
    raws = [base64.b64decode(p["body"]) for p in doc["parse"]]
    scans, models = engine.scan_bodies(raws)
    for p, raw, sc in zip(doc["parse"], raws, scans):
        want = 1 if p["detail_head"].startswith("Error reading") else 2 if p["detail_head"].startswith("Missing 'model") else 0
        assert int(sc["status"]) == want, raw
        if want == 0 and p["is_streaming"] is not None:
            assert bool(sc["stream_truthy"]) == p["is_streaming"], raw


def test_nonstream_response_tap(engine):
    """Row a8, non-stream mode (chat_logging.py:98-150): usage rows of non-streaming responses through lgw_documents_usage against the
    rows the unmodified ChunkProcessorThread wrote (goldens) and against the oracle on fuzzed documents."""
    from golden_io import canon_rows
    from oracle.sse_oracle import tap_nonstream
    doc = json.loads((GOLDEN / "response_cases.json").read_text())
    texts, want = [], []
    for c in doc["tap_cases"]:
        if c["text"]:
            texts.append(base64.b64decode(c["text"])); want.append(c["rows"])
    for c in doc["cases"]:
        if "tap_rows" in c:
            texts.append(base64.b64decode(c["body"])); want.append(c["tap_rows"])
    got = engine.documents_usage(texts)
    n_cmp = 0
    for t, (rows, exotic), w in zip(texts, got, want):
        if exotic:
            continue
        assert canon_rows(rows) == canon_rows(w), t[:100]
        n_cmp += 1
    assert n_cmp >= len(texts) - 3
    rng = random.Random(99)
    fuzz = []
    for it in range(1500):
        d = bc.rand_body(rng)
        if it % 3 == 0:
            d["usage"] = rng.choice([{"prompt_tokens": rng.randrange(10**6), "completion_tokens": rng.randrange(10**5), "total_tokens": rng.randrange(10**6),
                                      "cost": rng.choice([0, 1.5e-5, 0.25, 3]), "completion_tokens_details": rng.choice([{"reasoning_tokens": rng.randrange(50)}, None, {}]),
                                      "prompt_tokens_details": {"cached_tokens": rng.randrange(9)}}, None, [], "x", {"prompt_tokens": None}])
        if it % 5 == 0:
            d["choices"] = rng.choice([[{"message": {"content": "hi"}}], [{"delta": None}], "str", [], [{"message": {"content": 5}}], None])
        if it % 11 == 0:
            d["error"] = rng.choice([None, {"message": "x"}])
        if it % 4 == 0:
            d["model"] = rng.choice(["m", "café", 5, None]); d["provider"] = "P"
        fuzz.append(bc.spell(rng, d, plain_keys=True).encode("utf-8"))
    got = engine.documents_usage(fuzz)
    n_ok = 0
    for t, (rows, exotic) in zip(fuzz, got):
        if exotic:
            continue
        assert canon_rows(rows) == canon_rows(tap_nonstream([t]).rows), t[:200]
        n_ok += 1
    assert n_ok > 1300


def test_error_detail_on_the_device(engine):
    """lgw_documents_error_detail through the C ABI against CPython's evaluation of request_handler.py:167-169."""
    from llmapigateway_b200.responses import _error_detail
    rng = random.Random(7)
    docs = list(bc.error_detail_docs(rng, 1500)) + [b'{"error":{"message":"' + b"z" * 7000 + b'"},"pad":"' + b"p" * 3000 + b'"}']
    docs = [d for d in docs if ("error" in json.loads(d) or "detail" in json.loads(d))]
    got = engine.documents_error_detail(docs, 8192)
    n_text = 0
    for raw, (e, text) in zip(docs, got):
        doc = json.loads(raw)
        try:
            want = doc.get("error", {}).get("message") or doc.get("detail")
        except Exception as ex:
            want = f"Unexpected error during request to u: {str(ex)}"

        class One:                       # feed the already computed device answer to the host-side mapping
            def documents_error_detail(self, _docs, text_stride=4096):
                return [(e, text)]

        mapped, exotic = _error_detail(One(), raw, "u", rw.KIND_OBJ)
        if exotic is not None:
            assert isinstance(want, (dict, list)) or (isinstance(want, str) and any(0xD800 <= ord(ch) < 0xE000 for ch in want)), (raw, want, exotic)
            continue
        assert type(mapped) is type(want) and mapped == want, (raw, mapped, want)
        n_text += isinstance(want, str)
    assert n_text > 400
