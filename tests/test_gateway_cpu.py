"""CPU host-logic tests: the batcher and the make_llm_request seam (against the golden fixtures made
with the unmodified reference), stream sharding, and a 2-rank gloo run of the sharded path."""
import asyncio
import json
import os
import subprocess
import sys
from pathlib import Path

import httpx
import pytest

from fake_engine import FakeEngine
from golden_io import UNPINNED_DETAIL_PREFIX, canon_rows, load_sse_cases
from llmapigateway_b200.gateway import StreamBatcher, make_llm_request, shard_of

CASES = load_sse_cases()
ROOT = Path(__file__).resolve().parent.parent


class _Sink:
    def __init__(self):
        self.rows = []

    def insert_usage(self, u):
        self.rows.append(u)


def _client_factory(chunks, status):
    class _Body(httpx.AsyncByteStream):
        async def __aiter__(self):
            for c in chunks:
                yield c

    def handler(request):
        return httpx.Response(status, headers={"content-type": "text/event-stream"}, stream=_Body())

    return lambda **kw: httpx.AsyncClient(transport=httpx.MockTransport(handler), **kw)


async def _drive(case, batcher, sink):
    n0 = len(sink.rows)
    resp, err = await make_llm_request("http://upstream.test/v1/chat/completions", {}, {"model": "m", "messages": []}, True,
                                       batcher=batcher, client_factory=_client_factory(case["chunks"], case["http_status"]))
    if resp is None:
        return dict(failed=True, error_detail=err, emitted=[], rows=[])
    out = [bytes(c) async for c in resp.body_iterator]
    return dict(failed=False, error_detail=err, emitted=out, rows=sink.rows[n0:])


def test_make_llm_request_seam_matches_reference():
    async def go():
        sink = _Sink()
        batcher = StreamBatcher(FakeEngine(max_streams=32), window_s=0.0005, usage_sink=sink)
        n_ok = 0
        for case in CASES[:120]:
            got = await _drive(case, batcher, sink)
            assert got["failed"] == case["failed"], case["name"]
            if case["failed"]:
                if case["error_detail"].startswith(UNPINNED_DETAIL_PREFIX):
                    assert got["error_detail"].startswith(UNPINNED_DETAIL_PREFIX)
                else:
                    assert got["error_detail"] == case["error_detail"], case["name"]
                continue
            assert got["emitted"] == case["emitted"], case["name"]
            try:
                assert canon_rows(got["rows"]) == case["rows"], case["name"]
                n_ok += 1
            except TypeError:
                pass        # an Unrepresentable value (reported exotic shape) is not JSON-serialisable
        assert n_ok > 80
    asyncio.run(go())


@pytest.mark.parametrize("relay_from", ["device", "host"])
def test_batcher_interleaves_many_streams(relay_from):
    """relay_from="host" (verdicts-only steps): the relayed chunks are the upstream's own chunk objects, chosen by the engine's
    per-segment results -- same bytes, same rows as with the downloaded re-emit."""
    async def go():
        sink = _Sink()
        batcher = StreamBatcher(FakeEngine(max_streams=64), window_s=0.001, usage_sink=sink, relay_from=relay_from)
        picks = [c for c in CASES if not c["failed"] and c["chunks"]][:40]
        results = await asyncio.gather(*[_drive(c, batcher, _Sink()) for c in picks])
        for c, r in zip(picks, results):
            assert r["emitted"] == c["emitted"], c["name"]
        assert batcher.steps < sum(len(c["chunks"]) for c in picks)      # chunks of different streams shared steps
    asyncio.run(go())


def test_shard_of_is_stable_and_balanced():
    assert [shard_of(i, 8) for i in range(5)] == [shard_of(str(i), 8) for i in range(5)]
    counts = [0] * 8
    for i in range(8192):
        counts[shard_of(i, 8)] += 1
    assert min(counts) > 900 and max(counts) < 1150
    assert shard_of("abc", 1) == 0


def test_two_rank_gloo_sharded_path():
    """world_size 2 on CPU (gloo): each rank serves the streams shard_of() gives it, partial rollup
    tables merge with all_reduce(SUM); rank 0 checks the union against the goldens."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29517", PYTHONPATH=f"{ROOT}:{ROOT / 'tests'}")
    procs = [subprocess.Popen([sys.executable, str(ROOT / "tests" / "gloo_worker.py"), str(r), "2"], env=env,
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GLOO_OK" in outs[0], outs[0]


def test_non_streaming_seam_matches_reference_goldens():
    """make_llm_request(..., is_streaming=False): engine-rewritten body bytes go upstream as they are; the upstream
    response comes back as the reference would have rendered it (tests/golden/response_cases.json)"""
    import base64
    import body_cases as bc
    from golden_io import GOLDEN
    from llmapigateway_b200 import rewrite as rw
    doc = json.loads((GOLDEN / "response_cases.json").read_text())
    plans = rw.RulePlans(bc.RULES, fallback_provider="fb")
    url = "http://upstream.test/v1/chat/completions"

    async def go():
        batcher = StreamBatcher(FakeEngine(max_streams=8), window_s=0.0005)
        batcher.load_rules(plans)
        raw = json.dumps({"model": "gw/chain", "messages": [{"role": "user", "content": "hi"}]}).encode()
        (st, payload), = await batcher.rewrite_bodies([raw], [plans.plan_index("gw/chain", 1, stream=False)])
        assert st == rw.BODY_OK and payload.startswith(b'{model: "m-or"')            # json5.dumps form (request_handler.py:153)
        seen = []
        for c in doc["cases"]:
            content = base64.b64decode(c["content"])

            def handler(request, content=content, status=c["status"]):
                seen.append(request.content)
                return httpx.Response(status, headers={"content-type": "application/json"}, content=content)

            resp, err = await make_llm_request(url, {"Content-Type": "application/json"}, payload, False, batcher=batcher,
                                               client_factory=lambda **kw: httpx.AsyncClient(transport=httpx.MockTransport(handler), **kw))
            try:
                root = json.loads(content)
            except Exception:
                root = None
            handed_back = c["kind"] == "raise" or b"5e-324" in content
            if handed_back:                                     # documented hand-backs (see test_body_cpu): the seam reports a failed attempt
                assert resp is None and (c["kind"] == "fail" or err.startswith("Unexpected error during request"))
            elif c["kind"] == "ok":
                assert err is None and bytes(resp.body) == base64.b64decode(c["body"])
            else:
                assert resp is None
                d = c["detail"]
                if not isinstance(d, str) or not d.startswith("Unexpected error during request") or "has no attribute" in d:
                    assert err == d
        assert seen and all(s == payload for s in seen)          # the bytes on the wire are the engine's bytes
    asyncio.run(go())


# ---- config 4: the fallback-chain walker (host logic over the fake engine; the GPU twin is tests/test_chain_gpu.py) -----------
def test_chain_walker_matches_the_reference_goldens():
    """llmapigateway_b200.chat.chat_completions against the goldens of the unmodified chat.py:20: relayed bytes, 503/400
    status + detail, url / wire body / headers of every upstream attempt (rotation, retries with the log scrub, sub-providers)."""
    import chain_cases as cc
    from fake_engine import FakeEngine
    doc, ups = cc.load()
    cases = doc["cases"]
    got = cc.walk_product(lambda: FakeEngine(max_streams=8), cases, ups)
    for case, g in zip(cases, got):
        cc.check_against_golden(case, g)


@pytest.mark.parametrize("relay_from_host", [False, True])
def test_chain_batch_matches_the_oracle(relay_from_host):
    """ChainBatch (lock-step rounds) == the per-request oracle walk: served-by round, relayed bytes, 503 details, attempt count."""
    from fake_engine import FakeEngine
    from llmapigateway_b200 import chat, rewrite, synth
    from oracle import chain_oracle
    import chain_cases as cc
    n = 48
    providers, rules, fallback_provider = synth.chain_world()
    up = synth.ChainUpstream(n, 5, seed=4, p_fail=0.4)
    bodies = synth.chain_request_bodies(n, seed=4)
    bodies[7] = b'{"messages":[]}'
    bodies[9] = b'{"model":"gw/rotating","stream":true,"messages":[]}'
    bodies[11] = b'{"model":"gw/rotating","stream":true,"messages":[]}'
    bodies[13] = b'{"model":"gw/retrying","stream":true,"messages":[{"role":"user","content":"x"}]}'
    eng = FakeEngine(max_streams=n)
    plans = rewrite.RulePlans(rules, fallback_provider=fallback_provider, stream_mode=cc.stream_mode())
    eng.load_rules(plans)
    out = chat.ChainBatch(eng, plans, providers, rules, relay_from_host=relay_from_host).run(bodies, ["k"] * n, up)
    rot = chain_oracle.Rotation()
    n_attempts = 0
    for i in range(n):
        want = chain_oracle.walk(bodies[i], {"Authorization": "Bearer k"}, providers, rules, fallback_provider, lambda a: up.stream_chunks(i, a), rot, cc.stream_mode())
        n_attempts += len(want["attempts"])
        if want["kind"] == "stream":
            assert out.served_round[i] == len(want["attempts"]) - 1, i
            assert out.emitted(i) == want["emitted"], i
            assert out.detail[i] is None
        else:
            assert out.served_round[i] < 0 and out.detail[i] == want["detail"], (i, out.detail[i], want["detail"])
    assert out.attempts == n_attempts
    rows = dict(out.usage_rows())
    assert set(rows) == {i for i in range(n) if out.served_round[i] >= 0}


# ---- cancellation, hostile usage values, raising sinks ------------------------------------------------------
def test_cancelled_feed_never_reaches_the_next_user_of_the_slot():
    """A client disconnect cancels the coroutine awaiting feed(); its queued chunk must not be packed into the stream that
    gets the slot next (it would corrupt that stream's priming verdict), and slots are reused first-in first-out."""
    async def go():
        eng = FakeEngine(max_streams=2)
        seen = []
        real_step = eng.step

        def spy(data, chunk_off, seg_chunk, seg_slot, out=None):
            seen.append([(int(s), bytes(data[int(chunk_off[int(seg_chunk[k])]):int(chunk_off[int(seg_chunk[k + 1])])])) for k, s in enumerate(seg_slot)])
            return real_step(data, chunk_off, seg_chunk, seg_slot)

        eng.step = spy
        b = StreamBatcher(eng, window_s=0.02)
        a = await b.open_stream(200)
        t = asyncio.ensure_future(b.feed(a, b'data: {"error":"poison"}\n\n'))
        await asyncio.sleep(0)                       # queued, pump sleeping in its window
        t.cancel()
        await b.close_stream(a)                      # request torn down before the step ran
        other = await b.open_stream(200)
        assert other != a                            # FIFO free list: the slot just released is the last to be reused
        again = await b.open_stream(200)
        assert again == a
        r = await b.feed(again, b'data: {"choices":[{"delta":{"content":"hi"}}]}\n\n')
        assert r.emitted is not None and r.phase != 3
        assert all(b"poison" not in blob for step in seen for _, blob in step)
        assert b.dropped_stale >= 0
        await b.close_stream(again); await b.close_stream(other)
    asyncio.run(go())


def test_stale_chunk_queued_for_a_closed_slot_is_dropped():
    async def go():
        eng = FakeEngine(max_streams=1)
        b = StreamBatcher(eng, window_s=0.02)
        a = await b.open_stream(200)
        t = asyncio.ensure_future(b.feed(a, b'data: {"detail":"stale"}\n\n'))
        await asyncio.sleep(0)
        await b.close_stream(a)                      # purges the queue entry of the slot
        with pytest.raises(asyncio.CancelledError):
            await t
        a2 = await b.open_stream(200)
        r = await b.feed(a2, b'data: {"choices":[]}\n\n')
        assert r.phase != 3                           # not FAILED: the stale error event never reached this stream
        await b.close_stream(a2)
    asyncio.run(go())


def test_raising_sink_and_hostile_usage_values_do_not_stall_streams():
    class Bad:
        def insert_usage(self, u):
            raise RuntimeError("sink down")

    async def go():
        batcher = StreamBatcher(FakeEngine(max_streams=8), window_s=0.0005, usage_sink=Bad())
        case = next(c for c in CASES if not c["failed"] and c["rows"])
        got = await asyncio.wait_for(_drive(case, batcher, _Sink()), 20)
        assert got["emitted"] == case["emitted"]
    asyncio.run(go())


def test_usage_table_rejects_what_it_cannot_hold_and_stays_usable():
    from llmapigateway_b200.usage import UsageTable
    t = UsageTable.__new__(UsageTable)
    UsageTable.__init__(t, engine=None) if False else None
    import numpy as np
    t.eng = None; t._lib = None
    t._host = {c: np.zeros(0, dtype=d) for c, d in zip(UsageTable.COLS, UsageTable.DTYPES)}
    t._models, t._providers, t._dev, t._pending_rows, t.rejected = [], [], None, [], 0
    assert t.insert_usage({"prompt_tokens": 1, "completion_tokens": 2, "total_tokens": 3, "cost": 0.5, "model": "m"})
    assert t.insert_usage({"prompt_tokens": None, "completion_tokens": None, "total_tokens": None, "cost": None, "model": None})
    assert not t.insert_usage({"prompt_tokens": 2 ** 31})
    assert not t.insert_usage({"prompt_tokens": "7"})
    assert not t.insert_usage({"prompt_tokens": object()})
    assert t.insert_usage({"prompt_tokens": 2.0, "model": 17})
    assert t.rejected == 3 and len(t) == 3
    t._materialise()
    assert t._host["prompt_tokens"].tolist() == [1, 0, 2] and t._models == ["m", None, "17"]
    assert t.insert_usage({"prompt_tokens": 4})
    t._materialise()
    assert t._host["prompt_tokens"].tolist() == [1, 0, 2, 4]


def test_sqlite_sink_stores_what_the_reference_stores(tmp_path):
    """SqliteUsageSink against tokens_usage_db.py:118 `insert_usage` run unmodified (tests/golden/usage_sink_cases.json): same
    schema, same stored value and storage class per column, and the same rows lost (values SQLite cannot bind)."""
    import sqlite3
    from llmapigateway_b200.gateway import SqliteUsageSink
    doc = json.loads((ROOT / "tests" / "golden" / "usage_sink_cases.json").read_text())
    sink = SqliteUsageSink(str(tmp_path / "u.db"))
    sink.insert_many(doc["rows"][:4])
    for r in doc["rows"][4:]:
        sink.insert_usage(r)
    conn = sqlite3.connect(str(tmp_path / "u.db"))
    cols = ["prompt_tokens", "completion_tokens", "total_tokens", "reasoning_tokens", "cached_tokens", "cost", "model", "provider"]
    sel = ", ".join(f"{c}, typeof({c})" for c in cols)
    assert [list(r) for r in conn.execute(f"SELECT {sel} FROM tokens_usage ORDER BY id")] == doc["stored"]
    assert [list(r[1:4]) for r in conn.execute("PRAGMA table_info(tokens_usage)")] == doc["schema"]
    ts = [r[0] for r in conn.execute("SELECT timestamp FROM tokens_usage")]
    from datetime import datetime
    assert all(datetime.fromisoformat(t) for t in ts)


def test_stats_window_matches_the_reference():
    from datetime import datetime
    from llmapigateway_b200.usage import stats_window
    doc = json.loads((ROOT / "tests" / "golden" / "usage_sink_cases.json").read_text())
    now = datetime.fromisoformat(doc["now"])
    for period, (start, end) in doc["windows"].items():
        s, e = stats_window(period, now)
        assert [s.isoformat(), e.isoformat()] == [start, end], period


def test_nonstream_responses_are_tapped_by_the_middleware_mirror():
    """log_chat_completions (chat_logging.py:165-231 seam): non-streaming responses of /chat/completions are collected and handed
    to the engine's document tap; streams the engine already tapped and other routes pass through untouched."""
    import types
    from llmapigateway_b200.gateway import log_chat_completions

    class Eng(FakeEngine):
        def documents_usage(self, docs):
            return [([{"prompt_tokens": len(d)}], False) for d in docs]

    async def go():
        sink = _Sink()
        b = StreamBatcher(Eng(max_streams=2), usage_sink=sink)

        async def chunks():
            yield b'{"usage":'
            yield b'{"prompt_tokens":1}}'

        resp = types.SimpleNamespace(headers={"content-type": "application/json"}, body_iterator=chunks())
        req = types.SimpleNamespace(url=types.SimpleNamespace(path="/v1/chat/completions"))

        async def call_next(_):
            return resp

        out = await log_chat_completions(req, call_next, batcher=b)
        got = [c async for c in out.body_iterator]
        assert b"".join(got) == b'{"usage":{"prompt_tokens":1}}' and sink.rows == [{"prompt_tokens": 29}]
        tapped = types.SimpleNamespace(headers={"content-type": "text/event-stream"}, body_iterator=chunks(), lgw_tapped=True)

        async def call_next2(_):
            return tapped

        assert (await log_chat_completions(req, call_next2, batcher=b)).body_iterator is tapped.body_iterator
        other = types.SimpleNamespace(url=types.SimpleNamespace(path="/v1/models"))
        assert await log_chat_completions(other, call_next, batcher=b) is resp
        assert len(sink.rows) == 1
    asyncio.run(go())


def test_every_attempt_gives_its_slot_back_exactly_once():
    """Failed attempts, upstream errors while priming, a response whose close raises, a cancelled priming loop and a client that
    walks away mid-stream: afterwards every slot is free again, once (a slot freed twice would be handed to two streams)."""
    ok_case = next(c for c in CASES if not c["failed"] and len(c["chunks"]) > 3)
    bad_case = next(c for c in CASES if c["failed"] and c["http_status"] < 400)

    def factory(chunks, status=200, fail_after=None, close_raises=False, hang_after=None):
        class _Body(httpx.AsyncByteStream):
            async def __aiter__(self):
                for k, c in enumerate(chunks):
                    if fail_after is not None and k == fail_after:
                        raise httpx.ReadError("upstream went away")
                    if hang_after is not None and k == hang_after:
                        await asyncio.sleep(3600)
                    yield c

            async def aclose(self):
                if close_raises:
                    raise httpx.ReadError("close failed")

        return lambda **kw: httpx.AsyncClient(transport=httpx.MockTransport(
            lambda request: httpx.Response(status, headers={"content-type": "text/event-stream"}, stream=_Body())), **kw)

    async def go():
        n = 8
        batcher = StreamBatcher(FakeEngine(max_streams=n), window_s=0.0005, usage_sink=_Sink())
        url = "http://upstream.test/v1/chat/completions"

        async def attempt(f):
            return await make_llm_request(url, {}, {"model": "m", "messages": []}, True, batcher=batcher, client_factory=f)

        def all_free():
            return sorted(batcher._free) == list(range(n))

        for _ in range(3):
            resp, err = await attempt(factory(bad_case["chunks"], close_raises=True))         # fails while priming; closing raises too
            assert resp is None and err is not None and all_free()
            resp, err = await attempt(factory(ok_case["chunks"], fail_after=0))                # upstream dies before the first chunk
            assert resp is None and err.startswith("RequestError connecting to") and all_free()
            resp, err = await attempt(factory([b"x"], status=502))                             # HTTP error: no slot is ever taken
            assert resp is None and err == "x" and all_free()
            resp, err = await attempt(factory(ok_case["chunks"]))                              # served in full
            assert err is None
            assert b"".join([bytes(c) async for c in resp.body_iterator]) == b"".join(ok_case["emitted"])
            assert all_free()
            resp, err = await attempt(factory(ok_case["chunks"], close_raises=True))           # served in full, then the upstream close raises
            got = []                                                                           #   (the relay generator passes that on, as the
            with pytest.raises(httpx.ReadError):                                               #   reference's does; the slot is free all the same)
                async for c in resp.body_iterator:
                    got.append(bytes(c))
            assert b"".join(got) == b"".join(ok_case["emitted"]) and all_free()
            resp, err = await attempt(factory(ok_case["chunks"]))                              # the client walks away mid-stream
            it = resp.body_iterator
            await it.__anext__()
            await it.aclose()
            assert all_free()
            task = asyncio.ensure_future(attempt(factory([b": ping\n\n"] * 2, hang_after=1)))
            await asyncio.sleep(0.05)                                                           # cancelled inside the priming loop
            task.cancel()
            with pytest.raises(asyncio.CancelledError):
                await task
            assert all_free()
    asyncio.run(go())


def test_config1_plumbing_matches_the_reference_golden():
    """BASELINE config 1: one NON-streaming request, body of exactly 256 bytes, mock upstream with one fixed JSON
    document, through chat.chat_completions and the log_chat_completions mirror (host logic over the fake engine = host build of
    the device machines).  Against tests/golden/c1_case.json, recorded from the UNMODIFIED endpoint: the response bytes FastAPI
    sends, the attempt (url, headers, wire body -- the json5.dumps form, unpinned) and the usage row the tap stores."""
    import base64
    import types
    from golden_io import GOLDEN, canon_rows
    from llmapigateway_b200 import chat, rewrite as rw
    from llmapigateway_b200.gateway import log_chat_completions
    sys.path.insert(0, str(GOLDEN))
    import make_c1_golden as c1
    g = json.loads((GOLDEN / "c1_case.json").read_text())
    body = base64.b64decode(g["request_body"])
    assert len(body) == 256 and body == c1.c1_body()
    providers, rules, fallback_provider = c1.c1_world()
    attempts = []

    def handler(request):
        hdr = {k: v for k, v in request.headers.items() if k.lower() in ("authorization", "http-referer", "x-title", "content-type")}
        attempts.append(dict(url=str(request.url), body=base64.b64encode(request.content).decode(), headers=hdr))
        return httpx.Response(200, headers={"content-type": "application/json"}, content=base64.b64decode(g["upstream_doc"]))

    async def go():
        sink = _Sink()
        batcher = StreamBatcher(FakeEngine(max_streams=4), window_s=0.0005, usage_sink=sink)
        batcher.load_rules(rw.RulePlans(rules, fallback_provider=fallback_provider))
        loader = types.SimpleNamespace(providers_config=providers, fallback_rules=rules)
        request = types.SimpleNamespace(app=types.SimpleNamespace(state=types.SimpleNamespace(config_loader=loader)),
                                        headers={"Authorization": "Bearer client-key"}, url=types.SimpleNamespace(path="/v1/chat/completions"))

        async def _body():
            return body
        request.body = _body

        async def call_next(req):
            return await chat.chat_completions(req, batcher=batcher, client_factory=lambda **kw: httpx.AsyncClient(transport=httpx.MockTransport(handler), **kw))

        resp = await log_chat_completions(request, call_next, batcher=batcher)
        return resp, sink.rows

    resp, rows = asyncio.run(go())
    assert resp.status_code == g["status"] and bytes(resp.body) == base64.b64decode(g["response_body"])
    strip = lambda a: {k: v for k, v in a.items() if k != "body"}
    assert [strip(a) for a in attempts] == [strip(a) for a in g["attempts"]]
    # the wire body is json5.dumps' (request_handler.py:153) -- json5 is absent here, the golden holds the stand-in's stdlib form:
    # same document, and the engine's bytes equal the oracle's restatement of json5.dumps (unpinned)
    from oracle import body_oracle as bo
    sent = base64.b64decode(attempts[0]["body"])
    want_doc = json.loads(base64.b64decode(g["attempts"][0]["body"]))
    assert sent == bo.RENDERERS["json5"](want_doc)
    assert canon_rows(rows) == g["rows"]


def test_a_request_body_the_engine_does_not_model_is_handed_back_before_any_attempt():
    """Found by tools/fuzz_chain_live.py: a body with a value the engine reports but does not model (a float that needs 17
    significant digits inside `messages`) used to be walked as a FAILED ATTEMPT -- and the next attempt's retry plan drops
    `messages` (the log scrub of chat.py:150), the offending value with it, so a scrubbed body went upstream as the first real
    attempt.  Now: chat_completions raises RequestNotModelled (501) with no upstream call made; ChainBatch reports the request as
    handed back (served_round -3) and walks the others as before."""
    import types
    from fake_engine import FakeEngine
    from llmapigateway_b200 import chat, rewrite, synth
    import chain_cases as cc
    providers, rules, fallback_provider = synth.chain_world()
    odd = b'{"model":"gw/retrying","stream":true,"messages":[{"role":"user","content":"x","w":0.12345678901234567}]}'
    calls = []

    def handler(request):
        calls.append(request.content)
        return httpx.Response(500, content=b"never expected")

    async def go():
        batcher = StreamBatcher(FakeEngine(max_streams=4), window_s=0.0005)
        batcher.load_rules(rewrite.RulePlans(rules, fallback_provider=fallback_provider, stream_mode=cc.stream_mode()))
        loader = types.SimpleNamespace(providers_config=providers, fallback_rules=rules)
        with pytest.raises(chat.RequestNotModelled) as ei:
            await chat.chat_completions(cc.FakeRequest(odd, {}, loader), batcher=batcher,
                                        client_factory=lambda **kw: httpx.AsyncClient(transport=httpx.MockTransport(handler), **kw))
        assert ei.value.status_code == 501 and "exotic" in ei.value.detail and calls == []
    asyncio.run(go())

    n = 12
    up = synth.ChainUpstream(n, 4, seed=6, p_fail=0.5)
    bodies = synth.chain_request_bodies(n, seed=6)
    bodies[3] = odd
    bodies[8] = odd.replace(b"gw/retrying", b"gw/chain3" if b"gw/chain3" in bodies[0] else b"gw/retrying")
    eng = FakeEngine(max_streams=n)
    plans = rewrite.RulePlans(rules, fallback_provider=fallback_provider, stream_mode=cc.stream_mode())
    eng.load_rules(plans)
    ref = chat.ChainBatch(eng, plans, providers, rules).run([b for k, b in enumerate(bodies) if k not in (3, 8)], None, up,
                                                             stream_ids=[k for k in range(n) if k not in (3, 8)])
    out = chat.ChainBatch(eng, plans, providers, rules).run(bodies, None, up)
    assert list(out.served_round[[3, 8]]) == [-3, -3] and all("not modelled by the engine (exotic)" in out.detail[k] for k in (3, 8))
    others = [k for k in range(n) if k not in (3, 8)]
    assert list(out.served_round[others]) == list(ref.served_round) and [out.emitted(k) for k in others] == [ref.emitted(j) for j in range(len(others))]
    assert [out.detail[k] for k in others] == list(ref.detail) and out.attempts == ref.attempts


def test_request_error_text_of_a_document_the_engine_does_not_render():
    """Found by tools/fuzz_chain_live.py --broken: a valid JSON object WITHOUT a "model" key that also holds a value the engine does
    not re-render (a float outside 1e+-290, a duplicate key, NaN) answered "request body is not valid JSON"; the reference's text is
    the KeyError's, `'model'` (chat.py:35-39) -- and the TypeError texts for non-object roots of that kind."""
    import types
    from fake_engine import FakeEngine
    from llmapigateway_b200 import chat, rewrite, synth
    import chain_cases as cc
    from fastapi import HTTPException
    providers, rules, fallback_provider = synth.chain_world()
    want = [(b'{"mode":"x","a":1.5e+300,"stream":true}', "Error reading request body: 'model'"),
            (b'{"a":1,"a":2}', "Error reading request body: 'model'"),
            (b'{"a":NaN}', "Error reading request body: 'model'"),
            (b'[1.5e+300]', "Error reading request body: list indices must be integers or slices, not str"),
            (b'1.5e+300', "Error reading request body: 'float' object does not support item assignment")]

    async def go():
        batcher = StreamBatcher(FakeEngine(max_streams=4), window_s=0.0005)
        batcher.load_rules(rewrite.RulePlans(rules, fallback_provider=fallback_provider, stream_mode=cc.stream_mode()))
        loader = types.SimpleNamespace(providers_config=providers, fallback_rules=rules)
        for body, detail in want:
            with pytest.raises(HTTPException) as ei:
                await chat.chat_completions(cc.FakeRequest(body, {}, loader), batcher=batcher)
            assert ei.value.status_code == 400 and ei.value.detail == detail, body
    asyncio.run(go())
