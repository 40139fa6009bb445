"""GPU parity tests: the CUDA path through the C ABI against the golden fixtures (outputs of the
unmodified reference) and against the oracle on seeded synthetic streams.  Byte-exact."""
import json
import re
import random

import numpy as np
import pytest

from golden_io import canon_rows, load_sse_cases
from llmapigateway_b200 import _abi
from llmapigateway_b200.synth import pack_streams, sse_batch
from stream_compare import check_stream

pytestmark = pytest.mark.gpu

CASES = load_sse_cases()


@pytest.fixture(scope="module")
def engine():
    import llmapigateway_b200 as L
    e = L.Engine(max_streams=8192, max_step_chunks=1 << 20, max_step_bytes=64 << 20)
    yield e
    e.close_engine()


def _expect(case):
    return dict(failed=case["failed"], error_detail=case["error_detail"], emitted=case["emitted"],
                end_raises=case["end_raises"], rows=case["rows"], http_status=case["http_status"])


@pytest.mark.parametrize("mode", [1, 0], ids=["general", "fast"])
@pytest.mark.parametrize("stepping", ["one_step", "step_per_chunk", "random_steps"])
def test_golden_cases_batched(engine, mode, stepping):
    def schedule(c):
        n = len(c["chunks"])
        if stepping == "one_step" or n <= 1:
            return [0]
        if stepping == "step_per_chunk":
            return list(range(n))
        rng = random.Random(len(c["name"]) * 7919 + n)
        return [0] + sorted(set(rng.randrange(1, n) for _ in range(rng.randrange(0, 4))))

    engine.set_mode(mode)
    n = len(CASES)
    slots = np.arange(n, dtype=np.uint32)
    engine.open(slots, [c["http_status"] for c in CASES])
    bounds = [list(schedule(c)) + [len(c["chunks"])] for c in CASES]
    n_steps = max(len(b) - 1 for b in bounds)
    rows, emitted = [], [[] for _ in CASES]
    details = {}
    for k in range(n_steps):
        streams, who = [], []
        for i, c in enumerate(CASES):
            b = bounds[i]
            if k < len(b) - 1:
                streams.append(c["chunks"][b[k]:b[k + 1]]); who.append(i)
        pb = pack_streams(streams, slots=who)
        res = engine.step(pb.data, pb.chunk_off, pb.seg_chunk, pb.seg_slot)
        rows += res.rows
        for s, i in enumerate(who):
            c0, c1 = int(pb.seg_chunk[s]), int(pb.seg_chunk[s + 1])
            eb = int(res.segs["emit_chunk_begin"][s])
            assert c0 <= eb <= c1
            for ch in range(eb, c1):
                o0, o1 = int(pb.chunk_off[ch]), int(pb.chunk_off[ch + 1])
                if o1 > o0:
                    emitted[i].append(res.out[o0:o1].tobytes())
            if res.segs["verdict"][s] in (_abi.VERDICT_FAIL_EVENT, _abi.VERDICT_FAIL_PARSE) and i not in details:
                details[i] = engine.detail(i)
    states = engine.close(slots)
    n_exotic = 0
    for i, (c, st) in enumerate(zip(CASES, states)):
        label = f"{c['name']} mode={mode} {stepping}"
        failed = st.phase == _abi.PHASE_FAILED
        assert failed == c["failed"], label
        assert emitted[i] == c["emitted"], label
        if failed:
            if c["http_status"] >= 400:
                assert st.verdict == _abi.VERDICT_FAIL_HTTP, label
            elif c["error_detail"].startswith("Unexpected error during request to"):
                assert st.verdict == _abi.VERDICT_FAIL_PARSE, label
            else:
                assert st.verdict == _abi.VERDICT_FAIL_EVENT and details[i].decode() == c["error_detail"], label
            continue
        assert (not (st.flags & _abi.SF_A_USAGE_BOUND)) == c["end_raises"], label
        if st.n_exotic:
            n_exotic += 1
            continue
        got = [_abi.usage_rec_to_dict(ev.rec) for ev in sorted((r for r in rows if r.slot == i), key=lambda r: r.seq)]
        if st.flags & _abi.SF_EMITTED_ANY:
            got.append(_abi.usage_rec_to_dict(st.rec))
        assert canon_rows(got) == c["rows"], label
    assert n_exotic <= 6


@pytest.mark.parametrize("mode", [1, 0], ids=["general", "fast"])
@pytest.mark.parametrize("events_per_chunk", [1, 8])
def test_c3_small_vs_oracle(engine, mode, events_per_chunk, n_streams=96):
    """SURVEY 8(d) C3 shape at a size the oracle finishes in seconds: bytes, verdicts, usage rows."""
    from oracle.sse_oracle import run_stream
    engine.set_mode(mode)
    b = sse_batch(n_streams=n_streams, n_events=64, seed=3, events_per_chunk=events_per_chunk)
    engine.open(b.seg_slot)
    res = engine.step(b.data, b.chunk_off, b.seg_chunk, b.seg_slot)
    states = engine.close(b.seg_slot)
    assert np.array_equal(res.out, b.data)                       # every chunk relayed verbatim
    assert (res.segs["emit_chunk_begin"] == b.seg_chunk[:-1]).all()
    for s in range(n_streams):
        relay, tap = run_stream(b.stream_chunks(s))
        assert not relay.failed and not relay.end_raises
        st = states[s]
        assert st.phase == _abi.PHASE_COMMITTED and (st.flags & _abi.SF_A_USAGE_BOUND)
        assert canon_rows([_abi.usage_rec_to_dict(st.rec)]) == canon_rows(tap.rows)
        assert tap.rows[0] == b.truths[s].expected_row()
        assert st.n_chunks_emitted == len(relay.emitted) and st.bytes_emitted == sum(map(len, relay.emitted))


def test_c3_two_steps_with_mid_event_cut(engine):
    """A step boundary in the middle of an event: the carry must survive between launches."""
    from oracle.sse_oracle import run_stream
    engine.set_mode(0)
    b = sse_batch(n_streams=8, n_events=32, seed=5, events_per_chunk=1)
    streams = []
    for s in range(8):
        blob = b"".join(b.stream_chunks(s))
        cuts = list(range(37 + s, len(blob), 97 + s))
        streams.append([blob[i:j] for i, j in zip([0] + cuts, cuts + [len(blob)])])
    half = [len(x) // 2 for x in streams]
    engine.open(np.arange(8))
    outs = [[] for _ in range(8)]
    for part in (0, 1):
        pb = pack_streams([x[:h] if part == 0 else x[h:] for x, h in zip(streams, half)])
        res = engine.step(pb.data, pb.chunk_off, pb.seg_chunk, pb.seg_slot)
        for s in range(8):
            for ch in range(int(res.segs["emit_chunk_begin"][s]), int(pb.seg_chunk[s + 1])):
                outs[s].append(res.out[int(pb.chunk_off[ch]):int(pb.chunk_off[ch + 1])].tobytes())
    states = engine.close(np.arange(8))
    for s in range(8):
        relay, tap = run_stream(streams[s])
        assert outs[s] == relay.emitted
        assert canon_rows([_abi.usage_rec_to_dict(states[s].rec)]) == canon_rows(tap.rows)


@pytest.mark.parametrize("mode", [0], ids=["fast"])
def test_c3_full_size_properties(engine, mode):
    """BASELINE.json config 3 at full size (4096 x 512 x 64 B): size-independent properties."""
    import llmapigateway_b200 as L
    e = L.Engine(max_streams=4096, max_step_chunks=(4096 * 514) + 8, max_step_bytes=160 << 20)
    try:
        e.set_mode(mode)
        b = sse_batch(n_streams=4096, n_events=512, seed=3)
        e.open(b.seg_slot)
        res = e.step(b.data, b.chunk_off, b.seg_chunk, b.seg_slot)
        states = e.close(b.seg_slot)
        assert np.array_equal(res.out, b.data)                    # relay is the identity on committed streams
        assert (res.segs["emit_chunk_begin"] == b.seg_chunk[:-1]).all()
        assert (res.segs["phase"] == _abi.PHASE_COMMITTED).all()
        for s in range(0, 4096, 97):
            assert _abi.usage_rec_to_dict(states[s].rec) == b.truths[s].expected_row()
        assert sum(st.n_events_b for st in states) == 4096 * 513   # 512 deltas + usage event; [DONE] is not JSON
        assert sum(st.n_chunks_emitted for st in states) == b.n_chunks
        assert all(st.flags & _abi.SF_A_USAGE_BOUND for st in states)
        # 256 of the full-size streams (every 16th) through the oracle itself: relayed chunks, usage rows, end state
        from oracle import sse_oracle
        for s in range(0, 4096, 16):
            chunks = b.stream_chunks(s)
            relay, tap = sse_oracle.run_stream(chunks, 200)
            c0, c1 = int(b.seg_chunk[s]), int(b.seg_chunk[s + 1])
            assert not relay.failed and states[s].phase == _abi.PHASE_COMMITTED
            got = [res.out[int(b.chunk_off[c]):int(b.chunk_off[c + 1])].tobytes() for c in range(int(res.segs["emit_chunk_begin"][s]), c1)]
            assert got == relay.emitted, s
            rows = [r for r in res.rows if r.slot == int(b.seg_slot[s])]
            from stream_compare import rows_from_result
            assert canon_rows(rows_from_result(states[s], rows)) == canon_rows(tap.rows), s
    finally:
        e.close_engine()


def _random_streams(n, seed):
    """Mostly well-formed streams (so the bulk path is exercised) with a share of odd ones."""
    import sse_cases as sc
    rng = random.Random(seed)
    out = []
    for i in range(n):
        k = rng.random()
        n_ev = rng.randrange(1, 40)
        if k < 0.6:      # clean: deltas, usage, done
            evs = [sc.delta(rng.choice(["a", "hello", "\u00e9t\u00e9", "x" * rng.randrange(1, 200)])) for _ in range(n_ev)]
            if rng.random() < 0.8:
                u = dict(sc.USAGE); u["prompt_tokens"] = rng.randrange(10**6)
                evs.append(sc.ev({"choices": [], "usage": u, "model": "m%d" % i, "provider": "P"}))
            if rng.random() < 0.8:
                evs.append(sc.DONE)
        else:            # anything from the adversarial pool
            evs = [sc._random_event(rng) for _ in range(n_ev)]
            if rng.random() < 0.7:
                evs.insert(0, sc.delta("lead"))
        blob = b"".join(evs)
        m = rng.random()
        if m < 0.5:
            chunks = evs
        elif m < 0.75:
            chunks = sc.rechunk(blob, [rng.randrange(1, max(2, len(blob))) for _ in range(rng.randrange(0, 12))])
        else:
            step = rng.randrange(3, 150)
            chunks = sc.rechunk(blob, list(range(step, len(blob), step)))
        out.append([c for c in chunks if c])
    return out


def _run_all(engine, streams, mode, n_steps, seed):
    engine.set_mode(mode)
    n = len(streams)
    slots = np.arange(n, dtype=np.uint32)
    engine.open(slots)
    rng = random.Random(seed)
    bounds = []
    for st in streams:
        cuts = sorted(set(rng.randrange(0, len(st) + 1) for _ in range(n_steps - 1)))
        bounds.append([0] + cuts + [len(st)])
    emitted = [[] for _ in streams]
    rows = []
    for k in range(n_steps):
        parts, who = [], []
        for i, st in enumerate(streams):
            b = bounds[i]
            if k < len(b) - 1 and b[k + 1] > b[k]:
                parts.append(st[b[k]:b[k + 1]]); who.append(i)
        if not parts:
            continue
        pb = pack_streams(parts, slots=who)
        res = engine.step(pb.data, pb.chunk_off, pb.seg_chunk, pb.seg_slot)
        rows += [(r.slot, r.seq, canon_rows([_abi.usage_rec_to_dict(r.rec)]) if not r.rec.exotic else 'exotic') for r in res.rows]
        for s, i in enumerate(who):
            for ch in range(int(res.segs["emit_chunk_begin"][s]), int(pb.seg_chunk[s + 1])):
                emitted[i].append(res.out[int(pb.chunk_off[ch]):int(pb.chunk_off[ch + 1])].tobytes())
    states = engine.close(slots)
    return states, sorted(rows), emitted


@pytest.mark.parametrize("n_steps", [1, 3])
def test_bulk_path_equals_sequential_path_and_oracle(engine, n_steps, n_streams=1500, min_regular=1000):
    """Differential: bulk kernel (+ fix-up) vs the exact sequential kernel vs the oracle, on
    random streams with random chunking and random step boundaries."""
    from oracle.sse_oracle import run_stream
    streams = _random_streams(n_streams, seed=77 + n_steps)
    s_fast, r_fast, e_fast = _run_all(engine, streams, 0, n_steps, seed=5)
    s_seq, r_seq, e_seq = _run_all(engine, streams, 1, n_steps, seed=5)
    assert e_fast == e_seq
    assert r_fast == r_seq
    n_regular = 0
    for i, (a, b) in enumerate(zip(s_fast, s_seq)):
        assert bytes(a)[:64] == bytes(b)[:64], (i, streams[i][:3])          # header: phase, flags, carries, counters
        if a.flags & _abi.SF_REC_VALID:
            assert a.rec.exotic == b.rec.exotic and (a.rec.exotic or _abi.usage_rec_to_dict(a.rec) == _abi.usage_rec_to_dict(b.rec) or canon_rows([_abi.usage_rec_to_dict(a.rec)]) == canon_rows([_abi.usage_rec_to_dict(b.rec)])), i
        relay, tap = run_stream(streams[i])
        assert e_fast[i] == relay.emitted, i
        assert (a.phase == _abi.PHASE_FAILED) == relay.failed
        if not relay.failed and not a.n_exotic:
            got = [json.loads(r[2])[0] for r in r_fast if r[0] == i]
            if a.flags & _abi.SF_EMITTED_ANY:
                got.append(json.loads(canon_rows([_abi.usage_rec_to_dict(a.rec)]))[0])
            assert canon_rows(got) == canon_rows(tap.rows), i
            assert (not (a.flags & _abi.SF_A_USAGE_BOUND)) == relay.end_raises
            n_regular += 1
    assert n_regular > min_regular


@pytest.mark.parametrize("relay_from", ["device", "host"])
def test_gateway_seam_on_the_real_engine(engine, relay_from):
    """make_llm_request + StreamBatcher over the CUDA engine against the reference goldens (relay_from="host": verdicts-only steps,
    lgw_sse_step with out_bytes = NULL, the relayed chunks are the upstream's own)."""
    import asyncio
    from test_gateway_cpu import _Sink, _drive
    from llmapigateway_b200.gateway import StreamBatcher
    engine.set_mode(0)

    async def go():
        sink = _Sink()
        batcher = StreamBatcher(engine, window_s=0.0005, usage_sink=sink, relay_from=relay_from)
        picks = CASES[:150]
        results = await asyncio.gather(*[_drive(c, batcher, _Sink()) for c in picks])
        n_rows_ok = 0
        for c, r in zip(picks, results):
            assert r["failed"] == c["failed"], c["name"]
            if c["failed"]:
                if not c["error_detail"].startswith("Unexpected error during request to"):
                    assert r["error_detail"] == c["error_detail"], c["name"]
                continue
            assert r["emitted"] == c["emitted"], c["name"]
        assert batcher.steps > 0
    asyncio.run(go())


NUMBER_SPELLINGS = ["0", "-0", "01", "00", "7", "1.5", "1.", ".5", "-", "-12", "+1", "1e5", "1E+5", "1e", "1e+", "1.5e-3", "1.e3", "1e5.3",
                    "0x10", "1_0", "9" * 25, "12345678", "123456789012", "0.001234", "-0.0", "1-2", "1 2", "NaN", "-Infinity", "Infinity", "1e005"]


def _template_variant_streams(n_streams, seed):
    """Streams of near-identical events: the bulk kernel's template shortcut must accept exactly the
    variants that differ inside ONE string value by plain bytes, and fall back for everything else."""
    import sse_cases as sc
    rng = random.Random(seed)
    base = {"id": "chatcmpl-9x", "object": "chat.completion.chunk", "created": 1726900000, "model": "gpt-x",
            "choices": [{"index": 0, "delta": {"content": "hello"}, "finish_reason": None}]}
    plain = ["a", "hello", "hello world", "", "x" * 40, "café 中", "}{][,:", "data: {"]
    nasty = ['q"q', "b\\s", "line\nbreak", "tab\t", " ", "\x01", "a\\u0041", '\\"']
    out = []
    for s in range(n_streams):
        evs = [sc.ev(base), sc.ev(base)]
        for _ in range(rng.randrange(5, 40)):
            d = json.loads(json.dumps(base))
            k = rng.random()
            if k < 0.55:
                d["choices"][0]["delta"]["content"] = rng.choice(plain) + rng.choice(plain)
            elif k < 0.70:
                d["choices"][0]["delta"]["content"] = rng.choice(nasty) + rng.choice(plain)
            elif k < 0.75:
                d["id"] = "chatcmpl-" + rng.choice(plain)
            elif k < 0.80:
                d["created"] = rng.randrange(10**9)
            elif k < 0.84:
                d["choices"][0]["finish_reason"] = "stop"
            elif k < 0.88:
                d["usage"] = dict(sc.USAGE)
            elif k < 0.91:
                d["choices"][0]["delta"] = {"conten": "x"}
            elif k < 0.94:
                d = {"error": {"message": "late"}}
            text = json.dumps(d, separators=(",", ":"), ensure_ascii=bool(rng.random() < 0.3))
            m = rng.random()
            if m < 0.06:                      # raw mutations of the text
                i = rng.randrange(len(text))
                text = text[:i] + rng.choice(['"', "\\", "}", "{", ",", " ", "\n", "\x00", "x"]) + text[i + 1:]
            elif m < 0.09:
                text = text.replace('"content"', '"cont\\u0065nt"')
            elif m < 0.12:
                text = text + rng.choice([" ", "\t", "\x0c", "x"])
            elif m < 0.22 and '"created":' in text:      # number values: valid and invalid spellings (number spans of the templates)
                text = re.sub(r'"created":\d+', '"created":' + rng.choice(NUMBER_SPELLINGS), text)
            evs.append(sc.ev(text))
            if rng.random() < 0.03:
                evs.append(rng.choice([b": ping\n\n", sc.DONE, b"\n", b'{"usage":{"prompt_tokens":1}}\n\n']))
        blob = b"".join(evs)
        m = rng.random()
        if m < 0.6:
            chunks = evs
        elif m < 0.8:
            step = rng.randrange(20, 400)
            chunks = sc.rechunk(blob, list(range(step, len(blob), step)))
        else:
            chunks = sc.rechunk(blob, [rng.randrange(1, len(blob)) for _ in range(rng.randrange(1, 10))])
        out.append([c for c in chunks if c])
    return out


@pytest.mark.parametrize("n_steps", [1, 2])
def test_template_shortcut_is_exact(engine, n_steps, n_streams=1200, min_rows=900):
    from oracle.sse_oracle import run_stream
    streams = _template_variant_streams(n_streams, seed=4242 + n_steps)
    s_fast, r_fast, e_fast = _run_all(engine, streams, 0, n_steps, seed=9)
    s_seq, r_seq, e_seq = _run_all(engine, streams, 1, n_steps, seed=9)
    assert e_fast == e_seq and r_fast == r_seq
    n_rows = 0
    for i, (a, b) in enumerate(zip(s_fast, s_seq)):
        assert bytes(a)[:64] == bytes(b)[:64], (i, [c[:80] for c in streams[i][:4]])
        relay, tap = run_stream(streams[i])
        assert e_fast[i] == relay.emitted, i
        assert (a.phase == _abi.PHASE_FAILED) == relay.failed
        if not relay.failed and not a.n_exotic:
            got = [json.loads(r[2])[0] for r in r_fast if r[0] == i]
            if a.flags & _abi.SF_EMITTED_ANY:
                got.append(json.loads(canon_rows([_abi.usage_rec_to_dict(a.rec)]))[0])
            assert canon_rows(got) == canon_rows(tap.rows), i
            assert (not (a.flags & _abi.SF_A_USAGE_BOUND)) == relay.end_raises
            n_rows += 1
    assert n_rows > min_rows


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("events_per_chunk", [(1, 1), (1, 4)])
def test_openai_shaped_streams_vs_oracle(engine, mode, events_per_chunk, n_streams=48):
    """Realistic OpenAI chunks (id / created / model on every event, role and finish chunks, content pieces of varying length
    with escapes and raw UTF-8, usage chunk, [DONE]): multi-span templates, bytes / usage rows / counters against the oracle"""
    from llmapigateway_b200.synth import openai_batch
    from oracle.sse_oracle import run_stream
    engine.set_mode(mode)
    try:
        b = openai_batch(n_streams=n_streams, n_deltas=40, seed=17, events_per_chunk=events_per_chunk)
        engine.open(b.seg_slot)
        res = engine.step(b.data, b.chunk_off, b.seg_chunk, b.seg_slot)
        states = engine.close(b.seg_slot)
        assert np.array_equal(res.out, b.data)
        assert (res.segs["emit_chunk_begin"] == b.seg_chunk[:-1]).all()
        for s in range(n_streams):
            relay, tap = run_stream(b.stream_chunks(s))
            assert not relay.failed
            st = states[s]
            assert st.phase == _abi.PHASE_COMMITTED
            assert canon_rows([_abi.usage_rec_to_dict(st.rec)]) == canon_rows(tap.rows)
            assert tap.rows[-1] == b.truths[s].expected_row()
            assert st.n_chunks_emitted == len(relay.emitted) and st.bytes_emitted == sum(map(len, relay.emitted))
    finally:
        engine.set_mode(0)


@pytest.mark.parametrize("direct", ["out", "both", "0"])
def test_pinned_host_buffers_direct_modes(direct):
    """Host-buffer step with page-locked buffers: the bulk kernel's TMA stores (and, with LGW_DIRECT=both, its loads) go straight
    to / from host memory.  Same bytes, same segment results, same final states as the staged path on pageable buffers."""
    import os, subprocess, sys
    code = r'''
import numpy as np, sys
sys.path.insert(0, "tests")
import llmapigateway_b200 as L
from llmapigateway_b200 import _abi
from llmapigateway_b200.synth import sse_batch
b = sse_batch(256, 128, seed=17)
eng = L.Engine(max_streams=256, max_step_chunks=256 * 130 + 8, max_step_bytes=int(b.data.size) + 4096)
status = np.full(256, 200, np.int32)
eng.open(b.seg_slot, status)
ref = eng.step(b.data.copy(), b.chunk_off, b.seg_chunk, b.seg_slot)          # pageable: staged path
assert not eng.last_step_direct()
ref_states = [bytes(s) for s in eng.close(b.seg_slot)]
pin_in, pin_out = eng.alloc_pinned(int(b.data.size)), eng.alloc_pinned(int(b.data.size))
pin_in[:] = b.data; pin_out[:] = 0
eng.open(b.seg_slot, status)
got = eng.step(pin_in, b.chunk_off, b.seg_chunk, b.seg_slot, out=pin_out)
assert eng.last_step_direct() == (sys.argv[1] != "0"), eng.last_step_direct()
states = [bytes(s) for s in eng.close(b.seg_slot)]
assert np.array_equal(got.out, ref.out) and np.array_equal(pin_out, b.data)
assert got.segs.tobytes() == ref.segs.tobytes() and states == ref_states
assert _abi.usage_rec_to_dict(_abi.StreamState.from_buffer_copy(states[3]).rec) == b.truths[3].expected_row()
print("DIRECT_OK")
'''
    r = subprocess.run([sys.executable, "-c", code, direct], capture_output=True, text=True, timeout=300, env=dict(os.environ, LGW_DIRECT=direct),
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "DIRECT_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])


def test_verdicts_only_step_equals_the_full_step(engine):
    """lgw_sse_step with out_bytes = NULL (no download of the re-emitted bytes): segment results, row events and final states equal
    those of the ordinary step on the same streams, sliced pipeline (> 8 MiB) and single slice alike; the bytes the caller relays
    from its own buffer are the bytes the engine re-emits."""
    engine.set_mode(0)
    for n_streams, n_events in [(96, 24), (2048, 80)]:
        b = sse_batch(n_streams=n_streams, n_events=n_events, seed=31)
        got = []
        for host in (False, True):
            engine.open(b.seg_slot)
            half = [int(b.seg_chunk[s] + (b.seg_chunk[s + 1] - b.seg_chunk[s]) // 2) for s in range(n_streams)]
            parts = []
            for lo_of, hi_of in ((lambda s: int(b.seg_chunk[s]), lambda s: half[s]), (lambda s: half[s], lambda s: int(b.seg_chunk[s + 1]))):
                pb = pack_streams([[b.data[int(b.chunk_off[c]):int(b.chunk_off[c + 1])].tobytes() for c in range(lo_of(s), hi_of(s))] for s in range(n_streams)],
                                  slots=list(range(n_streams)))
                r = engine.step(pb.data, pb.chunk_off, pb.seg_chunk, pb.seg_slot, relay_from_host=host)
                emitted = [r.out[int(pb.chunk_off[int(r.segs["emit_chunk_begin"][s])]):int(pb.chunk_off[int(pb.seg_chunk[s + 1])])].tobytes() for s in range(n_streams)]
                parts.append((r.segs.tobytes(), sorted((e.slot, e.seq, canon_rows([_abi.usage_rec_to_dict(e.rec)])) for e in r.rows), emitted))
            states = engine.close(b.seg_slot)
            # (header bytes + the record's values: the padding inside the record's value cells is not part of the contract)
            got.append((parts, [(bytes(st)[:64], canon_rows([_abi.usage_rec_to_dict(st.rec)])) for st in states]))
        assert got[0] == got[1]


def test_container_valued_usage_fields_are_never_read_from_a_span(engine):
    """Regression (found by tools/fuzz_relay2_cpu.py): `"prompt_tokens":[1]` / `"model":["x"]` -- the container's last inner scalar
    used to be taken for the field's own value span, so a template learnt from such an event read `1` / "x" where the reference
    stores the list itself (get_token_usage hands it on, chat_logging.py:237-267).  Bulk path == sequential path == oracle."""
    from oracle.sse_oracle import run_stream
    import sse_cases as sc
    ev = ('data: {"choices":[],"usage":{"prompt_tokens":%s,"completion_tokens":%s,"total_tokens":3,"cost":0.5,'
          '"completion_tokens_details":{"reasoning_tokens":%s},"prompt_tokens_details":{"cached_tokens":0}},"model":%s,"provider":"P"}\n\n')
    shapes = [("[1]", "2", "0", '"m"'), ('{"a":7}', "2", "0", '"m"'), ("1", "[22]", "0", '"m"'), ("1", "2", "[5]", '"m"'), ("1", "2", "0", '["x"]'),
              ("1", "2", "0", '{"k":"x"}'), ("[1]", "[2]", "[3]", '["x"]'), ("1", "2", "0", '"m"')]
    streams = []
    for rep in range(6):                       # the same skeletons again and again: later streams follow the templates the first ones left
        for k, sh in enumerate(shapes):
            tail = (ev % sh).replace("[1]", "[%d]" % (rep + 1)).replace('["x"]', '["y%d"]' % rep).encode()
            streams.append([sc.delta("a"), sc.delta("hello %d" % k), tail, sc.DONE])
    engine.reset_templates() if hasattr(engine, "reset_templates") else None
    s_fast, r_fast, e_fast = _run_all(engine, streams, 0, 1, seed=1)
    s_seq, r_seq, e_seq = _run_all(engine, streams, 1, 1, seed=1)
    assert e_fast == e_seq and r_fast == r_seq
    for i, (a, b) in enumerate(zip(s_fast, s_seq)):
        assert bytes(a)[:64] == bytes(b)[:64], i
        assert a.rec.exotic == b.rec.exotic and str(_abi.usage_rec_to_dict(a.rec)) == str(_abi.usage_rec_to_dict(b.rec)), (i, streams[i][2])
        relay, tap = run_stream(streams[i])
        assert e_fast[i] == relay.emitted
        if not a.n_exotic:
            assert canon_rows([_abi.usage_rec_to_dict(a.rec)]) == canon_rows(tap.rows), i
