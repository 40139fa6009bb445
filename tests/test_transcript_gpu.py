"""Transcript tap (csrc/transcript.cuh) through the C ABI: the text the reference's tap accumulates in
`llm_response_accum` and the text at every write_log call, against the goldens of the UNMODIFIED ChunkProcessorThread
(tests/golden/sse_cases.json "transcripts") and against the oracle on seeded adversarial streams.  Byte-exact.
The bodies take the engine as an argument: tests/test_transcript_cpu.py runs them over the SIMT-emulated kernels."""
import json
import random

import numpy as np
import pytest

from golden_io import load_sse_cases
from llmapigateway_b200 import _abi
from llmapigateway_b200.synth import pack_streams, sse_batch
from llmapigateway_b200.transcripts import TranscriptBook, decode_text

pytestmark = pytest.mark.gpu

CASES = load_sse_cases()


@pytest.fixture(scope="module")
def engine():
    import llmapigateway_b200 as L
    e = L.Engine(max_streams=4096, carry_cap=16384, max_step_chunks=1 << 20, max_step_bytes=64 << 20)
    e.enable_transcripts()
    yield e
    e.close_engine()


def run_streams(engine, streams, schedules, http_status=None):
    """Feed `streams` (lists of chunks) in steps; -> per stream (transcripts at every write_log call incl. the final one,
    flags, emitted_any, regular_steps)."""
    n = len(streams)
    slots = np.arange(n, dtype=np.uint32)
    engine.open(slots, http_status)
    book = TranscriptBook()
    for i in range(n):
        book.open(i)
    bounds = [list(s) + [len(c)] for s, c in zip(schedules, streams)]
    n_steps = max(len(b) - 1 for b in bounds)
    snaps = [[] for _ in range(n)]
    seq_steps = [0] * n
    for k in range(n_steps):
        part, who = [], []
        for i, c in enumerate(streams):
            b = bounds[i]
            if k < len(b) - 1:
                part.append(c[b[k]:b[k + 1]]); who.append(i)
        pb = pack_streams(part, slots=who)
        engine.step(pb.data, pb.chunk_off, pb.seg_chunk, pb.seg_slot)
        st = engine.step_transcript()
        assert len(st.seg_off) == len(who) + 1 and int(st.seg_off[-1]) == len(st.text)
        for slot, seq, text in book.apply(pb.seg_slot, st):
            assert seq == len(snaps[slot]) + 1
            snaps[slot].append(text)
        for s, i in enumerate(who):
            if int(st.flags[s]) & _abi.TF_SEQUENTIAL:
                seq_steps[i] += 1
    states = engine.close(slots)
    out = []
    for i in range(n):
        text, flags = book.close(i)
        emitted_any = bool(states[i].flags & _abi.SF_EMITTED_ANY)
        if emitted_any:
            snaps[i].append(text)                      # the final write_log (chat_logging.py:150)
        out.append((snaps[i], flags, emitted_any, seq_steps[i]))
    return out


def _schedule(stepping, name, n):
    if stepping == "one_step" or n <= 1:
        return [0]
    if stepping == "step_per_chunk":
        return list(range(n))
    rng = random.Random(len(name) * 7919 + n)
    return [0] + sorted(set(rng.randrange(1, n) for _ in range(rng.randrange(0, 4))))


@pytest.mark.parametrize("stepping", ["one_step", "step_per_chunk", "random_steps"])
def test_golden_transcripts(engine, stepping):
    """The 236 reference-generated streams: llm_response_accum at every write_log call of the unmodified tap thread."""
    streams = [c["chunks"] for c in CASES]
    res = run_streams(engine, streams, [_schedule(stepping, c["name"], len(c["chunks"])) for c in CASES], [c["http_status"] for c in CASES])
    n_cmp = n_text = 0
    for c, (snaps, flags, emitted_any, _) in zip(CASES, res):
        label = f"{c['name']} {stepping}"
        if c["failed"]:
            assert snaps == [], label
            continue
        if flags & _abi.TF_EXOTIC:
            continue
        assert [decode_text(t) for t in snaps] == c["transcripts"], label
        n_cmp += 1
        n_text += any(c["transcripts"])
    assert n_cmp >= len(CASES) - 40 and n_text >= 50


# ---- adversarial content -------------------------------------------------------------------------------------------------
_PIECES = ["he", "llo", " wor", "ld", "", "\\n", "\\\\n", "\\\"q\\\"", "\\u00e9", "\\ud83d\\ude00", "\\u4e2d", "é", "中", "😀", "a\\/b", "\\t\\r\\b\\f",
           "\\ud83d", "\\ude00", "\\ud83d\\u0041", "\\uD83D\\uDE00x", "\\u0000", "{\\\"k\\\":1}", "data: {", "\\\\n\\\\n", "x" * 70]


def _content(rng):
    return "".join(rng.choice(_PIECES) for _ in range(rng.randrange(0, 4)))


def _choice(rng):
    r = rng.random()
    c = _content(rng)
    if r < 0.55:
        return '{"index":0,"delta":{"content":"%s"}}' % c
    if r < 0.65:
        return '{"index":0,"message":{"role":"assistant","content":"%s"}}' % c
    if r < 0.70:
        return '{"delta":{"role":"assistant"},"message":{"content":"%s"}}' % c           # delta without content: elif message
    if r < 0.74:
        return '{"delta":{"content":"%s"},"message":{"content":"never"}}' % c
    if r < 0.78:
        return '{"delta":{"content":"dup"},"delta":{"content":"%s"}}' % c                # last duplicate wins
    if r < 0.81:
        return '{"delta":{"content":"a","content":"%s"}}' % c
    if r < 0.84:
        return '{"delta":{"content":null}}'
    if r < 0.86:
        return '{"delta":{"content":0}}'
    if r < 0.88:
        return '{"delta":{"content":5}}'                                                  # str += int raises
    if r < 0.90:
        return '{"delta":null}'                                                            # "content" in None raises
    if r < 0.92:
        return '7'                                                                         # "delta" in 7 raises
    if r < 0.94:
        return '{"delta":{"content":{"x":"%s"}}}' % c                                      # truthy dict: raises
    if r < 0.96:
        return '{"delta":{"content":"%s","tool_calls":[{"function":{"content":"no"}}]}}' % c
    if r < 0.98:
        return '{"delta":{"nested":{"content":"no"}},"finish_reason":null}'
    return '{}'


def _event(rng):
    r = rng.random()
    n_ch = 1 if r < 0.7 else rng.randrange(0, 4)
    choices = "[" + ",".join(_choice(rng) for _ in range(n_ch)) + "]"
    head = '{"id":"c-%d","choices":%s' % (rng.randrange(100), choices)
    r = rng.random()
    if r < 0.05:
        head = '{"choices":[{"delta":{"content":"void"}}],"id":1,"choices":%s' % choices                    # repeated key: last wins
    elif r < 0.08:
        head = '{"choices":"text"'
    elif r < 0.10:
        head = '{"choices":null'
    elif r < 0.12:
        head = '{"nochoices":1'
    tail = "}"
    r = rng.random()
    if r < 0.06:
        tail = ',"error":{"message":"boom \\n %s"}}' % _content(rng)
    elif r < 0.10:
        tail = ',"usage":{"prompt_tokens":%d,"completion_tokens":2,"total_tokens":9}}' % rng.randrange(50)
    elif r < 0.13:
        tail = ',"x":[1,2}'                                                                                  # invalid: nothing is appended
    text = head + tail
    r = rng.random()
    if r < 0.85:
        pre = "data: "
    elif r < 0.92:
        pre = ""                                                                                             # "{" parts are tapped too
    elif r < 0.96:
        pre = "data:"
    else:
        pre = ": keep-alive"
        text = ""
    post = rng.choice(["", "", "", " ", "\t ", "\r", " \x0b"]) if pre == "data: " else ""
    return (pre + text + post + "\n\n").encode("utf-8")


def _stream(rng):
    n = rng.randrange(1, 14)
    blob_events = [b'data: {"id":"first","choices":[{"delta":{"content":"S"}}]}\n\n'] + [_event(rng) for _ in range(n)]
    style = rng.random()
    if style < 0.45:                       # one event per chunk: the one-chunk-per-lane path
        chunks = blob_events
    elif style < 0.6:                      # several whole events per chunk
        chunks, i = [], 0
        while i < len(blob_events):
            k = rng.randrange(1, 4)
            chunks.append(b"".join(blob_events[i:i + k])); i += k
    else:                                  # arbitrary cuts (also inside UTF-8 sequences and escapes)
        blob = b"".join(blob_events)
        cuts = sorted(set(rng.randrange(1, len(blob)) for _ in range(rng.randrange(1, 8))))
        chunks = [blob[a:b] for a, b in zip([0] + cuts, cuts + [len(blob)])]
    if rng.random() < 0.1:
        chunks.insert(rng.randrange(0, len(chunks) + 1), b"")
    if rng.random() < 0.08:
        chunks.insert(rng.randrange(1, len(chunks) + 1), b"\xff\xfe broken utf8\n\n")
    return chunks


@pytest.mark.parametrize("stepping", ["one_step", "random_steps"])
def test_adversarial_content_vs_oracle(engine, stepping, n_streams=600):
    from oracle.sse_oracle import run_stream
    rng = random.Random(77 if stepping == "one_step" else 78)
    streams = [_stream(rng) for _ in range(n_streams)]
    scheds = [_schedule(stepping, "s%d" % i, len(c)) for i, c in enumerate(streams)]
    res = run_streams(engine, streams, scheds)
    n_cmp = n_marks = n_regular = n_surr = 0
    for i, (chunks, (snaps, flags, emitted_any, seq_steps)) in enumerate(zip(streams, res)):
        relay, tap = run_stream(chunks)
        if relay.failed:
            assert snaps == [], i
            continue
        if flags & _abi.TF_EXOTIC:
            continue
        got = [decode_text(t) for t in snaps]
        assert got == tap.transcripts, (i, chunks)
        n_cmp += 1
        n_marks += len(snaps) - 1
        n_regular += seq_steps == 0
        has_surr = any(0xD800 <= ord(ch) <= 0xDFFF for ch in (tap.transcripts[-1] if tap.transcripts else ""))
        assert bool(flags & _abi.TF_LONE_SURROGATE) == has_surr, i
        n_surr += has_surr
    assert n_cmp >= n_streams * 0.8 and n_marks >= 20 and n_surr >= 5
    if stepping == "one_step":
        assert n_regular >= 30


@pytest.mark.parametrize("events_per_chunk", [1, 8])
def test_c3_shape_takes_the_lane_parallel_path(engine, events_per_chunk, n_streams=64):
    """BASELINE config 3 streams: every relayed chunk ends on LF LF, so no segment needs the sequential walk; the text is the
    concatenation of the 8-byte contents."""
    from oracle.sse_oracle import run_stream
    b = sse_batch(n_streams=n_streams, n_events=64, seed=9, events_per_chunk=events_per_chunk)
    streams = [b.stream_chunks(s) for s in range(n_streams)]
    res = run_streams(engine, streams, [[0]] * n_streams)
    for s in range(n_streams):
        snaps, flags, emitted_any, seq_steps = res[s]
        relay, tap = run_stream(streams[s])
        assert [decode_text(t) for t in snaps] == tap.transcripts and len(snaps[0]) == 64 * 8
        assert seq_steps == 0 and flags == 0


def test_mid_event_cuts_across_steps(engine):
    """Step boundaries inside events, escapes and UTF-8 sequences: the tap's carry lives on the device between steps."""
    from oracle.sse_oracle import run_stream
    rng = random.Random(5)
    streams = []
    for s in range(24):
        blob = b"".join([b'data: {"choices":[{"delta":{"content":"S"}}]}\n\n'] + [_event(rng) for _ in range(10)])
        cuts = list(range(11 + s, len(blob), 23 + s))
        streams.append([blob[i:j] for i, j in zip([0] + cuts, cuts + [len(blob)])])
    res = run_streams(engine, streams, [list(range(0, len(c), 2)) for c in streams])
    n = 0
    for chunks, (snaps, flags, _, _) in zip(streams, res):
        relay, tap = run_stream(chunks)
        if flags & _abi.TF_EXOTIC:
            continue
        assert [decode_text(t) for t in snaps] == tap.transcripts
        n += 1
    assert n >= 18


def test_transcript_pass_runs_once_per_step(engine):
    import llmapigateway_b200 as L
    pb = pack_streams([[b'data: {"choices":[{"delta":{"content":"x"}}]}\n\n']], slots=[0])
    engine.open([0])
    engine.step(pb.data, pb.chunk_off, pb.seg_chunk, pb.seg_slot)
    st = engine.step_transcript()
    assert st.segment(0) == b"x"
    if isinstance(engine, L.Engine):
        with pytest.raises(L.engine.EngineError):
            engine.step_transcript()
    engine.close([0])
