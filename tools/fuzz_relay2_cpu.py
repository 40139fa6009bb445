"""Differential fuzz campaign of the bulk SSE kernels on the CPU box (test aid, not a product path).

The kernels of csrc/relay2.cuh are run through the SIMT emulator (tests/host_relay.py) in their bulk mode (mode 0) and in the
exact sequential mode (mode 1) on streams from three generators -- the two of tests/test_sse_gpu.py with fresh seeds, and a
third one aimed at the usage-field read-out (template-following usage events whose numbers and strings take every spelling
JSON allows) -- and both are compared with the oracle (oracle/sse_oracle.py, the restatement of request_handler.py:21-150 and
chat_logging.py:87-150,233-272).  Engine geometries (blocks, tiles per warp) and step counts vary per round.

    python tools/fuzz_relay2_cpu.py --rounds 40 --procs 6 --seed 1000

Prints one line per round; a divergence is reported with the generator, seed and stream so that it can be replayed
(`--replay gen:seed:n_steps:geometry:cold`).
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import random
import sys
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

GEOMETRIES = [(2, 0), (3, 1), (1, 0), (5, 2), (4, 1)]
GENERATORS = ["random", "template", "usage", "usage", "openai", "c3", "skeleton", "skeleton"]

NUMBERS = ["0", "7", "12", "123456", "2147483647", "2147483648", "4294967296", "9007199254740993", "18446744073709551616",
           "-1", "-0", "0.0", "0.5", "1.25", "1e3", "1E3", "1e+3", "1e-3", "12.5e-2", "0.000001", "1.7976931348623157e308", "1e400",
           "123456789012345678901234567890", "0.1234567890123456789", "3.0", "5e0", "100000000000000000000.0", "4.9e-324", "1e-400"]
BAD_NUMBERS = ["01", "1.", ".5", "+1", "1e", "1e+", "--1", "0x10", "1_000", "NaN", "Infinity", "1.2.3", "1e1.5", ""]
STRINGS = ["m", "gpt-4o", "openai/gpt-4o-mini", "", "x" * 63, "x" * 64, "x" * 65, "x" * 200, "café", "中文", "\U0001F600",
           "a\\\"b", "tab\\there", "nl\\nx", "\\u0041\\u00e9", "\\ud83d\\ude00", "\\ud83d", "\\ude00x", "back\\\\slash", "sl\\/ash", "\\b\\f\\r"]
BAD_STRINGS = ["a\\qb", "\\u12", "\\u12G4", "ctl\x01x", 'q"q']


def usage_variant_streams(n_streams: int, seed: int):
    """Streams whose usage events share ONE skeleton (so that after the first few the kernel holds a usage template) while the
    values of the eight fields run through every number and string spelling; some events break the skeleton on purpose."""
    import sse_cases as sc
    rng = random.Random(seed)
    out = []
    for s in range(n_streams):
        evs = [sc.delta(rng.choice(["a", "hello", "x" * rng.randrange(1, 90)])) for _ in range(rng.randrange(1, 24))]
        n_usage = 1 if rng.random() < 0.8 else rng.randrange(0, 4)
        for _ in range(n_usage):
            def num():
                r = rng.random()
                return rng.choice(NUMBERS) if r < 0.85 else (rng.choice(BAD_NUMBERS) if r < 0.93 else rng.choice(["null", "true", '"7"', "[1]", "{}"]))

            def text():
                r = rng.random()
                return '"' + (rng.choice(STRINGS) if r < 0.9 else rng.choice(BAD_STRINGS)) + '"' if r < 0.96 else rng.choice(["null", "12", "[]"])
            shape = rng.random()
            if shape < 0.75:         # the C3 skeleton: every field present
                ev = ('{"choices":[],"usage":{"prompt_tokens":%s,"completion_tokens":%s,"total_tokens":%s,"cost":%s,'
                      '"completion_tokens_details":{"reasoning_tokens":%s},"prompt_tokens_details":{"cached_tokens":%s}},"model":%s,"provider":%s}'
                      % (num(), num(), num(), num(), num(), num(), text(), text()))
            elif shape < 0.85:       # a second skeleton: no details, no provider
                ev = '{"id":"u","choices":[],"usage":{"prompt_tokens":%s,"completion_tokens":%s,"total_tokens":%s},"model":%s}' % (num(), num(), num(), text())
            elif shape < 0.92:       # usage beside a delta (the choices walk meets value spans)
                ev = '{"choices":[{"index":0,"delta":{"content":%s}}],"usage":{"prompt_tokens":%s,"completion_tokens":%s,"total_tokens":%s,"cost":%s}}' % (text(), num(), num(), num(), num())
            else:                    # details present but null / not objects
                ev = '{"choices":[],"usage":{"prompt_tokens":%s,"completion_tokens":%s,"total_tokens":%s,"completion_tokens_details":%s,"prompt_tokens_details":%s},"model":%s}' % (
                    num(), num(), num(), rng.choice(["null", "{}", '{"reasoning_tokens":3}', "5"]), rng.choice(["null", "{}", '{"cached_tokens":4}', '"x"']), text())
            if rng.random() < 0.04:
                i = rng.randrange(len(ev)); ev = ev[:i] + rng.choice(['"', "\\", "}", " ", "\n", ","]) + ev[i + 1:]
            evs.append(("data: " + ev + "\n\n").encode("utf-8", errors="surrogatepass") if rng.random() < 0.97 else (ev + "\n\n").encode("utf-8", errors="surrogatepass"))
            if rng.random() < 0.3:
                evs.append(sc.delta("more"))
        if rng.random() < 0.8:
            evs.append(sc.DONE)
        blob = b"".join(evs)
        m = rng.random()
        if m < 0.55:
            chunks = evs
        elif m < 0.8:
            step = rng.randrange(5, 300)
            chunks = sc.rechunk(blob, list(range(step, len(blob), step)))
        else:
            chunks = sc.rechunk(blob, [rng.randrange(1, max(2, len(blob))) for _ in range(rng.randrange(0, 9))])
        out.append([c for c in chunks if c])
    return out


def _recut(rng, chunks):
    """the same bytes in other network chunks: whole events, a fixed stride, or a few arbitrary cuts"""
    import sse_cases as sc
    m = rng.random()
    if m < 0.5:
        return chunks
    blob = b"".join(chunks)
    if m < 0.75:
        step = rng.randrange(7, 700)
        return [c for c in sc.rechunk(blob, list(range(step, len(blob), step))) if c]
    return [c for c in sc.rechunk(blob, [rng.randrange(1, max(2, len(blob))) for _ in range(rng.randrange(1, 9))]) if c]


def shaped_streams(gen: str, n: int, seed: int):
    """`openai`: realistic OpenAI chunks (synth.openai_stream: id / created / model on every event, role and finish chunks, pieces
    of varying length with escapes and raw UTF-8); `c3`: the benchmark's own 64-byte deltas (synth.sse_batch).  Both with the
    usage event and [DONE], 1..k events per chunk or recut at arbitrary bytes."""
    import numpy as np
    from llmapigateway_b200 import synth
    rng = random.Random(seed)
    if gen == "openai":
        nrng = np.random.default_rng(seed)
        truths = synth._usage_truths(n, seed)
        out = []
        for s_ in range(n):
            epc = rng.choice([(1, 1), (1, 4), (2, 2), (1, 9)])
            out.append(_recut(rng, synth.openai_stream(nrng, rng.randrange(1, 60), truths[s_], "chatcmpl-%08x" % rng.randrange(2**32), epc)))
        return out
    n_events = rng.choice([8, 16, 48, 96])
    b = synth.sse_batch(n_streams=n, n_events=n_events, seed=seed, events_per_chunk=rng.choice([1, 1, 2, 8]),
                        with_usage=rng.random() < 0.9, with_done=rng.random() < 0.9)
    return [_recut(rng, b.stream_chunks(s_)) for s_ in range(n)]


SKELETONS = [   # %n = a number spelling, %s = a string value; every batch repeats a few of these so that they become templates
    '{"choices":[],"usage":{"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n},"model":%s,"provider":%s}',
    '{"choices": [], "usage": {"prompt_tokens": %n, "completion_tokens": %n, "total_tokens": %n, "cost": %n}, "model": %s}',
    '{"usage":{"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n,"cost":%n},"choices":[{"index":0,"delta":{"content":%s},"finish_reason":null}]}',
    '{"id":%s,"choices":[{"index":0,"delta":{},"finish_reason":"stop","usage":{"prompt_tokens":%n}}],"usage":{"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n}}',
    '{"x":{"usage":{"prompt_tokens":%n,"completion_tokens":%n}},"choices":[],"usage":{"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n},"model":%s}',
    '{"choices":[],"usage":{"prompt_tokens":%n,"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n},"model":%s}',
    '{"choices":[],"usage":{"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n},"usage":{"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n},"model":%s}',
    '{"choices":[],"model":%s,"model":%s,"usage":{"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n}}',
    '{"choices":[],"usage":{"completion_tokens_details":{"reasoning_tokens":%n,"prompt_tokens":%n},"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n,"prompt_tokens_details":{"cached_tokens":%n,"cost":%n}},"model":%s}',
    '{"choices":[],"usage":{"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n,"completion_tokens_details":null,"prompt_tokens_details":{"cached_tokens":%n}},"provider":%s}',
    '{"choices":[],"usage":null,"model":%s,"provider":%s,"cost":%n}',
    '{"choices":[],"usage":{"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n,"cost":%n,"cost_details":{"upstream_inference_cost":%n},"is_byok":false},"model":%s,"provider":%s}',
    '{"choices":[{"index":0,"message":{"role":"assistant","content":%s}}],"usage":{"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n}}',
    '{"choices":[],"usage":{"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n,"reasoning_tokens":%n,"cached_tokens":%n},"model":%s}',
    '{"choices":[],"usage":{"total_tokens":%n,"completion_tokens":%n,"prompt_tokens":%n,"cost":%n},"provider":%s,"model":%s}',
    '{"model":%s,"usage":{"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n,"completion_tokens_details":{"reasoning_tokens":%n,"accepted_prediction_tokens":%n},"prompt_tokens_details":{"cached_tokens":%n,"audio_tokens":%n}},"choices":[],"provider":%s}',
    '{"choices":[],"usage":{"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n},"model":{"name":%s},"provider":[%s]}',
    '{"choices":[],"usage":[{"prompt_tokens":%n}],"model":%s}',
    '{"choices":[],"usage":{"prompt_tokens":{"v":%n},"completion_tokens":[%n,%n],"total_tokens":%n},"model":%s}',
    '{"error":{"message":%s,"code":%n},"usage":{"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n}}',
    '{"choices":[],"usage":{"prompt_tokens":%n,"completion_tokens":%n,"total_tokens":%n},"code":%n,"model":%s}',
]


def skeleton_streams(n_streams: int, seed: int):
    """Usage events of odd shapes, each shape repeated across the batch with other values (so that it is learnt as a template and
    then FOLLOWED): keys in other places and orders, repeated keys, usage-like keys where get_token_usage does not look, containers
    and null where numbers go, spaces inside the event."""
    import re
    import sse_cases as sc
    rng = random.Random(seed)
    shapes = rng.sample(SKELETONS, rng.randrange(2, 5))
    out = []
    for s_ in range(n_streams):
        evs = [sc.delta(rng.choice(["a", "hello", "x" * rng.randrange(1, 70)])) for _ in range(rng.randrange(1, 16))]
        for _ in range(1 if rng.random() < 0.85 else 2):
            shape = rng.choice(shapes) if rng.random() < 0.93 else rng.choice(SKELETONS)

            def fill(m):
                r = rng.random()
                if m.group(0) == "%n":
                    return rng.choice(NUMBERS) if r < 0.9 else (rng.choice(BAD_NUMBERS) if r < 0.94 else rng.choice(["null", "true", '"7"', "[1]", "{}"]))
                return '"' + (rng.choice(STRINGS) if r < 0.93 else rng.choice(BAD_STRINGS)) + '"' if r < 0.97 else rng.choice(["null", "12", "[]"])
            ev = re.sub(r"%[ns]", fill, shape)
            evs.append(("data: " + ev + "\n\n").encode("utf-8", errors="surrogatepass"))
            if rng.random() < 0.25:
                evs.append(sc.delta("more"))
        if rng.random() < 0.8:
            evs.append(sc.DONE)
        out.append(_recut(rng, evs))
    return out


def make_streams(gen: str, n: int, seed: int):
    import test_sse_gpu as G
    if gen in ("openai", "c3"):
        return shaped_streams(gen, n, seed)
    if gen == "skeleton":
        return skeleton_streams(n, seed)
    if gen == "random":
        return G._random_streams(n, seed)
    if gen == "template":
        return G._template_variant_streams(n, seed)
    return usage_variant_streams(n, seed)


def one_round(job):
    gen, seed, n_steps, geo, n_streams, cold = job
    import test_sse_gpu as G
    from golden_io import canon_rows
    from host_relay import HostBulkEngine
    from llmapigateway_b200 import _abi
    from oracle.sse_oracle import run_stream
    tag = f"{gen}:{seed}:{n_steps}:{geo}:{int(cold)}"
    try:
        nb, tpw = GEOMETRIES[geo]
        cap = int(os.environ.get("LGW_FUZZ_CARRY_CAP", "4096"))     # a small carry capacity stresses the overflow paths (bulk vs sequential only)
        eng = HostBulkEngine(max_streams=2048, carry_cap=cap, n_blocks=nb, tiles_per_warp=tpw)
        try:
            streams = make_streams(gen, n_streams, seed)
            if not cold:                      # warm templates: a first batch of the same generator, other seed
                G._run_all(eng, make_streams(gen, 64, seed + 1), 0, 1, seed=3)
            s_fast, r_fast, e_fast = G._run_all(eng, streams, 0, n_steps, seed=seed)
            counters = eng.counters()
            s_seq, r_seq, e_seq = G._run_all(eng, streams, 1, n_steps, seed=seed)
            n_ok = 0
            for i, (a, b) in enumerate(zip(s_fast, s_seq)):
                where = f"{tag} stream {i}"
                assert e_fast[i] == e_seq[i], where + " emitted bytes: bulk != sequential"
                assert bytes(a)[:64] == bytes(b)[:64], where + " state header: bulk != sequential"
                if a.flags & _abi.SF_REC_VALID:
                    assert a.rec.exotic == b.rec.exotic, where + " exotic flag"
                    if not a.rec.exotic:
                        assert canon_rows([_abi.usage_rec_to_dict(a.rec)]) == canon_rows([_abi.usage_rec_to_dict(b.rec)]), where + " usage record: bulk != sequential"
                if (a.flags | b.flags) & _abi.SF_CARRY_OVERFLOW:      # an event outgrew the engine's carry capacity: the reference buffers without
                    continue                                           #   bound, so only bulk == sequential is checked for such a stream
                relay, tap = run_stream(streams[i])
                assert e_fast[i] == relay.emitted, where + " emitted bytes != oracle"
                assert (a.phase == _abi.PHASE_FAILED) == relay.failed, where + " verdict != oracle"
                if not relay.failed and not a.n_exotic:
                    got = [json.loads(r[2])[0] for r in r_fast if r[0] == i]
                    if a.flags & _abi.SF_EMITTED_ANY:
                        got.append(json.loads(canon_rows([_abi.usage_rec_to_dict(a.rec)]))[0])
                    assert canon_rows(got) == canon_rows(tap.rows), where + " usage rows != oracle"
                    assert (not (a.flags & _abi.SF_A_USAGE_BOUND)) == relay.end_raises, where + " end_raises != oracle"
                    n_ok += 1
            assert r_fast == r_seq, tag + " mid-stream rows: bulk != sequential"
            return tag, True, f"{n_ok}/{n_streams} compared with the oracle; {counters}"
        finally:
            eng.close_engine()
    except Exception:
        return tag, False, traceback.format_exc(limit=3)


def transcript_round(job):
    """The same generators through the transcript tap (k_text_extract / scan / pack on the emulator, TranscriptBook on the host)
    against the oracle's llm_response_accum at every write_log call."""
    gen, seed, n_steps, geo, n_streams, cold = job
    import test_transcript_gpu as T
    from host_relay import HostBulkEngine
    from llmapigateway_b200 import _abi
    from llmapigateway_b200.transcripts import decode_text
    from oracle.sse_oracle import run_stream
    tag = f"{gen}:{seed}:{n_steps}:{geo}:{int(cold)}"
    try:
        nb, tpw = GEOMETRIES[geo]
        eng = HostBulkEngine(max_streams=2048, carry_cap=8192, n_blocks=nb, tiles_per_warp=tpw)
        eng.enable_transcripts()
        try:
            streams = make_streams(gen, n_streams, seed)
            rng = random.Random(seed)
            scheds = [[0] + sorted(set(rng.randrange(1, max(2, len(c))) for _ in range(n_steps - 1))) if len(c) > 1 else [0] for c in streams]
            res = T.run_streams(eng, streams, scheds)
            n_cmp = 0
            for i, (chunks, (snaps, flags, emitted_any, seq_steps)) in enumerate(zip(streams, res)):
                relay, tap = run_stream(chunks)
                where = f"{tag} stream {i}"
                if relay.failed:
                    assert snaps == [], where + " transcript of a failed stream"
                    continue
                if flags & _abi.TF_EXOTIC:
                    continue
                assert [decode_text(t) for t in snaps] == tap.transcripts, where + " transcripts != oracle"
                n_cmp += 1
            return tag, True, f"{n_cmp}/{n_streams} transcripts compared with the oracle"
        finally:
            eng.close_engine()
    except Exception:
        return tag, False, traceback.format_exc(limit=3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=20)
    ap.add_argument("--procs", type=int, default=4)
    ap.add_argument("--seed", type=int, default=1000)
    ap.add_argument("--streams", type=int, default=300)
    ap.add_argument("--replay", default=None)
    ap.add_argument("--gens", default=None, help="comma-separated subset of the generators")
    ap.add_argument("--transcripts", action="store_true", help="check the transcript tap instead of the relay / usage results")
    args = ap.parse_args()
    if args.replay:
        gen, seed, n_steps, geo, *cold = args.replay.split(":")
        fn = transcript_round if args.transcripts else one_round
        print(*fn((gen, int(seed), int(n_steps), int(geo), args.streams, bool(cold and int(cold[0])))), sep="\n")
        return 0
    if args.gens:
        GENERATORS[:] = args.gens.split(",")
    rng = random.Random(args.seed)
    jobs = []
    for r in range(args.rounds):
        jobs.append((rng.choice(GENERATORS), args.seed * 1000 + r, rng.choice([1, 1, 2, 3, 5]), rng.randrange(len(GEOMETRIES)), args.streams, rng.random() < 0.25))
    from host_relay import lib
    lib()                                   # build the emulator library once, before the workers start
    # every round in a process of its own: the emulator ABORTS on a divergent collective or a deadlock (a finding, not a crash of
    # the campaign), and a pool worker that dies takes its job with it
    import subprocess
    from concurrent.futures import ThreadPoolExecutor

    def spawn(job):
        gen, seed, n_steps, geo, n_streams, cold = job
        tag = f"{gen}:{seed}:{n_steps}:{geo}:{int(cold)}"
        r = subprocess.run([sys.executable, __file__, "--replay", tag, "--streams", str(n_streams)] + (["--transcripts"] if args.transcripts else []),
                           capture_output=True, text=True)
        lines = [l for l in r.stdout.splitlines() if l.strip()]
        if r.returncode != 0 or len(lines) < 3:
            return tag, False, "process ended with status %d: %s" % (r.returncode, (r.stderr or r.stdout)[-400:])
        return tag, lines[1] == "True", "\n".join(lines[2:])
    bad = 0
    with ThreadPoolExecutor(args.procs) as pool:
        for tag, ok, msg in pool.map(spawn, jobs):
            print(("ok   " if ok else "FAIL ") + tag + "  " + msg, flush=True)
            bad += 0 if ok else 1
    print(f"{len(jobs) - bad}/{len(jobs)} rounds clean")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
