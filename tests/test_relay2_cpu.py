"""The bulk kernels of csrc/relay2.cuh (k_prime2 -> k_relay2 -> k_commit2) on the CPU box: the kernel
source is compiled with g++ over a SIMT emulator (tests/support/simt_emu.h, host_relay2.cpp) and driven through the
same test bodies as the GPU parity tests (tests/test_sse_gpu.py): goldens of the unmodified reference, the oracle, and
the exact sequential machine.  Test aid only -- the product has no CPU path."""
import numpy as np
import pytest

import test_sse_gpu as G
from golden_io import canon_rows
from host_relay import HostBulkEngine
from llmapigateway_b200 import _abi
from llmapigateway_b200.synth import pack_streams, sse_batch


@pytest.fixture(scope="module", params=[(2, 0), (3, 1)], ids=["spread", "one_tile_per_warp"])
def engine(request):
    n_blocks, tpw = request.param
    e = HostBulkEngine(max_streams=2048, n_blocks=n_blocks, tiles_per_warp=tpw)
    yield e
    e.close_engine()


@pytest.mark.parametrize("stepping", ["one_step", "step_per_chunk", "random_steps"])
def test_golden_cases(engine, stepping):
    G.test_golden_cases_batched(engine, 0, stepping)


def test_golden_cases_cold_templates(engine):
    engine.reset_templates()
    G.test_golden_cases_batched(engine, 0, "one_step")


@pytest.mark.parametrize("events_per_chunk", [1, 8])
def test_c3_small(engine, events_per_chunk):
    G.test_c3_small_vs_oracle(engine, 0, events_per_chunk, n_streams=40)


def test_c3_mid_event_cut(engine):
    G.test_c3_two_steps_with_mid_event_cut(engine)


@pytest.mark.parametrize("n_steps", [1, 3])
def test_bulk_vs_sequential_vs_oracle(engine, n_steps):
    G.test_bulk_path_equals_sequential_path_and_oracle(engine, n_steps, n_streams=500, min_regular=300)


@pytest.mark.parametrize("n_steps", [1, 2])
def test_template_shortcut(engine, n_steps):
    G.test_template_shortcut_is_exact(engine, n_steps, n_streams=300, min_rows=200)


@pytest.mark.parametrize("events_per_chunk", [(1, 1), (1, 4)])
def test_openai_shaped(engine, events_per_chunk):
    G.test_openai_shaped_streams_vs_oracle(engine, 0, events_per_chunk, n_streams=24)


def test_usage_fields_come_from_the_template_spans(engine):
    """C3 usage events follow a usage template after the first step: the record must then come from the matched value
    spans (no stash), and equal the oracle's row whatever the digits, exponents and string lengths are."""
    from oracle.sse_oracle import run_stream
    engine.reset_templates()
    b = sse_batch(n_streams=32, n_events=16, seed=21)
    engine.open(b.seg_slot); engine.step(b.data, b.chunk_off, b.seg_chunk, b.seg_slot); engine.close(b.seg_slot)     # templates learnt
    assert any(t["usage_ok"] for t in engine.templates())
    b = sse_batch(n_streams=64, n_events=16, seed=22)
    engine.open(b.seg_slot)
    c0 = engine.counters()
    engine.step(b.data, b.chunk_off, b.seg_chunk, b.seg_slot)
    c1 = engine.counters()
    assert c1["from_template"] - c0["from_template"] == 64 and c1["stashed"] == c0["stashed"] and c1["sequential"] == c0["sequential"]
    states = engine.state(b.seg_slot)
    for s in range(64):
        relay, tap = run_stream(b.stream_chunks(s))
        assert canon_rows([_abi.usage_rec_to_dict(states[s].rec)]) == canon_rows(tap.rows)
    engine.close(b.seg_slot)


def test_verdicts_only_step(engine):
    G.test_verdicts_only_step_equals_the_full_step(engine)


def test_container_valued_usage_fields(engine):
    G.test_container_valued_usage_fields_are_never_read_from_a_span(engine)


def test_lanes_agree_on_a_segment_another_warp_flags():
    """Regression (tools/fuzz_relay2_cpu.py): `plan[seg].irregular` is written by other warps of the launch; when every lane read
    it on its own, lanes could see different values and part ways before the next collective (the emulator aborts the process on
    that: "divergent collective").  This batch on this geometry is one the campaign found; it runs in a process of its own."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys
        sys.path[:0] = [%r, %r]
        import test_sse_gpu as G
        from host_relay import HostBulkEngine
        eng = HostBulkEngine(max_streams=2048, n_blocks=3, tiles_per_warp=1)
        G._run_all(eng, G._template_variant_streams(64, 11211), 0, 1, seed=3)
        print("clean")
    """) % (str(G.__file__).rsplit("/tests/", 1)[0], str(G.__file__).rsplit("/", 1)[0])
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "clean" in r.stdout, r.stderr[-500:]
