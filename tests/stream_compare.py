"""Compare an engine-style result (segment results + final state + row events) with the oracle /
golden expectation for one stream.  Shared by the CPU host-machine tests and the GPU tests."""
from __future__ import annotations

from golden_io import UNPINNED_DETAIL_PREFIX, canon_rows
from llmapigateway_b200 import _abi


def emitted_from_segments(chunks, step_chunk, segs):
    out = []
    for k, seg in enumerate(segs):
        end = int(step_chunk[k + 1])
        out += [c for c in chunks[int(seg.emit_chunk_begin):end] if c]
    return out


def rows_from_result(state, row_events):
    rows = [_abi.usage_rec_to_dict(ev.rec) for ev in sorted(row_events, key=lambda e: e.seq)]
    if state.flags & _abi.SF_EMITTED_ANY:
        rows.append(_abi.usage_rec_to_dict(state.rec))
    return rows


def check_stream(expect: dict, chunks, step_chunk, segs, state, detail: bytes, row_events, label=""):
    """expect: dict(failed, error_detail, emitted, end_raises, rows(canonical str), http_status)."""
    failed = state.phase == _abi.PHASE_FAILED
    assert failed == expect["failed"], label
    if failed:
        if expect.get("http_status", 200) >= 400:
            assert state.verdict == _abi.VERDICT_FAIL_HTTP, label
        elif expect["error_detail"].startswith(UNPINNED_DETAIL_PREFIX):
            assert state.verdict == _abi.VERDICT_FAIL_PARSE, label
        else:
            assert state.verdict == _abi.VERDICT_FAIL_EVENT, label
            assert detail.decode("utf-8") == expect["error_detail"], label
        assert emitted_from_segments(chunks, step_chunk, segs) == [], label
        return "failed"
    assert emitted_from_segments(chunks, step_chunk, segs) == expect["emitted"], label
    assert (not (state.flags & _abi.SF_A_USAGE_BOUND)) == expect["end_raises"], label
    if state.n_exotic:
        return "exotic"          # reported-but-unmodelled shape: rows are not compared
    assert canon_rows(rows_from_result(state, row_events)) == expect["rows"], label
    return "ok"
