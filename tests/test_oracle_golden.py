"""The oracle restatement against outputs of the UNMODIFIED reference (tests/golden/sse_cases.json).
CPU only."""
import pytest

from golden_io import UNPINNED_DETAIL_PREFIX, canon_rows, load_sse_cases
from oracle.sse_oracle import run_stream, split_events

CASES = load_sse_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference(case):
    relay, tap = run_stream(case["chunks"], case["http_status"])
    assert relay.failed == case["failed"]
    if case["failed"] and case["error_detail"].startswith(UNPINNED_DETAIL_PREFIX):
        # the tail of this message is the JSON library's exception text (request_handler.py:185)
        assert relay.error_detail.startswith(UNPINNED_DETAIL_PREFIX)
    else:
        assert relay.error_detail == case["error_detail"]
    assert relay.emitted == case["emitted"]
    assert relay.end_raises == case["end_raises"]
    assert canon_rows(tap.rows) == case["rows"]
    assert tap.transcripts == case["transcripts"]


def test_golden_has_the_adversarial_set():
    names = {c["name"] for c in CASES}
    for must in ("leading_comment_chunk", "first_event_split", "utf8_split_in_relay", "crlf_delimiters",
                 "first_event_error", "first_event_detail", "midstream_error_with_code", "usage_null",
                 "ctd_null", "no_usage_at_all", "duplicate_usage_events", "data_two_spaces", "data_no_space"):
        assert must in names
    assert sum(n.startswith("fuzz_") for n in names) >= 100


def test_split_rule():
    assert split_events("a\n\nb") == (["a"], "b")
    assert split_events("a\n\n") == (["a", ""], "")
    assert split_events("a\n\n\n") == (["a", "\n"], "")
    assert split_events("a\n") == ([], "a\n")


def test_nonstream_tap_oracle_matches_the_reference():
    """chat_logging.py:98-150 with is_real_streaming=False: rows of tests/golden/response_cases.json (made by the unmodified
    ChunkProcessorThread) against the restatement oracle.sse_oracle.tap_nonstream."""
    import base64
    import json
    from golden_io import GOLDEN, canon_rows
    from oracle.sse_oracle import tap_nonstream
    doc = json.loads((GOLDEN / "response_cases.json").read_text())
    assert len(doc["tap_cases"]) >= 20
    for c in doc["tap_cases"]:
        chunks = [base64.b64decode(x) for x in c["chunks"]] if "chunks" in c else ([base64.b64decode(c["text"])] if c["text"] else [])
        assert canon_rows(tap_nonstream(chunks).rows) == canon_rows(c["rows"]), base64.b64decode(c["text"])[:80]
    n = 0
    for c in doc["cases"]:
        if "tap_rows" in c:
            assert canon_rows(tap_nonstream([base64.b64decode(c["body"])]).rows) == canon_rows(c["tap_rows"])
            n += 1
    assert n >= 8


def test_chain_oracle_matches_the_reference():
    """oracle.chain_oracle.walk against the goldens of the unmodified chat.py:20 `chat_completions` (config 4): relayed bytes,
    503/400 status and detail, and the url / wire body / headers of every upstream attempt."""
    import chain_cases as cc
    from oracle import chain_oracle
    from llmapigateway_b200 import synth
    doc, ups = cc.load()
    providers, rules, fallback_provider = synth.chain_world()
    rot = chain_oracle.Rotation()
    kinds = set()
    for case in doc["cases"]:
        up, sid = ups[case["group"]], case["sid"]
        headers = {"Authorization": f"Bearer {case['api_key']}"} if case["api_key"] else {}
        got = chain_oracle.walk(cc.D64(case["body"]), headers, providers, rules, fallback_provider, lambda a: up.stream_chunks(sid, a), rot, cc.stream_mode())
        cc.check_against_golden(case, got)
        kinds.add((case["kind"], case.get("status"), len(case["attempts"])))
    assert ("http_exception", 503, 3) in kinds and ("stream", None, 1) in kinds and ("stream", None, 3) in kinds
