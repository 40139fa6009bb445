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
