"""CPU oracle for the streaming relay path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package. The product (llmapigateway_b200/) never does; it fails loudly when the
CUDA library is missing.

What is restated (reference = fabiojbg/LLMApiGateway, paths relative to /root/reference):

  RelayOracle      llm_gateway_core/services/request_handler.py:21-150
                   (stream_generator :21-63, priming loop :65-98, combined_generator :100-144)
  TapOracle        llm_gateway_core/middleware/chat_logging.py:87-150 (ChunkProcessorThread.run)
  token_usage      llm_gateway_core/middleware/chat_logging.py:233-272 (get_token_usage)

The reference is pure Python, so the restatement is Python too: `str.split`, `str.startswith`
and a JSON parser do the byte work exactly as the reference does.  The third-party parser on
this path is `json5` (PyPI, unpinned: requirements.txt:8) which is absent from this image and
from the GPU box; `loads` defaults to the stdlib `json.loads`, which agrees with `json5.loads`
on every strict RFC-8259 text (JSON5 is a superset), so parity is pinned for strict-JSON
inputs and UNPINNED for JSON5-only syntax (comments, single quotes, unquoted keys ...).

Pinning: tests/test_oracle_golden.py checks this file against tests/golden/sse_cases.json,
which was produced by running the UNMODIFIED reference modules (tests/golden/make_golden.py,
dev container only, library versions recorded in the fixture header).

The restatement is synchronous: the reference's three nested async generators advance in
lock-step (one upstream chunk at a time), so a plain loop visits the same states in the same
order.  It therefore runs a little FASTER than the reference (no asyncio, no logging calls),
which makes it a generous CPU baseline.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Callable, Iterable, Iterator

EVENT_SEP = "\n\n"
REAL_PREFIX = "data: {"       # request_handler.py:42,83,119
DATA_PREFIX = "data: "        # request_handler.py:49,85,122 ; chat_logging.py:120-121


def split_events(buffer: str) -> tuple[list[str], str]:
    """The carry rule every loop of the reference shares (request_handler.py:38-40,77-79,
    113-115; chat_logging.py:110-112): complete pieces, then the new carry."""
    pieces = buffer.split(EVENT_SEP)
    if buffer.endswith(EVENT_SEP):
        return pieces, ""          # trailing "" (or a stray "\n") stays in pieces, carry resets
    carry = pieces.pop()
    return pieces, carry


@dataclass
class RelayResult:
    failed: bool                       # make_llm_request returned (None, detail)
    error_detail: str | None
    emitted: list[bytes] = field(default_factory=list)
    end_raises: bool = False           # request_handler.py:144 UnboundLocalError (no usage ever bound)
    handler_usage: object = None       # the value bound at request_handler.py:134, if any
    first_kept_index: int | None = None  # index (into the input list) of the chunk kept by priming


class RelayOracle:
    """One upstream attempt with is_streaming=True."""

    def __init__(self, loads: Callable[[str], object] = json.loads,
                 url: str = "http://upstream.test/v1/chat/completions"):
        self.loads = loads
        self.url = url

    # -- stream_generator, request_handler.py:21-63 ------------------------------------
    def _sniff(self, chunks: Iterable[bytes], http_status: int, box: dict) -> Iterator[tuple[int, bytes]]:
        if http_status >= 400:                                   # :25-30
            box["detail"] = b"".join(chunks).decode("utf-8")
            box["error"] = True
            return
        text_buf = ""
        awaiting_first = True
        for idx, raw in enumerate(chunks):
            try:
                text_buf += raw.decode("utf-8")                  # :36-37
                pieces, text_buf = split_events(text_buf)
                for piece in pieces:
                    if not piece.startswith(REAL_PREFIX):        # :42
                        continue
                    if awaiting_first:                            # :46-54
                        awaiting_first = False
                        doc = self.loads(piece[len(DATA_PREFIX):])
                        if "error" in doc or "detail" in doc:
                            box["detail"] = piece
                            box["error"] = True
                            return
            except Exception:                                     # :55-58 swallowed
                pass
            if raw:                                               # :60-63
                yield idx, raw

    def run(self, chunks: list[bytes], http_status: int = 200) -> RelayResult:
        box = {"error": False, "detail": None}
        try:
            return self._run(chunks, http_status, box)
        except Exception as exc:                                  # :183-187
            return RelayResult(True, f"Unexpected error during request to {self.url}: {exc}")

    def _run(self, chunks, http_status, box) -> RelayResult:
        source = self._sniff(chunks, http_status, box)
        kept: tuple[int, bytes] | None = None
        text_buf = ""
        # -- priming loop, request_handler.py:65-95 -------------------------------------
        for idx, raw in source:
            try:
                text_buf += raw.decode("utf-8")
            except UnicodeDecodeError:                            # :94-95 chunk dropped
                continue
            pieces, text_buf = split_events(text_buf)
            hit = False
            for piece in pieces:
                if piece.startswith(REAL_PREFIX):
                    hit = True
                    doc = self.loads(piece[len(DATA_PREFIX):])   # :85 (errors escape to :183)
                    if "error" in doc or "detail" in doc:
                        box["detail"] = piece
                        box["error"] = True
                    else:
                        kept = (idx, raw)
                    break
            if hit:
                break
        if box["error"]:                                          # :97-98
            return RelayResult(True, box["detail"])

        # -- combined_generator, request_handler.py:100-144 -----------------------------
        res = RelayResult(False, None)
        if kept is not None:
            res.first_kept_index = kept[0]
            res.emitted.append(kept[1])
        usage_bound = False
        text_buf = ""                                             # :108 starts empty again
        for idx, raw in source:
            try:
                text_buf += raw.decode("utf-8")
                pieces, text_buf = split_events(text_buf)
                for piece in pieces:
                    if not piece.startswith(REAL_PREFIX):
                        continue
                    try:
                        doc = self.loads(piece[len(DATA_PREFIX):])
                        if "code" in doc:
                            # :123-131 -- line :129 formats an unbound local `e`; the
                            # UnboundLocalError lands in the handler at :135, so nothing
                            # after it runs for this piece.
                            raise UnboundLocalError("e")
                        if "usage" in doc:                        # :133-134
                            res.handler_usage = doc.get("usage")
                            usage_bound = True
                    except Exception:
                        pass
            except Exception:                                     # :138-139
                pass
            res.emitted.append(raw)                               # :141-142 original bytes
        res.end_raises = not usage_bound                          # :144
        return res


# ---------------------------------------------------------------------------------------
# Response tap: ChunkProcessorThread.run (streaming branch) + get_token_usage
# ---------------------------------------------------------------------------------------

USAGE_DEFAULTS = ("prompt_tokens", "completion_tokens", "total_tokens",
                  "reasoning_tokens", "cached_tokens", "cost")


def token_usage(doc) -> dict:
    """chat_logging.py:233-272.  Same order of operations, so the partially-filled dict that
    comes back when an exception interrupts it is the same."""
    rec = {k: 0 for k in USAGE_DEFAULTS}                          # :237-244
    try:
        if "usage" in doc and isinstance(doc["usage"], dict):
            u = doc["usage"]
            for name in ("prompt_tokens", "completion_tokens", "total_tokens", "cost"):  # :248-255
                if name in u:
                    rec[name] = u[name]
            for outer, inner, dest in (("completion_tokens_details", "reasoning_tokens", "reasoning_tokens"),
                                       ("prompt_tokens_details", "cached_tokens", "cached_tokens")):  # :256-261
                if outer in u and inner in u[outer]:
                    rec[dest] = u[outer][inner]
            if rec["reasoning_tokens"] > 0:                       # :262-263
                rec["completion_tokens"] = rec["completion_tokens"] - rec["reasoning_tokens"]
        for name in ("provider", "model"):                        # :264-267
            if name in doc:
                rec[name] = doc[name]
    except Exception:                                             # :268-270
        pass
    return rec


@dataclass
class TapResult:
    rows: list[dict]            # tokens_usage at each write_log call (:139, :150) -> insert_usage
    transcripts: list[str]      # llm_response_accum at those calls


class TapOracle:
    """chat_logging.py:87-150 with is_real_streaming=True, fed the chunks the relay emitted.
    The 5 s idle wait (:94) is not modelled: it only delays the final row."""

    def __init__(self, loads: Callable[[str], object] = json.loads):
        self.loads = loads

    def run(self, emitted: list[bytes]) -> TapResult:
        out = TapResult([], [])
        if not emitted:                      # no first chunk => no thread (:198-203) => no row
            return out
        usage = {k: 0 for k in USAGE_DEFAULTS}                    # :77-84
        accum = ""
        text_buf = ""
        for raw in emitted:
            try:
                text_buf += raw.decode("utf-8")                   # :108-112
                pieces, text_buf = split_events(text_buf)
                for piece in pieces:
                    try:
                        if not (piece.startswith(REAL_PREFIX) or piece.startswith("{")):  # :116-118
                            continue
                        if piece.startswith(DATA_PREFIX):          # :120-121
                            piece = piece[len(DATA_PREFIX):].strip()
                        doc = self.loads(piece)                    # :123
                        if "choices" in doc:                       # :124-133
                            for choice in doc["choices"]:
                                if "delta" in choice and "content" in choice["delta"]:
                                    frag = choice["delta"]["content"]
                                    if frag:
                                        accum += frag
                                elif "message" in choice and "content" in choice["message"]:
                                    frag = choice["message"]["content"]
                                    if frag:
                                        accum += frag
                        if "usage" in doc:                         # :134-135
                            usage = token_usage(doc)
                        if "error" in doc:                         # :137-139
                            accum += piece
                            out.rows.append(dict(usage)); out.transcripts.append(accum)
                    except Exception:                              # :140-141
                        pass
            except Exception:                                      # :142-143
                pass
        out.rows.append(dict(usage)); out.transcripts.append(accum)  # :150
        return out


def tap_nonstream(chunks: list[bytes], loads: Callable[[str], object] = json.loads) -> TapResult:
    """chat_logging.py:87-150 with is_real_streaming=False (every response that is not text/event-stream, :188): the chunks are
    concatenated first (:98-103), the whole text is ONE part (:105-106), then the same per-part code as the streaming branch."""
    out = TapResult([], [])
    if not chunks:                           # no first chunk => no thread (:198-203) => no row
        return out
    usage = {k: 0 for k in USAGE_DEFAULTS}                        # :77-84
    accum = ""
    try:
        buffer = ""
        for raw in chunks:
            buffer += raw.decode("utf-8")                         # :100-101 (a decode error escapes the thread's loop: no row at all)
    except UnicodeDecodeError:
        return out
    piece = buffer
    try:
        if piece.startswith(REAL_PREFIX) or piece.startswith("{"):    # :116-118
            if piece.startswith(DATA_PREFIX):                          # :120-121
                piece = piece[len(DATA_PREFIX):].strip()
            doc = loads(piece)                                         # :123
            if "choices" in doc:                                       # :124-133
                for choice in doc["choices"]:
                    if "delta" in choice and "content" in choice["delta"]:
                        frag = choice["delta"]["content"]
                        if frag:
                            accum += frag
                    elif "message" in choice and "content" in choice["message"]:
                        frag = choice["message"]["content"]
                        if frag:
                            accum += frag
            if "usage" in doc:                                         # :134-135
                usage = token_usage(doc)
            if "error" in doc:                                         # :137-139
                accum += piece
                out.rows.append(dict(usage)); out.transcripts.append(accum)
    except Exception:                                                  # :140-141
        pass
    out.rows.append(dict(usage)); out.transcripts.append(accum)       # :150
    return out


def run_stream(chunks: list[bytes], http_status: int = 200, loads=json.loads):
    """Whole per-stream hot path: relay then tap. Returns (RelayResult, TapResult)."""
    relay = RelayOracle(loads).run(chunks, http_status)
    tap = TapOracle(loads).run(relay.emitted) if not relay.failed else TapResult([], [])
    return relay, tap
