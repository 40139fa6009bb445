"""Chat transcript log: the host half of the reference's `write_log` (chat_logging.py:22-67).

The device (csrc/transcript.cuh) does the per-byte work of the tap -- split, parse, choices walk, JSON string decoding -- and
hands back, per step, the text every stream appended plus the positions of mid-stream `write_log` calls.  What stays here is
what the reference does once per response: `TranscriptBook` keeps each stream's `llm_response_accum`, `TranscriptLog` renders
the log file (same name, same blocks, same `\\n` replacement), inserts the usage row AFTER the file was written -- a failing
write skips the row exactly like the reference's single try block (:23-67) -- and prunes ./logs (:58-65).
"""
from __future__ import annotations

import glob
import logging
import os
from datetime import datetime
from pprint import pformat

import numpy as np

from . import _abi

logger = logging.getLogger(__name__)


class TranscriptBook:
    """llm_response_accum of every open stream, fed from `Engine.step_transcript()` results."""

    def __init__(self):
        self._text: dict[int, bytearray] = {}
        self._flags: dict[int, int] = {}

    def open(self, slot: int):
        self._text[int(slot)] = bytearray()
        self._flags[int(slot)] = 0

    def apply(self, seg_slot, step_text) -> list:
        """Append what the step's segments produced; returns [(slot, seq, text_so_far: str-or-bytes)] for the mid-stream
        write_log calls of the step (chat_logging.py:139), ordered per stream by seq."""
        off = step_text.seg_off
        raw = step_text.text
        for s, slot in enumerate(np.asarray(seg_slot).tolist()):
            a, b = int(off[s]), int(off[s + 1])
            buf = self._text.get(slot)
            if buf is None:
                buf = self._text[slot] = bytearray()
                self._flags[slot] = 0
            if b > a:
                buf += raw[a:b].tobytes()
            self._flags[slot] |= int(step_text.flags[s]) & ~_abi.TF_SEQUENTIAL
        marks = []
        for slot, seq, pos in sorted(step_text.marks, key=lambda m: (m[0], m[1])):
            marks.append((slot, seq, bytes(self._text[slot][:pos])))
        return marks

    def flags(self, slot: int) -> int:
        return self._flags.get(int(slot), 0)

    def text(self, slot: int) -> bytes:
        return bytes(self._text.get(int(slot), b""))

    def close(self, slot: int) -> tuple[bytes, int]:
        slot = int(slot)
        return bytes(self._text.pop(slot, b"")), self._flags.pop(slot, 0)


def decode_text(raw: bytes) -> str:
    """The Python str the reference holds: lone surrogates (json.loads keeps them) come back through 'surrogatepass'."""
    return raw.decode("utf-8", "surrogatepass")


def render_log(req_headers, req_body_str: str, llm_response_accum: str, tokens_usage: dict) -> str:
    """The log text of chat_logging.py:27-42, before the `\\n` replacement of :49."""
    division_line = "-" * 100
    model = f"Model: {tokens_usage['model']}\n" if "model" in tokens_usage else ""
    provider = f"Provider: {tokens_usage['provider']}\n\n" if "provider" in tokens_usage else ""
    return (
        f"{division_line}\nTokens Usage:\n-{division_line}\n\n"
        f"Input: {tokens_usage['prompt_tokens']}\n"
        f"Output: {tokens_usage['completion_tokens']}\n"
        f"Cached: {tokens_usage['cached_tokens']}\n"
        f"Reasoning: {tokens_usage['reasoning_tokens']}\n"
        f"Total: {tokens_usage['total_tokens']}\n"
        f"Cost: ${tokens_usage['cost']:0.6f}\n"
        f"{model}"
        f"{provider}"
        f"{division_line}\nRequest Headers:\n{division_line}\n\n{pformat(req_headers, indent=2)}\n\n"
        f"{division_line}\nRequest Body:\n-{division_line}\n\n{req_body_str}\n\n"
        f"{division_line}\nLLM Response:\n{division_line}\n\n{llm_response_accum}"
    )


class TranscriptLog:
    """write_log (chat_logging.py:22-67): file, then usage row, then pruning; never raises."""

    def __init__(self, log_dir: str = "./logs", log_file_limit: int | None = 50, usage_sink=None, clock=datetime.now):
        self.log_dir = log_dir
        self.log_file_limit = log_file_limit
        self.usage_sink = usage_sink            # callable(dict) or object with insert_usage(dict)
        self.clock = clock
        self.written = 0
        self.failed = 0

    def _insert(self, tokens_usage: dict):
        if self.usage_sink is None:
            return
        fn = getattr(self.usage_sink, "insert_usage", self.usage_sink)
        try:                                                             # :53-56
            fn(tokens_usage)
        except Exception as db_error:
            logger.error(f"Failed to insert token usage data into database: {db_error}", exc_info=True)

    def write_log(self, req_headers, req_body_str: str, llm_response_accum, tokens_usage: dict):
        try:
            if isinstance(llm_response_accum, (bytes, bytearray)):
                llm_response_accum = decode_text(bytes(llm_response_accum))
            log_time = self.clock()                                      # :25-26
            filename = log_time.strftime("%Y-%m-%d_%H-%M-%S") + (".%03d" % (log_time.microsecond // 1000)) + ".txt"
            log_content = render_log(req_headers, req_body_str, llm_response_accum, tokens_usage)
            os.makedirs(self.log_dir, exist_ok=True)                     # :43-44
            log_path = os.path.join(self.log_dir, filename)
            with open(log_path, "w", encoding="utf-8") as f:             # :47-50 (a lone surrogate raises here: no row either)
                log_content = log_content.replace("\\n\\n", "\r\n\r\n").replace("\\n", "\r\n")
                f.write(log_content)
            self.written += 1
            self._insert(tokens_usage)                                   # :52-56
            log_files = sorted(glob.glob(os.path.join(self.log_dir, "*.txt")), key=os.path.getmtime)     # :58-65
            max_logs = self.log_file_limit or 50
            while len(log_files) > max_logs:
                try:
                    os.remove(log_files.pop(0))
                except Exception:
                    pass
        except Exception as e:                                           # :66-67
            self.failed += 1
            logger.error(f"Failed to write chat log: {e}", exc_info=True)
