"""Non-streaming upstream responses.

Mirror of the tail of `make_llm_request` for `is_streaming=False`
(llm_gateway_core/services/request_handler.py:155-176) plus what FastAPI does with the dict it returns
(chat.py:146-148 -> Starlette JSONResponse.render): for a batch of upstream responses the engine parses each
body once on the GPU, probes the top level for `error` / `detail`, and re-renders the document compactly
(`ensure_ascii=False`, `(",", ":")`, floats by repr) -- the bytes the reference sends to the client.

    results = normalise_responses(engine, plans, contents, http_status, target_url)
    -> list of (body_bytes, None)           success: serve body_bytes as application/json
               (None, error_detail)          failed attempt, same convention as request_handler.py
"""
from __future__ import annotations

import json

import numpy as np

from . import rewrite as rw


class ExoticResponse(ValueError):
    """The engine does not model this document (duplicate keys, float with more than 15
    significant digits, NaN...).  The caller keeps the reference's own code path for it."""


def _error_detail(content: bytes, target_url: str):
    # request_handler.py:168 -- only reached for the rare failing response, so it runs on the host
    try:
        doc = json.loads(content)
        if "error" in doc or "detail" in doc:                # (raises for roots that do not support `in`)
            return doc.get("error", {}).get("message") or doc.get("detail")
        return None
    except Exception as e:                                   # request_handler.py:183-187
        return f"Unexpected error during request to {target_url}: {str(e)}"


def normalise_responses(engine, plans: rw.RulePlans, contents, http_status, target_url: str = "", strict: bool = True):
    """`plans` must be the table loaded into the engine (`engine.load_rules(plans)`)."""
    n = len(contents)
    todo = [i for i in range(n) if int(http_status[i]) < 400]
    out = [None] * n
    for i in range(n):
        if int(http_status[i]) >= 400:                       # request_handler.py:159-162
            out[i] = (None, contents[i].decode("utf-8", "replace"))
    if todo:
        got = engine.rewrite_bodies([contents[i] for i in todo], np.full(len(todo), plans.response_plan(), dtype=np.uint32), with_matched=True)
        for i, (st, body, matched, root_kind) in zip(todo, got):
            if st == rw.BODY_PARSE_ERROR:                    # request_handler.py:172-176 / :183-187 (detail text depends on json5)
                out[i] = (None, f"Invalid JSON response from {target_url}")
            elif st == rw.BODY_OK and (matched or root_kind not in (rw.KIND_OBJ, rw.KIND_ARR, rw.KIND_STR)):
                # request_handler.py:167-170: the probe hit (dict key, list element, substring of a string); on a number /
                # null / bool root `"error" in x` itself raises -> :183-187.  Rare: the text is recomputed on the host.
                out[i] = (None, _error_detail(contents[i], target_url))
            elif st == rw.BODY_OK and body in (b"{}", b"[]", b'""'):       # chat.py:146: falsy response_data -> failed attempt
                out[i] = (None, None)
            elif st == rw.BODY_OK:
                out[i] = (body, None)
            elif strict:
                raise ExoticResponse(f"response {i}: {rw.STATUS_NAMES[st]}")
            else:
                out[i] = ("exotic", rw.STATUS_NAMES[st])
    return out
