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

import numpy as np

from . import rewrite as rw


class ExoticResponse(ValueError):
    """The engine does not model this document (duplicate keys, float with more than 15
    significant digits, NaN...).  The caller keeps the reference's own code path for it."""


_ITER_ERRORS = {1: "int", 7: "int", 2: "float", 10: "float", 3: "NoneType", 4: "bool", 5: "bool"}     # lgw_kind of the root -> Python type name


def _error_detail(engine, content: bytes, target_url: str, root_kind: int):
    """request_handler.py:167-169 for a response whose probe hit: `response_json.get("error", {}).get("message") or
    response_json.get("detail")`, and the exception texts when the document does not support that (:183-187).  The document is
    walked ON THE DEVICE (lgw_documents_error_detail, csrc/error_detail.cuh): which value wins by Python's truthiness rules, its
    unescaped text when it is a string.  -> (detail, None) or (None, reason) when the winning value's str() is not modelled."""
    if root_kind == rw.KIND_ARR:                                  # `"error" in [..]` held: list.get does not exist
        return f"Unexpected error during request to {target_url}: 'list' object has no attribute 'get'", None
    if root_kind == rw.KIND_STR:                                  # substring test held
        return f"Unexpected error during request to {target_url}: 'str' object has no attribute 'get'", None
    if root_kind != rw.KIND_OBJ:                                  # `"error" in 5` itself raises
        name = _ITER_ERRORS.get(root_kind)
        if name is None:
            return None, "root value"
        return f"Unexpected error during request to {target_url}: argument of type '{name}' is not iterable", None
    e, text = engine.documents_error_detail([content])[0]
    if e.result == 0:
        return None, None
    if e.result == 1:
        return text.decode("utf-8"), None
    if e.result in (2, 3):
        return e.result == 2, None                                # the bool itself, as the reference returns it
    if e.result == 4:
        name = {1: "str", 2: "NoneType", 3: "bool", 4: "bool", 9: "list", 10: "list"}.get(e.error_kind)
        if name is None:                                          # a number: int or float by its spelling
            name = "float" if any(c in text for c in b".eE") else "int"
        return f"Unexpected error during request to {target_url}: '{name}' object has no attribute 'get'", None
    if e.result == 6:                                             # the number itself, as the reference returns it (a located literal, not a document)
        try:
            return (float(text) if any(c in text for c in b".eE") else int(text)), None
        except ValueError:
            return None, "number literal"
    return None, "error detail is not a string"


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
                # null / bool root `"error" in x` itself raises -> :183-187.  The winning value is read out on the device.
                detail, exotic = _error_detail(engine, contents[i], target_url, root_kind)
                if exotic is None:
                    out[i] = (None, detail)
                elif strict:
                    raise ExoticResponse(f"response {i}: {exotic}")
                else:
                    out[i] = ("exotic", exotic)
            elif st == rw.BODY_OK and body in (b"{}", b"[]", b'""'):       # chat.py:146: falsy response_data -> failed attempt
                out[i] = (None, None)
            elif st == rw.BODY_OK:
                out[i] = (body, None)
            elif strict:
                raise ExoticResponse(f"response {i}: {rw.STATUS_NAMES[st]}")
            else:
                out[i] = ("exotic", rw.STATUS_NAMES[st])
    return out
