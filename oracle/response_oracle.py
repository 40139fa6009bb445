"""CPU oracle for the non-streaming response path (row a12) -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates llm_gateway_core/services/request_handler.py:155-176 (status gate, response.json(), error/detail
probe), the success test of llm_gateway_core/api/v1/chat.py:146 (`if response_data and error_detail is None`)
and the bytes FastAPI makes of the returned dict (Starlette JSONResponse.render:
json.dumps(content, ensure_ascii=False, allow_nan=False, indent=None, separators=(",", ":")).encode("utf-8")).

Pinned by tests/golden/response_cases.json (tests/golden/make_response_golden.py drives the unmodified
make_llm_request(..., is_streaming=False) through httpx.MockTransport and renders with the installed
Starlette).  The error_detail TEXT of an invalid-JSON response depends on the absent json5 package
(request_handler.py:172 names json5.JSONDecodeError) -> only the verdict is pinned for that case.
"""
from __future__ import annotations

import json


def render(doc) -> bytes:
    return json.dumps(doc, ensure_ascii=False, allow_nan=False, indent=None, separators=(",", ":")).encode("utf-8")


def normalise(status_code: int, content: bytes, url: str = "http://upstream.test/v1/chat/completions"):
    """-> ("ok", body_bytes) | ("fail", error_detail | None) | ("raise", None) when the renderer raises."""
    if status_code >= 400:                                    # :159-162
        return "fail", content.decode("utf-8", "replace")
    try:
        doc = json.loads(content)                             # :166 response.json()
    except Exception:
        return "fail", "<invalid json: detail text unpinned>"  # :172-176 / :183-187
    try:
        if "error" in doc or "detail" in doc:                 # :167 (dict: key test)
            return "fail", doc.get("error", {}).get("message") or doc.get("detail")      # :168
    except Exception as e:                                     # :183-187 "Unexpected error during request ..."
        return "fail", f"Unexpected error during request to {url}: {str(e)}"
    if not doc:                                                # chat.py:146 falsy response_data
        return "fail", None
    try:
        return "ok", render(doc)
    except ValueError:                                         # NaN / Infinity with allow_nan=False
        return "raise", None
