"""Generate tests/golden/sse_cases.json by running the UNMODIFIED reference (dev container only).

    python tests/golden/make_golden.py

For every case in tests/sse_cases.py the real make_llm_request (request_handler.py:8) and the
real ChunkProcessorThread.run (chat_logging.py:87) are run through ref_driver.py; inputs and
observed outputs are stored.  The fixture header records the library versions that produced it
(`json5` is a stdlib-json shim: see ref_driver.py and SURVEY.md section 8(c)).
"""
from __future__ import annotations

import base64
import json
import platform
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE)); sys.path.insert(0, str(HERE.parent))

import ref_driver  # noqa: E402
import sse_cases  # noqa: E402


def b64(b: bytes) -> str:
    return base64.b64encode(b).decode("ascii")


def main():
    import httpx, starlette, fastapi
    cases = []
    for name, chunks, status in sse_cases.all_cases():
        chunks = [c for c in chunks if c]        # httpx never delivers empty chunks to aiter_bytes()
        relay = ref_driver.run_relay(chunks, status)
        rows, texts = ref_driver.run_tap(relay["emitted"]) if not relay["failed"] else ([], [])
        cases.append({
            "name": name, "http_status": status, "chunks": [b64(c) for c in chunks],
            "failed": relay["failed"], "error_detail": relay["error_detail"],
            "emitted": [b64(c) for c in relay["emitted"]],
            "end_raises": relay["end_exception"] is not None,
            "rows": json.dumps(rows, sort_keys=True), "transcripts": texts,
        })
    doc = {"generator": "tests/golden/make_golden.py", "reference_commit": "ade4090 (.SUBMODULES.json)",
           "json5": "ABSENT -> stdlib json shim (strict JSON inputs only)",
           "versions": {"python": platform.python_version(), "httpx": httpx.__version__,
                        "starlette": starlette.__version__, "fastapi": fastapi.__version__},
           "cases": cases}
    out = HERE / "sse_cases.json"
    out.write_text(json.dumps(doc, indent=0, ensure_ascii=True))
    print(f"wrote {out} ({len(cases)} cases, {out.stat().st_size} bytes)")


if __name__ == "__main__":
    main()
