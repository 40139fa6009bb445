"""Helpers to read the committed golden fixtures (tests/golden/*.json)."""
from __future__ import annotations

import base64
import json
from pathlib import Path

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_sse_cases():
    doc = json.loads((GOLDEN / "sse_cases.json").read_text())
    out = []
    for c in doc["cases"]:
        c = dict(c)
        c["chunks"] = [base64.b64decode(x) for x in c["chunks"]]
        c["emitted"] = [base64.b64decode(x) for x in c["emitted"]]
        out.append(c)
    return out


def canon_rows(rows) -> str:
    """Canonical text of a row list (NaN-safe comparison)."""
    return json.dumps(rows, sort_keys=True)


UNPINNED_DETAIL_PREFIX = "Unexpected error during request to"
