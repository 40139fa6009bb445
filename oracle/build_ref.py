"""Recipe for oracle/_ref: a byte-for-byte copy of the reference's OWN Python package, made from the sources where they lie
under /root/reference (dev container only).  TEST INFRASTRUCTURE.

The reference (fabiojbg/LLMApiGateway) is pure Python, so "building" it is copying it.  oracle/_ref/ is git-ignored (no
reference source enters the history) but not gpurun-ignored, so it travels to the GPU box, where bench.py's CPU legs run
the reference's real make_llm_request / ChunkProcessorThread (oracle/ref_arm.py) next to the GPU numbers.
__graft_entry__.build() calls this when /root/reference is present.
"""
from __future__ import annotations

import shutil
from pathlib import Path

REF = Path("/root/reference")
DST = Path(__file__).resolve().parent / "_ref"


def build() -> bool:
    """Refresh oracle/_ref from /root/reference.  False when the reference tree is not present (GPU box: use what travelled)."""
    if not (REF / "llm_gateway_core").is_dir():
        return DST.is_dir()
    if DST.exists():
        shutil.rmtree(DST)
    DST.mkdir(parents=True)
    shutil.copytree(REF / "llm_gateway_core", DST / "llm_gateway_core", ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.db"))
    for name in ("main.py", "requirements.txt", "pyproject.toml", "LICENSE"):
        if (REF / name).exists():
            shutil.copy2(REF / name, DST / name)
    (DST / "README.txt").write_text("Unmodified copy of /root/reference (fabiojbg/LLMApiGateway) made by oracle/build_ref.py; not tracked by git.\n")
    return True


if __name__ == "__main__":
    print("oracle/_ref", "ready" if build() else "unavailable")
