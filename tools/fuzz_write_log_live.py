"""Live differential of the chat log file against the UNMODIFIED reference `write_log` (chat_logging.py:22-67; dev container only).

Random request headers / request body text / transcript text (escaped and real newlines, non-ASCII, lone surrogates, control
characters) / usage dicts (missing keys, None, strings, floats where ints go) through the real `write_log` (in a temp cwd, the DB
insert captured) and through `llmapigateway_b200.transcripts.TranscriptLog.write_log`: the bytes of the file and whether the usage
row was handed to the DB must be the same (a failing write writes neither).

    python tools/fuzz_write_log_live.py --n 3000 --seed 1
"""
from __future__ import annotations

import argparse
import os
import random
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tests" / "golden"):
    sys.path.insert(0, str(p))

TEXTS = ["hello world", "", "line1\\nline2\\n\\nline3", "real\nnewline\r\nand\ttab", "é 中 😀", "bad \ud83d tail", "\x00\x01 ctl", "\\\\n double backslash", "x" * 300,
         'partial{"error":{"message":"boom \\n x"},"code":500}', "\\n", "\\n\\n", "\\n\\n\\n", "trailing\\", "{braces} %s %d {0}"]


def rand_usage(rng):
    u = {"prompt_tokens": rng.choice([0, 10, 3.5, None, "7", 2**40]), "completion_tokens": rng.choice([0, 5, None]), "total_tokens": rng.choice([0, 17, "x"]),
         "reasoning_tokens": rng.choice([0, 2]), "cached_tokens": rng.choice([0, 4, None]), "cost": rng.choice([0, 0.00123, 1, 1e-9, 123456.789, None, "0.5", float("nan"), float("inf"), -0.0, True])}
    if rng.random() < 0.6:
        u["model"] = rng.choice(["m-ok", "é", "", None, 5])
    if rng.random() < 0.5:
        u["provider"] = rng.choice(["P", "", None])
    if rng.random() < 0.08:
        del u[rng.choice(["prompt_tokens", "completion_tokens", "total_tokens", "reasoning_tokens", "cached_tokens", "cost"])]
    return u


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    if not Path("/root/reference").exists():
        print("needs /root/reference (dev container)"); return 2
    import logging
    import ref_driver
    from llmapigateway_b200.transcripts import TranscriptLog
    _, cl = ref_driver.load_reference()
    logging.disable(logging.CRITICAL)
    rng = random.Random(args.seed)
    bad = n_files = 0
    cwd = os.getcwd()
    for it in range(args.n):
        headers = {rng.choice(["host", "content-type", "x-title", "authorization", "h%02d" % rng.randrange(20)]): rng.choice(["gw", "é", "v" * 40, ""]) for _ in range(rng.randrange(0, 9))}
        body = rng.choice(['{"model":"m","messages":[{"role":"user","content":"hi"}]}', "{}", "", '{"messages":[{"content":"l1\\nl2\\n\\nl3"}]}', '{"k":"中"}', "not json \\n"])
        accum = "".join(rng.choice(TEXTS) for _ in range(rng.randrange(0, 4)))
        usage = rand_usage(rng)
        rows_ref, rows_our = [], []
        with tempfile.TemporaryDirectory() as d:
            os.chdir(d)
            real = cl.tokens_usage_db.insert_usage
            cl.tokens_usage_db.insert_usage = lambda u: rows_ref.append(dict(u))
            try:
                cl.write_log(headers, body, accum, usage)
            finally:
                cl.tokens_usage_db.insert_usage = real
                os.chdir(cwd)
            files = sorted(Path(d, "logs").glob("*.txt")) if Path(d, "logs").exists() else []
            ref_file = files[0].read_bytes() if files else None
        with tempfile.TemporaryDirectory() as d:
            TranscriptLog(log_dir=os.path.join(d, "logs"), usage_sink=lambda u: rows_our.append(dict(u))).write_log(headers, body, accum, usage)
            files = sorted(Path(d, "logs").glob("*.txt")) if Path(d, "logs").exists() else []
            our_file = files[0].read_bytes() if files else None
        # (a write that fails half-way leaves a truncated file behind on both sides: compare what is there)
        same_rows = [repr(sorted(r.items(), key=str)) for r in rows_ref] == [repr(sorted(r.items(), key=str)) for r in rows_our]
        if ref_file != our_file or not same_rows:
            bad += 1
            print(f"FAIL {it}: file equal {ref_file == our_file}, rows ref/ours {len(rows_ref)}/{len(rows_our)}; usage {usage!r} accum {accum[:80]!r}"[:700], flush=True)
        n_files += bool(rows_ref)
    print(f"{args.n - bad}/{args.n} write_log calls: same file bytes and same DB hand-off ({n_files} completed writes, {args.n - n_files} failed on both sides)")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
