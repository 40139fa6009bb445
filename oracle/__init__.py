"""CPU oracle (test infrastructure). See sse_oracle.py / rollup_oracle.py headers.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
