"""
ORACLE -- test infrastructure, not product code.

A CPU restatement of the reference renderer's hot path (stevenrobertson/cuburn),
used only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under cuburn_b200/ imports this package.

Parity status: the reference cannot be executed here (Python 2 + PyCUDA +
CUDA-4-era texture references).  The oracle is pinned
against the only known-answer material the reference holds for this path -- the
MWC recurrence model (code/mwc.py:90-129), the multiplier table
(code/primes.bin, regenerated and compared byte for byte), the pixel-format
assertions (code/tests/test_output.py) and the profile-time tests -- and is
otherwise a line-by-line restatement of the cited sources: interpolation,
palette, iterate, variations, filters: **parity unpinned by reference vectors**.
"""
