import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# Fuzz campaigns: LGW_FUZZ_SALT=<anything> re-seeds every `random.Random(seed)` of the test suite (seed -> "seed/salt"), so that
# the differential tests (device machines vs oracle / CPython / the sequential machine) see fresh inputs on every run:
#     for s in 1 2 3; do LGW_FUZZ_SALT=$s python -m pytest tests -q -m "not gpu" -p no:cacheprovider; done
# Unset (the default, and what the driver runs) the suite is deterministic.
import os as _os

_SALT = _os.environ.get("LGW_FUZZ_SALT")
if _SALT:
    import random as _random

    class _SaltedRandom(_random.Random):
        def seed(self, a=None, version=2):
            super().seed(a if a is None else f"{a!r}/{_SALT}", version)

    _random.Random = _SaltedRandom
