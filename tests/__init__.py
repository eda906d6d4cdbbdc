"""Test suite: `-m "not gpu"` runs on a CPU-only box, `-m gpu` needs a B200 (see tests/conftest.py)."""
