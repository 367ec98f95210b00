"""TEST INFRASTRUCTURE — the seeded synthetic inputs live in bench_data.py at the repository root (bench.py's own arm
must not import anything from oracle/); this module re-exports them for the golden generator and the tests."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench_data import probability_maps, rectangles_mask, train_batch  # noqa: E402,F401
