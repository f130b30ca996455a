#!/usr/bin/env python
"""Replay a recorded Basilisk trajectory through this repository and report the deviations.

    python scripts/compare_basilisk_trace.py trace.npz --backend oracle|gpu|both [--hill-cel-pun 1] [--json report.json]

The trace format is docs/TRACE_SCHEMA.md; scripts/record_basilisk_trace.py writes it on a machine that has Basilisk 1.x.
`gpu` replays through the CUDA path (C ABI), `oracle` through the CPU restatement.  Exit code 0 = every field within the
tolerances of tests/parity.py.  (The implementation lives in tests/trace_tool.py: it is test infrastructure.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.trace_tool import main  # noqa: E402

if __name__ == "__main__":
    sys.exit(main())
