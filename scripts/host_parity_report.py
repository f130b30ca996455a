#!/usr/bin/env python
"""Worst relative differences between the host-compiled device core and the oracle over a seeded batch
(development aid: shows how much of the 1e-9 parity budget a kernel change uses)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc
from tests import hostcore_binding as hostcore, parity
from tests.test_hostcore_parity import run_pair
orc.lib(); hostcore.build(); hostcore.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
rows = parity.sample_rows(orc, n, seed=77)
rows[:3, 15:18] = np.array([[2500., -2300., 2700.], [-2900., 2000., 1500.], [2999., 2999., -2999.]])
acts = np.random.RandomState(9).randint(0, 3, size=(steps, n))
worst = run_pair(orc, hostcore, rows, acts)
print({k: f"{v:.2e}" for k, v in worst.items()})
