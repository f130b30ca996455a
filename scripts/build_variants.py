#!/usr/bin/env python
"""Build tuning variants of libbskenv.so (different -D flags) under build/variants/ and print the
register / spill figures.  `BSKENV_LIB=<path> python bench.py ...` times one of them."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from basilisk_env_b200 import build as b
VARIANTS = dict(a.split("=", 1) for a in sys.argv[1:]) if len(sys.argv) > 1 else {}
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
for name, flags in VARIANTS.items():
    out = os.path.join(ROOT, "variants", f"libbskenv_{name}.so")
    fl = flags.split()
    log = subprocess.run([b.nvcc_path()] + b.NVCC_FLAGS + fl + ["-Xptxas", "-v", "-o", out, os.path.join(b.CSRC, "bskenv.cu")],
                         cwd=b.CSRC, capture_output=True, text=True)
    lines = log.stderr.splitlines()
    for i, l in enumerate(lines):
        if "leo_step_kernelILi3ELb0ELb1" in l and "Function properties" in l:
            print(name, flags, "|", lines[i + 1].strip(), "|", lines[i + 2].strip())
            break
    else:
        print(name, "build output:", log.stderr[-2000:])
