#!/usr/bin/env python
"""Build tuning variants of libbskenv.so (different -D flags for bskenv.cu) under variants/ and print the register /
spill figures of the reference-configuration step kernel.  `BSKENV_LIB=<path> python bench.py ...` times one of them.
    python scripts/build_variants.py name="-DFLAG=1 -DOTHER=2" name2="..."   """
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from basilisk_env_b200 import build as b
VARIANTS = dict(a.split("=", 1) for a in sys.argv[1:])
os.makedirs(os.path.join(ROOT, "variants"), exist_ok=True)
b.build()                                                   # default objects (opnav.o is shared by every variant)
objdir = os.path.join(b.HERE, "build", os.path.basename(b.LIB) + ".obj")
KERNEL = os.environ.get("VARIANT_KERNEL", "leo_step_kernelILi3ELi0ELb1ELb0")
VSRC = os.environ.get("VARIANT_SRC", "bskenv.cu")          # the translation unit that takes the flags; the others are shared


def one(item):
    name, flags = item
    out = os.path.join(ROOT, "variants", f"libbskenv_{name}.so")
    objs = []
    log = ""
    for src in b.SOURCES:
        if not os.path.exists(os.path.join(b.CSRC, src)):
            continue
        if src != VSRC:
            objs.append(os.path.join(objdir, src.replace(".cu", ".o"))); continue
        obj = os.path.join(ROOT, "variants", f"{name}_{src[:-3]}.o")
        r = subprocess.run([b.nvcc_path()] + b.NVCC_FLAGS + flags.split() + ["-Xptxas", "-v", "-c", "-o", obj, os.path.join(b.CSRC, src)],
                           cwd=b.CSRC, capture_output=True, text=True)
        if r.returncode:
            return f"{name}: BUILD FAILED\n{r.stderr[-3000:]}"
        log += r.stderr
        objs.append(obj)
    subprocess.check_call([b.nvcc_path(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs)
    lines = log.splitlines()
    res = []
    for i, l in enumerate(lines):
        if "Function properties" in l and any(k in l for k in KERNEL.split(",")):
            res.append(f"{name} [{flags}] {l.split('for ')[-1][:60]} | {lines[i + 1].strip()} | {lines[i + 2].strip()}")
    return "\n".join(res) or f"{name}: kernel {KERNEL} not found in ptxas output"


with ThreadPoolExecutor(max_workers=4) as ex:
    for r in ex.map(one, VARIANTS.items()):
        print(r, flush=True)
