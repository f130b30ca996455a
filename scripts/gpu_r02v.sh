#!/bin/bash
for v in ${VARIANTS:-old2 s0}; do
  echo "== $v"; BSKENV_LIB=$PWD/variants/libbskenv_$v.so timeout 300 python scripts/split_check.py --envs 4096 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['split'], 'bit_equal', d['outputs_bit_equal'], 'thread ms', round(d['ms_thread'],3), 'split ms', round(d['ms_split'],3))"
  echo -n "probe: "; BSKENV_LIB=$PWD/variants/libbskenv_$v.so timeout 300 python scripts/small_probe.py --envs 4096 2>&1 | tail -1 | cut -c1-80
done
