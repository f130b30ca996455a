#!/bin/bash
# Round times of 1/2/3 resident warps per SM sub-partition and the effect of the short-block quota.
for n in 18944 37888 56832; do
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --envs-per-gpu $n 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('envs',$n,'ms %.2f'%d['ms_per_step'])"
done
for q in 0 1 2 3 4; do
  BSKENV_QUOTA_SHORT=$q python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('quota',$q,'ms %.2f'%d['ms_per_step'])"
done
