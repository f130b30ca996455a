#!/bin/bash
for v in ${VARIANTS:-o0 onoise odyn ofilt}; do
  echo -n "$v: "; BSKENV_LIB=$PWD/variants/libbskenv_$v.so OPNAV_N=${OPNAV_N:-75776,4096} timeout 300 python scripts/opnav_quick.py 2>&1 | cut -c1-75 | tr '\n' ' '; echo
done
