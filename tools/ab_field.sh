#!/bin/bash
echo "== default"; python tools/perf_field_bwd.py 2>&1 | tail -2 | tr -d '\n'; echo
for so in iris_b200/_lib/ab/*.so; do
  echo "== $so"; IRIS_B200_LIB=$PWD/$so python tools/perf_field_bwd.py 2>&1 | tail -2 | tr -d '\n'; echo
done
