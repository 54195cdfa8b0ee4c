#!/bin/bash
# run tools/quick_perf.py once per A/B build found in iris_b200/_lib/ab (plus the default library)
echo "== default"; python tools/quick_perf.py 1000000 2>&1 | grep -E "Mrays|Msamples" | tr -d '\n'; echo
for so in iris_b200/_lib/ab/*.so; do
  echo "== $so"; IRIS_B200_LIB=$PWD/$so python tools/quick_perf.py 1000000 2>&1 | grep -E "Mrays|Msamples|Error|error" | tr -d '\n'; echo
done
