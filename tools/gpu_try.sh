#!/bin/bash
# parity + timing of the tensor-memory gather variants
python -m pytest tests/test_window_gather_gpu.py -x -q 2>&1 | tail -3
SPB_TMEM_PIPE=1 python -m pytest tests/test_window_gather_gpu.py -x -q 2>&1 | tail -3
for p in 0 1; do
  echo "== SPB_TMEM_PIPE=$p"
  SPB_TMEM_PIPE=$p python tools/sweep_gather.py --config c4 --variants tma,tmem 2>&1 | tail -2
  SPB_TMEM_PIPE=$p python tools/sweep_gather.py --config c2 --variants tmem 2>&1 | tail -1
done
