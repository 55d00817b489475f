#!/bin/bash
# parity + timing of the stage-1 gather kernels
python -m pytest tests/test_window_gather_gpu.py -x -q 2>&1 | tail -3
python tools/sweep_gather.py --config c4 --variants tma,tmem 2>&1 | tail -2
python tools/sweep_gather.py --config c2 --variants tma,tmem 2>&1 | tail -2
