#!/bin/bash
# 2 GPUs: full GPU test-suite on rank 0's device first, then sharded parity incl. the sharded bake, c5s
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/check_sharded.py > gpurun_out/check_sharded_2.log 2>&1; echo "check_sharded rc=$?"; grep "bake\|PARITY\|Error\|error" gpurun_out/check_sharded_2.log | tail
timeout 600 python bench.py --config c5s --verbose > gpurun_out/bench_c5s_1.json 2> gpurun_out/bench_c5s_1.err; echo "c5s x1 rc=$?"; tail -3 gpurun_out/bench_c5s_1.err
timeout 600 $TR bench.py --config c5s --gpus 2 --verbose > gpurun_out/bench_c5s_2.json 2> gpurun_out/bench_c5s_2.err; echo "c5s x2 rc=$?"; grep "bench\]\|Error" gpurun_out/bench_c5s_2.err | tail -3
python - <<'PY'
import json
for f in ['bench_c5s_1','bench_c5s_2']:
    try:
        d=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f, d['ms_per_step'], d['result_checksum'], d['bake'], d['exchange_peak_memory_gb_rank0'])
    except Exception as e: print(f, 'ERR', e)
PY
python bench.py --verbose > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench c4 rc=$?"; tail -2 gpurun_out/bench_c4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c4.json')); print(d['ms_per_step'], d['pipeline'])
PY
