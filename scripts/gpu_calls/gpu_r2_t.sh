#!/usr/bin/env bash
# speculative root launches: new tests first (own timeout), then the suite and the benches (e2e with / without)
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k speculative > $OUT/pytest_t_spec.log 2>&1; tail -15 $OUT/pytest_t_spec.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_t.log 2>&1; tail -4 $OUT/pytest_t.log
timeout 300 python bench.py --steps 30 --warmup 5 --out $OUT/bench_t.jsonl > $OUT/bench_t_glm.log 2>&1; tail -1 $OUT/bench_t_glm.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
timeout 300 python bench.py --shards 1 --steps 30 --warmup 5 --out $OUT/bench_t.jsonl > $OUT/bench_t_glm1.log 2>&1; tail -1 $OUT/bench_t_glm1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
timeout 300 python bench.py --config fp8 --shards 1 --steps 30 --warmup 5 --out $OUT/bench_t.jsonl > $OUT/bench_t_fp8.log 2>&1; tail -1 $OUT/bench_t_fp8.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
timeout 300 python bench.py --config linreg --steps 200 --warmup 20 --out $OUT/bench_t.jsonl > $OUT/bench_t_linreg.log 2>&1; tail -1 $OUT/bench_t_linreg.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
