#!/bin/bash
# Round 2: the multi-CTA-per-SM variants as the shipped tune space -- every GPU test, bench with extras, full C3 suite, fp32 and weighted probes.
O=gpurun_out; mkdir -p $O
echo "== pytest -m gpu"; timeout -s KILL 1500 python -m pytest tests -q -m gpu --timeout 300 --maxfail 20 > $O/r2o_t_gpu.log 2>&1; echo "rc=$?"; tail -5 $O/r2o_t_gpu.log | grep -v "Warning\|sparse_csr\|^$"
echo "== bench"; timeout -s KILL 1200 python bench.py > $O/r2o_bench_n1.json 2> $O/r2o_bench_n1.err; echo "rc=$?"; grep "timed loop\|ok'" $O/r2o_bench_n1.err | cut -c1-220
echo "== reference arm"; timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2o_bench_n1_ref.json 2> $O/r2o_bench_n1_ref.err; echo "rc=$?"; cut -c1-200 $O/r2o_bench_n1_ref.json
echo "== fp32 variants reddit"; timeout -s KILL 400 python scripts/time_models.py --workload reddit --dtype fp32 --only 4/12/6,3/12/6,3/24/8,1/32/8 2>&1 | grep "^model" | tee $O/r2o_tm_reddit_fp32.log
echo "== weighted probe"; timeout -s KILL 400 python scripts/weighted_probe.py --workload reddit 2>&1 | grep -v Warn | tail -9 | tee $O/r2o_weighted_reddit.log
echo "== C3 suite"; timeout -s KILL 1500 python scripts/suite.py --out $O/r2o_suite_c3.csv > $O/r2o_suite_c3.log 2>&1; echo "rc=$?"; tail -2 $O/r2o_suite_c3.log
echo "== shard cost probe"; timeout -s KILL 600 python scripts/shard_cost_probe.py > $O/r2o_shard_cost.log 2>&1; echo "rc=$?"; grep "shard 1/" $O/r2o_shard_cost.log
