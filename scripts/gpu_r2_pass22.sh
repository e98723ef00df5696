#!/bin/bash
# pass 22: row-batched CSR kernel -- focused parity test, the sparse end of the C3 suite, then the whole GPU suite
mkdir -p gpurun_out
echo "== low-degree parity"
timeout -s KILL 600 python -m pytest tests/test_spmm_gpu.py -x -q -m gpu -k "low_degree or weighted" > gpurun_out/r2w_t_lowdeg.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2w_t_lowdeg.log
echo "== CSR path on YeastH-shaped: events"
for cfg in "128 fp16" "512 fp16" "128 fp32"; do
  timeout -s KILL 300 python scripts/csr_stream_probe.py YeastH $cfg 2>&1 | tail -1
done
echo "== C3 suite, sparse end"
timeout -s KILL 900 python scripts/suite.py --datasets protein DD com-amazon Yeast YeastH ppi amazon0505 --out gpurun_out/r2w_suite_sparse.csv > gpurun_out/r2w_suite_sparse.log 2>&1; echo "rc=$?"
cat gpurun_out/r2w_suite_sparse.csv | cut -d, -f1,5,6,7,8,9,10,11,14
echo "== pytest -m gpu"
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2w_t_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2w_t_gpu.log
