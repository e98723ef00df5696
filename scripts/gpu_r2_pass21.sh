#!/bin/bash
# pass 21: compute-sanitizer over every SpMM path on a small graph; ncu on the CUDA-core CSR path (YeastH-shaped)
mkdir -p gpurun_out
echo "== plain run of the sanitizer probe (also warms the JIT cache)"
timeout -s KILL 600 python scripts/sanitize_probe.py > gpurun_out/r2v_sanitize_plain.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2v_sanitize_plain.log
for tool in memcheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout -s KILL 900 compute-sanitizer --tool $tool --error-exitcode 77 --print-limit 20 python scripts/sanitize_probe.py \
     > gpurun_out/r2v_sanitize_$tool.log 2>&1; echo "rc=$?"
  grep -c "^ok " gpurun_out/r2v_sanitize_$tool.log; grep "BAD\|SANITIZE_PROBE\|ERROR SUMMARY\|Invalid\|hazard" gpurun_out/r2v_sanitize_$tool.log | head -12
done
echo "== CSR path on YeastH-shaped: events"
for cfg in "128 fp16" "512 fp16" "128 fp32"; do
  timeout -s KILL 300 python scripts/csr_stream_probe.py YeastH $cfg 2>&1 | tail -1
done
echo "== ncu --set full on the CSR kernel (N=128 fp16)"
timeout -s KILL 600 ncu --set full --clock-control none -k regex:csr -s 2 -c 1 -o gpurun_out/r2v_csr_yeasth -f \
   python scripts/csr_stream_probe.py YeastH 128 fp16 4 > gpurun_out/r2v_ncu_csr.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/r2v_csr_yeasth.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r2v_csr_yeasth_raw.csv; wc -c gpurun_out/r2v_csr_yeasth_raw.csv
