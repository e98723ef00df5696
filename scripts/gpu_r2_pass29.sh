#!/bin/bash
# pass 29: probe -- narrow rows (N <= 128 fp16, N <= 64 fp32) through the warp-per-row kernel with four rows per warp
for flags in "" "-DVX_CSR_NARROW_RPW_PROBE"; do
  echo "== flags '$flags'"
  for cfg in "64 fp16" "128 fp16" "32 fp32" "64 fp32"; do
    VOLTRIX_EXTRA_NVCC_FLAGS="$flags" timeout -s KILL 300 python scripts/csr_stream_probe.py YeastH $cfg 2>&1 | tail -1 | cut -c1-150
  done
  VOLTRIX_EXTRA_NVCC_FLAGS="$flags" timeout -s KILL 300 python scripts/csr_stream_probe.py DD 64 fp32 2>&1 | tail -1 | cut -c1-150
  VOLTRIX_EXTRA_NVCC_FLAGS="$flags" timeout -s KILL 300 python scripts/csr_stream_probe.py com-amazon 128 fp16 2>&1 | tail -1 | cut -c1-150
done
