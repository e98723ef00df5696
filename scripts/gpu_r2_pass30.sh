#!/bin/bash
# pass 30: reschedule / tune_routing -- parity test, then the routing probe on the small end of the C3 suite
O=gpurun_out; mkdir -p $O
timeout -s KILL 600 python -m pytest tests/test_parity_shapes_gpu.py tests/test_spmm_gpu.py -x -q -m gpu > $O/r2ae_t.log 2>&1; echo "rc=$?"; tail -2 $O/r2ae_t.log
timeout -s KILL 420 python scripts/routing_probe.py --datasets ppi protein DD amazon0505 FraudYelp-RSR --feature_dims 128 256 --out $O/r2ae_routing.csv 2>&1 | grep -v "Warn\|warn" | tail -24 | cut -c1-260
