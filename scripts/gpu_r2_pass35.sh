#!/bin/bash
# pass 35: model 4 with one memset per call -- GPU suite, then fp32 timing on the small C3 shapes
O=gpurun_out; mkdir -p $O
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > $O/r2aj_t_gpu.log 2>&1; echo "rc=$?"; tail -2 $O/r2aj_t_gpu.log
timeout -s KILL 600 python scripts/suite.py --datasets protein ppi DD ddi com-amazon reddit --feature_dims 64 128 256 --out $O/r2aj_suite_small.csv > $O/r2aj_suite_small.log 2>&1; echo "rc=$?"
cut -d, -f1,5,6,8,11 $O/r2aj_suite_small.csv
