#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout -s KILL 500 python scripts/routing_probe.py --datasets com-amazon web-BerkStan amazon0601 Yeast protein DD ppi --feature_dims 32 64 512 --out $O/r2ag_routing.csv 2>&1 | grep -v "Warn\|warn" | tail -44 | cut -c1-200
