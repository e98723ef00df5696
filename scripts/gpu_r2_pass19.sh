#!/bin/bash
# pass 19: full GPU suite (model-4 narrow-width path), then a DTC-SpMM crash hunt with unbuffered stdio + faulthandler
mkdir -p gpurun_out
echo "== pytest -m gpu"
timeout -s KILL 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2t_t_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2t_t_gpu.log
echo "== DTC debug"
mkdir -p /tmp/dtcdbg && cd /tmp/dtcdbg
python $GRAFT_REPO_ROOT/bench/graph_gen.py --data_name ddi --num_feats 128 --mtx_max_nnz 0 > gen.log 2>&1; echo "gen rc=$?"
cat > dbg.py <<'PY'
import faulthandler, os, sys
faulthandler.enable()
import numpy as np, torch
sys.path.insert(0, os.path.join(os.environ["GRAFT_REPO_ROOT"], "bench", "_competitors", "dtc"))
import DTCSpMM
def say(*a): print(*a, flush=True)
col_h = torch.from_numpy(np.loadtxt("indices.csv", delimiter=",", dtype=np.int32))
ptr_h = torch.from_numpy(np.loadtxt("indptr.csv", delimiter=",", dtype=np.int32))
rows, nnz = ptr_h.numel() - 1, col_h.numel()
windows = (rows + 15) // 16
mode = sys.argv[1]
if mode == "gpu":
    col, ptr = col_h.cuda(), ptr_h.cuda()
    bp = torch.zeros(windows, dtype=torch.int32, device="cuda"); e2c = torch.zeros(nnz, dtype=torch.int32, device="cuda"); e2r = torch.zeros(nnz, dtype=torch.int32, device="cuda")
    say("calling preprocess_gpu")
    r = DTCSpMM.preprocess_gpu(col, ptr, rows, 16, 8, bp, e2c, e2r)
else:
    bp = torch.zeros(windows, dtype=torch.int32); e2c = torch.zeros(nnz, dtype=torch.int32); e2r = torch.zeros(nnz, dtype=torch.int32)
    say("calling preprocess (cpu)")
    r = DTCSpMM.preprocess(col_h, ptr_h, rows, 16, 8, bp, e2c, e2r)
    r = tuple(t.cuda() if torch.is_tensor(t) else t for t in r)
torch.cuda.synchronize(); say("preprocess returned", [tuple(t.shape) if torch.is_tensor(t) else t for t in r])
win_off, blk_row, tile_id, blk_off, a_to_x, _ = r
dense = torch.from_numpy(np.fromfile("feat.csv", dtype=np.float32)).cuda().view(rows, -1)
for plan in ("float4_split", "float_nonsplit"):
    say("run_DTCSpMM", plan)
    out = DTCSpMM.run_DTCSpMM(dense, win_off, tile_id, blk_off, a_to_x, rows, nnz, plan)[0]
    torch.cuda.synchronize(); say("done", plan)
    expected = np.fromfile("output_base.csv", dtype=np.float32).reshape(rows, -1)
    say("allclose", np.allclose(out.cpu().numpy(), expected, atol=1e-1))
say(open("DTCSpMM_exe_time_and_throughput.csv").read())
PY
for m in gpu cpu; do
  echo "-- mode $m"
  CUDA_LAUNCH_BLOCKING=1 timeout -s KILL 240 stdbuf -o0 -e0 python -X faulthandler dbg.py $m 2>&1 | tail -40; echo "rc=${PIPESTATUS[0]}"
done
