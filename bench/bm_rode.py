"""RoDe and Sputnik baselines (reference: bench/bm_rode.py, bench/bm_sputnik.py -- both run RoDe's eval driver).

`eval_spmm_f32_n<dim> data.mtx` (third-party/RoDe/eval, built by bench/competitors/build.py) times Sputnik, cuSPARSE and
RoDe on the MatrixMarket file and prints one comma-separated line: `<file>, <sputnik ms>, <gflops>, <cusparse ms>, <gflops>,
<rode ms>, <gflops>`, each time summed over its 10 timed runs.  Prints the two lines bench_all.py parses.
    python bm_rode.py --feat_dim 256 [--rode_home <dir with build/eval>] [--data data.mtx]"""
import argparse
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
RUNS_PER_FIGURE = 10


def run_eval(rode_home: str, feat_dim: int, data: str):
    exe = os.path.join(rode_home, "build", "eval", f"eval_spmm_f32_n{feat_dim}")
    if not os.path.exists(exe):
        raise FileNotFoundError(f"{exe} not found -- run `python bench/competitors/build.py` first")
    out = subprocess.run([exe, data], capture_output=True, text=True, stdin=subprocess.DEVNULL).stdout
    fields = [f.strip() for f in out.strip().splitlines()[-1].split(",")]
    sputnik_ms, cusparse_ms, rode_ms = (float(fields[i]) / RUNS_PER_FIGURE for i in (1, 3, 5))
    return {"Sputnik": sputnik_ms, "cuSPARSE (RoDe driver)": cusparse_ms, "RoDe": rode_ms}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rode_home", default=os.environ.get("RODE_HOME", os.path.join(HERE, "_competitors", "rode")))
    ap.add_argument("--feat_dim", type=int, required=True, choices=[32, 128, 256, 512, 1024])
    ap.add_argument("--data", default="data.mtx")
    a = ap.parse_args()
    for name, ms in run_eval(a.rode_home, a.feat_dim, a.data).items():
        print(f"[{name}] Elapsed time: {ms:.4f} ms")
