"""Sweep driver (reference: bench/bench_all.py:63-149): for every graph x feature dim, generate inputs with
graph_gen.py, run each method as a subprocess, parse its `[X] ... time: ` line, append results.csv
(`Method,Dataset,FeatDim,Reorder,Time (ms)`).  Differences: graphs are the synthetic suite of
voltrix.graphs.named_suite(); the N sweep is 32/64/128/256/512 (BASELINE.json configs[2]); methods are the ones
that exist on this box -- cuSPARSE and Voltrix (fp32 / fp16), plus the competitors bench/competitors/build.py compiled
for sm_100a from the reference's own sources into bench/_competitors/ (GE-SpMM, TC-GNN, RoDe, Sputnik, DTC-SpMM; fp32, like
the reference's columns)."""
import argparse
import csv
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
METHODS = {
    "cuSPARSE": ([sys.executable, os.path.join(HERE, "bm_sparse.py")], "[cuSPARSE] Elapsed time: "),
    "Voltrix": ([sys.executable, os.path.join(HERE, "bm_voltrix.py")], "[Voltrix] time: "),
    "Voltrix-fp16": ([sys.executable, os.path.join(HERE, "bm_voltrix.py"), "--dtype", "fp16"], "[Voltrix] time: "),
}
COMP = os.path.join(HERE, "_competitors")
RODE_DIMS = (32, 128, 256, 512, 1024)       # one eval binary per feature width (patches/rode_fix.patch adds 256-1024)
if os.path.exists(os.path.join(COMP, "gespmm")):
    METHODS["GE-SPMM"] = ([os.path.join(COMP, "gespmm")], "[GE-SPMM] Embedding time: ")
if os.path.exists(os.path.join(COMP, "tcgnn")):
    METHODS["TC-GNN"] = ([os.path.join(COMP, "tcgnn")], "[TC-GNN] Kernel time: ")
if os.path.exists(os.path.join(COMP, "dtc", "DTCSpMM.so")):
    METHODS["DTC-SPMM"] = ([sys.executable, os.path.join(HERE, "bm_dtc.py")], "[DTC-SPMM] Elapsed time: ")
if os.path.isdir(os.path.join(COMP, "rode", "build", "eval")):
    METHODS["RoDe"] = ([sys.executable, os.path.join(HERE, "bm_rode.py")], "[RoDe] Elapsed time: ")
    METHODS["Sputnik"] = ([sys.executable, os.path.join(HERE, "bm_sputnik.py")], "[Sputnik] Elapsed time: ")
FEATURE_DIMS = [32, 64, 128, 256, 512]


def parse_time(stdout: str, marker: str):
    for line in stdout.splitlines():
        if marker in line:
            return float(line.split(marker)[1].split()[0])
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--datasets", nargs="*", default=None)
    ap.add_argument("--feature_dims", nargs="*", type=int, default=FEATURE_DIMS)
    ap.add_argument("--results", default="results.csv")
    ap.add_argument("--reorder", action="store_true", help="also run the reordered variants")
    ap.add_argument("--no_competitors", action="store_true", help="only cuSPARSE and Voltrix")
    args = ap.parse_args()
    sys.path.insert(0, os.path.join(HERE, "..", "voltrix-spmm_b200"))
    from voltrix.graphs import named_suite
    datasets = args.datasets or [n for n, _, nnz in named_suite() if nnz <= 10_000_000]
    new = not os.path.exists(args.results)
    with open(args.results, "a", newline="") as fh:
        w = csv.writer(fh)
        if new:
            w.writerow(["Method", "Dataset", "FeatDim", "Reorder", "Time (ms)"])
        for ds in datasets:
            for fd in args.feature_dims:
                for reorder in ([False, True] if args.reorder else [False]):
                    with tempfile.TemporaryDirectory() as tmp:
                        need_mtx = "RoDe" in METHODS and fd in RODE_DIMS and not args.no_competitors
                        gen = [sys.executable, os.path.join(HERE, "graph_gen.py"), "--data_name", ds, "--num_feats",
                               str(fd), "--mtx_max_nnz", "20000000" if need_mtx else "0"] + (["--reorder"] if reorder else [])
                        r = subprocess.run(gen, cwd=tmp, capture_output=True, text=True)
                        if r.returncode != 0:
                            print(f"graph_gen failed for {ds}: {r.stderr[-300:]}")
                            continue
                        for method, (cmd, marker) in METHODS.items():
                            competitor = method in ("GE-SPMM", "TC-GNN", "RoDe", "Sputnik", "DTC-SPMM")
                            if competitor and args.no_competitors:
                                continue
                            if competitor and reorder != (method == "DTC-SPMM" and args.reorder):
                                continue     # the reference runs DTC-SpMM on the reordered graph, the others un-reordered
                            if method in ("RoDe", "Sputnik"):
                                if fd not in RODE_DIMS or not os.path.exists(os.path.join(tmp, "data.mtx")):
                                    continue
                                cmd = cmd + ["--feat_dim", str(fd)]
                            if method in ("GE-SPMM", "TC-GNN") and fd % 128 != 0:   # their kernels tile 128 features per warp
                                continue
                            try:
                                r = subprocess.run(cmd, cwd=tmp, capture_output=True, text=True, timeout=600)
                            except subprocess.TimeoutExpired:
                                print(f"{method:14s} {ds:14s} N={fd:4d}: TIMEOUT")
                                continue
                            t = parse_time(r.stdout, marker)
                            print(f"{method:14s} {ds:14s} N={fd:4d} reorder={reorder}: "
                                  f"{t if t is not None else 'FAILED ' + r.stderr[-200:]}")
                            if t is not None:
                                w.writerow([method, ds, fd, reorder, t])
                                fh.flush()


if __name__ == "__main__":
    main()
