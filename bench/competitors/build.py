"""Competitor harness for B200 (reference: bench/bench_all.py:11-33, bench/scripts/build.sh,
patches/rode_fix.patch, third-party/build_dtc.sh).

Builds, for sm_100a, the baselines BASELINE.json config 3 names -- from the reference's OWN sources, compiled where they
lie (nothing is copied into this repository; outputs go to the git-ignored ``bench/_competitors/``, which travels to the
GPU box like every other built artefact):

  gespmm, tcgnn        bench/scripts/gespmm.cu, tcgnn.cu  (GE-SpMM and TC-GNN kernels, stand-alone programs that read the
                        graph_gen.py files from the CWD; reference build line: bench/scripts/build.sh, -arch=sm_90a)
  rode/build/eval/eval_spmm_f32_n{32,128,256,512,1024}
                        third-party/RoDe: RoDe, Sputnik and cuSPARSE SpMM behind one driver per feature width.  The
                        reference builds it with cmake + glog + abseil after applying patches/rode_fix.patch; here the
                        patch is applied to a scratch copy and the five translation units are handed straight to nvcc,
                        with two small stand-in headers (bench/competitors/shims) for the CHECK macros and the uniform
                        random draw RoDe's matrix utilities take from glog / abseil.

  dtc/DTCSpMM.so       third-party/DTC-SpMM/DTC-SpMM: the DTC-SpMM torch extension (reference build: third-party/build_dtc.sh,
                        `setup.py install` against cmake builds of glog and Sputnik).  Here torch.utils.cpp_extension
                        compiles the two sources in place; the Sputnik headers it includes for one baseline entry point the
                        harness never calls are replaced by a stub (bench/competitors/shims/sputnik).

    python bench/competitors/build.py [--ref /root/reference] [--force] [--dtc]
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.abspath(os.path.join(HERE, "..", "_competitors"))
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
RODE_WIDTHS = (32, 128, 256, 512, 1024)


def nvcc() -> str:
    return os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "nvcc")


def _run(cmd, cwd=None):
    p = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd[:6])} ... failed:\n{p.stderr[-3000:]}")


def build_standalone(ref: str, name: str, force: bool) -> str:
    src = os.path.join(ref, "bench", "scripts", name + ".cu")
    dst = os.path.join(OUT, name)
    if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
        _run([nvcc(), *ARCH, "-std=c++17", "-O3", "-w", "--expt-relaxed-constexpr", "--expt-extended-lambda", src, "-o", dst,
              "-lcuda", "-lcublas", "-lcublasLt"])
    return dst


def build_rode(ref: str, force: bool):
    evald = os.path.join(OUT, "rode", "build", "eval")
    os.makedirs(evald, exist_ok=True)
    targets = [os.path.join(evald, f"eval_spmm_f32_n{w}") for w in RODE_WIDTHS]
    if not force and all(os.path.exists(t) for t in targets):
        return targets
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "RoDe")
        os.makedirs(src)
        for d in ("utils", "Sputnik_SpMM", "RoDe_SpMM", "cuSparse_SpMM", "eval"):
            shutil.copytree(os.path.join(ref, "third-party", "RoDe", d), os.path.join(src, d))
        shutil.copy(os.path.join(ref, "third-party", "RoDe", "CMakeLists.txt"), src)
        _run(["patch", "-p1", "-i", os.path.join(ref, "patches", "rode_fix.patch")], cwd=src)
        inc = [f"-I{os.path.join(HERE, 'shims')}"] + [f"-I{os.path.join(src, d)}" for d in
                                                     ("utils", "Sputnik_SpMM", "RoDe_SpMM", "cuSparse_SpMM")]
        common = [*ARCH, "-std=c++17", "-O3", "-w", "--expt-relaxed-constexpr", "-include", "iostream", "-include", "string", *inc]
        libs = {"utils": "utils/matrix_utils.cu", "sputnik": "Sputnik_SpMM/Sputnik_spmm.cu", "rode": "RoDe_SpMM/RoDeSpmm.cu",
                "cusparse": "cuSparse_SpMM/cuSPARSE_spmm.cu"}
        objs = {k: os.path.join(tmp, k + ".o") for k in libs}
        jobs = [[nvcc(), *common, "-c", os.path.join(src, v), "-o", objs[k]] for k, v in libs.items()]
        jobs += [[nvcc(), *common, "-c", os.path.join(src, "eval", f"eval_spmm_f32_n{w}.cu"), "-o",
                  os.path.join(tmp, f"eval{w}.o")] for w in RODE_WIDTHS]
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            list(pool.map(_run, jobs))
        for w, t in zip(RODE_WIDTHS, targets):
            _run([nvcc(), *ARCH, os.path.join(tmp, f"eval{w}.o"), *objs.values(), "-o", t, "-lcusparse"])
    return targets


def build_dtc(ref: str, force: bool) -> str:
    dst = os.path.join(OUT, "dtc", "DTCSpMM.so")
    src = os.path.join(ref, "third-party", "DTC-SpMM", "DTC-SpMM")
    if not force and os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(os.path.join(src, "DTCSpMM_kernel.cu")):
        return dst
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0a"
    # link with the distribution's g++: a toolchain wrapper whose libstdc++.so is missing falls back to libstdc++.a, and a
    # second copy of the locale facets inside the extension crashes the first `std::cout << number` in a torch process
    for var, tool in (("CC", "/usr/bin/gcc"), ("CXX", "/usr/bin/g++")):
        if os.path.exists(tool):
            os.environ[var] = tool
    from torch.utils.cpp_extension import load
    with tempfile.TemporaryDirectory() as tmp:
        load(name="DTCSpMM", sources=[os.path.join(src, "DTCSpMM.cpp"), os.path.join(src, "DTCSpMM_kernel.cu")],
             extra_include_paths=[os.path.join(HERE, "shims"), src], extra_cuda_cflags=["-O3", "-w", "--expt-relaxed-constexpr"],
             extra_cflags=["-O2", "-w"], build_directory=tmp, verbose=False, is_python_module=False)
        shutil.copy(os.path.join(tmp, "DTCSpMM.so"), dst)
    return dst


def build(ref: str = "/root/reference", force: bool = False, dtc: bool = False):
    """``dtc``: also (re)build the DTC-SpMM torch extension -- about four minutes of nvcc, so only on request."""
    if not os.path.isdir(os.path.join(ref, "bench", "scripts")):
        raise FileNotFoundError(f"reference tree not found at {ref}")
    os.makedirs(OUT, exist_ok=True)
    built = [build_standalone(ref, n, force) for n in ("gespmm", "tcgnn")]
    built += build_rode(ref, force)
    if dtc:
        built.append(build_dtc(ref, force))
    return built


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default=os.environ.get("VOLTRIX_REF", "/root/reference"))
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--dtc", action="store_true", help="also build the DTC-SpMM torch extension (~4 min)")
    a = ap.parse_args()
    for path in build(a.ref, a.force, a.dtc):
        print(os.path.relpath(path, os.path.join(HERE, "..", "..")))
