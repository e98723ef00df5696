"""Competitor harness: the stand-in headers compile, the build script produces every binary
bench_all.py looks for (only where the reference tree is mounted), and the wrappers parse RoDe's output line."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("VOLTRIX_REF", "/root/reference")


def test_shims_compile_as_host_code(tmp_path):
    src = tmp_path / "t.cc"
    src.write_text('#include "glog/logging.h"\n#include "absl/random/random.h"\n'
                   'int main() { absl::BitGen g; float x = absl::Uniform<float>(g, -1, 1); int k = absl::Uniform<int>(g, 0, 5);\n'
                   '  CHECK_GE(x, -1.0f) << "range"; CHECK_LT(k, 5) << "range"; CHECK_EQ(1, 1); return 0; }\n')
    exe = tmp_path / "t"
    p = subprocess.run(["g++", "-std=c++17", f"-I{os.path.join(ROOT, 'bench', 'competitors', 'shims')}", str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert subprocess.run([str(exe)]).returncode == 0


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "third-party", "RoDe", "eval")), reason="reference tree not mounted")
def test_build_script_produces_every_binary():
    sys.path.insert(0, os.path.join(ROOT, "bench", "competitors"))
    import build as competitors_build
    built = competitors_build.build(REF)          # incremental: a no-op after __graft_entry__.build()
    names = {os.path.basename(b) for b in built}
    assert {"gespmm", "tcgnn"} <= names and {f"eval_spmm_f32_n{w}" for w in (32, 128, 256, 512, 1024)} <= names
    assert all(os.access(b, os.X_OK) for b in built)


def test_rode_output_line_parsing(tmp_path, monkeypatch):
    sys.path.insert(0, os.path.join(ROOT, "bench"))
    import bm_rode
    home = tmp_path / "rode"
    evald = home / "build" / "eval"
    evald.mkdir(parents=True)
    exe = evald / "eval_spmm_f32_n256"
    exe.write_text("#!/bin/sh\necho 'loading...'\necho \"$1, 12.5, 100.0, 20.0, 62.5, 7.5, 166.6\"\n")
    exe.chmod(0o755)
    got = bm_rode.run_eval(str(home), 256, "data.mtx")
    assert got == {"Sputnik": 1.25, "cuSPARSE (RoDe driver)": 2.0, "RoDe": 0.75}
    with pytest.raises(FileNotFoundError):
        bm_rode.run_eval(str(home), 512, "data.mtx")


def test_plot_report_from_results_csv(tmp_path):
    """bench/plot.py (reference bench/plot.py:1-146 without matplotlib/seaborn): speed-ups over cuSPARSE per cell, a method
    with no row is 'n/a', cells without a cuSPARSE row are dropped, and the SVG parses."""
    import importlib.util
    import xml.dom.minidom
    spec = importlib.util.spec_from_file_location("vx_plot", os.path.join(ROOT, "bench", "plot.py"))
    plot = importlib.util.module_from_spec(spec); spec.loader.exec_module(plot)
    res = tmp_path / "results.csv"
    res.write_text("Method,Dataset,FeatDim,Reorder,Time (ms)\n"
                   "cuSPARSE,ddi,128,False,0.4\nVoltrix,ddi,128,False,0.1\nRoDe,ddi,128,False,0.2\n"
                   "cuSPARSE,ddi,256,False,0.8\nVoltrix,ddi,256,False,0.4\nRoDe,ddi,256,False,NAN\n"
                   "Voltrix,ppi,128,False,0.3\n"
                   "cuSPARSE,ddi,128,True,0.4\nVoltrix,ddi,128,True,0.05\n")
    sp = plot.speedups(plot.read_results(str(res)))
    assert set(sp) == {"ddi", "ddi.reorder", "ppi"} and sp["ppi"] == {}
    assert sp["ddi"][128] == {"cuSPARSE": 1.0, "Voltrix": 4.0, "RoDe": 2.0}
    assert "RoDe" not in sp["ddi"][256] and sp["ddi.reorder"][128]["Voltrix"] == 8.0
    md = plot.markdown(sp)
    assert "| ddi | 256 | 1.00x | n/a | 2.00x |" in md and "geomean" in md
    xml.dom.minidom.parseString(plot.svg(sp))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "bench", "_competitors", "dtc", "DTCSpMM.so")),
                    reason="DTC-SpMM extension not built (python bench/competitors/build.py --dtc)")
def test_dtc_extension_shares_the_process_libstdcxx():
    """A DTCSpMM.so that carries its own libstdc++.a (what a toolchain wrapper with a dangling libstdc++.so produces) holds
    a second set of locale facet ids and segfaults on its first `std::cout << number` inside a torch process."""
    so = os.path.join(ROOT, "bench", "_competitors", "dtc", "DTCSpMM.so")
    needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
    assert "libstdc++.so.6" in needed
