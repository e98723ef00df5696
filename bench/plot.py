"""Speed-up-over-cuSPARSE report from results.csv (reference: bench/plot.py:1-146, which draws the same grid with
matplotlib + seaborn -- neither is in this image, so the panels are written as plain SVG and the numbers as a markdown table).

One panel per dataset, one bar group per feature width, one bar per method, height = cuSPARSE time / method time; a method
with no row for a cell is marked "n/a" (the reference prints "CUDA ERROR" there); each panel is captioned with the
Voltrix min-max speed-up, like the reference's x-label.

    python bench/plot.py [--results results.csv] [--out results]      # -> results.svg, results.md
"""
import argparse
import csv
import math
from collections import defaultdict

METHODS = ["cuSPARSE", "Sputnik", "GE-SPMM", "RoDe", "TC-GNN", "DTC-SPMM", "Voltrix", "Voltrix-fp16"]
COLORS = ["#440154", "#46327e", "#365c8d", "#277f8e", "#1fa187", "#4ac16d", "#a0da39", "#fde725"]   # viridis, 8 steps


def read_results(path):
    """results.csv rows (Method,Dataset,FeatDim,Reorder,Time (ms)) -> {dataset: {featdim: {method: ms}}}; the last row of a
    (method, dataset, featdim) wins, reordered runs are kept under '<dataset>.reorder'."""
    data = defaultdict(lambda: defaultdict(dict))
    with open(path, newline="") as fh:
        for row in csv.DictReader(fh):
            try:
                t = float(row["Time (ms)"])
            except (TypeError, ValueError):
                continue
            if not math.isfinite(t) or t <= 0:
                continue
            name = row["Dataset"] + (".reorder" if str(row.get("Reorder", "False")) == "True" else "")
            data[name][int(row["FeatDim"])][row["Method"]] = t
    return data


def speedups(data):
    """{dataset: {featdim: {method: cuSPARSE_ms / method_ms}}} for the cells that have a cuSPARSE row."""
    out = {}
    for ds, by_n in data.items():
        out[ds] = {}
        for n, cell in sorted(by_n.items()):
            base = cell.get("cuSPARSE")
            if base:
                out[ds][n] = {m: base / t for m, t in cell.items()}
    return out


def markdown(sp):
    methods = [m for m in METHODS if any(m in c for d in sp.values() for c in d.values())]
    lines = ["| dataset | N | " + " | ".join(methods) + " |", "|---|---|" + "---|" * len(methods)]
    for ds in sorted(sp):
        for n, cell in sp[ds].items():
            lines.append(f"| {ds} | {n} | " + " | ".join(f"{cell[m]:.2f}x" if m in cell else "n/a" for m in methods) + " |")
    geo = {m: [c[m] for d in sp.values() for c in d.values() if m in c] for m in methods}
    lines.append("| **geomean** | | " + " | ".join(
        f"**{math.exp(sum(map(math.log, v)) / len(v)):.2f}x**" if v else "n/a" for v in geo.values()) + " |")
    return "\n".join(lines) + "\n"


def svg(sp, cols=4, pw=420, ph=260):
    names = sorted(sp)
    methods = [m for m in METHODS if any(m in c for d in sp.values() for c in d.values())]
    rows = max(1, math.ceil(len(names) / cols))
    W, H = cols * pw + 40, rows * ph + 60
    o = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{W}" height="{H}" font-family="sans-serif" font-size="11">',
         f'<rect width="{W}" height="{H}" fill="white"/>']
    for j, m in enumerate(methods):      # legend
        x = 50 + j * (W - 80) / len(methods)
        o.append(f'<rect x="{x:.0f}" y="12" width="22" height="10" fill="{COLORS[METHODS.index(m)]}" stroke="black"/>'
                 f'<text x="{x + 27:.0f}" y="21">{m}</text>')
    for i, ds in enumerate(names):
        x0, y0 = 40 + (i % cols) * pw, 40 + (i // cols) * ph
        gx, gy, gw, gh = x0 + 34, y0 + 22, pw - 54, ph - 74
        cells = sp[ds]
        top = max([v for c in cells.values() for v in c.values()] + [1.0]) * 1.1
        o.append(f'<text x="{gx + gw / 2:.0f}" y="{y0 + 14}" text-anchor="middle" font-size="13" font-weight="bold">{ds}</text>')
        o.append(f'<rect x="{gx}" y="{gy}" width="{gw}" height="{gh}" fill="none" stroke="#888"/>')
        for k in range(5):                # horizontal grid + y ticks
            v = top * k / 4
            y = gy + gh - gh * k / 4
            o.append(f'<line x1="{gx}" y1="{y:.1f}" x2="{gx + gw}" y2="{y:.1f}" stroke="#ccc" stroke-dasharray="3,3"/>'
                     f'<text x="{gx - 4}" y="{y + 4:.1f}" text-anchor="end">{v:.1f}</text>')
        ns = list(cells)
        group = gw / max(len(ns), 1)
        bar = group * 0.84 / len(methods)
        for g, n in enumerate(ns):
            for j, m in enumerate(methods):
                bx = gx + g * group + group * 0.08 + j * bar
                if m in cells[n]:
                    bh = gh * cells[n][m] / top
                    o.append(f'<rect x="{bx:.1f}" y="{gy + gh - bh:.1f}" width="{bar:.1f}" height="{bh:.1f}" '
                             f'fill="{COLORS[METHODS.index(m)]}" stroke="black" stroke-width="0.5"/>')
                else:
                    o.append(f'<text x="{bx + bar / 2:.1f}" y="{gy + gh - 4}" font-size="8" fill="{COLORS[METHODS.index(m)]}" '
                             f'transform="rotate(-90 {bx + bar / 2:.1f} {gy + gh - 4})">n/a</text>')
            o.append(f'<text x="{gx + (g + 0.5) * group:.0f}" y="{gy + gh + 14}" text-anchor="middle">{n}</text>')
        ours = [c[m] for c in cells.values() for m in ("Voltrix", "Voltrix-fp16") if m in c]
        cap = f"{min(ours):.2f}x - {max(ours):.2f}x speedup" if ours else "N/A"
        o.append(f'<text x="{gx + gw / 2:.0f}" y="{gy + gh + 32}" text-anchor="middle" font-size="12">{cap}</text>')
    o.append(f'<text x="14" y="{H / 2:.0f}" transform="rotate(-90 14 {H / 2:.0f})" text-anchor="middle" font-size="13">'
             'Speedup over cuSPARSE</text></svg>')
    return "\n".join(o) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--results", default="results.csv")
    ap.add_argument("--out", default="results", help="output stem: <out>.svg and <out>.md")
    args = ap.parse_args()
    sp = speedups(read_results(args.results))
    if not sp:
        raise SystemExit(f"{args.results}: no cell has a cuSPARSE row to normalise by")
    with open(args.out + ".svg", "w") as fh:
        fh.write(svg(sp))
    with open(args.out + ".md", "w") as fh:
        fh.write(markdown(sp))
    print(markdown(sp), end="")


if __name__ == "__main__":
    main()
