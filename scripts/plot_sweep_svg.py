"""Plots of the benchmark sweep (the reference's scripts/benchmark.py:178-180 saves forward time, forward+backward time and
peak memory against the number of queries as figures; this image has no matplotlib, so the three figures are written as
plain SVG): python scripts/plot_sweep_svg.py profiles/r2c_benchmark_sweep.csv  ->  <csv stem>_{fwd,fwdbwd,memory}.svg"""
import csv
import math
import sys
from pathlib import Path

COLORS = {"cuda": "#d62728", "torch": "#1f77b4", "reference_triton": "#2ca02c"}
W, H, ML, MR, MT, MB = 640, 420, 70, 20, 40, 55


def ticks(lo, hi):
    return [10.0 ** e for e in range(math.floor(math.log10(lo)), math.ceil(math.log10(hi)) + 1)]


def plot(rows, column, title, ylabel, out):
    series = {}
    for r in rows:
        series.setdefault(r["provider"], []).append((float(r["num_queries"]), float(r[column])))
    xs = [x for s in series.values() for x, _ in s]
    ys = [y for s in series.values() for _, y in s]
    xt, yt = ticks(min(xs), max(xs)), ticks(min(ys), max(ys))
    x0, x1, y0, y1 = math.log10(xt[0]), math.log10(xt[-1]), math.log10(yt[0]), math.log10(yt[-1])
    px = lambda x: ML + (math.log10(x) - x0) / (x1 - x0) * (W - ML - MR)          # noqa: E731
    py = lambda y: H - MB - (math.log10(y) - y0) / (y1 - y0) * (H - MT - MB)      # noqa: E731
    o = [f'<svg xmlns="http://www.w3.org/2000/svg" width="{W}" height="{H}" font-family="sans-serif" font-size="12">',
         f'<rect width="{W}" height="{H}" fill="white"/>',
         f'<text x="{W / 2}" y="22" text-anchor="middle" font-size="14">{title}</text>']
    for t in xt:
        o.append(f'<line x1="{px(t):.1f}" y1="{MT}" x2="{px(t):.1f}" y2="{H - MB}" stroke="#ddd"/>')
        o.append(f'<text x="{px(t):.1f}" y="{H - MB + 16}" text-anchor="middle">{t:g}</text>')
    for t in yt:
        o.append(f'<line x1="{ML}" y1="{py(t):.1f}" x2="{W - MR}" y2="{py(t):.1f}" stroke="#ddd"/>')
        o.append(f'<text x="{ML - 6}" y="{py(t) + 4:.1f}" text-anchor="end">{t:g}</text>')
    o.append(f'<rect x="{ML}" y="{MT}" width="{W - ML - MR}" height="{H - MT - MB}" fill="none" stroke="black"/>')
    o.append(f'<text x="{W / 2}" y="{H - 12}" text-anchor="middle">number of queries</text>')
    o.append(f'<text x="16" y="{H / 2}" text-anchor="middle" transform="rotate(-90 16 {H / 2})">{ylabel}</text>')
    for i, (name, pts) in enumerate(sorted(series.items())):
        pts.sort()
        c = COLORS.get(name, "#555")
        o.append('<polyline fill="none" stroke="%s" stroke-width="2" points="%s"/>'
                 % (c, " ".join(f"{px(x):.1f},{py(y):.1f}" for x, y in pts)))
        o += [f'<circle cx="{px(x):.1f}" cy="{py(y):.1f}" r="3" fill="{c}"/>' for x, y in pts]
        o.append(f'<rect x="{ML + 12}" y="{MT + 10 + 18 * i}" width="14" height="4" fill="{c}"/>')
        o.append(f'<text x="{ML + 32}" y="{MT + 16 + 18 * i}">{name}</text>')
    o.append("</svg>")
    Path(out).write_text("\n".join(o))
    print(out)


def main():
    src = Path(sys.argv[1] if len(sys.argv) > 1 else "profiles/r1_benchmark_sweep.csv")
    rows = list(csv.DictReader(open(src)))
    stem = str(src.with_suffix(""))
    plot(rows, "fwd_ms", "forward, B=4 H=8 D=32 L=4 K=4 (B200)", "ms", stem + "_fwd.svg")
    plot(rows, "fwd_bwd_ms", "forward + backward", "ms", stem + "_fwdbwd.svg")
    plot(rows, "peak_extra_memory_mb", "peak extra memory, forward + backward", "MB", stem + "_memory.svg")


if __name__ == "__main__":
    main()
