#!/usr/bin/env python3
"""Render bench.py / driver JSON (BENCH_rNN.json, SCALE_rNN.json, or raw bench lines) as the GLUP/s - roofline -
scaling tables.

    python tools/report.py BENCH_r02.json SCALE_r02.json            # driver files
    python bench.py | python tools/report.py -                       # a raw bench line on stdin
"""
import json
import sys


def lines_of(path):
    """Every bench JSON object found in a file: a raw line, a driver BENCH file ({'parsed': ...}) or a driver SCALE
    file ({'per_n': {N: {'parsed': ...}}})."""
    text = sys.stdin.read() if path == "-" else open(path).read()
    out = []
    try:
        doc = json.loads(text)
    except json.JSONDecodeError:
        for ln in text.splitlines():
            ln = ln.strip()
            if ln.startswith("{"):
                try:
                    out.append(json.loads(ln))
                except json.JSONDecodeError:
                    pass
        return out
    if isinstance(doc, dict) and "per_n" in doc:
        for n in sorted(doc["per_n"], key=lambda k: int(k)):
            p = doc["per_n"][n].get("parsed") or {}
            if p:
                out.append(p)
    elif isinstance(doc, dict) and isinstance(doc.get("parsed"), dict):
        out.append(doc["parsed"])
    elif isinstance(doc, dict) and "metric" in doc:
        out.append(doc)
    elif isinstance(doc, list):
        out += [d for d in doc if isinstance(d, dict) and "metric" in d]
    return out


def fmt(x, spec="{:.1f}"):
    return "-" if x is None else spec.format(x)


def table(rows, header):
    widths = [max(len(str(c)) for c in col) for col in zip(header, *rows)]
    def line(cells):
        return "| " + " | ".join(str(c).ljust(w) for c, w in zip(cells, widths)) + " |"
    return "\n".join([line(header), "|" + "|".join("-" * (w + 2) for w in widths) + "|"] + [line(r) for r in rows])


def main():
    benches = []
    for path in sys.argv[1:] or ["-"]:
        benches += lines_of(path)
    ours = [b for b in benches if b.get("impl") != "reference" and "value" in b]
    ref = [b for b in benches if b.get("impl") == "reference"]
    if not ours and not ref:
        raise SystemExit("no bench lines found")
    rows = []
    for b in ours:
        r, e, c = b.get("roofline") or {}, b.get("e2e") or {}, b.get("clocks") or {}
        rows.append([b.get("n_gpus"), b.get("dtype"), (b.get("config") or {}).get("workload", "")[:58], fmt(b.get("value")),
                     fmt(b.get("ms_per_step"), "{:.2f}"), fmt(r.get("frac"), "{:.2f}"), fmt(r.get("dram_frac_of_peak"), "{:.2f}"),
                     fmt(e.get("value")), fmt(c.get("sm_mhz"), "{:.0f}"), ",".join(c.get("reasons") or []) or "-"])
    if rows:
        print("## Headline (device-resident `value`, end-to-end `e2e`)\n")
        print(table(rows, ["GPUs", "dtype", "workload", "GLUP/s", "ms/step", "alg. frac", "DRAM frac", "e2e GLUP/s", "SM MHz", "throttle"]))
    scal = sorted((b for b in ours if b.get("n_gpus")), key=lambda b: b["n_gpus"])
    if len(scal) > 1 and scal[0]["n_gpus"] == 1:
        base, base_e = scal[0]["value"], (scal[0].get("e2e") or {}).get("value")
        rows = []
        for b in scal:
            n, e = b["n_gpus"], (b.get("e2e") or {}).get("value")
            par = (b.get("parity") or {}).get("multi_gpu_bitwise")
            ca = (b.get("e2e") or {}).get("matches_halo_exchange_path")
            rows.append([n, fmt(b["value"]), fmt(b["value"] / (n * base), "{:.3f}"), fmt(e),
                         fmt(None if not (e and base_e) else e / (n * base_e), "{:.3f}"), "-" if par is None else str(par),
                         "-" if ca is None else str(ca)])
        print("\n## Weak scaling (efficiency = value(N) / (N * value(1)))\n")
        print(table(rows, ["GPUs", "GLUP/s", "efficiency", "e2e GLUP/s", "e2e efficiency", "bitwise == 1 GPU",
                           "e2e: overlapped strips == halo exchange"]))
    var = [(b.get("n_gpus"), v) for b in ours for v in (b.get("variants") or [])]
    if var:
        rows = []
        conv = [(n, v) for n, v in var if v.get("unit") == "s"]
        var = [(n, v) for n, v in var if v.get("unit") != "s"]
        for n, v in var:
            r = v.get("roofline") or {}
            rows.append([n, v.get("name"), v.get("dtype"), fmt(v.get("value")), fmt(r.get("frac"), "{:.2f}"),
                         fmt(r.get("dram_frac_of_peak"), "{:.2f}"), fmt((v.get("e2e") or {}).get("value")), v.get("scaling", "-")])
        print("\n## Other BASELINE configs measured in the same run\n")
        print(table(rows, ["GPUs", "config", "dtype", "GLUP/s", "alg. frac", "DRAM frac", "e2e GLUP/s", "scaling"]))
        if conv:
            print("\n## Time to tolerance (4096^2 Poisson, known solution)\n")
            print(table([[v.get("name"), v.get("iterations"), v.get("iteration_unit"), fmt(v.get("value"), "{:.3f}"),
                          fmt(v.get("wall_seconds"), "{:.3f}"), fmt(v.get("max_error_vs_exact_solution"), "{:.2e}"),
                          fmt(v.get("discretisation_error_scale"), "{:.2e}")] for _, v in conv],
                        ["mode", "iterations", "unit", "kernel s", "wall s", "max error", "0.82 h^2"]))
        strong = {}
        for n, v in var:
            if v.get("scaling") == "strong":
                strong.setdefault(v["name"], []).append((n, v["value"]))
        for name, pts in strong.items():
            pts.sort()
            n0, v0 = pts[0]
            print(f"\nstrong scaling {name}: " + ", ".join(f"N={n}: {v:.0f} GLUP/s ({v / v0 * n0 / n:.2f} vs N={n0})" for n, v in pts))
    for b in ours[:1]:
        cb = b.get("cpu_baseline")
        if cb:
            print(f"\nCPU baseline ({cb.get('kind')}, {cb.get('cores')} threads): {cb['value']:.2f} GLUP/s -- {cb.get('sample', '')}")
            print(f"GPU / CPU: device-resident {b['value'] / cb['value']:.0f}x, end to end {((b.get('e2e') or {}).get('value') or 0) / cb['value']:.0f}x")
    for b in ref:
        print(f"\nreference arm: {b.get('value'):.2f} GLUP/s on {(b.get('cpu_baseline') or {}).get('cores')} threads")


if __name__ == "__main__":
    main()
