#!/usr/bin/env python3
"""profiles/r2_ncu_rb_stream_<dtype>_<depth>_<noise>_<src>.txt (tools/ncu_summary.py output of `ncu --set full`)
-> profiles/ncu_traffic.json: DRAM bytes per launch of the dominant kernel per variant, read by bench.py for
roofline.traffic (the ncu counters cannot be read while timing, so the capture of the same kernel on the same
workload, taken in the same round by tools/profile_all.sh, is what the bench line cites)."""
import glob, json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {}
for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r2_ncu_rb_stream_*.txt"))):
    m = re.search(r"rb_stream_(f64|f32)_(\d)_(\d)_(\d)\.txt$", path)
    if not m:
        continue
    dt, depth, noise, src = m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4))
    vals = {}
    for ln in open(path):
        f = ln.split()
        if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
            v, unit = float(f[1]), f[2]
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}[unit]
            vals[f[0]] = v * scale
        if ln.startswith("Kernel Name"):
            vals["kernel"] = re.sub(r"\s+", " ", ln[len("Kernel Name"):]).strip()
    rd, wr, t = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum"), vals.get("gpu__time_duration.sum")
    if rd is None or wr is None:
        continue
    out[f"{dt}_d{depth}_noise{noise}_src{src}_16384x16384"] = {
        "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
        "ncu_launch_seconds": t, "ncu_dram_gbs": (rd + wr) / t / 1e9 if t else None,
        "kernel": vals.get("kernel", "rb_stream_kernel"),
        "source": f"profiles/{os.path.basename(path)} (ncu --set full --clock-control none, one launch = {depth} fused sweep(s) of 16384^2)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
for k, v in out.items():
    print(f"{k}: {v['dram_bytes_per_launch'] / 1e9:.2f} GB per launch, {v['ncu_dram_gbs']:.0f} GB/s under ncu")
