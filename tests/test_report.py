"""CPU: the text-report helpers (analog_pde_solver/report.py, tools/report.py) -- the large-grid form of the
reference demo's ASCII heat-map and centre-column profile (examples/laplace_heat_demo.py:90-110 upstream)."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT
from analog_pde_solver.report import SHADES, ascii_heatmap, centre_column_profile, downsample


def test_small_grids_render_exactly_like_the_reference_demo():
    g = np.linspace(0.0, 1.0, 20)[:, None] * np.ones((20, 20))
    text = ascii_heatmap(g, "title").split("\n")
    assert text[0] == "title" and len(text) == 21
    for i, line in enumerate(text[1:]):
        idx = max(0, min(int(g[i, 0] * (len(SHADES) - 1)), len(SHADES) - 1))   # laplace_heat_demo.py:97-98
        assert line == SHADES[idx] * 40


def test_downsample_is_a_block_mean_that_keeps_the_total():
    rng = np.random.default_rng(0)
    g = rng.random((1000, 777))
    small = downsample(g, 32, 64)
    assert small.shape == (32, 64)
    r = np.linspace(0, 1000, 33).round().astype(int)
    c = np.linspace(0, 777, 65).round().astype(int)
    assert abs(small[3, 5] - g[r[3]:r[4], c[5]:c[6]].mean()) < 1e-12
    counts = np.outer(np.diff(r), np.diff(c))
    assert abs((small * counts).sum() - g.sum()) < 1e-6
    assert downsample(np.ones((8, 8))) .shape == (8, 8)


def test_large_grid_heatmap_and_profile_fit_a_terminal():
    g = np.zeros((4096, 4096))
    g[0] = 1.0
    g[:2048] += np.linspace(1, 0, 2048)[:, None] * 0.5
    text = ascii_heatmap(np.clip(g, 0, 1), "big").split("\n")
    assert "4096 x 4096 nodes shown as 32 x 64 block means" in text[0]
    assert len(text) == 33 and all(len(t) == 128 for t in text[1:])
    prof = centre_column_profile({"Digital": g, "Analog": g * 0.5}, max_rows=10).split("\n")
    assert "10 of 4096 rows" in prof[0] and len(prof) == 12
    assert prof[-1].split()[0] == "4095"


def test_report_tool_renders_bench_and_scale_files(tmp_path):
    line = {"metric": "GLUP/s", "value": 450.0, "unit": "GLUP/s", "n_gpus": 1, "dtype": "f64", "ms_per_step": 19.0,
            "config": {"workload": "16384x16384 Poisson"}, "roofline": {"frac": 1.6, "dram_frac_of_peak": 0.7},
            "e2e": {"value": 260.0}, "clocks": {"sm_mhz": 1800.0, "reasons": ["sw_power_cap"]},
            "cpu_baseline": {"value": 4.0, "cores": 16, "kind": "port", "sample": "tile"},
            "variants": [{"name": "config4_f32", "dtype": "f32", "value": 1000.0, "roofline": {"frac": 1.8}, "scaling": "weak"}]}
    two = dict(line, n_gpus=2, value=890.0, e2e={"value": 470.0}, parity={"multi_gpu_bitwise": True}, variants=[
        {"name": "config5_strong_f64", "dtype": "f64", "value": 700.0, "roofline": {"frac": 1.1}, "scaling": "strong"}])
    bench, scale = tmp_path / "BENCH.json", tmp_path / "SCALE.json"
    bench.write_text(json.dumps({"parsed": line}))
    scale.write_text(json.dumps({"per_n": {"1": {"parsed": line}, "2": {"parsed": two}}}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "report.py"), str(bench), str(scale)],
                         capture_output=True, text=True, check=True).stdout
    assert "Weak scaling" in out and "0.989" in out and "config5_strong_f64" in out and "GPU / CPU" in out
    assert "True" in out
