"""Host-side scheduling logic that needs no GPU: launch ranges, bench workload split."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_launch_ranges_cover_the_slab_once():
    from pystella_b200.fused import _launch_ranges
    for Nx in (4, 16, 31, 32, 33, 64, 200):
        for h in (1, 2, 4):
            for chunk in (16, 64):
                assert _launch_ranges(Nx, h, False, chunk) == [(0, Nx)]
                ranges = _launch_ranges(Nx, h, True, chunk)
                flat = []
                for r in ranges:
                    flat += list(r) if isinstance(r[0], tuple) else [r]
                covered = np.zeros(Nx, int)
                for b, e in flat:
                    assert 0 <= b < e <= Nx
                    covered[b:e] += 1
                assert (covered == 1).all(), (Nx, h, chunk, ranges)
                # the faces (outermost h planes on either side) are in the FIRST launch
                first = ranges[0]
                first = list(first) if isinstance(first[0], tuple) else [first]
                in_first = np.zeros(Nx, bool)
                for b, e in first:
                    in_first[b:e] = True
                assert in_first[:min(h, Nx)].all() and in_first[-min(h, Nx):].all()


def test_bench_workload_partition():
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    for world in (1, 2, 3, 8):
        shapes = [bench.workload(512, "f64", world, r, "strong")["rank_shape"]
                  for r in range(world)]
        assert sum(s[0] for s in shapes) == 512
        assert all(s[1:] == (512, 512) for s in shapes)
    w = bench.workload(1024, "f32", 8, 3, "weak", nx=1024)
    assert w["grid_shape"] == (1024, 1024, 1024) and w["rank_shape"] == (128, 1024, 1024)
    assert w["np_dtype"] == np.float32
    f, d = bench.initial_fields(bench.workload(16, "f64", 1, 0, "strong"))
    assert f.shape == (1, 20, 20, 20) and abs(f.mean() - 0.193) < 1e-5


def test_reference_arm_prints_contract_line():
    """bench.py --impl reference on a tiny lattice: one JSON line with the agreed keys."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl",
                          "reference", "--grid", "32", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                "higher_is_better", "scaling", "dtype", "data", "config", "cpu_baseline",
                "e2e"):
        assert key in line
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port"
    assert line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
