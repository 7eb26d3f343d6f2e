"""The reference arm of bench.py runs on the host cores, so its JSON contract can be checked without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["unit"] == "images/s" and d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and abs(cb["value"] - d["value"]) < 1e-9
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_metric_matches_baseline_json():
    """bench.py measures BASELINE.json's metric (fwd+bwd images/s of the middle head)."""
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "images/s" in src and "images/s" in json.dumps(base)


def test_roofline_work_model_names_real_entry_points():
    """Every row of bench.py's work model is a C-ABI entry point (scan_<name> in the ctypes table) or a `*_span` pseudo entry, and
    maps to a kernel name for the grouped roofline line where one is declared: a renamed entry point cannot silently drop out of
    the roofline table.  The GroupNorm rows count the streams the kernels actually move (2 forward, 5 backward per pixel row)."""
    sys.path.insert(0, ROOT)
    import bench
    from scan_b200 import _lib
    work = bench.work_model(8, 9, 8888, 7464, [34000, 8500, 2100, 520, 120])
    for name, (bound, amount) in work.items():
        assert bound in ("hbm", "tensor") and amount > 0, name
        assert name.endswith("_span") or ("scan_" + name) in _lib.SIGNATURES, name
    for name in bench.DOMINANT_KERNEL:
        assert name in work, name
    rows_bytes = 8 * bench.L_PER_IMAGE * 1024
    assert work["gn_relu_apply"][1] == 2 * 2 * 2 * rows_bytes        # 2 passes x 2 layers x (1 read + 1 write)
    assert work["gn_relu_bwd"][1] == 2 * 2 * 5 * rows_bytes          # x and dy read twice, dx written
    # the head_in forward launches move to the statistics-carrying entry point; the flops of the group stay the same
    total = work["conv3x3_rows"][1] + work["conv3x3_rows_gn"][1]
    assert total == 2 * 5 * (2 * 8 * bench.L_PER_IMAGE * 256 * 256 * 9)
