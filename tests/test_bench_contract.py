"""CPU: the pieces of bench.py's JSON contract that do not need a GPU - roofline arithmetic, the committed
full-size DRAM traffic, the pyarrow baseline object and the single-line stdout discipline."""
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class _FakeJoin(bench.JoinWorkload):
    def __init__(self):
        self.P, self.B, self.pairs = 1_000_000_000, 100_000_000, 1_000_000_000


def test_roofline_object_has_the_contract_keys_and_uses_algorithmic_bytes():
    wl = _FakeJoin()
    res = {"ms_per_step": 25.0, "kernels": {"join_part_probe": {"launches_per_step": 1.0, "ms_per_step": 13.8},
                                            "join_part_scatter": {"launches_per_step": 2.0, "ms_per_step": 5.7}}}
    roof = bench.roofline_for(res, wl, 6545.0, "MEASURED_PEAKS.json hbm_gbs (of measured)", scale=1.0)
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(roof)
    assert roof["kernel"] == "join_part_probe" and roof["bound"] == "hbm" and roof["unit"] == "GB/s"
    algorithmic = 12 * wl.P + 8 * wl.pairs                       # {key,tag} pairs in, index pairs out
    assert roof["algorithmic_bytes_per_launch"] == algorithmic
    assert abs(roof["achieved"] - algorithmic / 1e9 / 13.8e-3) < 1e-6
    assert abs(roof["frac"] - roof["achieved"] / 6545.0) < 1e-12
    assert roof["traffic"] == 17211175936 + 7982386688           # profiles/r01c_traffic_full_size.json
    assert bench.roofline_for(res, wl, 6545.0, "x", scale=0.25)["traffic"] is None   # only valid at full size


def test_traffic_file_matches_the_kernels_bench_reports():
    rec = json.load(open(os.path.join(ROOT, "profiles", "r01c_traffic_full_size.json")))
    for kernel in ("join_part_probe", "groupby_build_fast", "select"):
        assert rec[kernel]["dram_read"] + rec[kernel]["dram_write"] >= 0.95 * rec[kernel]["algorithmic"]


def test_pyarrow_baseline_object():
    out = bench.cpu_baseline_pyarrow(0.001)                      # tiny sample
    assert out["cores"] >= 1 and out["unit"] == "rows/s"
    for part in ("join", "groupby", "filter"):
        assert out[part]["value"] > 0 and isinstance(out[part]["sample"], str)


def test_emit_writes_exactly_one_json_line():
    buf = io.StringIO()
    old, bench._REAL_STDOUT = bench._REAL_STDOUT, buf
    try:
        bench.emit({"metric": "m", "value": 1.5})
    finally:
        bench._REAL_STDOUT = old
    lines = buf.getvalue().splitlines()
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "m", "value": 1.5}
