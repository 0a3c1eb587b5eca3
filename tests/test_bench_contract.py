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


def test_roofline_object_has_the_contract_keys_and_uses_algorithmic_bytes(tmp_path, monkeypatch):
    wl = _FakeJoin()
    res = {"ms_per_step": 18.0, "kernels": {"join_part_probe": {"launches_per_step": 1.0, "ms_per_step": 8.0},
                                            "join_part_scatter": {"launches_per_step": 2.0, "ms_per_step": 4.4}}}
    # a traffic record captured from THESE kernel sources is reported ...
    fresh = tmp_path / "traffic.json"
    fresh.write_text(json.dumps({"csrc_sha1": bench.csrc_sha1(),
                                 "entries": {"join": {"join_part_probe": {"dram_read": 10_200_000_000, "dram_write": 7_950_000_000}}}}))
    monkeypatch.setattr(bench, "TRAFFIC_FILE", str(fresh))
    roof = bench.roofline_for(res, wl, 6545.0, "MEASURED_PEAKS.json hbm_gbs (of measured)", scale=1.0, key="join")
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(roof)
    assert roof["kernel"] == "join_part_probe" and roof["bound"] == "hbm" and roof["unit"] == "GB/s"
    algorithmic = 8 * wl.P + 8 * wl.pairs                        # compact {key32,tag32} pairs in, index pairs out
    assert roof["algorithmic_bytes_per_launch"] == algorithmic
    assert abs(roof["achieved"] - algorithmic / 1e9 / 8.0e-3) < 1e-6
    assert abs(roof["frac"] - roof["achieved"] / 6545.0) < 1e-12
    assert roof["traffic"] == 10_200_000_000 + 7_950_000_000
    assert bench.roofline_for(res, wl, 6545.0, "x", scale=0.25, key="join")["traffic"] is None   # only valid at full size
    # ... one captured from other sources is stale and refused
    stale = tmp_path / "stale.json"
    stale.write_text(json.dumps({"csrc_sha1": "0" * 40, "entries": {"join": {"join_part_probe": {"dram_read": 1, "dram_write": 1}}}}))
    monkeypatch.setattr(bench, "TRAFFIC_FILE", str(stale))
    assert bench.roofline_for(res, wl, 6545.0, "x", scale=1.0, key="join")["traffic"] is None


def test_committed_traffic_file_is_well_formed():
    path = os.path.join(ROOT, "profiles", "r02_traffic_full_size.json")
    if not os.path.isfile(path):
        return
    rec = json.load(open(path))
    assert len(rec["csrc_sha1"]) == 40
    for wl, kernels in rec["entries"].items():
        for k, v in kernels.items():
            assert v["dram_read"] >= 0 and v["dram_write"] >= 0, (wl, k)


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
