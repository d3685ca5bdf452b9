"""profiles/traffic.json (read by bench.py for `roofline.traffic`) must be what profiles/make_traffic.py derives from the
committed ncu capture it names -- the figure cannot be edited by hand or go stale against the capture.  CPU only."""
import json
import os
import sys

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "profiles"))


def test_traffic_json_matches_the_committed_ncu_capture():
    import make_traffic
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    src = os.path.join(ROOT, t["source"].split(" ")[0])
    assert os.path.exists(src), src
    d = make_traffic.derive(src)
    assert d["iqbb_accum_f32_c2_bytes_per_launch"] == t["iqbb_accum_f32_c2_bytes_per_launch"]
    assert d["samples_per_launch"] == t["samples_per_launch"]
    # sanity: between 1.0x and 1.1x of the algorithmic bytes (8 B per sample + 4 B per output)
    alg = t["samples_per_launch"] * 8 + (t["samples_per_launch"] // 416) * 4
    assert 0.99 * alg < t["iqbb_accum_f32_c2_bytes_per_launch"] < 1.10 * alg
    # the summary that sits next to the details names the same kernel
    summ = src.replace("_details.csv", "_summary.md")
    assert os.path.exists(summ) and "iqbb_fold_f32_win_kernel" in open(summ).read()
