#!/usr/bin/env python
"""Regenerates profiles/traffic.json (the `roofline.traffic` figure bench.py reports) from the tracked ncu details of the
dominant kernel:

    python profiles/make_traffic.py [profiles/<stem>_details.csv]

traffic = dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the float accumulate kernel on the bench's pass
(256 buffers of 2^20 samples), exactly as `ncu --set full` recorded it (profiles/summarize_ncu.py wrote the csv from the
.ncu-rep).  tests/test_profiles_consistency.py asserts that the json equals what this script derives, so the number in
the bench line cannot drift away from the committed capture."""
import csv
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT = os.path.join(HERE, "r02_fold_win_kernel_details.csv")
MULT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def derive(details_csv):
    rows = {r[0]: (r[1], r[2]) for r in csv.reader(open(details_csv)) if len(r) >= 3}
    def num(key):
        unit, val = rows[key]
        return float(val.replace(",", "")) * MULT.get(unit, 1.0)
    grid = rows.get("launch__grid_size", ("", "0"))[1]
    return {"iqbb_accum_f32_c2_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
            "kernel": rows.get("Kernel Name", ("", ""))[1], "grid": grid,
            "kernel_us_under_ncu": num("gpu__time_duration.sum") if rows["gpu__time_duration.sum"][0] == "us" else None,
            "source": "profiles/%s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch)" % os.path.basename(details_csv),
            "samples_per_launch": 268435456}


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else DEFAULT
    out = derive(src)
    json.dump(out, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
