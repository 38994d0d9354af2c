#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of the dominant kernel, stamped with a hash of the kernel
sources it was taken from.  bench.py reports `roofline.traffic` from this file ONLY while that hash still matches
the sources in the tree (otherwise null with a note): a capture cannot silently outlive the code it measured.

    python tools/refresh_traffic.py gpurun_out/prof_r2.ncu-rep [profiles/r02_ncu_full.csv]
"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["lq_fq.cu", "lq_fq_ops.cuh", "lq_traverse.cuh", "lq_common.cuh"]     # what the dominant kernel is compiled from


def source_hash():
    h = hashlib.sha256()
    for name in SOURCES:
        with open(os.path.join(ROOT, "learned_quantization_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def main(rep, summary=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    col = {m: hdr.index(m) for m in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum")}
    units = rows[1]
    mult = lambda c: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[c]]
    picks = [r for r in rows[2:] if "BwdThresholdOp" in r[col["Kernel Name"]] and "row_kernel" in r[col["Kernel Name"]]]
    if not picks:
        raise SystemExit("no row_kernel<BwdThresholdOp> launch in this capture")
    r = picks[-1]
    rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * mult(col["dram__bytes_read.sum"])
    wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * mult(col["dram__bytes_write.sum"])
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out = {"kernel": r[col["Kernel Name"]][:80], "workload": "fakequant_fwd_bwd_channelwise_C512_inner512",
           "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "traffic_bytes_per_launch": int(rd + wr),
           "algorithmic_bytes_per_launch": 12 * 256 * 512 * 512,
           "duration_under_ncu": f"{r[col['gpu__time_duration.sum']]} {units[col['gpu__time_duration.sum']]}",
           "source_sha256": source_hash(), "sources": SOURCES, "commit_at_capture": commit,
           "capture": f"{os.path.basename(rep)} (scratch, not committed)" + (f"; summary in {summary}" if summary else ""),
           "command": "ncu --set full --clock-control none -k regex:row_kernel python bench.py --steps 5 --warmup 3 --no-qat"}
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:3])
