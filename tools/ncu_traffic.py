"""Turn an `ncu --set full` capture of the dominant kernel (exported with `ncu -i X.ncu-rep --page raw --csv`, or the text
summary tools/ncu_summary.py writes) into profiles/roofline_traffic.json, which bench.py reads for `roofline.traffic`.
The record carries a hash of the kernel's source so a capture from another build is reported as stale.

    python tools/ncu_traffic.py profiles/r2_ncu_full_conv_s0k11.txt "stage 0, k=11 conv, 768 ch, 15040 rows: 59 MB algorithmic incl. 13 MB weights"
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import _src_hash  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    path, what = sys.argv[1], sys.argv[2]
    total, found = 0.0, 0
    for ln in open(path):
        ln = ln.strip()
        for key in ("dram__bytes_read.sum [", "dram__bytes_write.sum ["):
            if ln.startswith(key):
                u = ln[len(key):ln.index("]")]
                total += float(ln.split("=")[1].replace(",", "")) * UNIT.get(u, 1.0)
                found += 1
    if found != 2:
        raise SystemExit(f"{path}: expected one dram__bytes_read.sum and one dram__bytes_write.sum line, found {found}")
    rec = {"dram_bytes": total, "capture": os.path.relpath(path, ROOT), "what": "ncu --set full, one launch (" + what + ")",
           "kernel_source_sha1_16": _src_hash("bvg_conv_umma.cu")}
    json.dump(rec, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
    print(rec)


if __name__ == "__main__":
    main()
