"""Selected metrics of an `ncu --set full` capture exported with `--page raw --csv`
(scripts/ncu_profile.sh -> gpurun_out/ncu/full_raw.csv) as a small JSON under profiles/.
Usage: python scripts/summarize_ncu_full.py <full_raw.csv> <out.json>"""
import csv
import json
import re
import sys

WANT = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__shared_mem_per_block_dynamic",
]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    cols = {}
    for w in WANT:
        for i, h in enumerate(hdr):
            if h == w or h.endswith("." + w):
                cols[w] = i
                break
    ik, ig, ib = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    out = []
    for r in rows[2:]:
        if not r or not r[0].isdigit():
            continue
        d = {"Kernel Name": r[ik][:160], "Grid Size": r[ig], "Block Size": r[ib]}
        for w, i in cols.items():
            try:
                d["%s [%s]" % (w, units[i])] = float(r[i].replace(",", ""))
            except ValueError:
                d["%s [%s]" % (w, units[i])] = r[i]
        m = re.search(r"(\w+_kernel)", r[ik])
        d["kernel"] = m.group(1) if m else r[ik][:40]
        out.append(d)
    json.dump({"source": "ncu --set full --clock-control none --nvtx-include mcn_profiled_step/ -k <conv|wgrad|stem|bn|"
                         "maxpool kernels> -s 6 -c 40 on `bench.py --steps 1 --warmup 3 --no-graph` "
                         "(scripts/ncu_profile.sh); cold-cache, serialised", "launches": out},
              open(dst, "w"), indent=1)
    print("wrote %s: %d launches" % (dst, len(out)))


if __name__ == "__main__":
    main()
