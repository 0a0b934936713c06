"""Turns the ncu launch list of one eager training step (scripts/ncu_profile.sh) into the files
committed under profiles/: per-launch table, per-kernel shares, and the per-class DRAM traffic that
bench.py reports as roofline.traffic.  Usage: python scripts/summarize_ncu.py <launches.csv> <tag>"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    m = re.search(r"(\w+_kernel)", name)
    return m.group(1) if m else name[:40]


def main():
    src, tag = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(src)) if r and r[0].isdigit() or (r and r[0] == "ID")]
    hdr = rows[0]
    iid, ik, im, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    igrid, iblk = hdr.index("Grid Size"), hdr.index("Block Size")
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(int(r[iid]), {"kernel": short(r[ik]), "grid": r[igrid], "block": r[iblk]})
        d[r[im]] = float(r[iv].replace(",", ""))
    # classes = bench.py's KERNEL_OF groups (launches are grouped by CUDA kernel family)
    out = []
    for i, d in launches.items():
        k = d["kernel"]
        if k in ("gemm_conv_kernel", "halo_conv_kernel"):
            cls = "conv_tc"
        elif k in ("wgrad_kernel", "wgrad_halo_kernel") or k.startswith("splitk_reduce"):
            cls = "wgrad_tc"
        elif k.startswith("bn_bwd_apply"):
            cls = "bn_bwd_apply"
        elif k.startswith("bn_bwd_reduce") or k.startswith("bn_reduce_pipe"):
            cls = "bn_bwd_reduce"
        elif k.startswith("bn_apply"):
            cls = "bn_apply"
        elif k.startswith("bn_stats"):
            cls = "bn_stats"
        elif k.startswith("stem_"):
            cls = "stem"
        elif k.startswith("maxpool"):
            cls = "maxpool"
        else:
            cls = k.replace("_kernel", "")
        d["class"] = cls
        out.append((i, d))
    total_ns = sum(d.get("gpu__time_duration.sum", 0.0) for _, d in out)
    prof = os.path.join(ROOT, "profiles")
    with open(os.path.join(prof, "%s_ncu_launches.csv" % tag), "w") as f:
        f.write("id,kernel,class,grid,block,duration_us,dram_read_MB,dram_write_MB\n")
        for i, d in out:
            f.write("%d,%s,%s,\"%s\",\"%s\",%.2f,%.3f,%.3f\n" % (
                i, d["kernel"], d["class"], d["grid"], d["block"], d.get("gpu__time_duration.sum", 0) / 1e3,
                d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6))
    agg = collections.OrderedDict()
    for _, d in out:
        a = agg.setdefault(d["class"], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    with open(os.path.join(prof, "%s_ncu_launch_shares.csv" % tag), "w") as f:
        f.write("class,launches,total_us,share_of_step,dram_MB_per_step,dram_GBps\n")
        for c, (n, ns, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.1f,%.4f,%.1f,%.0f\n" % (c, n, ns / 1e3, ns / total_ns, b / 1e6, b / ns if ns else 0))
    json.dump({"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                         "--clock-control none over one eager step of `bench.py --steps 1 --warmup 3 "
                         "--no-cpu-baseline --no-graph` (scripts/ncu_profile.sh); per-launch times are "
                         "cold-cache and serialised",
               "step_us_under_ncu": total_ns / 1e3,
               "classes": {c: {"launches": n, "dram_bytes_per_launch": b / n, "dram_bytes_per_step": b,
                               "share_of_step": ns / total_ns}
                           for c, (n, ns, b) in agg.items()}},
              open(os.path.join(prof, "ncu_traffic.json"), "w"), indent=1)
    import shutil
    shutil.copyfile(os.path.join(prof, "ncu_traffic.json"), os.path.join(prof, "%s_ncu_traffic.json" % tag))
    print("launches %d, step under ncu %.2f ms" % (len(out), total_ns / 1e6))
    for c, (n, ns, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
        print("%-18s n=%3d %8.1f us  %5.1f%%  %8.1f MB  %6.0f GB/s" % (c, n, ns / 1e3, 100 * ns / total_ns, b / 1e6, b / ns if ns else 0))


if __name__ == "__main__":
    main()
