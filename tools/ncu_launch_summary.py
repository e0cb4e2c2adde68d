"""Summarise an ncu launch list CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch) of
one denoise step into profiles/<tag>_ncu_launch_list_step_c3.txt and profiles/<tag>_engine_traffic.json.
usage: ncu_launch_summary.py gpurun_out/launches.csv r01b"""
import collections
import csv
import json
import re
import sys

path, tag = sys.argv[1], sys.argv[2]
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr = rows[0]
ik, im, iv, iu, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[iid], {"name": r[ik]})
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    if r[im] == "gpu__time_duration.sum":
        d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    else:
        d[r[im]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
agg = collections.OrderedDict()
for d in per.values():
    name = re.sub(r"^void ", "", d["name"])
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("i360::", "")
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("ms", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
lines = ["I360_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
         "--clock-control none python bench.py --steps 1 --warmup 1 --no-cpu-baseline",
         "exactly ONE dual-branch denoise step at 16x512x1024 (cudaProfilerStart/Stop around it); per-launch times are serialised + "
         "cold-cache: compare SHARES",
         f"launches {len(per)}, sum of kernel time {tot:.1f} ms", ""]
for name, (n, ms, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{ms:9.2f} ms {100 * ms / tot:5.1f}%  x{n:4d}  dram {by / 1e9:8.2f} GB  {name[:110]}")
open(f"profiles/{tag}_ncu_launch_list_step_c3.txt", "w").write("\n".join(lines) + "\n")
eng = [(n, ms, by) for name, (n, ms, by) in agg.items() if name.startswith("gemm_conv_kernel")]
n_e, ms_e, by_e = sum(e[0] for e in eng), sum(e[1] for e in eng), sum(e[2] for e in eng)
json.dump({"kernel": "gemm_conv_kernel", "launches": n_e, "dram_bytes_per_launch": by_e / max(1, n_e), "dram_bytes_per_step": by_e,
           "share_of_step_kernel_time": ms_e / tot, "source": f"profiles/{tag}_ncu_launch_list_step_c3.txt"},
          open(f"profiles/{tag}_engine_traffic.json", "w"), indent=1)
print("\n".join(lines[:30]))
