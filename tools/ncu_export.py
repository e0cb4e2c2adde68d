"""Export the key metrics + top stall lines of an .ncu-rep into a small text file under profiles/.
usage: ncu_export.py gpurun_out/x.ncu-rep profiles/r01b_ncu_x.txt "<note>" """
import csv, io, subprocess, sys, collections
rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, r = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
lines = [f"ncu --set full --clock-control none --import-source on  ({rep.split('/')[-1]})", note, ""]
for w in want:
    if w in h:
        i = h.index(w)
        lines.append(f"{w} = {r[i]} {u[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
hdr = srows[1]
isamp, isrc = hdr.index("# Samples"), hdr.index("Source")
stall = [(i, x) for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
tot = collections.Counter(); data = []
for k, row in enumerate(srows[2:]):
    try:
        s = float(row[isamp])
    except Exception:
        continue
    data.append((s, k, row))
    for i, x in stall:
        try:
            tot[x] += float(row[i] or 0)
        except Exception:
            pass
ssum = sum(tot.values()) or 1
lines += ["", "warp-state samples by reason: " + ", ".join(f"{k[6:]} {100 * v / ssum:.1f}%" for k, v in tot.most_common(8)), "",
          "top SASS lines by samples:"]
tsum = sum(d[0] for d in data) or 1
for s, k, row in sorted(data, key=lambda x: -x[0])[:14]:
    top = sorted(((float(row[i] or 0), x) for i, x in stall), reverse=True)[0]
    lines.append(f"  {100 * s / tsum:5.1f}%  #{k:5d} {row[isrc].strip()[:72]:72s} {top[1]}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:20]))
