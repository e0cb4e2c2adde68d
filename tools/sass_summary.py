"""Instruction histogram per kernel of the in-tree library (what proves a Blackwell-native kernel: UTC*MMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, HMMA = legacy mma.sync).  usage: sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "imagine360_b200", "libimagine360_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "HMMA", "SYNCS", "MUFU", "FFMA2", "FADD2", "FMUL2", "LDGSTS", "ATOM", "RED"]
kern, hist, order = None, collections.defaultdict(collections.Counter), []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("i360::", "")
        order.append(kern)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern]["_total"] += 1
        op = m.group(1)
        for k in KEY:
            if op.startswith(k):
                hist[kern][k] += 1
print(f"cuobjdump -sass {os.path.relpath(lib, root)}   (sm_100a only; {len(order)} kernels)")
print("totals: " + ", ".join(f"{k} {sum(h[k] for h in hist.values())}" for k in KEY))
print()
print(f"{'kernel':92s} {'instr':>6s}  " + " ".join(f"{k:>7s}" for k in KEY[:10]))
for k in order:
    h = hist[k]
    print(f"{k[:92]:92s} {h['_total']:6d}  " + " ".join(f"{h[x]:7d}" for x in KEY[:10]))
