"""Dynamic instruction mix + stall hot spots of one kernel launch from an ncu report.
   python tools/ncu_source_mix.py <report.ncu-rep> [launch_index]"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = raw.split('"Kernel Name",')[1:]
rows = list(csv.reader(io.StringIO(blocks[which].split("\n", 1)[1])))
hdr = rows[0]
iS, iN, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
mix = collections.Counter(); samp = collections.Counter(); tot = 0
lines = []
for r in rows[1:]:
    if len(r) <= iN or not r[iN].isdigit():
        continue
    op = re.sub(r"^@!?U?P\d+\s+", "", r[iS].strip()).split()[0].rstrip(";")
    n = int(r[iN]); s = int(r[iSm]) if r[iSm].isdigit() else 0
    base = op.split(".")[0]
    mix[base] += n; samp[base] += s; tot += n
    lines.append((s, n, r[iS].strip()))
print(f"total warp-instructions: {tot}")
for op, n in mix.most_common(22):
    print(f"{op:12s} {n:10d} {100*n/tot:6.2f}%   samples {samp[op]}")
print("--- top sampled SASS lines")
for s, n, src in sorted(lines, reverse=True)[:25]:
    print(s, n, src[:100])
