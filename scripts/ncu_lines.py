"""Aggregate an ncu source-page CSV (SASS level) per CUDA source line using nvdisasm line info.
usage: ncu_lines.py <report.ncu-rep> <kernel-substring> [top]"""
import csv, io, os, re, subprocess, sys, collections

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "mvoscalerecovery_b200", "csrc", "libmvosr.so")
tmp = "/tmp/ncu_lines_cub"
os.makedirs(tmp, exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
# per function: list of (offset, file, line)
cur_f, cur_line, table = None, None, collections.defaultdict(list)
for ln in dis.splitlines():
    m = re.match(r"\.text\.(\S+):", ln)
    if m:
        cur_f = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
    if m and cur_f:
        table[cur_f].append((int(m.group(1), 16), cur_line, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if kname not in name:
        continue
    body = b.split("\n", 1)[1]
    rows = list(csv.reader(io.StringIO(body)))
    hdr = rows[0]
    iS, iI, iT, iA = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Address")
    fn = [f for f in table if ("ILb1E" in f) == ("(bool)1" in name) and "frame_kernel" in f]
    tb = table[fn[0]] if fn else None
    base = int(rows[1][iA], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for k, r in enumerate(rows[1:]):
        if len(r) <= iT:
            continue
        off = int(r[iA], 16) - base
        line = tb[k][1] if tb and k < len(tb) else None
        v = (int(r[iS] or 0), int(r[iI] or 0), int(r[iT] or 0))
        for q in range(3):
            agg[line][q] += v[q]; tot[q] += v[q]
    print(name.strip()[:100])
    print("total samples %d  warp-instr %d  thread-instr %d  avg active %.1f" % (tot[0], tot[1], tot[2], tot[2] / max(tot[1], 1)))
    for line, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-28s samples %6.2f%%  winstr %6.2f%%  active %.1f" % ("%s:%s" % line if line else "?", 100.0 * v[0] / tot[0], 100.0 * v[1] / max(tot[1], 1), v[2] / max(v[1], 1)))
