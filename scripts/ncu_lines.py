"""Aggregate an ncu source-page CSV (SASS level) per CUDA source line using nvdisasm line info.
The SASS listing of a kernel in ncu is the kernel followed by its callees; functions are aligned by
matching instruction text.  usage: ncu_lines.py <report.ncu-rep> <kernel-substring> [top] [sass]
MVOSR_CALLER=1: instructions inlined from CUDA's own headers (shuffles, ballots, reductions, atomics) are attributed to the line
of OUR source that called them (nvdisasm -gi prints the inlining chain)."""
import csv, io, os, re, subprocess, sys, collections

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
show_sass = len(sys.argv) > 4
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "mvoscalerecovery_b200", "csrc", "libmvosr.so")
tmp = "/tmp/ncu_lines_cub"
os.makedirs(tmp, exist_ok=True)
for f in os.listdir(tmp):
    os.remove(os.path.join(tmp, f))
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
CALLER = bool(os.environ.get("MVOSR_CALLER"))
dis = "\n".join(subprocess.run(["nvdisasm", "-gi" if CALLER else "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
                for f in sorted(os.listdir(tmp)) if f.endswith(".cubin"))          # one cubin per translation unit
cur_f, cur_line, table = None, None, collections.defaultdict(list)
cur_path, in_chain = "", False
norm = lambda t: re.sub(r"\s+", " ", re.sub(r"`\([^)]*\)|0x[0-9a-f]+", "#", t.strip().rstrip(";"))).strip()
for ln in dis.splitlines():
    m = re.match(r"\.text\.(\S+):", ln)
    if m:
        cur_f = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        # -gi: the chain innermost -> outermost on consecutive lines; keep the innermost frame that is not a CUDA header
        if not (CALLER and in_chain and "/cuda/" not in cur_path):
            cur_path = m.group(1); cur_line = (os.path.basename(cur_path), int(m.group(2)))
        in_chain = True
        continue
    in_chain = False
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
    if m and cur_f:
        table[cur_f].append((int(m.group(1), 16), cur_line, norm(m.group(2))))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name",')
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if kname not in name:
        continue
    body = b.split("\n", 1)[1]
    rows = [r for r in csv.reader(io.StringIO(body))]
    hdr = rows[0]
    iS, iI, iT, iA, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Address"), hdr.index("Source")
    rows = [r for r in rows[1:] if len(r) > iT]
    texts = [norm(r[iSrc]) for r in rows]
    # align every function of the cubin inside the listing (greedy: first window of 12 matching instructions)
    line_of = [None] * len(rows)
    func_of = [None] * len(rows)
    used = 0
    for fn, tb in table.items():
        if not tb:
            continue
        key = [t[2] for t in tb[:12]]
        for s0 in range(0, len(texts) - len(tb) + 1):
            if line_of[s0] is None and texts[s0:s0 + len(key)] == key:
                # several template instances share prefixes: require the kernel name flavour to match for kernels
                if "frame_kernel" in fn and (("ILb1E" in fn) != ("(bool)1" in name)):
                    break
                for k, t in enumerate(tb):
                    line_of[s0 + k] = t[1]; func_of[s0 + k] = fn
                used += len(tb)
                break
    if used == 0:
        # the kernel's section holds all its callees: align by index when the lengths agree
        for fn, tb in table.items():
            if len(tb) == len(rows) and "frame_kernel" in fn and (("ILb1E" in fn) == ("(bool)1" in name)):
                for k, t in enumerate(tb):
                    line_of[k] = t[1]
                used = len(tb)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for k, r in enumerate(rows):
        v = (int(r[iS] or 0), int(r[iI] or 0), int(r[iT] or 0))
        for q in range(3):
            agg[line_of[k]][q] += v[q]; tot[q] += v[q]
    print(name.strip()[:100], " (aligned %d of %d SASS rows)" % (used, len(rows)))
    print("total samples %d  warp-instr %d  thread-instr %d  avg active %.1f" % (tot[0], tot[1], tot[2], tot[2] / max(tot[1], 1)))
    for line, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-28s samples %6.2f%%  winstr %6.2f%%  active %.1f" % ("%s:%s" % line if line else "?", 100.0 * v[0] / tot[0], 100.0 * v[1] / max(tot[1], 1), v[2] / max(v[1], 1)))
    if show_sass:
        order = sorted(range(len(rows)), key=lambda k: -int(rows[k][iI] or 0))[:top]
        for k in order:
            print("%10s %5.2f%% %-22s %s" % (rows[k][iI], 100.0 * int(rows[k][iI] or 0) / max(tot[1], 1), "%s:%s" % line_of[k] if line_of[k] else "?", rows[k][iSrc].strip()[:90]))
