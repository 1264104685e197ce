"""Regenerate the ncu summaries under profiles/ from gpurun_out/ (run in the build container after a profiling gpurun):
  profiles/ncu_r01_frame_kernel_bench.txt   selected metrics of the --set full capture (gpurun_out/prof_r01_final.ncu-rep)
  profiles/traffic.json                     dram bytes of that launch (read by bench.py for roofline.traffic)
  profiles/ncu_r01_frame_kernel_regions.txt instruction / sample shares per source region
  profiles/launches_r01_summary.txt         per-kernel shares of the launch list (profiles/launches_r01.csv)
traffic.json also records the warp-instruction count of the launch and the hash of the kernel sources the capture was taken from
(bench.kernel_source_hash): bench.py refuses the counters when the sources have changed since.
usage: python scripts/summarize_profiles.py [report.ncu-rep] [launches.csv] [round tag, e.g. r02]"""
import collections, csv, datetime, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                                            # kernel_source_hash
RT = sys.argv[3] if len(sys.argv) > 3 else "r02"
rep = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "prof_%s_final.ncu-rep" % RT)
launches = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "launches_%s.csv" % RT)
P = os.path.join(ROOT, "profiles")
WANT = ['Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__time_duration.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__block_size', 'launch__grid_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
        'sm__inst_executed.avg.per_cycle_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio']

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
out, d = [], {}
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        out.append('%-70s %-16s %s' % (w, units[i], vals[i])); d[w] = (units[i], vals[i])
tob = lambda u, v: float(v.replace(',', '')) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}[u]
r, w = tob(*d['dram__bytes_read.sum']), tob(*d['dram__bytes_write.sum'])
out.append('traffic %.1f' % (r + w))
open(os.path.join(P, 'ncu_%s_frame_kernel_bench.txt' % RT), 'w').write('\n'.join(out) + '\n')
json.dump({'frame_kernel_dram_bytes_per_launch': r + w,
           'frame_kernel_warp_instructions_per_launch': float(d['smsp__inst_executed.sum'][1].replace(',', '')),
           'kernel_source_hash': bench.kernel_source_hash(), 'captured': datetime.date.today().isoformat(),
           'source': 'ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum and smsp__inst_executed.sum, frame_kernel<FROM_CORR> on the '
                     'bench workload (4541 frames x 2500 correspondences), profiles/ncu_%s_frame_kernel_bench.txt' % RT, 'read_bytes': r, 'write_bytes': w},
          open(os.path.join(P, 'traffic.json'), 'w'), indent=1)

lines = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_lines.py"), rep, "frame_kernel", "400"], capture_output=True, text=True).stdout
open("/tmp/ncu_lines.txt", "w").write(lines)
reg = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_regions.py"), "/tmp/ncu_lines.txt"], capture_output=True, text=True).stdout
open(os.path.join(P, 'ncu_%s_frame_kernel_regions.txt' % RT), 'w').write(reg)

rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
h = rows[0]; ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.defaultdict(list)
for row in rows[1:]:
    v = float(row[vi].replace(',', '')); u = row[ui]
    agg[row[ki]].append(v / 1e3 if u in ('ns', 'nsecond') else (v * 1e3 if u in ('ms', 'msecond') else v))
tot = sum(sum(v) for v in agg.values())
ls = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 80 python bench.py --steps 2 --warmup 1 --cpu-sample 0",
      "(cold-cache, serialised launches: compare SHARES)", ""]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    ls.append('%-72s n=%3d  avg %10.1f us  share %5.2f%%' % (k[:72], len(v), sum(v) / len(v), 100 * sum(v) / tot))
open(os.path.join(P, 'launches_%s_summary.txt' % RT), 'w').write('\n'.join(ls) + '\n')
if os.path.abspath(launches) != os.path.join(P, 'launches_%s.csv' % RT):
    open(os.path.join(P, 'launches_%s.csv' % RT), 'w').write(open(launches).read())
print('\n'.join(out)); print(reg); print('\n'.join(ls[:6]))
