"""Sum the per-line output of ncu_lines.py over code regions of gstar.cuh / other files.  usage: ncu_regions.py <lines.txt>"""
import re, collections, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = open(sys.argv[1]).read()
GS = sys.argv[2] if len(sys.argv) > 2 else 'gstrip.cuh'          # the spatial-index header the build used (gstrip.cuh, or gstar.cuh with -DMVOSR_UNIFORM_GRID)
lines = open(os.path.join(root, 'mvoscalerecovery_b200/csrc', GS)).read().splitlines()
def find(*pats):
    for pat in pats:
        for i,l in enumerate(lines):
            if pat in l: return i+1
    raise KeyError(pats)
marks = {
 'consume': (find('struct FrameView'), find('// exact conflict of candidate')),
 'group path+fallback': (find('// exact conflict of candidate'), find('// wrap path: one warp per star')),
 'w_eval/key/circle': (find('struct WEval'), find('// One batch of candidates (one per lane)')),
 'w_batch': (find('// One batch of candidates (one per lane)'), find('// Stream the grid cells that the left cap', '// Stream the strips')),
 'w_stream': (find('// Stream the grid cells that the left cap', '// Stream the strips'), find('// The stars of list[0..n_list) (sorted positions), one per warp')),
 'wrap:fetch+load': (find('// The stars of list[0..n_list) (sorted positions), one per warp'), find('// ---- one step of the walk')),
 'wrap:step': (find('// ---- one step of the walk'), find('// ---- the star: lane i keeps')),
 'wrap:seeded walk': (find('// ---- the star: lane i keeps'), find('// ---- q0 = the nearest point')),
 'wrap:q0': (find('// ---- q0 = the nearest point'), find('// ---- the plain walk')),
 'wrap:walk': (find('// ---- the plain walk'), find('// ---- counter-clockwise slot order')),
 'wrap:finish': (find('// ---- counter-clockwise slot order'), find('// pair path: two stars per warp in lock step')),
 'pair': (find('// pair path: two stars per warp in lock step'), find('// All stars of the staged point set.')),
 'run_stars': (find('// All stars of the staged point set.'), 10**6),
}
agg = collections.defaultdict(lambda:[0.0,0.0])
for l in out.splitlines():
    m = re.match(r"(\S+):(\d+)\s+samples\s+([\d.]+)%\s+winstr\s+([\d.]+)%", l)
    if not m: continue
    f, ln, s, w = m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4))
    key = f
    if f == GS:
        key = 'gstar:other'
        for k,(a,b) in marks.items():
            if a <= ln < b: key = k
    agg[key][0] += s; agg[key][1] += w
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%-26s samples %6.2f%%  winstr %6.2f%%" % (k, v[0], v[1]))
