"""Like hot_lines.py, but aggregates an ncu SASS source page by CODE REGION (named line ranges), walking each instruction's inline chain
outwards until a frame falls in a region.

    python tools/hot_regions.py <report.ncu-rep> <mangled substring> "<ncu kernel name substring>" file:lo-hi=name ...
"""
import csv, io, re, subprocess, sys, os, glob, tempfile
rep, sub, name_sub = sys.argv[1], sys.argv[2], sys.argv[3]
regions = []
for a in sys.argv[4:]:
    loc, nm = a.split("="); f, rng = loc.split(":"); lo, hi = rng.split("-"); regions.append((f, int(lo), int(hi), nm))
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "netphys_b200", "libnetphys_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith("//---") and ".text." in l and sub in l)
end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith("//---") and ".text." in dis[i]), len(dis))
insts = []; cur = None; in_group = False
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:                                                    # a group of //## lines: innermost frame first, one enclosing frame per line
        if not in_group: cur = [(os.path.basename(m.group(1)), int(m.group(2)))]; in_group = True
        inl = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        if inl: cur.append((os.path.basename(inl.group(1)), int(inl.group(2))))
        continue
    in_group = False
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: insts.append((cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
blocks = []
for k, h in enumerate(hi):
    e = hi[k + 1] - 1 if k + 1 < len(hi) else len(rows)
    blocks.append((rows[h - 1] if h > 0 else [], rows[h], rows[h + 1:e]))
blk = next(b for b in blocks if b[0] and name_sub in " ".join(b[0]))
hdr, body = blk[1], [r for r in blk[2] if len(r) == len(blk[1])]
assert len(body) == len(insts), (len(body), len(insts))
ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); ith = hdr.index("Thread Instructions Executed")
agg = {}; tot = tott = tots = 0
for k in range(len(body)):
    r = body[k]; chain = insts[k][0] or []
    ie = int(r[ii]); s = int(r[isamp]); t = int(r[ith]); tot += ie; tott += t; tots += s
    nm = "other"
    for (f, l) in chain:
        hit = next((g[3] for g in regions if g[0] == f and g[1] <= l <= g[2]), None)
        if hit: nm = hit; break
    a = agg.setdefault(nm, [0, 0, 0]); a[0] += ie; a[1] += t; a[2] += s
print("warp-inst %d thread-inst %d avg active %.1f" % (tot, tott, tott / max(tot, 1)))
for nm, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%6.1f%% warp-inst %6.1f%% thread-inst %5.1f%% samples  act=%5.1f  %s" % (100 * a[0] / tot, 100 * a[1] / tott, 100 * a[2] / max(tots, 1), a[1] / max(a[0], 1), nm))
