"""Attribute an ncu SASS source page (--page source --csv) to CUDA source lines via nvdisasm -g line info.

    python tools/hot_lines.py <report.ncu-rep> <kernel mangled-name substring> [top]
"""
import csv, io, re, subprocess, sys, os, glob, tempfile

rep, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "netphys_b200", "libnetphys_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate function
start = next(i for i, l in enumerate(dis) if l.startswith("//---") and ".text." in l and sub in l)
end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith("//---") and ".text." in dis[i]), len(dis))
insts = []; cur = None
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        cur = (os.path.basename(m.group(1)), int(m.group(2)), tuple((os.path.basename(a), int(b)) for a, b in inl))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        insts.append((cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# pick the kernel block matching (first kernel in report by default)
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
blocks = []
for k, h in enumerate(hi):
    e = hi[k + 1] - 1 if k + 1 < len(hi) else len(rows)
    blocks.append((rows[h - 1] if h > 0 else [], rows[h], rows[h + 1:e]))
name_sub = sys.argv[4] if len(sys.argv) > 4 else None
blk = next((b for b in blocks if name_sub is None or (b[0] and name_sub in " ".join(b[0]))), blocks[0])
hdr, body = blk[1], [r for r in blk[2] if len(r) == len(blk[1])]
print("kernel:", " ".join(blk[0])[:120], "| sass rows", len(body), "| nvdisasm insts", len(insts))
ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); ith = hdr.index("Thread Instructions Executed")
agg = {}; tot = 0; tots = 0; tott = 0
n = min(len(body), len(insts))
for k in range(n):
    r = body[k]; li = insts[k][0]
    ie = int(r[ii]); s = int(r[isamp]); t = int(r[ith]); tot += ie; tots += s; tott += t
    key = li[:2] if li else ("?", 0)
    # attribute to outermost non-math frame too
    a = agg.setdefault(key, [0, 0, 0]); a[0] += ie; a[1] += s; a[2] += t
print("warp-inst %d  thread-inst %d  (avg active %.1f)  samples %d" % (tot, tott, tott / max(tot, 1), tots))
src_cache = {}
def src(f, l):
    p = os.path.join(os.path.dirname(lib), "csrc", f)
    if f not in src_cache:
        try: src_cache[f] = open(p).read().splitlines()
        except Exception: src_cache[f] = []
    L = src_cache[f]
    return L[l - 1].strip()[:95] if 0 < l <= len(L) else ""
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% samp act=%4.1f | %s:%d  %s" % (100 * a[0] / tot, 100 * a[1] / max(tots, 1), a[2] / max(a[0], 1), key[0], key[1], src(*key)))
