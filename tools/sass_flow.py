"""Dump an ncu SASS source page as a flow: one line per run of instructions with similar execution count."""
import csv, io, subprocess, sys
rep = sys.argv[1]; name_sub = sys.argv[2] if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
blocks = []
for k, h in enumerate(hi):
    e = hi[k + 1] - 1 if k + 1 < len(hi) else len(rows)
    blocks.append((rows[h - 1] if h > 0 else [], rows[h], rows[h + 1:e]))
blk = next((b for b in blocks if name_sub is None or (b[0] and name_sub in " ".join(b[0]))), blocks[0])
hdr, body = blk[1], [r for r in blk[2] if len(r) == len(blk[1])]
ii = hdr.index("Instructions Executed"); ith = hdr.index("Thread Instructions Executed"); isrc = hdr.index("Source")
run = None
def flush(run):
    if run: print("%5d..%5d n=%4d  exec=%10d  act=%5.1f  first: %s" % (run[0], run[1], run[1] - run[0] + 1, run[2], run[3] / max(run[4], 1), run[5][:70]))
for idx, r in enumerate(body):
    ie = int(r[ii]); it = int(r[ith])
    if run and abs(ie - run[2]) <= 0.02 * max(ie, run[2], 1):
        run[1] = idx; run[3] += it; run[4] += ie
    else:
        flush(run); run = [idx, idx, ie, it, ie, r[isrc].strip()]
flush(run)
