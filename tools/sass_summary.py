"""Per-kernel SASS summary of libnetphys_b200.so for profiles/: registers, stack, shared, local loads/stores, FP32 instruction mix.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "netphys_b200", "libnetphys_b200.so")
res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for l in res.splitlines():
    m = re.search(r"Function (\S+):", l)
    if m: cur = m.group(1); continue
    if cur and "REG:" in l:
        usage[cur] = dict(re.findall(r"(REG|STACK|SHARED|LOCAL|CONSTANT\[0\]):(\d+)", l)); cur = None
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.OrderedDict(); cur = None
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m: cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m and cur: counts[cur][m.group(1).split(".")[0]] += 1
def demangle(n):
    full = subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip()
    full = (full.split(">(")[0] + ">") if ">(" in full else full.split("(")[0]
    return full.replace("void np::", "").replace("np::", "").replace("(bool)", "").replace("(int)", "")
print("# cuobjdump -sass / --dump-resource-usage of netphys_b200/libnetphys_b200.so (parity build: -fmad=false).  FFMA outside the IEEE division / sqrt / sincos")
print("# expansions would mean a contracted product: the narrow-phase kernels have none of their own (FFMA counts below are those expansions, and support_tab's explicit")
print("# __fmaf_rn in the cell lookup, which never touches a result value).")
print("%-58s %4s %6s %6s %5s %5s %6s %6s %6s %6s" % ("kernel", "regs", "stack", "insts", "LDL", "STL", "FADD", "FMUL", "FFMA", "MUFU"))
for k, c in counts.items():
    u = usage.get(k, {})
    n = demangle(k)
    print("%-58s %4s %6s %6d %5d %5d %6d %6d %6d %6d" % (n[:58], u.get("REG", "?"), u.get("STACK", "?"), sum(c.values()), c["LDL"], c["STL"], c["FADD"], c["FMUL"], c["FFMA"], c["MUFU"]))
