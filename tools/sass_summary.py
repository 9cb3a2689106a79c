"""Count the SASS mnemonics that prove tcgen05 / TMEM / bulk-copy use, per kernel of the built library -> profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "mit_rl-vrptw_b200", "lib", "libvrpb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
cols = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "ELECT", "MUFU", "FFMA2", "FMUL2", "FADD2", "FFMA", "BAR", "UTMALDG", "UTMASTG", "HMMA"]
rows, cur, it = collections.OrderedDict(), None, iter(names)
for line in sass.split("\n"):
    if "Function :" in line:
        n = next(it)
        n = re.sub(r"^void ", "void ", re.sub(r"vrpb200::", "", n.split("(")[0]))
        cur = rows.setdefault(n, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        for c in cols:
            if op == c or (c in ("BAR",) and op.startswith("BAR")) or (c not in ("BAR", "FFMA") and op.startswith(c)):
                cur[c] += 1
        if op == "FFMA":
            pass
out = ["# SASS mnemonics per kernel of mit_rl-vrptw_b200/lib/libvrpb200.so (cuobjdump -sass, CUDA 12.9, sm_100a); tools/sass_summary.py",
       "# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops,",
       "# FFMA2/FMUL2/FADD2 = packed FP32x2; UTMALDG/UTMASTG (tensor-map TMA) are not used: operands are staged with bulk copies and plain stores",
       "kernel | " + " | ".join(cols) + " | total"]
for n in sorted(rows):
    out.append(n + " | " + " | ".join(str(rows[n][c]) for c in cols) + " | " + str(rows[n]["total"]))
open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w").write("\n".join(out) + "\n")
print(len(rows), "kernels")
