"""profiles/rNN_sass.md: per-kernel counts of the SASS instructions that show which hardware paths are used.
    python tools/sass_census.py nafae_b200/libnafae_b200.so profiles/r02_sass.md"""
import re
import subprocess
import sys

lib, dst = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
pats = ["UBLKCP", "UBLKRED", "SYNCS", "UTCHMMA", "UTMALDG", "LDTM", "LDGMC", "LDGSTS", "ATOMS", r"REDG|ATOMG",
        r"(?<![A-Z])HMMA"]
heads = ["UBLKCP", "UBLKRED", "SYNCS", "UTCHMMA", "UTMALDG", "LDTM", "LDGMC", "LDGSTS", "ATOMS", "REDG/ATOMG", "HMMA"]
names = [f.split("\n", 1)[0].strip() for f in funcs]
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.strip().split("\n")
rows = []
for f, d in zip(funcs, dem):
    ins = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4,6}\*/\s", l)]
    cnt = [sum(1 for l in ins if re.search(p, l)) for p in pats]
    d = d.replace("(anonymous namespace)::", "").replace("nafae::", "").replace("void ", "")
    d = re.sub(r"\((?!anonymous).*$", "", d)
    rows.append((d, len(ins), cnt))
out = ["# SASS instruction census of libnafae_b200.so (cuobjdump -sass, sm_100a)", "",
       "Which hardware paths each kernel uses. `UBLKCP` / `UBLKRED` = 1-D bulk async copy / bulk reduce-add on the TMA "
       "engine (`cp.async.bulk`, `cp.reduce.async.bulk`); `UTMALDG` = tensor-map TMA load; `SYNCS` = mbarrier operations; "
       "`UTCHMMA` = `tcgen05.mma` (kind::f16 / kind::tf32); `LDTM` = `tcgen05.ld` (TMEM -> registers); `LDGMC` = "
       "`multimem.ld_reduce` (NVLS in-switch reduction; `multimem.st` is an ordinary `STG.E.128.STRONG.SYS` to the "
       "multicast address); `LDGSTS` = `cp.async`; `ATOMS` = shared-memory atomics; `REDG`/`ATOMG` = global reductions / "
       "atomics; `HMMA` = legacy mma.sync (none).  Counts are static instructions (unrolled copies included).", "",
       "| kernel | instructions | " + " | ".join(heads) + " |", "|---|---:|" + "---:|" * len(pats)]
for d, n, cnt in sorted(rows, key=lambda r: -r[1]):
    if n < 100 and not any(cnt):
        continue
    out.append("| `%s` | %d | " % (d[:80], n) + " | ".join(str(c) if c else "" for c in cnt) + " |")
tot = [sum(r[2][i] for r in rows) for i in range(len(pats))]
out.append("| **whole library** | %d | " % sum(r[1] for r in rows) + " | ".join(str(c) for c in tot) + " |")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out[-8:]))
