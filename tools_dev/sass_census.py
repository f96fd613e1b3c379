"""SASS opcode census of the in-tree library: per kernel, how many tcgen05 MMA (UTCHMMA), TMEM load (LDTM), TMA load/store
(UTMALDG/UTMASTG), tcgen05 commit/barrier (UTCBAR), legacy tensor-core (HMMA), warp-shuffle (SHFL), packed-fp32 (FFMA2) and
mbarrier (SYNCS) instructions it contains.  Usage: python tools_dev/sass_census.py [out.md]   (needs cuobjdump; no GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diffusynth_b200", "libdiffusynth_b200.so")
OPS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "HMMA", "LDSM", "SHFL", "FFMA2", "SYNCS", "MUFU"]


def main():
    txt = subprocess.check_output(["cuobjdump", "-sass", LIB], text=True)
    names = subprocess.check_output(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), text=True).splitlines()
    blocks = re.split(r"\n\s*Function : \S+\n", txt)[1:]
    rows = collections.OrderedDict()
    for name, body in zip(names, blocks):
        short = re.sub(r"\(.*", "", name).replace("ds::", "")
        c = rows.setdefault(short, collections.Counter())
        c["_inst"] += len(re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/", body, flags=re.M))
        for op in OPS:
            c[op] += len(re.findall(r"\b%s\b|\b%s\." % (op, op), body))
        c["2CTA"] += len(re.findall(r"UTCHMMA\.2CTA", body))
    out = ["# SASS opcode census of diffusynth_b200/libdiffusynth_b200.so (cuobjdump -sass, sm_100a)", "",
           "`UTCHMMA` = tcgen05.mma kind::f16 (`2CTA` = cta_group::2), `LDTM` = tcgen05.ld, `UTMALDG/UTMASTG` = cp.async.bulk.tensor load/store, "
           "`UTCBAR` = tcgen05.commit, `HMMA` = legacy mma.sync, `LDSM` = ldmatrix, `SHFL` = warp shuffle, `FFMA2` = packed fp32 FMA, `SYNCS` = mbarrier ops.", "",
           "| kernel | SASS instructions | " + " | ".join(OPS) + " | UTCHMMA.2CTA |", "|---|---|" + "---|" * (len(OPS) + 1)]
    for k, c in rows.items():
        out.append(f"| `{k}` | {c['_inst']} | " + " | ".join(str(c[o]) for o in OPS) + f" | {c['2CTA']} |")
    text = "\n".join(out) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
