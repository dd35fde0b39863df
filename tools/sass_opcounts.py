"""SASS mnemonic counts of the library's kernels: python tools/sass_opcounts.py > profiles/rNN_sass_opcounts.txt
(cuobjdump -sass of the in-tree .so, names demangled with cu++filt; only kernels that hold one of the listed mnemonics are printed)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "saigegds_b200", "libsaigegds_b200.so")
KEYS = ["UTCIMMA", "UTCIMMA.2CTA", "UTCBAR", "UTCBAR.2CTA.MULTICAST", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMALDG.2CTA", "UBLKCP",
        "IMMA", "LDSM", "SYNCS", "USETMAXREG", "LDGSTS", "LDG.SYS", "STG.SYS", "RED", "DADD", "DFMA", "DMUL", "MUFU", "LOP3", "SHF"]


def classify(op):
    base = op.split(".")[0]
    if base in ("UTCIMMA", "UTCBAR", "UTMALDG"):
        if ".2CTA" in op:
            return base + (".2CTA.MULTICAST" if "MULTICAST" in op else ".2CTA")
        return base
    if base in ("LDG", "LD") and ".SYS" in op:
        return "LDG.SYS"
    if base in ("STG", "ST") and ".SYS" in op:
        return "STG.SYS"
    if base in ("RED", "REDG"):
        return "RED"
    return base


def kernel_counts():
    """{demangled kernel name: Counter of classified mnemonics}, total instruction counts"""
    counts, total, names, dem = _scan()
    return {d: counts[n] for n, d in zip(names, dem)}, {d: total[n] for n, d in zip(names, dem)}


def _scan():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    counts, total, name = {}, {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            counts[name] = collections.Counter()
            total[name] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            total[name] += 1
            counts[name][classify(m.group(1))] += 1
    names = list(counts)
    dem = subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return counts, total, names, dem


def main():
    counts, total, names, dem = _scan()
    print("SASS op counts (cuobjdump -sass saigegds_b200/libsaigegds_b200.so, nvcc 12.9, -gencode arch=compute_100a,code=sm_100a); tools/sass_opcounts.py")
    print("tcgen05.mma -> UTCIMMA, tcgen05.commit -> UTCBAR, tcgen05.alloc -> UTCATOMSWS, tcgen05.ld/st -> LDTM/STTM, cp.async.bulk.tensor -> UTMALDG, "
          "cp.async.bulk -> UBLKCP,\nmma.sync u8/s8 -> IMMA, ldmatrix -> LDSM, mbarrier -> SYNCS, setmaxnreg -> USETMAXREG, cp.async -> LDGSTS, "
          "ld/st at system scope (volatile flags, peer memory) -> LDG.SYS / STG.SYS\n")
    rows = []
    for n, d in zip(names, dem):
        depth, cut = 0, len(d)
        for k in range(len(d) - 1, -1, -1):            # drop the trailing parameter list (template arguments may hold parentheses)
            if d[k] == ")":
                depth += 1
            elif d[k] == "(":
                depth -= 1
                if depth == 0:
                    cut = k
                    break
        short = d[:cut].replace("void ", "").replace("sgb::", "").replace("(anonymous namespace)::", "")
        short = re.sub(r"<unnamed>::", "", short)
        parts = [f"{k} {counts[n][k]}" for k in KEYS if counts[n][k]]
        rows.append(f"{short}: {total[n]} instructions; " + ", ".join(parts))
    for r in sorted(rows):
        print(r)


if __name__ == "__main__":
    sys.exit(main())
