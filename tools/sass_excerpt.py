"""profiles/r02_sass.txt: which kernels of libselfrecon_b200.so carry tcgen05 / TMEM / TMA-engine instructions
(cuobjdump -sass), as counts per kernel plus the first occurrence of each mnemonic with its SASS line."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "selfreconcode_b200", "lib", "libselfrecon_b200.so")
PAT = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTCQMMA|UTCBAR(?:\.[A-Z0-9.]+)?|LDTM(?:\.[a-zA-Z0-9.]+)?|STTM|UBLKCP(?:\.[A-Z.]+)?|"
                 r"UTMALDG|UTCATOMSWS(?:\.[A-Z_.0-9]+)?|SYNCS(?:\.[A-Z_.0-9]+)?)")


def main():
    out = subprocess.check_output(["cuobjdump", "-sass", LIB]).decode(errors="replace")
    counts = collections.OrderedDict()
    first = {}
    fn = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.check_output(["c++filt", m.group(1)]).decode().strip()
            fn = re.sub(r"\(anonymous namespace\)::", "", fn)
            counts[fn] = collections.Counter()
            continue
        m = PAT.search(line)
        if m and fn:
            key = m.group(1)
            counts[fn][key] += 1
            first.setdefault(key, (fn, re.sub(r"\s+", " ", line.strip())))
    lines = ["# cuobjdump -sass selfreconcode_b200/lib/libselfrecon_b200.so (sm_100a), tcgen05 / TMEM / TMA-engine mnemonics",
             "# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA engine),",
             "# UTCBAR = tcgen05.commit -> mbarrier, UTCATOMSWS = tcgen05.alloc / dealloc, SYNCS = mbarrier ops", ""]
    for fn, c in counts.items():
        if any(k.startswith(("UTC", "LDTM", "UBLKCP", "STTM", "UTMA")) for k in c):
            lines.append("%-110s %s" % (fn[:110], "  ".join("%s x%d" % kv for kv in sorted(c.items()))))
    lines += ["", "# first occurrence of each mnemonic:"]
    for k, (fn, l) in sorted(first.items()):
        lines.append("%-28s %s\n%28s %s" % (k, fn[:100], "", l[:150]))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", "r02_sass.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    main()
