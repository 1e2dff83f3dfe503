"""Per-kernel Blackwell evidence from the built library: counts of the SASS mnemonics that only tcgen05 / TMEM / TMA code produces
(B200_PROFILING.md "What proves a Blackwell-native kernel").

    python profiles/sass_summary.py > profiles/sass_summary.md
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "few-shot-music-generation_b200" / "fsmg" / "libfsmg.so"
PATTERNS = collections.OrderedDict([
    ("UTCHMMA (tcgen05.mma kind::f16)", r"\bUTCHMMA"), ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"), ("LDTM (tcgen05.ld)", r"\bLDTM"),
    ("STTM (tcgen05.st)", r"\bSTTM"), ("UTMALDG (TMA tensor load)", r"\bUTMALDG"), ("UTMAPF (TMA prefetch)", r"\bUTMAPF|UTMACCTL"),
    ("SYNCS (mbarrier)", r"\bSYNCS"), ("HMMA (legacy mma.sync)", r"\bHMMA"), ("REDG/RED (global reductions)", r"\bRED\b|\bREDG"),
])


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PATTERNS.items():
            if re.search(pat, line):
                kernels[cur][name] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS evidence per kernel of `libfsmg.so` (sm_100a)\n")
    print("`cuobjdump -sass few-shot-music-generation_b200/fsmg/libfsmg.so`, instruction counts per kernel (static, not dynamic).  "
          "`UTCHMMA` = `tcgen05.mma`, `LDTM`/`STTM` = `tcgen05.ld`/`st` (TMEM), `UTMALDG` = `cp.async.bulk.tensor` (TMA), `SYNCS` = mbarrier; "
          "no `HMMA` (legacy `mma.sync`) anywhere.\n")
    cols = list(PATTERNS)
    print("| kernel | " + " | ".join(c.split(" ")[0] for c in cols) + " |")
    print("|---|" + "---|" * len(cols))
    tot = collections.Counter()
    for (mangled, cnt), name in zip(kernels.items(), demangled):
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("fsmg::", "")
        if not any(cnt.values()):
            continue
        print(f"| `{short}` | " + " | ".join(str(cnt[c]) for c in cols) + " |")
        tot.update(cnt)
    print("| **total** | " + " | ".join(f"**{tot[c]}**" for c in cols) + " |")
    plain = [re.sub(r"\(.*", "", n).replace("void ", "").replace("fsmg::", "") for (k, c), n in zip(kernels.items(), demangled) if not any(c.values())]
    print(f"\n{len(kernels)} kernels in the library; {len(plain)} SIMT-only kernels (element-wise, gather/scatter, reductions, sort, unigram) carry none of the above: "
          + ", ".join(f"`{p}`" for p in plain[:40]) + ("…" if len(plain) > 40 else "") + ".")


if __name__ == "__main__":
    main()
