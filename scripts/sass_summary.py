"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UBLKCP / UTMALDG / UTMASTG (bulk / tensor TMA), HMMA (legacy mma.sync),
plus the FFMA count of the fp32 engines.  Needs no GPU:

    python scripts/sass_summary.py [lrs_pnp_dip_b200/csrc/liblrs_pnp.so] > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "lrs_pnp_dip_b200", "csrc", "liblrs_pnp.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
PAT = collections.OrderedDict([("UTC*MMA", r"\bUTC[A-Z]*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"), ("UBLKCP", r"\bUBLKCP\b"),
                               ("UTMALDG", r"\bUTMALDG\b"), ("UTMASTG", r"\bUTMASTG\b"), ("HMMA", r"\bHMMA\b"),
                               ("FFMA", r"\bFFMA\b"), ("FFMA2", r"\bFFMA2\b"), ("DFMA", r"\bDFMA\b"), ("LDGSTS", r"\bLDGSTS\b"),
                               ("SYNCS", r"\bSYNCS\b")])
kern, counts, arch = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", kern)
        kern = re.sub(r"\(.*", "", kern)[:90]
        counts[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        cur_arch = m.group(1)
    if kern is None:
        continue
    arch.setdefault(kern, cur_arch if "cur_arch" in dir() else "?")
    for name, pat in PAT.items():
        if re.search(pat, line):
            counts[kern][name] += 1
print(f"# {os.path.relpath(so, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass); arch of every cubin: "
      f"{sorted(set(arch.values()))}")
print(f"{'kernel':92s}" + "".join(f"{n:>9s}" for n in PAT))
for k, c in counts.items():
    print(f"{k:92s}" + "".join(f"{c[n]:9d}" for n in PAT))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"{'TOTAL':92s}" + "".join(f"{tot[n]:9d}" for n in PAT))
