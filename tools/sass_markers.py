"""Counts the Blackwell-only SASS mnemonics per kernel of libthunder_b200.so (cuobjdump -sass): UTCHMMA (tcgen05.mma),
UTMALDG / UTMASTG (TMA load / store), LDTM / STTM (TMEM <-> registers), UTCBAR (tcgen05.commit), SYNCS (mbarrier).

    python tools/sass_markers.py > profiles/r02_sass_markers.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "thunder_speech_b200", "libthunder_b200.so")
MARK = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "UBLKCP", "HMMA", "FFMA", "LDGSTS"]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
cur, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k in MARK:
        if re.search(r"\b" + k + r"\b|\b" + k + r"\.", line):
            counts[cur][k] += 1
print(f"# {os.path.basename(so)}: SASS markers per kernel (sm_100a)")
print(f"{'kernel':70s} " + " ".join(f"{k:>8s}" for k in MARK))
for k, c in counts.items():
    print(f"{k[:70]:70s} " + " ".join(f"{c[m]:8d}" for m in MARK))
