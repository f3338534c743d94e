#!/usr/bin/env python
"""Static SASS instruction counts per object (tcgen05 / TMEM / TMA evidence): python scripts/sass_summary.py > table.md
Reads moleculardiffusion_mivit_b200/build/*.o with cuobjdump (run build() first)."""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pats = [("UTCHMMA", r"\bUTCHMMA"), ("of which .2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("UTCBAR", r"\bUTCBAR"),
        ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("LDGSTS", r"\bLDGSTS"),
        ("warp-level HMMA (tf32 `mma.sync`)", r"\bHMMA"), ("IMMA / HGMMA", r"\b(IMMA|HGMMA)")]
print("| object | " + " | ".join(p[0] for p in pats) + " |")
print("|---|" + "---:|" * len(pats))
for obj in sorted(glob.glob(os.path.join(ROOT, "moleculardiffusion_mivit_b200", "build", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    counts = [len(re.findall(rx, sass)) for _, rx in pats]
    print("| `%s.cu` | " % os.path.basename(obj)[:-2] + " | ".join(str(c) for c in counts) + " |")
