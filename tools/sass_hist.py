"""Per-kernel SASS opcode histogram of libvti.so (static counts) + the Blackwell-specific mnemonics found:
    python tools/sass_hist.py > profiles/r2/sass_opcode_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "vision_textile_inspection_b200", "libvti.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::", "", name).split("(")[0].replace("void ", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0] if not m.group(1).startswith(("UTC", "LDTM", "UTMA", "UBLKCP", "SYNCS", "IDP")) else m.group(1)] += 1
special = ("UTC", "LDTM", "STTM", "UTMA", "UBLKCP", "SYNCS", "REDUX", "IDP")
print("# SASS of", os.path.relpath(lib, ROOT), "(cuobjdump -sass, static instruction counts per kernel)\n")
for k, c in hist.items():
    tot = sum(c.values())
    sp = {o: n for o, n in c.items() if o.startswith(special)}
    print(f"{k}: {tot} instructions")
    print("   top: " + ", ".join(f"{o} {n}" for o, n in c.most_common(14)))
    if sp:
        print("   blackwell / async / dot-product: " + ", ".join(f"{o} {n}" for o, n in sorted(sp.items())))
