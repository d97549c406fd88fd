"""Opcode evidence from the shipped library: per kernel, how many tcgen05 / TMA / mma.sync instructions its SASS holds.

    python tools/sass_summary.py [lib.so] > profiles/r2_sass_summary.txt

UTCHMMA = tcgen05.mma (kind::f16), UTCHMMA.2CTA = cta_group::2, LDTM = tcgen05.ld, UTMALDG / UTMASTG / UTMAREDG = TMA load /
store / reduce, UTCBAR = tcgen05.commit, HMMA.16816 = mma.sync (B200_PROFILING.md lists the mnemonics)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "focused-attention-vit_b200", "libfavit_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTCQMMA\S*|LDTM\S*|STTM\S*|UTMALDG\S*|UTMASTG\S*|UTMAREDG\S*|UTMAPF\S*|UTCBAR\S*|"
                 r"UTCATOM\S*|HMMA\.\d+\S*|SYNCS\.\S+|LDGSTS\S*|MUFU\.\S+|RED\.\S+|ATOM\S*|REDG\S*)")
kern, counts, lines = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "")
        kern = re.sub(r"^void ", "", re.sub(r"\(.*", "", kern))
        while kern in counts:          # template instantiations demangle to the same prefix only if cut too early
            kern += "'"
        counts[kern] = collections.Counter()
        continue
    if kern and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        lines[kern] += 1
        for op in pat.findall(line):
            op = re.sub(r"^(SYNCS)\.(\w+).*", r"\1.\2", op)
            op = re.sub(r"^(MUFU\.\w+).*", r"\1", op)
            op = re.sub(r"^(UTMALDG\.\dD)(\.\w+)*", lambda mm: mm.group(0), op)
            counts[kern][op] += 1
total = collections.Counter()
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}: {len(counts)} kernels, {sum(lines.values())} SASS lines")
print("# tensor-core / TMA / barrier opcodes per kernel (kernels without any are listed by name only at the end)\n")
plain = []
for k, c in counts.items():
    keys = [o for o in c if o.startswith(("UTC", "LDTM", "STTM", "UTMA", "HMMA"))]
    if not keys:
        plain.append(k)
        continue
    print(f"{k}   [{lines[k]} lines]")
    print("    " + "  ".join(f"{o} x{c[o]}" for o in sorted(c)))
    total.update({o: c[o] for o in keys})
print("\n# totals over the library")
print("    " + "  ".join(f"{o} x{n}" for o, n in sorted(total.items())))
print(f"\n# {len(plain)} kernels without tensor-core / TMA opcodes (SIMT / plain ld-st):")
for k in plain:
    print("    " + k)
