"""Static SASS evidence for profiles/: per kernel of libevk.so, the counts of the memory instructions that carry the design
(global reductions, shared-memory atomics, TMA bulk operations, evict-first 16-byte loads, multicast load-reduce).

    python tools/sass_summary.py > profiles/sass_r2.md        # needs cuobjdump + c++filt, no GPU
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "event_utils_b200", "libevk.so")
KEEP = re.compile(r"^(REDG|RED\.|ATOMG|ATOM\.|ATOMS|UBLKRED|UBLKCP|UTMA|LDGMC|LDG\.E(\.EF)?\.128|LDG\.E\.EF|STG\.E\.128|MATCH|BAR\.RED|NANOSLEEP|LDS\.128|STS\.128|SYNCS|UCGABAR)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            if KEEP.match(op):
                cur[op] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS evidence, round 2 (`cuobjdump -sass event_utils_b200/libevk.so`, sm_100a; `tools/sass_summary.py`)\n")
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print("Whole library: " + ", ".join("%s x%d" % kv for kv in sorted(tot.items()) if kv[0].startswith(("UBLK", "UTMA", "ATOMS", "LDGMC"))) + "\n")
    print("What to look for: `UBLKRED` = `cp.reduce.async.bulk.global.shared::cta.add.f32` (TMA bulk add-reduction of a finished "
          "shared-memory tile: the contrast-maximisation image of `cmax_onchip_kernel`, the tiles of `voxel_routed_kernel`); "
          "`ATOMS.ADD` = native integer shared-memory atomics (fixed-point accumulators; the f32 form would be an "
          "`ATOMS.CAST.SPIN` CAS loop); `REDG.E.ADD.F32x4` = the 16-byte vector reduction to L2; `LDG.E.EF.128` = evict-first "
          "16-byte event loads; `LDGMC` = NVLS multicast load-reduce of the multi-GPU fold.\n")
    for mangled, name in zip(kernels, names):
        c = kernels[mangled]
        if not c:
            continue
        short = re.sub(r"\(.*", "", name).replace("void ", "")
        print("* `%s`: %s" % (short, ", ".join("%s ×%d" % kv for kv in sorted(c.items()))))


if __name__ == "__main__":
    sys.exit(main())
