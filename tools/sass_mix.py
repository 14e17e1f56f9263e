"""Static instruction mix of a sweep kernel's steady-state loop, from the object files of the in-tree build
(no GPU needed): the numbers behind DESIGN.md section 9 ("~N non-FP64 instructions per pair").

    python tools/sass_mix.py [d] [kernel-substring]        # default: d = 11, dmma_sweep_kernelILi0E (Matern32)

The steady-state region is found structurally: the longest branch-free run of instructions that contains DMMAs
(the fully unrolled n-tiles of one half tile).  Pairs per lane in that region = DMMA chains / KS * 2 / ... is
reported as (number of MUFU.RSQ64H) for Matern32, one reciprocal square root per evaluated pair."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    cur, body = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur:
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
            if m:
                body[cur].append(m.group(1))
    return body


def steady_region(ops):
    best, start = (0, 0, 0), 0
    for i, op in enumerate(ops + ["BRA"]):
        if op.split(".")[0] in ("BRA", "EXIT", "BSYNC", "BSSY", "RET", "CALL"):
            n_mma = sum(1 for o in ops[start:i] if o.startswith("DMMA"))
            if n_mma and i - start > best[0]:
                best = (i - start, start, i)
            start = i + 1
    return best[1], best[2]


def main():
    d = int(sys.argv[1]) if len(sys.argv) > 1 else 11
    pat = sys.argv[2] if len(sys.argv) > 2 else "dmma_sweep_kernelILi0E"
    obj = os.path.join(ROOT, "cglb_b200", "csrc", "build", f"kmv_d{d}.o")
    for name, ops in functions(obj).items():
        if pat not in name:
            continue
        a, b = steady_region(ops)
        c = collections.Counter(o.split(".")[0] for o in ops[a:b])
        full = collections.Counter(ops[a:b])
        fp64 = c["DFMA"] + c["DMUL"] + c["DADD"]
        pairs = full["MUFU.RSQ64H"] or None
        other = (b - a) - fp64 - c["DMMA"]
        print(f"{name}\n  steady region: {b - a} instructions; DFMA {c['DFMA']} DMUL {c['DMUL']} DADD {c['DADD']} "
              f"(FP64 {fp64}), DMMA {c['DMMA']}, other {other}")
        if pairs:
            print(f"  per evaluated pair ({pairs} pairs per lane in the region): FP64 {fp64 / pairs:.2f}, DMMA {c['DMMA'] / pairs:.2f} "
                  f"(= {8 * c['DMMA'] / pairs:.1f} DFMA-equivalent pipe slots), other {other / pairs:.2f}")
        print("  other: " + ", ".join(f"{k} {v}" for k, v in sorted(full.items(), key=lambda kv: -kv[1])
                                      if k.split('.')[0] not in ("DFMA", "DMUL", "DADD", "DMMA")))


if __name__ == "__main__":
    main()
