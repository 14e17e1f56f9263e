"""CPU: the instruction budget DESIGN.md quotes for the headline sweeps, read back from the compiled objects
(cuobjdump -sass of cglb_b200/csrc/build/kmv_d11.o through tools/sass_mix.py; no GPU).  Guards the claims
"27 FP64-pipe slots per evaluated pair" (forward, Matern32, d = 11) and "47" (backward) against source or compiler
drift, and checks that the hot loops contain what they are said to contain: DMMA, the MUFU.RSQ64H seed, no libm
calls, no local-memory spills."""
import importlib.util
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "cglb_b200", "csrc", "build", "kmv_d11.o")

pytestmark = pytest.mark.skipif(not (os.path.exists(OBJ) and shutil.which("cuobjdump")),
                                reason="needs the in-tree object files (python -m cglb_b200.build) and cuobjdump")


@pytest.fixture(scope="module")
def mix():
    spec = importlib.util.spec_from_file_location("sass_mix", os.path.join(ROOT, "tools", "sass_mix.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, mod.functions(OBJ)


def _region(mod, ops):
    a, b = mod.steady_region(ops)
    return ops[a:b]


def _count(region, *prefixes):
    return sum(1 for o in region if o.split(".")[0] in prefixes)


def test_forward_dmma_sweep_budget(mix):
    mod, fns = mix
    (name, ops), = [(k, v) for k, v in fns.items() if "dmma_sweep_kernelILi0E" in k]
    reg = _region(mod, ops)
    pairs = sum(1 for o in reg if o == "MUFU.RSQ64H")
    assert pairs >= 16
    fp64 = _count(reg, "DFMA", "DMUL", "DADD")
    dmma = _count(reg, "DMMA")
    assert fp64 == 15 * pairs                     # 5 (sqrt) + 7 (exp) + 1 ((1+s) e) + 2 (y_i, y_j)
    assert 2 * dmma == 3 * pairs                  # K = 12 = d + 1: three k-steps per 8 x 8 tile, two pairs per lane
    assert fp64 + 8 * dmma == 27 * pairs          # a DMMA.8x8x4 holds the pipe as long as 8 warp-wide DFMAs
    other = len(reg) - fp64 - dmma
    assert other <= 15 * pairs                    # DESIGN.md section 9 / profiles/dsweep_sass_mix_r01.md: 14.25
    assert not any(o.startswith("CALL") for o in reg)                      # no libm in the loop (the item cursor's sqrt is
    assert not any(o.startswith(("LDL", "STL")) for o in ops)              # the only call, once per work item); no spills


def test_backward_dmma_sweep_budget(mix):
    mod, fns = mix
    (name, ops), = [(k, v) for k, v in fns.items() if "dmma_bwd_kernelILi0E" in k]
    reg = _region(mod, ops)
    pairs = sum(1 for o in reg if o == "MUFU.RSQ64H")
    assert pairs >= 16
    fp64 = _count(reg, "DFMA", "DMUL", "DADD")
    dmma = _count(reg, "DMMA")
    assert 2 * dmma == 7 * pairs                  # 3 distance + 2 x 2 cross-term DMMAs per 8 x 8 tile
    assert fp64 + 8 * dmma <= 47 * pairs
    assert not any(o.startswith("CALL") for o in reg)
    # no spill traffic in the tile loop; the kernel sits at the 255-register limit and parks ONE word per work item on the
    # stack (stored when an item starts, reloaded for the item-end flush of the per-CTA sums)
    assert not any(o.startswith(("LDL", "STL")) for o in reg)
    assert sum(1 for o in ops if o.startswith(("LDL", "STL"))) <= 2


def test_tma_ring_is_in_the_binary(mix):
    _, fns = mix
    (name, ops), = [(k, v) for k, v in fns.items() if "dmma_sweep_kernelILi0E" in k]
    assert any(o.startswith("UBLKCP") for o in ops)                        # cp.async.bulk
    assert any(o.startswith("SYNCS") for o in ops)                         # mbarrier


def test_register_budget_leaves_one_cta_per_sm_without_spills():
    out = subprocess.run(["cuobjdump", "-res-usage", OBJ], capture_output=True, text=True, check=True).stdout
    lines = out.splitlines()
    seen = 0
    for i, l in enumerate(lines):
        if "Function" in l and ("dmma_sweep_kernel" in l or "dmma_bwd_kernel" in l):
            usage = lines[i + 1]
            regs = int(usage.split("REG:")[1].split()[0])
            stack = int(usage.split("STACK:")[1].split()[0])
            assert regs <= 255 and stack <= (8 if "dmma_bwd_kernel" in l else 0) and "LOCAL:0" in usage, usage
            seen += 1
    assert seen == 4
