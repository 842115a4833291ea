"""Properties of the compiled sm_100a code that the measured performance depends on (CPU only: ptxas logs and SASS of the in-tree build).
Each one guards a regression that was hit while developing:
  * the quad kernel calls its scene scan (one copy of the scan instead of three inlined ones: it was instruction-fetch bound) and must fit
    3 CTAs x 128 threads per SM (<= 168 registers; forcing 128 registers / 4 CTAs was measured: fused 0.91 -> 0.95 ms on default1080);
  * the persistent kernel runs one 640-thread CTA per SM (<= 96 registers after allocation granularity) and must not spill inside its loops
    more than it does today;
  * the scene is staged with a TMA bulk copy and the solver uses the packed FFMA2 instruction (DESIGN.md sections 3 and 4)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "raytracing-opengl_b200", "csrc")
LIB = os.path.join(ROOT, "raytracing-opengl_b200", "librtb200.so")


def _ptxas(log, entry):
    """(registers, spill store bytes, spill load bytes) of one kernel in a `-Xptxas -v` log"""
    text = open(os.path.join(CSRC, log)).read()
    m = re.search(r"Compiling entry function '[^']*" + re.escape(entry) + r"[^']*' for 'sm_100a'.*?(\d+) bytes spill stores, (\d+) bytes spill loads.*?Used (\d+) registers",
                  text, re.S)
    assert m, entry
    return int(m.group(3)), int(m.group(1)), int(m.group(2))


@pytest.mark.skipif(not os.path.isfile(os.path.join(CSRC, "ptxas_strict.log")), reason="ptxas logs are written by the build (make -C csrc)")
@pytest.mark.parametrize("log", ["ptxas_strict.log", "ptxas_fast.log"])
def test_register_budgets_of_the_two_kernels(log):
    for counted in ("Lb0", "Lb1"):
        regs, st, ld = _ptxas(log, "quad_kernelI" + counted)
        if counted == "Lb0":                                 # (the counting variant is an untimed instrumentation build)
            assert regs <= 168, f"quad_kernel<{counted}> uses {regs} registers: fewer than 3 CTAs of 128 threads fit an SM"
        assert st <= 128 and ld <= 128, (st, ld)
        regs, st, ld = _ptxas(log, "persistent_kernelI" + counted + "ELi640E")
        assert regs <= 96, f"persistent_kernel<{counted}, 640> uses {regs} registers: a 640-thread CTA no longer fits"
        assert st <= 1024 and ld <= 1024, (st, ld)           # spills exist, outside the scan loops (checked in the SASS when they change)
        regs, st, ld = _ptxas(log, "persistent_kernelI" + counted + "ELi768E")      # the 24-warp variant of scenes without tori
        assert regs <= 80, f"persistent_kernel<{counted}, 768> uses {regs} registers: a 768-thread CTA no longer fits"


@pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.isfile(LIB), reason="needs cuobjdump and the built library")
def test_sass_has_the_tma_bulk_copy_and_the_packed_fp32_instructions():
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", "_ZN10rtb_strict17persistent_kernelILb0ELi640EEEv11FrameParams", LIB],
                          capture_output=True, text=True, timeout=600).stdout
    assert "sm_100a" in sass
    assert sass.count("UBLKCP") >= 1                         # cp.async.bulk of the packed scene (stage_scene_tma)
    assert sass.count("SYNCS.PHASECHK") >= 1                 # the mbarrier wait that publishes it
    assert sass.count("FFMA2") >= 400                        # rotate2, the box / quadric tests, the Durand-Kerner solver
    assert "F32x2.LO_HI" in sass                             # the free half swap of the packed complex product (cmul_pp)
    assert sass.count("FMNMX3") >= 8                         # 3-input min/max of the Durand-Kerner witnesses
    assert "HMMA" not in sass and "UTCHMMA" not in sass      # no tensor-core work on this path (north_star)
