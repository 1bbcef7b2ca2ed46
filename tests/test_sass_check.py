"""The build's SASS check of the episode kernel's Tensor-Memory loads (learning_drones_b200/build.py,
verify_tmem_scoreboards): csrc/ldrones.cu streams the brain with tcgen05.ld and no tcgen05.wait::ld between the loads, which
is only correct because ptxas tracks every load on a hardware scoreboard and makes the first consumer wait for it.  The
check must pass on the shipped libraries and must catch a consumer that does not wait.  CPU only (cuobjdump)."""
import shutil
import subprocess

import pytest

from learning_drones_b200 import build as B

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None and not B.os.path.exists("/usr/local/cuda/bin/cuobjdump"),
                                reason="cuobjdump not available")


def _functions(lib, pattern):
    cuobjdump = B.os.path.join(B.os.path.dirname(B.nvcc_path()), "cuobjdump")
    listing = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, check=True).stdout
    for block in listing.split("Function : ")[1:]:
        if pattern in block.split("\n", 1)[0]:
            yield B.parse_sass(block)


def test_shipped_libraries_pass():
    B.build()
    for lib in (B.LIB, B.LIB_CHECKED):
        assert B.verify_tmem_scoreboards(lib) == []
    assert B.verify_tmem_scoreboards(B.LIB, "fly_mlp_f32_kernel") == []


def test_unknown_kernel_is_an_error():
    assert B.verify_tmem_scoreboards(B.LIB, "no_such_kernel")


def test_a_consumer_that_does_not_wait_is_caught():
    ins = next(_functions(B.LIB, "fly_mlp_f64_kernelILb0ELi0ELb0ELi4E"))
    assert B.check_sass_tmem_loads(ins) == []
    loads = [k for k, (_, t, w) in enumerate(ins) if B._sass_opcode(t) == "LDTM" and ((w >> 46) & 7) != 7]
    assert len(loads) >= 32  # the hot loop's x16 stream + the staging code
    broken = 0
    for k in loads[::5]:
        sb = (ins[k][2] >> 46) & 7
        # the first later instruction that waits on this load's scoreboard: strip that wait
        j = next(i for i in range(k + 1, len(ins)) if (ins[i][2] >> 52) & 0x3F & (1 << sb))
        tampered = list(ins)
        a, t, w = ins[j]
        tampered[j] = (a, t, w & ~((1 << sb) << 52))
        if B._sass_registers(t) & B._sass_registers(ins[k][1]):  # j is a real consumer of the load
            assert any(f"{ins[k][0]:#x}" in b for b in B.check_sass_tmem_loads(tampered)), (ins[k], ins[j])
            broken += 1
    assert broken >= 3


def test_register_operand_widths():
    assert B._sass_registers("DFMA R16, R158.reuse, R52, R16") == {16, 17, 158, 159, 52, 53}
    assert B._sass_registers("LDTM.x16 R20, tmem[UR9+0x10]") == set(range(20, 36))
    assert B._sass_registers("LDS.128 R36, [R85+0x20]") >= {36, 37, 38, 39, 85}
    assert B._sass_registers("@P6 DADD R12, -R156, R174") == {12, 13, 156, 157, 174, 175}
    assert B._sass_registers("R2UR UR9, R126") == {126}
    assert B._sass_registers("FSEL R176, R26, R176, P5") == {176, 26}
