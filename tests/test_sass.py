"""CPU: the built library really contains the Blackwell instructions the design rests on (no GPU needed: cuobjdump reads the
sm_100a SASS out of libnvbio_b200.so).  This is the static half of the evidence; the dynamic half is profiles/*ncu*."""
import os
import re
import shutil
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    from nvbio_b200 import build
    lib = build.build()
    out = subprocess.run([exe, "-sass", lib], capture_output=True, text=True).stdout
    funcs = {}
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)
    assert "sm_100a" in out or "SM100" in out.upper() or funcs, "no sm_100a code in the library"
    return {k: "\n".join(v) for k, v in funcs.items()}


def _find(sass, *needles):
    return [k for k in sass if all(n in k for n in needles)]


def test_banded_pair_kernel_uses_dpx(sass):
    """gotoh_pair_kernel<31,LOCAL>: three VIADDMNMX.S16x2 per cell (93 of them in the unrolled row) + VIMNMX3 + PRMT look-ups + LDS.U16"""
    k = _find(sass, "gotoh_pair_kernel", "Li31ELi1E")
    assert k, "gotoh_pair_kernel<31,1> not found"
    body = sass[k[0]]
    assert body.count("VIADDMNMX.S16x2") >= 90
    assert body.count("VIMNMX3.S16x2") >= 29 and body.count("VIMNMX3.U16x2") >= 15
    assert body.count("PRMT") >= 31 and body.count("LDS.U16") >= 31
    assert "STL" not in body and "LDL" not in body, "the band spilled to local memory"


def test_full_matrix_kernels_use_dpx_and_shuffles(sass):
    for name in ("gotoh_full_pair_kernel", "gotoh_full_warp_kernel"):
        ks = _find(sass, name)
        assert ks, name
        assert all("VIADDMNMX.S16x2" in sass[k] for k in ks), name
    warp = _find(sass, "gotoh_full_warp_kernel")
    assert all("SHFL.UP" in sass[k] and "SHFL.IDX" in sass[k] for k in warp)


def test_fm_kernels_load_blocks_with_one_256bit_instruction(sass):
    """{bwt, occ} block = one LDG.E.ENL2.LTC64B.256 (ld.global.nc.L2::64B.v8.u32)"""
    for name in ("fm_match_kernel", "pipe_seed_match_kernel", "fm_rank_kernel", "fm_locate_kernel"):
        ks = _find(sass, name)
        assert ks, name
        assert all(re.search(r"LDG\.E\.ENL2\.LTC64B\.256", sass[k]) for k in ks), name
