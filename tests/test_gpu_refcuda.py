"""-m gpu: the B200 kernels against the reference's OWN CUDA code recompiled for sm_100a (oracle/_ref/ref_cuda_bench, built from
/root/reference by oracle/Makefile; the prebuilt binary travels with the snapshot):
  * nvb_fm_match_approx == nvBowtie's device function detail::map<true> (nvBowtie/bowtie2/cuda/mapping_inl.h:128-220), which has no
    host build: every pushed SA range in push order, the push counts and the range sums, incl. seeds with N's;
  * nvb_banded_gotoh_score == batched_banded_alignment_score_kernel, nvb_fm_match / filter == FMIndexFilterDevice::rank / locate."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import nvbio_b200 as nb
from nvbio_b200 import synth
from tests.gpu_util import require_gpu

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tool():
    require_gpu()
    spec = importlib.util.spec_from_file_location("compare_ref_cuda", os.path.join(ROOT, "tools", "compare_ref_cuda.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    if not os.path.exists(m.BIN):
        pytest.skip("oracle/_ref/ref_cuda_bench not built (needs /root/reference at build time)")
    return m


@pytest.fixture(scope="module")
def index2m():
    require_gpu()
    n = 2_000_000
    gw = synth.random_genome_words(n, seed=4242)
    fmi, _ = nb.FMIndexDevice.from_text(gw, n)
    return n, gw, fmi


@pytest.mark.parametrize("L,len1", [(22, 11), (22, 0), (20, 10), (22, 22), (16, 5)])
def test_match_approx_equals_nvbowtie_map(tool, index2m, L, len1):
    n, gw, fmi = index2m
    r = tool.approx_check(fmi, gw, n, nq=20000, L=L, len1=len1, max_out=80, with_n=True, seed=100 + L + len1)
    assert "error" not in r, r
    assert r["bit_identical_counts"] and r["bit_identical_range_sums"] and r["bit_identical_ranges_in_push_order"], r
    assert r["pushes"] > 20000 * (0.3 if len1 < L else 0.2), r
    assert r["max_pushes_per_seed"] <= 80, r


def test_match_approx_equals_nvbowtie_map_repetitive(tool):
    """a genome with long repeats: wide ranges, many substitutions survive"""
    require_gpu()
    n = 600_000
    rng = np.random.default_rng(5)
    unit = rng.integers(0, 4, 3000)
    sym = np.concatenate([unit if (i % 3) else rng.integers(0, 4, 3000) for i in range(n // 3000)]).astype(np.uint8)
    from nvbio_b200.strings import pack_symbols
    gw = torch.from_numpy(pack_symbols(sym, 2, True).view(np.int32)).cuda()
    fmi, _ = nb.FMIndexDevice.from_text(gw, n)
    r = tool.approx_check(fmi, gw, n, nq=10000, L=14, len1=7, max_out=96, with_n=True, seed=9)
    assert "error" not in r, r
    assert r["bit_identical_counts"] and r["bit_identical_range_sums"] and r["bit_identical_ranges_in_push_order"], r
