"""-m gpu: the drop-in boundary on the reference's side.  tests/shim/shim_harness.cu is ONE program
written against nvbio's own types -- io::FMIndexDataDevice::fm_index_type, FMIndexFilterDevice::rank/locate over an InfixSet of
4-bit reads, aln::batch_banded_alignment_score<31>, BatchedBandedAlignmentScore<...>::enact on a user-defined stream,
aln::batch_alignment_score -- compiled twice against /root/reference (tests/shim/Makefile): once as is (the reference's own
templates and kernels) and once with include/nvbio_b200/shim/nvbio_shim.h + libnvbio_b200.so.  Both prebuilt binaries travel
with the snapshot; here they run on the same deterministic inputs and every result array must be bit-identical, and the shim
build must really have gone through the B200 kernels (call counters, no fall-backs)."""
import filecmp
import json
import os
import subprocess

import pytest

from tests.gpu_util import require_gpu

pytestmark = pytest.mark.gpu

BIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim", "_bin")
REF, B200 = os.path.join(BIN, "shim_harness_ref"), os.path.join(BIN, "shim_harness_b200")


@pytest.fixture(scope="module")
def binaries():
    require_gpu()
    if not (os.path.exists(REF) and os.path.exists(B200)):
        pytest.skip("tests/shim/_bin not built (needs /root/reference at build time)")
    return REF, B200


def _run(binary, args, out):
    os.makedirs(out, exist_ok=True)
    r = subprocess.run([binary] + args[:1] + [out] + args[1:], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, "%s failed: %s" % (os.path.basename(binary), (r.stderr or r.stdout)[-800:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def _same_dumps(a, b):
    fa, fb = sorted(os.listdir(a)), sorted(os.listdir(b))
    assert fa == fb and fa, (fa, fb)
    diff = [f for f in fa if not filecmp.cmp(os.path.join(a, f), os.path.join(b, f), shallow=False)]
    assert not diff, "arrays differ between the reference build and the shim build: %s" % diff
    return fa


def test_fmmap_program_identical_with_and_without_shim(binaries, tmp_path):
    ref = _run(REF, ["fmmap", "3000000", "40000", "100"], str(tmp_path / "ref"))
    b2 = _run(B200, ["fmmap", "3000000", "40000", "100"], str(tmp_path / "b200"))
    files = _same_dumps(str(tmp_path / "ref"), str(tmp_path / "b200"))
    assert {"fmmap_ranges.bin", "fmmap_slots.bin", "fmmap_hits.bin", "fmmap_scores_i16.bin", "fmmap_sinks_i32.bin", "fmmap_best_i32.bin"} <= set(files)
    assert ref["hits"] == b2["hits"] and ref["hits"] > ref["reads"]
    assert ref["shim"] == 0 and b2["shim"] == 1
    c = b2["b200_calls"]
    assert c["fm_rank"] == 3 and c["fm_locate"] == 6 and c["banded"] == 6 and c["fallbacks"] == 0, c
    print("fmmap: reference ms", {k: ref[k] for k in ref if k.endswith("_ms")}, "shim ms", {k: b2[k] for k in b2 if k.endswith("_ms")})


def test_batch_program_identical_with_and_without_shim(binaries, tmp_path):
    ref = _run(REF, ["batch", "65536"], str(tmp_path / "ref"))
    b2 = _run(B200, ["batch", "65536"], str(tmp_path / "b200"))
    files = _same_dumps(str(tmp_path / "ref"), str(tmp_path / "b200"))
    assert len(files) == 7 + 6
    c = b2["b200_calls"]
    assert c["banded"] == 21 and c["full"] == 9 and c["fallbacks"] == 0, c
    print("batch GCUPS: reference", {k: ref[k] for k in ref if k.endswith("gcups")}, "shim", {k: b2[k] for k in b2 if k.endswith("gcups")})


def test_nvbowtie_mapping_entry_points_identical(tmp_path):
    """boundary B-A2: nvBowtie's non-template map_exact / map_approx / map / gather_ranges (nvBowtie/bowtie2/cuda/mapping.h) called on
    nvBowtie's own PODs by ONE harness object, linked once against nvBowtie's mapping.cu (compiled where it lies) and once against
    tests/shim/nvbowtie_mapping_b200.cu + libnvbio_b200.so: deque sizes, the SeedHits of every read (canonical order), the range sizes
    in pop_top() order, reseed flags and gather_ranges totals are identical; both leave valid priority deques behind"""
    require_gpu()
    ref, b2 = os.path.join(BIN, "nvbowtie_harness_ref"), os.path.join(BIN, "nvbowtie_harness_b200")
    if not (os.path.exists(ref) and os.path.exists(b2)):
        pytest.skip("tests/shim/_bin/nvbowtie_harness_* not built (needs /root/reference at build time)")
    outs = {}
    for name, binary in (("ref", ref), ("b200", b2)):
        d = str(tmp_path / name); os.makedirs(d)
        r = subprocess.run([binary, d, "400000", "4000"], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, "%s failed: %s" % (name, (r.stderr or r.stdout)[-800:])
        outs[name] = json.loads(r.stdout.strip().splitlines()[-1])
    files = _same_dumps(str(tmp_path / "ref"), str(tmp_path / "b200"))
    assert "exact_r0_hits.bin" in files and "approx_r0_hits.bin" in files and "exact_tiny_sizes.bin" in files
    for cfg in ("exact_r0", "exact_r1", "exact_fw", "approx_r0", "approx_rc", "map_sub", "map_nosub", "exact_tiny"):
        a, b = outs["ref"][cfg], outs["b200"][cfg]
        assert a["hits"] == b["hits"] > 0 and a["range_total"] == b["range_total"] and a["full_deques"] == b["full_deques"], (cfg, a, b)
        assert a["pop_order_ok"] == 1 and b["pop_order_ok"] == 1, cfg
    assert outs["ref"]["exact_tiny"]["full_deques"] > 100
    print("nvBowtie mapping ms: reference", {k: v["ms"] for k, v in outs["ref"].items() if isinstance(v, dict)},
          "b200", {k: v["ms"] for k, v in outs["b200"].items() if isinstance(v, dict)})


def test_nvbowtie_score_best_identical_with_and_without_shim(tmp_path):
    """boundary B-B2: nvBowtie's best-score extension (detail::banded_score_best = the body of score_best_t / score_best,
    score_best_inl.h:154-234) on nvBowtie's own types -- reversed DNA_N reads with base qualities, 2-bit genome, HitQueues,
    SmithWatermanScoringScheme<> (--local preset and end-to-end default, quality-dependent mismatch penalties), ParamsPOD -- compiled
    as is and with include/nvbio_b200/shim/nvbowtie_scoring.h: hit.score (incl. the worst_score clamp) and hit.sink of every hit equal"""
    require_gpu()
    ref, b2 = os.path.join(BIN, "nvbowtie_score_harness_ref"), os.path.join(BIN, "nvbowtie_score_harness_b200")
    if not (os.path.exists(ref) and os.path.exists(b2)):
        pytest.skip("tests/shim/_bin/nvbowtie_score_harness_* not built (needs /root/reference at build time)")
    outs = {}
    for name, binary in (("ref", ref), ("b200", b2)):
        d = str(tmp_path / name); os.makedirs(d)
        r = subprocess.run([binary, d, "2000000", "20000", "200000"], capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, "%s failed: %s" % (name, (r.stderr or r.stdout)[-800:])
        outs[name] = json.loads(r.stdout.strip().splitlines()[-1])
    files = _same_dumps(str(tmp_path / "ref"), str(tmp_path / "b200"))
    assert len(files) == 8
    import numpy as np
    sc = np.fromfile(str(tmp_path / "ref" / "local_b31_score.bin"), np.int32)
    assert (sc > 60).mean() > 0.5 and (sc != 12345).all()
    assert outs["b200"]["b200_calls"] == {"banded": 12, "fallbacks": 0}
    print("nvBowtie score_best ms: reference", {k: v for k, v in outs["ref"].items() if k.endswith("_ms")}, "b200", {k: v for k, v in outs["b200"].items() if k.endswith("_ms")})
