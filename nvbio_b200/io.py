"""FM-index files in the reference's (BWA-compatible) format, as written by nvBWT and read by
io::FMIndexDataHost::load (nvBWT/nvBWT.cu:314-351, nvbio/io/fmindex/fmindex_impl.cu:120-260):

  <prefix>.bwt : uint32 primary, uint32 cumFreq[4] (cumFreq[3] = text length), packed 2-bit big-endian BWT words
  <prefix>.sa  : uint32 primary, uint32 cumFreq[4], uint32 SA_INT, uint32 text length, uint32 ssa[1..]  (ssa[0] = -1 implied)

The loader hands the BWT words to the device, where the occurrence table is built and interleaved
(nvb_fm_build_occ) -- the reference does that serially on the host (fmindex_impl.cu:263-331)."""
import numpy as np
import torch
from .fmindex import FMIndexDevice


def save_index(prefix: str, bwt_occ: np.ndarray, ssa: np.ndarray, L2, length: int, primary: int, sa_interval: int = 16):
    """write <prefix>.bwt / <prefix>.sa from host arrays in the in-memory layout (32-byte {bwt,occ} blocks)"""
    L2 = np.asarray(L2, dtype=np.uint32)
    hdr = np.array([primary, L2[1], L2[2], L2[3], L2[4]], dtype=np.uint32)
    blocks = np.ascontiguousarray(bwt_occ, dtype=np.uint32).reshape(-1, 8)
    words = np.ascontiguousarray(blocks[:, :4]).reshape(-1)
    with open(prefix + ".bwt", "wb") as f:
        hdr.tofile(f)
        words.tofile(f)
    with open(prefix + ".sa", "wb") as f:
        hdr.tofile(f)
        np.array([sa_interval, length], dtype=np.uint32).tofile(f)
        np.ascontiguousarray(ssa, dtype=np.uint32)[1:].tofile(f)


def load_index(prefix: str, device="cuda") -> FMIndexDevice:
    """read <prefix>.bwt (+ <prefix>.sa when present) and build the device index"""
    raw = np.fromfile(prefix + ".bwt", dtype=np.uint32)
    primary, cum = int(raw[0]), raw[1:5]
    n = int(cum[3])
    seq_words = ((n + 63) // 64) * 4                    # whole 64-symbol blocks, as the reference pads (align<4>)
    words = np.zeros(seq_words + 4, dtype=np.uint32)
    body = raw[5:5 + seq_words]
    if len(body) < (n + 15) // 16:
        raise IOError("%s.bwt is truncated" % prefix)
    words[:len(body)] = body
    ssa, interval = None, 16
    try:
        sa_raw = np.fromfile(prefix + ".sa", dtype=np.uint32)
    except FileNotFoundError:
        sa_raw = None
    if sa_raw is not None:
        if int(sa_raw[0]) != primary or int(sa_raw[6]) != n:
            raise IOError("SA file mismatch: primary/length differ from the .bwt")     # the reference throws file_mismatch
        interval = int(sa_raw[5])
        n_items = (n + interval) // interval
        ssa_h = np.empty(n_items, dtype=np.uint32)
        ssa_h[0] = 0xFFFFFFFF
        ssa_h[1:] = sa_raw[7:7 + n_items - 1]
        ssa = torch.from_numpy(ssa_h.view(np.int32)).to(device)
    d_words = torch.from_numpy(words.view(np.int32)).to(device)
    fmi = FMIndexDevice.from_bwt(d_words, n, primary, ssa, sa_interval=interval)
    got = np.array(fmi.L2[1:], dtype=np.uint64)
    if not np.array_equal(got, cum.astype(np.uint64)):
        raise IOError("cumulative symbol counts of %s.bwt do not match its header" % prefix)
    return fmi
