"""TEST / BASELINE INFRASTRUCTURE: the seed+extend composition on the CPU, vectorised, on top of either
checker (`orc.Ref` = the reference's own templates with OpenMP over all host cores, or `orc.Oracle` = the
single-threaded C port).  Used by bench.py's cpu_baseline / --impl reference legs and by tests.
Only the C calls (match, locate, banded score) are timed; the numpy glue between them is excluded, which
favours the CPU baseline."""
import time
import numpy as np


def gather_2bit(words_u32: np.ndarray, pos: np.ndarray) -> np.ndarray:
    """symbols of a 2-bit big-endian stream at int64 positions (any shape)"""
    w = words_u32[pos >> 4]
    sh = (30 - 2 * (pos & 15)).astype(np.uint32)
    return ((w >> sh) & 3).astype(np.uint8)


def cpu_seed_extend(E, idx, genome_words_u32, reads_sym, seed_len=20, seed_interval=10, band=31, typ=1,
                    scheme=(2, -2, -5, -3), both_strands=True, max_seed_hits=100, count_blocks_with=None, blocks_from_step=0):
    """reads_sym: uint8 [n, L] (fixed length).  Returns dict with best_score[n], timings and counts."""
    n, L = reads_sym.shape
    glen = idx.n
    if both_strands:
        strings = np.empty((2 * n, L), np.uint8)
        strings[0::2] = reads_sym
        strings[1::2] = np.where(reads_sym < 4, 3 - reads_sym, reads_sym)[:, ::-1]
    else:
        strings = reads_sym
    strands = 2 if both_strands else 1
    ns = strings.shape[0]
    K = (L - seed_len) // seed_interval + 1
    starts = np.arange(K) * seed_interval
    cols = starts[:, None] + np.arange(seed_len)[None, :]                  # [K, seed_len]
    q = np.ascontiguousarray(strings[:, cols].reshape(ns * K * seed_len))
    nq = ns * K
    off = (np.arange(nq, dtype=np.uint32) * seed_len).astype(np.uint32)
    ln = np.full(nq, seed_len, np.uint32)

    t0 = time.perf_counter()
    ranges, _ = E.match(idx, q, off, ln)
    t_match = time.perf_counter() - t0
    blocks = blocks_tail = None
    if count_blocks_with is not None:
        _, blocks = count_blocks_with.match(idx, q, off, ln)
        if blocks_from_step:
            _, blocks_tail = count_blocks_with.match(idx, q, off, ln, blocks_from_step=blocks_from_step)

    x = ranges[:, 0].astype(np.int64); y = ranges[:, 1].astype(np.int64)
    sizes = np.where(x <= y, np.minimum(y - x + 1, max_seed_hits), 0)
    excl = np.cumsum(sizes) - sizes
    total = int(sizes.sum())
    qid = np.repeat(np.arange(nq), sizes)
    rows = (np.repeat(x, sizes) + (np.arange(total) - np.repeat(excl, sizes))).astype(np.uint32)

    t0 = time.perf_counter()
    pos = E.locate(idx, rows).astype(np.int64) if total else np.zeros(0, np.int64)
    t_locate = time.perf_counter() - t0

    s_id = qid // K
    sb = (qid % K) * seed_interval
    diag = np.where(pos > sb, pos - sb, 0)
    gb = np.where(diag > band // 2, diag - band // 2, 0)
    ge = np.minimum(gb + L + band, glen)
    wl = (ge - gb).astype(np.int64)
    W = L + band
    # gather each window's symbols into a dense [total, W] buffer (positions past the window end are never read)
    if total:
        wpos = gb[:, None] + np.arange(W)[None, :]
        np.minimum(wpos, glen - 1, out=wpos)
        txt = np.ascontiguousarray(gather_2bit(genome_words_u32, wpos).reshape(-1))
    else:
        txt = np.zeros(1, np.uint8)
    pat = np.ascontiguousarray(strings.reshape(-1))
    p_off = (s_id * L).astype(np.uint32)
    p_len = np.full(total, L, np.uint32)
    t_off = (np.arange(total, dtype=np.int64) * W).astype(np.uint32)
    t_len = wl.astype(np.uint32)

    t0 = time.perf_counter()
    if total:
        score, sx, sy, _ = E.banded_gotoh(band, typ, scheme, pat, p_off, p_len, txt, t_off, t_len)
    else:
        score = np.zeros(0, np.int32); sx = sy = np.zeros(0, np.uint32)
    t_dp = time.perf_counter() - t0

    best = np.full(n, -2**31, np.int64)
    if total:
        np.maximum.at(best, s_id // strands, score.astype(np.int64))
    return dict(best_score=best, n_seeds=nq, n_hits=total, t_match=t_match, t_locate=t_locate, t_dp=t_dp,
                t_total=t_match + t_locate + t_dp, blocks=blocks, blocks_tail=blocks_tail, hit_score=score, hit_read=s_id // strands,
                cells=int(total) * L * band)
