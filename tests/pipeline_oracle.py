"""Oracle composition of the seed+extend pipeline (test infrastructure): the same stages as
nvbio_b200/csrc/pipeline.cu, each stage computed by the CPU oracle / reference library."""
import numpy as np


def seed_extend_oracle(O, idx, genome_sym, reads, params):
    """reads: list of uint8 arrays.  Returns dict(best_score, best_pos, hit_string, hit_window, hit_score, hit_sink, n_hits)."""
    L, I, B = params.seed_len, params.seed_interval, params.band_len
    n = idx.n
    strands = 2 if params.both_strands else 1
    strings = []
    for r in reads:
        strings.append(r)
        if strands == 2:
            rc = np.where(r < 4, 3 - r, r)[::-1].astype(np.uint8)
            strings.append(rc)
    max_len = max(len(r) for r in reads)
    K = (max_len - L) // I + 1
    # seeds
    q, off, ln, valid = [], [], [], []
    o = 0
    for s in strings:
        for k in range(K):
            p = k * I
            if p + L <= len(s):
                q.append(s[p:p + L]); off.append(o); ln.append(L); o += L; valid.append(True)
            else:
                off.append(o); ln.append(0); valid.append(False)
    qcat = np.concatenate(q) if q else np.zeros(1, np.uint8)
    ranges, _ = O.match(idx, qcat, np.array(off, np.uint32), np.array(ln, np.uint32))
    hit_string, rows, seed_k = [], [], []
    for qi, ((x, y), v) in enumerate(zip(ranges, valid)):
        if not v or x > y:
            continue
        sz = min(int(y) - int(x) + 1, params.max_seed_hits)
        for j in range(sz):
            hit_string.append(qi // K); seed_k.append(qi % K); rows.append(int(x) + j)
    rows = np.array(rows, np.uint32)
    pos = O.locate(idx, rows) if len(rows) else np.zeros(0, np.uint32)
    p_sym, p_off, p_len, t_off, t_len, wins = [], [], [], [], [], []
    po = 0
    for s, k, p in zip(hit_string, seed_k, pos):
        ln_s = len(strings[s])
        sb = k * I
        diag = int(p) - sb if int(p) > sb else 0
        gb = diag - B // 2 if diag > B // 2 else 0
        ge = min(gb + ln_s + B, n)
        p_sym.append(strings[s]); p_off.append(po); p_len.append(ln_s); po += ln_s
        t_off.append(gb); t_len.append(ge - gb); wins.append((gb, ge))
    sch = params.scheme
    scheme = (sch.match, sch.mismatch, sch.gap_open, sch.gap_ext)
    if hit_string:
        score, sx, sy, _ = O.banded_gotoh(B, params.type, scheme, np.concatenate(p_sym), np.array(p_off, np.uint32), np.array(p_len, np.uint32),
                                          genome_sym, np.array(t_off, np.uint32), np.array(t_len, np.uint32))
    else:
        score = np.zeros(0, np.int32); sx = sy = np.zeros(0, np.uint32)
    best_score = np.full(len(reads), -2**31, np.int64)
    best_pos = np.full(len(reads), 0xFFFFFFFF, np.int64)
    for h, s in enumerate(hit_string):
        r = s // strands
        if int(score[h]) > best_score[r]:
            best_score[r] = int(score[h]); best_pos[r] = t_off[h] + int(sx[h])
    return dict(best_score=best_score, best_pos=best_pos, hit_string=np.array(hit_string, np.int64),
                hit_window=np.array(wins, np.int64).reshape(-1, 2), hit_score=score.astype(np.int64),
                hit_sink=np.stack([sx.astype(np.int64), sy.astype(np.int64)], axis=1) if len(sx) else np.zeros((0, 2), np.int64),
                n_hits=len(hit_string))
