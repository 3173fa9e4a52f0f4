// gotoh_full_core.cuh -- full-matrix (un-banded) Gotoh scoring, one alignment per thread.
//
// Semantics: aln::alignment_score with GotohAligner<TYPE,scheme,PatternBlockingTag>
// (nvbio/alignment/gotoh/gotoh_inl.h:459-960): rows follow the TEXT, columns the PATTERN; E runs along the pattern and
// F along the text, both with the PATTERN gap costs; first column H = text_gap_open + i*text_gap_ext for GLOBAL else 0,
// E = 0 for LOCAL else -inf (:75-89); top row H = Go + (j-1)*Ge for GLOBAL and SEMI_GLOBAL, 0 for LOCAL, F = -inf
// (:688-696); -inf = SHRT_MIN - min(Go,Ge) (:665).  sink = (text end, pattern end).  LOCAL reports every cell in stripes
// of 8 pattern columns (gotoh_bandlen_selector, :1491-1495), row by row inside a stripe, so a tie resolves to the last
// maximal cell in (stripe, row, column) order; SEMI_GLOBAL reports H[i][M] of every row, GLOBAL H[N][M].
//
// Formulation here: 32-column pattern stripes held in registers, the text swept once per stripe, the stripe's right-hand
// boundary column (H,E per text row) kept in HBM scratch between stripes; four 8-column LOCAL trackers per stripe merged
// in order reproduce the reference's tie-breaking exactly.
#pragma once
#include "gotoh_core.cuh"

namespace nvb {

constexpr int FULL_W = 32;     // pattern columns per stripe

template <int TYPE>
__host__ __device__ inline SinkResult gotoh_full(const GotohScheme& S,
        const uint32_t* __restrict__ pwords, uint32_t pbits, uint32_t pbe, uint32_t poff, uint32_t M,
        const uint32_t* __restrict__ twords, uint32_t tbits, uint32_t tbe, uint32_t toff, uint32_t N,
        int2* __restrict__ col, size_t col_stride)
{
    SinkResult res; res.score = INT_MIN; res.x = 0xFFFFFFFFu; res.y = 0xFFFFFFFFu;
    if (M == 0 || N == 0) return res;            // outside the supported domain (see header)
    const int32_t Go = S.pgo, Ge = S.pge;
    const int32_t INF = SHRT_MIN - (Go < Ge ? Go : Ge);

    for (uint32_t b = 0; b < M; b += FULL_W) {
        const bool first = (b == 0), last = (b + FULL_W >= M);
        uint32_t q[FULL_W];
        {
            SymReaderRT pr(pwords, pbits, pbe);
#pragma unroll
            for (int j = 0; j < FULL_W; ++j) q[j] = (b + j < M) ? pr.get(poff + b + j) : 256u;     // 256 never equals a text symbol
        }
        int32_t H[FULL_W + 1], F[FULL_W + 1];
#pragma unroll
        for (int j = 0; j <= FULL_W; ++j) {
            H[j] = (TYPE != NVB_LOCAL) ? ((b + j > 0) ? Go + Ge * (int32_t)(b + j - 1) : 0) : 0;
            F[j] = INF;
        }
        int32_t diag_next = H[0];                                    // H[0][b]
        int32_t tb[4] = { INT_MIN, INT_MIN, INT_MIN, INT_MIN };      // LOCAL trackers of the four 8-column sub-stripes
        uint32_t tp[4] = { 0, 0, 0, 0 };                             // (row << 16) | column, 1-based
        SymReaderRT tr(twords, tbits, tbe);
        for (uint32_t r = 0; r < N; ++r) {
            const uint32_t g = tr.get(toff + r);
            int32_t Hl, E;
            if (first) { Hl = (TYPE == NVB_GLOBAL) ? S.tgo + S.tge * (int32_t)r : 0; E = (TYPE == NVB_LOCAL) ? 0 : INF; }
            else       { const int2 c = col[(size_t)r * col_stride]; Hl = c.x; E = c.y; }
            int32_t Hd = diag_next;
            diag_next = Hl;
            H[0] = Hl;
#pragma unroll
            for (int j = 1; j <= FULL_W; ++j) {
                F[j] = imax2(F[j] + Ge, H[j] + Go);                   // H[j] still holds the previous row
                E    = imax2(E + Ge, H[j - 1] + Go);                  // H[j-1] is already this row
                int32_t h = imax2(imax2(E, F[j]), Hd + ((g == q[j - 1]) ? S.match : S.mismatch));
                if (TYPE == NVB_LOCAL) h = imax2(h, 0);
                Hd = H[j];
                H[j] = h;
                if (TYPE == NVB_LOCAL && b + (uint32_t)j <= M) {
                    const int k = (j - 1) >> 3;
                    if (tb[k] <= h) { tb[k] = h; tp[k] = ((r + 1u) << 16) | (b + (uint32_t)j); }
                }
            }
            if (!last) col[(size_t)r * col_stride] = make_int2(H[FULL_W], E);
            if (TYPE == NVB_SEMI_GLOBAL && last) {
                int32_t hM = H[1];
#pragma unroll
                for (int j = 1; j <= FULL_W; ++j) if (b + (uint32_t)j == M) hM = H[j];
                if (res.score <= hM) { res.score = hM; res.x = r + 1u; res.y = M; }
            }
        }
        if (TYPE == NVB_GLOBAL && last) {
            int32_t hM = H[1];
#pragma unroll
            for (int j = 1; j <= FULL_W; ++j) if (b + (uint32_t)j == M) hM = H[j];
            res.score = hM; res.x = N; res.y = M;
        }
        if (TYPE == NVB_LOCAL) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (b + 8u * (uint32_t)k < M && res.score <= tb[k]) { res.score = tb[k]; res.x = tp[k] >> 16; res.y = tp[k] & 0xFFFFu; }
        }
    }
    return res;
}

} // namespace nvb
