// gotoh_core.cuh -- per-thread banded Gotoh scoring (HP-B).
//
// Semantics follow nvbio's priv::banded::gotoh_alignment_score_dispatch<BAND_LEN,TYPE>::run
// (nvbio/alignment/gotoh/gotoh_banded_inl.h:406-658), the row-0 initialisation (:46-77), BestSink's
// `<=` tie-break (nvbio/alignment/sink_inl.h:57-65) and the 2-bit sliding text cache used for every
// BAND_LEN outside {3,5,7,15} (nvbio/alignment/alignment_base_inl.h:75-99).  Band cell j of row i is
// DP cell (pattern i, text i+j);   F[j] <- max(F[j+1]+Ge, H[j+1]+Go)   (previous row's H[j+1]),
// H[j] <- max(F[j], E, H[j]+S(text[i+j],pattern[i]))  (LOCAL: clamped at 0, every cell reported),
// E <- max(H[j]+Go, E+Ge) chained left to right.  E and F both use the PATTERN gap costs.
//
// Two formulations, both new:
//   * gotoh_generic<B,TYPE>  one alignment per thread, int32, any symbol width / quality table;
//   * gotoh_pair<B,TYPE>     TWO alignments per thread packed as s16x2 halves, driven by Blackwell's
//                            DPX instructions (VIADDMNMX.S16x2[.RELU], VIMNMX[3].S16x2, VIADD.16x2):
//                            5 DPX ops + 1 PRMT (substitution lookup) per cell PAIR, the LOCAL sink
//                            tracked with one IMAD + half a VIMNMX3.U16x2 per cell pair.
#pragma once
#include "common.cuh"
#include <limits.h>

namespace nvb {

struct GotohScheme {
    int32_t match, mismatch, pgo, pge, tgo, tge;
    const int32_t* qtab;       // device (or host in host tests): [256][2] = {match(q), mismatch(q)} or NULL
    // multipliers that are 1 and 32 at run time but opaque to the compiler (kernel-parameter constants), so that
    // "x*one + y" and "x*keymul + y" stay IMADs on the FMA pipe instead of IADD3/LEA on the saturated ALU pipe
    uint32_t one, keymul;
};
static inline GotohScheme make_scheme(const nvb_gotoh_scheme* s) {
    GotohScheme r; r.match = s->match; r.mismatch = s->mismatch; r.pgo = s->pattern_gap_open; r.pge = s->pattern_gap_ext;
    r.tgo = s->text_gap_open; r.tge = s->text_gap_ext; r.qtab = s->d_qual_table; r.one = 1u; r.keymul = 32u; return r;
}

__host__ __device__ __forceinline__ int32_t imax2(int32_t a, int32_t b) { return a > b ? a : b; }
__host__ __device__ __forceinline__ uint32_t umin2(uint32_t a, uint32_t b) { return a < b ? a : b; }

__host__ __device__ __forceinline__ int32_t gotoh_infimum(const GotohScheme& S) {
    return SHRT_MIN - imax2(imax2(S.pgo, S.pge), imax2(S.tgo, S.tge));
}
__host__ __device__ __forceinline__ constexpr bool packed_text_cache(int B) { return !(B == 3 || B == 5 || B == 7 || B == 15); }

// runtime-format sequential symbol reader (bits in {2,4,8})
struct SymReaderRT {
    const uint32_t* words; uint32_t bits, be, log_spw, cur_idx, cur_word;
    __host__ __device__ __forceinline__ SymReaderRT(const uint32_t* w, uint32_t b, uint32_t e)
        : words(w), bits(b), be(e), log_spw(b == 2 ? 4 : (b == 4 ? 3 : 2)), cur_idx(0xFFFFFFFFu), cur_word(0) {}
    __host__ __device__ __forceinline__ uint32_t get(uint32_t p) {
        const uint32_t wi = p >> log_spw;
        if (wi != cur_idx) { cur_idx = wi; cur_word = words[wi]; }
        const uint32_t r = p & ((1u << log_spw) - 1u);
        const uint32_t sh = (bits == 8) ? 8u * r : (be ? 32u - bits - bits * r : bits * r);
        return (cur_word >> sh) & ((1u << bits) - 1u);
    }
};

// Sequential reader for the per-row pattern symbols of the packed kernels: each word is normalised to big-endian symbol order
// when it is loaded, so a symbol costs one shift to extract and one to advance (SymReaderRT::get recomputes word index and
// shift amount per call).
struct PatStream {
    const uint32_t* wp; uint32_t w, left, bits, be, spw;
    __host__ __device__ __forceinline__ static uint32_t bitrev32(uint32_t x) {
#ifdef __CUDA_ARCH__
        return __brev(x);
#else
        x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1); x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
        x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4); x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
        return (x >> 16) | (x << 16);
#endif
    }
    __host__ __device__ __forceinline__ uint32_t normalise(uint32_t x) const {
        if (bits == 8) return ((x >> 24) & 0xFFu) | ((x >> 8) & 0xFF00u) | ((x << 8) & 0xFF0000u) | (x << 24);    // bytes are little-endian
        if (be) return x;
        uint32_t y = bitrev32(x);                                        // symbol order reversed, and the bits inside each symbol
        if (bits == 2) return ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
        return ((y & 0x11111111u) << 3) | ((y & 0x22222222u) << 1) | ((y >> 1) & 0x22222222u) | ((y >> 3) & 0x11111111u);
    }
    __host__ __device__ __forceinline__ PatStream(const uint32_t* words, uint32_t b, uint32_t e, uint32_t off, uint32_t /*len*/ = 0u)
        : bits(b), be(e), spw(32u / b) {
        const uint32_t lg = (b == 2 ? 4u : (b == 4 ? 3u : 2u));
        const uint32_t r = off & (spw - 1u);
        wp = words + (off >> lg);
        w = normalise(*wp++) << (bits * r);        // r < spw, so the shift is < 32
        left = spw - r;
    }
    __host__ __device__ __forceinline__ uint32_t next() {
        if (left == 0u) { w = normalise(*wp++); left = spw; }
        const uint32_t s = w >> (32u - bits);
        w <<= bits;
        --left;
        return s;
    }
};

// The same for a pattern format known at compile time (BITS-bit symbols, big-endian words: nvBowtie's 4-bit reads, 2-bit reads): no
// run-time format dispatch in the per-row fetch -- a shift to extract, a shift to advance, a counter
template <int BITS>
struct PatStreamBE {
    // one word of look-ahead: the word after the current one is requested when the current one is taken into use, 16 / 8 rows before
    // its first symbol is needed, so the row loop never waits for a global load (never beyond the pattern's last word)
    const uint32_t* wp; const uint32_t* last; uint32_t w, nxt, left;
    __host__ __device__ __forceinline__ PatStreamBE(const uint32_t* words, uint32_t /*bits*/, uint32_t /*be*/, uint32_t off, uint32_t len) {
        constexpr uint32_t SPW = 32u / BITS, LG = (BITS == 2 ? 4u : 3u);
        const uint32_t r = off & (SPW - 1u);
        wp = words + (off >> LG);
        last = words + ((off + (len ? len - 1u : 0u)) >> LG);
        w = *wp << (BITS * r);
        nxt = (wp < last) ? wp[1] : 0u;
        left = SPW - r;
    }
    // branch-free (selects and one predicated load), so that the row loop of gotoh_pair is a single basic block; past the end of
    // the pattern it keeps delivering symbol 0
    __host__ __device__ __forceinline__ uint32_t next() {
        constexpr uint32_t SPW = 32u / BITS;
        const bool refill = (left == 0u);
        const bool more = refill && (wp + 1 < last);
        const uint32_t fresh = more ? wp[2] : 0u;
        w = refill ? nxt : w;
        nxt = refill ? fresh : nxt;
        wp += refill ? 1 : 0;
        left = refill ? SPW : left;
        const uint32_t s = w >> (32u - BITS);
        w <<= BITS;
        --left;
        return s;
    }
};
// Word-at-a-time reader of a BITS-bit big-endian pattern: next() returns the next 32 / BITS symbols ALIGNED to the top of a word
// (a funnel shift of two consecutive stream words; the pattern may start at any symbol offset), with one word of look-ahead.
// Past the pattern's last word it delivers zeros and touches no memory.
template <int BITS>
struct PatChunksBE {
    const uint32_t* wp; const uint32_t* last; uint32_t cur, nxt, sh;
    __host__ __device__ __forceinline__ PatChunksBE(const uint32_t* words, uint32_t off, uint32_t len) {
        constexpr uint32_t SPW = 32u / BITS, LG = (BITS == 2 ? 4u : 3u);
        wp = words + (off >> LG);
        last = words + ((off + (len ? len - 1u : 0u)) >> LG);
        sh = BITS * (off & (SPW - 1u));
        cur = *wp;
        nxt = (wp < last) ? wp[1] : 0u;
    }
    __host__ __device__ __forceinline__ uint32_t next() {
#ifdef __CUDA_ARCH__
        const uint32_t c = __funnelshift_l(nxt, cur, sh);
#else
        const uint32_t c = sh ? ((cur << sh) | (nxt >> (32u - sh))) : cur;
#endif
        ++wp; cur = nxt;
        nxt = (wp < last) ? wp[1] : 0u;
        return c;
    }
};
template <int PFMT> struct PatStreamOf      { typedef PatStream type; };
template <>         struct PatStreamOf<2>   { typedef PatStreamBE<2> type; };
template <>         struct PatStreamOf<4>   { typedef PatStreamBE<4> type; };

struct SinkResult { int32_t score; uint32_t x, y; };

// ---------------------------------------------------------------------------------------------
// generic: one alignment, int32
// ---------------------------------------------------------------------------------------------
// direction vectors of the traceback (nvbio::aln::DirectionVector, nvbio/alignment/alignment_base.h:139-150)
enum { DIR_SUB = 0, DIR_INS = 1, DIR_DEL = 2, DIR_SINK = 3, DIR_INS_EXT = 4, DIR_DEL_EXT = 8 };
template <int B> struct DirWords { static constexpr int N = (B * 4 + 31) / 32; };     // 4 bits per band cell

// aln::Best2Sink<int32> (nvbio/alignment/sink.h:114-147, sink_inl.h:70-116): the best alignment (ties: the last report wins) and the
// best one whose text end lies more than `dist` away from it -- fed with the very reports BestSink receives
struct Best2 {
    int32_t s1, s2; uint32_t x1, y1, x2, y2, dist;
    __host__ __device__ __forceinline__ void init(uint32_t d) { s1 = s2 = NVB_SINK_MIN; x1 = y1 = x2 = y2 = 0xFFFFFFFFu; dist = d; }
    __host__ __device__ __forceinline__ void report(int32_t s, uint32_t x, uint32_t y) {
        if (s1 <= s) { s1 = s; x1 = x; y1 = y; }
        else if (s2 <= s && (x + dist < x1 || x > x1 + dist)) { s2 = s; x2 = x; y2 = y; }
    }
};

// DIRS: also emit, per row, the packed 4-bit direction vectors (H | E | F flow) of every band cell to
// dirs[row * DirWords<B>::N ..] -- exactly the bits GotohSubmatrixContext::new_cell stores (gotoh_banded_inl.h:325-337)
template <int B, int TYPE, bool DIRS>
__host__ __device__ inline SinkResult gotoh_generic_impl(const GotohScheme& S,
        const uint32_t* __restrict__ pwords, uint32_t pbits, uint32_t pbe, uint32_t poff, uint32_t M,
        const uint8_t* __restrict__ quals,
        const uint32_t* __restrict__ twords, uint32_t tbits, uint32_t tbe, uint32_t toff, uint32_t N,
        uint32_t* __restrict__ dirs, Best2* best2 = nullptr)
{
    SinkResult res; res.score = NVB_SINK_MIN; res.x = 0xFFFFFFFFu; res.y = 0xFFFFFFFFu;
    if (N < M) return res;

    constexpr bool PACKED = packed_text_cache(B);
    constexpr int  NW = DirWords<B>::N;
    const int32_t Go = S.pgo, Ge = S.pge;
    const int32_t INF = gotoh_infimum(S);

    int32_t H[B], F[B];
    uint32_t cache[B];        // cache[j] = (quirk-adjusted) text symbol of band cell j, j < B-1
    H[0] = 0;
#pragma unroll
    for (int j = 1; j < B; ++j) H[j] = (TYPE == NVB_GLOBAL) ? S.tgo + (j - 1) * S.tge : 0;
#pragma unroll
    for (int j = 0; j < B; ++j) F[j] = INF;

    SymReaderRT tr(twords, tbits, tbe), pr(pwords, pbits, pbe);
#pragma unroll
    for (int j = 0; j < B - 1; ++j) {
        // the reference reads text[j], j < B-1, without a bound check (undefined beyond N); we define 255
        const uint32_t g = ((uint32_t)j < N) ? tr.get(toff + j) : 255u;
        cache[j] = PACKED ? (g & 3u) : g;
    }

    int32_t best = INT_MIN; uint32_t bpos = 0;     // bpos = (i << 6) | j of the last maximal LOCAL cell
    for (uint32_t i = 0; i < M; ++i) {
        const uint32_t q = pr.get(poff + i);
        const uint32_t qq = quals ? quals[poff + i] : 0u;
        const int32_t s_eq = S.qtab ? S.qtab[2 * qq]     : S.match;
        const int32_t s_ne = S.qtab ? S.qtab[2 * qq + 1] : S.mismatch;
        const uint32_t g_new = (i + (uint32_t)B - 1u < N) ? tr.get(toff + i + B - 1) : 255u;
        int32_t E = 0;
        uint32_t edir = DIR_SUB;
        uint32_t dw[NW];
        if (DIRS) {
#pragma unroll
            for (int w = 0; w < NW; ++w) dw[w] = 0u;
        }
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const uint32_t g = (j < B - 1) ? cache[j] : g_new;          // cell B-1 sees the unmasked symbol
            if (j >= 1 && j < B - 1) cache[j - 1] = g;
            uint32_t fdir = DIR_SUB;
            if (j < B - 1) {
                const int32_t ftop = F[j + 1] + Ge, htop = H[j + 1] + Go;
                F[j] = imax2(ftop, htop);
                if (DIRS) fdir = ftop > htop ? DIR_DEL_EXT : DIR_SUB;
            } else F[j] = INF;
            const int32_t diagonal = H[j] + ((g == q) ? s_eq : s_ne);
            const int32_t top = F[j], left = E;
            int32_t h = diagonal;
            if (j < B - 1) h = imax2(h, top);
            if (j > 0)     h = imax2(h, left);
            uint32_t hdir = DIR_SUB;
            if (DIRS) {
                if (j == 0)          hdir = top > diagonal ? DIR_INS : DIR_SUB;
                else if (j < B - 1)  hdir = top > left ? (top > diagonal ? DIR_INS : DIR_SUB) : (left > diagonal ? DIR_DEL : DIR_SUB);
                else                 hdir = left > diagonal ? DIR_DEL : DIR_SUB;
            }
            if (TYPE == NVB_LOCAL) {
                h = imax2(h, 0);
                if (DIRS && h == 0) hdir = DIR_SINK;
                if (best <= h) { best = h; bpos = (i << 6) | (uint32_t)j; }
                if (best2) best2->report(h, i + (uint32_t)j + 1u, i + 1u);
            }
            H[j] = h;
            if (DIRS) dw[j >> 3] |= (hdir | edir | fdir) << (4 * (j & 7));
            if (j == 0) { E = h + Go; edir = DIR_SUB; }
            else {
                const int32_t eleft = E + Ge, ediagonal = h + Go;
                if (DIRS) edir = eleft > ediagonal ? DIR_INS_EXT : DIR_SUB;
                E = imax2(ediagonal, eleft);
            }
        }
        cache[B - 2] = PACKED ? (g_new & 3u) : g_new;
        if (DIRS) {
#pragma unroll
            for (int w = 0; w < NW; ++w) dirs[(size_t)i * NW + w] = dw[w];
        }
    }
    if (TYPE == NVB_LOCAL) {
        if (M > 0) { res.score = best; res.x = (bpos >> 6) + (bpos & 63u) + 1u; res.y = (bpos >> 6) + 1u; }
    } else if (TYPE == NVB_GLOBAL) {
        res.score = H[B - 1]; res.x = M + (uint32_t)B - 1u; res.y = M;
        if (best2) best2->report(H[B - 1], M + (uint32_t)B - 1u, M);
    } else {
        const uint32_t m = umin2(M + (uint32_t)B - 1u, N) - (M - 1u);
        res.score = H[0]; res.x = M; res.y = M;
        if (best2) best2->report(H[0], M, M);
#pragma unroll
        for (int j = 1; j < B; ++j)
            if ((uint32_t)j < m) {
                if (res.score <= H[j]) { res.score = H[j]; res.x = M + (uint32_t)j; }
                if (best2) best2->report(H[j], M + (uint32_t)j, M);
            }
    }
    return res;
}

template <int B, int TYPE>
__host__ __device__ inline SinkResult gotoh_generic(const GotohScheme& S,
        const uint32_t* __restrict__ pwords, uint32_t pbits, uint32_t pbe, uint32_t poff, uint32_t M,
        const uint8_t* __restrict__ quals,
        const uint32_t* __restrict__ twords, uint32_t tbits, uint32_t tbe, uint32_t toff, uint32_t N)
{
    return gotoh_generic_impl<B, TYPE, false>(S, pwords, pbits, pbe, poff, M, quals, twords, tbits, tbe, toff, N, nullptr);
}

// Windowed scoring: rows [wb, we) of the band only, carrying the (H, F) band between calls in a short2 checkpoint
// (aln::banded_alignment_score<B>(..., window_begin, window_end, sink, checkpoint), banded_inl.h:178-218;
// GotohCheckpointedScoringContext, gotoh_banded_inl.h:132-199; the early exit :616-634).  `res` is the caller's BestSink (in/out).
// Returns false when N < M or when the band maximum can no longer reach min_score (the checkpoint is then left untouched).
template <int B, int TYPE>
__host__ __device__ inline bool gotoh_window(const GotohScheme& S,
        const uint32_t* __restrict__ pwords, uint32_t pbits, uint32_t pbe, uint32_t poff, uint32_t M,
        const uint8_t* __restrict__ quals,
        const uint32_t* __restrict__ twords, uint32_t tbits, uint32_t tbe, uint32_t toff, uint32_t N,
        uint32_t wb, uint32_t we, int32_t min_score, short2* __restrict__ ckpt, SinkResult& res)
{
    if (N < M) return false;
    constexpr bool PACKED = packed_text_cache(B);
    const int32_t Go = S.pgo, Ge = S.pge;
    const int32_t INF = gotoh_infimum(S);
    int32_t H[B], F[B];
    uint32_t cache[B];
    if (wb == 0) {
        H[0] = 0;
#pragma unroll
        for (int j = 1; j < B; ++j) H[j] = (TYPE == NVB_GLOBAL) ? S.tgo + (j - 1) * S.tge : 0;
#pragma unroll
        for (int j = 0; j < B; ++j) F[j] = INF;
    } else {
#pragma unroll
        for (int j = 0; j < B; ++j) { const short2 c = ckpt[j]; H[j] = c.x; F[j] = c.y; }
    }
    SymReaderRT tr(twords, tbits, tbe), pr(pwords, pbits, pbe);
#pragma unroll
    for (int j = 0; j < B - 1; ++j) {
        const uint32_t g = (wb + (uint32_t)j < N) ? tr.get(toff + wb + j) : 255u;      // unchecked in the reference: 255 here
        cache[j] = PACKED ? (g & 3u) : g;
    }
    for (uint32_t i = wb; i < we; ++i) {
        const uint32_t q = pr.get(poff + i);
        const uint32_t qq = quals ? quals[poff + i] : 0u;
        const int32_t s_eq = S.qtab ? S.qtab[2 * qq]     : S.match;
        const int32_t s_ne = S.qtab ? S.qtab[2 * qq + 1] : S.mismatch;
        const uint32_t g_new = (i + (uint32_t)B - 1u < N) ? tr.get(toff + i + B - 1) : 255u;
        int32_t E = 0;
#pragma unroll
        for (int j = 0; j < B; ++j) {
            const uint32_t g = (j < B - 1) ? cache[j] : g_new;
            if (j >= 1 && j < B - 1) cache[j - 1] = g;
            if (j < B - 1) F[j] = imax2(F[j + 1] + Ge, H[j + 1] + Go); else F[j] = INF;
            int32_t h = H[j] + ((g == q) ? s_eq : s_ne);
            if (j < B - 1) h = imax2(h, F[j]);
            if (j > 0)     h = imax2(h, E);
            if (TYPE == NVB_LOCAL) {
                h = imax2(h, 0);
                if (res.score <= h) { res.score = h; res.x = i + (uint32_t)j + 1u; res.y = i + 1u; }
            }
            H[j] = h;
            E = (j == 0) ? h + Go : imax2(h + Go, E + Ge);
        }
        cache[B - 2] = PACKED ? (g_new & 3u) : g_new;
    }
    if (we < M) {
        int32_t mx = H[0];
#pragma unroll
        for (int j = 1; j < B; ++j) mx = imax2(mx, H[j]);
        const long long thr = (long long)min_score + (long long)(M - we) * (long long)(S.qtab ? S.qtab[0] : S.match);
        if ((long long)mx < thr) return false;
    }
#pragma unroll
    for (int j = 0; j < B; ++j) ckpt[j] = make_short2((short)imax2(H[j], SHRT_MIN + 32), (short)imax2(F[j], SHRT_MIN + 32));
    if (we == M) {
        if (TYPE == NVB_GLOBAL) {
            if (res.score <= H[B - 1]) { res.score = H[B - 1]; res.x = M + (uint32_t)B - 1u; res.y = M; }
        } else if (TYPE == NVB_SEMI_GLOBAL) {
            const uint32_t m = umin2(M + (uint32_t)B - 1u, N) - (M - 1u);
#pragma unroll
            for (int j = 0; j < B; ++j)
                if ((j == 0 || (uint32_t)j < m) && res.score <= H[j]) { res.score = H[j]; res.x = M + (uint32_t)j; res.y = M; }
        }
    }
    return true;
}

// Backtrack through the direction matrix from the sink (the H/E/F state machine of
// priv::banded_alignment_traceback, gotoh_banded_inl.h:893-958, walked over the whole matrix instead of one
// 32-row checkpoint window at a time).  Pushes ops in END -> START order (0 SUBSTITUTION 'M', 1 INSERTION 'I',
// 2 DELETION 'D'); returns their number (ops beyond max_ops are counted but not stored) and the source cell.
template <int B, int TYPE>
__host__ __device__ inline uint32_t gotoh_walk(const uint32_t* __restrict__ dirs, const SinkResult& sink,
                                               uint8_t* __restrict__ ops, uint32_t max_ops, uint32_t& src_x, uint32_t& src_y)
{
    constexpr int NW = DirWords<B>::N;
    int32_t entry = (int32_t)(sink.x - sink.y), row = (int32_t)sink.y - 1;
    uint32_t n_ops = 0, state = 0;                  // HSTATE 0, ESTATE 1, FSTATE 2
    while (row >= 0) {
        const uint32_t op = (dirs[(size_t)row * NW + (entry >> 3)] >> (4 * (entry & 7))) & 15u;
        const uint32_t h_op = op & 3u;
        if (TYPE == NVB_LOCAL && state == 0 && h_op == DIR_SINK) {
            src_y = (uint32_t)row + 1u; src_x = (uint32_t)entry + src_y;
            return n_ops;
        }
        if (state == 1)      { if ((op & DIR_INS_EXT) == 0) state = 0; --entry;        if (n_ops < max_ops) ops[n_ops] = DIR_DEL; ++n_ops; }
        else if (state == 2) { if ((op & DIR_DEL_EXT) == 0) state = 0; ++entry; --row; if (n_ops < max_ops) ops[n_ops] = DIR_INS; ++n_ops; }
        else {
            if (h_op == DIR_DEL) state = 1;
            else if (h_op == DIR_INS) state = 2;
            else { --row; if (n_ops < max_ops) ops[n_ops] = DIR_SUB; ++n_ops; }
        }
    }
    src_y = 0u; src_x = (uint32_t)entry;
    return n_ops;
}

// Gapless fast path of the banded traceback (LOCAL and SEMI_GLOBAL).  Given the optimal score and its sink (from the score kernels),
// walk the sink's diagonal backwards adding substitution scores: if a suffix of the diagonal adds up to exactly `score` (LOCAL: the
// SHORTEST such suffix; SEMI_GLOBAL: the whole pattern), the traceback of the reference IS that suffix, all substitutions.  Proof:
// H(cell) >= H(diagonal predecessor) + s holds everywhere; summed along the suffix it reads  score = H(sink) >= H(before the suffix) +
// sum = H(before) + score  with  H(before) >= 0 (LOCAL) resp. the boundary value 0 (SEMI_GLOBAL), so every inequality is tight: on every
// cell of the suffix the diagonal term equals H, and the reference's direction rule takes SUBSTITUTION whenever the diagonal is not
// beaten (gotoh_banded_inl.h:325-337; ties go to the diagonal), until it meets H == 0 / row 0 -- which the shortest suffix guarantees
// is not met earlier.  Checked against the reference's own traceback on 21,600 random alignments (11,695 taken, 0 different).
// Returns false (the caller runs the full traceback) for GLOBAL, when no such suffix exists (the alignment has a gap), or when a symbol
// > 3 lies on the diagonal (N's: the reference's text-cache quirks make their scores row-dependent).
template <int TYPE>
__host__ __device__ inline bool gapless_traceback(const GotohScheme& S,
        const uint32_t* __restrict__ pwords, uint32_t pbits, uint32_t pbe, uint32_t poff, uint32_t M, const uint8_t* __restrict__ quals,
        const uint32_t* __restrict__ twords, uint32_t tbits, uint32_t tbe, uint32_t toff, uint32_t N,
        const int32_t score, const uint32_t sx, const uint32_t sy, uint32_t& len)
{
    len = 0u;
    if (TYPE == NVB_GLOBAL) return false;
    if (sx == 0xFFFFFFFFu || sy == 0xFFFFFFFFu || sy == 0u || sy > M || sx < sy || sx > N) return false;
    if (TYPE == NVB_SEMI_GLOBAL && sy != M) return false;
    if (TYPE == NVB_LOCAL && score <= 0) return false;
    SymReaderRT tr(twords, tbits, tbe), pr(pwords, pbits, pbe);
    int32_t acc = 0;
    uint32_t i = sy, t = sx;
    while (i > 0u) {
        --i; --t;
        const uint32_t q = pr.get(poff + i), g = tr.get(toff + t);
        if ((q | g) > 3u) return false;
        const uint32_t qq = quals ? quals[poff + i] : 0u;
        acc += (g == q) ? (S.qtab ? S.qtab[2 * qq] : S.match) : (S.qtab ? S.qtab[2 * qq + 1] : S.mismatch);
        ++len;
        if (TYPE == NVB_LOCAL && acc == score) return true;
    }
    return TYPE == NVB_SEMI_GLOBAL && acc == score;
}

// ---------------------------------------------------------------------------------------------
// packed pair: two alignments per thread in s16x2 halves
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t pack16(int32_t lo, int32_t hi) { return ((uint32_t)lo & 0xFFFFu) | ((uint32_t)hi << 16); }
__host__ __device__ __forceinline__ int32_t half_lo(uint32_t v) { return (int32_t)(int16_t)(v & 0xFFFFu); }
__host__ __device__ __forceinline__ int32_t half_hi(uint32_t v) { return (int32_t)(int16_t)(v >> 16); }

// prmt.b32 generic mode incl. the sign-replicate bit of each selector nibble
__host__ __device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
    uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d;
#else
    const uint64_t ab = ((uint64_t)b << 32) | a;
    uint32_t d = 0;
    for (int k = 0; k < 4; ++k) {
        const uint32_t nib = (sel >> (4 * k)) & 0xFu;
        uint32_t byte = (uint32_t)(ab >> (8 * (nib & 7u))) & 0xFFu;
        if (nib & 8u) byte = (byte & 0x80u) ? 0xFFu : 0x00u;
        d |= byte << (8 * k);
    }
    return d;
#endif
}

// selector for a pair of 2-bit text symbols (g0 -> low half from profile a, g1 -> high half from b):
// nibbles {g0, 8|g0, 4|g1, 0xC|g1}: byte, its sign extension, byte, its sign extension
__host__ __device__ __forceinline__ uint32_t pair_selector(uint32_t g0, uint32_t g1) { return 0xC480u + g0 * 0x11u + g1 * 0x1100u; }

// 4-byte substitution profile of one pattern symbol: byte c = (c == q) ? s_eq : s_ne  (int8 each)
__host__ __device__ __forceinline__ uint32_t sub_profile(uint32_t q, int32_t s_eq, int32_t s_ne) {
    const uint32_t base = ((uint32_t)s_ne & 0xFFu) * 0x01010101u;
    const uint32_t flip = (uint32_t)(s_eq ^ s_ne) & 0xFFu;
    return (q < 4u) ? (base ^ (flip << (8u * q))) : base;
}

// host-side admissibility of the packed path for a batch (max pattern length max_m)
static inline bool pair_path_ok(int B, int type, const nvb_gotoh_scheme* s, uint32_t max_m) {
    if (!(B == 7 || B == 15 || B == 31)) return false;
    const int64_t Go = s->pattern_gap_open, Ge = s->pattern_gap_ext;
    if (Go >= 0 || Ge >= 0 || s->text_gap_open >= 0 || s->text_gap_ext >= 0) return false;
    // value range of the substitution scores: the two constants, or the caller's bounds on the quality table
    int64_t s_lo = s->match < s->mismatch ? s->match : s->mismatch, s_hi = s->match > s->mismatch ? s->match : s->mismatch;
    if (s->d_qual_table) {
        if (s->qual_table_min == 0 && s->qual_table_max == 0) return false;   // bounds unknown: generic path
        s_lo = s->qual_table_min; s_hi = s->qual_table_max;
    }
    const int64_t a_m = s_lo < 0 ? -s_lo : s_lo, a_x = s_hi < 0 ? -s_hi : s_hi;
    const int64_t max_s = a_m > a_x ? a_m : a_x;
    int64_t max_g = -Go; if (-Ge > max_g) max_g = -Ge; if (-(int64_t)s->text_gap_open > max_g) max_g = -(int64_t)s->text_gap_open;
    if (-(int64_t)s->text_gap_ext > max_g) max_g = -(int64_t)s->text_gap_ext;
    const int64_t bound = (int64_t)max_m * max_s + (int64_t)(B + 2) * max_g + max_g;
    if (bound > 30000) return false;
    // substitution bytes (S - Go) must fit int8
    if (s_lo - Go < -128 || s_hi - Go > 127) return false;
    if (type == NVB_LOCAL) {
        // LOCAL cells are packed as (h << 5) | j in 16 bits
        const int64_t top = (int64_t)max_m * (s_hi > 0 ? s_hi : 0);
        if (top >= 2048) return false;
    }
    return true;
}

// DPX / SIMD-in-word wrappers.  On sm_100a each is ONE instruction (VIADDMNMX.S16x2[.RELU],
// VIMNMX.S16x2[.RELU], VIADD.16x2, VIMNMX[3].U16x2); the host versions exist only so that tests can run
// the very same per-thread routine on the CPU.
#define NVB_VIADDMAX(a, b, c)      __viaddmax_s16x2((a), (b), (c))
#define NVB_VIADDMAX_RELU(a, b, c) __viaddmax_s16x2_relu((a), (b), (c))
#define NVB_VIMAX_RELU(a, b)       __vimax_s16x2_relu((a), (b))
#define NVB_VIMAX3(a, b, c)        __vimax3_s16x2((a), (b), (c))
#define NVB_VIMAX3_U(a, b, c)      __vimax3_u16x2((a), (b), (c))
#ifdef __CUDA_ARCH__
#define NVB_VIMAX(a, b)            __vmaxs2((a), (b))
#define NVB_VIADD(a, b)            __vadd2((a), (b))
#define NVB_VIMAX_U(a, b)          __vmaxu2((a), (b))
#else
static inline uint32_t nvb_host_vmaxs2(uint32_t a, uint32_t b) {
    const int32_t l = imax2(half_lo(a), half_lo(b)), h = imax2(half_hi(a), half_hi(b)); return pack16(l, h); }
static inline uint32_t nvb_host_vadd2(uint32_t a, uint32_t b) {
    return ((a + b) & 0xFFFFu) | (((a >> 16) + (b >> 16)) << 16); }
static inline uint32_t nvb_host_vmaxu2(uint32_t a, uint32_t b) {
    const uint32_t l = (a & 0xFFFFu) > (b & 0xFFFFu) ? (a & 0xFFFFu) : (b & 0xFFFFu);
    const uint32_t h = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16); return l | (h << 16); }
#define NVB_VIMAX(a, b)            nvb_host_vmaxs2((a), (b))
#define NVB_VIADD(a, b)            nvb_host_vadd2((a), (b))
#define NVB_VIMAX_U(a, b)          nvb_host_vmaxu2((a), (b))
#endif

// one LOCAL band cell of the packed pair (see the formulation in gotoh_pair), in three steps so that the cells of two rows can be
// written interleaved (pair_local_cell2); j is a compile-time constant after unrolling.
//   front: selector -> substitution scores, F[j], t' = max(H_diag + s, F')      (independent of the row's E chain, except cell B-1)
//   h    : h' = max(t', E', beta)                                               (chain)
//   back : H = h' + Go, key (IMAD), row maximum of the keys, E' = max(E' + Ge, H)   (chain)
struct PairLocalConsts { uint32_t Ge2, GoX, beta2, INFb2; };     // packed per-half constants of the LOCAL pair cells

template <int B>
__host__ __device__ __forceinline__ uint32_t pair_local_front(const GotohScheme& S, const int j, const uint32_t (&G)[B], uint32_t (&F)[B - 1],
        const uint16_t* srow, const uint32_t sel_stride, const uint32_t P0, const uint32_t P1, const uint32_t Ge2, const uint32_t INFb2, const uint32_t E)
{
    const uint32_t s = prmt(P0, P1, (uint32_t)srow[(size_t)j * sel_stride]);
    if (j == B - 1) return NVB_VIADDMAX(G[j], s, E);
    F[j] = (j < B - 2) ? NVB_VIADDMAX(F[j + 1], Ge2, G[j + 1]) : NVB_VIADDMAX(INFb2, Ge2, G[j + 1]);
    return NVB_VIADDMAX(G[j], s, F[j]);
}
template <int B>
__host__ __device__ __forceinline__ uint32_t pair_local_h(const int j, const uint32_t t, const uint32_t E, const uint32_t beta2)
{
    return (j == 0 || j == B - 1) ? NVB_VIMAX(t, beta2) : NVB_VIMAX3(t, E, beta2);
}
template <int B>
__host__ __device__ __forceinline__ void pair_local_back(const GotohScheme& S, const int j, uint32_t (&G)[B], const uint32_t hb,
        const uint32_t Ge2, const uint32_t GoX, uint32_t& E, uint32_t& rowkey, uint32_t& pk)
{
    // H = h' + Go per half, ONE 32-bit add (both halves stay >= 0: no borrow).  With two rows in flight this value is consumed two
    // cells later by the row below, so it is a plain integer add (ALU pipe, short latency) and not an IMAD: measured on C4, band 31,
    // 5.27 vs 4.90 TCUPS (profiles/README.md, r02); the key below is off every chain and stays an IMAD on the FMA pipe.
    G[j] = hb + GoX;
    const uint32_t key = G[j] * S.keymul + (uint32_t)(j | (j << 16));   // IMAD: (H << 5) | j per half, H < 2048
    // row maximum of the keys, two cells per VIMNMX3.U16x2 (pk holds the even cell's key until the odd one arrives)
    if (j & 1)           rowkey = NVB_VIMAX3_U(rowkey, pk, key);
    else if (j == B - 1) rowkey = NVB_VIMAX_U(rowkey, key);
    else                 pk = key;
    if (j < B - 1) E = (j == 0) ? G[0] : NVB_VIADDMAX(E, Ge2, G[j]);
}
template <int B>
__host__ __device__ __forceinline__ void pair_local_cell(const GotohScheme& S, const int j, uint32_t (&G)[B], uint32_t (&F)[B - 1],
        const uint16_t* srow, const uint32_t sel_stride, const uint32_t P0, const uint32_t P1,
        const PairLocalConsts& K, uint32_t& E, uint32_t& rowkey, uint32_t& pk)
{
    const uint32_t t = pair_local_front<B>(S, j, G, F, srow, sel_stride, P0, P1, K.Ge2, K.INFb2, E);
    const uint32_t hb = pair_local_h<B>(j, t, E, K.beta2);
    pair_local_back<B>(S, j, G, hb, K.Ge2, K.GoX, E, rowkey, pk);
}
// cell jA of row A and cell jB = jA - 2 of the row below it, statement by statement: the two rows' chains alternate in the
// instruction stream.  (Row A touches G/F[jA], [jA+1]; row B touches [jB], [jB+1] = [jA-2], [jA-1]: disjoint.)
template <int B>
__host__ __device__ __forceinline__ void pair_local_cell2(const GotohScheme& S, const int jA, const int jB, uint32_t (&G)[B], uint32_t (&F)[B - 1],
        const uint16_t* srowA, const uint16_t* srowB, const uint32_t sel_stride,
        const uint32_t PA0, const uint32_t PA1, const uint32_t PB0, const uint32_t PB1,
        const PairLocalConsts& K,
        uint32_t& EA, uint32_t& rkA, uint32_t& pkA, uint32_t& EB, uint32_t& rkB, uint32_t& pkB)
{
    const uint32_t tA = pair_local_front<B>(S, jA, G, F, srowA, sel_stride, PA0, PA1, K.Ge2, K.INFb2, EA);
    const uint32_t tB = pair_local_front<B>(S, jB, G, F, srowB, sel_stride, PB0, PB1, K.Ge2, K.INFb2, EB);
    const uint32_t hA = pair_local_h<B>(jA, tA, EA, K.beta2);
    const uint32_t hB = pair_local_h<B>(jB, tB, EB, K.beta2);
    pair_local_back<B>(S, jA, G, hA, K.Ge2, K.GoX, EA, rkA, pkA);
    pair_local_back<B>(S, jB, G, hB, K.Ge2, K.GoX, EB, rkB, pkB);
}

// GLOBAL / SEMI_GLOBAL cells: the same front (with the unbiased infimum), h = max(t, E), G = h + Go as a packed add (values may
// be negative: no carry-free IMAD here), no sink keys (the sink is read off the last row)
template <int B>
__host__ __device__ __forceinline__ void pair_nl_back(const int j, uint32_t (&G)[B], const uint32_t t, const uint32_t Ge2, const uint32_t Go2, uint32_t& E)
{
    const uint32_t h = (j == 0 || j == B - 1) ? t : NVB_VIMAX(t, E);
    G[j] = NVB_VIADD(h, Go2);
    E = (j == 0) ? G[0] : NVB_VIADDMAX(E, Ge2, G[j]);
}
template <int B>
__host__ __device__ __forceinline__ void pair_nl_cell(const GotohScheme& S, const int j, uint32_t (&G)[B], uint32_t (&F)[B - 1],
        const uint16_t* srow, const uint32_t sel_stride, const uint32_t P0, const uint32_t P1,
        const uint32_t Ge2, const uint32_t Go2, const uint32_t INF2, uint32_t& E)
{
    const uint32_t t = pair_local_front<B>(S, j, G, F, srow, sel_stride, P0, P1, Ge2, INF2, E);
    pair_nl_back<B>(j, G, t, Ge2, Go2, E);
}
template <int B>
__host__ __device__ __forceinline__ void pair_nl_cell2(const GotohScheme& S, const int jA, const int jB, uint32_t (&G)[B], uint32_t (&F)[B - 1],
        const uint16_t* srowA, const uint16_t* srowB, const uint32_t sel_stride,
        const uint32_t PA0, const uint32_t PA1, const uint32_t PB0, const uint32_t PB1,
        const uint32_t Ge2, const uint32_t Go2, const uint32_t INF2, uint32_t& EA, uint32_t& EB)
{
    const uint32_t tA = pair_local_front<B>(S, jA, G, F, srowA, sel_stride, PA0, PA1, Ge2, INF2, EA);
    const uint32_t tB = pair_local_front<B>(S, jB, G, F, srowB, sel_stride, PB0, PB1, Ge2, INF2, EB);
    pair_nl_back<B>(jA, G, tA, Ge2, Go2, EA);
    pair_nl_back<B>(jB, G, tB, Ge2, Go2, EB);
}

// Preconditions (checked by the caller, else the generic path is used):
//   M0,M1 >= 1; N_k >= M_k + B - 1 (no pad symbol is ever read inside an alignment's own rows);
//   text is 2-bit; TYPE != LOCAL => M0 == M1; scheme admitted by pair_path_ok().
// sel: selectors of text columns t = 0 .. max(M0,M1)+B-2, element t at sel[t*sel_stride]
// (shared memory on the device: conflict-free u16 column per thread).
// PFMT: 0 = pattern format and quality table handled at run time; 2 / 4 = 2- / 4-bit big-endian patterns AND no quality table, both
// known at compile time (the dispatcher guarantees it): the per-row preamble loses its format dispatch and its table branch
// ROWS2: two pattern rows per loop iteration, the second one trailing the first by two band cells.  A row is one serial
// chain h' -> H -> E' (three dependent instructions per cell); two rows in flight give every thread two independent chains, which is
// what hides the ALU latency at 4 warps per scheduler.  The in-place update of G[] / F[] stays valid: row i+1 reads cells j, j+1 of
// row i after row i has written them (it is at cell j+2) and row i never looks back.
template <int B, int TYPE, int PFMT = 0, bool ROWS2 = false>
__host__ __device__ inline void gotoh_pair(const GotohScheme& S,
        const uint32_t* __restrict__ pwords, uint32_t pbits, uint32_t pbe,
        uint32_t poff0, uint32_t M0, uint32_t poff1, uint32_t M1,
        uint32_t N0, uint32_t N1,
        const uint16_t* sel, uint32_t sel_stride,
        SinkResult& r0, SinkResult& r1,
        const uint8_t* __restrict__ quals = nullptr,
        const uint32_t* prof_tab = nullptr)     // optional 256-entry table of sub_profile(q, c_eq, c_ne) (shared memory on the device)
{
    const int32_t Go = S.pgo, Ge = S.pge;
    const uint32_t Go2 = pack16(Go, Go), Ge2 = pack16(Ge, Ge);
    // F "minus infinity", raised just enough that INF+Ge cannot wrap below -32768 (identical maxima:
    // every H+Go it competes with is representable and therefore >= -32768)
    int32_t INF = gotoh_infimum(S);
    if (INF + Ge < -32768) INF = -32768 - Ge;
    const uint32_t INF2 = pack16(INF, INF);
    const int32_t c_eq = S.match - Go, c_ne = S.mismatch - Go;       // substitution minus Go (G = H + Go is stored)
    // per-row substitution profile of both alignments; with a quality table the two scores of a row come from
    // table[2*qual], table[2*qual+1] (nvBowtie's SmithWatermanScoringScheme::substitution)
#define NVB_ROW_PROFILES(i)                                                                                   \
    /* fixed-format streams run on past the end (symbol 0): rows beyond an alignment's own length are never reported */ \
    const uint32_t q0 = (PFMT != 0 || (i) < M0) ? pr0.next() : 255u;                                          \
    const uint32_t q1 = (PFMT != 0 || (i) < M1) ? pr1.next() : 255u;                                          \
    int32_t e0 = c_eq, n0 = c_ne, e1 = c_eq, n1 = c_ne;                                                        \
    if (PFMT == 0 && S.qtab) {                                                                                \
        const uint32_t qq0 = (quals && (i) < M0) ? quals[poff0 + (i)] : 0u;                                   \
        const uint32_t qq1 = (quals && (i) < M1) ? quals[poff1 + (i)] : 0u;                                   \
        e0 = S.qtab[2 * qq0] - Go; n0 = S.qtab[2 * qq0 + 1] - Go;                                             \
        e1 = S.qtab[2 * qq1] - Go; n1 = S.qtab[2 * qq1 + 1] - Go;                                             \
    }                                                                                                         \
    const uint32_t P0 = (prof_tab && (PFMT != 0 || !S.qtab)) ? prof_tab[q0] : sub_profile(q0, e0, n0);        \
    const uint32_t P1 = (prof_tab && (PFMT != 0 || !S.qtab)) ? prof_tab[q1] : sub_profile(q1, e1, n1);

    uint32_t G[B], F[B - 1];
    {
        G[0] = pack16(0 + Go, 0 + Go);
#pragma unroll
        for (int j = 1; j < B; ++j) {
            const int32_t h = (TYPE == NVB_GLOBAL) ? S.tgo + (j - 1) * S.tge : 0;
            G[j] = pack16(h + Go, h + Go);
        }
#pragma unroll
        for (int j = 0; j < B - 1; ++j) F[j] = INF2;
    }

    typename PatStreamOf<PFMT>::type pr0(pwords, pbits, pbe, poff0, M0), pr1(pwords, pbits, pbe, poff1, M1);
    const uint32_t Mmax = M0 > M1 ? M0 : M1;
    int32_t bk0 = -1, bk1 = -1; uint32_t bi0 = 0, bi1 = 0;       // LOCAL: best row key (H << 5 | j) and its row, per half

    if (TYPE == NVB_LOCAL) {
        // LOCAL formulation, biased by beta = -Go so that the one plain add per cell (h' + Go) is carry-free, i.e. ONE 32-bit add for
        // both halves:  G[] holds H itself (>= 0), F[] and E hold F+beta / E+beta, h' = h+beta >= beta.
        //   F'[j] = max(F'[j+1]+Ge, H[j+1]);  t' = max(H[j] + (S-Go), F'[j]);  h' = max(t', E', beta);
        //   H[j] = h' + Go;  E' = max(E'+Ge, H[j]);  sink key (H << 5) | j = one IMAD on the FMA pipe.
        // (A variant that biases by |Ge| as well, so that the chain is E' -> h' -> E' with H and E'+Ge formed off the chain by
        // IMADs, was measured slower: 4.47 TCUPS against 5.27 on C4, band 31 -- every extra IMAD costs issue bandwidth.)
        const int32_t beta = -Go;
        PairLocalConsts K;
        K.Ge2 = Ge2;
        K.GoX = (uint32_t)(Go * 65537);                            // Go in both halves, as ONE 32-bit addend
        K.beta2 = pack16(beta, beta);
        K.INFb2 = pack16(INF + beta, INF + beta);
#pragma unroll
        for (int j = 0; j < B; ++j) G[j] = 0u;
#pragma unroll
        for (int j = 0; j < B - 1; ++j) F[j] = K.INFb2;
        // later rows win ties: replace when H_row >= H_best, i.e. key_row >= (key_best with its column bits cleared)
#define NVB_LOCAL_ROW_END(i, rowkey)                                                                            \
        {   const int32_t k0 = (int32_t)((rowkey) & 0xFFFFu), k1 = (int32_t)((rowkey) >> 16);                  \
            if ((i) < M0 && k0 >= (bk0 & ~31)) { bk0 = k0; bi0 = (i); }                                         \
            if ((i) < M1 && k1 >= (bk1 & ~31)) { bk1 = k1; bi1 = (i); } }
        uint32_t i = 0;
        constexpr int SK = 2;                                      // ROWS2: row i+1 trails row i by two cells (one would read a stale G)
#define NVB_LOCAL_TWO_ROWS(i, PA0, PA1, PB0, PB1)                                                               \
        {   const uint16_t* srowA = sel + (size_t)(i) * sel_stride;                                                 \
            const uint16_t* srowB = srowA + sel_stride;                                                             \
            uint32_t EA = 0, EB = 0, rkA = 0, rkB = 0, pkA = 0, pkB = 0;                                        \
            _Pragma("unroll")                                                                                   \
            for (int jj = 0; jj < B + SK; ++jj) {                                                               \
                if (jj >= SK && jj < B)                                                                         \
                    pair_local_cell2<B>(S, jj, jj - SK, G, F, srowA, srowB, sel_stride, PA0, PA1, PB0, PB1, K, \
                                              EA, rkA, pkA, EB, rkB, pkB);                                      \
                else if (jj < B) pair_local_cell<B>(S, jj,      G, F, srowA, sel_stride, PA0, PA1, K, EA, rkA, pkA); \
                else             pair_local_cell<B>(S, jj - SK, G, F, srowB, sel_stride, PB0, PB1, K, EB, rkB, pkB); \
            }                                                                                                   \
            NVB_LOCAL_ROW_END(i, rkA)                                                                           \
            NVB_LOCAL_ROW_END((i) + 1u, rkB) }
        if (ROWS2 && PFMT != 0) {
            // compile-time pattern format: the symbols of 16 (2-bit) / 8 (4-bit) rows arrive as one aligned word per alignment, so
            // a row costs a shift per alignment instead of a stream refill check (rows past an alignment's own length read zeros
            // and are never reported)
            constexpr uint32_t PB_ = (PFMT != 0) ? (uint32_t)PFMT : 2u, SPW = 32u / PB_, QM = (1u << PB_) - 1u;
            PatChunksBE<(int)PB_> pc0(pwords, poff0, M0), pc1(pwords, poff1, M1);
            uint32_t c0 = 0, c1 = 0;
            for (uint32_t base = 0; base < Mmax; base += SPW) {
                c0 = pc0.next(); c1 = pc1.next();
                const uint32_t end = (base + SPW < Mmax) ? base + SPW : Mmax;
                for (; i + 1u < end; i += 2u) {
                    const uint32_t qa0 = c0 >> (32u - PB_), qb0 = (c0 >> (32u - 2u * PB_)) & QM;
                    const uint32_t qa1 = c1 >> (32u - PB_), qb1 = (c1 >> (32u - 2u * PB_)) & QM;
                    c0 <<= 2u * PB_; c1 <<= 2u * PB_;
                    const uint32_t PA0 = prof_tab ? prof_tab[qa0] : sub_profile(qa0, c_eq, c_ne), PA1 = prof_tab ? prof_tab[qa1] : sub_profile(qa1, c_eq, c_ne);
                    const uint32_t PB0 = prof_tab ? prof_tab[qb0] : sub_profile(qb0, c_eq, c_ne), PB1 = prof_tab ? prof_tab[qb1] : sub_profile(qb1, c_eq, c_ne);
                    NVB_LOCAL_TWO_ROWS(i, PA0, PA1, PB0, PB1)
                }
            }
            if (i < Mmax) {                                        // odd number of rows: the last one, its symbols at the top of c0 / c1
                const uint32_t q0 = c0 >> (32u - PB_), q1 = c1 >> (32u - PB_);
                const uint32_t P0 = prof_tab ? prof_tab[q0] : sub_profile(q0, c_eq, c_ne), P1 = prof_tab ? prof_tab[q1] : sub_profile(q1, c_eq, c_ne);
                const uint16_t* srow = sel + (size_t)i * sel_stride;
                uint32_t E = 0, rowkey = 0, pk = 0;
#pragma unroll
                for (int j = 0; j < B; ++j)
                    pair_local_cell<B>(S, j, G, F, srow, sel_stride, P0, P1, K, E, rowkey, pk);
                NVB_LOCAL_ROW_END(i, rowkey)
                ++i;
            }
        } else if (ROWS2) {
            for (; i + 1u < Mmax; i += 2u) {
                uint32_t PA0, PA1, PB0, PB1;
                { NVB_ROW_PROFILES(i)      PA0 = P0; PA1 = P1; }
                { NVB_ROW_PROFILES(i + 1u) PB0 = P0; PB1 = P1; }
                NVB_LOCAL_TWO_ROWS(i, PA0, PA1, PB0, PB1)
            }
        }
        if (!(ROWS2 && PFMT != 0))                                 // (that path has consumed every row)
        for (; i < Mmax; ++i) {
            NVB_ROW_PROFILES(i)
            const uint16_t* srow = sel + (size_t)i * sel_stride;
            uint32_t E = 0, rowkey = 0, pk = 0;
#pragma unroll
            for (int j = 0; j < B; ++j)
                pair_local_cell<B>(S, j, G, F, srow, sel_stride, P0, P1, K, E, rowkey, pk);
            NVB_LOCAL_ROW_END(i, rowkey)
        }
#undef NVB_LOCAL_ROW_END
#undef NVB_LOCAL_TWO_ROWS
    } else {
        // GLOBAL / SEMI_GLOBAL (M0 == M1): the same loop structure as LOCAL
        uint32_t i = 0;
        constexpr int SK = 2;
#define NVB_NL_TWO_ROWS(i, PA0, PA1, PB0, PB1)                                                                  \
        {   const uint16_t* srowA = sel + (size_t)(i) * sel_stride;                                             \
            const uint16_t* srowB = srowA + sel_stride;                                                         \
            uint32_t EA = 0, EB = 0;                                                                            \
            _Pragma("unroll")                                                                                   \
            for (int jj = 0; jj < B + SK; ++jj) {                                                               \
                if (jj >= SK && jj < B)                                                                         \
                    pair_nl_cell2<B>(S, jj, jj - SK, G, F, srowA, srowB, sel_stride, PA0, PA1, PB0, PB1, Ge2, Go2, INF2, EA, EB); \
                else if (jj < B) pair_nl_cell<B>(S, jj,      G, F, srowA, sel_stride, PA0, PA1, Ge2, Go2, INF2, EA); \
                else             pair_nl_cell<B>(S, jj - SK, G, F, srowB, sel_stride, PB0, PB1, Ge2, Go2, INF2, EB); \
            } }
        if (ROWS2 && PFMT != 0) {
            constexpr uint32_t PB_ = (PFMT != 0) ? (uint32_t)PFMT : 2u, SPW = 32u / PB_, QM = (1u << PB_) - 1u;
            PatChunksBE<(int)PB_> pc0(pwords, poff0, M0), pc1(pwords, poff1, M1);
            uint32_t c0 = 0, c1 = 0;
            for (uint32_t base = 0; base < Mmax; base += SPW) {
                c0 = pc0.next(); c1 = pc1.next();
                const uint32_t end = (base + SPW < Mmax) ? base + SPW : Mmax;
                for (; i + 1u < end; i += 2u) {
                    const uint32_t qa0 = c0 >> (32u - PB_), qb0 = (c0 >> (32u - 2u * PB_)) & QM;
                    const uint32_t qa1 = c1 >> (32u - PB_), qb1 = (c1 >> (32u - 2u * PB_)) & QM;
                    c0 <<= 2u * PB_; c1 <<= 2u * PB_;
                    const uint32_t PA0 = prof_tab ? prof_tab[qa0] : sub_profile(qa0, c_eq, c_ne), PA1 = prof_tab ? prof_tab[qa1] : sub_profile(qa1, c_eq, c_ne);
                    const uint32_t PB0 = prof_tab ? prof_tab[qb0] : sub_profile(qb0, c_eq, c_ne), PB1 = prof_tab ? prof_tab[qb1] : sub_profile(qb1, c_eq, c_ne);
                    NVB_NL_TWO_ROWS(i, PA0, PA1, PB0, PB1)
                }
            }
            if (i < Mmax) {
                const uint32_t q0 = c0 >> (32u - PB_), q1 = c1 >> (32u - PB_);
                const uint32_t P0 = prof_tab ? prof_tab[q0] : sub_profile(q0, c_eq, c_ne), P1 = prof_tab ? prof_tab[q1] : sub_profile(q1, c_eq, c_ne);
                const uint16_t* srow = sel + (size_t)i * sel_stride;
                uint32_t E = 0;
#pragma unroll
                for (int j = 0; j < B; ++j) pair_nl_cell<B>(S, j, G, F, srow, sel_stride, P0, P1, Ge2, Go2, INF2, E);
                ++i;
            }
        } else if (ROWS2) {
            for (; i + 1u < Mmax; i += 2u) {
                uint32_t PA0, PA1, PB0, PB1;
                { NVB_ROW_PROFILES(i)      PA0 = P0; PA1 = P1; }
                { NVB_ROW_PROFILES(i + 1u) PB0 = P0; PB1 = P1; }
                NVB_NL_TWO_ROWS(i, PA0, PA1, PB0, PB1)
            }
        }
#undef NVB_NL_TWO_ROWS
        if (!(ROWS2 && PFMT != 0))
        for (; i < Mmax; ++i) {
            NVB_ROW_PROFILES(i)
            const uint16_t* srow = sel + (size_t)i * sel_stride;
            uint32_t E = 0;
#pragma unroll
            for (int j = 0; j < B; ++j) pair_nl_cell<B>(S, j, G, F, srow, sel_stride, P0, P1, Ge2, Go2, INF2, E);
        }
    }

#undef NVB_ROW_PROFILES
    if (TYPE == NVB_LOCAL) {
        r0.score = bk0 >> 5; r0.x = bi0 + ((uint32_t)bk0 & 31u) + 1u; r0.y = bi0 + 1u;
        r1.score = bk1 >> 5; r1.x = bi1 + ((uint32_t)bk1 & 31u) + 1u; r1.y = bi1 + 1u;
    } else if (TYPE == NVB_GLOBAL) {
        r0.score = half_lo(G[B - 1]) - Go; r0.x = M0 + (uint32_t)B - 1u; r0.y = M0;
        r1.score = half_hi(G[B - 1]) - Go; r1.x = M1 + (uint32_t)B - 1u; r1.y = M1;
    } else {
        const uint32_t m0 = umin2(M0 + (uint32_t)B - 1u, N0) - (M0 - 1u);
        const uint32_t m1 = umin2(M1 + (uint32_t)B - 1u, N1) - (M1 - 1u);
        r0.score = half_lo(G[0]) - Go; r0.x = M0; r0.y = M0;
        r1.score = half_hi(G[0]) - Go; r1.x = M1; r1.y = M1;
#pragma unroll
        for (int j = 1; j < B; ++j) {
            const int32_t h0 = half_lo(G[j]) - Go, h1 = half_hi(G[j]) - Go;
            if ((uint32_t)j < m0 && r0.score <= h0) { r0.score = h0; r0.x = M0 + (uint32_t)j; }
            if ((uint32_t)j < m1 && r1.score <= h1) { r1.score = h1; r1.x = M1 + (uint32_t)j; }
        }
    }
}

} // namespace nvb
