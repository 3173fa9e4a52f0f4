// oracle/ref_nvbowtie.cpp -- TEST INFRASTRUCTURE (part of oracle/_ref/libnvbio_ref.so): nvBowtie's OWN scoring scheme, compiled
// from /root/reference where it lies, so that the host-evaluated 256x2 substitution table of nvbio_b200.aln.QualityGotohScheme
// and the DP it drives are pinned against the real SmithWatermanScoringScheme<QualCost<int>,ConstantCost<int>>
// (nvBowtie/bowtie2/cuda/scoring.h:86-105 QualCost, :203-317 the scheme, scoring_inl.h:74-147 presets) instead of a model of it.
#include <nvbio/basic/types.h>
#include <nvbio/basic/vector_view.h>
#include <nvbio/alignment/alignment.h>
#include <nvbio/alignment/utils.h>
#include <nvBowtie/bowtie2/cuda/scoring.h>
#include <climits>
#include <omp.h>

using namespace nvbio;

namespace {
typedef bowtie2::cuda::SmithWatermanScoringScheme<>   scheme_t;     // <QualCost<int>, ConstantCost<int>>: nvBowtie's production scheme
typedef vector_view<const uint8*>                     str_view;

// preset 0: the given constants; 1: scheme_t::local() (nvBowtie --local defaults); 2: scheme_t() (the end-to-end defaults)
scheme_t make_scheme(int preset, int match_bonus, int mm_min, int mm_max, int read_gap_const, int read_gap_coeff, int ref_gap_const, int ref_gap_coeff)
{
    if (preset == 1) return scheme_t::local();
    scheme_t s;
    if (preset == 2) return s;
    s.m_match = scheme_t::MatchCost( match_bonus, match_bonus );
    s.m_mmp   = scheme_t::MismatchCost( mm_min, mm_max );
    s.m_read_gap_const = read_gap_const; s.m_read_gap_coeff = read_gap_coeff;
    s.m_ref_gap_const  = ref_gap_const;  s.m_ref_gap_coeff  = ref_gap_coeff;
    s.m_monotone = (match_bonus == 0);
    return s;
}

template <uint32 BAND, aln::AlignmentType TYPE>
void run(const scheme_t scheme, const uint8* pat, const uint8* qual, const uint32* p_off, const uint32* p_len,
         const uint8* txt, const uint32* t_off, const uint32* t_len, uint32 n, int32* score, uint32* sink_x, uint32* sink_y)
{
    #pragma omp parallel for schedule(static)
    for (int64 i = 0; i < int64(n); ++i)
    {
        aln::BestSink<int32> sink;
        aln::banded_alignment_score<BAND>(
            aln::make_gotoh_aligner<TYPE>( scheme ),
            str_view( p_len[i], pat + p_off[i] ),
            str_view( p_len[i], qual + p_off[i] ),
            str_view( t_len[i], txt + t_off[i] ),
            INT_MIN,
            sink );
        score[i] = sink.score; sink_x[i] = sink.sink.x; sink_y[i] = sink.sink.y;
    }
}
} // anonymous namespace

extern "C" {

// table512[2q] = substitution on a match, table512[2q+1] on a mismatch, for base quality q; gaps4 = pattern open/ext, text open/ext;
// limits2 = {worst_score, perfect_score(100)}
void ref_nvbowtie_scheme(int preset, int match_bonus, int mm_min, int mm_max, int read_gap_const, int read_gap_coeff, int ref_gap_const, int ref_gap_coeff,
                         int32* table512, int32* gaps4, int32* limits2)
{
    const scheme_t s = make_scheme( preset, match_bonus, mm_min, mm_max, read_gap_const, read_gap_coeff, ref_gap_const, ref_gap_coeff );
    for (uint32 q = 0; q < 256u; ++q)
    {
        table512[2u*q]      = s.substitution( 0u, 0u, uint8(1), uint8(1), uint8(q) );
        table512[2u*q + 1u] = s.substitution( 0u, 0u, uint8(0), uint8(1), uint8(q) );
    }
    gaps4[0] = s.pattern_gap_open(); gaps4[1] = s.pattern_gap_extension(); gaps4[2] = s.text_gap_open(); gaps4[3] = s.text_gap_extension();
    limits2[0] = scheme_t::worst_score; limits2[1] = s.perfect_score( 100u );
}

// banded DP (band 15 / 31; type 1 LOCAL, 2 SEMI_GLOBAL = nvBowtie's two modes) with the real scheme and per-base qualities
int ref_nvbowtie_banded(int band, int type, int preset, int match_bonus, int mm_min, int mm_max, int read_gap_const, int read_gap_coeff,
                        int ref_gap_const, int ref_gap_coeff,
                        const uint8* pat, const uint8* qual, const uint32* p_off, const uint32* p_len,
                        const uint8* txt, const uint32* t_off, const uint32* t_len, uint32 n, int32* score, uint32* sink_x, uint32* sink_y)
{
    const scheme_t s = make_scheme( preset, match_bonus, mm_min, mm_max, read_gap_const, read_gap_coeff, ref_gap_const, ref_gap_coeff );
    if (band == 15 && type == 1) { run<15,aln::LOCAL>      ( s, pat,qual,p_off,p_len, txt,t_off,t_len, n, score,sink_x,sink_y ); return 0; }
    if (band == 15 && type == 2) { run<15,aln::SEMI_GLOBAL>( s, pat,qual,p_off,p_len, txt,t_off,t_len, n, score,sink_x,sink_y ); return 0; }
    if (band == 31 && type == 1) { run<31,aln::LOCAL>      ( s, pat,qual,p_off,p_len, txt,t_off,t_len, n, score,sink_x,sink_y ); return 0; }
    if (band == 31 && type == 2) { run<31,aln::SEMI_GLOBAL>( s, pat,qual,p_off,p_len, txt,t_off,t_len, n, score,sink_x,sink_y ); return 0; }
    return -1;
}

} // extern "C"
