// nvbio_b200/shim/nvbowtie_scoring.h -- nvBowtie's best-score extension stage (boundary B-B2) over the B200 DP kernels.
//
// nvBowtie scores its queued seed hits with
//     detail::banded_score_best(band_len, pipeline, aligner, params)   ->  BatchedBandedAlignmentScore<B, BestScoreStream<aligner,pipeline>,
//                                                                          DeviceThreadScheduler>::enact          score_best_inl.h:53-200
// called from score_best_t / the non-template score_best (score_best_inl.h:212-234, score.h:52-60).  Including this header BEFORE those
// templates are instantiated binds BestScoreStream to the engine of batched_alignment.h:
//   * patterns: nvBowtie reads a read reversed or complemented through its ReadLoader (alignment_utils.h:197-220): the engine
//     materialises pattern symbols and base qualities through the stream's OWN load_strings();
//   * texts: the window [genome_begin, genome_end) of the 2-bit genome stream, read in place;
//   * scheme: SmithWatermanScoringScheme<MMCost,NCost> (scoring.h:203-317) becomes the 256 x 2 substitution table of the C ABI, filled
//     ON THE DEVICE by the scheme object's own substitution() so that the float arithmetic of QualCost (scoring.h:96-100) runs under
//     exactly the flags the reference kernel would run it with;
//   * init_context() / output() stay nvBowtie's: hit queues, min_score, hit.score = max(score, worst_score), hit.sink (:94-145).
#pragma once

#include <nvbio_b200/shim/batched_alignment.h>
#include <nvBowtie/bowtie2/cuda/scoring.h>
#include <nvBowtie/bowtie2/cuda/alignment_utils.h>
#include <nvBowtie/bowtie2/cuda/score_best_inl.h>

namespace nvbio {
namespace aln {
namespace b200 {

template <typename scheme_type>
__global__ void qual_table_kernel(const scheme_type scheme, int32* table)
{
    const uint32 q = threadIdx.x;                                            // one thread per base quality
    table[2u*q]      = scheme.substitution( 0u, 0u, uint8(1), uint8(1), uint8(q) );
    table[2u*q + 1u] = scheme.substitution( 0u, 0u, uint8(0), uint8(1), uint8(q) );
}

template <AlignmentType TYPE_T, typename MMCost, typename NCost, typename algorithm_tag>
struct gotoh_scheme_of< GotohAligner<TYPE_T, bowtie2::cuda::SmithWatermanScoringScheme<MMCost,NCost>, algorithm_tag> >
{
    typedef bowtie2::cuda::SmithWatermanScoringScheme<MMCost,NCost> scheme_type;
    static const bool supported = true;
    static const int  TYPE      = int(TYPE_T);
    static nvb_gotoh_scheme get(const GotohAligner<TYPE_T,scheme_type,algorithm_tag>& a)
    {
        int32* table = (int32*)nvbio::b200::scratch( nvbio::b200::SCRATCH_TABLE, 512u * sizeof(int32) );
        qual_table_kernel<<<1,256>>>( a.scheme, table );
        nvb_gotoh_scheme s;
        s.match = a.scheme.match( 0 );                         s.mismatch = a.scheme.mismatch( 0 );
        s.pattern_gap_open = a.scheme.pattern_gap_open();      s.pattern_gap_ext = a.scheme.pattern_gap_extension();
        s.text_gap_open    = a.scheme.text_gap_open();         s.text_gap_ext    = a.scheme.text_gap_extension();
        s.d_qual_table = table;
        // bounds of the table's values for the packed 16-bit path: the host evaluation, widened by one for float rounding differences
        int32 lo = 0, hi = 0;
        for (uint32 q = 0; q < 256u; ++q)
        {
            const int32 m = a.scheme.match( uint8(q) ), x = a.scheme.mismatch( uint8(q) );
            if (q == 0u) { lo = nvbio::min( m, x ); hi = nvbio::max( m, x ); }
            lo = nvbio::min( lo, nvbio::min( m, x ) ); hi = nvbio::max( hi, nvbio::max( m, x ) );
        }
        s.qual_table_min = lo - 1; s.qual_table_max = hi + 1;
        return s;
    }
};

template <typename AlignerType, typename PipelineType>
struct stream_binding< bowtie2::cuda::detail::BestScoreStream<AlignerType,PipelineType> >
{
    typedef bowtie2::cuda::detail::BestScoreStream<AlignerType,PipelineType>    stream_type;
    typedef typename stream_type::context_type                                  context_type;
    typedef typename PipelineType::genome_iterator                              genome_iterator;
    typedef nvbio::vector_view<genome_iterator>                                 text_string;

    static const bool bound = nvbio::b200::packed_iterator<genome_iterator>::supported;
    static const bool materialise_patterns = true;

    // (the stream's own max_pattern_length() does not compile: see binding_defaults)
    static uint32 max_pattern_length(const stream_type& s) { return s.m_pipeline.reads.max_sequence_len(); }
    static uint32 max_text_length   (const stream_type& s) { return s.m_pipeline.reads.max_sequence_len() + s.m_band_len; }

    static const uint32* text_words(const stream_type& s) { return nvbio::b200::packed_iterator<genome_iterator>::words( s.m_pipeline.genome ); }
    NVBIO_FORCEINLINE NVBIO_HOST_DEVICE static text_string text(const stream_type& s, const uint32 i, const context_type* context)
    {
        return text_string( context->genome_end - context->genome_begin, s.m_pipeline.genome + context->genome_begin );
    }
};

} // namespace b200

/// the batch class nvBowtie's banded_score_best instantiates (score_best_inl.h:160-200), every band length
template <uint32 BLOCKDIM, uint32 MINBLOCKS, uint32 BAND_LEN, typename AlignerType, typename PipelineType>
struct BatchedBandedAlignmentScore< BAND_LEN, bowtie2::cuda::detail::BestScoreStream<AlignerType,PipelineType>, DeviceThreadBlockScheduler<BLOCKDIM,MINBLOCKS> > :
    public b200::BandedScoreBatch< BAND_LEN, bowtie2::cuda::detail::BestScoreStream<AlignerType,PipelineType>, DeviceThreadBlockScheduler<BLOCKDIM,MINBLOCKS> > {};

} // namespace aln
} // namespace nvbio
