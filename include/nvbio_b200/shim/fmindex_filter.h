// nvbio_b200/shim/fmindex_filter.h -- nvbio::FMIndexFilter<device_tag, fm_index_type> for the production FM-index layout,
// forwarding rank() / locate() to the B200 kernels (nvb_fm_filter_rank / nvb_fm_filter_locate).
//
// Replaces the primary template's bodies at nvbio/fmindex/filter_inl.h:268-300 (rank: thrust::transform of
// fmindex::rank_functor + inclusive scan of the range sizes) and :306-402 (locate: filter_results -> locate_ssa_results ->
// lookup_ssa_results) for every
//     fm_index< rank_dictionary<2,64, PackedStream<deinterleaved_iterator<2,0,P>,uint8,2,true>, deinterleaved_iterator<2,1,P>, C>,
//               SSA_index_multiple_context<K,S> >
// i.e. io::FMIndexDataDevice::fm_index_type (nvbio/io/fmindex/fmindex.h:302-319) and its raw-pointer twins.  The class
// keeps the primary template's public interface (filter.h:145-214): rank, locate, n_hits, ranges, ranks and the typedefs
// FMIndexFilterDevice<fm_index_type> re-exports (filter.h:236-254).  Include this header BEFORE the first use of
// FMIndexFilterDevice<...> in a translation unit (a partial specialisation must be visible before it is instantiated).
// String sets whose storage the shim cannot see through (packed_string_set<...>::supported == false) take the
// reference's own functors, chosen at compile time.
#pragma once

#include <nvbio_b200/shim/views.h>
#include <nvbio/fmindex/filter.h>
#include <thrust/transform.h>
#include <thrust/scan.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <cstdio>

namespace nvbio {

template <typename BwtOccPtr, typename CountTable, uint32 SA_INT, typename SsaPtr>
struct FMIndexFilter<
    device_tag,
    fm_index<
        rank_dictionary< 2u, 64u,
            PackedStream< deinterleaved_iterator<2,0,BwtOccPtr>, uint8, 2u, true, uint32 >,
            deinterleaved_iterator<2,1,BwtOccPtr>,
            CountTable >,
        SSA_index_multiple_context<SA_INT,SsaPtr>,
        null_type > >
{
    typedef fm_index<
        rank_dictionary< 2u, 64u,
            PackedStream< deinterleaved_iterator<2,0,BwtOccPtr>, uint8, 2u, true, uint32 >,
            deinterleaved_iterator<2,1,BwtOccPtr>,
            CountTable >,
        SSA_index_multiple_context<SA_INT,SsaPtr>,
        null_type >                                         fm_index_type;

    typedef device_tag                                      system_tag;     ///< the backend system
    typedef fm_index_type                                   index_type;     ///< the index type

    typedef typename index_type::index_type                 coord_type;     ///< uint32
    static const uint32                                     coord_dim = vector_traits<coord_type>::DIM;
    typedef typename vector_type<coord_type,2>::type        range_type;     ///< uint2, inclusive SA range
    static const uint32                                     hit_dim = coord_dim*2;
    typedef typename vector_type<coord_type,hit_dim>::type  hit_type;       ///< uint2 = (text position, string id)

    typedef b200::fm_index_view<fm_index_type>              view_type;

    /// enact the filter on an FM-index and a string-set; returns the total number of hits
    template <typename string_set_type>
    uint64 rank(const fm_index_type& index, const string_set_type& string_set)
    {
        m_n_queries = string_set.size();
        m_index     = index;
        m_ranges.resize( m_n_queries );
        m_slots.resize( m_n_queries );
        if (m_n_queries == 0u)
            return m_n_occurrences = 0u;

        rank_dispatch( string_set, std::integral_constant<bool, view_type::supported && b200::packed_string_set<string_set_type>::supported>() );
        return m_n_occurrences;
    }

    /// enumerate the hits [begin,end) as (text position, string id) pairs
    template <typename hits_iterator>
    void locate(const uint64 begin, const uint64 end, hits_iterator hits)
    {
        if (end <= begin) return;
        locate_dispatch( begin, end, hits, std::integral_constant<bool, view_type::supported>() );
    }

    uint64            n_hits() const { return m_n_occurrences; }
    const range_type* ranges() const { return nvbio::plain_view( m_ranges ); }
    const uint64*     ranks()  const { return nvbio::plain_view( m_slots ); }

    uint32                              m_n_queries;
    index_type                          m_index;
    uint64                              m_n_occurrences;
    thrust::device_vector<range_type>   m_ranges;
    thrust::device_vector<uint64>       m_slots;
    thrust::device_vector<hit_type>     m_hits;
    thrust::device_vector<uint8>        d_temp_storage;
    thrust::device_vector<uint32>       m_layout;           // string offsets + lengths of the current query set
    nvb_fm_index                        m_view;

private:
    // B200 path: one kernel matches every query (nvb_fm_match semantics == nvbio::match(), fmindex_inl.h:280-341), then the scan
    template <typename string_set_type>
    void rank_dispatch(const string_set_type& string_set, std::true_type)
    {
        typedef b200::packed_string_set<string_set_type> set_traits;
        m_view = view_type::get( m_index );

        m_layout.resize( 2u * size_t( m_n_queries ) );
        uint32* d_off = thrust::raw_pointer_cast( m_layout.data() );
        uint32* d_len = d_off + m_n_queries;
        b200::string_set_layout_kernel<<< (m_n_queries + 255u) / 256u, 256u >>>( string_set, m_n_queries, d_off, d_len );

        nvb_string_set q;
        q.d_words    = (const uint32_t*)set_traits::words( string_set );
        q.bits       = set_traits::BITS;
        q.big_endian = set_traits::BE;
        q.d_offsets  = d_off;
        q.d_lengths  = d_len;
        q.stride     = 0u;
        q.length     = 0u;

        nvb_uint2* d_ranges = (nvb_uint2*)thrust::raw_pointer_cast( m_ranges.data() );
        uint64_t*  d_slots  = (uint64_t*)thrust::raw_pointer_cast( m_slots.data() );
        size_t   temp_bytes = 0u;
        uint64_t n_occ      = 0u;
        const int r = nvb_fm_filter_rank( &m_view, &q, m_n_queries, 0u, d_ranges, d_slots, NULL, NULL, &temp_bytes, NULL );
        if (r != NVB_E_TEMP_SIZE) b200::check( r, "nvb_fm_filter_rank (size query)" );
        if (d_temp_storage.size() < temp_bytes) d_temp_storage.resize( temp_bytes );
        temp_bytes = d_temp_storage.size();
        b200::check( nvb_fm_filter_rank( &m_view, &q, m_n_queries, 0u, d_ranges, d_slots, &n_occ,
                                         thrust::raw_pointer_cast( d_temp_storage.data() ), &temp_bytes, NULL ), "nvb_fm_filter_rank" );
        m_n_occurrences = n_occ;
        b200::stats().fm_rank++;
    }
    // reference path (a string set the shim cannot see through): the reference's own functors
    template <typename string_set_type>
    void rank_dispatch(const string_set_type& string_set, std::false_type)
    {
        thrust::transform(
            thrust::make_counting_iterator<uint32>(0u),
            thrust::make_counting_iterator<uint32>(0u) + m_n_queries,
            m_ranges.begin(),
            fmindex::rank_functor<fm_index_type,string_set_type>( m_index, string_set ) );
        thrust::inclusive_scan(
            thrust::make_transform_iterator( m_ranges.begin(), fmindex::range_size<range_type>() ),
            thrust::make_transform_iterator( m_ranges.begin(), fmindex::range_size<range_type>() ) + m_n_queries,
            m_slots.begin(),
            thrust::plus<uint64>() );
        m_n_occurrences = m_slots[ m_n_queries-1 ];
        m_view.d_bwt_occ = NULL;
        b200::stats().fallbacks++;
    }

    template <typename hits_iterator>
    void locate_dispatch(const uint64 begin, const uint64 end, hits_iterator hits, std::true_type)
    {
        if (m_view.d_bwt_occ == NULL)
            m_view = view_type::get( m_index );
        const int r = nvb_fm_filter_locate( &m_view,
                                            (const nvb_uint2*)thrust::raw_pointer_cast( m_ranges.data() ),
                                            (const uint64_t*)thrust::raw_pointer_cast( m_slots.data() ),
                                            m_n_queries, begin, end,
                                            (nvb_uint2*)b200::raw_device_pointer( hits ), NULL );
        if (r != NVB_OK)
        {
            char msg[512];
            snprintf( msg, sizeof(msg), "nvb_fm_filter_locate(index{blocks %p, ssa %p, n %u, primary %u, sa_interval %u}, ranges %p, slots %p, queries %u, [%llu, %llu), hits %p)",
                      m_view.d_bwt_occ, (const void*)m_view.d_ssa, m_view.length, m_view.primary, m_view.sa_interval,
                      (const void*)thrust::raw_pointer_cast( m_ranges.data() ), (const void*)thrust::raw_pointer_cast( m_slots.data() ), m_n_queries,
                      (unsigned long long)begin, (unsigned long long)end, (const void*)b200::raw_device_pointer( hits ) );
            b200::check( r, msg );
        }
        b200::stats().fm_locate++;
    }
    template <typename hits_iterator>
    void locate_dispatch(const uint64 begin, const uint64 end, hits_iterator hits, std::false_type)
    {
        const uint32 n_hits = uint32( end - begin );
        if (m_hits.size() < n_hits) { m_hits.clear(); m_hits.resize( n_hits ); }
        thrust::transform(
            thrust::make_counting_iterator<uint64>(0u) + begin,
            thrust::make_counting_iterator<uint64>(0u) + end,
            device_iterator( hits ),
            fmindex::filter_results<range_type>( m_n_queries, nvbio::plain_view( m_slots ), nvbio::plain_view( m_ranges ) ) );
        thrust::transform(
            device_iterator( hits ), device_iterator( hits ) + n_hits, m_hits.begin(),
            fmindex::locate_ssa_results<fm_index_type>( m_index ) );
        thrust::transform(
            device_iterator( hits ), device_iterator( hits ) + n_hits, m_hits.begin(), device_iterator( hits ),
            fmindex::lookup_ssa_results<fm_index_type>( m_index ) );
        b200::stats().fallbacks++;
    }
};

} // namespace nvbio
