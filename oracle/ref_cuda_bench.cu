// oracle/ref_cuda_bench.cu -- BASELINE INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The reference's own CUDA path (its sm_35-era kernels, header templates included from /root/reference where
// they lie) recompiled for sm_100a, driven through the reference's public batch APIs:
//   banded : aln::batch_banded_alignment_score<31>( make_gotoh_aligner<LOCAL>(SimpleGotohScheme), patterns, texts,
//            sinks, DeviceThreadScheduler(), ... )      nvbio/alignment/batched_inl.h:1067-1101
//            -> batched_banded_alignment_score_kernel     nvbio/alignment/batched_banded_inl.h:78-162
//   fm     : FMIndexFilterDevice<fm_index_type>::rank + locate   nvbio/fmindex/filter_inl.h:268-402
//   approx : nvBowtie's detail::map<true> (one-mismatch seed search, a __device__ function)   nvBowtie/bowtie2/cuda/mapping_inl.h:128-220
// Inputs are the arrays bench.py / the tests dump (same workload as the B200-native kernels); outputs are
// written back so that the results can be compared bit for bit.
//
// Build (oracle/Makefile, only where /root/reference exists):
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -use_fast_math -Xcompiler -fopenmp \
//        -I/root/reference -I/root/reference/contrib oracle/ref_cuda_bench.cu -o oracle/_ref/ref_cuda_bench
#include <nvbio/basic/types.h>
#include <nvbio/basic/packedstream.h>
#include <nvbio/basic/deinterleaved_iterator.h>
#include <nvbio/basic/cuda/ldg.h>
#include <nvbio/strings/string_set.h>
#include <nvbio/fmindex/bwt.h>
#include <nvbio/fmindex/fmindex.h>
#include <nvbio/fmindex/ssa.h>
#include <nvbio/fmindex/filter.h>
#include <nvbio/alignment/alignment.h>
#include <nvbio/alignment/batched.h>
#include <nvBowtie/bowtie2/cuda/mapping_inl.h>          // detail::map<find_exact>: nvBowtie's one-mismatch seed search (device function)
#include <nvbio/io/sequence/sequence.h>
#include <nvbio/io/sequence/sequence_access.h>
#include <thrust/device_vector.h>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

using namespace nvbio;

template <typename T>
static std::vector<T> read_file(const std::string& path)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
    fseek(f, 0, SEEK_END); const long bytes = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<T> v(bytes / sizeof(T));
    if (fread(v.data(), 1, bytes, f) != (size_t)bytes) { fprintf(stderr, "short read %s\n", path.c_str()); exit(2); }
    fclose(f);
    return v;
}
template <typename T>
static void write_file(const std::string& path, const T* p, size_t n)
{
    FILE* f = fopen(path.c_str(), "wb"); fwrite(p, sizeof(T), n, f); fclose(f);
}

// --------------------------------------------------------------------------------------------------
// full-matrix Gotoh (the DP sw-benchmark times): n patterns of M symbols against n texts of N symbols (2-bit big-endian,
// string i at symbols [i*stride, +len)), LOCAL, aln::batch_alignment_score with the DeviceThreadScheduler
// --------------------------------------------------------------------------------------------------
static int run_full(const std::string& dir)
{
    const std::vector<uint32> meta = read_file<uint32>(dir + "/meta.bin");    // n, M, Mstride, N, Nstride, match, mismatch, go, ge, reps
    const uint32 n = meta[0], M = meta[1], Mstride = meta[2], N = meta[3], Nstride = meta[4];
    const int32 s_match = (int32)meta[5], s_mm = (int32)meta[6], s_go = (int32)meta[7], s_ge = (int32)meta[8];
    const uint32 reps = meta[9];
    const std::vector<uint32> h_pat = read_file<uint32>(dir + "/pat_words.bin");
    const std::vector<uint32> h_txt = read_file<uint32>(dir + "/txt_words.bin");
    thrust::device_vector<uint32> d_pat(h_pat), d_txt(h_txt);
    std::vector<uint2> h_pr(n), h_tr(n);
    for (uint32 i = 0; i < n; ++i) { h_pr[i] = make_uint2(i * Mstride, i * Mstride + M); h_tr[i] = make_uint2(i * Nstride, i * Nstride + N); }
    thrust::device_vector<uint2> d_pr(h_pr), d_tr(h_tr);
    thrust::device_vector< aln::BestSink<int32> > d_sinks(n);

    typedef nvbio::cuda::ldg_pointer<uint32>                    storage_it;
    typedef PackedStream<storage_it, uint8, 2u, true>           stream_t;
    typedef SparseStringSet<stream_t, const uint2*>             set_t;
    const stream_t pat_stream( storage_it( thrust::raw_pointer_cast(d_pat.data()) ) );
    const stream_t txt_stream( storage_it( thrust::raw_pointer_cast(d_txt.data()) ) );
    const set_t patterns( n, pat_stream, thrust::raw_pointer_cast(d_pr.data()) );
    const set_t texts   ( n, txt_stream, thrust::raw_pointer_cast(d_tr.data()) );

    const aln::SimpleGotohScheme scheme( s_match, s_mm, s_go, s_ge );
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best_ms = 1e30f;
    for (uint32 r = 0; r < reps + 1; ++r)
    {
        cudaEventRecord(e0);
        aln::batch_alignment_score(
            aln::make_gotoh_aligner<aln::LOCAL>( scheme ),
            patterns, texts,
            thrust::raw_pointer_cast(d_sinks.data()),
            aln::DeviceThreadScheduler(),
            M, N );
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best_ms) best_ms = ms;
    }
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(err)); return 3; }
    std::vector< aln::BestSink<int32> > h_sinks(n);
    cudaMemcpy(h_sinks.data(), thrust::raw_pointer_cast(d_sinks.data()), sizeof(aln::BestSink<int32>) * n, cudaMemcpyDeviceToHost);
    std::vector<int32> score(n); std::vector<uint2> sink(n);
    for (uint32 i = 0; i < n; ++i) { score[i] = h_sinks[i].score; sink[i] = h_sinks[i].sink; }
    write_file(dir + "/ref_scores.bin", score.data(), n);
    write_file(dir + "/ref_sinks.bin", sink.data(), n);
    printf("{\"what\": \"reference CUDA batched_alignment_score_kernel (full-matrix LOCAL Gotoh, DeviceThreadScheduler), sm_100a\", \"n\": %u, \"M\": %u, \"N\": %u, "
           "\"ms\": %.4f, \"gcups\": %.2f}\n", n, M, N, best_ms, double(n) * M * N / (best_ms * 1e-3) / 1e9);
    return 0;
}

// --------------------------------------------------------------------------------------------------
// banded Gotoh: n patterns of M symbols (2-bit big-endian, pattern i at symbols [i*Mstride, +M)) against
// genome windows [begin,end)
// --------------------------------------------------------------------------------------------------
static int run_banded(const std::string& dir)
{
    const std::vector<uint32> meta = read_file<uint32>(dir + "/meta.bin");    // n, M, Mstride, match, mismatch(neg as int), go, ge, reps
    const uint32 n = meta[0], M = meta[1], Mstride = meta[2];
    const int32 s_match = (int32)meta[3], s_mm = (int32)meta[4], s_go = (int32)meta[5], s_ge = (int32)meta[6];
    const uint32 reps = meta[7];
    const std::vector<uint32> h_pat = read_file<uint32>(dir + "/pat_words.bin");
    const std::vector<uint32> h_gen = read_file<uint32>(dir + "/genome_words.bin");
    const std::vector<uint2>  h_win = read_file<uint2>(dir + "/windows.bin");

    thrust::device_vector<uint32> d_pat(h_pat), d_gen(h_gen);
    thrust::device_vector<uint2>  d_win(h_win);
    std::vector<uint2> h_prange(n);
    for (uint32 i = 0; i < n; ++i) h_prange[i] = make_uint2(i * Mstride, i * Mstride + M);
    thrust::device_vector<uint2> d_prange(h_prange);
    thrust::device_vector< aln::BestSink<int32> > d_sinks(n);

    typedef nvbio::cuda::ldg_pointer<uint32>                    storage_it;
    typedef PackedStream<storage_it, uint8, 2u, true>           stream_t;
    typedef SparseStringSet<stream_t, const uint2*>             set_t;

    const stream_t pat_stream( storage_it( thrust::raw_pointer_cast(d_pat.data()) ) );
    const stream_t gen_stream( storage_it( thrust::raw_pointer_cast(d_gen.data()) ) );
    const set_t patterns( n, pat_stream, thrust::raw_pointer_cast(d_prange.data()) );
    const set_t texts   ( n, gen_stream, thrust::raw_pointer_cast(d_win.data()) );

    const aln::SimpleGotohScheme scheme( s_match, s_mm, s_go, s_ge );
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best_ms = 1e30f;
    for (uint32 r = 0; r < reps + 1; ++r)
    {
        cudaEventRecord(e0);
        aln::batch_banded_alignment_score<31u>(
            aln::make_gotoh_aligner<aln::LOCAL>( scheme ),
            patterns, texts,
            thrust::raw_pointer_cast(d_sinks.data()),
            aln::DeviceThreadScheduler(),
            M, M + 31u + 8u );
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best_ms) best_ms = ms;              // first run is the warm-up
    }
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(err)); return 3; }
    std::vector< aln::BestSink<int32> > h_sinks(n);
    cudaMemcpy(h_sinks.data(), thrust::raw_pointer_cast(d_sinks.data()), sizeof(aln::BestSink<int32>) * n, cudaMemcpyDeviceToHost);
    std::vector<int32> score(n); std::vector<uint2> sink(n);
    for (uint32 i = 0; i < n; ++i) { score[i] = h_sinks[i].score; sink[i] = h_sinks[i].sink; }
    write_file(dir + "/ref_scores.bin", score.data(), n);
    write_file(dir + "/ref_sinks.bin", sink.data(), n);
    printf("{\"what\": \"reference CUDA batched_banded_alignment_score_kernel<128,1,31> (LOCAL Gotoh), sm_100a\", \"n\": %u, \"M\": %u, "
           "\"ms\": %.4f, \"gcups\": %.2f}\n", n, M, best_ms, double(n) * M * 31.0 / (best_ms * 1e-3) / 1e9);
    return 0;
}

// --------------------------------------------------------------------------------------------------
// FM-index: nq seeds of L symbols (2-bit big-endian, seed i at [i*Lstride, +L))
// --------------------------------------------------------------------------------------------------
static int run_fm(const std::string& dir)
{
    const std::vector<uint32> meta = read_file<uint32>(dir + "/meta.bin");    // length, primary, L2[5], nq, L, Lstride, reps
    const uint32 length = meta[0], primary = meta[1], nq = meta[7], L = meta[8], Lstride = meta[9], reps = meta[10];
    const std::vector<uint32> h_bwt_occ = read_file<uint32>(dir + "/bwt_occ.bin");
    const std::vector<uint32> h_ssa     = read_file<uint32>(dir + "/ssa.bin");
    const std::vector<uint32> h_seeds   = read_file<uint32>(dir + "/seed_words.bin");
    thrust::device_vector<uint32> d_bwt_occ(h_bwt_occ), d_ssa(h_ssa), d_seeds(h_seeds), d_L2(meta.begin() + 2, meta.begin() + 7);
    std::vector<uint32> h_ct(256); gen_bwt_count_table(h_ct.data());
    thrust::device_vector<uint32> d_ct(h_ct);
    std::vector<uint2> h_range(nq);
    for (uint32 i = 0; i < nq; ++i) h_range[i] = make_uint2(i * Lstride, i * Lstride + L);
    thrust::device_vector<uint2> d_range(h_range);

    // the production device index types (nvbio/io/fmindex/fmindex.h:302-319)
    typedef nvbio::cuda::ldg_pointer<uint4>                              bwt_occ_type;
    typedef deinterleaved_iterator<2,0,bwt_occ_type>                     bwt_type;
    typedef deinterleaved_iterator<2,1,bwt_occ_type>                     occ_type;
    typedef nvbio::cuda::ldg_pointer<uint32>                             u32_ldg;
    typedef PackedStream<bwt_type,uint8,2u,true>                         bwt_stream_type;
    typedef SSA_index_multiple_context<16u,u32_ldg>                      ssa_type;
    typedef rank_dictionary<2u,64u,bwt_stream_type,occ_type,u32_ldg>     rank_dict_type;
    typedef fm_index<rank_dict_type,ssa_type>                            fm_index_type;

    const bwt_occ_type p( (const uint4*)thrust::raw_pointer_cast(d_bwt_occ.data()) );
    const fm_index_type fmi( length, primary, thrust::raw_pointer_cast(d_L2.data()),
        rank_dict_type( bwt_stream_type( bwt_type(p) ), occ_type(p), u32_ldg( thrust::raw_pointer_cast(d_ct.data()) ) ),
        ssa_type( u32_ldg( thrust::raw_pointer_cast(d_ssa.data()) ) ) );

    typedef PackedStream<u32_ldg, uint8, 2u, true>          stream_t;
    typedef SparseStringSet<stream_t, const uint2*>         set_t;
    const set_t seeds( nq, stream_t( u32_ldg( thrust::raw_pointer_cast(d_seeds.data()) ) ), thrust::raw_pointer_cast(d_range.data()) );

    FMIndexFilterDevice<fm_index_type> filter;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float rank_ms = 1e30f, locate_ms = 1e30f; uint64 n_hits = 0;
    thrust::device_vector<uint2> d_hits;
    for (uint32 r = 0; r < reps + 1; ++r)
    {
        cudaEventRecord(e0);
        n_hits = filter.rank( fmi, seeds );
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < rank_ms) rank_ms = ms;
        const uint64 n_loc = n_hits < (uint64(1) << 26) ? n_hits : (uint64(1) << 26);
        d_hits.resize(n_loc);
        cudaEventRecord(e0);
        filter.locate( 0, n_loc, d_hits.begin() );
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < locate_ms) locate_ms = ms;
    }
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(err)); return 3; }
    std::vector<uint2> h_ranges(nq);
    cudaMemcpy(h_ranges.data(), filter.ranges(), sizeof(uint2) * nq, cudaMemcpyDeviceToHost);
    write_file(dir + "/ref_ranges.bin", h_ranges.data(), nq);
    std::vector<uint2> h_hits(d_hits.size());
    cudaMemcpy(h_hits.data(), thrust::raw_pointer_cast(d_hits.data()), sizeof(uint2) * d_hits.size(), cudaMemcpyDeviceToHost);
    write_file(dir + "/ref_hits.bin", h_hits.data(), h_hits.size());
    printf("{\"what\": \"reference CUDA FMIndexFilterDevice::rank/locate (thrust::transform of match()/locate), sm_100a\", \"nq\": %u, \"L\": %u, "
           "\"n_hits\": %llu, \"rank_ms\": %.4f, \"mseeds_per_s\": %.2f, \"locate_ms\": %.4f}\n",
           nq, L, (unsigned long long)n_hits, rank_ms, double(nq) / (rank_ms * 1e-3) / 1e6, locate_ms);
    return 0;
}

// --------------------------------------------------------------------------------------------------
// one-mismatch seed search: nvBowtie's own detail::map<true> (nvBowtie/bowtie2/cuda/mapping_inl.h:128-220) run over nq seeds of
// L unpacked symbols (seed i at bytes [i*L, +L), read forwards), exact region [0, len1); every pushed SeedHit range is recorded
// in push order (inclusive ranges), up to max_out per seed, with the push count and range_sum
// --------------------------------------------------------------------------------------------------
struct ApproxCollector
{
    uint2* out; uint32 n, cap;
    NVBIO_FORCEINLINE NVBIO_DEVICE uint32 size() const { return 0u; }                   // never "full": nothing is ever popped
    NVBIO_FORCEINLINE NVBIO_DEVICE void pop_bottom() {}
    NVBIO_FORCEINLINE NVBIO_DEVICE void push(const bowtie2::cuda::SeedHit hit)
    {
        const uint2 r = hit.get_range();                                                // exclusive end (inclusive_to_exclusive)
        if (n < cap) out[n] = make_uint2( r.x, r.y - 1u );
        ++n;
    }
};
struct SeedBytes
{
    const uint8* p;
    NVBIO_FORCEINLINE NVBIO_HOST_DEVICE uint8 operator[] (const uint32 i) const { return p[i]; }
};

template <typename fm_index_type>
__global__ void ref_map_approx_kernel(const fm_index_type fmi, const uint8* seeds, const uint32 nq, const uint32 L, const uint32 len1,
                                      const uint32 max_out, uint2* ranges, uint32* counts, uint32* sums)
{
    const uint32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    SeedBytes q; q.p = seeds + uint64(i) * L;
    ApproxCollector heap; heap.out = ranges + uint64(i) * max_out; heap.n = 0; heap.cap = max_out;
    uint32 range_sum = 0, range_count = 0;
    bowtie2::cuda::detail::map<true>( q, len1, L, fmi,
        bowtie2::cuda::SeedHit::build_flags( STANDARD, FORWARD, 0u ),
        heap, 0xFFFFFFFFu, range_sum, range_count );
    counts[i] = heap.n; sums[i] = range_sum;
}

static int run_approx(const std::string& dir)
{
    const std::vector<uint32> meta = read_file<uint32>(dir + "/meta.bin");    // length, primary, L2[5], nq, L, len1, max_out
    const uint32 length = meta[0], primary = meta[1], nq = meta[7], L = meta[8], len1 = meta[9], max_out = meta[10];
    const std::vector<uint32> h_bwt_occ = read_file<uint32>(dir + "/bwt_occ.bin");
    const std::vector<uint32> h_ssa     = read_file<uint32>(dir + "/ssa.bin");
    const std::vector<uint8>  h_seeds   = read_file<uint8>(dir + "/seed_bytes.bin");
    thrust::device_vector<uint32> d_bwt_occ(h_bwt_occ), d_ssa(h_ssa), d_L2(meta.begin() + 2, meta.begin() + 7);
    thrust::device_vector<uint8>  d_seeds(h_seeds);
    std::vector<uint32> h_ct(256); gen_bwt_count_table(h_ct.data());
    thrust::device_vector<uint32> d_ct(h_ct);

    typedef nvbio::cuda::ldg_pointer<uint4>                              bwt_occ_type;
    typedef deinterleaved_iterator<2,0,bwt_occ_type>                     bwt_type;
    typedef deinterleaved_iterator<2,1,bwt_occ_type>                     occ_type;
    typedef nvbio::cuda::ldg_pointer<uint32>                             u32_ldg;
    typedef PackedStream<bwt_type,uint8,2u,true>                         bwt_stream_type;
    typedef SSA_index_multiple_context<16u,u32_ldg>                      ssa_type;
    typedef rank_dictionary<2u,64u,bwt_stream_type,occ_type,u32_ldg>     rank_dict_type;
    typedef fm_index<rank_dict_type,ssa_type>                            fm_index_type;
    const bwt_occ_type p( (const uint4*)thrust::raw_pointer_cast(d_bwt_occ.data()) );
    const fm_index_type fmi( length, primary, thrust::raw_pointer_cast(d_L2.data()),
        rank_dict_type( bwt_stream_type( bwt_type(p) ), occ_type(p), u32_ldg( thrust::raw_pointer_cast(d_ct.data()) ) ),
        ssa_type( u32_ldg( thrust::raw_pointer_cast(d_ssa.data()) ) ) );

    thrust::device_vector<uint2>  d_ranges( uint64(nq) * max_out );
    thrust::device_vector<uint32> d_counts(nq), d_sums(nq);
    ref_map_approx_kernel<<<(nq + 127) / 128, 128>>>( fmi, thrust::raw_pointer_cast(d_seeds.data()), nq, L, len1, max_out,
        thrust::raw_pointer_cast(d_ranges.data()), thrust::raw_pointer_cast(d_counts.data()), thrust::raw_pointer_cast(d_sums.data()) );
    cudaDeviceSynchronize();
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(err)); return 3; }
    std::vector<uint2> h_ranges( d_ranges.size() ); std::vector<uint32> h_counts(nq), h_sums(nq);
    cudaMemcpy(h_ranges.data(), thrust::raw_pointer_cast(d_ranges.data()), sizeof(uint2) * h_ranges.size(), cudaMemcpyDeviceToHost);
    cudaMemcpy(h_counts.data(), thrust::raw_pointer_cast(d_counts.data()), sizeof(uint32) * nq, cudaMemcpyDeviceToHost);
    cudaMemcpy(h_sums.data(),   thrust::raw_pointer_cast(d_sums.data()),   sizeof(uint32) * nq, cudaMemcpyDeviceToHost);
    write_file(dir + "/ref_ranges.bin", h_ranges.data(), h_ranges.size());
    write_file(dir + "/ref_counts.bin", h_counts.data(), nq);
    write_file(dir + "/ref_sums.bin", h_sums.data(), nq);
    printf("{\"what\": \"nvBowtie detail::map<true> (one-mismatch seed search), sm_100a\", \"nq\": %u, \"L\": %u, \"len1\": %u}\n", nq, L, len1);
    return 0;
}

// --------------------------------------------------------------------------------------------------
// nvBowtie's seed mapping stage: its own map_queues_kernel<EXACT_MAPPING | APPROX_MAPPING> (nvBowtie/bowtie2/cuda/mapping_inl.h:539-591,
// seed_mapper<>::enact :229-366) over a DNA_N read batch, the production index type, a PingPong input queue and a real
// SeedHitDequeArray; every read's deque (as stored: interval-heap order), its size and the reseed flags are written back
// --------------------------------------------------------------------------------------------------
template <bowtie2::cuda::detail::MappingAlgorithm ALGO, typename batch_type, typename fm_index_type>
static void launch_mapq(const batch_type reads, const fm_index_type fmi, const uint32 retry, const nvbio::cuda::PingPongQueuesView<uint32> queues,
                        uint8* reseed, bowtie2::cuda::SeedHitDequeArrayDeviceView hits, const bowtie2::cuda::ParamsPOD params, const bool fw, const bool rc)
{
    const uint32 blocks = (queues.in_size + bowtie2::cuda::BLOCKDIM - 1) / bowtie2::cuda::BLOCKDIM;
    bowtie2::cuda::detail::map_queues_kernel<ALGO> <<<blocks, bowtie2::cuda::BLOCKDIM>>>( reads, fmi, fmi, retry, queues, reseed, hits, params, fw, rc );
}

static int run_mapq(const std::string& dir)
{
    // meta: length, primary, L2[5], n_reads, n_queue, algo, seed_len, seed_freq, max_hits, max_reseed, rep_seeds, subseed_len, min_read_len, retry, fw, rc, arena
    const std::vector<uint32> meta = read_file<uint32>(dir + "/meta.bin");
    const uint32 length = meta[0], primary = meta[1], n_reads = meta[7], n_queue = meta[8], algo = meta[9];
    const std::vector<uint32> h_bwt_occ = read_file<uint32>(dir + "/bwt_occ.bin");
    const std::vector<uint32> h_ssa     = read_file<uint32>(dir + "/ssa.bin");
    const std::vector<uint32> h_reads   = read_file<uint32>(dir + "/read_words.bin");      // DNA_N: 4 bits per symbol, big-endian
    const std::vector<uint32> h_index   = read_file<uint32>(dir + "/read_index.bin");      // n_reads + 1 symbol offsets
    const std::vector<uint32> h_queue   = read_file<uint32>(dir + "/queue.bin");
    thrust::device_vector<uint32> d_bwt_occ(h_bwt_occ), d_ssa(h_ssa), d_L2(meta.begin() + 2, meta.begin() + 7), d_reads(h_reads), d_index(h_index), d_queue(h_queue);
    std::vector<uint32> h_ct(256); gen_bwt_count_table(h_ct.data());
    thrust::device_vector<uint32> d_ct(h_ct);

    typedef nvbio::cuda::ldg_pointer<uint4>                              bwt_occ_type;
    typedef deinterleaved_iterator<2,0,bwt_occ_type>                     bwt_type;
    typedef deinterleaved_iterator<2,1,bwt_occ_type>                     occ_type;
    typedef nvbio::cuda::ldg_pointer<uint32>                             u32_ldg;
    typedef PackedStream<bwt_type,uint8,2u,true>                         bwt_stream_type;
    typedef SSA_index_multiple_context<16u,u32_ldg>                      ssa_type;
    typedef rank_dictionary<2u,64u,bwt_stream_type,occ_type,u32_ldg>     rank_dict_type;
    typedef fm_index<rank_dict_type,ssa_type>                            fm_index_type;
    const bwt_occ_type p( (const uint4*)thrust::raw_pointer_cast(d_bwt_occ.data()) );
    const fm_index_type fmi( length, primary, thrust::raw_pointer_cast(d_L2.data()),
        rank_dict_type( bwt_stream_type( bwt_type(p) ), occ_type(p), u32_ldg( thrust::raw_pointer_cast(d_ct.data()) ) ),
        ssa_type( u32_ldg( thrust::raw_pointer_cast(d_ssa.data()) ) ) );

    uint32 max_len = 0u; for (uint32 i = 0; i < n_reads; ++i) max_len = nvbio::max( max_len, h_index[i+1] - h_index[i] );
    io::SequenceDataInfo info;
    info.m_alphabet = DNA_N; info.m_n_seqs = n_reads; info.m_name_stream_len = 0u;
    info.m_sequence_stream_len = h_index[n_reads]; info.m_sequence_stream_words = uint32( h_reads.size() );
    info.m_has_qualities = 0u; info.m_min_sequence_len = 1u; info.m_max_sequence_len = max_len; info.m_avg_sequence_len = max_len;
    const io::ConstSequenceDataView view( info, thrust::raw_pointer_cast(d_reads.data()), thrust::raw_pointer_cast(d_index.data()), NULL, NULL, NULL );
    typedef io::SequenceDataAccess<DNA_N> batch_type;
    const batch_type reads( view );

    bowtie2::cuda::ParamsPOD params;
    memset( &params, 0, sizeof(params) );
    params.seed_len = meta[10];
    params.seed_freq = bowtie2::cuda::SimpleFunc( bowtie2::cuda::SimpleFunc::LinearFunc, float(meta[11]), 0.0f );     // a constant
    params.max_hits = meta[12]; params.max_reseed = meta[13]; params.rep_seeds = meta[14]; params.subseed_len = meta[15]; params.min_read_len = meta[16];
    const uint32 retry = meta[17]; const bool fw = meta[18] != 0u, rc = meta[19] != 0u; const uint32 arena = meta[20];

    // the deque array: global arena + per-read counts / indices + the bump-allocator pools (seed_hit_deque_array.h:62-75,157-204)
    thrust::device_vector<bowtie2::cuda::SeedHit> d_hits( arena );
    thrust::device_vector<uint32> d_counts( n_reads, 0u ), d_hindex( n_reads, 0u ), d_pindex( n_reads, 0u ), d_pool( 1, 0u ), d_ppool( 1, 0u );
    thrust::device_vector<float>  d_probs( 2u * arena + 1024u );
    bowtie2::cuda::SeedHitDequeArrayDeviceView hits(
        nvbio::device_view( d_counts ), nvbio::device_view( d_hindex ), nvbio::device_view( d_hits ), nvbio::device_view( d_pindex ),
        nvbio::device_view( d_probs ), nvbio::device_view( d_pool ), nvbio::device_view( d_ppool ) );
    thrust::device_vector<uint8> d_reseed( n_queue, uint8(7) );

    nvbio::cuda::PingPongQueuesView<uint32> queues;
    queues.in_size   = n_queue;
    queues.in_queue  = thrust::raw_pointer_cast( d_queue.data() );
    queues.out_size  = NULL;
    queues.out_queue = NULL;

    if (algo == 0u) launch_mapq<bowtie2::cuda::detail::EXACT_MAPPING>( reads, fmi, retry, queues, thrust::raw_pointer_cast(d_reseed.data()), hits, params, fw, rc );
    else            launch_mapq<bowtie2::cuda::detail::APPROX_MAPPING>( reads, fmi, retry, queues, thrust::raw_pointer_cast(d_reseed.data()), hits, params, fw, rc );
    cudaDeviceSynchronize();
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(err)); return 3; }

    std::vector<uint32> h_counts( n_reads ), h_hindex( n_reads ), h_pool( 1 ); std::vector<uint8> h_reseed( n_queue );
    std::vector<uint2> h_hits( arena );                                                  // SeedHit = 8 bytes: (range_begin, packed bits)
    cudaMemcpy( h_counts.data(), thrust::raw_pointer_cast(d_counts.data()), 4u * n_reads, cudaMemcpyDeviceToHost );
    cudaMemcpy( h_hindex.data(), thrust::raw_pointer_cast(d_hindex.data()), 4u * n_reads, cudaMemcpyDeviceToHost );
    cudaMemcpy( h_pool.data(),   thrust::raw_pointer_cast(d_pool.data()),   4u, cudaMemcpyDeviceToHost );
    cudaMemcpy( h_reseed.data(), thrust::raw_pointer_cast(d_reseed.data()), n_queue, cudaMemcpyDeviceToHost );
    cudaMemcpy( h_hits.data(),   thrust::raw_pointer_cast(d_hits.data()),   8u * size_t(arena), cudaMemcpyDeviceToHost );
    write_file( dir + "/ref_counts.bin", h_counts.data(), n_reads );
    write_file( dir + "/ref_index.bin",  h_hindex.data(), n_reads );
    write_file( dir + "/ref_reseed.bin", h_reseed.data(), n_queue );
    write_file( dir + "/ref_hits.bin",   h_hits.data(),   arena );
    printf("{\"what\": \"nvBowtie map_queues_kernel<%s>, sm_100a\", \"reads\": %u, \"queue\": %u, \"arena_used\": %u, \"sizeof_SeedHit\": %u}\n",
           algo == 0u ? "EXACT_MAPPING" : "APPROX_MAPPING", n_reads, n_queue, h_pool[0], uint32(sizeof(bowtie2::cuda::SeedHit)));
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: ref_cuda_bench banded|fm <dir>\n"); return 1; }
    const std::string mode = argv[1], dir = argv[2];
    if (mode == "banded") return run_banded(dir);
    if (mode == "fm")     return run_fm(dir);
    if (mode == "full")   return run_full(dir);
    if (mode == "approx") return run_approx(dir);
    if (mode == "mapq")   return run_mapq(dir);
    return 1;
}
