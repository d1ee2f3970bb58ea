// TEST INFRASTRUCTURE — not part of the product path.
//
// Drives the UNMODIFIED reference index code (included from $(REF) where it
// lies) to (1) build + freeze a dict_freq_index for one of the three DINT types
// and run the reference's own verify_collection on it, and (2) dump every
// docid/freq the reference enumerator yields, as the golden output for the
// block path. src/create_freq_index.cpp itself needs real Boost.PP (via
// index_types.hpp -> dec_time_prediction.hpp), so the three DINT aliases of
// include/index_types.hpp:54-80 are re-declared here.
//
//   ref_index build <type> <collection_basename> <out.bin>
//   ref_index dict  <type> <collection_basename>            (only builds ./dict.* in CWD)
//   ref_index dump  <type> <index.bin> <out_docs.u32> <out_freqs.u32>
//   ref_index scan  <type> <index.bin> [repeat]      (CPU baseline: full enumerator scan)
//   ref_index and   <type> <index.bin> <queries> [repeat]   (the reference's and_query<false> per query line:
//                                                           prints one result count per query, then a JSON timing line)
#include <cassert>
#include <cstring>
#include <fstream>
#include <iostream>
#include <numeric>
#include <unordered_map>

#include <succinct/mapper.hpp>
#include "configuration.hpp"
#include "semiasync_queue.hpp"
#include "binary_freq_collection.hpp"
#include "block_codecs.hpp"
#include "dint_configuration.hpp"
#include "dint_codecs.hpp"
#include "dictionary_types.hpp"
#include "dictionary_builders.hpp"
#include "block_statistics.hpp"
#include "dict_freq_index.hpp"
#include "block_freq_index.hpp"

namespace ds2i {
using adjusted_collector_type = adjusted<constants::max_entry_size>;
using adjusted_block_stats_type = block_statistics<adjusted_collector_type>;
using adjusted_block_multi_stats_type = block_multi_statistics<adjusted_collector_type>;
using single_rectangular_builder =
    decreasing_static_frequencies<single_dictionary_rectangular_type, adjusted_block_stats_type>;
using single_packed_builder =
    decreasing_static_frequencies<single_dictionary_packed_type, adjusted_block_stats_type>;
using multi_packed_builder =
    decreasing_static_frequencies<multi_dictionary_packed_type, adjusted_block_multi_stats_type>;
using single_rect_dint_index = dict_freq_index<single_rectangular_builder, opt_dint_single_dict_block>;
using single_packed_dint_index = dict_freq_index<single_packed_builder, opt_dint_single_dict_block>;
using multi_packed_dint_index = dict_freq_index<multi_packed_builder, opt_dint_multi_dict_block>;
// include/index_types.hpp:26: block_interpolative — the first of the other block codecs of block_freq_index
// (SURVEY §8f rank 4); same container, no dictionaries, every block binary-interpolative coded
using block_interpolative_index = block_freq_index<interpolative_block>;
using block_vbyte_index = block_freq_index<vbyte_block>;  // index_types.hpp:45
using block_u32_index = block_freq_index<u32_block>;      // index_types.hpp:44
}  // namespace ds2i

#include "util.hpp"
#include "verify_collection.hpp"
// include/ds2i/queries.hpp (and_query, read_query) pulls in include/index_types.hpp, which needs real Boost.PP and
// every codec of the zoo; the three DINT aliases it would provide are declared above, so it is pre-empted here
// (same include guard idiom: #pragma once files cannot be skipped, hence the empty stand-ins under oracle/shim/stubs).
#include "queries.hpp"

using namespace ds2i;

template <typename Index>
int build(const char* basename, const char* out) {
    binary_freq_collection input(basename);
    global_parameters params;
    typename Index::builder builder(input.num_docs(), params);
    builder.build_model(basename);
    for (auto const& plist : input) {
        uint64_t n = plist.docs.size();
        builder.add_posting_list(n, plist.docs.begin(), plist.freqs.begin(), 0);
    }
    Index coll;
    builder.build(coll);
    succinct::mapper::freeze(coll, out);
    verify_collection<binary_freq_collection, Index>(input, out);
    return 0;
}

template <typename Index>
int dict_only(const char* basename) {
    binary_freq_collection input(basename);
    global_parameters params;
    typename Index::builder builder(input.num_docs(), params);
    builder.build_model(basename);  // dict_freq_index.hpp:53-68 -> build_or_load_dict stores ./dict.*
    return 0;
}

template <typename Index>
int dump(const char* bin, const char* out_docs, const char* out_freqs) {
    Index idx;
    boost::iostreams::mapped_file_source m(bin);
    succinct::mapper::map(idx, m);
    std::ofstream fd(out_docs, std::ios::binary), ff(out_freqs, std::ios::binary);
    std::vector<uint32_t> d, f;
    uint64_t total = 0;
    for (size_t i = 0; i < idx.size(); ++i) {
        auto e = idx[i];
        d.clear();
        f.clear();
        for (size_t j = 0; j < e.size(); ++j, e.next()) {
            d.push_back(uint32_t(e.docid()));
            f.push_back(uint32_t(e.freq()));
        }
        fd.write(reinterpret_cast<char const*>(d.data()), d.size() * 4);
        ff.write(reinterpret_cast<char const*>(f.data()), f.size() * 4);
        total += d.size();
    }
    std::cerr << "dumped " << idx.size() << " lists, " << total << " postings" << std::endl;
    return 0;
}

template <typename Index>
int scan(const char* bin, int repeat) {
    Index idx;
    boost::iostreams::mapped_file_source m(bin);
    succinct::mapper::map(idx, m);
    for (int r = 0; r < repeat; ++r) {
        double t0 = get_time_usecs();
        uint64_t sum = 0, cnt = 0;
        for (size_t i = 0; i < idx.size(); ++i) {
            auto e = idx[i];
            for (size_t j = 0; j < e.size(); ++j, e.next()) {
                sum += e.docid() + e.freq();
                ++cnt;
            }
        }
        double t1 = get_time_usecs();
        std::cout << "{\"postings\": " << cnt << ", \"ints\": " << 2 * cnt << ", \"seconds\": " << (t1 - t0) / 1e6
                  << ", \"ints_x_sec\": " << uint64_t(2 * cnt / ((t1 - t0) / 1e6)) << ", \"checksum\": " << sum
                  << "}" << std::endl;
    }
    return 0;
}

template <typename Index>
int and_queries(const char* bin, const char* qfile, int repeat) {
    Index index;
    boost::iostreams::mapped_file_source m(bin);
    succinct::mapper::map(index, m);
    std::vector<term_id_vec> queries;
    {
        std::ifstream in(qfile);
        term_id_vec q;
        while (read_query(q, in)) queries.push_back(q);
    }
    std::vector<uint64_t> counts(queries.size());
    double best = 1e30;
    for (int r = 0; r < std::max(repeat, 1); ++r) {
        double t0 = get_time_usecs();
        for (size_t i = 0; i < queries.size(); ++i) counts[i] = and_query<false>()(index, queries[i]);
        best = std::min(best, (get_time_usecs() - t0) / 1e6);
    }
    for (uint64_t c : counts) std::cout << c << "\n";
    std::cout << "{\"queries\": " << queries.size() << ", \"seconds\": " << best << "}" << std::endl;
    return 0;
}

template <typename Index>
int run(int argc, char** argv) {
    std::string mode = argv[1];
    if (mode == "build" && argc >= 5) return build<Index>(argv[3], argv[4]);
    if (mode == "dict" && argc >= 4) return dict_only<Index>(argv[3]);
    if (mode == "dump" && argc >= 6) return dump<Index>(argv[3], argv[4], argv[5]);
    if (mode == "scan" && argc >= 4) return scan<Index>(argv[3], argc > 4 ? atoi(argv[4]) : 1);
    if (mode == "and" && argc >= 5) return and_queries<Index>(argv[3], argv[4], argc > 5 ? atoi(argv[5]) : 1);
    return 2;
}

int main(int argc, char** argv) {
    if (argc < 4) {
        std::cerr << "usage: ref_index build|dump|scan <type> ..." << std::endl;
        return 2;
    }
    std::string t = argv[2];
    if (t == "single_rect_dint") return run<single_rect_dint_index>(argc, argv);
    if (t == "single_packed_dint") return run<single_packed_dint_index>(argc, argv);
    if (t == "multi_packed_dint") return run<multi_packed_dint_index>(argc, argv);
    if (t == "block_interpolative") return run<block_interpolative_index>(argc, argv);
    if (t == "block_vbyte") return run<block_vbyte_index>(argc, argv);
    if (t == "block_u32") return run<block_u32_index>(argc, argv);
    std::cerr << "unknown type " << t << std::endl;
    return 2;
}
