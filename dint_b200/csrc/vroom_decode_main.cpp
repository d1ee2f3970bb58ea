// dint_b200_decode — host driver of the B200 DINT decoder: the counterpart of the reference's
// vroom_env/decode.cpp (decode_dint<Decoder, Dictionary>, :95-155) on top of the C ABI of
// libdint_b200.so. Plain C++; nothing here touches CUDA directly.
//
//   dint_b200_decode <type> <encoded_data_filename> --dict <dictionary_filename>
//                    [--repeat N] [--device D] [--mode raw|docids|freqs] [--sums <file>] [--host-plan]
//
// Same positional arguments, the same failure behaviour (std::runtime_error for a missing
// dictionary / unreadable file, "ERROR: unknown type" for an unknown codec) and the same JSON
// statistics line on stdout (vroom_env/statistics.hpp:26-34: filename, num_sequences,
// num_integers, type, tot_elapsed_time, ns_x_int, ints_x_sec), followed by the GPU extras.
// Where the reference times each Decoder::decode call with std::chrono and sums the deltas
// (decode.cpp:142-147), the whole batch decode is one stream-ordered call here, so the timed
// region is that call (+ synchronisation), averaged over --repeat passes over the resident file;
// the first pass, which includes the host->device copy of the file, is reported separately.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/dint_b200.h"

namespace {

std::vector<uint8_t> read_file(const char* path, const char* what) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error(what);
    std::vector<uint8_t> buf(size_t(f.tellg()));
    f.seekg(0);
    f.read(reinterpret_cast<char*>(buf.data()), std::streamsize(buf.size()));
    return buf;
}

void check(int rc, const char* what) {
    if (rc != DINT_OK) throw std::runtime_error(std::string(what) + ": " + dint_last_error() + " (" + std::to_string(rc) + ")");
}

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 3) {
        std::cerr << "Usage " << argv[0] << ":\n"
                  << "\t<type> <encoded_data_filename> --dict <dictionary_filename> [--repeat N] [--device D] "
                     "[--mode raw|docids|freqs] [--sums <out file>]"
                  << std::endl;
        return 1;
    }
    std::string type = argv[1];
    const char* encoded_data_filename = argv[2];
    const char* dictionary_filename = nullptr;
    const char* sums_filename = nullptr;
    int repeat = 10, device = 0, mode = DINT_MODE_RAW;
    bool host_plan = false;
    for (int i = 3; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "--dict" && i + 1 < argc) dictionary_filename = argv[++i];
        else if (a == "--repeat" && i + 1 < argc) repeat = std::max(1, atoi(argv[++i]));
        else if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
        else if (a == "--sums" && i + 1 < argc) sums_filename = argv[++i];
        else if (a == "--host-plan") host_plan = true;
        else if (a == "--mode" && i + 1 < argc) {
            std::string m = argv[++i];
            mode = m == "docids" ? DINT_MODE_DOCIDS : m == "freqs" ? DINT_MODE_FREQS : DINT_MODE_RAW;
        } else throw std::runtime_error("unknown parameter");  // decode.cpp:230
    }
    int kind;
    if (type == "single_rect_dint") kind = DINT_DICT_RECT;
    else if (type == "single_packed_dint") kind = DINT_DICT_PACKED;
    else if (type == "multi_packed_dint") kind = DINT_DICT_MULTI;
    else {
        std::cerr << "ERROR: unknown type '" << type << "'" << std::endl;  // decode.cpp:259
        return 0;
    }
    if (!dictionary_filename) throw std::runtime_error("dictionary_filename must be specified");  // decode.cpp:98-100

    std::vector<uint8_t> bin = read_file(encoded_data_filename, "Error opening index file");  // decode.cpp:104-106
    std::vector<uint8_t> dict_bytes = read_file(dictionary_filename, "Error opening dictionary file");

    dint_dict* dict = nullptr;
    check(dint_dict_create(kind, dict_bytes.data(), dict_bytes.size(), device, &dict), "dint_dict_create");
    // load time: file -> HBM, and the GPU finds every list / block / strip start in it (no host walk of the codewords;
    // --host-plan: the serial host walk dint_vroom_plan_create instead)
    double t0 = now();
    dint_vroom_plan* plan = nullptr;
    dint_vroom_ctx* ctx = nullptr;
    double discover_ms = 0;
    if (host_plan) {
        check(dint_vroom_plan_create(dict, bin.data(), bin.size(), 0, 0, &plan), "dint_vroom_plan_create");
        check(dint_vroom_ctx_create(dict, plan, 0, &ctx), "dint_vroom_ctx_create");
    } else {
        check(dint_vroom_ctx_create_discover(dict, bin.data(), bin.size(), 0, nullptr, 0, 0, nullptr, &ctx),
              "dint_vroom_ctx_create_discover");
        discover_ms = dint_device_plan_discover_ms(dint_vroom_ctx_device_plan(ctx));
    }
    double t_plan = now() - t0;
    const uint64_t n_lists = dint_vroom_ctx_num_lists(ctx), n_ints = dint_vroom_ctx_num_ints(ctx);

    std::vector<uint64_t> sums(n_lists ? n_lists : 1);
    // pass 0: file -> HBM, decode, per-list checksums back (the end-to-end call)
    t0 = now();
    check(dint_vroom_decode_host(ctx, bin.data(), bin.size(), mode, nullptr, sums.data(), nullptr), "dint_vroom_decode_host");
    check(dint_vroom_ctx_sync(ctx, nullptr), "sync");
    double t_first = now() - t0;
    // the timed passes: the file is resident, like the reference's mmap after the first touch
    t0 = now();
    for (int r = 0; r < repeat; ++r) check(dint_vroom_decode_resident(ctx, mode, nullptr), "dint_vroom_decode_resident");
    check(dint_vroom_ctx_sync(ctx, nullptr), "sync");
    double tot_elapsed = (now() - t0) / repeat;

    uint64_t checksum = 0;
    for (uint64_t i = 0; i < n_lists; ++i) checksum += sums[i];
    if (sums_filename) {
        std::ofstream f(sums_filename, std::ios::binary);
        f.write(reinterpret_cast<const char*>(sums.data()), std::streamsize(n_lists * 8));
    }
    static const uint64_t billion = 1000000000;
    double ns_x_int = n_ints ? tot_elapsed * billion / double(n_ints) : 0.0;
    uint64_t ints_x_sec = ns_x_int > 0 ? uint64_t(1 / ns_x_int * billion) : 0;
    std::cerr << "elapsed time " << tot_elapsed << " [sec]\n" << ns_x_int << " [ns] x int\n" << ints_x_sec << " ints x [sec]" << std::endl;
    std::cout << "{";
    std::cout << "\"filename\": \"" << encoded_data_filename << "\", ";
    std::cout << "\"num_sequences\": \"" << n_lists << "\", ";
    std::cout << "\"num_integers\": \"" << n_ints << "\", ";
    std::cout << "\"type\": \"" << type << "\", ";
    std::cout << "\"tot_elapsed_time\": \"" << tot_elapsed << "\", ";
    std::cout << "\"ns_x_int\": \"" << ns_x_int << "\", ";
    std::cout << "\"ints_x_sec\": \"" << ints_x_sec << "\", ";
    std::cout << "\"device\": " << device << ", \"repeat\": " << repeat << ", \"mode\": " << mode
              << ", \"plan_sec\": " << t_plan << ", \"plan\": \"" << (host_plan ? "host walk" : "device discovery")
              << "\", \"discover_device_ms\": " << discover_ms << ", \"first_pass_with_h2d_sec\": " << t_first
              << ", \"kernel_launches\": " << dint_kernel_launches() << ", \"checksum\": \"" << checksum << "\"";
    std::cout << "}" << std::endl;

    dint_vroom_ctx_destroy(ctx);
    dint_vroom_plan_destroy(plan);
    dint_dict_destroy(dict);
    return 0;
}
