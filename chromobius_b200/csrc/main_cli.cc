// Command line front end of the B200 decoder: `chromobius_b200 predict | benchmark | describe_decoder | help`.
//
// Mirrors the reference CLI (citations relative to /root/reference/src/chromobius):
//   commands/main_all.cc:27-88               command dispatch and the help text's command list
//   commands/main_predict.cc:22-81           predict: detection-event records in, predicted observable records out
//   commands/main_benchmark.cc:23-132        benchmark: mistake / timing report (same field names and layout)
//   commands/main_describe_decoder.cc:22-50  describe_decoder: text dump of the compiled tables
//   ../main.cc:21-26                         std::invalid_argument -> message on stderr, exit status 1
// The reference decodes one record at a time; here records are read in batches, decoded by one cb200_decode_batch call
// (the C ABI in include/chromobius_b200.h) and written back in order. Record formats (stim's, restated: SURVEY App. C.3):
//   01    one text line of '0' / '1' per record
//   b8    ceil(bits / 8) bytes per record, bit k = (byte[k >> 3] >> (k & 7)) & 1
//   dets  "shot" followed by D<k> (detectors) and L<k> (observables) tokens, one line per record
//   hits  comma separated indices of the set bits, one line per record
// There is no CPU fallback: without a CUDA device cb200_compile fails and the command exits with status 1.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/chromobius_b200.h"

namespace {

[[noreturn]] void fail(const std::string &msg) { throw std::invalid_argument(msg); }

const char *HELP = R"HELP(

Available chromobius_b200 commands:

    # Print usage information.
    chromobius_b200 help

    # Predict observable flips from detection event data.
    chromobius_b200 predict \
        [--dem FILEPATH] \                     # where to read detector error model from
        [--in] \                               # where to read detection event data (defaults to stdin)
        [--in_format 01|b8|dets|hits] \        # format of input detection event data
        [--in_includes_appended_observables] \ # if set, input data includes observables as extra detectors to ignore
        [--out FILEPATH] \                     # where to write predictions to (defaults to stdout)
        [--out_format 01|b8|dets|hits] \       # format to use when writing predictions
        [--device N] [--batch SHOTS]           # GPU to use, records decoded per call

    # Print accuracy and timing statistics collected while decoding.
    chromobius_b200 benchmark
        [--dem FILEPATH] \                     # where to read detector error model from
        [--in] \                               # where to read detection event data (defaults to stdin)
        [--in_format 01|b8|dets|hits] \        # format of input detection event data
        [--in_includes_appended_observables] \ # if set, observables are extra detectors in detection event data
        [--obs_in FILEPATH] \                  # if set, observables are read from a separate file
        [--obs_in_format 01|b8|dets|hits] \    # format of separate observable data
        [--out FILEPATH] \                     # where to write results (defaults to stdout)
        [--device N] [--batch SHOTS]

    # Describes the internal representations used to decode a given dem.
    chromobius_b200 describe_decoder \
        [--in] \           # where to read a detector error model from (defaults to stdin)
        [--out FILEPATH] \ # where to write output (defaults to stdout)
        [--device N]
)HELP";

// ---- arguments ------------------------------------------------------------------------------------------------------
struct Args {
    std::string command;
    std::vector<std::pair<std::string, std::string>> kv;  // "--name" -> value ("" for flags)

    Args(int argc, const char **argv, const std::vector<std::string> &known, const std::vector<std::string> &flags) {
        command = argv[1];
        for (int k = 2; k < argc; k++) {
            std::string a = argv[k];
            bool is_flag = false, is_known = false;
            for (const auto &f : flags) is_flag |= f == a;
            for (const auto &f : known) is_known |= f == a;
            if (is_flag) {
                kv.push_back({a, "1"});
            } else if (is_known) {
                if (k + 1 >= argc) fail("Command line argument " + a + " needs a value.");
                kv.push_back({a, argv[++k]});
            } else {
                std::stringstream ss;
                ss << "Unrecognized command line argument " << a << " for `chromobius_b200 " << command << "`.\n";
                ss << "Recognized command line arguments for `chromobius_b200 " << command << "`:\n";
                for (const auto &f : known) ss << "    " << f << "\n";
                for (const auto &f : flags) ss << "    " << f << "\n";
                fail(ss.str());
            }
        }
    }
    const char *get(const std::string &name) const {
        for (const auto &p : kv)
            if (p.first == name) return p.second.c_str();
        return nullptr;
    }
    bool flag(const std::string &name) const { return get(name) != nullptr; }
    uint64_t get_u64(const std::string &name, uint64_t dflt) const {
        const char *v = get(name);
        if (!v) return dflt;
        char *end = nullptr;
        unsigned long long x = strtoull(v, &end, 10);
        if (end == v || *end) fail("Command line argument " + name + " needs a non-negative integer, not '" + v + "'.");
        return x;
    }
};

FILE *open_arg(const Args &a, const std::string &name, FILE *dflt, const char *mode) {
    const char *path = a.get(name);
    if (!path) {
        if (!dflt) fail("Missing command line argument: '" + name + "'");
        return dflt;
    }
    FILE *f = fopen(path, mode);
    if (!f) fail(std::string("Failed to open '") + path + "'");
    return f;
}

std::string read_all(FILE *f) {
    std::string out;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out.append(buf, n);
    return out;
}

// ---- record formats -------------------------------------------------------------------------------------------------
enum Format { FMT_01, FMT_B8, FMT_DETS, FMT_HITS };

Format parse_format(const Args &a, const std::string &name, const char *dflt) {
    const char *v = a.get(name);
    std::string s = v ? v : dflt;
    if (s == "01") return FMT_01;
    if (s == "b8") return FMT_B8;
    if (s == "dets") return FMT_DETS;
    if (s == "hits") return FMT_HITS;
    fail("Unrecognized value '" + s + "' for enum flag '" + name + "'.\nRecognized values are:\n    '01'\n    'b8'\n    'dets'\n    'hits'\n");
}

// Reads records of `num_dets` detector bits followed by `num_obs` observable bits.
struct RecordReader {
    FILE *f;
    Format fmt;
    size_t num_dets, num_obs;
    std::vector<uint8_t> tmp;
    std::string line;

    RecordReader(FILE *file, Format format, size_t nd, size_t no) : f(file), fmt(format), num_dets(nd), num_obs(no) {
        tmp.resize((nd + no + 7) / 8);
    }
    bool read_line() {
        line.clear();
        int c;
        bool any = false;
        while ((c = fgetc(f)) != EOF) {
            any = true;
            if (c == '\n') return true;
            if (c != '\r') line.push_back((char)c);
        }
        return any;
    }
    // dets_row: ceil(num_dets / 8) zeroed bytes (nullable when num_dets == 0); obs: receives the observable bits.
    bool next(uint8_t *dets_row, uint64_t *obs) {
        const size_t total = num_dets + num_obs;
        *obs = 0;
        auto set_bit = [&](size_t k) {
            if (k < num_dets) dets_row[k >> 3] |= (uint8_t)(1u << (k & 7));
            else *obs |= uint64_t{1} << (k - num_dets);
        };
        switch (fmt) {
            case FMT_B8: {
                if (tmp.empty()) fail("Zero-bit records cannot be read from the b8 format.");
                size_t n = fread(tmp.data(), 1, tmp.size(), f);
                if (n == 0) return false;
                if (n != tmp.size()) fail("b8 data ended in the middle of a record.");
                for (size_t k = 0; k < total; k++)
                    if ((tmp[k >> 3] >> (k & 7)) & 1) set_bit(k);
                return true;
            }
            case FMT_01: {
                if (!read_line()) return false;
                if (line.size() != total) {
                    if (line.empty() && feof(f)) return false;
                    fail("01 data didn't have the expected number of bits per record (" + std::to_string(total) + "): '" + line + "'");
                }
                for (size_t k = 0; k < total; k++) {
                    if (line[k] == '1') set_bit(k);
                    else if (line[k] != '0') fail("Unexpected character in 01 format data: '" + std::string(1, line[k]) + "'");
                }
                return true;
            }
            case FMT_HITS: {
                if (!read_line()) return false;
                size_t pos = 0;
                while (pos < line.size()) {
                    size_t end = line.find(',', pos);
                    if (end == std::string::npos) end = line.size();
                    std::string tok = line.substr(pos, end - pos);
                    pos = end + 1;
                    if (tok.empty()) continue;
                    char *e = nullptr;
                    unsigned long long k = strtoull(tok.c_str(), &e, 10);
                    if (*e) fail("Unexpected token in hits format data: '" + tok + "'");
                    if (k >= total) fail("hits data has a bit index " + std::to_string(k) + " beyond the record size " + std::to_string(total) + ".");
                    set_bit(k);
                }
                return true;
            }
            case FMT_DETS: {
                // skip blank lines; a record line starts with "shot"
                while (true) {
                    if (!read_line()) return false;
                    size_t a = line.find_first_not_of(" \t");
                    if (a != std::string::npos) {
                        line = line.substr(a);
                        break;
                    }
                }
                std::stringstream ss(line);
                std::string tok;
                ss >> tok;
                if (tok != "shot") fail("dets format data must start each record with 'shot', got '" + line + "'");
                while (ss >> tok) {
                    char kind = tok[0];
                    char *e = nullptr;
                    unsigned long long k = strtoull(tok.c_str() + 1, &e, 10);
                    if (tok.size() < 2 || *e) fail("Unexpected token in dets format data: '" + tok + "'");
                    if (kind == 'D') {
                        if (k >= num_dets) fail("dets data has a detector D" + std::to_string(k) + " beyond the " + std::to_string(num_dets) + " detectors of the model.");
                        set_bit(k);
                    } else if (kind == 'L') {
                        if (k >= num_obs) fail("dets data has an observable L" + std::to_string(k) + " beyond the expected " + std::to_string(num_obs) + " observables.");
                        set_bit(num_dets + k);
                    } else {
                        fail("Unexpected token in dets format data: '" + tok + "'");
                    }
                }
                return true;
            }
        }
        return false;
    }
};

// Writes records of `bits` bits of result type `kind` ('L' for predictions).
struct RecordWriter {
    FILE *f;
    Format fmt;
    size_t bits;
    char kind;
    std::string buf;
    void write(uint64_t mask) {
        switch (fmt) {
            case FMT_B8:
                for (size_t b = 0; b < (bits + 7) / 8; b++) buf.push_back((char)((mask >> (8 * b)) & 0xFF));
                break;
            case FMT_01:
                for (size_t k = 0; k < bits; k++) buf.push_back(((mask >> k) & 1) ? '1' : '0');
                buf.push_back('\n');
                break;
            case FMT_HITS: {
                bool first = true;
                for (size_t k = 0; k < bits; k++)
                    if ((mask >> k) & 1) {
                        if (!first) buf.push_back(',');
                        first = false;
                        buf += std::to_string(k);
                    }
                buf.push_back('\n');
                break;
            }
            case FMT_DETS:
                buf += "shot";
                for (size_t k = 0; k < bits; k++)
                    if ((mask >> k) & 1) {
                        buf.push_back(' ');
                        buf.push_back(kind);
                        buf += std::to_string(k);
                    }
                buf.push_back('\n');
                break;
        }
        if (buf.size() > (1u << 20)) flush();
    }
    void flush() {
        if (!buf.empty() && fwrite(buf.data(), 1, buf.size(), f) != buf.size()) fail("Failed to write output.");
        buf.clear();
    }
};

struct DecoderHandle {
    cb200_decoder *d = nullptr;
    DecoderHandle(const std::string &dem_text, int device) {
        if (cb200_compile(dem_text.c_str(), device, &d) != CB200_OK) fail(cb200_last_error());
    }
    ~DecoderHandle() { cb200_free(d); }
};

// Reads up to `batch` records, decodes them in one call. Returns the number of records read.
struct BatchDecoder {
    DecoderHandle &dec;
    size_t nd, no, nb, ob;
    std::vector<uint8_t> dets, pred;
    std::vector<uint64_t> obs_in;
    uint64_t detection_events = 0;
    explicit BatchDecoder(DecoderHandle &d)
        : dec(d), nd(cb200_num_detectors(d.d)), no(cb200_num_observables(d.d)), nb(cb200_num_detector_bytes(d.d)),
          ob(cb200_num_observable_bytes(d.d)) {}
    size_t run(RecordReader &reader, RecordReader *obs_reader, size_t batch) {
        dets.assign(batch * nb, 0);
        obs_in.assign(batch, 0);
        size_t n = 0;
        while (n < batch) {
            uint64_t o = 0;
            if (!reader.next(nb ? dets.data() + n * nb : nullptr, &o)) break;
            if (obs_reader && !obs_reader->next(nullptr, &o)) fail("Obs data ended before shot data ended.");
            obs_in[n++] = o;
        }
        if (n == 0) return 0;
        pred.assign(n * ob, 0);
        cb200_batch_stats st{};
        static uint8_t dummy = 0;
        if (cb200_decode_batch(dec.d, nb ? dets.data() : &dummy, n, nb, ob ? pred.data() : &dummy, nullptr, nullptr,
                               CB200_IO_HOST, &st) != CB200_OK)
            fail(cb200_last_error());
        // the reference counts every set bit of the record, ignored detectors included (main_benchmark.cc:87)
        for (size_t k = 0; k < n * nb; k++) detection_events += (uint64_t)__builtin_popcount(dets[k]);
        return n;
    }
    uint64_t prediction(size_t k) const {
        uint64_t m = 0;
        for (size_t b = 0; b < ob; b++) m |= (uint64_t)pred[k * ob + b] << (8 * b);
        return m;
    }
};

// commands/main_predict.cc:22-81
int main_predict(int argc, const char **argv) {
    Args a(argc, argv, {"--in", "--in_format", "--out", "--out_format", "--dem", "--device", "--batch"},
           {"--in_includes_appended_observables"});
    FILE *shots_in = open_arg(a, "--in", stdin, "rb");
    FILE *predictions_out = open_arg(a, "--out", stdout, "wb");
    FILE *dem_file = open_arg(a, "--dem", nullptr, "rb");
    Format in_format = parse_format(a, "--in_format", "b8");
    Format out_format = parse_format(a, "--out_format", "01");
    bool append_obs = a.flag("--in_includes_appended_observables");
    std::string dem_text = read_all(dem_file);
    fclose(dem_file);
    DecoderHandle dec(dem_text, (int)a.get_u64("--device", 0));
    BatchDecoder bd(dec);
    RecordReader reader(shots_in, in_format, bd.nd, append_obs ? bd.no : 0);
    RecordWriter writer{predictions_out, out_format, bd.no, 'L', {}};
    const size_t batch = (size_t)std::max<uint64_t>(1, a.get_u64("--batch", 1u << 16));
    while (size_t n = bd.run(reader, nullptr, batch))
        for (size_t k = 0; k < n; k++) writer.write(bd.prediction(k));
    writer.flush();
    if (predictions_out != stdout) fclose(predictions_out);
    if (shots_in != stdin) fclose(shots_in);
    return EXIT_SUCCESS;
}

// commands/main_benchmark.cc:23-132
int main_benchmark(int argc, const char **argv) {
    auto time_config_starts = std::chrono::steady_clock::now();
    Args a(argc, argv, {"--in", "--in_format", "--obs_in", "--obs_in_format", "--out", "--dem", "--device", "--batch"},
           {"--in_includes_appended_observables"});
    FILE *shots_in = open_arg(a, "--in", stdin, "rb");
    FILE *obs_in = a.get("--obs_in") ? open_arg(a, "--obs_in", nullptr, "rb") : nullptr;
    FILE *stats_out = open_arg(a, "--out", stdout, "wb");
    FILE *dem_file = open_arg(a, "--dem", nullptr, "rb");
    Format in_format = parse_format(a, "--in_format", "01");
    Format obs_in_format = parse_format(a, "--obs_in_format", "01");
    bool append_obs = a.flag("--in_includes_appended_observables");
    if (!append_obs && obs_in == nullptr) fail("Must specify --in_includes_appended_observables or --obs_in.");
    std::string dem_text = read_all(dem_file);
    fclose(dem_file);
    DecoderHandle dec(dem_text, (int)a.get_u64("--device", 0));
    BatchDecoder bd(dec);
    RecordReader reader(shots_in, in_format, bd.nd, append_obs ? bd.no : 0);
    RecordReader obs_reader(obs_in, obs_in_format, 0, bd.no);
    const size_t batch = (size_t)std::max<uint64_t>(1, a.get_u64("--batch", 1u << 16));
    size_t num_mistakes = 0, num_shots = 0;
    auto time_config_ends_decoding_starts = std::chrono::steady_clock::now();
    while (size_t n = bd.run(reader, obs_in ? &obs_reader : nullptr, batch)) {
        for (size_t k = 0; k < n; k++)
            if (bd.obs_in[k] != bd.prediction(k)) num_mistakes++;
        num_shots += n;
    }
    auto time_decoding_ends = std::chrono::steady_clock::now();
    auto decoding_microseconds =
        (double)std::chrono::duration_cast<std::chrono::microseconds>(time_decoding_ends - time_config_ends_decoding_starts).count();
    auto config_microseconds =
        (double)std::chrono::duration_cast<std::chrono::microseconds>(time_config_ends_decoding_starts - time_config_starts).count();
    const uint64_t num_detection_events = bd.detection_events;
    auto total_detectors = (uint64_t)bd.nd * (uint64_t)num_shots;
    std::stringstream output;
    output << "                                num_shots = " << num_shots << "\n";
    output << "                             num_mistakes = " << num_mistakes << "\n";
    output << "                        mistakes_per_shot = " << (num_shots ? (double)num_mistakes / (double)num_shots : 0) << "\n";
    output << "\n";
    output << "                     num_detection_events = " << num_detection_events << "\n";
    output << "                   num_detectors_per_shot = " << bd.nd << "\n";
    output << "                       detection_fraction = " << ((double)num_detection_events / (double)total_detectors) << "\n";
    output << "\n";
    output << "                            setup_seconds = " << (config_microseconds / 1000000) << "\n";
    output << "                         decoding_seconds = " << (decoding_microseconds / 1000000) << "\n";
    output << "           decoding_microseconds_per_shot = " << (decoding_microseconds / num_shots) << "\n";
    output << "decoding_microseconds_per_detection_event = " << (decoding_microseconds / num_detection_events) << "\n";
    fprintf(stats_out, "%s", output.str().c_str());
    if (stats_out != stdout) fclose(stats_out);
    if (shots_in != stdin) fclose(shots_in);
    if (obs_in != nullptr) fclose(obs_in);
    return EXIT_SUCCESS;
}

// commands/main_describe_decoder.cc:22-50 (the --circuit input needs stim's circuit analyser and is not offered)
int main_describe_decoder(int argc, const char **argv) {
    Args a(argc, argv, {"--in", "--out", "--device"}, {});
    FILE *dem_in = open_arg(a, "--in", stdin, "rb");
    FILE *out = open_arg(a, "--out", stdout, "wb");
    std::string dem_text = read_all(dem_in);
    if (dem_in != stdin) fclose(dem_in);
    DecoderHandle dec(dem_text, (int)a.get_u64("--device", 0));
    char *text = cb200_describe(dec.d);
    if (!text) fail(cb200_last_error());
    fprintf(out, "%s\n", text);
    cb200_free_string(text);
    if (out != stdout) fclose(out);
    return EXIT_SUCCESS;
}

// commands/main_all.cc:27-88
int main_all(int argc, const char **argv) {
    const char *command = argc >= 2 ? argv[1] : "";
    if (strcmp(command, "describe_decoder") == 0) return main_describe_decoder(argc, argv);
    if (strcmp(command, "predict") == 0) return main_predict(argc, argv);
    if (strcmp(command, "benchmark") == 0) return main_benchmark(argc, argv);
    if (strcmp(command, "help") == 0 || strcmp(command, "--help") == 0 || strcmp(command, "-help") == 0 || strcmp(command, "-h") == 0) {
        fputs(HELP, stdout);
        return EXIT_SUCCESS;
    }
    std::stringstream ss;
    if (command[0] == '\0') ss << "Specify a chromobius_b200 command to run.\n";
    else ss << "Unrecognized chromobius_b200 command '" << command << "'.\n";
    ss << HELP;
    fail(ss.str());
}

}  // namespace

// ../main.cc:21-26
int main(int argc, const char **argv) {
    try {
        return main_all(argc, argv);
    } catch (const std::invalid_argument &ex) {
        fprintf(stderr, "%s\n", ex.what());
        return EXIT_FAILURE;
    }
}
