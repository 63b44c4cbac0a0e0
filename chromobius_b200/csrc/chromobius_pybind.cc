// pybind11 module named `chromobius`: the reference's Python surface on top of the C ABI (include/chromobius_b200.h).
//
// Mirrors /root/reference/src/chromobius/pybind/chromobius.pybind.cc:26-118, 120-440 (CompiledDecoder,
// compile_decoder_for_dem, __version__) and /root/reference/src/chromobius/pybind/sinter_compat.pybind.cc:23-241
// (_ChromobiusSinterDecoder, _ChromobiusSinterCompiledDecoder, sinter_decoders): same names, argument checks, error
// messages and return types. Differences, all behind the same signatures:
//   * the serial loop over Decoder::decode_detection_events becomes ONE cb200_decode_batch call (CUDA, sm_100a);
//   * the GIL is released while the GPU decodes;
//   * `device=` (keyword only, optional) picks the GPU -- an index, a list of indices or 'all' (the batch is then split
//     over those GPUs inside the one call, cb200_compile_multi); the default follows CHROMOBIUS_B200_DEVICE / LOCAL_RANK / 0.
// There is no CPU fallback: without a CUDA device compile_decoder_for_dem raises.
//
// Build (see __graft_entry__.build_pybind): g++ -shared $(python -m pybind11 --includes) ... -l:libchromobius_b200.so
#include <pybind11/numpy.h>
#include <pybind11/operators.h>
#include <pybind11/pybind11.h>

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/chromobius_b200.h"

namespace py = pybind11;

namespace {

[[noreturn]] void raise_last(int rc) {
    std::string msg = cb200_last_error();
    if (rc == CB200_ERR_CUDA) throw std::runtime_error(msg);
    throw std::invalid_argument(msg);  // pybind11 turns this into ValueError, as the reference's exceptions are
}

int default_device() {
    for (const char *key : {"CHROMOBIUS_B200_DEVICE", "LOCAL_RANK"}) {
        const char *v = getenv(key);
        if (v && *v) {
            char *end = nullptr;
            long d = strtol(v, &end, 10);
            if (end && *end == 0) return (int)d;
        }
    }
    return 0;
}

// `device`: None (default device), an int, a list / tuple of ints, or 'all' (every visible GPU).
struct DeviceSpec {
    bool multi = false;
    std::vector<int> devices;  // multi && empty: all
};
DeviceSpec device_spec_from(const py::object &device) {
    DeviceSpec spec;
    if (device.is_none()) {
        spec.devices.push_back(default_device());
    } else if (py::isinstance<py::str>(device)) {
        if (py::cast<std::string>(device) != "all") throw std::invalid_argument("device must be an int, a list of ints or 'all'");
        spec.multi = true;
    } else if (py::isinstance<py::list>(device) || py::isinstance<py::tuple>(device)) {
        spec.multi = true;
        for (auto item : device) spec.devices.push_back(py::cast<int>(item));
        if (spec.devices.empty()) throw std::invalid_argument("device list must not be empty");
    } else {
        spec.devices.push_back(py::cast<int>(device));
    }
    return spec;
}

// Owns one decoder, or one decoder replicated over several GPUs (cb200_compile_multi).
struct Handle {
    cb200_decoder *dec = nullptr;  // the decoder that answers queries (replica 0 of a multi handle)
    cb200_multi *multi = nullptr;
    Handle(const std::string &dem_text, const DeviceSpec &spec) {
        int rc;
        if (spec.multi) {
            rc = cb200_compile_multi(dem_text.c_str(), spec.devices.empty() ? nullptr : spec.devices.data(), (int)spec.devices.size(), &multi);
            if (rc == CB200_OK) dec = cb200_multi_replica(multi, 0);
        } else {
            rc = cb200_compile(dem_text.c_str(), spec.devices[0], &dec);
        }
        if (rc != CB200_OK) raise_last(rc);
    }
    ~Handle() {
        if (multi) cb200_multi_free(multi);
        else if (dec) cb200_free(dec);
    }
    Handle(const Handle &) = delete;
    Handle &operator=(const Handle &) = delete;
};

DeviceSpec default_spec() {
    DeviceSpec spec;
    spec.devices.push_back(default_device());
    return spec;
}

// One GPU batch; the GIL is released for the duration of the call.
void decode_rows(const Handle &h, const uint8_t *rows, uint64_t num_shots, uint64_t stride, uint8_t *obs_out, float *weight_out) {
    int rc;
    {
        py::gil_scoped_release release;
        rc = h.multi ? cb200_multi_decode_batch(h.multi, rows, num_shots, stride, obs_out, weight_out, nullptr, nullptr)
                     : cb200_decode_batch(h.dec, rows, num_shots, stride, obs_out, weight_out, nullptr, CB200_IO_HOST, nullptr);
    }
    if (rc != CB200_OK) raise_last(rc);
}

// chromobius.pybind.cc:26-118
struct CompiledDecoder {
    std::shared_ptr<Handle> h;
    uint64_t num_detectors;
    uint64_t num_detector_bytes;
    uint64_t num_observable_bytes;

    static CompiledDecoder from_text(const std::string &dem_text, const DeviceSpec &spec) {
        auto h = std::make_shared<Handle>(dem_text, spec);
        return CompiledDecoder{h, cb200_num_detectors(h->dec), cb200_num_detector_bytes(h->dec), cb200_num_observable_bytes(h->dec)};
    }

    static CompiledDecoder from_dem(const py::object &dem, const py::object &device) {
        auto type_name = py::str(py::type::of(dem));
        if (!type_name.contains("stim.") || !type_name.contains(".DetectorErrorModel")) {
            throw std::invalid_argument("dem must be a stim.DetectorErrorModel.");
        }
        return from_text(py::cast<std::string>(py::str(dem)), device_spec_from(device));
    }

    py::object predict(const py::object &dets_obj, bool include_weight) {
        if (!py::isinstance<py::array_t<uint8_t>>(dets_obj)) {
            throw std::invalid_argument("Expected bit packed detection event data, but dets.dtype wasn't np.uint8.");
        }
        py::array_t<uint8_t> dets = py::cast<py::array_t<uint8_t>>(dets_obj);
        size_t num_shots, det_shape;
        py::ssize_t shot_stride;
        py::array_t<uint8_t> result_buf;
        py::array_t<float> weight_buf;
        if (dets.ndim() == 2) {
            num_shots = dets.shape(0);
            shot_stride = dets.strides(0);
            det_shape = dets.shape(1);
            result_buf = py::array_t<uint8_t>({(py::ssize_t)num_shots, (py::ssize_t)num_observable_bytes});
            if (include_weight) weight_buf = py::array_t<float>((py::ssize_t)num_shots);
            if (dets.strides(1) != 1 && dets.shape(1) > 1) {
                std::stringstream ss;
                ss << "Bit packed shot data must be contiguous in memory, but dets.stride[1] wasn't equal to 1.\n";
                ss << "It was " << dets.strides(1) << ".";
                throw std::invalid_argument(ss.str());
            }
        } else if (dets.ndim() == 1) {
            num_shots = 1;
            shot_stride = (py::ssize_t)num_detector_bytes;
            det_shape = dets.shape(0);
            result_buf = py::array_t<uint8_t>((py::ssize_t)num_observable_bytes);
            if (include_weight) weight_buf = py::array_t<float>(std::vector<py::ssize_t>{});
            if (dets.strides(0) != 1 && dets.shape(0) > 1) dets = py::array_t<uint8_t, py::array::c_style | py::array::forcecast>(dets);
        } else {
            throw std::invalid_argument("dets.shape not in [1, 2]");
        }
        if (det_shape != num_detector_bytes) {
            std::stringstream ss;
            ss << "Expected dets.shape[1]=" << det_shape;
            ss << " == num_detector_bytes=" << num_detector_bytes;
            ss << " because dets.shape == 2 and dets.dtype==np.uint8 indicating bit packed shots.";
            throw std::invalid_argument(ss.str());
        }
        if (shot_stride < (py::ssize_t)num_detector_bytes) {
            // broadcast or reversed rows: the reference reads them through the strides, one GPU copy wants ascending rows
            dets = py::array_t<uint8_t, py::array::c_style | py::array::forcecast>(dets);
            shot_stride = (py::ssize_t)num_detector_bytes;
        }
        decode_rows(*h, dets.data(), num_shots, (uint64_t)shot_stride, result_buf.mutable_data(),
                    include_weight ? weight_buf.mutable_data() : nullptr);
        if (include_weight) return py::make_tuple(result_buf, weight_buf);
        return std::move(result_buf);
    }
};

// sinter_compat.pybind.cc:23-72
struct SinterCompiledDecoder {
    std::shared_ptr<Handle> h;
    uint64_t num_detector_bytes;
    uint64_t num_observable_bytes;

    py::array_t<uint8_t> decode_shots_bit_packed(const py::array_t<uint8_t> &data) {
        if (data.ndim() != 2) throw std::invalid_argument("bit_packed_detection_event_data.ndim() != 2");
        if (data.strides(1) != 1 && data.shape(1) > 1) throw std::invalid_argument("bit_packed_detection_event_data.strides(1) != 1");
        if (data.shape(1) != (py::ssize_t)num_detector_bytes) {
            throw std::invalid_argument("bit_packed_detection_event_data.shape(1) != num_detector_bytes");
        }
        py::array_t<uint8_t> rows = data;
        py::ssize_t stride = rows.strides(0);
        if (stride < (py::ssize_t)num_detector_bytes) {
            rows = py::array_t<uint8_t, py::array::c_style | py::array::forcecast>(data);
            stride = (py::ssize_t)num_detector_bytes;
        }
        size_t num_shots = rows.shape(0);
        py::array_t<uint8_t> out({(py::ssize_t)num_shots, (py::ssize_t)num_observable_bytes});
        decode_rows(*h, rows.data(), num_shots, (uint64_t)stride, out.mutable_data(), nullptr);
        return out;
    }
};

std::string read_text_file(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) throw std::invalid_argument("Failed to open '" + path + "'");
    std::string text;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
    fclose(f);
    return text;
}

// sinter_compat.pybind.cc:74-147
struct SinterDecoder {
    uint8_t sub_decoder;
    explicit SinterDecoder(uint8_t s) : sub_decoder(s) {}
    bool operator==(const SinterDecoder &) const { return true; }
    bool operator!=(const SinterDecoder &o) const { return !(*this == o); }

    SinterCompiledDecoder compile_decoder_for_dem(const py::object &dem) {
        // no type check on this path (sinter_compat.pybind.cc:134-136)
        auto h = std::make_shared<Handle>(py::cast<std::string>(py::str(dem)), default_spec());
        return SinterCompiledDecoder{h, cb200_num_detector_bytes(h->dec), cb200_num_observable_bytes(h->dec)};
    }

    // b8 records in, b8 records out (sinter_compat.pybind.cc:96-132); the whole file is one GPU batch.
    void decode_via_files(uint64_t num_shots, uint64_t num_dets, uint64_t num_obs, const py::object &dem_path,
                          const py::object &dets_b8_in_path, const py::object &obs_predictions_b8_out_path, const py::object &) {
        Handle h(read_text_file(py::cast<std::string>(py::str(dem_path))), default_spec());
        const uint64_t nb = (num_dets + 7) / 8, ob = (num_obs + 7) / 8, dob = cb200_num_observable_bytes(h.dec);
        if (nb != cb200_num_detector_bytes(h.dec)) throw std::invalid_argument("num_dets does not match the detector error model");
        std::vector<uint8_t> dets(num_shots * nb), pred(num_shots * dob), out(num_shots * ob, 0);
        const std::string in_path = py::cast<std::string>(py::str(dets_b8_in_path));
        FILE *fin = fopen(in_path.c_str(), "rb");
        if (!fin) throw std::invalid_argument("Failed to open '" + in_path + "'");
        size_t got = fread(dets.data(), 1, dets.size(), fin);
        fclose(fin);
        if (got != dets.size()) throw std::invalid_argument("dets_b8_in_path holds fewer than num_shots records");
        if (num_shots) decode_rows(h, dets.data(), num_shots, nb, pred.data(), nullptr);
        const uint8_t last_mask = (num_obs % 8) ? (uint8_t)((1u << (num_obs % 8)) - 1) : 0xFF;
        for (uint64_t s = 0; s < num_shots; s++) {
            for (uint64_t k = 0; k < ob && k < dob; k++) out[s * ob + k] = pred[s * dob + k];
            if (ob) out[s * ob + ob - 1] &= last_mask;
        }
        const std::string out_path = py::cast<std::string>(py::str(obs_predictions_b8_out_path));
        FILE *fout = fopen(out_path.c_str(), "wb");
        if (!fout) throw std::invalid_argument("Failed to open '" + out_path + "'");
        fwrite(out.data(), 1, out.size(), fout);
        fclose(fout);
    }
};

}  // namespace

PYBIND11_MODULE(chromobius, m) {
    m.attr("__version__") = "1.2.dev0+b200";
    m.doc() = "chromobius: the mobius color code decoder, batched on a B200 (drop-in for the reference module's decode path).";

    auto sinter_decoder = py::class_<SinterDecoder>(m, "_ChromobiusSinterDecoder", "An object that implements the sinter.Decoder API.");
    auto sinter_compiled =
        py::class_<SinterCompiledDecoder>(m, "_ChromobiusSinterCompiledDecoder", "An object that implements the sinter.CompiledDecoder API.");
    sinter_decoder.def(py::pickle([](const SinterDecoder &self) -> py::object { return py::cast(self.sub_decoder); },
                                  [](const py::object &obj) -> SinterDecoder { return SinterDecoder(py::cast<uint8_t>(obj)); }));
    sinter_decoder.def(py::self == py::self);
    sinter_decoder.def(py::self != py::self);
    sinter_decoder.def(py::init([](uint8_t sub_decoder) { return SinterDecoder(sub_decoder); }), "Creates a chromobius.ChromobiusSinterDecoder.");
    sinter_decoder.def("decode_via_files", &SinterDecoder::decode_via_files, py::kw_only(), py::arg("num_shots"), py::arg("num_dets"),
                       py::arg("num_obs"), py::arg("dem_path"), py::arg("dets_b8_in_path"), py::arg("obs_predictions_b8_out_path"),
                       py::arg("tmp_dir"), "Decodes data on disk, to disk.");
    sinter_decoder.def("compile_decoder_for_dem", &SinterDecoder::compile_decoder_for_dem, py::kw_only(), py::arg("dem"),
                       "Creates a chromobius decoder preconfigured for the given detector error model.");
    sinter_compiled.def("decode_shots_bit_packed", &SinterCompiledDecoder::decode_shots_bit_packed, py::kw_only(),
                        py::arg("bit_packed_detection_event_data"), "Predicts observable flips from the given detection events.");
    m.def(
        "sinter_decoders",
        []() -> py::object {
            auto result = py::dict();
            result["chromobius"] = SinterDecoder(0);
            return std::move(result);
        },
        "A dictionary describing chromobius to sinter: {'chromobius': <an object compatible with sinter.Decoder>}.");

    auto compiled = py::class_<CompiledDecoder>(m, "CompiledDecoder",
                                                "A chromobius decoder ready to predict observable flips from detection events.");
    compiled.def(
        "predict_obs_flips_from_dets_bit_packed",
        [](CompiledDecoder &self, const py::object &dets) { return self.predict(dets, false); }, py::arg("dets"),
        "Predicts observable flips from bit packed detection event data: uint8[num_shots, ceil(num_detectors / 8)] (or one "
        "1-D shot) -> uint8[num_shots, ceil(num_observables / 8)].");
    compiled.def(
        "predict_weighted_obs_flips_from_dets_bit_packed",
        [](CompiledDecoder &self, const py::object &dets) { return self.predict(dets, true); }, py::arg("dets"),
        "Like predict_obs_flips_from_dets_bit_packed, and also returns the float32 matching weight of every shot.");
    compiled.def_static("from_dem", &CompiledDecoder::from_dem, py::arg("dem"), py::kw_only(), py::arg("device") = py::none(),
                        "Compiles a decoder for a stim detector error model.");
    m.def(
        "compile_decoder_for_dem", [](const py::object &dem, const py::object &device) { return CompiledDecoder::from_dem(dem, device); },
        py::arg("dem"), py::kw_only(), py::arg("device") = py::none(), "Compiles a decoder for a stim detector error model.");
}
