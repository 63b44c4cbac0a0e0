// On-device detector-error-model sampler (sm_100a): the step in front of the decode path.
// Takes the role stim::DemSampler plays in the reference's tests and benchmarks
// (/root/reference/src/chromobius/decode/decoder_integration.test.cc:55-68): every error mechanism fires independently
// with its probability; a shot is the XOR of the symptoms of the mechanisms that fired.
//
// One warp per shot. Mechanisms are sorted by probability class; inside a class each lane owns a contiguous slice and
// jumps from hit to hit with geometric gaps (no per-mechanism random number). The row is assembled in shared memory
// with atomicXor and written out once, so HBM traffic is the bit-packed output itself.
// RNG: counter-based splitmix64 keyed by (seed, shot, lane, class); runs are reproducible for a given seed.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/chromobius_b200.h"
#include "host_compile.h"

using namespace cb200;

namespace {

struct ProbClass {
    double inv_log1m;  // 1 / log(1 - p)
    uint32_t start, end;
};

__device__ __forceinline__ uint64_t splitmix64(uint64_t &state) {
    uint64_t z = (state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

constexpr int SAMPLER_WARPS = 4;

__global__ void __launch_bounds__(SAMPLER_WARPS * 32) k_sample_dem(
    const ProbClass *__restrict__ classes, uint32_t num_classes, const uint32_t *__restrict__ det_ptr,
    const uint32_t *__restrict__ dets, const uint64_t *__restrict__ err_obs, uint32_t nwords, uint32_t num_det_bytes,
    uint64_t seed, uint64_t first_shot, uint32_t num_shots, uint8_t *__restrict__ out, uint64_t stride,
    uint64_t *__restrict__ obs_out) {
    extern __shared__ uint32_t smem[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *row = smem + warp * nwords;
    const uint32_t warps_total = gridDim.x * SAMPLER_WARPS;
    for (uint32_t shot = blockIdx.x * SAMPLER_WARPS + warp; shot < num_shots; shot += warps_total) {
        for (uint32_t w = lane; w < nwords; w += 32) row[w] = 0;
        __syncwarp();
        uint64_t obs = 0;
        for (uint32_t c = 0; c < num_classes; c++) {
            ProbClass pc = classes[c];
            uint64_t len = pc.end - pc.start;
            uint32_t lo = pc.start + (uint32_t)(len * lane / 32), hi = pc.start + (uint32_t)(len * (lane + 1) / 32);
            if (lo >= hi) continue;
            // One stream per (seed, shot, class, lane): the start state goes through full mixers, so that no two streams
            // are shifted copies of each other (a start state that merely ADDS shot * the splitmix increment makes draw j
            // of shot s + 1 equal draw j + 1 of shot s).
            uint64_t st = mix64(mix64(seed ^ mix64(first_shot + shot)) + (((uint64_t)c << 32) | lane));
            int64_t pos = (int64_t)lo - 1;
            while (true) {
                double u = ((double)(splitmix64(st) >> 11) + 1.0) * (1.0 / 9007199254740992.0);  // (0, 1]
                double gap = floor(log(u) * pc.inv_log1m);
                if (gap >= (double)(hi - lo)) break;
                pos += 1 + (int64_t)gap;
                if (pos >= (int64_t)hi) break;
                for (uint32_t k = det_ptr[pos]; k < det_ptr[pos + 1]; k++) {
                    uint32_t d = dets[k];
                    atomicXor(&row[d >> 5], 1u << (d & 31));
                }
                obs ^= err_obs[pos];
            }
        }
        for (int o = 16; o > 0; o >>= 1) obs ^= __shfl_xor_sync(0xFFFFFFFFu, obs, o);
        __syncwarp();
        uint8_t *dst = out + (uint64_t)shot * stride;
        const uint8_t *src = (const uint8_t *)row;
        for (uint32_t k = lane; k < num_det_bytes; k += 32) dst[k] = src[k];
        if (lane == 0 && obs_out) obs_out[shot] = obs;
        __syncwarp();
    }
}

thread_local std::string g_serr;

}  // namespace

struct cb200_sampler {
    int device = 0;
    uint32_t num_dets = 0, num_obs = 0, num_classes = 0;
    uint64_t num_errors = 0;
    double expected_events = 0;
    ProbClass *d_classes = nullptr;
    uint32_t *d_det_ptr = nullptr, *d_dets = nullptr;
    uint64_t *d_obs = nullptr;
    uint8_t *d_out = nullptr;
    uint64_t *d_obs_out = nullptr;
    size_t out_cap = 0, obs_cap = 0;
};

extern "C" {

const char *cb200_sampler_last_error(void) { return g_serr.c_str(); }

void cb200_sampler_free(cb200_sampler *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaFree(s->d_classes);
    cudaFree(s->d_det_ptr);
    cudaFree(s->d_dets);
    cudaFree(s->d_obs);
    cudaFree(s->d_out);
    cudaFree(s->d_obs_out);
    delete s;
}

int cb200_sampler_create(const char *dem_text, int device, cb200_sampler **out) {
    *out = nullptr;
    cb200_sampler *s = nullptr;
    try {
        FlatDem dem = FlatDem::parse(dem_text);
        s = new cb200_sampler();
        s->device = device;
        s->num_dets = (uint32_t)dem.num_detectors;
        s->num_obs = (uint32_t)dem.num_observables;
        struct Err {
            double p;
            std::vector<uint32_t> dets;
            uint64_t obs;
        };
        std::vector<Err> errs;
        for (const auto &e : dem.errors) {
            if (!(e.p > 0)) continue;
            Err x{e.p, {}, 0};
            for (uint32_t k = 0; k < e.num_targets; k++) {
                uint64_t t = dem.targets[e.first_target + k];
                if (t & TGT_SEPARATOR) continue;
                if (t & TGT_OBS_FLAG) x.obs ^= uint64_t{1} << ((t & ~TGT_OBS_FLAG) & 63);
                else x.dets.push_back((uint32_t)t);
            }
            std::sort(x.dets.begin(), x.dets.end());
            std::vector<uint32_t> kept;
            for (uint32_t d : x.dets) {
                if (!kept.empty() && kept.back() == d) kept.pop_back();
                else kept.push_back(d);
            }
            x.dets = kept;
            s->expected_events += e.p * (double)kept.size();
            errs.push_back(std::move(x));
        }
        std::stable_sort(errs.begin(), errs.end(), [](const Err &a, const Err &b) { return a.p < b.p; });
        std::vector<ProbClass> classes;
        std::vector<uint32_t> det_ptr{0}, dets;
        std::vector<uint64_t> obs;
        for (size_t i = 0; i < errs.size(); i++) {
            if (i == 0 || errs[i].p != errs[i - 1].p) {
                if (!classes.empty()) classes.back().end = (uint32_t)i;
                double p = std::min(errs[i].p, 1.0 - 1e-12);
                classes.push_back(ProbClass{1.0 / std::log1p(-p), (uint32_t)i, (uint32_t)i});
            }
            dets.insert(dets.end(), errs[i].dets.begin(), errs[i].dets.end());
            det_ptr.push_back((uint32_t)dets.size());
            obs.push_back(errs[i].obs);
        }
        if (!classes.empty()) classes.back().end = (uint32_t)errs.size();
        s->num_classes = (uint32_t)classes.size();
        s->num_errors = errs.size();
        auto check = [&](cudaError_t e) {
            if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e));
        };
        check(cudaSetDevice(device));
        auto up = [&](auto **dptr, const auto &vec) {
            check(cudaMalloc((void **)dptr, std::max<size_t>(1, vec.size()) * sizeof(vec[0])));
            if (!vec.empty()) check(cudaMemcpy(*dptr, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice));
        };
        up(&s->d_classes, classes);
        up(&s->d_det_ptr, det_ptr);
        up(&s->d_dets, dets);
        up(&s->d_obs, obs);
        *out = s;
        return CB200_OK;
    } catch (const std::invalid_argument &e) {
        g_serr = e.what();
        cb200_sampler_free(s);
        return CB200_ERR_INVALID_ARGUMENT;
    } catch (const std::exception &e) {
        g_serr = e.what();
        cb200_sampler_free(s);
        return CB200_ERR_CUDA;
    }
}

uint64_t cb200_sampler_num_detectors(const cb200_sampler *s) { return s->num_dets; }
uint64_t cb200_sampler_num_observables(const cb200_sampler *s) { return s->num_obs; }
double cb200_sampler_expected_events_per_shot(const cb200_sampler *s) { return s->expected_events; }

int cb200_sampler_sample(cb200_sampler *s, uint64_t num_shots, uint64_t seed, uint64_t first_shot, uint8_t *dets_out,
                         uint64_t row_stride, uint64_t *obs_out, uint32_t flags) {
    try {
        auto check = [&](cudaError_t e) {
            if (e != cudaSuccess) throw std::runtime_error(std::string("CUDA: ") + cudaGetErrorString(e));
        };
        if (num_shots == 0) return CB200_OK;
        if (num_shots > 0x7FFFFFF0ull) throw std::invalid_argument("too many shots in one call");
        const uint32_t nb = (s->num_dets + 7) / 8;
        if (row_stride < nb) throw std::invalid_argument("row_stride is smaller than the number of detector bytes");
        check(cudaSetDevice(s->device));
        const bool dev = (flags & CB200_IO_DEVICE) != 0;
        uint8_t *d_out = dets_out;
        uint64_t *d_obs = obs_out;
        uint64_t stride = row_stride;
        if (!dev) {
            size_t need = (size_t)num_shots * std::max<uint32_t>(nb, 1);
            if (need > s->out_cap) {
                cudaFree(s->d_out);
                s->d_out = nullptr;
                check(cudaMalloc((void **)&s->d_out, need));
                s->out_cap = need;
            }
            if (num_shots > s->obs_cap) {
                cudaFree(s->d_obs_out);
                s->d_obs_out = nullptr;
                check(cudaMalloc((void **)&s->d_obs_out, num_shots * 8));
                s->obs_cap = num_shots;
            }
            d_out = s->d_out;
            d_obs = s->d_obs_out;
            stride = nb;
        }
        uint32_t nwords = (s->num_dets + 31) / 32;
        size_t smem = (size_t)SAMPLER_WARPS * std::max<uint32_t>(nwords, 1) * 4;
        if (smem > 200 * 1024) throw std::invalid_argument("detector error model too large for the sampler's shared-memory row");
        check(cudaFuncSetAttribute(k_sample_dem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        uint32_t blocks = (uint32_t)std::min<uint64_t>((num_shots + SAMPLER_WARPS - 1) / SAMPLER_WARPS, 148 * 16);
        k_sample_dem<<<blocks, SAMPLER_WARPS * 32, smem>>>(s->d_classes, s->num_classes, s->d_det_ptr, s->d_dets, s->d_obs, nwords,
                                                           nb, seed, first_shot, (uint32_t)num_shots, d_out, stride, d_obs);
        check(cudaGetLastError());
        if (!dev) {
            if (nb) check(cudaMemcpy2D(dets_out, row_stride, d_out, nb, nb, num_shots, cudaMemcpyDeviceToHost));
            if (obs_out) check(cudaMemcpy(obs_out, d_obs, num_shots * 8, cudaMemcpyDeviceToHost));
        }
        check(cudaDeviceSynchronize());
        return CB200_OK;
    } catch (const std::invalid_argument &e) {
        g_serr = e.what();
        return CB200_ERR_INVALID_ARGUMENT;
    } catch (const std::exception &e) {
        g_serr = e.what();
        return CB200_ERR_CUDA;
    }
}

}  // extern "C"
