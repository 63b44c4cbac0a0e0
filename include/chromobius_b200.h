/* chromobius_b200.h -- C ABI of the B200 batched colour-code decoder.
 *
 * This is the drop-in boundary for ONE path of quantumlib/chromobius: the per-shot decode loop behind
 * `CompiledDecoder.predict_obs_flips_from_dets_bit_packed`. The reference has no FFI of its own; its boundary is the
 * C++ `chromobius::Decoder` behind pybind11. Every entry point below names the reference interface it replaces
 * (paths relative to /root/reference/src/chromobius). A binding that keeps the reference's Python surface on top of
 * this ABI is shown in INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes, no exceptions, no global state besides a thread-local error string.
 * Every function that can fail returns 0 on success and a negative CB200_ERR_* code otherwise;
 * cb200_last_error() then describes the failure (the reference throws std::invalid_argument with that text).
 * A decoder handle owns device memory on one GPU and must not be used from two threads at once
 * (same rule as the reference's Decoder, decode/decoder.h:64-144).
 */
#ifndef CHROMOBIUS_B200_H
#define CHROMOBIUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CB200_OK 0
#define CB200_ERR_INVALID_ARGUMENT (-1) /* what the reference reports as std::invalid_argument */
#define CB200_ERR_CUDA (-2)             /* CUDA runtime failure (no GPU, out of memory, launch error) */
#define CB200_ERR_DECODE (-3)           /* a shot could not be decoded; see cb200_batch_stats.first_failed_shot */

/* cb200_decode_batch flags */
#define CB200_IO_HOST 0u   /* dets / outputs are host pointers; copies are part of the call */
#define CB200_IO_DEVICE 1u /* dets / outputs are device pointers on the decoder's GPU */

typedef struct cb200_decoder cb200_decoder;

typedef struct cb200_batch_stats {
    uint64_t num_shots;
    uint64_t num_detection_events; /* after masking ignored detectors and pad bits */
    uint64_t num_retried_shots;    /* shots re-run with enlarged scratch */
    uint64_t num_failed_shots;
    int64_t first_failed_shot;     /* -1 if none */
    uint32_t first_failed_status;  /* cb200 ShotStatus of that shot */
    uint32_t max_events_per_shot;
    uint32_t group_threads;        /* threads cooperating on one shot in the matcher kernel (32 = warp, else block) */
    uint32_t num_groups;
    uint64_t kernel_launches;      /* kernels launched by this call */
    float ms_h2d, ms_count, ms_decode, ms_d2h, ms_total; /* CUDA-event times on the decoder's stream */
    uint64_t matcher_rounds;       /* sum over shots of certified-matching rounds */
} cb200_batch_stats;

/* Thread-local description of the most recent failure in this thread. */
const char *cb200_last_error(void);

/* Compile a decoder from detector-error-model text and upload its tables to `device`.
 * Replaces Decoder::from_dem (decode/decoder.cc:24-69) as reached from CompiledDecoder::from_dem
 * (pybind/chromobius.pybind.cc:32-47), which also hands the DEM over as text. */
int cb200_compile(const char *dem_text, int device, cb200_decoder **out);
/* Models the reference accepts and this library rejects with CB200_ERR_INVALID_ARGUMENT (the reference hands them to
 * PyMatching, which flips negative-weight edges and tolerates zero weights):
 *   - a Moebius error with probability above 0.5 (negative edge weight), and
 *   - an edge whose discretised weight rounds to 0 (p = 0.5, or a weight below 2^-24 of the largest one).
 * Other limits: at most 64 observables (as the reference), at most 2^24 Moebius nodes, and matching components of at
 * most 2000 detection events per shot (a larger one fails the call with CB200_ERR_DECODE, status 6). */

/* Releases host and device memory. */
void cb200_free(cb200_decoder *dec);

/* Decoder::node_colors.size() / num_observables, as exposed through pybind/chromobius.pybind.cc:40-46. */
uint64_t cb200_num_detectors(const cb200_decoder *dec);
uint64_t cb200_num_observables(const cb200_decoder *dec);
uint64_t cb200_num_detector_bytes(const cb200_decoder *dec);   /* ceil(num_detectors / 8) */
uint64_t cb200_num_observable_bytes(const cb200_decoder *dec); /* ceil(num_observables / 8) */
/* PymatchingMatcher::weight_scale = flooder.graph.normalising_constant (decode/pymatcher.cc:23). */
double cb200_normalising_constant(const cb200_decoder *dec);

/* Text dump of the compiled tables in the layout of Decoder's operator<< (decode/decoder.cc:279-291),
 * i.e. what `chromobius describe_decoder` prints. Returns a malloc'd string; free with cb200_free_string. */
char *cb200_describe(const cb200_decoder *dec);
void cb200_free_string(char *s);

/* Decode a batch of bit-packed shots.
 * Replaces the serial loop over Decoder::decode_detection_events (decode/decoder.cc:249-277) in
 * CompiledDecoder::predict_obs_flips_from_dets_bit_packed (pybind/chromobius.pybind.cc:49-117) and in
 * ChromobiusSinterCompiledDecoder::decode_shots_bit_packed (pybind/sinter_compat.pybind.cc:30-72).
 *
 *   dets            num_shots rows of cb200_num_detector_bytes() bytes, `row_stride` bytes apart
 *                   (bit d of a shot = (row[d >> 3] >> (d & 7)) & 1). Pad bits and ignored detectors are masked.
 *   obs_out         num_shots x cb200_num_observable_bytes() bytes, little-endian observable mask per shot
 *   weight_out      nullable; float32 per shot = integer matching weight / normalising constant
 *                   (PymatchingMatcher::match_edges, decode/pymatcher.cc:32-41)
 *   weight_int_out  nullable; int64 per shot, the exact integer matching weight (parity hook)
 *   flags           CB200_IO_HOST or CB200_IO_DEVICE
 *   stats           nullable
 * A shot that cannot be decoded fails the whole call with CB200_ERR_DECODE, as the reference's exception does.
 * Large host batches (three chunks of 96 MB or more) are streamed: the decode kernel starts after the first chunk and a
 * second stream delivers the rest while it runs. That needs launches and copies to overlap, so it is switched off when
 * CUDA_LAUNCH_BLOCKING=1 or a profiler / sanitizer is injected into the process (or with CB200_NO_PIPELINE=1). */
int cb200_decode_batch(cb200_decoder *dec, const uint8_t *dets, uint64_t num_shots, uint64_t row_stride, uint8_t *obs_out,
                       float *weight_out, int64_t *weight_int_out, uint32_t flags, cb200_batch_stats *stats);

/* Decode one host shot and also return the matched graph edges (the MatcherInterface::match_edges contract,
 * decode/matcher_interface.h:29-45: interleaved node pairs of the Moebius graph). Test / debugging hook.
 * `*num_edge_ints` receives the number of int64 values the full list needs (may exceed edges_cap). */
int cb200_decode_shot_edges(cb200_decoder *dec, const uint8_t *dets_row, int64_t *edges_out, uint64_t edges_cap,
                            uint64_t *num_edge_ints, uint64_t *obs_mask_out, int64_t *weight_int_out);

/* Tuning knobs (0 keeps the automatic choice): threads per shot in the decode kernel (32, 128, 192, 256 or 512) and the
 * log2 size of the per-warp shared-memory ball table. */
int cb200_set_tuning(cb200_decoder *dec, uint32_t group_threads, uint32_t ht_log2);
/* Caps the number of shots in flight per SM (0 = as many as fit). Fewer shots in flight keep the per-shot scratch
 * inside L2. */
int cb200_set_groups_per_sm(cb200_decoder *dec, uint32_t groups_per_sm);

/* Phase profiling of the decode kernel (clock64 of thread 0 of every group, summed over shots). Enables/disables it
 * for later calls and copies out the counters of the most recent profiled call:
 * [0] ingest [1] first ball growth [2] component building [3] component solves [4] ball re-growth
 * [5] path extraction [6] Euler tours + discharge [7] number of shots
 * [8..11] blossom cycles for components of <= 8, <= 16, <= 32, > 32 events, [12..15] how many of each
 * [16] closed-form component solves [17] per-warp blossoms [18] block-wide blossoms (the three parts of [3]). */
int cb200_profile_phases(cb200_decoder *dec, int enable, unsigned long long *cycles_out24);

/* Sizes of the device-resident tables, for DESIGN.md / bench reporting. */
uint64_t cb200_table_bytes(const cb200_decoder *dec);
uint64_t cb200_graph_half_edges(const cb200_decoder *dec);

/* ---- One decoder replicated over several GPUs of one box ---------------------------------------------------------
 * BASELINE's largest configuration shards 10 M shots over 8 B200; shots are independent, so a batch is split into
 * contiguous row ranges, one per device, each decoded by that device's replica on its own host thread and stream.
 * There is no data-path collective: every range's predictions land directly in the caller's host buffers. This keeps
 * the reference's single call (CompiledDecoder::predict_obs_flips_from_dets_bit_packed, pybind/chromobius.pybind.cc:49-117)
 * a single call on N GPUs. The model is compiled once on the host; every device receives a copy of the tables.
 *   devices / num_devices   GPU indices to use (NULL / 0: every visible GPU)
 * cb200_multi_decode_batch takes HOST pointers only; `stats` (nullable) sums the per-device counters, with the times
 * being the maximum over devices. cb200_multi_replica() exposes the per-device decoders (queries, tuning). */
typedef struct cb200_multi cb200_multi;
int cb200_compile_multi(const char *dem_text, const int *devices, int num_devices, cb200_multi **out);
void cb200_multi_free(cb200_multi *m);
int cb200_multi_num_devices(const cb200_multi *m);
cb200_decoder *cb200_multi_replica(cb200_multi *m, int index);
int cb200_multi_decode_batch(cb200_multi *m, const uint8_t *dets, uint64_t num_shots, uint64_t row_stride, uint8_t *obs_out,
                             float *weight_out, int64_t *weight_int_out, cb200_batch_stats *stats);

/* ---- On-device syndrome sampler: the step in front of the decode path --------------
 * Plays the role of stim::DemSampler in the reference's tests (decode/decoder_integration.test.cc:55-68): each error
 * mechanism of the DEM fires independently with its probability, a shot is the XOR of the fired symptoms.
 * Output layout equals the decoder's input layout, so sampled shots can be decoded without leaving HBM. */
typedef struct cb200_sampler cb200_sampler;
const char *cb200_sampler_last_error(void);
int cb200_sampler_create(const char *dem_text, int device, cb200_sampler **out);
void cb200_sampler_free(cb200_sampler *s);
uint64_t cb200_sampler_num_detectors(const cb200_sampler *s);
uint64_t cb200_sampler_num_observables(const cb200_sampler *s);
double cb200_sampler_expected_events_per_shot(const cb200_sampler *s);
/* Shots [first_shot, first_shot + num_shots) of the stream identified by `seed`.
 * dets_out: num_shots rows, row_stride bytes apart; obs_out (nullable): one u64 observable mask per shot.
 * flags: CB200_IO_HOST or CB200_IO_DEVICE for both outputs. */
int cb200_sampler_sample(cb200_sampler *s, uint64_t num_shots, uint64_t seed, uint64_t first_shot, uint8_t *dets_out,
                         uint64_t row_stride, uint64_t *obs_out, uint32_t flags);

#ifdef __cplusplus
}
#endif

#endif
