/* include/tp3.h — C ABI of the B200-native 3photons hot path (libtp3.so).
 *
 * The reference (HadrienG2/3photons-rust) has no FFI: its narrowest seam is the
 * kernel closure handed to the scheduler,
 *     scheduling::run_simulation(num_events, simulate_events)      src/scheduling/mod.rs:31-34
 *     simulate_events = |num_events, &mut rng| -> ResultsAccumulator  src/main.rs:103-128
 * This header is what a Rust `-sys` crate would bind in place of that closure
 * (see INTEGRATION.md for the extern "C" block).  Plain pointers and sizes only;
 * no C++/torch types; every entry point returns 0 on success or a TP3_E_* code and
 * never throws or aborts across the boundary.  A context is not re-entrant: one
 * host thread drives it.
 *
 * There is NO CPU fallback: every compute entry point fails with TP3_E_CUDA /
 * TP3_E_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef TP3_H
#define TP3_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TP3_ABI_VERSION 2

/* Size of an event batch — src/scheduling/mod.rs:21 (EVENT_BATCH_SIZE) */
#define TP3_EVENT_BATCH_SIZE 10000u

/* Feature flags: the reference's cargo features (Cargo.toml:9-21), as run-time flags */
#define TP3_F32               (1u << 0) /* numeric.rs:6-13        single precision           */
#define TP3_FASTER_EVGEN      (1u << 1) /* evgen.rs:143-173       unit-disc rejection evgen   */
#define TP3_FASTER_THREADING  (1u << 2) /* multi_threading.rs:66-69  per-batch rng.jump()     */
#define TP3_MULTI_THREADING   (1u << 3) /* scheduling/mod.rs:4-7  (host scheduling only)      */
#define TP3_NO_PHOTON_SORTING (1u << 4) /* evgen.rs:109, event.rs:96                          */
#define TP3_STANDARD_RANDOM   (1u << 5) /* random/standard.rs     xoshiro256+ / xoshiro128+   */

/* Kernel variants (tp3_params.kernel) */
#define TP3_KERNEL_FAST    0u /* algebraically reduced amplitudes, FMA, survivor compaction */
#define TP3_KERNEL_LITERAL 1u /* operation-for-operation transcription of spinor.rs/matelems.rs */

/* Error codes */
#define TP3_OK            0
#define TP3_E_INVALID     1 /* bad argument / unsupported combination */
#define TP3_E_NO_DEVICE   2 /* no usable sm_100 CUDA device */
#define TP3_E_CUDA        3 /* CUDA runtime error, see tp3_last_error() */
#define TP3_E_CONFIG      4 /* `valeurs` parse / validation error (config.rs:57-128) */
#define TP3_E_IO          5

/* Everything the per-event kernel reads.  Replaces the closure's captured state
 * (&cfg, &couplings, &evgen; main.rs:103-115).  All constants are computed by the
 * host in the run's Float precision and widened exactly to double. */
typedef struct tp3_params {
    uint64_t num_events_total;   /* cfg.num_events (config.rs:10); only for bookkeeping */
    double   e_total;            /* config.rs:13                                        */
    double   acut, bcut, e_min, sincut; /* EventCut, evcut.rs:11-23                     */
    double   g_a, g_beta_p, g_beta_m;   /* Couplings, coupling.rs:10-19                 */
    double   sigma_contribs[5];  /* ResultsAccumulator::new, resacc.rs:94-100           */
    uint32_t flags;              /* TP3_F32 | TP3_FASTER_EVGEN | ...                    */
    uint32_t kernel;             /* TP3_KERNEL_*                                        */
} tp3_params;

/* One ResultsAccumulator (resacc.rs:16-34): 13 scalars, 104 bytes.
 * Under TP3_F32 the device accumulates in f32 and widens exactly. */
typedef struct tp3_acc {
    uint64_t selected_events;
    double   spm2[5];
    double   vars[5];
    double   sigma;
    double   variance;
} tp3_acc;

typedef struct tp3_ctx tp3_ctx;

/* ---- life cycle ------------------------------------------------------------------ */
int  tp3_abi_version(void);
/* n_dev devices (dev_ids may be NULL = 0..n_dev-1) driven from this one host thread. */
int  tp3_create(const tp3_params* params, int n_dev, const int* dev_ids, tp3_ctx** out);
void tp3_destroy(tp3_ctx* ctx);
const char* tp3_last_error(const tp3_ctx* ctx); /* ctx may be NULL: last create error */
/* Launch on a caller-owned CUDA stream (cudaStream_t cast to void*) of device slot i. */
int  tp3_set_stream(tp3_ctx* ctx, int dev_slot, void* cuda_stream);

/* ---- the hot path: replaces simulate_events over a RANGE of batches ---------------- */
/* Simulates batches [first_batch, first_batch+n_batches) of the run; every batch has
 * TP3_EVENT_BATCH_SIZE events except the last of the range, which has last_batch_len
 * (1..10000; pass 10000 for a full one).  Batch b's random stream is positioned exactly
 * where the sequential reference has it (sequential.rs:24-36), or after b jump()s under
 * TP3_FASTER_THREADING (multi_threading.rs:66-69).  Writes one accumulator per batch to
 * the caller-owned HOST array (device->host copy included).  With several devices the
 * range is split into contiguous sub-ranges, one per device. */
int  tp3_simulate_batches(tp3_ctx* ctx, uint64_t first_batch, uint64_t n_batches,
                          uint32_t last_batch_len, tp3_acc* out_per_batch);
/* Same launch, but results stay in device memory (asynchronous; for HBM-resident timing). */
int  tp3_simulate_batches_device(tp3_ctx* ctx, uint64_t first_batch, uint64_t n_batches,
                                 uint32_t last_batch_len);
/* Copy the accumulators of the last *_device launch to the host and synchronise. */
int  tp3_fetch(tp3_ctx* ctx, tp3_acc* out_per_batch, uint64_t n_batches);
/* Same range, but the per-batch accumulators are left-folded in batch order ON DEVICE
 * (ResultsAccumulator::merge, resacc.rs:133-139; the fold of sequential.rs:24-36 and
 * multi_threading.rs:107-126) and only the merged one comes back (104 bytes device->host).
 * The fold runs inside the simulation kernel, in strict batch order, so the result is bit-identical
 * to tp3_simulate_batches + tp3_fold_batches for any launch shape.  With several devices in the
 * context each device folds its contiguous sub-range and the per-device results are merged in
 * device order: reproducible for a fixed device count (the reference's order-insensitive
 * FastAccumulator, multi_threading.rs:130-190, makes the same trade). */
int  tp3_simulate_merged(tp3_ctx* ctx, uint64_t first_batch, uint64_t n_batches,
                         uint32_t last_batch_len, tp3_acc* out_merged);
/* Both at once: one accumulator per batch in the caller's host array (copied while the kernel runs when the range
 * is long) AND the left fold of those accumulators in batch order, done by the same launch on the device -- the host does
 * not need to fold (sequential.rs:24-36, multi_threading.rs:107-126).  On one device *out_merged equals
 * tp3_fold_batches(out_per_batch) bit for bit; with several devices see tp3_simulate_merged. */
int  tp3_simulate_batches_merged(tp3_ctx* ctx, uint64_t first_batch, uint64_t n_batches, uint32_t last_batch_len,
                                 tp3_acc* out_per_batch, tp3_acc* out_merged);
/* Asynchronous form for multi-GPU runs (single-device contexts): the merged accumulator is left in
 * the caller's DEVICE buffer as 13 doubles {selected_events, spm2[5], vars[5], sigma, variance}
 * (the count is exact as a double below 2^53), i.e. the operand of one ncclReduce(sum) over the
 * ranks -- the whole inter-GPU exchange of a run. */
int  tp3_simulate_merged_device(tp3_ctx* ctx, uint64_t first_batch, uint64_t n_batches,
                                uint32_t last_batch_len, double* device_out13);
/* faster-evgen on several GPUs (sequential RANF stream; single-device contexts).  Under this feature the position of
 * an event in the random stream depends on every earlier event, so a rank cannot start at "its" batches without walking
 * everything before them (the reference's scheduler thread does exactly that: evgen.rs:257-267, multi_threading.rs:59-64).
 * Instead the STREAM is sharded: this call simulates every event that STARTS in RANF rounds
 * [first_round, first_round + n_rounds) (a round = 55 numbers, ranf.rs:106-119; multiples of 64), at most max_events of
 * them (0 = no limit; n_rounds = 0 = no round limit), in groups of 10 000 consecutive events counted from the tile's first
 * event, and leaves the merged accumulator in the caller's device buffer as 13 doubles (see tp3_simulate_merged_device).
 * *events_done = the number of events simulated (synchronises the walk, not the physics).  Consecutive tiles partition
 * the events of the run exactly: ranks take one tile each, exchange their event counts (8 bytes), the last rank adds
 * the remaining events with a second call limited by max_events, and one reduce(sum) of 13 doubles finishes the run
 * (run_simulation_tiles in the Python mirror).  The events, the selected-event count and every sum are those of the
 * sequential run up to the order of the additions. */
int  tp3_fe_tile_device(tp3_ctx* ctx, uint64_t first_round, uint64_t n_rounds, uint64_t max_events,
                        double* device_out13, uint64_t* events_done);
/* Host left fold of per-batch accumulators in batch order, starting FROM the first one
 * (sequential.rs:24-36), in the run's Float (flags & TP3_F32). */
int  tp3_fold_batches(const tp3_acc* per_batch, uint64_t n_batches, uint32_t flags, tp3_acc* out);
int  tp3_synchronize(tp3_ctx* ctx);
/* Number of kernel launches issued by this context so far. */
uint64_t tp3_launch_count(const tp3_ctx* ctx);
/* Bytes of kernel arguments one launch of the fused kernel carries host->device (the only input). */
size_t tp3_kernel_arg_bytes(void);

/* ---- test / A-B switches ------------------------------------------------------------
 * Explicit, per context (nothing is read from the environment).  None of them changes a result
 * beyond the order of floating-point additions, and none selects a CPU path for the simulation:
 *   "unit_batches"     consecutive batches per scheduling unit of the fused kernel (0 = by launch size)
 *   "grid_warps"       warps in the grid (0 = as many as the device holds at once)
 *   "sched_dynamic"    1 (default): one unit per warp, dispatched by the hardware; 0: static balanced schedule
 *   "ramp_units"       dynamic schedule: that many units of 1, 2, .., 8 batches at the head of the launch (0 = none, the default)
 *   "taper_units"      dynamic schedule: units per taper stage (half units, then quarter units) between the big units and the
 *                      single batches at the end of a launch (0 = half a wave each, the default; -1 = no taper stages)
 *   "tail_singles"     dynamic schedule: single-batch units at the end of the launch (0 = by launch size, the default)
 *   "align_units"      dynamic schedule: 1 (default): the number of big units is rounded down to a multiple of 4 x SMs, so that every
 *                      sub-partition of the device leaves the big units in the same state; 0: as the sizes fall
 *   "batch_parts"      sequential RANF stream, default event generator: every batch is cut into that many equal parts (1 = whole
 *                      batches, the default; 2, 5, 10; 0 = the largest of those for which the launch still fits one
 *                      wave of resident warps), one warp each, and the parts of a batch are added in part order: a small run
 *                      (the default 1e7 events are 1000 batches for 2368 warp slots) then fills the device.  Same events,
 *                      same selected-event counts; the sums of a batch differ between settings by the order of additions,
 *                      so a caller that compares sub-ranges bit for bit keeps one setting.  tp3_run uses 0.
 *   "f32_scalar"       f32: one event per lane instead of the packed two-events-per-lane kernel
 *   "fe_split"         faster-evgen: 1 = one thread per batch, 32 = one lane per 313 events, 0 = by launch size
 *   "fe_host_scan"     faster-evgen: batch start states from the reference's own method, the event-by-event
 *                      walk of the master generator on the host (evgen.rs:257-267), instead of the GPU scans:
 *                      the cross-check of those scans
 *   "fe_legacy"        faster-evgen on the sequential RANF stream: 1 = the round-1 pipeline (scan over transition maps +
 *                      batch kernel with private generators) instead of the stream pipeline (walk -> records -> physics)
 *   "fe_pass_segments" stream pipeline: segments (= lanes of the walk) per pass (0 = one full wave of the device)
 *   "fe_seg_rounds"    stream pipeline: RANF rounds per segment (0 = 64 .. 512 by run size)
 *   "fe_warm"          stream pipeline: warm-up rounds before a segment (0 = 24); small values force the redo path
 *   "fe_serial"        stream pipeline: 1 = passes one after the other (the walk of pass k + 1 not next to the physics of pass k)
 *   "fe_xo_seg_units"  faster-evgen + xoshiro scan: segment length in units of 2048 outputs (0 = by launch size)
 *   "fe_timing"        print the scan phases of every call to stderr */
int  tp3_set_option(tp3_ctx* ctx, const char* name, int64_t value);
/* "fe_xo_pass_b": how many times pass B of the last xoshiro faster-evgen scan ran; "fe_passes", "fe_redone": passes of the
 * last stream-pipeline call and segments it had to redo. */
int  tp3_get_stat(tp3_ctx* ctx, const char* name, int64_t* value);

/* ---- per-event observables (the reference's empty hook) ---------------------------------------------
 * The reference parses `num_bins` (config.rs:45-46,102) and marks, without implementing them, the places
 * where the FORTRAN code filled and normalised its histograms (main.rs:117-122,133).  This fills them in the
 * fused kernel: every selected event adds its weight  w = m . sigma_contribs  (the quantity
 * ResultsAccumulator::integrate adds to sigma, resacc.rs:126-128) to one bin of each of
 * TP3_HIST_OBSERVABLES distributions, photons in the event's order (decreasing energy unless
 * TP3_NO_PHOTON_SORTING, evgen.rs:109-118):
 *    observable k     (k = 0,1,2):  x_k    = 2 E_k / e_total       in [0, 1]
 *    observable 3 + k (k = 0,1,2):  cos(theta_k) = p_x,k / E_k     in [-1, 1]   (the beam is along X, evgen.rs:66-72)
 * with num_bins uniform bins over the stated range, bin = min(num_bins - 1, floor(t * num_bins)), t the value
 * mapped to [0, 1].  Histograms accumulate over all simulate calls of the context (all devices) until reset.
 * Event counts are exact; the weight sums are accumulated with floating-point reductions (order not fixed:
 * equal to ~1e-13 relative between runs).  Available for the fast kernel: default event generator (any generator and
 * seeding) and faster-evgen on the sequential RANF stream. */
#define TP3_HIST_OBSERVABLES 6
#define TP3_HIST_MAX_BINS 1024
/* num_bins = 0 switches the epilogue off again. */
int  tp3_histograms_enable(tp3_ctx* ctx, uint32_t num_bins);
int  tp3_histograms_reset(tp3_ctx* ctx);
/* counts / weights: [TP3_HIST_OBSERVABLES][num_bins], summed over the devices; synchronises. */
int  tp3_histograms_fetch(tp3_ctx* ctx, uint64_t* counts, double* weights);

/* ---- parity hooks ------------------------------------------------------------------ */
/* Raw integer random stream of batch `batch`, in the order the reference consumes it,
 * produced by the SAME per-warp/per-lane stream code the simulation kernel uses.
 * RANF words are the i32 values (ranf.rs:95-101) zero-extended; xoshiro words are the
 * raw u64 (f64) / u32 (f32) outputs.  n_words <= 12 * TP3_EVENT_BATCH_SIZE. */
int  tp3_rng_dump(tp3_ctx* ctx, uint64_t batch, uint32_t n_words, uint64_t* out_words);
/* Per-event outputs of the first n events of batch `batch`: sorted photon momenta
 * [n][3][4] (X,Y,Z,E), cut decision [n], helicity-summed matrix elements [n][5]. */
int  tp3_events_dump(tp3_ctx* ctx, uint64_t batch, uint32_t n, double* momenta, int32_t* kept,
                     double* m2_sums);

/* The hand-written FP64 functions of the fast kernel, evaluated on the device: out[i] = f(in[i]).
 * which: 0 -log x, 1 sin(2 pi x), 2 cos(2 pi x), 3 sqrt x, 4 1/x, 5 1/sqrt x, 6 sqrt x (paired form),
 *        7 raw MUFU.RCP64H seed, 8 raw MUFU.RSQ64H seed, 9 / 10 (u32)x * 1e-9 / 256e-9 by the single-FMA route. */
int  tp3_fastmath_probe(tp3_ctx* ctx, int which, uint32_t n, const double* in, double* out);

/* ---- measurement ------------------------------------------------------------------- */
/* Sustained FMA throughput of the FP64 (which=0) or FP32 (which=1) pipe, TFLOP/s. */
int  tp3_peak_probe(tp3_ctx* ctx, int which, double* tflops);

/* ---- host side of the reference surface (config.rs, coupling.rs, resacc.rs, resfin.rs,
 *      output.rs), so a caller can go valeurs -> res.data without the Rust crate ----- */
typedef struct tp3_config {        /* Configuration, config.rs:8-53 (values widened to double) */
    uint64_t num_events;
    double   e_total, beam_photons_cut, photon_photon_cut, e_min, beam_photon_plane_cut;
    double   alpha, alpha_z, gev2_to_picobarn, m_z0, g_z0, sin2_weinberg, branching_ep_em;
    double   beta_plus, beta_minus;
    int32_t  num_bins;
    int32_t  impr, plot;
} tp3_config;

typedef struct tp3_final {         /* FinalResults, resfin.rs:26-62 */
    uint64_t selected_events;
    double   spm2[2][5], vars[2][5];
    double   sigma, prec, variance, beta_min, ss_p, inc_ss_p, ss_m, inc_ss_m;
} tp3_final;

/* Parse the text of a `valeurs` file (config.rs:57-128). err_buf receives the message. */
int  tp3_config_parse(const char* valeurs_text, uint32_t flags, tp3_config* out, char* err_buf,
                      size_t err_cap);
/* Couplings::new + EventGenerator::new + ResultsAccumulator::new -> kernel parameters. */
int  tp3_params_from_config(const tp3_config* cfg, uint32_t flags, uint32_t kernel, tp3_params* out);
/* ResultsAccumulator::merge (resacc.rs:133-139), in the run's Float precision. */
int  tp3_merge(tp3_acc* into, const tp3_acc* other, uint32_t flags);
/* ResultsAccumulator::finalize (resacc.rs:142-223). */
int  tp3_finalize(const tp3_config* cfg, uint32_t flags, const tp3_acc* merged, tp3_final* out);
/* Text of res.data (output.rs:64-142) and of the stdout report (config.rs:131-155,
 * evgen.rs:47, resfin.rs:66-194).  Return the number of bytes needed (excl. NUL). */
size_t tp3_format_res_data(const tp3_config* cfg, uint32_t flags, const tp3_final* fin, char* buf,
                           size_t cap);
size_t tp3_format_stdout(const tp3_config* cfg, uint32_t flags, const tp3_final* fin, char* buf,
                         size_t cap);
/* The whole program (main.rs:75-145) on n_dev GPUs: reads valeurs_path, writes res.data,
 * res.times and appends pil.mc in out_dir, returns the stdout text in stdout_buf. */
int  tp3_run(const char* valeurs_path, const char* out_dir, uint32_t flags, uint32_t kernel, int n_dev,
             char* stdout_buf, size_t stdout_cap, double* elapsed_seconds);
/* The same, with the wall time of its stages in stages[TP3_RUN_STAGES] (seconds): 0 read + parse `valeurs`,
 * 1 context creation (device tables, streams), 2 simulation (launch to merged accumulator on the host),
 * 3 finalize, 4 formatting + writing res.data / res.times / pil.mc, 5 context destruction.  The reference's own
 * timed region (main.rs:83-85,138) is stages 1 + 2 + 3 = *elapsed_seconds. */
#define TP3_RUN_STAGES 6
int  tp3_run_stages(const char* valeurs_path, const char* out_dir, uint32_t flags, uint32_t kernel, int n_dev,
                    char* stdout_buf, size_t stdout_cap, double* elapsed_seconds, double* stages);

/* Host mirror of the device RANF jump-ahead (no GPU needed): state of the generator's
 * 55-word round `round` (0 = after seeding + warm-up), numbers[1..55] -> out[0..54]. */
int  tp3_host_ranf_round(int32_t seed, uint64_t round, uint32_t* out55);
/* Host mirror of the xoshiro jump-ahead: state after `n_steps` outputs from
 * seed_from_u64(12345); f32 != 0 selects xoshiro128+ (4 x u32 widened). */
int  tp3_host_xoshiro_state(int f32, uint64_t n_steps, uint64_t n_jumps, uint64_t* out4);

#ifdef __cplusplus
}
#endif
#endif /* TP3_H */
