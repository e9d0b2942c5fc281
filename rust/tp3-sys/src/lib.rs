//! Raw bindings to `include/tp3.h`. UNCOMPILED in this repository (no Rust toolchain in the image); the symbol names and
//! arities of the extern block are held to the header by tests/test_host_surface.py::test_rust_sys_crate_matches_header.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const TP3_ABI_VERSION: i32 = 2;
pub const TP3_EVENT_BATCH_SIZE: u32 = 10_000; // scheduling/mod.rs:21

pub const TP3_F32: u32 = 1 << 0;
pub const TP3_FASTER_EVGEN: u32 = 1 << 1;
pub const TP3_FASTER_THREADING: u32 = 1 << 2;
pub const TP3_MULTI_THREADING: u32 = 1 << 3;
pub const TP3_NO_PHOTON_SORTING: u32 = 1 << 4;
pub const TP3_STANDARD_RANDOM: u32 = 1 << 5;

pub const TP3_KERNEL_FAST: u32 = 0;
pub const TP3_KERNEL_LITERAL: u32 = 1;

pub const TP3_OK: c_int = 0;

/// `tp3_params`: everything the per-event kernel reads (the closure's captured state, main.rs:103-115).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct tp3_params {
    pub num_events_total: u64,
    pub e_total: f64,
    pub acut: f64,
    pub bcut: f64,
    pub e_min: f64,
    pub sincut: f64,
    pub g_a: f64,
    pub g_beta_p: f64,
    pub g_beta_m: f64,
    pub sigma_contribs: [f64; 5],
    pub flags: u32,
    pub kernel: u32,
}

/// `tp3_acc`: the 13 sums of one `ResultsAccumulator` (resacc.rs:19-34).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct tp3_acc {
    pub selected_events: u64,
    pub spm2: [f64; 5],
    pub vars: [f64; 5],
    pub sigma: f64,
    pub variance: f64,
}

#[repr(C)]
pub struct tp3_ctx {
    _private: [u8; 0],
}

/// `tp3_config`: `Configuration` (config.rs:8-53), values widened to f64.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct tp3_config {
    pub num_events: u64,
    pub e_total: f64,
    pub beam_photons_cut: f64,
    pub photon_photon_cut: f64,
    pub e_min: f64,
    pub beam_photon_plane_cut: f64,
    pub alpha: f64,
    pub alpha_z: f64,
    pub gev2_to_picobarn: f64,
    pub m_z0: f64,
    pub g_z0: f64,
    pub sin2_weinberg: f64,
    pub branching_ep_em: f64,
    pub beta_plus: f64,
    pub beta_minus: f64,
    pub num_bins: i32,
    pub impr: i32,
    pub plot: i32,
}

/// `tp3_final`: `FinalResults` (resfin.rs:26-62).
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct tp3_final {
    pub selected_events: u64,
    pub spm2: [[f64; 5]; 2],
    pub vars: [[f64; 5]; 2],
    pub sigma: f64,
    pub prec: f64,
    pub variance: f64,
    pub beta_min: f64,
    pub ss_p: f64,
    pub inc_ss_p: f64,
    pub ss_m: f64,
    pub inc_ss_m: f64,
}

pub const TP3_HIST_OBSERVABLES: usize = 6;
pub const TP3_HIST_MAX_BINS: u32 = 1024;

extern "C" {
    pub fn tp3_abi_version() -> c_int;
    pub fn tp3_create(params: *const tp3_params, n_dev: c_int, dev_ids: *const c_int, out: *mut *mut tp3_ctx) -> c_int;
    pub fn tp3_destroy(ctx: *mut tp3_ctx);
    pub fn tp3_last_error(ctx: *const tp3_ctx) -> *const c_char;
    pub fn tp3_set_stream(ctx: *mut tp3_ctx, dev_slot: c_int, cuda_stream: *mut c_void) -> c_int;
    pub fn tp3_simulate_batches(ctx: *mut tp3_ctx, first_batch: u64, n_batches: u64, last_batch_len: u32,
                                out_per_batch: *mut tp3_acc) -> c_int;
    pub fn tp3_simulate_batches_device(ctx: *mut tp3_ctx, first_batch: u64, n_batches: u64, last_batch_len: u32) -> c_int;
    pub fn tp3_fetch(ctx: *mut tp3_ctx, out_per_batch: *mut tp3_acc, n_batches: u64) -> c_int;
    pub fn tp3_simulate_merged(ctx: *mut tp3_ctx, first_batch: u64, n_batches: u64, last_batch_len: u32,
                               out_merged: *mut tp3_acc) -> c_int;
    pub fn tp3_simulate_batches_merged(ctx: *mut tp3_ctx, first_batch: u64, n_batches: u64, last_batch_len: u32,
                                       out_per_batch: *mut tp3_acc, out_merged: *mut tp3_acc) -> c_int;
    pub fn tp3_simulate_merged_device(ctx: *mut tp3_ctx, first_batch: u64, n_batches: u64, last_batch_len: u32,
                                      device_out13: *mut f64) -> c_int;
    pub fn tp3_fe_tile_device(ctx: *mut tp3_ctx, first_round: u64, n_rounds: u64, max_events: u64, device_out13: *mut f64,
                              events_done: *mut u64) -> c_int;
    pub fn tp3_fold_batches(per_batch: *const tp3_acc, n_batches: u64, flags: u32, out: *mut tp3_acc) -> c_int;
    pub fn tp3_synchronize(ctx: *mut tp3_ctx) -> c_int;
    pub fn tp3_launch_count(ctx: *const tp3_ctx) -> u64;
    pub fn tp3_kernel_arg_bytes() -> usize;
    pub fn tp3_set_option(ctx: *mut tp3_ctx, name: *const c_char, value: i64) -> c_int;
    pub fn tp3_get_stat(ctx: *mut tp3_ctx, name: *const c_char, value: *mut i64) -> c_int;
    /// Per-event observables, the hook main.rs:117-122,133 leaves empty: [TP3_HIST_OBSERVABLES][num_bins] histograms.
    pub fn tp3_histograms_enable(ctx: *mut tp3_ctx, num_bins: u32) -> c_int;
    pub fn tp3_histograms_reset(ctx: *mut tp3_ctx) -> c_int;
    pub fn tp3_histograms_fetch(ctx: *mut tp3_ctx, counts: *mut u64, weights: *mut f64) -> c_int;
    pub fn tp3_rng_dump(ctx: *mut tp3_ctx, batch: u64, n_words: u32, out_words: *mut u64) -> c_int;
    pub fn tp3_events_dump(ctx: *mut tp3_ctx, batch: u64, n: u32, momenta: *mut f64, kept: *mut i32, m2_sums: *mut f64) -> c_int;
    pub fn tp3_fastmath_probe(ctx: *mut tp3_ctx, which: c_int, n: u32, input: *const f64, out: *mut f64) -> c_int;
    pub fn tp3_peak_probe(ctx: *mut tp3_ctx, which: c_int, tflops: *mut f64) -> c_int;
    // host side of the reference surface (config.rs, coupling.rs, resacc.rs, resfin.rs, output.rs)
    pub fn tp3_config_parse(valeurs_text: *const c_char, flags: u32, out: *mut tp3_config, err_buf: *mut c_char, err_cap: usize) -> c_int;
    pub fn tp3_params_from_config(cfg: *const tp3_config, flags: u32, kernel: u32, out: *mut tp3_params) -> c_int;
    pub fn tp3_merge(into: *mut tp3_acc, other: *const tp3_acc, flags: u32) -> c_int;
    pub fn tp3_finalize(cfg: *const tp3_config, flags: u32, merged: *const tp3_acc, out: *mut tp3_final) -> c_int;
    pub fn tp3_format_res_data(cfg: *const tp3_config, flags: u32, fin: *const tp3_final, buf: *mut c_char, cap: usize) -> usize;
    pub fn tp3_format_stdout(cfg: *const tp3_config, flags: u32, fin: *const tp3_final, buf: *mut c_char, cap: usize) -> usize;
    pub fn tp3_run(valeurs_path: *const c_char, out_dir: *const c_char, flags: u32, kernel: u32, n_dev: c_int,
                   stdout_buf: *mut c_char, stdout_cap: usize, elapsed_seconds: *mut f64) -> c_int;
    pub fn tp3_run_stages(valeurs_path: *const c_char, out_dir: *const c_char, flags: u32, kernel: u32, n_dev: c_int,
                          stdout_buf: *mut c_char, stdout_cap: usize, elapsed_seconds: *mut f64, stages: *mut f64) -> c_int;
    pub fn tp3_host_ranf_round(seed: i32, round: u64, out55: *mut u32) -> c_int;
    pub fn tp3_host_xoshiro_state(f32_: c_int, n_steps: u64, n_jumps: u64, out4: *mut u64) -> c_int;
}
