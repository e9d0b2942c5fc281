//! `scheduling::run_simulation` on B200s (UNCOMPILED here). In the reference crate this file would be
//! `src/scheduling/gpu.rs`, selected by a cargo feature next to `sequential.rs` and `multi_threading.rs`;
//! `Configuration`, `Couplings`, `EventGenerator`, `ResultsAccumulator` are the reference's own types.
use std::ffi::CStr;
use tp3_sys::*;

/// The 13 per-batch sums, ready to be merged in batch order (sequential.rs:24-36).
pub type BatchSums = tp3_acc;

/// What the reference passes to its closure by reference (main.rs:103-115), flattened for the kernel.
pub struct KernelInputs {
    pub num_events: usize,       // cfg.num_events
    pub e_total: f64,            // cfg.e_total
    pub cuts: [f64; 4],          // beam_photons_cut, photon_photon_cut, e_min, beam_photon_plane_cut (evcut.rs:11-23)
    pub couplings: [f64; 3],     // g_a, g_beta_p, g_beta_m (coupling.rs:23-33)
    pub sigma_contribs: [f64; 5] // ResultsAccumulator::new (resacc.rs:94-100)
}

fn feature_flags() -> u32 {
    let mut f = 0;
    if cfg!(feature = "f32") { f |= TP3_F32; }
    if cfg!(feature = "faster-evgen") { f |= TP3_FASTER_EVGEN; }
    if cfg!(all(feature = "multi-threading", feature = "faster-threading")) { f |= TP3_FASTER_THREADING; }
    if cfg!(feature = "multi-threading") { f |= TP3_MULTI_THREADING; }
    if cfg!(feature = "no-photon-sorting") { f |= TP3_NO_PHOTON_SORTING; }
    if cfg!(feature = "standard-random") { f |= TP3_STANDARD_RANDOM; }
    f
}

fn kernel_params(inputs: &KernelInputs) -> tp3_params {
    tp3_params {
        num_events_total: inputs.num_events as u64,
        e_total: inputs.e_total,
        acut: inputs.cuts[0], bcut: inputs.cuts[1], e_min: inputs.cuts[2], sincut: inputs.cuts[3],
        g_a: inputs.couplings[0], g_beta_p: inputs.couplings[1], g_beta_m: inputs.couplings[2],
        sigma_contribs: inputs.sigma_contribs,
        flags: feature_flags(),
        kernel: TP3_KERNEL_FAST,
    }
}

/// The whole run, merged in batch order on the device (tp3_simulate_merged): what `run_simulation` needs before `finalize`.
pub fn simulate_run_merged(inputs: &KernelInputs, n_gpus: i32) -> Result<BatchSums, String> {
    let params = kernel_params(inputs);
    let batch = TP3_EVENT_BATCH_SIZE as usize;
    let n_batches = (inputs.num_events + batch - 1) / batch;       // multi_threading.rs:25
    let last_len = inputs.num_events - (n_batches - 1) * batch;    // multi_threading.rs:47
    let mut ctx = std::ptr::null_mut();
    let mut out = BatchSums::default();
    unsafe {
        if tp3_create(&params, n_gpus, std::ptr::null(), &mut ctx) != TP3_OK {
            return Err(CStr::from_ptr(tp3_last_error(std::ptr::null())).to_string_lossy().into_owned());
        }
        let rc = tp3_simulate_merged(ctx, 0, n_batches as u64, last_len as u32, &mut out);
        let err = if rc != TP3_OK { Some(CStr::from_ptr(tp3_last_error(ctx)).to_string_lossy().into_owned()) } else { None };
        tp3_destroy(ctx);
        if let Some(e) = err { return Err(e); }
    }
    Ok(out)
}

/// Per-batch sums in batch order AND their ordered fold from the same launch (tp3_simulate_batches_merged): nothing left to fold
/// on the host (sequential.rs:24-36).
pub fn simulate_all_batches_merged(inputs: &KernelInputs, n_gpus: i32) -> Result<(Vec<BatchSums>, BatchSums), String> {
    let params = kernel_params(inputs);
    let batch = TP3_EVENT_BATCH_SIZE as usize;
    let n_batches = (inputs.num_events + batch - 1) / batch;       // multi_threading.rs:25
    let last_len = inputs.num_events - (n_batches - 1) * batch;    // multi_threading.rs:47
    let mut ctx = std::ptr::null_mut();
    let mut out = vec![BatchSums::default(); n_batches];
    let mut merged = BatchSums::default();
    unsafe {
        if tp3_create(&params, n_gpus, std::ptr::null(), &mut ctx) != TP3_OK {
            return Err(CStr::from_ptr(tp3_last_error(std::ptr::null())).to_string_lossy().into_owned());
        }
        let rc = tp3_simulate_batches_merged(ctx, 0, n_batches as u64, last_len as u32, out.as_mut_ptr(), &mut merged);
        let err = if rc != TP3_OK { Some(CStr::from_ptr(tp3_last_error(ctx)).to_string_lossy().into_owned()) } else { None };
        tp3_destroy(ctx);
        if let Some(e) = err { return Err(e); }
    }
    Ok((out, merged))
}

/// Simulates every batch of the run on `n_gpus` devices and returns the per-batch sums in batch order.
pub fn simulate_all_batches(inputs: &KernelInputs, n_gpus: i32) -> Result<Vec<BatchSums>, String> {
    let params = kernel_params(inputs);
    let batch = TP3_EVENT_BATCH_SIZE as usize;
    let n_batches = (inputs.num_events + batch - 1) / batch;       // multi_threading.rs:25
    let last_len = inputs.num_events - (n_batches - 1) * batch;    // multi_threading.rs:47
    let mut ctx = std::ptr::null_mut();
    let mut out = vec![BatchSums::default(); n_batches];
    unsafe {
        if tp3_create(&params, n_gpus, std::ptr::null(), &mut ctx) != TP3_OK {
            return Err(CStr::from_ptr(tp3_last_error(std::ptr::null())).to_string_lossy().into_owned());
        }
        let rc = tp3_simulate_batches(ctx, 0, n_batches as u64, last_len as u32, out.as_mut_ptr());
        let err = if rc != TP3_OK { Some(CStr::from_ptr(tp3_last_error(ctx)).to_string_lossy().into_owned()) } else { None };
        tp3_destroy(ctx);
        if let Some(e) = err { return Err(e); }
    }
    Ok(out)
}
