//! Raw bindings to `include/diffurch_b200.h` (ABI version 2).  One `#[repr(C)]` struct per C struct,
//! field for field; see the header for the reference interface each item replaces.
//! NOT compiled in this repository (no Rust toolchain in the build image).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const DFB_ABI_VERSION: i32 = 3;
pub const DFB_MAX_STAGES: usize = 7;
pub const DFB_MAX_INTERP: usize = 5;
pub const DFB_MAX_STATE: usize = 12;
pub const DFB_MAX_PARAMS: usize = 12;
pub const DFB_MAX_AUX: usize = 4;
pub const DFB_MAX_LOCATORS: usize = 4;
pub const DFB_MAX_DISCO: usize = 8;
pub const DFB_MAX_FILTER_TIMES: usize = 8;
pub const DFB_SYS_USER_BASE: i32 = 1000;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct dfb_tableau {
    pub S: i32,
    pub I: i32,
    pub order: i32,
    pub order_embedded: i32,
    pub order_interpolant: i32,
    pub _pad: i32,
    pub a: [f64; DFB_MAX_STAGES * DFB_MAX_STAGES],
    pub b: [f64; DFB_MAX_STAGES],
    pub b2: [f64; DFB_MAX_STAGES],
    pub c: [f64; DFB_MAX_STAGES],
    pub bi: [f64; DFB_MAX_STAGES * DFB_MAX_INTERP],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct dfb_locator {
    pub kind: i32,
    pub detector_id: i32,
    pub detection: i32,
    pub location: i32,
    pub callback_id: i32,
    pub dedup: i32,
    pub smoothing_order: i32,
    pub filter: i32,
    pub filter_n: i32,
    pub _pad: i32,
    pub period: f64,
    pub offset: f64,
    pub delay: f64,
    pub filter_a: f64,
    pub filter_b: f64,
    pub filter_times: [i32; DFB_MAX_FILTER_TIMES],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct dfb_auto_stepsize {
    pub stepsize: f64,
    pub stepsize_min: f64,
    pub stepsize_max: f64,
    pub atol: [f64; DFB_MAX_STATE],
    pub rtol: [f64; DFB_MAX_STATE],
    pub order: u32,
    pub _pad: u32,
    pub fac: f64,
    pub fac_min: f64,
    pub fac_max: f64,
    pub initial_stepsize: f64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct dfb_problem {
    pub abi_version: i32,
    pub system_id: i32,
    pub history_id: i32,
    pub arith: i32,
    pub rk: dfb_tableau,
    pub t0: f64,
    pub t1: f64,
    pub stepsize_kind: i32,
    pub n_loc: i32,
    pub h: f64,
    pub automatic: dfb_auto_stepsize,
    pub max_delay: f64,
    pub n_disco: i32,
    pub on_start_callback_id: i32,
    pub disco_t: [f64; DFB_MAX_DISCO],
    pub disco_order: [i32; DFB_MAX_DISCO],
    pub loc: [dfb_locator; DFB_MAX_LOCATORS],
    pub trace_every: i32,
    pub trace_capacity: i32,
    pub event_capacity: i32,
    pub ring_capacity: i32,
    pub max_steps: i64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct dfb_system_info {
    pub n_state: i32,
    pub n_params: i32,
    pub n_aux: i32,
    pub n_detectors: i32,
    pub n_callbacks: i32,
    pub uses_history: i32,
}

#[repr(C)]
pub struct dfb_stats {
    pub steps: *mut u64,
    pub rejected: *mut u64,
    pub events: *mut u64,
    pub rhs_evals: *mut u64,
    pub status: *mut u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct dfb_totals {
    pub n_traj: u64,
    pub steps: u64,
    pub rejected: u64,
    pub events: u64,
    pub rhs_evals: u64,
    pub n_failed: u64,
    pub kernel_ms: f64,
    pub kernel_launches: u32,
    pub kernel_kind: u32,
    pub locate_iters: u64,
}

#[repr(C)]
pub struct dfb_handle {
    _private: [u8; 0],
}

extern "C" {
    pub fn dfb_abi_version() -> c_int;
    pub fn dfb_last_error() -> *const c_char;
    pub fn dfb_tableau_by_name(name: *const c_char, out: *mut dfb_tableau) -> c_int;
    pub fn dfb_tableau_generic_order_2(c2: f64, out: *mut dfb_tableau) -> c_int;
    pub fn dfb_tableau_generic_order_3(c2: f64, c3: f64, has_b3: c_int, b3: f64, out: *mut dfb_tableau) -> c_int;
    pub fn dfb_problem_default(p: *mut dfb_problem) -> c_int;
    pub fn dfb_system_info_get(system_id: i32, out: *mut dfb_system_info) -> c_int;
    pub fn dfb_register_system_from_source(cuda_source: *const c_char, struct_name: *const c_char, system_id_out: *mut i32) -> c_int;
    pub fn dfb_nvrtc_stats(kernels_compiled: *mut u32, disk_cache_hits: *mut u32, compile_ms_total: *mut f64, last_compile_ms: *mut f64) -> c_int;
    pub fn dfb_nvrtc_prebuild(system_id: i32, arith: i32, rk: *const dfb_tableau, members_per_thread: i32, trace: i32) -> c_int;
    pub fn dfb_register_plugin(so_path: *const c_char, system_id_out: *mut i32) -> c_int;
    pub fn dfb_create(desc: *const dfb_problem, n_traj: usize, device: c_int, out: *mut *mut dfb_handle) -> c_int;
    pub fn dfb_destroy(h: *mut dfb_handle);
    pub fn dfb_set_initial(h: *mut dfb_handle, y0: *const f64) -> c_int;
    pub fn dfb_set_params(h: *mut dfb_handle, params: *const f64) -> c_int;
    pub fn dfb_set_initial_device(h: *mut dfb_handle, d_y0: *const f64) -> c_int;
    pub fn dfb_set_params_device(h: *mut dfb_handle, d_params: *const f64) -> c_int;
    pub fn dfb_run(h: *mut dfb_handle, stream: *mut c_void) -> c_int;
    pub fn dfb_synchronize(h: *mut dfb_handle) -> c_int;
    pub fn dfb_get_final(h: *mut dfb_handle, t: *mut f64, y: *mut f64, aux: *mut f64) -> c_int;
    pub fn dfb_get_stats(h: *mut dfb_handle, out: *mut dfb_stats) -> c_int;
    pub fn dfb_get_totals(h: *mut dfb_handle, out: *mut dfb_totals) -> c_int;
    pub fn dfb_final_device(h: *mut dfb_handle, d_t: *mut *const f64, d_y: *mut *const f64, d_aux: *mut *const f64) -> c_int;
    pub fn dfb_get_trace(h: *mut dfb_handle, n_samples: *mut u32, t: *mut f64, y: *mut f64) -> c_int;
    pub fn dfb_get_events(h: *mut dfb_handle, n_events: *mut u32, t: *mut f64, index: *mut i32) -> c_int;
    pub fn dfb_gather_final(h: *mut dfb_handle, nccl_comm: *mut c_void, n_ranks: c_int, t_all: *mut f64, y_all: *mut f64, aux_all: *mut f64) -> c_int;
    pub fn dfb_measure_fp64_peak(device: c_int, tflops: *mut f64, sm_mhz_est: *mut f64) -> c_int;
}
