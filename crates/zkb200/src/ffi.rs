//! `extern "C"` declarations of include/zkb200.h (the subset the shim calls; `bindgen include/zkb200.h` gives the rest).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_uint, c_void};

#[repr(C)]
pub struct zk_ctx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct zk_bases {
    _p: [u8; 0],
}
#[repr(C)]
pub struct zk_srs {
    _p: [u8; 0],
}

pub const ZK_FP: c_int = 0;
pub const ZK_FQ: c_int = 1;
pub const ZK_PALLAS: c_int = 0;
pub const ZK_VESTA: c_int = 1;
pub const ZK_OK: c_int = 0;
pub const ZK_ERR_LENGTH: c_int = -4;

#[repr(C)]
pub struct zk_open_poly {
    pub data: *const u64,
    pub len: usize,
    pub domain_size: usize,
    pub blinders: *const u64,
    pub n_blinders: usize,
}

#[repr(C)]
pub struct zk_open_transcript {
    pub user: *mut c_void,
    pub u_base: unsafe extern "C" fn(user: *mut c_void, cip: *const u64, out_u_xy: *mut u64) -> c_int,
    pub round: unsafe extern "C" fn(user: *mut c_void, round: c_uint, l_xy: *const u64, r_xy: *const u64, out_u: *mut u64) -> c_int,
    pub final_challenge: unsafe extern "C" fn(user: *mut c_void, delta_xy: *const u64, out_c: *mut u64) -> c_int,
}

pub const ZK_EXPR_CONST: u32 = 0;
pub const ZK_EXPR_CELL: u32 = 1;
pub const ZK_EXPR_DUP: u32 = 2;
pub const ZK_EXPR_POW: u32 = 3;
pub const ZK_EXPR_ADD: u32 = 4;
pub const ZK_EXPR_MUL: u32 = 5;
pub const ZK_EXPR_SUB: u32 = 6;
pub const ZK_EXPR_STORE: u32 = 7;
pub const ZK_EXPR_LOAD: u32 = 8;

#[repr(C)]
#[derive(Copy, Clone)]
pub struct zk_expr_token {
    pub op: u32,
    pub arg: u32,
}

#[repr(C)]
#[derive(Copy, Clone)]
pub struct zk_expr_column {
    pub d_evals: *const c_void,
    pub len: u64,
    pub domain_mult: u32,
    pub reserved: u32,
}

extern "C" {
    pub fn zk_last_error() -> *const c_char;
    pub fn zk_ctx_create(device_id: c_int, out: *mut *mut zk_ctx) -> c_int;
    pub fn zk_ctx_destroy(ctx: *mut zk_ctx);

    pub fn zk_srs_create(ctx: *mut zk_ctx, curve_id: c_int, g_xy: *const u64, n: usize, h_xy: *const u64, window_bits: c_int,
                         out: *mut *mut zk_srs) -> c_int;
    pub fn zk_srs_destroy(srs: *mut zk_srs);
    pub fn zk_srs_lagrange_basis(srs: *mut zk_srs, domain_size: usize, window_bits: c_int) -> c_int;
    pub fn zk_srs_lagrange_basis_chunks(srs: *const zk_srs, domain_size: usize) -> usize;
    pub fn zk_srs_get_lagrange_basis(srs: *mut zk_srs, domain_size: usize, out_xy: *mut u64, capacity_points: usize) -> c_int;
    pub fn zk_srs_commit_non_hiding(srs: *mut zk_srs, coeffs_mont: *const u64, len: usize, num_chunks: usize, out_xy: *mut u64,
                                    out_capacity: usize, out_chunks: *mut usize) -> c_int;
    pub fn zk_srs_commit_evaluations_non_hiding(srs: *mut zk_srs, domain_size: usize, evals_mont: *const u64, evals_domain_size: usize,
                                                out_xy: *mut u64) -> c_int;
    pub fn zk_srs_mask_custom(srs: *mut zk_srs, chunks_xy: *const u64, n_chunks: usize, blinders_mont: *const u64, n_blinders: usize,
                              out_xy: *mut u64) -> c_int;
    pub fn zk_srs_open(srs: *mut zk_srs, polys: *const zk_open_poly, n_polys: usize, elm_mont: *const u64, n_elm: usize,
                       polyscale: *const u64, evalscale: *const u64, rng_scalars: *const u64, n_rng_scalars: usize,
                       transcript: *const zk_open_transcript, out_lr_xy: *mut u64, lr_capacity_rounds: usize, out_rounds: *mut usize,
                       out_delta_xy: *mut u64, out_z1: *mut u64, out_z2: *mut u64, out_sg_xy: *mut u64) -> c_int;

    pub fn zk_dev_alloc(ctx: *mut zk_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn zk_dev_free(ctx: *mut zk_ctx, d_ptr: *mut c_void) -> c_int;
    pub fn zk_dev_upload(ctx: *mut zk_ctx, d_dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    pub fn zk_dev_download(ctx: *mut zk_ctx, dst: *mut c_void, d_src: *const c_void, bytes: usize) -> c_int;
    pub fn zk_expr_eval_dev(ctx: *mut zk_ctx, field_id: c_int, tokens: *const zk_expr_token, n_tokens: usize, constants_mont: *const u64,
                            n_constants: usize, cols: *const zk_expr_column, n_cols: usize, out_len: u64, out_domain_mult: c_uint,
                            accumulate: c_int, d_out: *mut c_void) -> c_int;

    pub fn zk_ntt_dev(ctx: *mut zk_ctx, field_id: c_int, d_data: *mut c_void, log_n: c_uint, batch: usize, in_len: usize, inverse: c_int,
                      coset: c_int) -> c_int;
    pub fn zk_ntt_dev_oop(ctx: *mut zk_ctx, field_id: c_int, d_in: *const c_void, in_stride: usize, in_len: usize, d_out: *mut c_void,
                          log_n: c_uint, batch: usize, inverse: c_int, coset: c_int) -> c_int;
    pub fn zk_poly_add_dev(ctx: *mut zk_ctx, field_id: c_int, d_dst: *mut c_void, d_src: *const c_void, len: usize) -> c_int;
    pub fn zk_poly_divide_by_vanishing_dev(ctx: *mut zk_ctx, field_id: c_int, d_f: *const c_void, len: usize, log_n: c_uint,
                                           d_quot: *mut c_void, remainder_is_zero: *mut c_int) -> c_int;

    pub fn zk_ntt_batch(ctx: *mut zk_ctx, field_id: c_int, data: *mut u64, log_n: c_uint, batch: usize, in_len: usize, inverse: c_int,
                        coset: c_int) -> c_int;
}
