//! B200 back end for Kimchi's two proving-time hot paths, behind the reference's own seams:
//!
//! * [`GpuSRS`] implements `poly_commitment::SRS<G>` (poly-commitment/src/lib.rs:61-241) — every MSM-bearing method runs on the
//!   device, the rest follows `ipa::SRS` (poly-commitment/src/ipa.rs:596-800);
//! * [`GpuOpeningProof`] implements `poly_commitment::OpenProof<G, FULL_ROUNDS>` (lib.rs:254-298) with the serde layout of
//!   `ipa::OpeningProof` (ipa.rs:1175-1191): `open` is one `zk_srs_open` call, `verify` is the reference's verifier;
//! * [`GpuRadix2Domain`] implements `ark_poly::EvaluationDomain<F>` by delegation to `Radix2EvaluationDomain`, with
//!   `fft_in_place` / `ifft_in_place` on field elements sent to `zk_ntt_batch`.
//!
//! * [`expr::DeviceColumns`] runs `Expr::evaluations` (kimchi/src/circuits/expr.rs:1938-2190) — the gate and lookup constraints of
//!   the quotient — as RPN programs over device-resident columns (`zk_expr_eval_dev`).
//!
//! Everything called is declared in include/zkb200.h and exported by libzkb200.so; there is no CPU fallback inside the library
//! (`Ctx::new` fails without a CUDA device) — code that must also run without a GPU keeps using `ipa::SRS`.
pub mod domain;
pub mod expr;
pub mod ffi;
pub mod marshal;
pub mod open;
pub mod srs;

pub use domain::GpuRadix2Domain;
pub use open::GpuOpeningProof;
pub use srs::{Ctx, GpuCurve, GpuSRS};
