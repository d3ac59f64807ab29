//! `GpuRadix2Domain<F>`: `ark_poly::EvaluationDomain<F>` by delegation to `Radix2EvaluationDomain<F>` — same fields, same
//! (de)serialisation — with the two hot methods on field elements sent to the device (`zk_ntt_batch`).  Drop-in: the alias
//! `use ark_poly::Radix2EvaluationDomain as D` (kimchi/src/prover.rs:41, kimchi/src/circuits/domains.rs:4,
//! poly-commitment/src/lib.rs:49) becomes `use zkb200::GpuRadix2Domain as D`.
//! (Trait shape: ark-poly 0.5.0, Cargo.lock:171-279 — the crate is not vendored in the reference tree.)
use crate::{ffi::*, marshal::Limbs4, srs::{check, Ctx}};
use ark_ff::FftField;
use ark_poly::{domain::DomainCoeff, EvaluationDomain, Radix2EvaluationDomain};
use ark_serialize::{CanonicalDeserialize, CanonicalSerialize};
use core::any::TypeId;

/// field_id of the ABI for the two Pasta fields
pub trait GpuField: FftField + Limbs4 {
    const FIELD_ID: i32;
}
impl GpuField for mina_curves::pasta::Fp {
    const FIELD_ID: i32 = ZK_FP;
}
impl GpuField for mina_curves::pasta::Fq {
    const FIELD_ID: i32 = ZK_FQ;
}

#[derive(Copy, Clone, Hash, Eq, PartialEq, Debug, CanonicalSerialize, CanonicalDeserialize)]
pub struct GpuRadix2Domain<F: FftField>(pub Radix2EvaluationDomain<F>);

impl<F: GpuField> GpuRadix2Domain<F> {
    /// Below this size the PCIe round trip costs more than the CPU transform
    pub const MIN_LOG_SIZE: u32 = 12;

    fn device_transform(&self, v: &mut Vec<F>, inverse: bool) {
        let n = self.0.size();
        let in_len = v.len().min(n);
        v.resize(n, F::zero());                                   // ark: coeffs.resize(self.size(), T::zero())
        let mut limbs: Vec<u64> = Vec::with_capacity(4 * n);
        for x in v.iter() {
            limbs.extend_from_slice(&x.to_limbs());
        }
        let coset = (self.0.coset_offset() != F::one()) as i32;
        assert!(coset == 0 || self.0.coset_offset() == F::GENERATOR, "only the plain domain and the default coset are on the device");
        let ctx = Ctx::global();
        check(unsafe {
            zk_ntt_batch(ctx.0, F::FIELD_ID, limbs.as_mut_ptr(), self.0.log_size_of_group, 1, if inverse { 0 } else { in_len }, inverse as i32, coset)
        })
        .expect("zkb200: ntt");
        for (x, l) in v.iter_mut().zip(limbs.chunks_exact(4)) {
            *x = F::from_limbs([l[0], l[1], l[2], l[3]]);
        }
    }
}

impl<F: GpuField> EvaluationDomain<F> for GpuRadix2Domain<F> {
    type Elements = <Radix2EvaluationDomain<F> as EvaluationDomain<F>>::Elements;

    fn new(num_coeffs: usize) -> Option<Self> {
        Radix2EvaluationDomain::new(num_coeffs).map(Self)
    }
    fn get_coset(&self, offset: F) -> Option<Self> {
        self.0.get_coset(offset).map(Self)
    }
    fn compute_size_of_domain(num_coeffs: usize) -> Option<usize> {
        Radix2EvaluationDomain::<F>::compute_size_of_domain(num_coeffs)
    }
    fn size(&self) -> usize {
        self.0.size()
    }
    fn log_size_of_group(&self) -> u64 {
        self.0.log_size_of_group()
    }
    fn size_as_field_element(&self) -> F {
        self.0.size_as_field_element()
    }
    fn size_inv(&self) -> F {
        self.0.size_inv()
    }
    fn group_gen(&self) -> F {
        self.0.group_gen()
    }
    fn group_gen_inv(&self) -> F {
        self.0.group_gen_inv()
    }
    fn coset_offset(&self) -> F {
        self.0.coset_offset()
    }
    fn coset_offset_inv(&self) -> F {
        self.0.coset_offset_inv()
    }
    fn coset_offset_pow_size(&self) -> F {
        self.0.coset_offset_pow_size()
    }
    fn elements(&self) -> Self::Elements {
        self.0.elements()
    }

    /// T == F and a domain worth the PCIe round trip: the device; group elements (the Lagrange-basis iFFT of ipa.rs:1161, which
    /// `GpuSRS` replaces by `zk_srs_lagrange_basis`) and tiny domains: arkworks.
    fn fft_in_place<T: DomainCoeff<F>>(&self, coeffs: &mut Vec<T>) {
        if TypeId::of::<T>() == TypeId::of::<F>() && self.0.log_size_of_group >= Self::MIN_LOG_SIZE && self.0.log_size_of_group <= 30 {
            // SAFETY: T and F are the same type
            let v: &mut Vec<F> = unsafe { &mut *(coeffs as *mut Vec<T>).cast::<Vec<F>>() };
            self.device_transform(v, false);
        } else {
            self.0.fft_in_place(coeffs);
        }
    }
    fn ifft_in_place<T: DomainCoeff<F>>(&self, evals: &mut Vec<T>) {
        if TypeId::of::<T>() == TypeId::of::<F>() && self.0.log_size_of_group >= Self::MIN_LOG_SIZE && self.0.log_size_of_group <= 30 {
            let v: &mut Vec<F> = unsafe { &mut *(evals as *mut Vec<T>).cast::<Vec<F>>() };
            self.device_transform(v, true);
        } else {
            self.0.ifft_in_place(evals);
        }
    }
}
