//! Limb-level marshalling between arkworks values and the ABI's raw form (include/zkb200.h "Conventions"):
//! field element = 4 x u64 little-endian MONTGOMERY limbs — exactly `Fp<MontBackend<_, 4>, 4>.0.0`, the representation the
//! reference itself reads and writes as raw limbs in kimchi/src/cached_prover_index.rs:486-530; affine point = x || y, the
//! identity all zeros.
use ark_ec::{short_weierstrass::{Affine, SWCurveConfig}, AffineRepr};
use ark_ff::{BigInt, Fp, MontBackend, MontConfig, Zero};

/// A 256-bit Montgomery field of arkworks (both Pasta fields).
pub trait Limbs4: Sized + Copy {
    fn to_limbs(&self) -> [u64; 4];
    fn from_limbs(l: [u64; 4]) -> Self;
}
impl<T: MontConfig<4>> Limbs4 for Fp<MontBackend<T, 4>, 4> {
    #[inline]
    fn to_limbs(&self) -> [u64; 4] {
        (self.0).0
    }
    #[inline]
    fn from_limbs(l: [u64; 4]) -> Self {
        // the limbs ARE the Montgomery representation: no conversion (ark_ff::Fp::new_unchecked)
        Fp::new_unchecked(BigInt(l))
    }
}

/// Copy of a slice of field elements as one flat limb vector (Montgomery).
pub fn limbs_of<F: Limbs4>(v: &[F]) -> Vec<u64> {
    let mut out = Vec::with_capacity(4 * v.len());
    for x in v {
        out.extend_from_slice(&x.to_limbs());
    }
    out
}
pub fn fields_of<F: Limbs4>(l: &[u64]) -> Vec<F> {
    l.chunks_exact(4).map(|c| F::from_limbs([c[0], c[1], c[2], c[3]])).collect()
}

/// x || y in Montgomery limbs; the identity is eight zeros.
pub fn point_limbs<P: SWCurveConfig>(p: &Affine<P>) -> [u64; 8]
where
    P::BaseField: Limbs4,
{
    let mut out = [0u64; 8];
    if let Some((x, y)) = p.xy() {
        out[..4].copy_from_slice(&x.to_limbs());
        out[4..].copy_from_slice(&y.to_limbs());
    }
    out
}
pub fn points_limbs<P: SWCurveConfig>(ps: &[Affine<P>]) -> Vec<u64>
where
    P::BaseField: Limbs4,
{
    let mut out = Vec::with_capacity(8 * ps.len());
    for p in ps {
        out.extend_from_slice(&point_limbs(p));
    }
    out
}
pub fn point_of<P: SWCurveConfig>(l: &[u64]) -> Affine<P>
where
    P::BaseField: Limbs4,
{
    if l[..8].iter().all(|w| *w == 0) {
        return Affine::<P>::zero();
    }
    // the library returns points of the group: no on-curve / subgroup re-check (cofactor 1)
    Affine::<P>::new_unchecked(
        P::BaseField::from_limbs([l[0], l[1], l[2], l[3]]),
        P::BaseField::from_limbs([l[4], l[5], l[6], l[7]]),
    )
}
