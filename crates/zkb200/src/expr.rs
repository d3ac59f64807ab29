//! `Expr::evaluations` on the device: kimchi's constraint expressions (kimchi/src/circuits/expr.rs) lowered to the RPN program
//! `zk_expr_eval_dev` runs (include/zkb200.h, "constraint evaluator").
//!
//! The prover computes, for every gate argument and every lookup constraint, `constraint.evaluations(&env)` over d4 or d8 and adds
//! the result into t4 / t8 (kimchi/src/prover.rs:794-892).  With this module the same loop reads
//!
//! ```ignore
//! let cols = DeviceColumns::upload(&ctx, &env)?;                       // once per proof: witness8, coefficients8, selectors, z, ...
//! for gate in gates { cols.accumulate(&gate.combined_constraints(&all_alphas, &mut cache), &env, &mut t4_dev, &mut t8_dev)?; }
//! ```
//!
//! and no intermediate `Evaluations` is materialised: the expression's own flat form (`Expr::to_polish`, expr.rs:1644-1649) is
//! translated token by token (`lower`), constants and challenges become literals, cells become indices into the resident columns.
use crate::{
    domain::GpuField,
    ffi::*,
    marshal::limbs_of,
    srs::{check, Ctx},
};
use ark_ff::{FftField, Zero};
use ark_poly::{EvaluationDomain, Evaluations, Radix2EvaluationDomain as D};
use core::ops::Index;
use kimchi::circuits::{
    expr::{ColumnEnvironment, ConstantTerm, Domain, Expr, ExprInner, PolishToken, Variable},
    gate::CurrOrNext,
};
use std::collections::HashMap;
use std::hash::Hash;

/// A resident array of evaluations: `domain_mult` x |d1| Montgomery field elements in device memory.
pub struct DeviceEvals {
    pub ptr: *mut core::ffi::c_void,
    pub len: u64,
    pub domain_mult: u32,
}

fn mult_of(d: Domain) -> u32 {
    d as u32 // Domain::{D1 = 1, D2 = 2, D4 = 4, D8 = 8} (expr.rs `pub enum Domain`)
}

/// The program handed to the library: tokens, literal table, column table.
#[derive(Default)]
pub struct Program {
    pub tokens: Vec<zk_expr_token>,
    pub constants: Vec<u64>, // 4 limbs per literal
    pub columns: Vec<zk_expr_column>,
}

impl Program {
    fn tok(&mut self, op: u32, arg: u32) {
        self.tokens.push(zk_expr_token { op, arg });
    }
    fn literal<F: GpuField>(&mut self, x: F) {
        let k = (self.constants.len() / 4) as u32;
        self.constants.extend_from_slice(&x.to_limbs());
        self.tok(ZK_EXPR_CONST, k);
    }
    fn column(&mut self, e: &DeviceEvals) -> u32 {
        // one table entry per distinct array
        if let Some(k) = self.columns.iter().position(|c| c.d_evals == e.ptr as *const _) {
            return k as u32;
        }
        self.columns.push(zk_expr_column { d_evals: e.ptr as *const _, len: e.len, domain_mult: e.domain_mult, reserved: 0 });
        (self.columns.len() - 1) as u32
    }
}

/// The environment's arrays, uploaded once per proof (or taken from the prover-index cache, `zk_index_cache_section`).
pub struct DeviceColumns<'c, Column: Eq + Hash> {
    ctx: &'c Ctx,
    cols: HashMap<Column, DeviceEvals>,
    vanishes: DeviceEvals,
    lagrange: HashMap<(i32, u32), DeviceEvals>, // unnormalized Lagrange bases by (offset, domain multiple)
}

impl<'c, Column: Copy + Eq + Hash> DeviceColumns<'c, Column> {
    fn upload_evals<F: GpuField>(ctx: &Ctx, e: &Evaluations<F, D<F>>, d1: u64) -> Result<DeviceEvals, String> {
        let limbs = limbs_of(&e.evals);
        let mut ptr = core::ptr::null_mut();
        check(unsafe { zk_dev_alloc(ctx.0, 8 * limbs.len(), &mut ptr) })?;
        check(unsafe { zk_dev_upload(ctx.0, ptr, limbs.as_ptr().cast(), 8 * limbs.len()) })?;
        Ok(DeviceEvals { ptr, len: e.evals.len() as u64, domain_mult: (e.evals.len() as u64 / d1) as u32 })
    }

    /// Uploads every column `columns` names (the witness, coefficient, selector, z, lookup columns of the environment).
    pub fn upload<'a, F, ChallengeTerm, Challenges, Env>(ctx: &'c Ctx, env: &Env, columns: &[Column]) -> Result<Self, String>
    where
        F: GpuField,
        Challenges: Index<ChallengeTerm, Output = F>,
        Env: ColumnEnvironment<'a, F, ChallengeTerm, Challenges, Column = Column>,
    {
        let d1 = env.get_domain(Domain::D1).size;
        let mut cols = HashMap::new();
        for c in columns {
            if let Some(e) = env.get_column(c) {
                cols.insert(*c, Self::upload_evals(ctx, e, d1)?);
            }
        }
        let vanishes = Self::upload_evals(ctx, env.vanishes_on_zero_knowledge_and_previous_rows(), d1)?;
        Ok(Self { ctx, cols, vanishes, lagrange: HashMap::new() })
    }

    /// Token-by-token translation of `PolishToken` (expr.rs:819-836) following `PolishToken::evaluate` (expr.rs:856-940) for the
    /// stack machine and `Expr::evaluations_helper` (expr.rs:1992-2160) for what a cell, a missing column and the two special
    /// atoms mean over a domain.
    pub fn lower<'a, F, ChallengeTerm, Challenges, Env>(
        &mut self,
        toks: &[PolishToken<F, Column, ChallengeTerm>],
        env: &Env,
        d: Domain,
    ) -> Result<Program, String>
    where
        F: GpuField,
        ChallengeTerm: Copy,
        Challenges: Index<ChallengeTerm, Output = F>,
        Env: ColumnEnvironment<'a, F, ChallengeTerm, Challenges, Column = Column>,
    {
        let consts = env.get_constants();
        let mut p = Program::default();
        for t in toks {
            match t {
                PolishToken::Challenge(c) => p.literal(env.get_challenges()[*c]),
                PolishToken::Constant(ConstantTerm::EndoCoefficient) => p.literal(consts.endo_coefficient),
                PolishToken::Constant(ConstantTerm::Mds { row, col }) => p.literal(consts.mds[*row][*col]),
                PolishToken::Constant(ConstantTerm::Literal(x)) => p.literal(*x),
                PolishToken::Cell(Variable { col, row }) => match self.cols.get(col) {
                    // a column the environment does not have evaluates to zero (expr.rs:2096-2101)
                    None => p.literal(F::zero()),
                    Some(e) => {
                        let k = p.column(e);
                        p.tok(ZK_EXPR_CELL, k | if matches!(row, CurrOrNext::Next) { 1 << 31 } else { 0 });
                    }
                },
                PolishToken::VanishesOnZeroKnowledgeAndPreviousRows => {
                    let k = p.column(&self.vanishes);
                    p.tok(ZK_EXPR_CELL, k);
                }
                PolishToken::UnnormalizedLagrangeBasis(i) => {
                    // the reference builds this array per use (unnormalized_lagrange_evals, expr.rs:1055-1130); built with the
                    // reference's own code on the host, once per (offset, domain), then resident
                    let offset = if i.zk_rows { -(consts.zk_rows as i32) + i.offset } else { i.offset };
                    let key = (offset, mult_of(d));
                    if !self.lagrange.contains_key(&key) {
                        let atom: Expr<ConstantTerm<F>, Column> = Expr::Atom(ExprInner::UnnormalizedLagrangeBasis(*i));
                        let evals = atom.evaluations(env); // Evals { domain: d, .. } for this atom
                        let d1 = env.get_domain(Domain::D1).size;
                        let e = Self::upload_evals(self.ctx, &evals, d1)?;
                        self.lagrange.insert(key, e);
                    }
                    let k = p.column(&self.lagrange[&key]);
                    p.tok(ZK_EXPR_CELL, k);
                }
                PolishToken::Dup => p.tok(ZK_EXPR_DUP, 0),
                PolishToken::Pow(n) => p.tok(ZK_EXPR_POW, u32::try_from(*n).map_err(|_| String::from("Pow exponent above 2^32"))?),
                PolishToken::Add => p.tok(ZK_EXPR_ADD, 0),
                PolishToken::Mul => p.tok(ZK_EXPR_MUL, 0),
                PolishToken::Sub => p.tok(ZK_EXPR_SUB, 0),
                PolishToken::Store => p.tok(ZK_EXPR_STORE, 0),
                PolishToken::Load(k) => p.tok(ZK_EXPR_LOAD, *k as u32),
                // FeatureFlag::is_enabled is `todo!()` in the reference (expr.rs:592-594): expressions reach the prover with
                // their feature flags already applied (Expr::apply_feature_flags), so these never appear here
                PolishToken::SkipIf(..) | PolishToken::SkipIfNot(..) => {
                    return Err(String::from("apply the feature flags before lowering (Expr::apply_feature_flags)"))
                }
            }
        }
        Ok(p)
    }

    /// `t += expr.evaluations(env)` with t4 / t8 resident: the domain is chosen as Expr::evaluations does (expr.rs:1949-1960).
    pub fn accumulate<'a, F, ChallengeTerm, Challenges, Env>(
        &mut self,
        expr: &Expr<ConstantTerm<F>, Column>,
        env: &Env,
        t4: &DeviceEvals,
        t8: &DeviceEvals,
    ) -> Result<(), String>
    where
        F: GpuField,
        ChallengeTerm: Copy,
        Challenges: Index<ChallengeTerm, Output = F>,
        Env: ColumnEnvironment<'a, F, ChallengeTerm, Challenges, Column = Column>,
        Column: core::fmt::Debug,
    {
        let d1_size = env.get_domain(Domain::D1).size;
        let deg = expr.degree(d1_size, env.get_constants().zk_rows);
        let (d, target) = if deg <= 4 * d1_size { (Domain::D4, t4) } else if deg <= 8 * d1_size { (Domain::D8, t8) } else {
            panic!("constraint had degree {deg} > d8 ({})", 8 * d1_size) // the reference's panic, expr.rs:1958
        };
        let toks = expr.to_polish();
        let p = self.lower(&toks, env, d)?;
        check(unsafe {
            zk_expr_eval_dev(
                self.ctx.0, F::FIELD_ID, p.tokens.as_ptr(), p.tokens.len(), p.constants.as_ptr(), p.constants.len() / 4,
                p.columns.as_ptr(), p.columns.len(), target.len, target.domain_mult, /* accumulate = */ 1, target.ptr,
            )
        })
    }
}

impl<'c, Column: Eq + Hash> Drop for DeviceColumns<'c, Column> {
    fn drop(&mut self) {
        for e in self.cols.values().chain(self.lagrange.values()).chain(core::iter::once(&self.vanishes)) {
            unsafe { zk_dev_free(self.ctx.0, e.ptr) };
        }
    }
}

/// The tail of the quotient with t4 / t8 resident (kimchi/src/prover.rs:905-918):
/// `f = t4.interpolate() + t8.interpolate() + public`, `(quotient, res) = f.divide_by_vanishing_poly(d1)`, `quotient += bnd`.
/// `public` and `bnd` are resident coefficient vectors of n and 7n elements; returns the quotient's 7n coefficients on the device, or
/// the prover's error when the remainder does not vanish.
pub fn quotient_tail<F: GpuField>(
    ctx: &Ctx,
    t4: &DeviceEvals,
    t8: &DeviceEvals,
    public: &DeviceEvals,
    bnd: &DeviceEvals,
    log_n: u32,
) -> Result<DeviceEvals, String> {
    let n = 1u64 << log_n;
    check(unsafe { zk_ntt_dev(ctx.0, F::FIELD_ID, t4.ptr, log_n + 2, 1, 0, 1, 0) })?; // t4.interpolate()
    check(unsafe { zk_ntt_dev(ctx.0, F::FIELD_ID, t8.ptr, log_n + 3, 1, 0, 1, 0) })?; // t8.interpolate()
    check(unsafe { zk_poly_add_dev(ctx.0, F::FIELD_ID, t8.ptr, t4.ptr, (4 * n) as usize) })?;
    check(unsafe { zk_poly_add_dev(ctx.0, F::FIELD_ID, t8.ptr, public.ptr, public.len as usize) })?; // f += &public_poly
    let mut q = core::ptr::null_mut();
    check(unsafe { zk_dev_alloc(ctx.0, (7 * n * 32) as usize, &mut q) })?;
    let mut zero_rem = 0;
    check(unsafe { zk_poly_divide_by_vanishing_dev(ctx.0, F::FIELD_ID, t8.ptr, (8 * n) as usize, log_n, q, &mut zero_rem) })?;
    if zero_rem == 0 {
        unsafe { zk_dev_free(ctx.0, q) };
        return Err(String::from("rest of division by vanishing polynomial")); // ProverError::Prover, prover.rs:910-914
    }
    check(unsafe { zk_poly_add_dev(ctx.0, F::FIELD_ID, q, bnd.ptr, bnd.len as usize) })?; // quotient += &bnd
    Ok(DeviceEvals { ptr: q, len: 7 * n, domain_mult: 0 })
}

