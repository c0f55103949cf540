//! B200-native K-Means behind linfa's `Fit` / `PredictInplace` / `Transformer` traits.
//!
//! Drop-in for `linfa_clustering::KMeans` with `D = L2Dist`: same `KMeans::params(k)` builder
//! (`KMeansParams` and `KMeansValidParams` are linfa-clustering's own types, re-exported), same
//! `KMeansInit` / `KMeansAlgorithm` / `KMeansError`, same accessors.  All arithmetic runs in the
//! sm_100a CUDA library behind `include/lkm.h`; there is no CPU fallback — `fit` returns
//! `KMeansError::LinfaError` when no CUDA device is usable.
//!
//! NOTE: written against the C ABI in this repository; it could not be compiled in the authoring
//! environment (no rustc/cargo there).  The Python binding `linfa_b200/_ffi.py` binds the very
//! same symbols and is what the test-suite drives.
#![allow(non_camel_case_types)]

use std::ffi::{c_char, c_void, CStr};
use std::marker::PhantomData;

use linfa::dataset::{DatasetBase, Records};
use linfa::traits::{Fit, PredictInplace, Transformer};
use linfa::Float;
pub use linfa_clustering::{KMeansAlgorithm, KMeansError, KMeansInit, KMeansParams, KMeansParamsError, KMeansValidParams};
use linfa_nn::distance::L2Dist;
use ndarray::{Array1, Array2, ArrayBase, Data, Ix1, Ix2};
use ndarray_rand::rand::{Rng, RngCore};

// ---------------------------------------------------------------------------------------------
// FFI (include/lkm.h)
// ---------------------------------------------------------------------------------------------
#[repr(C)]
pub struct lkm_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct lkm_rng {
    user: *mut c_void,
    next_u32: unsafe extern "C" fn(*mut c_void) -> u32,
    next_u64: unsafe extern "C" fn(*mut c_void) -> u64,
}

#[repr(C)]
pub struct lkm_params {
    n_clusters: u64,
    n_runs: u64,
    max_n_iterations: u64,
    tolerance: f64,
    init_method: i32,
    algorithm: i32,
    pick_mode: i32,
    reserved: i32,
    precomputed: *const c_void,
    precomputed_rows: u64,
    precomputed_cols: u64,
}

#[repr(C)]
#[derive(Default)]
pub struct lkm_stats {
    pub n_iter: u64,
    pub last_shift: f64,
    pub inertia_total: f64,
    pub device_ms: f32,
    pub kernel_launches: u32,
    pub assign_path: u32,
    pub fallback_rows: u64,
    pub main_kernel_ms: f32,
    pub reserved: u32,
}

const LKM_OK: i32 = 0;
const LKM_ERR_INERTIA: i32 = 1;
const LKM_ERR_N_CLUSTERS: i32 = 10;
const LKM_ERR_N_RUNS: i32 = 11;
const LKM_ERR_TOLERANCE: i32 = 12;
const LKM_ERR_MAX_ITERATIONS: i32 = 13;

extern "C" {
    fn lkm_create(out: *mut *mut lkm_ctx, device: i32) -> i32;
    fn lkm_destroy(ctx: *mut lkm_ctx) -> i32;
    fn lkm_last_error(ctx: *const lkm_ctx) -> *const c_char;
    fn lkm_set_data_f32(ctx: *mut lkm_ctx, x: *const f32, n: u64, d: u64, row_stride: i64) -> i32;
    fn lkm_set_data_f64(ctx: *mut lkm_ctx, x: *const f64, n: u64, d: u64, row_stride: i64) -> i32;
    fn lkm_fit_f32(ctx: *mut lkm_ctx, p: *const lkm_params, rng: lkm_rng, c: *mut f32, cnt: *mut f32, inertia: *mut f32, st: *mut lkm_stats) -> i32;
    fn lkm_fit_f64(ctx: *mut lkm_ctx, p: *const lkm_params, rng: lkm_rng, c: *mut f64, cnt: *mut f64, inertia: *mut f64, st: *mut lkm_stats) -> i32;
    fn lkm_predict_f32(ctx: *mut lkm_ctx, c: *const f32, k: u64, x: *const f32, n: u64, d: u64, row_stride: i64, labels: *mut u64) -> i32;
    fn lkm_predict_f64(ctx: *mut lkm_ctx, c: *const f64, k: u64, x: *const f64, n: u64, d: u64, row_stride: i64, labels: *mut u64) -> i32;
    fn lkm_transform_f32(ctx: *mut lkm_ctx, c: *const f32, k: u64, x: *const f32, n: u64, d: u64, row_stride: i64, out: *mut f32) -> i32;
    fn lkm_transform_f64(ctx: *mut lkm_ctx, c: *const f64, k: u64, x: *const f64, n: u64, d: u64, row_stride: i64, out: *mut f64) -> i32;
}

/// Dispatch on `F` ∈ {f32, f64} — `linfa::Float` is implemented for exactly these two
/// (src/dataset/mod.rs:68-74).
pub trait GpuFloat: Float {
    unsafe fn set_data(ctx: *mut lkm_ctx, x: *const Self, n: u64, d: u64, stride: i64) -> i32;
    unsafe fn fit(ctx: *mut lkm_ctx, p: *const lkm_params, rng: lkm_rng, c: *mut Self, cnt: *mut Self, inertia: *mut Self, st: *mut lkm_stats) -> i32;
    unsafe fn predict(ctx: *mut lkm_ctx, c: *const Self, k: u64, x: *const Self, n: u64, d: u64, stride: i64, labels: *mut u64) -> i32;
    unsafe fn transform(ctx: *mut lkm_ctx, c: *const Self, k: u64, x: *const Self, n: u64, d: u64, stride: i64, out: *mut Self) -> i32;
}
macro_rules! gpu_float {
    ($t:ty, $set:ident, $fit:ident, $pred:ident, $tr:ident) => {
        impl GpuFloat for $t {
            unsafe fn set_data(ctx: *mut lkm_ctx, x: *const Self, n: u64, d: u64, s: i64) -> i32 { $set(ctx, x, n, d, s) }
            unsafe fn fit(ctx: *mut lkm_ctx, p: *const lkm_params, rng: lkm_rng, c: *mut Self, cnt: *mut Self, i: *mut Self, st: *mut lkm_stats) -> i32 { $fit(ctx, p, rng, c, cnt, i, st) }
            unsafe fn predict(ctx: *mut lkm_ctx, c: *const Self, k: u64, x: *const Self, n: u64, d: u64, s: i64, l: *mut u64) -> i32 { $pred(ctx, c, k, x, n, d, s, l) }
            unsafe fn transform(ctx: *mut lkm_ctx, c: *const Self, k: u64, x: *const Self, n: u64, d: u64, s: i64, o: *mut Self) -> i32 { $tr(ctx, c, k, x, n, d, s, o) }
        }
    };
}
gpu_float!(f32, lkm_set_data_f32, lkm_fit_f32, lkm_predict_f32, lkm_transform_f32);
gpu_float!(f64, lkm_set_data_f64, lkm_fit_f64, lkm_predict_f64, lkm_transform_f64);

/// RAII handle of one `lkm_ctx` (one GPU).  `Send`: a context may move between threads; it is
/// used by one thread at a time (`&mut`).
struct Context(*mut lkm_ctx);
unsafe impl Send for Context {}
impl Context {
    fn new(device: i32) -> Result<Self, KMeansError> {
        let mut h = std::ptr::null_mut();
        let st = unsafe { lkm_create(&mut h, device) };
        if st != LKM_OK {
            return Err(native_error(std::ptr::null(), st));
        }
        Ok(Context(h))
    }
}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { lkm_destroy(self.0) };
    }
}

fn native_error(ctx: *const lkm_ctx, status: i32) -> KMeansError {
    match status {
        LKM_ERR_INERTIA => KMeansError::InertiaError,
        LKM_ERR_N_CLUSTERS => KMeansParamsError::NClusters.into(),
        LKM_ERR_N_RUNS => KMeansParamsError::NRuns.into(),
        LKM_ERR_TOLERANCE => KMeansParamsError::Tolerance.into(),
        LKM_ERR_MAX_ITERATIONS => KMeansParamsError::MaxIterations.into(),
        _ => {
            let msg = if ctx.is_null() {
                "lkm_create failed: no usable CUDA device".to_string()
            } else {
                unsafe { CStr::from_ptr(lkm_last_error(ctx)) }.to_string_lossy().into_owned()
            };
            KMeansError::LinfaError(linfa::Error::Parameters(format!("lkm status {status}: {msg}")))
        }
    }
}

// trampolines into the caller's `R: Rng` (KM/hyperparams.rs:37)
unsafe extern "C" fn tramp_u32<R: RngCore>(user: *mut c_void) -> u32 {
    (*(user as *mut R)).next_u32()
}
unsafe extern "C" fn tramp_u64<R: RngCore>(user: *mut c_void) -> u64 {
    (*(user as *mut R)).next_u64()
}

/// The row stride (in elements) of a view whose rows are contiguous, else `None`.
fn row_stride<F, S: Data<Elem = F>>(x: &ArrayBase<S, Ix2>) -> Option<i64> {
    let s = x.strides();
    if x.ncols() <= 1 || s[1] == 1 {
        if x.nrows() <= 1 { Some(x.ncols() as i64) } else if s[0] >= x.ncols() as isize { Some(s[0] as i64) } else { None }
    } else {
        None
    }
}

// ---------------------------------------------------------------------------------------------
// model (KM/algorithm.rs:202-241)
// ---------------------------------------------------------------------------------------------
#[derive(Debug, Clone, PartialEq)]
pub struct KMeans<F: Float, D = L2Dist> {
    centroids: Array2<F>,
    cluster_count: Array1<F>,
    inertia: F,
    device: i32,
    dist_fn: PhantomData<D>,
}

impl<F: GpuFloat> KMeans<F, L2Dist> {
    /// `KMeans::params(k)` (KM/algorithm.rs:210-212): default rng `Xoshiro256Plus::seed_from_u64(42)`.
    pub fn params(nclusters: usize) -> GpuKMeansParams<F, rand_xoshiro::Xoshiro256Plus> {
        GpuKMeansParams(linfa_clustering::KMeans::params(nclusters), 0)
    }
    /// `KMeans::params_with_rng(k, rng)` (KM/algorithm.rs:214-216)
    pub fn params_with_rng<R: Rng>(nclusters: usize, rng: R) -> GpuKMeansParams<F, R> {
        GpuKMeansParams(linfa_clustering::KMeans::params_with_rng(nclusters, rng), 0)
    }
    pub fn centroids(&self) -> &Array2<F> {
        &self.centroids
    }
    pub fn cluster_count(&self) -> &Array1<F> {
        &self.cluster_count
    }
    pub fn inertia(&self) -> F {
        self.inertia
    }
}

/// linfa-clustering's `KMeansParams` plus the one extra knob of this crate (the CUDA device).
/// It derefs to the reference builder, so `.n_runs(..).tolerance(..).max_n_iterations(..)
/// .init_method(..).algorithm(..)` and `ParamGuard::{check, check_ref, check_unwrap}` are linfa's own.
pub struct GpuKMeansParams<F: Float, R: Rng>(KMeansParams<F, R, L2Dist>, i32);

impl<F: GpuFloat, R: Rng + Clone> GpuKMeansParams<F, R> {
    pub fn n_runs(self, n: usize) -> Self { GpuKMeansParams(self.0.n_runs(n), self.1) }
    pub fn tolerance(self, t: F) -> Self { GpuKMeansParams(self.0.tolerance(t), self.1) }
    pub fn max_n_iterations(self, m: u64) -> Self { GpuKMeansParams(self.0.max_n_iterations(m), self.1) }
    pub fn init_method(self, i: KMeansInit<F>) -> Self { GpuKMeansParams(self.0.init_method(i), self.1) }
    pub fn algorithm(self, a: KMeansAlgorithm) -> Self { GpuKMeansParams(self.0.algorithm(a), self.1) }
    /// extra knob, default 0
    pub fn device(self, device: i32) -> Self { GpuKMeansParams(self.0, device) }
}

impl<F: GpuFloat, R: Rng + Clone, DA: Data<Elem = F>, T> Fit<ArrayBase<DA, Ix2>, T, KMeansError> for GpuKMeansParams<F, R> {
    type Object = KMeans<F, L2Dist>;

    /// `Fit::fit` (KM/algorithm.rs:370-388 through src/param_guard.rs:55-66): the whole of
    /// `fit_lloyd` — seeding, Lloyd loop, best-of-n_runs, cluster_count, mean inertia — runs
    /// behind `lkm_fit_*`.  `Hamerly` is executed as Lloyd (same fixed point, KM/init.rs:46-47).
    fn fit(&self, dataset: &DatasetBase<ArrayBase<DA, Ix2>, T>) -> Result<Self::Object, KMeansError> {
        use linfa::ParamGuard;
        let v = self.0.check_ref()?; // the four KMeansParamsError, reference order
        let obs = dataset.records();
        let owned;
        let (ptr, stride) = match row_stride(obs) {
            Some(s) => (obs.as_ptr(), s),
            None => {
                owned = obs.as_standard_layout().into_owned(); // arbitrary ndarray layout -> standard
                (owned.as_ptr(), owned.ncols() as i64)
            }
        };
        let (n, d, k) = (obs.nrows() as u64, obs.ncols() as u64, v.n_clusters());
        let ctx = Context::new(self.1)?;
        let st = unsafe { F::set_data(ctx.0, ptr, n, d, stride) };
        if st != LKM_OK {
            return Err(native_error(ctx.0, st));
        }
        let (init_method, pre): (i32, Option<Array2<F>>) = match v.init_method() {
            KMeansInit::Random => (0, None),
            KMeansInit::Precomputed(c) => (1, Some(c.as_standard_layout().into_owned())),
            KMeansInit::KMeansPlusPlus => (2, None),
            KMeansInit::KMeansPara => (3, None),
            _ => (2, None),
        };
        let p = lkm_params {
            n_clusters: k as u64,
            n_runs: v.n_runs() as u64,
            max_n_iterations: v.max_n_iterations(),
            tolerance: v.tolerance().to_f64().unwrap(),
            init_method,
            algorithm: 0,
            pick_mode: 0,
            reserved: 0,
            precomputed: pre.as_ref().map_or(std::ptr::null(), |c| c.as_ptr() as *const c_void),
            precomputed_rows: pre.as_ref().map_or(0, |c| c.nrows() as u64),
            precomputed_cols: pre.as_ref().map_or(0, |c| c.ncols() as u64),
        };
        let mut rng = v.rng().clone(); // `let mut rng = self.rng().clone()` (KM/algorithm.rs:298)
        let tramp = lkm_rng { user: &mut rng as *mut R as *mut c_void, next_u32: tramp_u32::<R>, next_u64: tramp_u64::<R> };
        let mut centroids = Array2::<F>::zeros((k, d as usize));
        let mut counts = Array1::<F>::zeros(k);
        let mut inertia = F::zero();
        let mut stats = lkm_stats::default();
        let st = unsafe { F::fit(ctx.0, &p, tramp, centroids.as_mut_ptr(), counts.as_mut_ptr(), &mut inertia, &mut stats) };
        if st != LKM_OK {
            return Err(native_error(ctx.0, st));
        }
        Ok(KMeans { centroids, cluster_count: counts, inertia, device: self.1, dist_fn: PhantomData })
    }
}

impl<F: GpuFloat, DA: Data<Elem = F>> PredictInplace<ArrayBase<DA, Ix2>, Array1<usize>> for KMeans<F, L2Dist> {
    /// KM/algorithm.rs:743-756 -> update_cluster_memberships (:838-849)
    fn predict_inplace(&self, observations: &ArrayBase<DA, Ix2>, memberships: &mut Array1<usize>) {
        assert_eq!(observations.nrows(), memberships.len(), "The number of data points must match the number of memberships.");
        let x = observations.as_standard_layout();
        let ctx = Context::new(self.device).expect("CUDA device");
        let mut labels = vec![0u64; x.nrows()];
        let st = unsafe {
            F::predict(ctx.0, self.centroids.as_ptr(), self.centroids.nrows() as u64, x.as_ptr(), x.nrows() as u64,
                       x.ncols() as u64, x.ncols() as i64, labels.as_mut_ptr())
        };
        assert_eq!(st, LKM_OK, "lkm_predict failed");
        for (m, l) in memberships.iter_mut().zip(labels) {
            *m = l as usize;
        }
    }
    fn default_target(&self, x: &ArrayBase<DA, Ix2>) -> Array1<usize> {
        Array1::zeros(x.nrows())
    }
}

impl<F: GpuFloat, DA: Data<Elem = F>> PredictInplace<ArrayBase<DA, Ix1>, usize> for KMeans<F, L2Dist> {
    /// KM/algorithm.rs:763-777
    fn predict_inplace(&self, observation: &ArrayBase<DA, Ix1>, membership: &mut usize) {
        let x = observation.to_owned().insert_axis(ndarray::Axis(0));
        let mut out = Array1::zeros(1);
        PredictInplace::<Array2<F>, Array1<usize>>::predict_inplace(self, &x, &mut out);
        *membership = out[0];
    }
    fn default_target(&self, _x: &ArrayBase<DA, Ix1>) -> usize {
        0
    }
}

impl<F: GpuFloat, DA: Data<Elem = F>> Transformer<&ArrayBase<DA, Ix2>, Array1<F>> for KMeans<F, L2Dist> {
    /// KM/algorithm.rs:718-733 -> update_min_dists (:852-863): min SQUARED distance per row
    fn transform(&self, observations: &ArrayBase<DA, Ix2>) -> Array1<F> {
        let x = observations.as_standard_layout();
        let ctx = Context::new(self.device).expect("CUDA device");
        let mut out = Array1::<F>::zeros(x.nrows());
        let st = unsafe {
            F::transform(ctx.0, self.centroids.as_ptr(), self.centroids.nrows() as u64, x.as_ptr(), x.nrows() as u64,
                         x.ncols() as u64, x.ncols() as i64, out.as_mut_ptr())
        };
        assert_eq!(st, LKM_OK, "lkm_transform failed");
        out
    }
}

#[cfg(test)]
mod tests {
    use super::*;
    use ndarray::array;

    // mirrors KM/algorithm.rs:961-969
    #[test]
    fn autotraits() {
        fn has_autotraits<T: Send + Sync + Sized + Unpin>() {}
        has_autotraits::<KMeans<f64, L2Dist>>();
    }

    // mirrors KM/algorithm.rs:1003-1010 and :1150-1160 (needs a B200 at test time)
    #[test]
    fn kat_min_dists_and_argmin() {
        let m = KMeans::<f64, L2Dist> { centroids: array![[0.0, 1.0], [40.0, 10.0]], cluster_count: array![0., 0.], inertia: 0., device: 0, dist_fn: PhantomData };
        let d = m.transform(&array![[3.0, 4.0], [1.0, 3.0], [25.0, 15.0]]);
        assert_eq!(d, array![18.0, 5.0, 250.0]);
    }
}
