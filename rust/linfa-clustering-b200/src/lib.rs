//! B200-native K-Means behind linfa's `ParamGuard` / `Fit` / `FitWith` / `PredictInplace` / `Transformer` traits.
//!
//! Drop-in for `linfa_clustering::KMeans` with `D = L2Dist`:
//!
//! ```ignore
//! use linfa_clustering_b200::KMeans;            // instead of linfa_clustering::KMeans
//! let model = KMeans::params_with_rng(k, rng).max_n_iterations(200).tolerance(1e-5).fit(&dataset)?;
//! let model = KMeans::params(k).check().unwrap().fit(&dataset)?;   // the GMM call site (gaussian_mixture/algorithm.rs:137-146)
//! let dataset = model.predict(dataset);
//! ```
//!
//! * `GpuKMeansParams` is the unchecked builder; `impl ParamGuard` gives `check / check_ref / check_unwrap` and, through
//!   linfa's blanket impls (src/param_guard.rs:55-86), `Fit` and `FitWith` on the unchecked type as well.
//! * `GpuKMeansValidParams` is the checked set; `Fit` and `FitWith` are implemented on it.
//! * `KMeansInit`, `KMeansAlgorithm`, `KMeansError`, `KMeansParamsError`, `IncrKMeansError` are linfa-clustering's own
//!   types, re-exported — existing `match` arms keep compiling.
//! * All arithmetic runs in the sm_100a CUDA library behind `include/lkm.h`; there is no CPU fallback — `fit` returns
//!   `KMeansError::LinfaError` when no CUDA device is usable.
//! * `devices(&[0, 1, ..])` (the one knob linfa does not have) shards the rows over several GPUs from this one call.
//!
//! NOTE: written against the C ABI in this repository; it could not be compiled in the authoring environment (no
//! rustc / cargo there).  `tests/harness/lkm_harness.c` drives the same symbols in the same order from plain C, and the
//! Python binding `linfa_b200/_ffi.py` binds them too; both run in the test-suite.
#![allow(non_camel_case_types)]

use std::collections::HashMap;
use std::ffi::{c_char, c_void, CStr};
use std::marker::PhantomData;
use std::sync::{Mutex, OnceLock};

use linfa::dataset::DatasetBase;
use linfa::traits::{Fit, FitWith, PredictInplace, Transformer};
use linfa::{Float, ParamGuard};
pub use linfa_clustering::{IncrKMeansError, KMeansAlgorithm, KMeansError, KMeansInit, KMeansParamsError};
use linfa_nn::distance::L2Dist;
use ndarray::{Array1, Array2, ArrayBase, Data, Ix1, Ix2};
use ndarray_rand::rand::{Rng, RngCore, SeedableRng};
use rand_xoshiro::Xoshiro256Plus;

// ---------------------------------------------------------------------------------------------
// FFI (include/lkm.h)
// ---------------------------------------------------------------------------------------------
#[repr(C)]
pub struct lkm_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct lkm_multi {
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
    pub update_path: u32,
    pub rows_scanned: u64,
    pub rows_tightened: u64,
}

const LKM_OK: i32 = 0;
const LKM_ERR_INERTIA: i32 = 1;
const LKM_ERR_N_CLUSTERS: i32 = 10;
const LKM_ERR_N_RUNS: i32 = 11;
const LKM_ERR_TOLERANCE: i32 = 12;
const LKM_ERR_MAX_ITERATIONS: i32 = 13;
const LKM_ERR_INCREMENTAL_HAMERLY: i32 = 14;
const LKM_NOT_CONVERGED: i32 = 20;

extern "C" {
    fn lkm_create(out: *mut *mut lkm_ctx, device: i32) -> i32;
    fn lkm_destroy(ctx: *mut lkm_ctx) -> i32;
    fn lkm_last_error(ctx: *const lkm_ctx) -> *const c_char;
    fn lkm_create_multi(out: *mut *mut lkm_multi, device_ids: *const i32, n_devices: i32) -> i32;
    fn lkm_destroy_multi(m: *mut lkm_multi) -> i32;
    fn lkm_multi_last_error(m: *const lkm_multi) -> *const c_char;
    fn lkm_set_data_f32(ctx: *mut lkm_ctx, x: *const f32, n: u64, d: u64, row_stride: i64) -> i32;
    fn lkm_set_data_f64(ctx: *mut lkm_ctx, x: *const f64, n: u64, d: u64, row_stride: i64) -> i32;
    fn lkm_multi_set_data_f32(m: *mut lkm_multi, x: *const f32, n: u64, d: u64, row_stride: i64) -> i32;
    fn lkm_multi_set_data_f64(m: *mut lkm_multi, x: *const f64, n: u64, d: u64, row_stride: i64) -> i32;
    fn lkm_fit_f32(ctx: *mut lkm_ctx, p: *const lkm_params, rng: lkm_rng, c: *mut f32, cnt: *mut f32, inertia: *mut f32, st: *mut lkm_stats) -> i32;
    fn lkm_fit_f64(ctx: *mut lkm_ctx, p: *const lkm_params, rng: lkm_rng, c: *mut f64, cnt: *mut f64, inertia: *mut f64, st: *mut lkm_stats) -> i32;
    fn lkm_multi_fit_f32(m: *mut lkm_multi, p: *const lkm_params, rng: lkm_rng, c: *mut f32, cnt: *mut f32, inertia: *mut f32, st: *mut lkm_stats) -> i32;
    fn lkm_multi_fit_f64(m: *mut lkm_multi, p: *const lkm_params, rng: lkm_rng, c: *mut f64, cnt: *mut f64, inertia: *mut f64, st: *mut lkm_stats) -> i32;
    fn lkm_fit_with_f32(ctx: *mut lkm_ctx, p: *const lkm_params, rng: lkm_rng, has_model: i32, rows: u64, cols: u64, c: *mut f32, cnt: *mut f32, inertia: *mut f32) -> i32;
    fn lkm_fit_with_f64(ctx: *mut lkm_ctx, p: *const lkm_params, rng: lkm_rng, has_model: i32, rows: u64, cols: u64, c: *mut f64, cnt: *mut f64, inertia: *mut f64) -> i32;
    fn lkm_predict_f32(ctx: *mut lkm_ctx, c: *const f32, k: u64, x: *const f32, n: u64, d: u64, row_stride: i64, labels: *mut u64) -> i32;
    fn lkm_predict_f64(ctx: *mut lkm_ctx, c: *const f64, k: u64, x: *const f64, n: u64, d: u64, row_stride: i64, labels: *mut u64) -> i32;
    fn lkm_transform_f32(ctx: *mut lkm_ctx, c: *const f32, k: u64, x: *const f32, n: u64, d: u64, row_stride: i64, out: *mut f32) -> i32;
    fn lkm_transform_f64(ctx: *mut lkm_ctx, c: *const f64, k: u64, x: *const f64, n: u64, d: u64, row_stride: i64, out: *mut f64) -> i32;
}

/// Dispatch on `F` ∈ {f32, f64} — `linfa::Float` is implemented for exactly these two (src/dataset/mod.rs:68-74).
pub trait GpuFloat: Float {
    unsafe fn set_data(ctx: *mut lkm_ctx, x: *const Self, n: u64, d: u64, stride: i64) -> i32;
    unsafe fn multi_set_data(m: *mut lkm_multi, x: *const Self, n: u64, d: u64, stride: i64) -> i32;
    unsafe fn fit(ctx: *mut lkm_ctx, p: *const lkm_params, rng: lkm_rng, c: *mut Self, cnt: *mut Self, inertia: *mut Self, st: *mut lkm_stats) -> i32;
    unsafe fn multi_fit(m: *mut lkm_multi, p: *const lkm_params, rng: lkm_rng, c: *mut Self, cnt: *mut Self, inertia: *mut Self, st: *mut lkm_stats) -> i32;
    unsafe fn fit_with(ctx: *mut lkm_ctx, p: *const lkm_params, rng: lkm_rng, has_model: i32, rows: u64, cols: u64, c: *mut Self, cnt: *mut Self, inertia: *mut Self) -> i32;
    unsafe fn predict(ctx: *mut lkm_ctx, c: *const Self, k: u64, x: *const Self, n: u64, d: u64, stride: i64, labels: *mut u64) -> i32;
    unsafe fn transform(ctx: *mut lkm_ctx, c: *const Self, k: u64, x: *const Self, n: u64, d: u64, stride: i64, out: *mut Self) -> i32;
}
macro_rules! gpu_float {
    ($t:ty, $set:ident, $mset:ident, $fit:ident, $mfit:ident, $fw:ident, $pred:ident, $tr:ident) => {
        impl GpuFloat for $t {
            unsafe fn set_data(ctx: *mut lkm_ctx, x: *const Self, n: u64, d: u64, s: i64) -> i32 { $set(ctx, x, n, d, s) }
            unsafe fn multi_set_data(m: *mut lkm_multi, x: *const Self, n: u64, d: u64, s: i64) -> i32 { $mset(m, x, n, d, s) }
            unsafe fn fit(ctx: *mut lkm_ctx, p: *const lkm_params, rng: lkm_rng, c: *mut Self, cnt: *mut Self, i: *mut Self, st: *mut lkm_stats) -> i32 { $fit(ctx, p, rng, c, cnt, i, st) }
            unsafe fn multi_fit(m: *mut lkm_multi, p: *const lkm_params, rng: lkm_rng, c: *mut Self, cnt: *mut Self, i: *mut Self, st: *mut lkm_stats) -> i32 { $mfit(m, p, rng, c, cnt, i, st) }
            unsafe fn fit_with(ctx: *mut lkm_ctx, p: *const lkm_params, rng: lkm_rng, h: i32, r: u64, co: u64, c: *mut Self, cnt: *mut Self, i: *mut Self) -> i32 { $fw(ctx, p, rng, h, r, co, c, cnt, i) }
            unsafe fn predict(ctx: *mut lkm_ctx, c: *const Self, k: u64, x: *const Self, n: u64, d: u64, s: i64, l: *mut u64) -> i32 { $pred(ctx, c, k, x, n, d, s, l) }
            unsafe fn transform(ctx: *mut lkm_ctx, c: *const Self, k: u64, x: *const Self, n: u64, d: u64, s: i64, o: *mut Self) -> i32 { $tr(ctx, c, k, x, n, d, s, o) }
        }
    };
}
gpu_float!(f32, lkm_set_data_f32, lkm_multi_set_data_f32, lkm_fit_f32, lkm_multi_fit_f32, lkm_fit_with_f32, lkm_predict_f32, lkm_transform_f32);
gpu_float!(f64, lkm_set_data_f64, lkm_multi_set_data_f64, lkm_fit_f64, lkm_multi_fit_f64, lkm_fit_with_f64, lkm_predict_f64, lkm_transform_f64);

// ---------------------------------------------------------------------------------------------
// contexts: one per device, created on first use and kept (a context owns a stream, pinned
// buffers and device workspaces; creating one per predict / transform call would dominate
// small calls).  A context is used by one thread at a time: the Mutex is held for the call.
// ---------------------------------------------------------------------------------------------
struct Context(*mut lkm_ctx);
unsafe impl Send for Context {}
impl Drop for Context {
    fn drop(&mut self) {
        unsafe { lkm_destroy(self.0) };
    }
}

fn with_context<T>(device: i32, f: impl FnOnce(*mut lkm_ctx) -> Result<T, KMeansError>) -> Result<T, KMeansError> {
    static CONTEXTS: OnceLock<Mutex<HashMap<i32, Context>>> = OnceLock::new();
    let mut map = CONTEXTS.get_or_init(|| Mutex::new(HashMap::new())).lock().unwrap_or_else(|e| e.into_inner());
    if !map.contains_key(&device) {
        let mut h = std::ptr::null_mut();
        let st = unsafe { lkm_create(&mut h, device) };
        if st != LKM_OK {
            return Err(native_error(std::ptr::null(), st));
        }
        map.insert(device, Context(h));
    }
    f(map[&device].0)
}

struct Group(*mut lkm_multi);
impl Drop for Group {
    fn drop(&mut self) {
        unsafe { lkm_destroy_multi(self.0) };
    }
}

fn native_message(ctx: *const lkm_ctx, status: i32) -> String {
    let msg = if ctx.is_null() {
        "no usable CUDA device (this path has no CPU fallback)".to_string()
    } else {
        unsafe { CStr::from_ptr(lkm_last_error(ctx)) }.to_string_lossy().into_owned()
    };
    format!("lkm status {status}: {msg}")
}

fn param_error(status: i32) -> Option<KMeansParamsError> {
    match status {
        LKM_ERR_N_CLUSTERS => Some(KMeansParamsError::NClusters),
        LKM_ERR_N_RUNS => Some(KMeansParamsError::NRuns),
        LKM_ERR_TOLERANCE => Some(KMeansParamsError::Tolerance),
        LKM_ERR_MAX_ITERATIONS => Some(KMeansParamsError::MaxIterations),
        LKM_ERR_INCREMENTAL_HAMERLY => Some(KMeansParamsError::IncrementalHamerly),
        _ => None,
    }
}

fn native_error(ctx: *const lkm_ctx, status: i32) -> KMeansError {
    match (status, param_error(status)) {
        (_, Some(e)) => e.into(),
        (LKM_ERR_INERTIA, _) => KMeansError::InertiaError,
        _ => KMeansError::LinfaError(linfa::Error::Parameters(native_message(ctx, status))),
    }
}

// trampolines into the caller's `R: Rng` (KM/hyperparams.rs:37)
unsafe extern "C" fn tramp_u32<R: RngCore>(user: *mut c_void) -> u32 {
    (*(user as *mut R)).next_u32()
}
unsafe extern "C" fn tramp_u64<R: RngCore>(user: *mut c_void) -> u64 {
    (*(user as *mut R)).next_u64()
}

/// (pointer, row stride in elements) of a view whose rows are contiguous; any other ndarray layout is copied to the
/// standard layout first (`owned` keeps the copy alive).
fn rows_of<'a, F: Float, S: Data<Elem = F>>(x: &'a ArrayBase<S, Ix2>, owned: &'a mut Option<Array2<F>>) -> (*const F, i64) {
    let s = x.strides();
    let contiguous_rows = x.ncols() <= 1 || s[1] == 1;
    if contiguous_rows && (x.nrows() <= 1 || s[0] >= x.ncols() as isize) {
        (x.as_ptr(), if x.nrows() <= 1 { x.ncols() as i64 } else { s[0] as i64 })
    } else {
        *owned = Some(x.as_standard_layout().into_owned());
        let o = owned.as_ref().unwrap();
        (o.as_ptr(), o.ncols() as i64)
    }
}

// ---------------------------------------------------------------------------------------------
// hyper-parameters (KM/hyperparams.rs)
// ---------------------------------------------------------------------------------------------
/// The checked set (KM/hyperparams.rs:19-42) plus this crate's device list.
#[derive(Clone, Debug, PartialEq)]
pub struct GpuKMeansValidParams<F: Float, R: Rng> {
    n_runs: usize,
    tolerance: F,
    max_n_iterations: u64,
    n_clusters: usize,
    init: KMeansInit<F>,
    rng: R,
    algorithm: KMeansAlgorithm,
    devices: Vec<i32>,
}

/// The unchecked builder (KM/hyperparams.rs:50).  Defaults (:71-82): n_runs 10, tolerance 1e-4, max_n_iterations 300,
/// KMeansPlusPlus, Lloyd; device 0.
#[derive(Clone, Debug, PartialEq)]
pub struct GpuKMeansParams<F: Float, R: Rng>(GpuKMeansValidParams<F, R>);

impl<F: Float, R: Rng> GpuKMeansParams<F, R> {
    pub fn new(n_clusters: usize, rng: R) -> Self {
        Self(GpuKMeansValidParams {
            n_runs: 10,
            tolerance: F::cast(1e-4),
            max_n_iterations: 300,
            n_clusters,
            init: KMeansInit::KMeansPlusPlus,
            rng,
            algorithm: KMeansAlgorithm::Lloyd,
            devices: vec![0],
        })
    }
    pub fn n_runs(mut self, n_runs: usize) -> Self {
        self.0.n_runs = n_runs;
        self
    }
    pub fn tolerance(mut self, tolerance: F) -> Self {
        self.0.tolerance = tolerance;
        self
    }
    pub fn max_n_iterations(mut self, max_n_iterations: u64) -> Self {
        self.0.max_n_iterations = max_n_iterations;
        self
    }
    pub fn init_method(mut self, init: KMeansInit<F>) -> Self {
        self.0.init = init;
        self
    }
    pub fn algorithm(mut self, algorithm: KMeansAlgorithm) -> Self {
        self.0.algorithm = algorithm;
        self
    }
    /// Not in linfa: the CUDA devices to use.  More than one device shards the rows (contiguous blocks) and exchanges
    /// the per-device partial sums over NVLink every iteration.
    pub fn devices(mut self, devices: &[i32]) -> Self {
        self.0.devices = devices.to_vec();
        self
    }
}

impl<F: Float, R: Rng> ParamGuard for GpuKMeansParams<F, R> {
    type Checked = GpuKMeansValidParams<F, R>;
    type Error = KMeansParamsError;

    /// KM/hyperparams.rs:124-136, same order
    fn check_ref(&self) -> Result<&Self::Checked, Self::Error> {
        if self.0.n_clusters == 0 {
            Err(KMeansParamsError::NClusters)
        } else if self.0.n_runs == 0 {
            Err(KMeansParamsError::NRuns)
        } else if self.0.tolerance <= F::zero() {
            Err(KMeansParamsError::Tolerance)
        } else if self.0.max_n_iterations == 0 {
            Err(KMeansParamsError::MaxIterations)
        } else {
            Ok(&self.0)
        }
    }
    fn check(self) -> Result<Self::Checked, Self::Error> {
        self.check_ref()?;
        Ok(self.0)
    }
}

impl<F: Float, R: Rng> GpuKMeansValidParams<F, R> {
    pub fn n_runs(&self) -> usize {
        self.n_runs
    }
    pub fn tolerance(&self) -> F {
        self.tolerance
    }
    pub fn max_n_iterations(&self) -> u64 {
        self.max_n_iterations
    }
    pub fn n_clusters(&self) -> usize {
        self.n_clusters
    }
    pub fn init_method(&self) -> &KMeansInit<F> {
        &self.init
    }
    pub fn rng(&self) -> &R {
        &self.rng
    }
    pub fn algorithm(&self) -> &KMeansAlgorithm {
        &self.algorithm
    }
    pub fn dist_fn(&self) -> &L2Dist {
        &L2Dist
    }

    /// `precomputed` keeps the standard-layout copy of `KMeansInit::Precomputed` alive while the C struct points at it.
    fn native(&self, precomputed: &mut Option<Array2<F>>) -> lkm_params {
        let init_method = match &self.init {
            KMeansInit::Random => 0,
            KMeansInit::Precomputed(c) => {
                *precomputed = Some(c.as_standard_layout().into_owned());
                1
            }
            KMeansInit::KMeansPlusPlus => 2,
            KMeansInit::KMeansPara => 3,
            #[allow(unreachable_patterns)]
            _ => 2, // KMeansInit is #[non_exhaustive]
        };
        lkm_params {
            n_clusters: self.n_clusters as u64,
            n_runs: self.n_runs as u64,
            max_n_iterations: self.max_n_iterations,
            tolerance: self.tolerance.to_f64().unwrap(),
            init_method,
            algorithm: if self.algorithm == KMeansAlgorithm::Hamerly { 1 } else { 0 },
            pick_mode: 0, // rand 0.8.5 WeightedIndex bit for bit: same rows as linfa for the same rng
            reserved: 0,
            precomputed: precomputed.as_ref().map_or(std::ptr::null(), |c| c.as_ptr() as *const c_void),
            precomputed_rows: precomputed.as_ref().map_or(0, |c| c.nrows() as u64),
            precomputed_cols: precomputed.as_ref().map_or(0, |c| c.ncols() as u64),
        }
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
    pub fn params(nclusters: usize) -> GpuKMeansParams<F, Xoshiro256Plus> {
        GpuKMeansParams::new(nclusters, Xoshiro256Plus::seed_from_u64(42))
    }
    /// `KMeans::params_with_rng(k, rng)` (KM/algorithm.rs:214-216)
    pub fn params_with_rng<R: Rng>(nclusters: usize, rng: R) -> GpuKMeansParams<F, R> {
        GpuKMeansParams::new(nclusters, rng)
    }
    /// `KMeans::params_with(k, rng, dist_fn)` (KM/algorithm.rs:218-222); only `L2Dist` runs on the GPU
    pub fn params_with<R: Rng>(nclusters: usize, rng: R, _dist_fn: L2Dist) -> GpuKMeansParams<F, R> {
        GpuKMeansParams::new(nclusters, rng)
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

impl<F: GpuFloat, R: Rng + Clone, DA: Data<Elem = F>, T> Fit<ArrayBase<DA, Ix2>, T, KMeansError> for GpuKMeansValidParams<F, R> {
    type Object = KMeans<F, L2Dist>;

    /// `Fit::fit` (KM/algorithm.rs:370-388): fit_lloyd / fit_hamerly — seeding, the iteration loop, best of n_runs,
    /// cluster_count, mean inertia — run behind `lkm_fit_*` (or `lkm_multi_fit_*` over a device group).
    fn fit(&self, dataset: &DatasetBase<ArrayBase<DA, Ix2>, T>) -> Result<Self::Object, KMeansError> {
        let obs = dataset.records();
        let mut owned = None;
        let (ptr, stride) = rows_of(obs, &mut owned);
        let (n, d, k) = (obs.nrows() as u64, obs.ncols() as u64, self.n_clusters);
        let mut pre = None;
        let p = self.native(&mut pre);
        let mut rng = self.rng.clone(); // `let mut rng = self.rng().clone()` (KM/algorithm.rs:298)
        let tramp = lkm_rng { user: &mut rng as *mut R as *mut c_void, next_u32: tramp_u32::<R>, next_u64: tramp_u64::<R> };
        let mut centroids = Array2::<F>::zeros((k, d as usize));
        let mut counts = Array1::<F>::zeros(k);
        let mut inertia = F::zero();
        let mut stats = lkm_stats::default();
        let device = *self.devices.first().unwrap_or(&0);
        if self.devices.len() > 1 {
            let mut h = std::ptr::null_mut();
            let st = unsafe { lkm_create_multi(&mut h, self.devices.as_ptr(), self.devices.len() as i32) };
            if st != LKM_OK {
                return Err(native_error(std::ptr::null(), st));
            }
            let group = Group(h);
            let multi_err = |st: i32| match (st, param_error(st)) {
                (_, Some(e)) => KMeansError::from(e),
                (LKM_ERR_INERTIA, _) => KMeansError::InertiaError,
                _ => KMeansError::LinfaError(linfa::Error::Parameters(format!(
                    "lkm status {st}: {}",
                    unsafe { CStr::from_ptr(lkm_multi_last_error(group.0)) }.to_string_lossy()
                ))),
            };
            let st = unsafe { F::multi_set_data(group.0, ptr, n, d, stride) };
            if st != LKM_OK {
                return Err(multi_err(st));
            }
            let st = unsafe { F::multi_fit(group.0, &p, tramp, centroids.as_mut_ptr(), counts.as_mut_ptr(), &mut inertia, &mut stats) };
            if st != LKM_OK {
                return Err(multi_err(st));
            }
        } else {
            with_context(device, |ctx| {
                let st = unsafe { F::set_data(ctx, ptr, n, d, stride) };
                if st != LKM_OK {
                    return Err(native_error(ctx, st));
                }
                let st = unsafe { F::fit(ctx, &p, tramp, centroids.as_mut_ptr(), counts.as_mut_ptr(), &mut inertia, &mut stats) };
                if st != LKM_OK {
                    return Err(native_error(ctx, st));
                }
                Ok(())
            })?;
        }
        Ok(KMeans { centroids, cluster_count: counts, inertia, device, dist_fn: PhantomData })
    }
}

impl<'a, F: GpuFloat + std::fmt::Debug, R: Rng + Clone, DA: Data<Elem = F>, T> FitWith<'a, ArrayBase<DA, Ix2>, T, IncrKMeansError<KMeans<F, L2Dist>>>
    for GpuKMeansValidParams<F, R>
{
    type ObjectIn = Option<KMeans<F, L2Dist>>;
    type ObjectOut = KMeans<F, L2Dist>;

    /// `FitWith::fit_with` (KM/algorithm.rs:635-715): one mini-batch step behind `lkm_fit_with_*`.  With a model the
    /// model's own shape rules (the reference never looks at n_clusters then); Hamerly is rejected (:640-644).
    fn fit_with(&self, model: Self::ObjectIn, dataset: &'a DatasetBase<ArrayBase<DA, Ix2>, T>) -> Result<Self::ObjectOut, IncrKMeansError<Self::ObjectOut>> {
        if self.algorithm == KMeansAlgorithm::Hamerly {
            return Err(IncrKMeansError::InvalidParams(KMeansParamsError::IncrementalHamerly));
        }
        let obs = dataset.records();
        let mut owned = None;
        let (ptr, stride) = rows_of(obs, &mut owned);
        let (n, d) = (obs.nrows() as u64, obs.ncols() as u64);
        let mut pre = None;
        let p = self.native(&mut pre);
        let mut rng = self.rng.clone();
        let tramp = lkm_rng { user: &mut rng as *mut R as *mut c_void, next_u32: tramp_u32::<R>, next_u64: tramp_u64::<R> };
        let has_model = model.is_some();
        let device = model.as_ref().map_or(*self.devices.first().unwrap_or(&0), |m| m.device);
        let (mut centroids, mut counts) = match model {
            Some(m) => (m.centroids.as_standard_layout().into_owned(), m.cluster_count.to_owned()),
            None => (Array2::<F>::zeros((self.n_clusters, d as usize)), Array1::<F>::zeros(self.n_clusters)),
        };
        let mut inertia = F::zero();
        let lin = |msg: String| IncrKMeansError::LinfaError(linfa::Error::Parameters(msg));
        let st = with_context(device, |ctx| {
            let st = unsafe { F::set_data(ctx, ptr, n, d, stride) };
            if st != LKM_OK {
                return Err(native_error(ctx, st));
            }
            let st = unsafe {
                F::fit_with(ctx, &p, tramp, has_model as i32, centroids.nrows() as u64, centroids.ncols() as u64, centroids.as_mut_ptr(),
                            counts.as_mut_ptr(), &mut inertia)
            };
            if st != LKM_OK && st != LKM_NOT_CONVERGED {
                return Err(native_error(ctx, st));
            }
            Ok(st)
        })
        .map_err(|e| match e {
            KMeansError::InvalidParams(p) => IncrKMeansError::InvalidParams(p),
            other => lin(other.to_string()),
        })?;
        let out = KMeans { centroids, cluster_count: counts, inertia, device, dist_fn: PhantomData };
        if st == LKM_OK {
            Ok(out)
        } else {
            Err(IncrKMeansError::NotConverged(out))
        }
    }
}

impl<F: GpuFloat, DA: Data<Elem = F>> PredictInplace<ArrayBase<DA, Ix2>, Array1<usize>> for KMeans<F, L2Dist> {
    /// KM/algorithm.rs:743-756 -> update_cluster_memberships (:838-849)
    fn predict_inplace(&self, observations: &ArrayBase<DA, Ix2>, memberships: &mut Array1<usize>) {
        assert_eq!(observations.nrows(), memberships.len(), "The number of data points must match the number of memberships.");
        let mut owned = None;
        let (ptr, stride) = rows_of(observations, &mut owned);
        let mut labels = vec![0u64; observations.nrows()];
        with_context(self.device, |ctx| {
            let st = unsafe {
                F::predict(ctx, self.centroids.as_ptr(), self.centroids.nrows() as u64, ptr, observations.nrows() as u64, observations.ncols() as u64,
                           stride, labels.as_mut_ptr())
            };
            if st != LKM_OK {
                return Err(native_error(ctx, st));
            }
            Ok(())
        })
        .expect("lkm_predict failed"); // the reference panics on shape mismatches here too (:744-748)
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
        let mut owned = None;
        let (ptr, stride) = rows_of(observations, &mut owned);
        let mut out = Array1::<F>::zeros(observations.nrows());
        with_context(self.device, |ctx| {
            let st = unsafe {
                F::transform(ctx, self.centroids.as_ptr(), self.centroids.nrows() as u64, ptr, observations.nrows() as u64, observations.ncols() as u64,
                             stride, out.as_mut_ptr())
            };
            if st != LKM_OK {
                return Err(native_error(ctx, st));
            }
            Ok(())
        })
        .expect("lkm_transform failed");
        out
    }
}

#[cfg(test)]
mod tests {
    use super::*;
    use linfa::traits::Predict;
    use ndarray::array;

    // mirrors KM/algorithm.rs:961-969
    #[test]
    fn autotraits() {
        fn has_autotraits<T: Send + Sync + Sized + Unpin>() {}
        has_autotraits::<KMeans<f64, L2Dist>>();
        has_autotraits::<GpuKMeansParams<f64, Xoshiro256Plus>>();
        has_autotraits::<GpuKMeansValidParams<f64, Xoshiro256Plus>>();
    }

    // hyperparams.rs:204-232
    #[test]
    fn param_guard() {
        assert!(matches!(KMeans::<f64>::params(0).check(), Err(KMeansParamsError::NClusters)));
        assert!(matches!(KMeans::<f64>::params(1).n_runs(0).check(), Err(KMeansParamsError::NRuns)));
        assert!(matches!(KMeans::<f64>::params(1).tolerance(0.0).check(), Err(KMeansParamsError::Tolerance)));
        assert!(matches!(KMeans::<f64>::params(1).max_n_iterations(0).check(), Err(KMeansParamsError::MaxIterations)));
    }

    // the GMM call site (gaussian_mixture/algorithm.rs:137-146): `.check().unwrap().fit(dataset)?` then `predict`
    #[test]
    fn gmm_call_site_compiles_and_runs() {
        let data = array![[0.0, 0.1], [0.1, 0.0], [10.0, 10.1], [10.1, 10.0]];
        let dataset = DatasetBase::from(data.clone());
        let model = KMeans::params_with_rng(2, Xoshiro256Plus::seed_from_u64(42)).check().unwrap().fit(&dataset).unwrap();
        let labels = model.predict(&data);
        assert_eq!(labels[0], labels[1]);
        assert_ne!(labels[0], labels[2]);
    }

    // mirrors KM/algorithm.rs:1003-1013 (needs a B200 at test time)
    #[test]
    fn kat_min_dists() {
        let m = KMeans::<f64, L2Dist> { centroids: array![[0.0, 1.0], [40.0, 10.0]], cluster_count: array![0., 0.], inertia: 0., device: 0, dist_fn: PhantomData };
        let d = m.transform(&array![[3.0, 4.0], [1.0, 3.0], [25.0, 15.0]]);
        assert_eq!(d, array![18.0, 5.0, 250.0]);
    }

    // mirrors fit_with_rejects_hamerly (KM/algorithm.rs:1204-1218)
    #[test]
    fn fit_with_rejects_hamerly() {
        let dataset = DatasetBase::from(array![[0.0, 0.0], [1.0, 1.0]]);
        let err = KMeans::params(1).algorithm(KMeansAlgorithm::Hamerly).check().unwrap().fit_with(None, &dataset).unwrap_err();
        assert!(matches!(err, IncrKMeansError::InvalidParams(KMeansParamsError::IncrementalHamerly)));
    }
}
