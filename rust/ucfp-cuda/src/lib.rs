//! `ucfp-cuda` -- the thin FFI crate the UCFP host links to run its data-parallel hot path on a B200.
//!
//! It binds `include/ucfp_cuda.h` one to one (`mod sys`) and offers the two seams of the reference:
//!   * `image::fingerprint_batch`  replaces the calls into `imgfprint` at `src/modality/image.rs:68-70,175-179`
//!     after host-side decode, and builds the same 168 / 536-byte `Record::fingerprint` blobs;
//!   * `GpuIndexBackend: IndexBackend` replaces `EmbeddedBackend::knn` (`src/index/embedded/mod.rs:268-360`)
//!     and adds `HashIndex::{hamming_knn, jaccard_knn}`, which the reference lacks.
//!
//! This crate is NOT compiled in the repository's development container (no Rust toolchain there). The Python
//! package `ucfp_b200` binds the same ABI through ctypes and is what the tests and the benchmark run.
#![allow(non_camel_case_types)]

pub mod sys {
    use std::os::raw::{c_char, c_int, c_void};
    #[repr(C)] pub struct ucfp_ctx { _p: [u8; 0] }
    #[repr(C)] pub struct ucfp_corpus { _p: [u8; 0] }
    #[repr(C)] #[derive(Copy, Clone)]
    pub struct ucfp_image_desc { pub pixels: *const u8, pub width: u32, pub height: u32, pub stride: u64 }
    #[repr(C)] #[derive(Copy, Clone, bytemuck::Zeroable, bytemuck::Pod)]
    pub struct ucfp_hash17 { pub global_hash: u64, pub block_hashes: [u64; 16] }
    #[repr(C)] #[derive(Copy, Clone, bytemuck::Zeroable, bytemuck::Pod)]
    pub struct ucfp_image_hashes { pub ahash: ucfp_hash17, pub phash: ucfp_hash17, pub dhash: ucfp_hash17 }
    pub const UCFP_OK: c_int = 0;
    pub const UCFP_E_UNSUPPORTED: c_int = -4;
    pub const UCFP_ALGO_AHASH: u32 = 1; pub const UCFP_ALGO_PHASH: u32 = 2; pub const UCFP_ALGO_DHASH: u32 = 4; pub const UCFP_ALGO_MULTI: u32 = 7;
    pub const UCFP_KIND_HAMMING64: c_int = 1; pub const UCFP_KIND_MINHASH128: c_int = 2; pub const UCFP_KIND_COSINE: c_int = 3;
    pub const UCFP_ID_NONE: u64 = u64::MAX;
    unsafe extern "C" {
        pub fn ucfp_init(device: c_int, out: *mut *mut ucfp_ctx) -> c_int;
        pub fn ucfp_destroy(ctx: *mut ucfp_ctx);
        pub fn ucfp_last_error() -> *const c_char;
        pub fn ucfp_ctx_synchronize(ctx: *mut ucfp_ctx) -> c_int;
        pub fn ucfp_image_hash_batch(ctx: *mut ucfp_ctx, imgs: *const ucfp_image_desc, n: usize, algo_mask: u32,
                                     out: *mut ucfp_image_hashes, status: *mut i32) -> c_int;
        pub fn ucfp_corpus_create(ctx: *mut ucfp_ctx, kind: c_int, dim: u32, capacity: u64, out: *mut *mut ucfp_corpus) -> c_int;
        pub fn ucfp_corpus_destroy(c: *mut ucfp_corpus);
        pub fn ucfp_corpus_append(c: *mut ucfp_corpus, ids: *const u64, rows: *const c_void, n: u64) -> c_int;
        pub fn ucfp_corpus_clear(c: *mut ucfp_corpus) -> c_int;
        pub fn ucfp_corpus_size(c: *const ucfp_corpus) -> u64;
        pub fn ucfp_scan_hamming(c: *mut ucfp_corpus, q: *const u64, nq: usize, k: usize, ids: *mut u64, dist: *mut u32) -> c_int;
        pub fn ucfp_scan_jaccard(c: *mut ucfp_corpus, q: *const u64, nq: usize, k: usize, ids: *mut u64, matches: *mut u32) -> c_int;
        pub fn ucfp_scan_cosine(c: *mut ucfp_corpus, q: *const f32, nq: usize, k: usize, ids: *mut u64, score: *mut f32) -> c_int;
        pub fn ucfp_merge_topk_u32(ctx: *mut ucfp_ctx, ids_in: *const u64, keys_in: *const u32, parts: usize, nq: usize, k: usize,
                                   descending: c_int, ids_out: *mut u64, keys_out: *mut u32) -> c_int;
        pub fn ucfp_merge_topk_f32(ctx: *mut ucfp_ctx, ids_in: *const u64, scores_in: *const f32, parts: usize, nq: usize, k: usize,
                                   ids_out: *mut u64, scores_out: *mut f32) -> c_int;
    }
}

use std::{collections::HashMap, ffi::CStr, sync::Mutex};
use bytes::Bytes;
use ucfp::core::{Hit, HitSource, Modality, Record};
use ucfp::error::{Error, Result};

fn last_error() -> String { unsafe { CStr::from_ptr(sys::ucfp_last_error()) }.to_string_lossy().into_owned() }
fn check_index(rc: i32) -> Result<()> { if rc == 0 { Ok(()) } else { Err(Error::Index(last_error())) } }

/// One per (process, GPU).  `Send + Sync`: every ABI entry point takes the context's own lock.
pub struct Gpu { ctx: *mut sys::ucfp_ctx }
unsafe impl Send for Gpu {} unsafe impl Sync for Gpu {}
impl Gpu {
    pub fn new(device: i32) -> Result<Self> {
        let mut ctx = std::ptr::null_mut();
        check_index(unsafe { sys::ucfp_init(device, &mut ctx) })?;
        Ok(Self { ctx })
    }
}
impl Drop for Gpu { fn drop(&mut self) { unsafe { sys::ucfp_destroy(self.ctx) } } }

/// Hashing seam.  `DecodedRgb` is what the host's decoder (the `image` crate, unchanged) produces.
pub mod image {
    use super::*;
    pub struct DecodedRgb<'a> { pub pixels: &'a [u8], pub width: u32, pub height: u32, pub encoded: &'a [u8] }
    pub use ucfp::image::{ALGORITHM_AHASH, ALGORITHM_DHASH, ALGORITHM_MULTIHASH, ALGORITHM_PHASH};

    /// Batched `ucfp::image::fingerprint*`: one GPU call for the whole batch, one `Result<Record>` per image
    /// (errors map to `Error::Modality`, as at `src/modality/image.rs:70`).
    pub fn fingerprint_batch(gpu: &Gpu, imgs: &[DecodedRgb<'_>], algo_mask: u32, tenant_id: u32, record_ids: &[u64]) -> Vec<Result<Record>> {
        let descs: Vec<_> = imgs.iter().map(|i| sys::ucfp_image_desc { pixels: i.pixels.as_ptr(), width: i.width, height: i.height, stride: 3 * i.width as u64 }).collect();
        let mut out = vec![bytemuck::Zeroable::zeroed(); imgs.len()];
        let mut status = vec![0i32; imgs.len()];
        let rc = unsafe { sys::ucfp_image_hash_batch(gpu.ctx, descs.as_ptr(), descs.len(), algo_mask, out.as_mut_ptr(), status.as_mut_ptr()) };
        imgs.iter().enumerate().map(|(i, img)| {
            if rc != 0 || status[i] != 0 { return Err(Error::Modality(last_error())); }
            let exact = *blake3::hash(img.encoded).as_bytes();           // host side, as in imgfprint
            let h: &sys::ucfp_image_hashes = &out[i];
            let single = |h17: &sys::ucfp_hash17| { let mut b = Vec::with_capacity(168); b.extend_from_slice(&exact); b.extend_from_slice(bytemuck::bytes_of(h17)); b };
            let (tag, blob) = match algo_mask {
                sys::UCFP_ALGO_MULTI => { let mut b = Vec::with_capacity(536); b.extend_from_slice(&exact); for h17 in [&h.ahash, &h.phash, &h.dhash] { b.extend_from_slice(&single(h17)); } (ALGORITHM_MULTIHASH, b) }
                sys::UCFP_ALGO_PHASH => (ALGORITHM_PHASH, single(&h.phash)),
                sys::UCFP_ALGO_DHASH => (ALGORITHM_DHASH, single(&h.dhash)),
                _ => (ALGORITHM_AHASH, single(&h.ahash)),
            };
            Ok(Record { tenant_id, record_id: record_ids[i], modality: Modality::Image, format_version: 1, algorithm: tag.into(),
                        config_hash: 0, fingerprint: Bytes::from(blob), embedding: None, model_id: None, metadata: Bytes::new(), text: None })
        }).collect()
    }
}

/// Scan seam: cosine k-NN behind `IndexBackend::knn`, one HBM corpus per (tenant, dim); Hamming corpora per
/// (tenant, algorithm tag) and MinHash corpora per tenant behind `HashIndex`.
pub struct GpuIndexBackend {
    gpu: Gpu,
    vectors: Mutex<HashMap<(u32, usize), *mut sys::ucfp_corpus>>,
    hashes: Mutex<HashMap<(u32, String), *mut sys::ucfp_corpus>>,
    signatures: Mutex<HashMap<u32, *mut sys::ucfp_corpus>>,
}

/// The queries the reference stores fingerprints for but cannot ask (`src/index/mod.rs:29-35` has `knn` only).
/// `Hit::score` = `1 - dist / 64` (Hamming) or `matches / 128` (Jaccard), `source = HitSource::Vector`.
pub trait HashIndex {
    fn hamming_knn_blocking(&self, tenant_id: u32, algorithm: &str, code: u64, k: usize) -> Result<Vec<Hit>>;
    fn jaccard_knn_blocking(&self, tenant_id: u32, signature: &[u64; 128], k: usize) -> Result<Vec<Hit>>;
}

fn hit(tenant_id: u32, record_id: u64, score: f32) -> Hit {
    Hit { tenant_id, record_id, score, source: HitSource::Vector, vector_score: None, bm25_score: None, vector_rank: None,
          bm25_rank: None, term_hits: Vec::new() }
}

impl HashIndex for GpuIndexBackend {
    fn hamming_knn_blocking(&self, tenant_id: u32, algorithm: &str, code: u64, k: usize) -> Result<Vec<Hit>> {
        if k == 0 { return Ok(Vec::new()); }
        let Some(&corpus) = self.hashes.lock().unwrap().get(&(tenant_id, algorithm.to_string())) else { return Ok(Vec::new()) };
        let (mut ids, mut dist) = (vec![0u64; k], vec![0u32; k]);
        check_index(unsafe { sys::ucfp_scan_hamming(corpus, &code, 1, k, ids.as_mut_ptr(), dist.as_mut_ptr()) })?;
        Ok(ids.into_iter().zip(dist).filter(|(id, _)| *id != sys::UCFP_ID_NONE).map(|(id, d)| hit(tenant_id, id, 1.0 - d as f32 / 64.0)).collect())
    }
    fn jaccard_knn_blocking(&self, tenant_id: u32, signature: &[u64; 128], k: usize) -> Result<Vec<Hit>> {
        if k == 0 { return Ok(Vec::new()); }
        let Some(&corpus) = self.signatures.lock().unwrap().get(&tenant_id) else { return Ok(Vec::new()) };
        let (mut ids, mut m) = (vec![0u64; k], vec![0u32; k]);
        check_index(unsafe { sys::ucfp_scan_jaccard(corpus, signature.as_ptr(), 1, k, ids.as_mut_ptr(), m.as_mut_ptr()) })?;
        Ok(ids.into_iter().zip(m).filter(|(id, _)| *id != sys::UCFP_ID_NONE).map(|(id, x)| hit(tenant_id, id, x as f32 / 128.0)).collect())
    }
}
unsafe impl Send for GpuIndexBackend {} unsafe impl Sync for GpuIndexBackend {}

impl GpuIndexBackend {
    pub fn knn_blocking(&self, tenant_id: u32, query: &[f32], k: usize) -> Result<Vec<Hit>> {
        if query.is_empty() || k == 0 { return Ok(Vec::new()); }                    // embedded/mod.rs:275
        let Some(&corpus) = self.vectors.lock().unwrap().get(&(tenant_id, query.len())) else { return Ok(Vec::new()) };
        let (mut ids, mut scores) = (vec![0u64; k], vec![0f32; k]);
        check_index(unsafe { sys::ucfp_scan_cosine(corpus, query.as_ptr(), 1, k, ids.as_mut_ptr(), scores.as_mut_ptr()) })?;
        Ok(ids.into_iter().zip(scores).filter(|(id, _)| *id != sys::UCFP_ID_NONE).map(|(record_id, score)| Hit {
            tenant_id, record_id, score, source: HitSource::Vector, vector_score: None, bm25_score: None, vector_rank: None,
            bm25_rank: None, term_hits: Vec::new() }).collect())
    }
}
// `impl ucfp::IndexBackend for GpuIndexBackend` forwards `knn` to `knn_blocking` inside `spawn_blocking`
// (as EmbeddedBackend does, embedded/mod.rs:282) and delegates upsert/delete/bm25/flush/get_record_metadata to
// the wrapped `EmbeddedBackend`, mirroring vectors into HBM on upsert: see INTEGRATION.md.
