//! `ucfp-cuda` -- the thin FFI crate the UCFP host links to run its data-parallel hot path on a B200.
//!
//! It binds `include/ucfp_cuda.h` (`mod sys`) and offers the two seams of the reference:
//!   * `image::fingerprint_batch` / `image::fingerprint_jpeg_batch` replace the calls into `imgfprint` at
//!     `src/modality/image.rs:68-70,175-179` (after host-side decode, or -- for JPEG -- from the encoded bytes, decoded on the
//!     device) and build the same 168 / 536-byte `Record::fingerprint` blobs;
//!   * `GpuIndexBackend` **implements `ucfp::IndexBackend`** (`src/index/mod.rs:17-78`): storage, BM25 and metadata are
//!     delegated to the wrapped backend (the embedded redb one), every upsert / delete is mirrored into HBM corpora
//!     (`ucfp_corpus_upsert` / `ucfp_corpus_delete`: insert-or-replace and idempotent delete on the device, nothing is
//!     rebuilt), and `knn` is answered by the GPU scan through a query batcher that coalesces the single-query requests of the
//!     tokio workers (<= 512 in flight, `src/bin/ucfp.rs:262-267`) into tensor-path batches;
//!   * `HashIndex` adds the queries the reference stores fingerprints for but cannot ask: Hamming, MinHash-Jaccard and the
//!     block-hash-aware multi-hash re-rank.
//!
//! STATUS: written against the reference's crate layout, NOT compiled in this repository's development container (no Rust
//! toolchain there -- `cargo`, `rustc` absent, no network).  The Python package `ucfp_b200` binds the same ABI through ctypes
//! with the same semantics and is what the test-suite and the benchmark run; treat this file as the binding a maintainer
//! drops into `crates/ucfp-cuda/` and compiles, not as tested code.
#![allow(non_camel_case_types)]

pub mod sys {
    use std::os::raw::{c_char, c_int, c_void};
    #[repr(C)] pub struct ucfp_ctx { _p: [u8; 0] }
    #[repr(C)] pub struct ucfp_corpus { _p: [u8; 0] }
    #[repr(C)] pub struct ucfp_batcher { _p: [u8; 0] }
    #[repr(C)] pub struct ucfp_group { _p: [u8; 0] }
    #[repr(C)] #[derive(Copy, Clone)]
    pub struct ucfp_image_desc { pub pixels: *const u8, pub width: u32, pub height: u32, pub stride: u64 }
    #[repr(C)] #[derive(Copy, Clone, bytemuck::Zeroable, bytemuck::Pod)]
    pub struct ucfp_hash17 { pub global_hash: u64, pub block_hashes: [u64; 16] }
    #[repr(C)] #[derive(Copy, Clone, bytemuck::Zeroable, bytemuck::Pod)]
    pub struct ucfp_image_hashes { pub ahash: ucfp_hash17, pub phash: ucfp_hash17, pub dhash: ucfp_hash17 }
    #[repr(C)] #[derive(Copy, Clone)]
    pub struct ucfp_multihash_config { pub ahash_weight: f32, pub phash_weight: f32, pub dhash_weight: f32, pub global_weight: f32,
                                       pub block_weight: f32, pub block_distance_threshold: u32 }
    pub const UCFP_OK: c_int = 0;
    pub const UCFP_E_UNSUPPORTED: c_int = -4;
    pub const UCFP_ALGO_AHASH: u32 = 1; pub const UCFP_ALGO_PHASH: u32 = 2; pub const UCFP_ALGO_DHASH: u32 = 4; pub const UCFP_ALGO_MULTI: u32 = 7;
    pub const UCFP_KIND_HAMMING64: c_int = 1; pub const UCFP_KIND_MINHASH128: c_int = 2; pub const UCFP_KIND_COSINE: c_int = 3;
    pub const UCFP_KIND_MULTIHASH: c_int = 4;
    pub const UCFP_ID_NONE: u64 = u64::MAX;
    unsafe extern "C" {
        pub fn ucfp_init(device: c_int, out: *mut *mut ucfp_ctx) -> c_int;
        pub fn ucfp_destroy(ctx: *mut ucfp_ctx);
        pub fn ucfp_last_error() -> *const c_char;
        pub fn ucfp_ctx_synchronize(ctx: *mut ucfp_ctx) -> c_int;
        pub fn ucfp_image_hash_batch(ctx: *mut ucfp_ctx, imgs: *const ucfp_image_desc, n: usize, algo_mask: u32,
                                     out: *mut ucfp_image_hashes, status: *mut i32) -> c_int;
        pub fn ucfp_image_hash_jpeg_batch(ctx: *mut ucfp_ctx, jpegs: *const *const u8, lengths: *const usize, n: usize, algo_mask: u32,
                                          out: *mut ucfp_image_hashes, status: *mut i32, dims_out: *mut u32, pixels_out: *mut u8,
                                          pixels_capacity: usize) -> c_int;
        pub fn ucfp_corpus_create(ctx: *mut ucfp_ctx, kind: c_int, dim: u32, capacity: u64, out: *mut *mut ucfp_corpus) -> c_int;
        pub fn ucfp_corpus_destroy(c: *mut ucfp_corpus);
        pub fn ucfp_corpus_append_strided(c: *mut ucfp_corpus, ids: *const u64, records: *const c_void, record_stride: u64,
                                          field_offset: u64, n: u64) -> c_int;
        pub fn ucfp_corpus_upsert(c: *mut ucfp_corpus, ids: *const u64, rows: *const c_void, n: u64, n_replaced: *mut u64) -> c_int;
        pub fn ucfp_corpus_delete(c: *mut ucfp_corpus, ids: *const u64, n: u64, n_removed: *mut u64) -> c_int;
        pub fn ucfp_corpus_size(c: *const ucfp_corpus) -> u64;
        pub fn ucfp_scan_hamming(c: *mut ucfp_corpus, q: *const u64, nq: usize, k: usize, ids: *mut u64, dist: *mut u32) -> c_int;
        pub fn ucfp_scan_jaccard(c: *mut ucfp_corpus, q: *const u64, nq: usize, k: usize, ids: *mut u64, matches: *mut u32) -> c_int;
        pub fn ucfp_scan_cosine(c: *mut ucfp_corpus, q: *const f32, nq: usize, k: usize, ids: *mut u64, score: *mut f32) -> c_int;
        pub fn ucfp_scan_multihash(c: *mut ucfp_corpus, q: *const ucfp_image_hashes, nq: usize, k_prime: usize, k: usize,
                                   cfg: *const ucfp_multihash_config, ids: *mut u64, score: *mut f32) -> c_int;
        pub fn ucfp_batcher_create(c: *mut ucfp_corpus, max_batch: u32, max_delay_us: u32, out: *mut *mut ucfp_batcher) -> c_int;
        pub fn ucfp_batcher_destroy(b: *mut ucfp_batcher);
        pub fn ucfp_batcher_query(b: *mut ucfp_batcher, query: *const c_void, k: usize, ids: *mut u64, keys: *mut c_void) -> c_int;
        pub fn ucfp_group_create(devices: *const c_int, n: c_int, out: *mut *mut ucfp_group) -> c_int;
        pub fn ucfp_group_destroy(g: *mut ucfp_group);
        pub fn ucfp_group_ctx(g: *mut ucfp_group, local_rank: c_int) -> *mut ucfp_ctx;
        pub fn ucfp_group_scan_hamming(g: *mut ucfp_group, corpora: *const *mut ucfp_corpus, q: *const u64, nq: usize, k: usize,
                                       ids: *mut u64, dist: *mut u32) -> c_int;
        // gauges of the most recent scan on a context: queries whose candidate list overflowed (and were re-scanned), the longest list,
        // queries that reached the exact multi-pass selection -- worth exporting next to the reference's own request metrics
        pub fn ucfp_ctx_last_scan_stats(ctx: *mut ucfp_ctx, queries_recomputed: *mut u64, max_list_fill: *mut u64) -> c_int;
        pub fn ucfp_ctx_last_scan_exact_selects(ctx: *mut ucfp_ctx, queries: *mut u64) -> c_int;
    }
}

use std::{collections::HashMap, ffi::CStr, sync::{Arc, RwLock}};
use bytes::Bytes;
use ucfp::core::{FingerprintMeta, Hit, HitSource, Modality, Record};
use ucfp::error::{Error, Result};
use ucfp::index::IndexBackend;

fn last_error() -> String { unsafe { CStr::from_ptr(sys::ucfp_last_error()) }.to_string_lossy().into_owned() }
fn check_index(rc: i32) -> Result<()> { if rc == 0 { Ok(()) } else { Err(Error::Index(last_error())) } }

/// One per (process, GPU).  The library is thread-safe (per-call lanes: private stream + scratch), so `&Gpu` is shared freely.
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
    // Spec-v1 hashes carry their own tags until bit parity with imgfprint is pinned by golden vectors (docs/HASH_SPEC.md;
    // ucfp_b200/image.py IMGFPRINT_PARITY_VERIFIED): an index keys its Hamming corpora by tag, mixing definitions is silent garbage.
    pub const ALGORITHM_MULTIHASH: &str = "ucfp-b200-multihash-v1";
    pub const ALGORITHM_PHASH: &str = "ucfp-b200-phash-v1";
    pub const ALGORITHM_DHASH: &str = "ucfp-b200-dhash-v1";
    pub const ALGORITHM_AHASH: &str = "ucfp-b200-ahash-v1";

    fn record(exact: [u8; 32], h: &sys::ucfp_image_hashes, algo_mask: u32, tenant_id: u32, record_id: u64) -> Record {
        let single = |h17: &sys::ucfp_hash17| { let mut b = Vec::with_capacity(168); b.extend_from_slice(&exact); b.extend_from_slice(bytemuck::bytes_of(h17)); b };
        let (tag, blob) = match algo_mask {
            sys::UCFP_ALGO_MULTI => { let mut b = Vec::with_capacity(536); b.extend_from_slice(&exact); for h17 in [&h.ahash, &h.phash, &h.dhash] { b.extend_from_slice(&single(h17)); } (ALGORITHM_MULTIHASH, b) }
            sys::UCFP_ALGO_PHASH => (ALGORITHM_PHASH, single(&h.phash)),
            sys::UCFP_ALGO_DHASH => (ALGORITHM_DHASH, single(&h.dhash)),
            _ => (ALGORITHM_AHASH, single(&h.ahash)),
        };
        Record { tenant_id, record_id, modality: Modality::Image, format_version: 1, algorithm: tag.into(), config_hash: 0,
                 fingerprint: Bytes::from(blob), embedding: None, model_id: None, metadata: Bytes::new(), text: None }
    }

    /// Batched `ucfp::image::fingerprint*`: one GPU call for the whole batch, one `Result<Record>` per image
    /// (errors map to `Error::Modality`, as at `src/modality/image.rs:70`).
    pub fn fingerprint_batch(gpu: &Gpu, imgs: &[DecodedRgb<'_>], algo_mask: u32, tenant_id: u32, record_ids: &[u64]) -> Vec<Result<Record>> {
        let descs: Vec<_> = imgs.iter().map(|i| sys::ucfp_image_desc { pixels: i.pixels.as_ptr(), width: i.width, height: i.height, stride: 3 * i.width as u64 }).collect();
        let mut out: Vec<sys::ucfp_image_hashes> = vec![bytemuck::Zeroable::zeroed(); imgs.len()];
        let mut status = vec![0i32; imgs.len()];
        let rc = unsafe { sys::ucfp_image_hash_batch(gpu.ctx, descs.as_ptr(), descs.len(), algo_mask, out.as_mut_ptr(), status.as_mut_ptr()) };
        let batch_err = if rc != 0 { Some(last_error()) } else { None };
        imgs.iter().enumerate().map(|(i, img)| {
            if let Some(e) = &batch_err { return Err(Error::Modality(e.clone())); }
            if status[i] != 0 { return Err(Error::Modality(format!("image {i}: hash status {}", status[i]))); }
            Ok(record(*blake3::hash(img.encoded).as_bytes(), &out[i], algo_mask, tenant_id, record_ids[i]))
        }).collect()
    }

    /// From ENCODED bytes: JPEGs are decoded on the device (nvJPEG) and hashed by the same call.  `Err(Error::Unsupported)` for
    /// an entry means "not a JPEG / refused by nvJPEG": decode it with the `image` crate and use `fingerprint_batch`.
    pub fn fingerprint_jpeg_batch(gpu: &Gpu, encoded: &[&[u8]], algo_mask: u32, tenant_id: u32, record_ids: &[u64]) -> Vec<Result<Record>> {
        let ptrs: Vec<*const u8> = encoded.iter().map(|e| e.as_ptr()).collect();
        let lens: Vec<usize> = encoded.iter().map(|e| e.len()).collect();
        let mut out: Vec<sys::ucfp_image_hashes> = vec![bytemuck::Zeroable::zeroed(); encoded.len()];
        let mut status = vec![0i32; encoded.len()];
        let rc = unsafe { sys::ucfp_image_hash_jpeg_batch(gpu.ctx, ptrs.as_ptr(), lens.as_ptr(), ptrs.len(), algo_mask, out.as_mut_ptr(),
                                                          status.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut(), 0) };
        let batch_err = if rc != 0 { Some(last_error()) } else { None };
        encoded.iter().enumerate().map(|(i, e)| {
            if let Some(m) = &batch_err { return Err(Error::Modality(m.clone())); }
            match status[i] {
                0 => Ok(record(*blake3::hash(e).as_bytes(), &out[i], algo_mask, tenant_id, record_ids[i])),
                sys::UCFP_E_UNSUPPORTED => Err(Error::Unsupported("not decodable on the device: decode on the host".into())),
                s => Err(Error::Modality(format!("image {i}: decode/hash status {s}"))),
            }
        }).collect()
    }
}

/// One HBM corpus plus the batcher that serves single-query callers from it.
struct Shelf { corpus: *mut sys::ucfp_corpus, batcher: *mut sys::ucfp_batcher }
unsafe impl Send for Shelf {} unsafe impl Sync for Shelf {}
impl Shelf {
    fn new(gpu: &Gpu, kind: i32, dim: u32) -> Result<Self> {
        let (mut corpus, mut batcher) = (std::ptr::null_mut(), std::ptr::null_mut());
        check_index(unsafe { sys::ucfp_corpus_create(gpu.ctx, kind, dim, 1024, &mut corpus) })?;   // upsert grows it
        if let Err(e) = check_index(unsafe { sys::ucfp_batcher_create(corpus, 512, 200, &mut batcher) }) {
            unsafe { sys::ucfp_corpus_destroy(corpus) };
            return Err(e);
        }
        Ok(Self { corpus, batcher })
    }
    fn len(&self) -> usize { unsafe { sys::ucfp_corpus_size(self.corpus) as usize } }
}
impl Drop for Shelf { fn drop(&mut self) { unsafe { sys::ucfp_batcher_destroy(self.batcher); sys::ucfp_corpus_destroy(self.corpus) } } }

/// Scan seam: the HBM mirror of what the wrapped backend stores.  One corpus per (tenant, dim) for vectors, per
/// (tenant, algorithm tag) for 64-bit hash codes and for 51-word multi bundles, per tenant for MinHash signatures.
pub struct GpuIndexBackend {
    gpu: Gpu,
    store: Arc<dyn IndexBackend>,                      // redb tables, BM25, metadata: unchanged, out of scope for the GPU
    vectors: RwLock<HashMap<(u32, usize), Arc<Shelf>>>,
    hashes: RwLock<HashMap<(u32, String), Arc<Shelf>>>,
    bundles: RwLock<HashMap<(u32, String), Arc<Shelf>>>,
    signatures: RwLock<HashMap<u32, Arc<Shelf>>>,
}

const K_LIMIT_COSINE: usize = 1024;   // include/ucfp_cuda.h "Limits"
const K_LIMIT_HASH: usize = 2048;

fn hit(tenant_id: u32, record_id: u64, score: f32) -> Hit {
    Hit { tenant_id, record_id, score, source: HitSource::Vector, vector_score: None, bm25_score: None, vector_rank: None,
          bm25_rank: None, term_hits: Vec::new() }
}

/// `EmbeddedBackend::knn` returns min(k, N) hits; the scans cap k per call.
fn clamp_k(k: usize, rows: usize, limit: usize) -> Result<usize> {
    let k = k.min(rows);
    if k > limit { Err(Error::Unsupported(format!("k = {k} exceeds the scan limit of {limit} results per query"))) } else { Ok(k) }
}

fn shelf_of<K: std::hash::Hash + Eq + Clone>(map: &RwLock<HashMap<K, Arc<Shelf>>>, key: &K, gpu: &Gpu, kind: i32, dim: u32) -> Result<Arc<Shelf>> {
    if let Some(s) = map.read().unwrap().get(key) { return Ok(s.clone()); }
    let mut w = map.write().unwrap();
    if let Some(s) = w.get(key) { return Ok(s.clone()); }
    let s = Arc::new(Shelf::new(gpu, kind, dim)?);
    w.insert(key.clone(), s.clone());
    Ok(s)
}

impl GpuIndexBackend {
    pub fn new(device: i32, store: Arc<dyn IndexBackend>) -> Result<Self> {
        Ok(Self { gpu: Gpu::new(device)?, store, vectors: Default::default(), hashes: Default::default(), bundles: Default::default(),
                  signatures: Default::default() })
    }

    /// Mirrors a batch into HBM: rows grouped per shelf, one `ucfp_corpus_upsert` per shelf (insert-or-replace on the device).
    fn mirror_upsert(&self, batch: &[Record]) -> Result<()> {
        let mut groups: HashMap<*mut sys::ucfp_corpus, (Arc<Shelf>, Vec<u64>, Vec<u8>)> = HashMap::new();
        let mut push = |s: Arc<Shelf>, id: u64, row: &[u8]| {
            let e = groups.entry(s.corpus).or_insert_with(|| (s.clone(), Vec::new(), Vec::new()));
            e.1.push(id); e.2.extend_from_slice(row);
        };
        for r in batch {
            if let Some(v) = r.embedding.as_ref().filter(|v| !v.is_empty()) {
                push(shelf_of(&self.vectors, &(r.tenant_id, v.len()), &self.gpu, sys::UCFP_KIND_COSINE, v.len() as u32)?, r.record_id, bytemuck::cast_slice(v));
            }
            let fp = &r.fingerprint;
            let is_multi = r.algorithm.ends_with("-multihash-v1") && fp.len() == 536;
            let is_single = (r.algorithm.ends_with("-phash-v1") || r.algorithm.ends_with("-dhash-v1") || r.algorithm.ends_with("-ahash-v1")) && fp.len() == 168;
            if is_multi || is_single {
                let off = if is_multi { 232 } else { 32 };   // PHash global hash of a bundle / global_hash of an ImageFingerprint
                push(shelf_of(&self.hashes, &(r.tenant_id, r.algorithm.clone()), &self.gpu, sys::UCFP_KIND_HAMMING64, 0)?, r.record_id, &fp[off..off + 8]);
            }
            if is_multi {
                let mut words = Vec::with_capacity(408);
                for a in 0..3 { words.extend_from_slice(&fp[64 + 168 * a..200 + 168 * a]); }
                push(shelf_of(&self.bundles, &(r.tenant_id, r.algorithm.clone()), &self.gpu, sys::UCFP_KIND_MULTIHASH, 0)?, r.record_id, &words);
            }
            if r.algorithm == "minhash-h128" && fp.len() == 1032 && fp[..8] == [1, 0, 0, 0, 0, 0, 0, 0] {
                push(shelf_of(&self.signatures, &r.tenant_id, &self.gpu, sys::UCFP_KIND_MINHASH128, 0)?, r.record_id, &fp[8..]);
            }
        }
        for (_, (shelf, ids, rows)) in groups {
            check_index(unsafe { sys::ucfp_corpus_upsert(shelf.corpus, ids.as_ptr(), rows.as_ptr().cast(), ids.len() as u64, std::ptr::null_mut()) })?;
        }
        Ok(())
    }

    fn mirror_delete(&self, tenant_id: u32, ids: &[u64]) -> Result<()> {
        let mut shelves: Vec<Arc<Shelf>> = Vec::new();
        shelves.extend(self.vectors.read().unwrap().iter().filter(|(k, _)| k.0 == tenant_id).map(|(_, s)| s.clone()));
        shelves.extend(self.hashes.read().unwrap().iter().filter(|(k, _)| k.0 == tenant_id).map(|(_, s)| s.clone()));
        shelves.extend(self.bundles.read().unwrap().iter().filter(|(k, _)| k.0 == tenant_id).map(|(_, s)| s.clone()));
        shelves.extend(self.signatures.read().unwrap().get(&tenant_id).cloned());
        for s in shelves {
            check_index(unsafe { sys::ucfp_corpus_delete(s.corpus, ids.as_ptr(), ids.len() as u64, std::ptr::null_mut()) })?;
        }
        Ok(())
    }

    /// Blocking single-query cosine k-NN through the batcher (call it inside `spawn_blocking`, as the embedded backend does).
    pub fn knn_blocking(&self, tenant_id: u32, query: &[f32], k: usize) -> Result<Vec<Hit>> {
        if query.is_empty() || k == 0 { return Ok(Vec::new()); }                    // embedded/mod.rs:275
        let Some(shelf) = self.vectors.read().unwrap().get(&(tenant_id, query.len())).cloned() else { return Ok(Vec::new()) };
        let k = clamp_k(k, shelf.len(), K_LIMIT_COSINE)?;
        if k == 0 { return Ok(Vec::new()); }
        let (mut ids, mut scores) = (vec![0u64; k], vec![0f32; k]);
        check_index(unsafe { sys::ucfp_batcher_query(shelf.batcher, query.as_ptr().cast(), k, ids.as_mut_ptr(), scores.as_mut_ptr().cast()) })?;
        Ok(ids.into_iter().zip(scores).filter(|(id, _)| *id != sys::UCFP_ID_NONE).map(|(id, s)| hit(tenant_id, id, s)).collect())
    }
}

#[async_trait::async_trait]
impl IndexBackend for GpuIndexBackend {
    async fn upsert(&self, batch: &[Record]) -> Result<()> {
        self.store.upsert(batch).await?;      // the durable write first (one redb transaction per batch, embedded/mod.rs:157-227)
        // a record that changes shape (new dimension / algorithm) must leave its old corpora: insert-or-replace is per (tenant, record)
        let mut by_tenant: HashMap<u32, Vec<u64>> = HashMap::new();
        for r in batch { by_tenant.entry(r.tenant_id).or_default().push(r.record_id); }
        for (t, ids) in &by_tenant { self.mirror_delete(*t, ids)?; }
        self.mirror_upsert(batch)
    }
    async fn delete(&self, tenant_id: u32, ids: &[u64]) -> Result<()> {
        self.store.delete(tenant_id, ids).await?;
        self.mirror_delete(tenant_id, ids)
    }
    async fn knn(&self, tenant_id: u32, query: &[f32], k: usize, _filter: Option<&Bytes>) -> Result<Vec<Hit>> {
        // `filter` is ignored exactly as in the embedded backend (embedded/mod.rs:268-273).  The FFI call blocks its thread
        // until the batch it rode in has been scanned: run it off the async workers, like EmbeddedBackend (:282).
        let (q, this) = (query.to_vec(), self as *const Self as usize);
        tokio::task::spawn_blocking(move || unsafe { &*(this as *const Self) }.knn_blocking(tenant_id, &q, k))
            .await.map_err(|e| Error::Index(e.to_string()))?
    }
    async fn bm25(&self, tenant_id: u32, terms: &[&str], k: usize, filter: Option<&Bytes>) -> Result<Vec<Hit>> { self.store.bm25(tenant_id, terms, k, filter).await }
    async fn bm25_explain(&self, tenant_id: u32, terms: &[&str], k: usize, filter: Option<&Bytes>) -> Result<Vec<Hit>> { self.store.bm25_explain(tenant_id, terms, k, filter).await }
    async fn flush(&self) -> Result<()> { check_index(unsafe { sys::ucfp_ctx_synchronize(self.gpu.ctx) })?; self.store.flush().await }
    async fn get_record_metadata(&self, tenant_id: u32, record_id: u64) -> Result<FingerprintMeta> { self.store.get_record_metadata(tenant_id, record_id).await }
}

/// The queries the reference stores fingerprints for but cannot ask (`src/index/mod.rs:29-35` has `knn` only).
/// `Hit::score` = `1 - dist / 64` (Hamming), `matches / 128` (Jaccard) or the blended multi-hash similarity; `source = Vector`.
pub trait HashIndex {
    fn hamming_knn_blocking(&self, tenant_id: u32, algorithm: &str, code: u64, k: usize) -> Result<Vec<Hit>>;
    fn jaccard_knn_blocking(&self, tenant_id: u32, signature: &[u64; 128], k: usize) -> Result<Vec<Hit>>;
    /// docs/HASH_SPEC.md section 10: the k' nearest PHash global hashes re-ranked with the block hashes; `cfg = None` = the defaults
    /// of `MultiHashConfigDto` (src/server/dto.rs:462-480).
    fn multihash_knn_blocking(&self, tenant_id: u32, algorithm: &str, bundle: &[u8; 536], k: usize, k_prime: usize,
                              cfg: Option<sys::ucfp_multihash_config>) -> Result<Vec<Hit>>;
}

impl HashIndex for GpuIndexBackend {
    fn hamming_knn_blocking(&self, tenant_id: u32, algorithm: &str, code: u64, k: usize) -> Result<Vec<Hit>> {
        let Some(shelf) = self.hashes.read().unwrap().get(&(tenant_id, algorithm.to_string())).cloned() else { return Ok(Vec::new()) };
        let k = clamp_k(k, shelf.len(), K_LIMIT_HASH)?;
        if k == 0 { return Ok(Vec::new()); }
        let (mut ids, mut dist) = (vec![0u64; k], vec![0u32; k]);
        check_index(unsafe { sys::ucfp_batcher_query(shelf.batcher, (&code as *const u64).cast(), k, ids.as_mut_ptr(), dist.as_mut_ptr().cast()) })?;
        Ok(ids.into_iter().zip(dist).filter(|(id, _)| *id != sys::UCFP_ID_NONE).map(|(id, d)| hit(tenant_id, id, 1.0 - d as f32 / 64.0)).collect())
    }
    fn jaccard_knn_blocking(&self, tenant_id: u32, signature: &[u64; 128], k: usize) -> Result<Vec<Hit>> {
        let Some(shelf) = self.signatures.read().unwrap().get(&tenant_id).cloned() else { return Ok(Vec::new()) };
        let k = clamp_k(k, shelf.len(), K_LIMIT_HASH)?;
        if k == 0 { return Ok(Vec::new()); }
        let (mut ids, mut m) = (vec![0u64; k], vec![0u32; k]);
        check_index(unsafe { sys::ucfp_batcher_query(shelf.batcher, signature.as_ptr().cast(), k, ids.as_mut_ptr(), m.as_mut_ptr().cast()) })?;
        Ok(ids.into_iter().zip(m).filter(|(id, _)| *id != sys::UCFP_ID_NONE).map(|(id, x)| hit(tenant_id, id, x as f32 / 128.0)).collect())
    }
    fn multihash_knn_blocking(&self, tenant_id: u32, algorithm: &str, bundle: &[u8; 536], k: usize, k_prime: usize,
                              cfg: Option<sys::ucfp_multihash_config>) -> Result<Vec<Hit>> {
        let Some(shelf) = self.bundles.read().unwrap().get(&(tenant_id, algorithm.to_string())).cloned() else { return Ok(Vec::new()) };
        let k = clamp_k(k, shelf.len(), K_LIMIT_HASH)?;
        if k == 0 { return Ok(Vec::new()); }
        let kp = k_prime.max(k).min(shelf.len()).min(K_LIMIT_HASH);
        let mut q: sys::ucfp_image_hashes = bytemuck::Zeroable::zeroed();
        for a in 0..3 { bytemuck::bytes_of_mut(&mut q)[136 * a..136 * a + 136].copy_from_slice(&bundle[64 + 168 * a..200 + 168 * a]); }
        let (mut ids, mut sc) = (vec![0u64; k], vec![0f32; k]);
        check_index(unsafe { sys::ucfp_scan_multihash(shelf.corpus, &q, 1, kp, k, cfg.as_ref().map_or(std::ptr::null(), |c| c as *const _),
                                                      ids.as_mut_ptr(), sc.as_mut_ptr()) })?;
        Ok(ids.into_iter().zip(sc).filter(|(id, _)| *id != sys::UCFP_ID_NONE).map(|(id, s)| hit(tenant_id, id, s)).collect())
    }
}
