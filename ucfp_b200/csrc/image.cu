// image.cu -- batched perceptual hashing of decoded RGB8 images (sm_100a): the multi bundle
// AHash + PHash + DHash, each as 1 global + 16 block hashes (docs/HASH_SPEC.md sections 1-5).
// Replaces the calls into imgfprint at src/modality/image.rs:68-70,175-179 of the reference.
//
// Compiled with -fmad=false: the spec fixes separately rounded f32 multiply and add in a fixed
// summation order, so that hash bits are identical on every implementation.
//
// Two kernels, one CTA per image:
//   image_stream_kernel<CPT>  -- the fast path for ordinary shapes.  Every thread owns CPT adjacent
//       columns and walks the image top to bottom exactly once (the only HBM traffic: 3*w*h bytes).
//       Per pixel: integer luma, then the four vertical Triangle passes (whole->32, whole->8,
//       block->32, block->8) as sliding accumulators in registers -- a row contributes to at most three
//       output rows of a pass.  A finished vertical row goes to a shared-memory row buffer; at the end of
//       each band of rows the CTA runs the horizontal passes (->32, ->9, ->8) over the buffered rows and
//       keeps only the u8 grids (17 regions x (32x32 + 9x8 + 8x8) bytes) in shared memory.
//   image_generic_kernel     -- any shape (tiny, up-scaling, extreme aspect, very wide): output-major
//       passes over a gray plane in global scratch.
// Both end in hash_regions(): 17 x {8x32 . 32x32 . 32x8 partial DCT-II, median, mean, gradient, ballot}.
#include <stdlib.h>

#include <algorithm>
#include <map>

#include "common.cuh"
#include "dct_table.h"

namespace ucfp {
namespace {

__constant__ float c_dct_cos[8][32] = { UCFP_DCT_VALUES };  // same literals as the host table

constexpr int kRegions = 17;                 // whole image + 4x4 blocks
constexpr int kHOuts = 49;                   // 32 + 9 + 8 horizontal outputs per region column set
constexpr int kVOuts = 40;                   // 32 + 8 vertical outputs per region row set
constexpr size_t kRowBufBudget = 32 * 1024;  // bytes of shared memory for buffered vertical rows
constexpr int kMaxBandRows = 64;

struct Taps { int left, n, woff; };                         // one output sample of a 1-D Triangle pass
struct __align__(16) VEntry { float w[3]; uint8_t n_active, n_finish; uint16_t pad; };   // host-side, per (pass, row)
// Device-side per-row table of the four vertical passes: weights of the two lowest active outputs, the third
// weight, and flags (byte s: bit 0 = three outputs active, bits 1..3 = outputs completed by this row).
struct __align__(16) RowEntry { float w01[4][2]; float w2[4]; uint32_t flags; uint32_t run; uint32_t pad[2]; };   // 64 bytes; run = rows from here on with flags == 0
struct FinDesc { uint8_t stream, o, r, pad; };               // stream: 0 whole->32, 1 whole->8, 2 block->32, 3 block->8

struct ShapeDev {
    int w, h;
    int by[5], bx[5];
    const Taps *hout;      // [5][49]: column set 0 = whole width, 1 + c = block column c; absolute x
    const Taps *vout;      // [5][40]: row set 0 = whole height, 1 + r = block row r; absolute y
    const float *wts;
    const RowEntry *rowtab; // [h] (stream kernel)
    const FinDesc *fin;    // finished rows in emission order
    const int *band_off;   // [nbands + 1] into fin
    const uint32_t *hitems; // stream kernel: every horizontal output of every finished row, band after band, in emission order:
                            //   tap index into hout << 24 | row's slot in the band's row buffer << 16 | byte offset into RegionGrids
    const int *hitem_off;  // [nbands + 1] into hitems
    int band_rows, nbands, max_fin;
};

struct RegionGrids {         // shared memory, per CTA, lives for the whole image
    uint8_t g32[kRegions][1024];
    uint8_t g98[kRegions][72];
    uint8_t g8[kRegions][64];
};
struct HashScratch {         // shared memory used only by hash_regions (may alias streaming buffers)
    float R[kRegions][32][8];
    float D[kRegions][64];
    float Ct[32][8];             // DCT cosines transposed: Ct[x][v] = C[v][x], four frequencies per LDS.128
    float med_lo[kRegions], med_hi[kRegions];
};
constexpr size_t kGridsBytes = (sizeof(RegionGrids) + 127) & ~size_t(127);

struct ImgDev { const uint8_t *pixels; uint64_t stride; uint32_t out_index; uint32_t aligned4; /* 0 no, 1 = 4 B, 2 = 16 B rows (bulk) */ };

// ---------------------------------------------------------------------------------------------
// Host: tap tables.  Same f32 arithmetic as `image` 0.25 imageops::sample (spec section 2).
// ---------------------------------------------------------------------------------------------
struct HostTaps { int left, n; std::vector<float> w; };

std::vector<HostTaps> make_taps(int src, int dst) {
    std::vector<HostTaps> t(dst);
    volatile float ratio = (float)src / (float)dst;  // volatile: keep every step rounded to binary32
    float sratio = ratio < 1.0f ? 1.0f : ratio;
    float support = 1.0f * sratio;
    for (int o = 0; o < dst; ++o) {
        volatile float in = ((float)o + 0.5f) * ratio;
        long long left = (long long)floorf(in - support);
        if (left < 0) left = 0;
        if (left > src - 1) left = src - 1;
        long long right = (long long)ceilf(in + support);
        if (right < left + 1) right = left + 1;
        if (right > src) right = src;
        volatile float inc = in - 0.5f;
        int n = (int)(right - left);
        t[o].left = (int)left; t[o].n = n; t[o].w.resize(n);
        volatile float sum = 0.0f;
        for (int i = 0; i < n; ++i) {
            volatile float x = ((float)(left + i) - inc) / sratio;
            float a = fabsf(x);
            float wv = a < 1.0f ? 1.0f - a : 0.0f;
            t[o].w[i] = wv;
            sum = sum + wv;
        }
        for (int i = 0; i < n; ++i) { volatile float q = t[o].w[i] / sum; t[o].w[i] = q; }
    }
    return t;
}

struct StreamSmem { uint32_t off_rowbuf, off_ring, off_bars, off_rowtab, stage_bytes, stages, chunk_rows, row_pitch, row_words; };

constexpr size_t kShapeCacheMaxEntries = 256, kShapeCacheMaxBytes = size_t(256) << 20;

struct ShapeTables {
    ShapeDev dev{};
    void *blob = nullptr;
    size_t blob_bytes = 0;
    uint32_t in_use = 0;      // calls hashing this shape right now (never evicted while > 0)
    uint64_t last_use = 0;
    bool streamable = false;
    int threads = 0, cpt = 0;
    StreamSmem lay{};
    size_t stream_smem = 0;
};

struct ShapeCache { std::map<uint64_t, ShapeTables> m; uint64_t tick = 0; size_t bytes = 0; };   // lives in ucfp_ctx::image_cache, guarded by ucfp_ctx::image_mu

int build_shape(ucfp_lane *ctx, int w, int h, ShapeTables &st) {
    std::vector<Taps> hout(5 * kHOuts), vout(5 * kVOuts);
    std::vector<float> wts;
    ShapeDev &d = st.dev;
    d.w = w; d.h = h;
    for (int i = 0; i <= 4; ++i) { d.by[i] = (int)((long long)i * h / 4); d.bx[i] = (int)((long long)i * w / 4); }
    // Identical weight vectors are stored once (for integer ratios every interior output shares one vector), so
    // that the lanes of a warp working on different outputs read the same addresses (a broadcast).
    std::map<std::vector<uint32_t>, int> seen;
    auto emit = [&](std::vector<Taps> &dst, int base, const std::vector<HostTaps> &t, int offset) {
        for (size_t o = 0; o < t.size(); ++o) {
            std::vector<uint32_t> key(t[o].w.size());
            memcpy(key.data(), t[o].w.data(), key.size() * 4);
            auto it = seen.find(key);
            if (it == seen.end()) {
                it = seen.emplace(std::move(key), (int)wts.size()).first;
                wts.insert(wts.end(), t[o].w.begin(), t[o].w.end());
            }
            dst[base + o] = Taps{t[o].left + offset, t[o].n, it->second};
        }
    };
    // vertical tap sets (also the source of the streaming tables)
    std::vector<std::vector<HostTaps>> vsets(10);  // [set][0: ->32, 1: ->8]
    for (int s = 0; s < 5; ++s) {
        int y0 = s == 0 ? 0 : d.by[s - 1], len = s == 0 ? h : d.by[s] - d.by[s - 1];
        vsets[2 * s] = make_taps(len, 32);
        vsets[2 * s + 1] = make_taps(len, 8);
        emit(vout, s * kVOuts, vsets[2 * s], y0);
        emit(vout, s * kVOuts + 32, vsets[2 * s + 1], y0);
    }
    for (int s = 0; s < 5; ++s) {
        int x0 = s == 0 ? 0 : d.bx[s - 1], len = s == 0 ? w : d.bx[s] - d.bx[s - 1];
        emit(hout, s * kHOuts, make_taps(len, 32), x0);
        emit(hout, s * kHOuts + 32, make_taps(len, 9), x0);
        emit(hout, s * kHOuts + 41, make_taps(len, 8), x0);
    }

    // ---- streaming tables: per row and pass, the <= 3 active outputs and how many finish
    std::vector<VEntry> vtab((size_t)4 * h);
    std::vector<FinDesc> fin_rows;
    std::vector<int> fin_row_y;
    bool ok = w >= 32 && h >= 32;
    for (int stream = 0; stream < 4 && ok; ++stream) {
        for (int seg = 0; seg < (stream < 2 ? 1 : 4) && ok; ++seg) {
            int set = stream < 2 ? 0 : 1 + seg;
            const std::vector<HostTaps> &t = vsets[2 * set + (stream & 1)];
            int y0 = set == 0 ? 0 : d.by[set - 1], len = set == 0 ? h : d.by[set] - d.by[set - 1];
            int nout = (int)t.size();
            for (int o = 1; o < nout; ++o)  // windows must advance monotonically
                if (t[o].left < t[o - 1].left || t[o].left + t[o].n < t[o - 1].left + t[o - 1].n) ok = false;
            int o0 = 0;
            for (int yl = 0; yl < len && ok; ++yl) {
                while (o0 < nout && t[o0].left + t[o0].n <= yl) o0++;
                VEntry e{};
                int n = 0;
                for (int o = o0; o < nout && t[o].left <= yl; ++o) {
                    if (n == 3) { ok = false; break; }
                    e.w[n++] = t[o].w[yl - t[o].left];
                }
                // every output below the active range must already be complete, and active ones contiguous
                for (int o = o0; o < o0 + n; ++o) if (!(t[o].left <= yl && yl < t[o].left + t[o].n)) ok = false;
                int nf = 0;
                for (int o = o0; o < o0 + n; ++o) if (t[o].left + t[o].n - 1 == yl) nf++;
                for (int j = 0; j < nf; ++j) if (t[o0 + j].left + t[o0 + j].n - 1 != yl) ok = false;  // lowest finish first
                e.n_active = (uint8_t)n; e.n_finish = (uint8_t)nf;
                vtab[(size_t)stream * h + y0 + yl] = e;
            }
            if (ok) {  // every output must start no earlier than all lower outputs are accounted for
                for (int o = 0; o < nout; ++o) if (t[o].n < 1) ok = false;
            }
        }
    }
    std::vector<int> band_off;
    int band_rows = 0, max_fin = 0;
    if (ok) {
        // emission order: y ascending, stream ascending, finishing outputs ascending
        std::vector<int> next_out(4 * 5, 0);
        for (int y = 0; y < h; ++y)
            for (int stream = 0; stream < 4; ++stream) {
                int r = 0;
                if (stream >= 2) { while (r < 3 && y >= d.by[r + 1]) r++; }
                int key = stream * 5 + (stream >= 2 ? 1 + r : 0);
                for (int f = 0; f < vtab[(size_t)stream * h + y].n_finish; ++f) {
                    fin_rows.push_back(FinDesc{(uint8_t)stream, (uint8_t)next_out[key]++, (uint8_t)r, 0});
                    fin_row_y.push_back(y);
                }
            }
        if ((int)fin_rows.size() != 32 + 8 + 4 * 40) ok = false;
        // shared-memory budgets (tunable for experiments through the environment; defaults chosen from measurements)
        static const long env_rowbuf = getenv("UCFP_IMG_ROWBUF_KB") ? atol(getenv("UCFP_IMG_ROWBUF_KB")) : 0;
        const size_t rowbuf_budget = env_rowbuf > 0 ? (size_t)env_rowbuf * 1024 : (w <= 512 ? 14 * 1024 : kRowBufBudget);
        for (int cand = kMaxBandRows; cand >= 1 && ok; cand >>= 1) {
            int nb = (h + cand - 1) / cand, mx = 0;
            std::vector<int> cnt(nb, 0);
            for (int y : fin_row_y) cnt[y / cand]++;
            for (int c : cnt) mx = c > mx ? c : mx;
            if ((size_t)mx * (w + w / 32 + 1) * 4 <= rowbuf_budget || cand == 1) {
                if ((size_t)mx * (w + w / 32 + 1) * 4 > rowbuf_budget) { ok = false; break; }
                band_rows = cand; max_fin = mx;
                band_off.assign(nb + 1, 0);
                for (int b = 0; b < nb; ++b) band_off[b + 1] = band_off[b] + cnt[b];
                break;
            }
        }
    }
    // flat list of the horizontal outputs per band: a thread of the stream kernel takes items tid, tid + nt, ... and finds
    // row, taps and destination byte in one word instead of walking every finished row of the band
    std::vector<uint32_t> hitems;
    std::vector<int> hitem_off;
    if (ok) {
        if (max_fin > 256) ok = false;
        hitem_off.push_back(0);
        for (size_t b = 0; ok && b + 1 < band_off.size(); ++b) {
            for (int fi = band_off[b]; fi < band_off[b + 1]; ++fi) {
                const FinDesc f = fin_rows[fi];
                const int items = f.stream == 0 ? 32 : f.stream == 1 ? 17 : f.stream == 2 ? 128 : 68;   // fin_items()
                for (int item = 0; item < items; ++item) {   // same mapping as hpass_item()
                    int colset = 0, sub = item;
                    if (f.stream >= 2) { const int per = f.stream == 2 ? 32 : 17; colset = 1 + item / per; sub = item % per; }
                    const int region = f.stream < 2 ? 0 : 1 + 4 * f.r + (colset - 1);
                    size_t dest; int tap;
                    if ((f.stream & 1) == 0) { dest = offsetof(RegionGrids, g32) + (size_t)region * 1024 + f.o * 32 + sub; tap = colset * kHOuts + sub; }
                    else if (sub < 9) { dest = offsetof(RegionGrids, g98) + (size_t)region * 72 + f.o * 9 + sub; tap = colset * kHOuts + 32 + sub; }
                    else { dest = offsetof(RegionGrids, g8) + (size_t)region * 64 + f.o * 8 + (sub - 9); tap = colset * kHOuts + 41 + (sub - 9); }
                    hitems.push_back((uint32_t)tap << 24 | (uint32_t)(fi - band_off[b]) << 16 | (uint32_t)dest);
                }
            }
            // Within a band the order of the items is free (each writes its own byte): sort them by tap count, so that the lanes
            // of a warp run the same number of taps.  Unsorted, a warp straddling 4-tap (block -> 32) and 15/16-tap (block -> 9/8)
            // outputs executed every unrolled remainder of the tap loop: 170 instead of ~75 instructions per 4-tap item.
            std::stable_sort(hitems.begin() + hitem_off.back(), hitems.end(),
                             [&](uint32_t a, uint32_t b) { return hout[a >> 24].n < hout[b >> 24].n; });
            hitem_off.push_back((int)hitems.size());
        }
    }
    static_assert(sizeof(RegionGrids) < 65536 && 5 * kHOuts <= 256, "hitems packing");
    static const long env_cpt2 = getenv("UCFP_IMG_NO_CPT2") ? 0 : 1;   // developer switch: one column per thread up to 512 px as before
    int cpt = w <= 512 ? ((env_cpt2 && w >= 128 && w % 2 == 0) ? 2 : 1) : (w <= 1024 ? 4 : 8);
    int threads = ((w + cpt - 1) / cpt + 31) / 32 * 32;
    if (threads > 512) ok = false;  // wider than 4096 px: generic kernel
    if (threads < 128) threads = 128;
    st.streamable = ok; st.threads = threads; st.cpt = cpt;
    if (ok) {   // shared-memory plan of image_stream_kernel
        auto up = [](size_t x, size_t a) { return (x + a - 1) / a * a; };
        StreamSmem &L = st.lay;
        L.row_pitch = (uint32_t)up(3 * (size_t)w, 16);
        static const long env_stage = getenv("UCFP_IMG_STAGE_KB") ? atol(getenv("UCFP_IMG_STAGE_KB")) : 0;
        static const long env_stages = getenv("UCFP_IMG_STAGES") ? atol(getenv("UCFP_IMG_STAGES")) : 0;
        const size_t stage_budget = env_stage > 0 ? (size_t)env_stage * 1024 : (w <= 512 ? 6 * 1024 : 24 * 1024);
        uint32_t rs = 1;
        while (rs * 2 <= (uint32_t)band_rows && (size_t)rs * 2 * L.row_pitch <= stage_budget) rs *= 2;
        L.chunk_rows = rs;
        L.stage_bytes = (uint32_t)up((size_t)rs * L.row_pitch, 128);
        L.off_rowbuf = (uint32_t)kGridsBytes;
        L.row_words = (uint32_t)(w + (w >> 5) + 1);       // skewed row: one pad word per 32
        L.off_ring = (uint32_t)(L.off_rowbuf + up((size_t)max_fin * L.row_words * 4, 128));
        L.stages = env_stages > 0 ? (uint32_t)env_stages : (w <= 512 ? 2 : ((L.off_ring + 3 * (size_t)L.stage_bytes + 2 * (size_t)band_rows * sizeof(RowEntry) <= 106 * 1024) ? 3 : 2));
        L.off_bars = L.off_ring + L.stages * L.stage_bytes;          // full[stages], empty[stages]
        L.off_rowtab = (uint32_t)up(L.off_bars + 16 * L.stages, 16);   // [2][band_rows] RowEntry, double-buffered
        size_t total = L.off_rowtab + 2 * (size_t)band_rows * sizeof(RowEntry);
        size_t hash_end = L.off_rowbuf + sizeof(HashScratch);
        st.stream_smem = up(total > hash_end ? total : hash_end, 128);
        if (st.stream_smem > 200 * 1024) st.streamable = false;
    }

    // ---- upload one blob
    auto align16 = [](size_t x) { return (x + 15) & ~size_t(15); };
    size_t off_h = 0, off_v = align16(off_h + hout.size() * sizeof(Taps)), off_w = align16(off_v + vout.size() * sizeof(Taps));
    std::vector<RowEntry> rowtab(ok ? h : 0);
    for (int y = 0; ok && y < h; ++y) {
        RowEntry e{};
        for (int st4 = 0; st4 < 4; ++st4) {
            const VEntry &v = vtab[(size_t)st4 * h + y];
            e.w01[st4][0] = v.w[0]; e.w01[st4][1] = v.w[1]; e.w2[st4] = v.w[2];
            e.flags |= (uint32_t)((v.n_active == 3 ? 1u : 0u) | ((uint32_t)v.n_finish << 1)) << (8 * st4);
        }
        rowtab[y] = e;
    }
    for (int y = h - 1; ok && y >= 0; --y)   // length of the run of flag-free rows starting at y
        rowtab[y].run = rowtab[y].flags ? 0u : 1u + (y + 1 < h ? rowtab[y + 1].run : 0u);
    size_t off_t = align16(off_w + wts.size() * 4), off_f = align16(off_t + rowtab.size() * sizeof(RowEntry));
    size_t off_b = align16(off_f + (ok ? fin_rows.size() * sizeof(FinDesc) : 0));
    size_t off_hi = align16(off_b + (ok ? band_off.size() * 4 : 0)), off_ho = align16(off_hi + (ok ? hitems.size() * 4 : 0));
    size_t total = align16(off_ho + (ok ? hitem_off.size() * 4 : 0)) + 16;
    std::vector<uint8_t> host(total, 0);
    memcpy(&host[off_h], hout.data(), hout.size() * sizeof(Taps));
    memcpy(&host[off_v], vout.data(), vout.size() * sizeof(Taps));
    memcpy(&host[off_w], wts.data(), wts.size() * 4);
    if (ok) {
        memcpy(&host[off_t], rowtab.data(), rowtab.size() * sizeof(RowEntry));
        memcpy(&host[off_f], fin_rows.data(), fin_rows.size() * sizeof(FinDesc));
        memcpy(&host[off_b], band_off.data(), band_off.size() * 4);
        memcpy(&host[off_hi], hitems.data(), hitems.size() * 4);
        memcpy(&host[off_ho], hitem_off.data(), hitem_off.size() * 4);
    }
    UCFP_CUDA_TRY(cudaMalloc(&st.blob, total));
    st.blob_bytes = total;
    UCFP_CUDA_TRY(cudaMemcpyAsync(st.blob, host.data(), total, cudaMemcpyHostToDevice, ctx->stream));
    UCFP_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // `host` dies at return
    uint8_t *b = static_cast<uint8_t *>(st.blob);
    d.hout = reinterpret_cast<const Taps *>(b + off_h);
    d.vout = reinterpret_cast<const Taps *>(b + off_v);
    d.wts = reinterpret_cast<const float *>(b + off_w);
    d.rowtab = ok ? reinterpret_cast<const RowEntry *>(b + off_t) : nullptr;
    d.fin = ok ? reinterpret_cast<const FinDesc *>(b + off_f) : nullptr;
    d.band_off = ok ? reinterpret_cast<const int *>(b + off_b) : nullptr;
    d.hitems = ok ? reinterpret_cast<const uint32_t *>(b + off_hi) : nullptr;
    d.hitem_off = ok ? reinterpret_cast<const int *>(b + off_ho) : nullptr;
    d.band_rows = band_rows; d.nbands = ok ? (h + band_rows - 1) / band_rows : 0; d.max_fin = max_fin;
    return UCFP_OK;
}

// ---------------------------------------------------------------------------------------------
// Device: shared pieces
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t luma_u8(uint32_t r, uint32_t g, uint32_t b) {
    // `image` 0.25 rgb_to_luma: (2126 R + 7152 G + 722 B) / 10000, truncating.  The sum is < 2^22, for
    // which floor(s / 10000) == (s * 429497) >> 32 exactly (error s * 2704 / 2^32 / 10000 < 1e-4).
    uint32_t s = 2126u * r + 7152u * g + 722u * b;
    return __umulhi(s, 429497u);
}

// One output of a horizontal pass over a row of f32 vertical sums: sequential sum, clamp, round half away.
// kSkew: the row is stored with one pad word per 32 (element x at x + (x >> 5)), so that lanes reading at a
// constant stride (the resize ratio, usually a power of two) fall into different shared-memory banks.
__device__ __forceinline__ int skew(int x) { return x + (x >> 5); }

template <bool kSkew>
__device__ __forceinline__ uint8_t hsample(const float *row, Taps t, const float *__restrict__ wts) {
    float acc = 0.0f;
    const float *wp = wts + t.woff;
    for (int i = 0; i < t.n; ++i) {
        const int x = t.left + i;
        acc = acc + row[kSkew ? skew(x) : x] * __ldg(wp + i);
    }
    acc = acc < 0.0f ? 0.0f : acc;
    acc = acc > 255.0f ? 255.0f : acc;
    float fl = truncf(acc);
    if (acc - fl >= 0.5f) fl += 1.0f;
    return (uint8_t)fl;
}

// Horizontal passes for one finished vertical row.  `item` enumerates the outputs of that row:
// whole->32 rows: 32 items; whole->8 rows: 9 + 8; block rows: x4 block columns.
__device__ __forceinline__ int fin_items(int stream) { return stream == 0 ? 32 : stream == 1 ? 17 : stream == 2 ? 128 : 68; }

template <bool kSkew>
__device__ __forceinline__ void hpass_item(RegionGrids &G, const ShapeDev &S, FinDesc f, int item, const float *row) {
    int colset = 0, sub = item;
    if (f.stream >= 2) { int per = f.stream == 2 ? 32 : 17; colset = 1 + item / per; sub = item % per; }
    int region = f.stream < 2 ? 0 : 1 + 4 * f.r + (colset - 1);
    const Taps *ho = S.hout + colset * kHOuts;
    if ((f.stream & 1) == 0) {
        G.g32[region][f.o * 32 + sub] = hsample<kSkew>(row, ho[sub], S.wts);
    } else if (sub < 9) {
        G.g98[region][f.o * 9 + sub] = hsample<kSkew>(row, ho[32 + sub], S.wts);
    } else {
        G.g8[region][f.o * 8 + (sub - 9)] = hsample<kSkew>(row, ho[41 + (sub - 9)], S.wts);
    }
}

// PHash / AHash / DHash of the 17 regions from the u8 grids in shared memory (spec sections 3-5).
__device__ void hash_regions(RegionGrids &G, HashScratch &H, uint32_t algo_mask, uint64_t *out /* 51 words */) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < 256; i += nt) H.Ct[i & 31][i >> 5] = c_dct_cos[i >> 5][i & 31];
    __syncthreads();
    if (algo_mask & UCFP_ALGO_PHASH) {
        // R[y][v] = sum_x g[y][x] * C[v][x], x ascending.  One thread = one grid row and FOUR frequencies: a pixel is converted
        // once for four products and the four cosines come with one LDS.128 from the transposed table.
        for (int it = tid; it < kRegions * 64; it += nt) {
            const int reg = it >> 6, y = (it >> 1) & 31, vh = it & 1;
            const uint32_t *g = reinterpret_cast<const uint32_t *>(&G.g32[reg][y * 32]);
            float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;
#pragma unroll
            for (int xw = 0; xw < 8; ++xw) {
                const uint32_t w = g[xw];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const float gv = (float)((w >> (8 * b)) & 255u);
                    const float4 c = *reinterpret_cast<const float4 *>(&H.Ct[4 * xw + b][4 * vh]);
                    t0 = t0 + gv * c.x; t1 = t1 + gv * c.y; t2 = t2 + gv * c.z; t3 = t3 + gv * c.w;
                }
            }
            *reinterpret_cast<float4 *>(&H.R[reg][y][4 * vh]) = make_float4(t0, t1, t2, t3);
        }
        __syncthreads();
        // D[u][v] = sum_y C[u][y] * R[y][v], y ascending; one thread = one u and four v
        for (int it = tid; it < kRegions * 16; it += nt) {
            const int reg = it >> 4, u = (it >> 1) & 7, vh = it & 1;
            float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;
#pragma unroll 8
            for (int y = 0; y < 32; ++y) {
                const float c = H.Ct[y][u];
                const float4 r = *reinterpret_cast<const float4 *>(&H.R[reg][y][4 * vh]);
                t0 = t0 + c * r.x; t1 = t1 + c * r.y; t2 = t2 + c * r.z; t3 = t3 + c * r.w;
            }
            *reinterpret_cast<float4 *>(&H.D[reg][u * 8 + 4 * vh]) = make_float4(t0, t1, t2, t3);
        }
        __syncthreads();
        for (int it = tid; it < kRegions * 64; it += nt) {        // ranks 31 and 32 of the 64 coefficients
            int reg = it >> 6, i = it & 63;
            float val = H.D[reg][i];
            int rank = 0;
            for (int j = 0; j < 64; ++j) { float o = H.D[reg][j]; rank += (o < val) || (o == val && j < i); }
            if (rank == 31) H.med_lo[reg] = val;
            if (rank == 32) H.med_hi[reg] = val;
        }
        __syncthreads();
    }
    const int warp = tid >> 5, lane = tid & 31, nwarps = nt >> 5;
    for (int reg = warp; reg < kRegions; reg += nwarps) {
        uint64_t a = 0, p = 0, dh = 0;
        if (algo_mask & UCFP_ALGO_AHASH) {
            uint32_t v0 = G.g8[reg][lane], v1 = G.g8[reg][lane + 32];
            uint32_t sum = v0 + v1;
            for (int s = 16; s > 0; s >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
            uint32_t lo = __ballot_sync(0xffffffffu, 64u * v0 > sum), hi = __ballot_sync(0xffffffffu, 64u * v1 > sum);
            a = (uint64_t)hi << 32 | lo;
        }
        if (algo_mask & UCFP_ALGO_PHASH) {
            float med = (H.med_lo[reg] + H.med_hi[reg]) * 0.5f;
            uint32_t lo = __ballot_sync(0xffffffffu, H.D[reg][lane] > med), hi = __ballot_sync(0xffffffffu, H.D[reg][lane + 32] > med);
            p = (uint64_t)hi << 32 | lo;
        }
        if (algo_mask & UCFP_ALGO_DHASH) {
            int i0 = lane, i1 = lane + 32;  // bit 8r+c compares g[r][c] > g[r][c+1] on the 9-wide grid
            const uint8_t *g = G.g98[reg];
            uint32_t lo = __ballot_sync(0xffffffffu, g[(i0 >> 3) * 9 + (i0 & 7)] > g[(i0 >> 3) * 9 + (i0 & 7) + 1]);
            uint32_t hi = __ballot_sync(0xffffffffu, g[(i1 >> 3) * 9 + (i1 & 7)] > g[(i1 >> 3) * 9 + (i1 & 7) + 1]);
            dh = (uint64_t)hi << 32 | lo;
        }
        if (lane == 0) { out[reg] = a; out[17 + reg] = p; out[34 + reg] = dh; }
    }
}

// ---------------------------------------------------------------------------------------------
// Generic kernel: gray plane and vertical sums in global scratch, output-major everywhere.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
image_generic_kernel(ShapeDev S, const ImgDev *__restrict__ imgs, uint32_t n, uint32_t algo_mask, uint8_t *scratch,
                     size_t scratch_per_cta, uint64_t *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RegionGrids &G = *reinterpret_cast<RegionGrids *>(smem_raw);
    HashScratch &H = *reinterpret_cast<HashScratch *>(smem_raw + kGridsBytes);
    const int w = S.w, h = S.h, tid = threadIdx.x, nt = blockDim.x;
    uint8_t *gray = scratch + (size_t)blockIdx.x * scratch_per_cta;
    float *T = reinterpret_cast<float *>(gray + (((size_t)w * h + 15) & ~size_t(15)));  // [5][40][w]
    for (uint32_t im = blockIdx.x; im < n; im += gridDim.x) {
        const ImgDev I = imgs[im];
        for (size_t p = tid; p < (size_t)w * h; p += nt) {
            int y = (int)(p / w), x = (int)(p % w);
            const uint8_t *px = I.pixels + (size_t)y * I.stride + 3 * (size_t)x;
            gray[p] = (uint8_t)luma_u8(px[0], px[1], px[2]);
        }
        __syncthreads();
        for (size_t it = tid; it < (size_t)5 * kVOuts * w; it += nt) {   // vertical passes
            int x = (int)(it % w), vo = (int)(it / w);
            Taps t = S.vout[vo];
            float acc = 0.0f;
            for (int i = 0; i < t.n; ++i) acc = acc + (float)gray[(size_t)(t.left + i) * w + x] * __ldg(S.wts + t.woff + i);
            T[it] = acc;
        }
        __syncthreads();
        for (int it = tid; it < 5 * kVOuts * 128; it += nt) {             // horizontal passes
            int vo = it >> 7, item = it & 127, set = vo / kVOuts, o = vo % kVOuts;
            FinDesc f;
            f.stream = (uint8_t)((set ? 2 : 0) + (o >= 32 ? 1 : 0));
            f.o = (uint8_t)(o >= 32 ? o - 32 : o);
            f.r = (uint8_t)(set ? set - 1 : 0); f.pad = 0;
            if (item < fin_items(f.stream)) hpass_item<false>(G, S, f, item, T + (size_t)vo * w);
        }
        __syncthreads();
        hash_regions(G, H, algo_mask, out + (size_t)I.out_index * 51);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Streaming kernel: one pass over the pixels, vertical sums in registers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier (UBLKCP in SASS)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// luma of the 4 pixels packed in the 12 bytes {a, b, c}: two IDP.2A per pixel, no byte extraction
#define UCFP_W16(lo, hi) ((uint32_t)(lo) | ((uint32_t)(hi) << 16))
__device__ __forceinline__ float luma_to_f32(uint32_t s) {
    // floor(s / 10000) as in luma_u8, then u32 -> f32 exactly by the 2^23 trick (value < 256)
    return __uint_as_float(__umulhi(s, 429497u) | 0x4B000000u) - 8388608.0f;
}
__device__ __forceinline__ void luma4(uint32_t a, uint32_t b, uint32_t c, float &v0, float &v1, float &v2, float &v3) {
    v0 = luma_to_f32(__dp2a_hi(UCFP_W16(722, 0), a, __dp2a_lo(UCFP_W16(2126, 7152), a, 0u)));     // R G B = a0 a1 a2
    v1 = luma_to_f32(__dp2a_lo(UCFP_W16(7152, 722), b, __dp2a_hi(UCFP_W16(0, 2126), a, 0u)));     //         a3 b0 b1
    v2 = luma_to_f32(__dp2a_lo(UCFP_W16(722, 0), c, __dp2a_hi(UCFP_W16(2126, 7152), b, 0u)));     //         b2 b3 c0
    v3 = luma_to_f32(__dp2a_hi(UCFP_W16(7152, 722), c, __dp2a_lo(UCFP_W16(0, 2126), c, 0u)));     //         c1 c2 c3
}

// {v * w0, v * w1} with one FMUL2.  The products feed SCALAR adds: ptxas contracts mul.f32x2 + add.f32x2 into
// FFMA2 even for .rn operands and -fmad=false, which would break the spec's separately rounded mul and add.
__device__ __forceinline__ void mul2(float v, uint64_t w01, float &p0, float &p1) {
    uint64_t vv, r;
    asm("mov.b64 %0, {%1, %1};" : "=l"(vv) : "f"(v));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(vv), "l"(w01));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(p0), "=f"(p1) : "l"(r));
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

template <int CPT, int MAXT, bool BULK>
__global__ void __launch_bounds__(MAXT)
image_stream_kernel(ShapeDev S, StreamSmem L, const ImgDev *__restrict__ imgs, uint32_t algo_mask, uint64_t *out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RegionGrids &G = *reinterpret_cast<RegionGrids *>(smem_raw);
    float *rowbuf = reinterpret_cast<float *>(smem_raw + L.off_rowbuf);          // [max_fin][row_words], skewed rows
    unsigned char *ring = smem_raw + L.off_ring;                                 // [stages][chunk_rows * row_pitch]
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + L.off_bars);        // [stages] TMA bytes landed
    uint64_t *empty = full + L.stages;                                           // [stages] every warp is done reading
    RowEntry *rtab = reinterpret_cast<RowEntry *>(smem_raw + L.off_rowtab);      // [2][band_rows]
    HashScratch &H = *reinterpret_cast<HashScratch *>(smem_raw + L.off_rowbuf);  // aliases the streaming buffers
    const int w = S.w, h = S.h, tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const ImgDev I = imgs[blockIdx.x];
    // threads past the last column group redo the last group's arithmetic and simply never store
    const int x0 = min(tid * CPT, ((w - 1) / CPT) * CPT);
    const bool owner = tid * CPT < w;
    const int RS = (int)L.chunk_rows, nchunks = (h + RS - 1) / RS, NS = (int)L.stages, BR = S.band_rows;
    const uint32_t row_bytes = 3u * (uint32_t)w;
    const size_t pitch = BULK ? (I.stride == row_bytes ? row_bytes : L.row_pitch) : I.stride;
    const int row_words = (int)L.row_words;

    auto issue = [&](int c) {            // thread 0: stage chunk c with TMA bulk copies
        const int s = c % NS, rows = min(RS, h - c * RS);
        unsigned char *dst = ring + (size_t)s * L.stage_bytes;
        const uint8_t *src = I.pixels + (size_t)c * RS * I.stride;
        mbar_expect_tx(&full[s], rows * row_bytes);
        if (I.stride == row_bytes) bulk_g2s(dst, src, rows * row_bytes, &full[s]);
        else for (int r = 0; r < rows; ++r) bulk_g2s(dst + (size_t)r * L.row_pitch, src + (size_t)r * I.stride, row_bytes, &full[s]);
    };
    auto load_rowtab = [&](int band) {   // all threads: per-row pass tables of `band` into its shared-memory slot
        const int y0 = band * BR, rows = min(BR, h - y0);
        if (rows <= 0) return;
        const uint4 *src = reinterpret_cast<const uint4 *>(S.rowtab + y0);
        uint4 *dst = reinterpret_cast<uint4 *>(rtab + (size_t)(band & 1) * BR);
        for (int i = tid; i < rows * 4; i += nt) dst[i] = __ldg(src + i);
    };
    if (BULK && tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nt / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    load_rowtab(0);
    __syncthreads();
    if (BULK && tid == 0) for (int c = 0; c < min(NS, nchunks); ++c) issue(c);

    float acc[4][3][CPT];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int c = 0; c < CPT; ++c) acc[s][j][c] = 0.0f;

    int c = 0;   // chunk counter over the whole image
    for (int band = 0; band < S.nbands; ++band) {
        const int band_lo = band * BR, band_hi = min(h, band_lo + BR);
        const RowEntry *rt = rtab + (size_t)(band & 1) * BR;
        int slot = 0;
        for (int y_lo = band_lo; y_lo < band_hi; y_lo += RS, ++c) {
            const int y_hi = min(band_hi, y_lo + RS);
            const int s_idx = c % NS;
            if (BULK) mbar_wait(&full[s_idx], (c / NS) & 1);
            const uint8_t *base = BULK ? ring + (size_t)s_idx * L.stage_bytes + 3 * (size_t)x0
                                       : I.pixels + (size_t)y_lo * I.stride + 3 * (size_t)x0;
            // One row: luma of this thread's pixels, then the two lowest active outputs of each of the four passes.
            auto row_body = [&](int y, const RowEntry &e, float (&v)[CPT]) {
                const ulonglong2 wa = *reinterpret_cast<const ulonglong2 *>(&e.w01[0][0]);   // passes 0, 1 (broadcast LDS.128)
                const ulonglong2 wb = *reinterpret_cast<const ulonglong2 *>(&e.w01[2][0]);   // passes 2, 3
                const uint64_t w01[4] = {wa.x, wa.y, wb.x, wb.y};
                const uint8_t *px = base + (size_t)(y - y_lo) * pitch;
                if (CPT % 4 == 0 && (BULK || (I.aligned4 && w % CPT == 0))) {   // bulk staging implies w % 16 == 0
                    const uint32_t *p32 = reinterpret_cast<const uint32_t *>(px);
#pragma unroll
                    for (int g = 0; g < CPT / 4; ++g)   // 4 pixels = 12 bytes = 3 words
                        luma4(p32[3 * g], p32[3 * g + 1], p32[3 * g + 2], v[(4 * g) % CPT], v[(4 * g + 1) % CPT], v[(4 * g + 2) % CPT],
                              v[(4 * g + 3) % CPT]);
                } else if (CPT == 2 && (BULK || I.aligned4)) {   // cpt 2 is chosen for even widths only: 2 pixels = 6 bytes at an even offset
                    const uint16_t *p16 = reinterpret_cast<const uint16_t *>(px);
                    const uint32_t a = (uint32_t)p16[0] | (uint32_t)p16[1] << 16, b = p16[2];   // a = R0 G0 B0 R1, b = G1 B1
                    v[0] = luma_to_f32(__dp2a_hi(UCFP_W16(722, 0), a, __dp2a_lo(UCFP_W16(2126, 7152), a, 0u)));
                    v[1 % CPT] = luma_to_f32(__dp2a_lo(UCFP_W16(7152, 722), b, __dp2a_hi(UCFP_W16(0, 2126), a, 0u)));
                } else {
#pragma unroll
                    for (int cc = 0; cc < CPT; ++cc) {
                        const int xc = min(cc, w - 1 - x0);   // a ragged last group repeats its last column
                        v[cc] = luma_to_f32(2126u * px[3 * xc] + 7152u * px[3 * xc + 1] + 722u * px[3 * xc + 2]);
                    }
                }
#pragma unroll
                for (int s = 0; s < 4; ++s) {
#pragma unroll
                    for (int cc = 0; cc < CPT; ++cc) {
                        float p0, p1;
                        mul2(v[cc], w01[s], p0, p1);
                        acc[s][0][cc] = acc[s][0][cc] + p0;
                        acc[s][1][cc] = acc[s][1][cc] + p1;
                    }
                }
            };
            int y = y_lo;
            while (y < y_hi) {
                if constexpr (CPT >= 8) {
                    // a run of rows on which no output completes and no third output is active: nothing but arithmetic.  (With
                    // <= 4 columns per thread one loop body is faster -- 2.19 -> 2.29 M img/s at 256x256: the second copy of the
                    // row arithmetic costs more in register moves at the joins than the skipped flag test saves.)
                    const int run = min((int)rt[y - band_lo].run, y_hi - y);
                    for (int i = 0; i < run; ++i, ++y) {
                        float v[CPT];
                        row_body(y, rt[y - band_lo], v);
                    }
                    if (y >= y_hi) break;
                }
                const RowEntry &e = rt[y - band_lo];
                float v[CPT];
                row_body(y, e, v);
                const uint32_t flags = e.flags;   // uniform: a third active output, or outputs completed by this row
                if (flags) {
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const uint32_t fl = (flags >> (8 * s)) & 255u;
                        if (fl & 1u) {
                            const float w2 = e.w2[s];
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc) acc[s][2][cc] = acc[s][2][cc] + v[cc] * w2;
                        }
                        auto finish = [&]() {   // the lowest active output of this pass is complete
                            float *dst = rowbuf + (size_t)slot * row_words;
#pragma unroll
                            for (int cc = 0; cc < CPT; ++cc) {
                                if (owner && x0 + cc < w) dst[skew(x0 + cc)] = acc[s][0][cc];
                                acc[s][0][cc] = acc[s][1][cc]; acc[s][1][cc] = acc[s][2][cc]; acc[s][2][cc] = 0.0f;
                            }
                            slot++;
                        };
                        const uint32_t nf = fl >> 1;   // <= 3: a pass has at most three active outputs (build_shape checks)
                        if constexpr (CPT >= 8) {   // 96 accumulators: three inlined copies of the rotation spill
                            for (uint32_t f = 0; f < nf; ++f) finish();
                        } else if (nf) {
                            finish();
                            if (nf > 1) { finish(); if (nf > 2) finish(); }
                        }
                    }
                }
                ++y;
            }
            if (BULK) {   // hand the stage back: every warp arrives, thread 0 refills once all have
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s_idx]);
                if (tid == 0 && c + NS < nchunks) { mbar_wait(&empty[s_idx], (c / NS) & 1); issue(c + NS); }
            }
        }
        __syncthreads();   // finished rows of this band are visible to everyone
        load_rowtab(band + 1);
        {   // horizontal passes over the rows finished in this band: one flat item list, lanes on neighbouring outputs of a row
            const int i_lo = S.hitem_off[band], i_hi = S.hitem_off[band + 1];
            uint8_t *grids = reinterpret_cast<uint8_t *>(&G);
            int i = i_lo + tid;
            uint32_t d = i < i_hi ? __ldg(S.hitems + i) : 0u;
            while (i < i_hi) {   // the next item's word is in flight while this one's taps are summed
                const int in = i + nt;
                const uint32_t dn = in < i_hi ? __ldg(S.hitems + in) : 0u;
                const Taps t = S.hout[d >> 24];
                grids[d & 0xFFFFu] = hsample<true>(rowbuf + (size_t)((d >> 16) & 255u) * row_words, t, S.wts);
                d = dn; i = in;
            }
        }
        __syncthreads();
    }
    hash_regions(G, H, algo_mask, out + (size_t)I.out_index * 51);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Host driver
// ---------------------------------------------------------------------------------------------
int image_hash_batch(ucfp_lane *ctx, const ucfp_image_desc *descs, size_t n, uint32_t algo_mask, ucfp_image_hashes *out_dev,
                     int32_t *status, int pixels_mem) {
    cudaStream_t st = ctx->stream;
    UCFP_CUDA_TRY(cudaMemsetAsync(out_dev, 0, sizeof(ucfp_image_hashes) * n, st));
    static_assert(sizeof(ucfp_image_hashes) == 51 * 8, "bundle layout");

    // ---- validate, stage host pixels, group by shape
    struct Item { size_t idx; const uint8_t *dev_pixels; uint64_t stride; };
    std::map<uint64_t, std::vector<Item>> groups;
    size_t stage_bytes = 0;
    std::vector<size_t> stage_off(n, SIZE_MAX);
    std::vector<uint64_t> pitch(n, 0);
    // A pointer query per image (cudaPointerGetAttributes) was a third of the wall time of a 16 K-image device batch.  The
    // uniform entry point knows where its one buffer lives (pixels_mem); otherwise an image that starts where the previous
    // device image ended is taken to be in the same allocation.
    const uint8_t *dev_next = nullptr;
    auto is_device = [&](const ucfp_image_desc &d) {
        bool dev;
        if (pixels_mem >= 0) dev = pixels_mem == 1;
        else if (d.pixels == dev_next) dev = true;
        else dev = classify(d.pixels) == Mem::Device;
        dev_next = dev ? d.pixels + d.stride * (uint64_t)d.height : nullptr;
        return dev;
    };
    for (size_t i = 0; i < n; ++i) {
        const ucfp_image_desc &d = descs[i];
        status[i] = UCFP_OK;
        if (!d.pixels || d.width < 4 || d.height < 4 || d.stride < 3ull * d.width || d.width > 65535 || d.height > 65535) {
            status[i] = UCFP_E_INVALID;
            continue;
        }
        if (!is_device(d)) {
            pitch[i] = (3ull * d.width + 15) & ~15ull;
            stage_off[i] = stage_bytes;
            stage_bytes += pitch[i] * d.height;
        }
    }
    const size_t kStageLimit = size_t(3) << 30;  // larger host batches are hashed in slices
    if (stage_bytes > kStageLimit && n > 1) {
        size_t half = n / 2;
        UCFP_TRY(image_hash_batch(ctx, descs, half, algo_mask, out_dev, status, pixels_mem));
        UCFP_CUDA_TRY(cudaStreamSynchronize(st));
        return image_hash_batch(ctx, descs + half, n - half, algo_mask, out_dev + half, status + half, pixels_mem);
    }
    if (stage_bytes) UCFP_TRY(ctx->img_stage_dev.reserve(stage_bytes));
    for (size_t i = 0; i < n; ++i) {
        if (status[i] != UCFP_OK) continue;
        const ucfp_image_desc &d = descs[i];
        Item it{i, d.pixels, d.stride};
        if (stage_off[i] != SIZE_MAX) {
            uint8_t *dst = ctx->img_stage_dev.as<uint8_t>() + stage_off[i];
            // runs of tightly packed, contiguous host images of one shape go up as a single 2-D copy
            size_t run = 1;
            if (d.stride == 3ull * d.width) {
                while (i + run < n && status[i + run] == UCFP_OK && stage_off[i + run] != SIZE_MAX &&
                       descs[i + run].width == d.width && descs[i + run].height == d.height &&
                       descs[i + run].stride == d.stride &&
                       descs[i + run].pixels == d.pixels + run * d.stride * d.height)
                    run++;
            }
            UCFP_CUDA_TRY(cudaMemcpy2DAsync(dst, pitch[i], d.pixels, d.stride, 3ull * d.width, (size_t)d.height * run,
                                            cudaMemcpyHostToDevice, st));
            for (size_t j = 0; j < run; ++j) {
                Item jt{i + j, dst + j * pitch[i] * d.height, pitch[i]};
                groups[(uint64_t)d.width << 32 | d.height].push_back(jt);
            }
            i += run - 1;
            continue;
        }
        groups[(uint64_t)d.width << 32 | d.height].push_back(it);
    }

    // ---- tap tables per shape: looked up (or built) under the context's image mutex and pinned by a use count, so that
    // the bounded cache below can evict shapes nobody is hashing right now without pulling tables from under a running kernel
    struct GroupRef { uint64_t key; ShapeTables T; std::vector<Item> *items; size_t desc_off; };
    std::vector<GroupRef> refs;
    refs.reserve(groups.size());
    ucfp_ctx *own = ctx->owner;
    auto release_refs = [&]() {
        std::lock_guard<std::mutex> lk(own->image_mu);
        ShapeCache &cache = *static_cast<ShapeCache *>(own->image_cache);
        for (GroupRef &g : refs) { auto f = cache.m.find(g.key); if (f != cache.m.end() && f->second.in_use) f->second.in_use--; }
    };
    size_t n_items = 0;
    {
        std::lock_guard<std::mutex> lk(own->image_mu);
        if (!own->image_cache) own->image_cache = new (std::nothrow) ShapeCache();
        UCFP_REQUIRE(own->image_cache != nullptr, UCFP_E_OOM, "out of host memory");
        ShapeCache &cache = *static_cast<ShapeCache *>(own->image_cache);
        for (auto &kv : groups) {
            const int w = (int)(kv.first >> 32), h = (int)(kv.first & 0xffffffffu);
            auto found = cache.m.find(kv.first);
            if (found == cache.m.end()) {
                ShapeTables stb;
                int rc = build_shape(ctx, w, h, stb);
                if (rc != UCFP_OK) { for (GroupRef &g : refs) cache.m[g.key].in_use--; return rc; }
                found = cache.m.emplace(kv.first, stb).first;
                cache.bytes += stb.blob_bytes;
                // bounded: least recently used shapes that no call is using go first (a service hashing arbitrary upload
                // sizes would otherwise grow this map, and the device blobs behind it, without limit)
                while ((cache.m.size() > kShapeCacheMaxEntries || cache.bytes > kShapeCacheMaxBytes)) {
                    auto victim = cache.m.end();
                    for (auto it2 = cache.m.begin(); it2 != cache.m.end(); ++it2)
                        if (it2 != found && it2->second.in_use == 0 && (victim == cache.m.end() || it2->second.last_use < victim->second.last_use)) victim = it2;
                    if (victim == cache.m.end()) break;
                    if (victim->second.blob) cudaFree(victim->second.blob);
                    cache.bytes -= victim->second.blob_bytes;
                    cache.m.erase(victim);
                }
            }
            found->second.in_use++;
            found->second.last_use = ++cache.tick;
            refs.push_back(GroupRef{kv.first, found->second, &kv.second, n_items});
            n_items += kv.second.size();
        }
    }
    // ---- one descriptor array for the whole batch (pinned staging owned by the lane, one upload), groups launch back to back
    int rc = UCFP_OK;
    do {
        if ((rc = ctx->pin_a.reserve(sizeof(ImgDev) * n_items)) != UCFP_OK) break;
        if ((rc = ctx->img_desc_dev.reserve(sizeof(ImgDev) * n_items)) != UCFP_OK) break;
        ImgDev *hostdesc = ctx->pin_a.as<ImgDev>();
        for (GroupRef &g : refs) {
            const int w = (int)(g.key >> 32);
            for (size_t j = 0; j < g.items->size(); ++j) {
                const Item &itm = (*g.items)[j];
                const uintptr_t base = (uintptr_t)itm.dev_pixels;
                uint32_t al = (base % 4 == 0 && itm.stride % 4 == 0) ? 1u : 0u;
                if (base % 16 == 0 && itm.stride % 16 == 0 && (3 * (size_t)w) % 16 == 0) al = 2u;   // TMA bulk staging
                hostdesc[g.desc_off + j] = ImgDev{itm.dev_pixels, itm.stride, (uint32_t)itm.idx, al};
            }
        }
        if (cudaMemcpyAsync(ctx->img_desc_dev.ptr, hostdesc, sizeof(ImgDev) * n_items, cudaMemcpyHostToDevice, st) != cudaSuccess) {
            set_error("descriptor upload failed: %s", cudaGetErrorString(cudaGetLastError())); rc = UCFP_E_CUDA; break;
        }
        uint64_t *out_words = reinterpret_cast<uint64_t *>(out_dev);
        for (GroupRef &g : refs) {
            const int w = (int)(g.key >> 32), h = (int)(g.key & 0xffffffffu);
            ShapeTables &T = g.T;
            const size_t cnt = g.items->size();
            const ImgDev *descs_dev = ctx->img_desc_dev.as<ImgDev>() + g.desc_off;
            const double units = (3.0 * w * h + 408.0) * (double)cnt;
            if (T.streamable) {
                const size_t smem = T.stream_smem;
                ProfScope ps(ctx, UCFP_PROF_IMAGE_HASH, units);
                auto launch = [&](auto kern) { kern<<<(unsigned)cnt, T.threads, smem, st>>>(T.dev, T.lay, descs_dev, algo_mask, out_words); };
                bool all_bulk = true;
                for (size_t j = 0; j < cnt; ++j) all_bulk = all_bulk && hostdesc[g.desc_off + j].aligned4 == 2;
                if (T.cpt == 1) { if (all_bulk) launch(image_stream_kernel<1, 512, true>); else launch(image_stream_kernel<1, 512, false>); }
                else if (T.cpt == 2) { if (all_bulk) launch(image_stream_kernel<2, 256, true>); else launch(image_stream_kernel<2, 256, false>); }
                else if (T.cpt == 4) { if (all_bulk) launch(image_stream_kernel<4, 256, true>); else launch(image_stream_kernel<4, 256, false>); }
                else { if (all_bulk) launch(image_stream_kernel<8, 512, true>); else launch(image_stream_kernel<8, 512, false>); }
            } else {
                size_t per = (((size_t)w * h + 15) & ~size_t(15)) + (size_t)5 * kVOuts * w * 4 + 256;
                size_t grid = cnt;
                size_t budget = size_t(1) << 30;
                if (grid * per > budget) grid = budget / per ? budget / per : 1;
                if (grid > (size_t)ctx->sm_count * 4) grid = (size_t)ctx->sm_count * 4;
                // generic-kernel groups of one batch share this scratch: they run one after the other on the lane's stream
                if ((rc = ctx->img_tables_dev.reserve(grid * per)) != UCFP_OK) break;
                ProfScope ps(ctx, UCFP_PROF_IMAGE_HASH, units);
                image_generic_kernel<<<(unsigned)grid, 256, kGridsBytes + sizeof(HashScratch), st>>>(T.dev, descs_dev, (uint32_t)cnt, algo_mask,
                                                                                                    ctx->img_tables_dev.as<uint8_t>(), per, out_words);
            }
            count_launch(ctx);
            if ((rc = check_launch("image hash")) != UCFP_OK) break;
        }
    } while (false);
    // one synchronisation per batch: the pinned descriptors may be rewritten and the shapes evicted after it
    cudaError_t se = cudaStreamSynchronize(st);
    release_refs();
    if (rc == UCFP_OK && se != cudaSuccess) { set_error("image hash failed: %s", cudaGetErrorString(se)); rc = UCFP_E_CUDA; }
    return rc;
}

int image_device_init(ucfp_ctx *) {
    const int optin = 200 * 1024;
    UCFP_CUDA_TRY(cudaFuncSetAttribute(image_stream_kernel<1, 512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(image_stream_kernel<1, 512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(image_stream_kernel<2, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(image_stream_kernel<2, 256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(image_stream_kernel<4, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(image_stream_kernel<4, 256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(image_stream_kernel<8, 512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(image_stream_kernel<8, 512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    UCFP_CUDA_TRY(cudaFuncSetAttribute(image_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kGridsBytes + sizeof(HashScratch))));
    return UCFP_OK;
}

void image_cache_destroy(ucfp_ctx *ctx) {
    if (!ctx->image_cache) return;
    ShapeCache *cache = static_cast<ShapeCache *>(ctx->image_cache);
    for (auto &kv : cache->m) if (kv.second.blob) cudaFree(kv.second.blob);
    delete cache;
    ctx->image_cache = nullptr;
}

}  // namespace ucfp
