// jpeg.cu -- batch ingest from ENCODED bytes: JPEG bitstreams are decoded on the device by nvJPEG
// (GPU-assisted Huffman + IDCT + colour conversion) straight into the RGB8 layout the hashing kernels stage, and hashed by
// the same call -- decoded pixels never cross PCIe.  The reference decodes every upload on a tokio worker thread with
// the `image` crate and commits one redb transaction per image (src/server/handlers.rs:232-302,
// src/modality/image.rs:68-70); its only benchmark includes that decode (benches/end_to_end.rs:40-53).
// PNG / WebP / GIF / BMP stay on the host (decode there, then ucfp_image_hash_batch); so does any JPEG nvJPEG refuses
// (per-image status UCFP_E_UNSUPPORTED tells the host which ones).
// nvJPEG is library code (it ships with the CUDA toolkit) and is resolved with dlopen at first use.
#include <dlfcn.h>
#include <nvjpeg.h>

#include "api_util.cuh"

namespace {

struct NvjpegApi {
    void *lib = nullptr;
    nvjpegStatus_t (*CreateEx)(nvjpegBackend_t, nvjpegDevAllocator_t *, nvjpegPinnedAllocator_t *, unsigned int, nvjpegHandle_t *) = nullptr;
    nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
    nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t *) = nullptr;
    nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
    nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char *, size_t, int *, nvjpegChromaSubsampling_t *, int *, int *) = nullptr;
    nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char *, size_t, nvjpegOutputFormat_t, nvjpegImage_t *, cudaStream_t) = nullptr;
    nvjpegStatus_t (*DecodeBatchedInitialize)(nvjpegHandle_t, nvjpegJpegState_t, int, int, nvjpegOutputFormat_t) = nullptr;
    nvjpegStatus_t (*DecodeBatched)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char *const *, const size_t *, nvjpegImage_t *, cudaStream_t) = nullptr;
};
std::mutex g_mu;
NvjpegApi g_api;

int nvjpeg_load() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_api.lib) return UCFP_OK;
    void *h = nullptr;
    for (const char *name : {"libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so.12", "/usr/local/cuda/targets/x86_64-linux/lib/libnvjpeg.so.12", "libnvjpeg.so"})
        if ((h = dlopen(name, RTLD_NOW | RTLD_LOCAL))) break;
    UCFP_REQUIRE(h != nullptr, UCFP_E_UNSUPPORTED, "nvJPEG is not available: %s", dlerror());
    NvjpegApi a;
    a.lib = h;
#define UCFP_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name))
    UCFP_SYM(CreateEx, "nvjpegCreateEx"); UCFP_SYM(Destroy, "nvjpegDestroy"); UCFP_SYM(JpegStateCreate, "nvjpegJpegStateCreate");
    UCFP_SYM(JpegStateDestroy, "nvjpegJpegStateDestroy"); UCFP_SYM(GetImageInfo, "nvjpegGetImageInfo"); UCFP_SYM(Decode, "nvjpegDecode");
    UCFP_SYM(DecodeBatchedInitialize, "nvjpegDecodeBatchedInitialize"); UCFP_SYM(DecodeBatched, "nvjpegDecodeBatched");
#undef UCFP_SYM
    UCFP_REQUIRE(a.CreateEx && a.Destroy && a.JpegStateCreate && a.JpegStateDestroy && a.GetImageInfo && a.Decode && a.DecodeBatchedInitialize && a.DecodeBatched,
                 UCFP_E_UNSUPPORTED, "libnvjpeg lacks a required symbol");
    g_api = a;
    return UCFP_OK;
}

// per-context decoder: one nvJPEG handle + state, one decode batch at a time (ucfp_ctx::jpeg_mu)
struct JpegDecoder {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    int backend = 0;
    ucfp::DevBuf pixels;
};

int decoder_of(ucfp_ctx *ctx, JpegDecoder **out) {
    if (!ctx->jpeg) {
        UCFP_TRY(nvjpeg_load());
        JpegDecoder *d = new JpegDecoder();
        // GPU-assisted Huffman decode for batches; the default (hybrid, host Huffman) backend when the GPU one is refused
        for (nvjpegBackend_t b : {NVJPEG_BACKEND_GPU_HYBRID, NVJPEG_BACKEND_DEFAULT}) {
            if (g_api.CreateEx(b, nullptr, nullptr, 0, &d->handle) == NVJPEG_STATUS_SUCCESS) { d->backend = (int)b; break; }
            d->handle = nullptr;
        }
        if (!d->handle || g_api.JpegStateCreate(d->handle, &d->state) != NVJPEG_STATUS_SUCCESS) {
            if (d->handle) g_api.Destroy(d->handle);
            delete d;
            ucfp::set_error("nvJPEG could not create a decoder on device %d", ctx->device);
            return UCFP_E_CUDA;
        }
        ctx->jpeg = d;
    }
    *out = static_cast<JpegDecoder *>(ctx->jpeg);
    return UCFP_OK;
}

}  // namespace

namespace ucfp {
void jpeg_destroy(ucfp_ctx *ctx) {
    if (!ctx->jpeg) return;
    JpegDecoder *d = static_cast<JpegDecoder *>(ctx->jpeg);
    if (d->state) g_api.JpegStateDestroy(d->state);
    if (d->handle) g_api.Destroy(d->handle);
    d->pixels.release();
    delete d;
    ctx->jpeg = nullptr;
}
}  // namespace ucfp

using namespace ucfp;

extern "C" {

int ucfp_image_hash_jpeg_batch(ucfp_ctx *ctx, const uint8_t *const *jpegs, const size_t *lengths, size_t n, uint32_t algo_mask,
                               ucfp_image_hashes *out, int32_t *status, uint32_t *dims_out, uint8_t *pixels_out, size_t pixels_capacity) {
    UCFP_API_BEGIN
    UCFP_REQUIRE(ctx != nullptr, UCFP_E_INVALID, "null context");
    UCFP_LEASE(ctx);
    if (n == 0) return UCFP_OK;
    UCFP_REQUIRE(jpegs && lengths && out && status, UCFP_E_INVALID, "NULL bitstreams, lengths, output or status");
    UCFP_REQUIRE(classify(status) != Mem::Device, UCFP_E_INVALID, "status must be host memory");
    UCFP_REQUIRE((algo_mask & ~UCFP_ALGO_MULTI) == 0 && algo_mask != 0, UCFP_E_INVALID, "bad algo_mask 0x%x", algo_mask);
    std::lock_guard<std::mutex> jl(ctx->jpeg_mu);
    JpegDecoder *dec = nullptr;
    UCFP_TRY(decoder_of(ctx, &dec));
    cudaStream_t st = lane->stream;

    // ---- 1. headers: dimensions, per-image verdict, layout of the decoded batch in one device buffer
    std::vector<uint32_t> w(n, 0), h(n, 0);
    std::vector<size_t> off(n, 0), pitch(n, 0);
    std::vector<size_t> ok;
    size_t total = 0;
    for (size_t i = 0; i < n; ++i) {
        status[i] = UCFP_E_INVALID;
        if (dims_out) { dims_out[2 * i] = 0; dims_out[2 * i + 1] = 0; }
        if (!jpegs[i] || lengths[i] < 4 || jpegs[i][0] != 0xFF || jpegs[i][1] != 0xD8) { status[i] = jpegs[i] && lengths[i] ? UCFP_E_UNSUPPORTED : UCFP_E_INVALID; continue; }   // not a JPEG: host decoders
        int ncomp = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
        nvjpegChromaSubsampling_t sub;
        const nvjpegStatus_t r = g_api.GetImageInfo(dec->handle, jpegs[i], lengths[i], &ncomp, &sub, ws, hs);
        if (r != NVJPEG_STATUS_SUCCESS) { status[i] = r == NVJPEG_STATUS_BAD_JPEG || r == NVJPEG_STATUS_INCOMPLETE_BITSTREAM ? UCFP_E_INVALID : UCFP_E_UNSUPPORTED; continue; }
        if ((ncomp != 1 && ncomp != 3) || ws[0] < 4 || hs[0] < 4 || ws[0] > 65535 || hs[0] > 65535) { status[i] = UCFP_E_UNSUPPORTED; continue; }
        w[i] = (uint32_t)ws[0]; h[i] = (uint32_t)hs[0];
        if (dims_out) { dims_out[2 * i] = w[i]; dims_out[2 * i + 1] = h[i]; }
        pitch[i] = (3 * (size_t)w[i] + 15) & ~size_t(15);   // 16-byte rows: the hashing kernel's TMA bulk staging applies when 3 w is a multiple of 16
        off[i] = total;
        total += (pitch[i] * h[i] + 255) & ~size_t(255);
        status[i] = UCFP_OK;
        ok.push_back(i);
    }
    // ---- 2. decode on the device
    if (!ok.empty()) {
        UCFP_TRY(dec->pixels.reserve(total));
        uint8_t *base = dec->pixels.as<uint8_t>();
        std::vector<const unsigned char *> data(ok.size());
        std::vector<size_t> lens(ok.size());
        std::vector<nvjpegImage_t> dst(ok.size());
        for (size_t j = 0; j < ok.size(); ++j) {
            const size_t i = ok[j];
            data[j] = jpegs[i]; lens[j] = lengths[i];
            memset(&dst[j], 0, sizeof(nvjpegImage_t));
            dst[j].channel[0] = base + off[i];
            dst[j].pitch[0] = pitch[i];
        }
        bool batched = g_api.DecodeBatchedInitialize(dec->handle, dec->state, (int)ok.size(), 1, NVJPEG_OUTPUT_RGBI) == NVJPEG_STATUS_SUCCESS &&
                       g_api.DecodeBatched(dec->handle, dec->state, data.data(), lens.data(), dst.data(), st) == NVJPEG_STATUS_SUCCESS;
        if (!batched) {   // one image the batched decoder refuses fails the whole batch call: decode one by one, mark the refused ones
            cudaGetLastError();
            for (size_t j = 0; j < ok.size(); ++j) {
                const nvjpegStatus_t r = g_api.Decode(dec->handle, dec->state, data[j], lens[j], NVJPEG_OUTPUT_RGBI, &dst[j], st);
                if (r != NVJPEG_STATUS_SUCCESS) status[ok[j]] = r == NVJPEG_STATUS_BAD_JPEG || r == NVJPEG_STATUS_INCOMPLETE_BITSTREAM ? UCFP_E_INVALID : UCFP_E_UNSUPPORTED;
            }
        }
        // the bitstreams are host memory the caller may free on return, and nvJPEG reads them asynchronously
        UCFP_CUDA_TRY(cudaStreamSynchronize(st));
    }
    // ---- 3. hash the decoded pixels where they are
    std::vector<ucfp_image_desc> descs(n);
    for (size_t i = 0; i < n; ++i)
        descs[i] = status[i] == UCFP_OK ? ucfp_image_desc{dec->pixels.as<uint8_t>() + off[i], w[i], h[i], pitch[i]} : ucfp_image_desc{nullptr, 0, 0, 0};
    void *out_dev = nullptr;
    bool out_host = false;
    UCFP_TRY(stage_out(lane->img_out_dev, out, sizeof(ucfp_image_hashes) * n, &out_dev, &out_host));
    std::vector<int32_t> hst(n, 0);
    UCFP_TRY(image_hash_batch(lane, descs.data(), n, algo_mask, static_cast<ucfp_image_hashes *>(out_dev), hst.data(), /*pixels_mem=*/1));   // the decoder's own device buffer
    for (size_t i = 0; i < n; ++i)
        if (status[i] == UCFP_OK) status[i] = hst[i];
    if (out_host) UCFP_TRY(copy_back(lane, out, out_dev, sizeof(ucfp_image_hashes) * n));
    // ---- 4. optional: the decoded pixels, tightly packed (3 w bytes per row), image after image -- parity checks and hosts that
    // want to keep thumbnails; images that failed contribute nothing
    if (pixels_out) {
        size_t at = 0;
        const bool dev_dst = classify(pixels_out) == Mem::Device;
        for (size_t i = 0; i < n; ++i) {
            if (status[i] != UCFP_OK) continue;
            const size_t bytes = 3 * (size_t)w[i] * h[i];
            UCFP_REQUIRE(at + bytes <= pixels_capacity, UCFP_E_CAPACITY, "pixels_out holds %zu bytes, the decoded batch needs more", pixels_capacity);
            UCFP_CUDA_TRY(cudaMemcpy2DAsync(pixels_out + at, 3 * (size_t)w[i], dec->pixels.as<uint8_t>() + off[i], pitch[i], 3 * (size_t)w[i], h[i],
                                            dev_dst ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
            at += bytes;
        }
    }
    // the decoder's pixel buffer is reused by the next batch: always complete before the decoder lock is dropped
    UCFP_CUDA_TRY(cudaStreamSynchronize(st));
    return UCFP_OK;
    UCFP_API_END
}

}  // extern "C"
