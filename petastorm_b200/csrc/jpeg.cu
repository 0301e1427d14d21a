// K9: JPEG through nvJPEG (NVIDIA library, batched API) -> interleaved RGB u8 [n, H, W, 3] in HBM.
// Replaces cv2.imdecode (libjpeg-turbo) of CompressedImageCodec('jpeg').decode -- petastorm/codecs.py:102-116.
// nvJPEG is dlopen'ed so libpst_b200.so loads on hosts without it; a missing library is a hard error at call time
// (no CPU fallback).  Results differ from libjpeg-turbo by small LSB amounts (different IDCT / upsampling), which is the
// tolerance the north star allows for this codec.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvjpeg.h>

#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/pst_b200.h"
#include "host_state.h"

namespace {

struct NvjpegApi {
    void *lib = nullptr;
    decltype(&nvjpegCreateEx) CreateEx = nullptr;
    decltype(&nvjpegCreateSimple) CreateSimple = nullptr;
    decltype(&nvjpegDestroy) Destroy = nullptr;
    decltype(&nvjpegJpegStateCreate) JpegStateCreate = nullptr;
    decltype(&nvjpegJpegStateDestroy) JpegStateDestroy = nullptr;
    decltype(&nvjpegDecodeBatchedInitialize) DecodeBatchedInitialize = nullptr;
    decltype(&nvjpegDecodeBatched) DecodeBatched = nullptr;
    decltype(&nvjpegGetImageInfo) GetImageInfo = nullptr;
    bool ok = false;
    std::string why;
};

NvjpegApi &api() {
    static NvjpegApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so.12", "libnvjpeg.so"};
        for (const char *nm : names) {
            a.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (a.lib) break;
        }
        if (!a.lib) {
            a.why = std::string("cannot dlopen libnvjpeg: ") + (dlerror() ? dlerror() : "?");
            return;
        }
#define LOAD(field, sym)                                                    \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.lib, sym));       \
    if (!a.field) {                                                         \
        a.why = std::string("libnvjpeg lacks symbol ") + sym;               \
        return;                                                             \
    }
        LOAD(CreateEx, "nvjpegCreateEx")
        LOAD(CreateSimple, "nvjpegCreateSimple")
        LOAD(Destroy, "nvjpegDestroy")
        LOAD(JpegStateCreate, "nvjpegJpegStateCreate")
        LOAD(JpegStateDestroy, "nvjpegJpegStateDestroy")
        LOAD(DecodeBatchedInitialize, "nvjpegDecodeBatchedInitialize")
        LOAD(DecodeBatched, "nvjpegDecodeBatched")
        LOAD(GetImageInfo, "nvjpegGetImageInfo")
#undef LOAD
        a.ok = true;
    });
    return a;
}

struct JpegState {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;
    int batch = 0;
    int backend = -1;
    bool tried = false;
    cudaEvent_t last = nullptr;    // end of the previous batch decoded with this state (its scratch buffers are reused)
    std::mutex mu;
};

JpegState &jstate() {
    static JpegState s;
    return s;
}
// handles whose batched decode reads the bitstreams from DEVICE memory (no blob D2H); two of them, so that the
// resolver threads of a pool can have two batches in nvJPEG at the same time (its host-side work per batch is serial)
constexpr int kDeviceStates = 2;
JpegState *device_states() {
    static JpegState s[kDeviceStates];
    return s;
}

void fill_outputs(std::vector<nvjpegImage_t> &outs, uint64_t dst, int64_t n, int height, int width) {
    uint8_t *d = (uint8_t *)dst;
    const size_t img_bytes = (size_t)height * width * 3;
    for (int64_t i = 0; i < n; i++) {
        for (int k = 0; k < NVJPEG_MAX_COMPONENT; k++) {
            outs[i].channel[k] = nullptr;
            outs[i].pitch[k] = 0;
        }
        outs[i].channel[0] = d + (size_t)i * img_bytes;
        outs[i].pitch[0] = (size_t)width * 3;
    }
}

}  // namespace

extern "C" {

int pst_jpeg_available(void) { return api().ok ? 1 : 0; }

int pst_jpeg_backend(void) { return jstate().backend; }

// creates the device-bitstream handle of one state on first use: the hardware JPEG engines first, then GPU-assisted
// Huffman; returns its nvjpegBackend_t or -1 (caller holds js.mu)
static int ensure_device_state(NvjpegApi &a, JpegState &js) {
    if (!js.tried) {
        js.tried = true;
        const char *skip_hw = getenv("PST_JPEG_NO_HW");
        const nvjpegBackend_t order[] = {NVJPEG_BACKEND_HARDWARE_DEVICE, NVJPEG_BACKEND_GPU_HYBRID_DEVICE};
        for (nvjpegBackend_t b : order) {
            if (b == NVJPEG_BACKEND_HARDWARE_DEVICE && skip_hw && skip_hw[0] == '1') continue;
            if (a.CreateEx(b, nullptr, nullptr, 0, &js.handle) == NVJPEG_STATUS_SUCCESS) {
                if (a.JpegStateCreate(js.handle, &js.state) == NVJPEG_STATUS_SUCCESS) {
                    js.backend = (int)b;
                    break;
                }
                a.Destroy(js.handle);
            }
            js.handle = nullptr;
        }
    }
    return js.backend;
}

int pst_jpeg_device_backend(void) {
    NvjpegApi &a = api();
    if (!a.ok) return -1;
    JpegState &js = device_states()[0];
    std::lock_guard<std::mutex> g(js.mu);
    return ensure_device_state(a, js);
}

int pst_jpeg_batch_device(pst_ctx *c, uint64_t base, const int64_t *host_offs, const int32_t *host_lens, int64_t n,
                          int height, int width, uint64_t dst, uint64_t stream) {
    (void)c;
    try {
        NvjpegApi &a = api();
        if (!a.ok) throw std::runtime_error("nvJPEG unavailable: " + a.why);
        if (n <= 0) return 0;
        // a free state, else wait for the first one
        JpegState *states = device_states();
        std::unique_lock<std::mutex> g;
        JpegState *jsp = nullptr;
        for (int k = 0; k < kDeviceStates && !jsp; k++) {
            std::unique_lock<std::mutex> t(states[k].mu, std::try_to_lock);
            if (t.owns_lock()) {
                g = std::move(t);
                jsp = &states[k];
            }
        }
        if (!jsp) {
            g = std::unique_lock<std::mutex>(states[0].mu);
            jsp = &states[0];
        }
        JpegState &js = *jsp;
        if (ensure_device_state(a, js) < 0)
            throw std::runtime_error("this nvJPEG has no backend that decodes device-resident bitstreams");
        if (js.batch != (int)n) {
            nvjpegStatus_t st = a.DecodeBatchedInitialize(js.handle, js.state, (int)n, 1, NVJPEG_OUTPUT_RGBI);
            if (st != NVJPEG_STATUS_SUCCESS)
                throw std::runtime_error("nvjpegDecodeBatchedInitialize (device bitstreams) failed " + std::to_string((int)st));
            js.batch = (int)n;
        }
        std::vector<const unsigned char *> ptrs((size_t)n);
        std::vector<size_t> lens((size_t)n);
        for (int64_t i = 0; i < n; i++) {
            ptrs[i] = (const unsigned char *)base + host_offs[i];
            lens[i] = (size_t)host_lens[i];
        }
        std::vector<nvjpegImage_t> outs((size_t)n);
        fill_outputs(outs, dst, n, height, width);
        // batches of different row-groups run on different streams: order them on the state's own buffers
        if (!js.last) cudaEventCreateWithFlags(&js.last, cudaEventDisableTiming);
        else cudaStreamWaitEvent((cudaStream_t)stream, js.last, 0);
        nvjpegStatus_t st = a.DecodeBatched(js.handle, js.state, ptrs.data(), lens.data(), outs.data(), (cudaStream_t)stream);
        if (js.last) cudaEventRecord(js.last, (cudaStream_t)stream);
        if (st != NVJPEG_STATUS_SUCCESS) {
            js.batch = 0;
            throw std::runtime_error("nvjpegDecodeBatched (device bitstreams, backend " + std::to_string(js.backend) +
                                     ") failed with status " + std::to_string((int)st));
        }
        return 0;
    } catch (const std::exception &e) {
        pst::set_error(e.what());
        return 1;
    }
}

int pst_jpeg_batch(pst_ctx *c, const uint8_t *const *host_blobs, const size_t *host_lens, int64_t n, int height,
                   int width, uint64_t dst, uint64_t stream) {
    (void)c;
    try {
        NvjpegApi &a = api();
        if (!a.ok) throw std::runtime_error("nvJPEG unavailable: " + a.why);
        if (n <= 0) return 0;
        JpegState &js = jstate();
        std::lock_guard<std::mutex> g(js.mu);
        if (!js.handle) {
            // GPU-assisted Huffman first (batched decode is then entirely on the device), then the hybrid backend
            const nvjpegBackend_t order[] = {NVJPEG_BACKEND_GPU_HYBRID, NVJPEG_BACKEND_HYBRID, NVJPEG_BACKEND_DEFAULT};
            nvjpegStatus_t st = NVJPEG_STATUS_NOT_INITIALIZED;
            for (nvjpegBackend_t b : order) {
                st = a.CreateEx(b, nullptr, nullptr, 0, &js.handle);
                if (st == NVJPEG_STATUS_SUCCESS) {
                    js.backend = (int)b;
                    break;
                }
                js.handle = nullptr;
            }
            if (!js.handle) throw std::runtime_error("nvjpegCreateEx failed with status " + std::to_string((int)st));
            st = a.JpegStateCreate(js.handle, &js.state);
            if (st != NVJPEG_STATUS_SUCCESS) throw std::runtime_error("nvjpegJpegStateCreate failed " + std::to_string((int)st));
        }
        if (js.batch != (int)n) {
            nvjpegStatus_t st = a.DecodeBatchedInitialize(js.handle, js.state, (int)n, 1, NVJPEG_OUTPUT_RGBI);
            if (st != NVJPEG_STATUS_SUCCESS)
                throw std::runtime_error("nvjpegDecodeBatchedInitialize failed " + std::to_string((int)st));
            js.batch = (int)n;
        }
        // geometry check of every stream against the field shape (cheap header parse on the host)
        for (int64_t i = 0; i < n; i++) {
            int comps = 0;
            nvjpegChromaSubsampling_t ss;
            int ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
            nvjpegStatus_t st = a.GetImageInfo(js.handle, host_blobs[i], host_lens[i], &comps, &ss, ws, hs);
            if (st != NVJPEG_STATUS_SUCCESS)
                throw std::runtime_error("image " + std::to_string(i) + " is not a decodable JPEG (status " + std::to_string((int)st) + ")");
            if (ws[0] != width || hs[0] != height)
                throw std::runtime_error("JPEG " + std::to_string(i) + " is " + std::to_string(hs[0]) + "x" + std::to_string(ws[0]) +
                                         ", expected " + std::to_string(height) + "x" + std::to_string(width));
        }
        std::vector<nvjpegImage_t> outs((size_t)n);
        fill_outputs(outs, dst, n, height, width);
        nvjpegStatus_t st = a.DecodeBatched(js.handle, js.state, host_blobs, host_lens, outs.data(), (cudaStream_t)stream);
        if (st != NVJPEG_STATUS_SUCCESS) {
            js.batch = 0;  // the batch state must be re-initialised after a failure
            throw std::runtime_error("nvjpegDecodeBatched failed with status " + std::to_string((int)st));
        }
        return 0;
    } catch (const std::exception &e) {
        pst::set_error(e.what());
        return 1;
    }
}

}  // extern "C"
