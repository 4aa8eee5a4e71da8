// Library-level entry points: version, last error, device query.
#include "common.cuh"
#include <string.h>
#include <mutex>

namespace gnb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int cached = 0;
    if (cached > 0) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    cached = n;
    return n;
}

namespace {
struct BankState {
    std::mutex mu;
    bool used[64] = {false};
    cudaStream_t last[64] = {nullptr};
    cudaEvent_t done[64] = {nullptr};
};
BankState g_banks[BANK_COUNT];
}  // namespace

// A stream that is being captured into a CUDA graph (pipeline.py captures the static front part of predict()) must not wait on
// or record the guard's events: inside one captured stream the launches are ordered anyway, and a replayed graph is ordered
// against other work by the stream it is launched on.  The guard then only serialises the host side.
static bool stream_is_capturing(cudaStream_t st) {
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &status) != cudaSuccess) { cudaGetLastError(); return false; }
    return status != cudaStreamCaptureStatusNone;
}

ConstBankGuard::ConstBankGuard(ConstBank bank, cudaStream_t st) : bank_(bank), dev_(0), st_(st) {
    BankState& b = g_banks[bank_];
    b.mu.lock();
    if (cudaGetDevice(&dev_) != cudaSuccess || dev_ < 0 || dev_ >= 64) dev_ = 0;
    if (stream_is_capturing(st_)) return;
    if (b.done[dev_] == nullptr) cudaEventCreateWithFlags(&b.done[dev_], cudaEventDisableTiming);
    if (b.used[dev_] && b.last[dev_] != st_) cudaStreamWaitEvent(st_, b.done[dev_], 0);
}

ConstBankGuard::~ConstBankGuard() {
    BankState& b = g_banks[bank_];
    if (!stream_is_capturing(st_)) {
        if (b.done[dev_] != nullptr) cudaEventRecord(b.done[dev_], st_);
        b.last[dev_] = st_;
        b.used[dev_] = true;
    }
    b.mu.unlock();
}

// ---- fp16 operand range flag ---------------------------------------------------------------------------------------------
// The tensor-core paths split fp32 operands into fp16 hi + lo with SATURATING conversions: an activation beyond +-65504 would
// silently become 65504.  One 32-bit flag per device records that this happened: set by the normalise-and-split pass of the
// UNet and by gnb_f16_range_check (run by the pipeline on the grids the decoders read), fetched by gnb_f16_overflow_fetch.
uint32_t* f16_flag_ptr() {
    static uint32_t* flags[64] = {nullptr};
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    std::lock_guard<std::mutex> lock(mu);
    if (flags[dev] == nullptr) {
        if (cudaMalloc(&flags[dev], sizeof(uint32_t)) != cudaSuccess) return nullptr;
        cudaMemset(flags[dev], 0, sizeof(uint32_t));
    }
    return flags[dev];
}

__global__ void __launch_bounds__(256)
f16_range_check_kernel(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ flag) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool bad = false;   // !(|v| <= limit) is true for out-of-range values, infinities AND NaNs
    const int64_t n4 = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) ? n / 4 : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        bad |= !(fabsf(v.x) <= 65504.f) | !(fabsf(v.y) <= 65504.f) | !(fabsf(v.z) <= 65504.f) | !(fabsf(v.w) <= 65504.f);
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) bad |= !(fabsf(x[i]) <= 65504.f);
    if (bad) *flag = 1u;
}

// ---- small device -> host transfers without the copy engine ----------------------------------------------------------------
// A few hundred bytes the host is WAITING for (marching-cubes totals, the range flag) must not queue behind the bulk
// device -> host transfers of the previous batch on the copy engine (HostPredictor: 316 MB per batch; with eight ranks sharing
// one host those transfers occupy the engine for most of a step, and a cudaMemcpy of 2 KB waited milliseconds behind them).
// Pinned host memory is device-addressable under unified addressing: a kernel stores the words straight into it.
__global__ void __launch_bounds__(256)
copy_words_to_host_kernel(const uint32_t* __restrict__ src, int64_t src_pitch_words, uint32_t* __restrict__ host_dst,
                          int64_t dst_pitch_words, int width_words, int64_t rows) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * width_words) return;
    const int64_t r = t / width_words;
    const int c = (int)(t - r * width_words);
    host_dst[r * dst_pitch_words + c] = src[r * src_pitch_words + c];
    __threadfence_system();
}

__global__ void f16_flag_to_host_kernel(uint32_t* __restrict__ flag, uint32_t* __restrict__ host_dst, int reset) {
    *host_dst = *flag;
    if (reset) *flag = 0u;
    __threadfence_system();
}

}  // namespace gnb

extern "C" {

int32_t gnb_copy_to_pinned_host(const void* src, int64_t src_pitch, void* pinned_host_dst, int64_t dst_pitch,
                                int64_t width_bytes, int64_t rows, void* stream) {
    GNB_REQUIRE(src && pinned_host_dst, "gnb_copy_to_pinned_host: null pointer");
    GNB_REQUIRE(width_bytes >= 0 && rows >= 0 && width_bytes <= (1 << 20) && width_bytes * rows <= (1ll << 24),
                "gnb_copy_to_pinned_host: meant for small records (<= 16 MiB)");
    GNB_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(pinned_host_dst) | (uintptr_t)src_pitch |
                  (uintptr_t)dst_pitch | (uintptr_t)width_bytes) & 3) == 0,
                "gnb_copy_to_pinned_host: pointers, pitches and width must be multiples of 4 bytes");
    if (width_bytes == 0 || rows == 0) return GNB_OK;
    cudaPointerAttributes attr;
    GNB_CUDA(cudaPointerGetAttributes(&attr, pinned_host_dst));
    GNB_REQUIRE(attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr,
                "gnb_copy_to_pinned_host: destination is not pinned (device-addressable) host memory");
    const int64_t words = width_bytes / 4 * rows;
    gnb::copy_words_to_host_kernel<<<(unsigned)((words + 255) / 256), 256, 0, gnb::as_stream(stream)>>>(
        static_cast<const uint32_t*>(src), src_pitch / 4, static_cast<uint32_t*>(attr.devicePointer), dst_pitch / 4,
        (int)(width_bytes / 4), rows);
    return gnb::check_launch("gnb_copy_to_pinned_host");
}

int32_t gnb_f16_range_check(const float* x, int64_t n, void* stream) {
    GNB_REQUIRE(x || n == 0, "gnb_f16_range_check: null pointer");
    if (n <= 0) return GNB_OK;
    uint32_t* flag = gnb::f16_flag_ptr();
    GNB_REQUIRE(flag != nullptr, "gnb_f16_range_check: flag allocation failed");
    int64_t blocks = (n / 4 + 255) / 256;
    const int64_t cap = (int64_t)gnb::sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    gnb::f16_range_check_kernel<<<(unsigned)blocks, 256, 0, gnb::as_stream(stream)>>>(x, n, flag);
    return gnb::check_launch("gnb_f16_range_check");
}

int32_t gnb_f16_overflow_fetch(int32_t reset, void* stream) {
    uint32_t* flag = gnb::f16_flag_ptr();
    GNB_REQUIRE(flag != nullptr, "gnb_f16_overflow_fetch: flag allocation failed");
    cudaStream_t st = gnb::as_stream(stream);
    uint32_t h = 0;
    GNB_CUDA(cudaMemcpyAsync(&h, flag, sizeof(h), cudaMemcpyDeviceToHost, st));
    if (reset) GNB_CUDA(cudaMemsetAsync(flag, 0, sizeof(h), st));
    GNB_CUDA(cudaStreamSynchronize(st));
    return h ? 1 : 0;
}

int32_t gnb_f16_overflow_fetch_async(uint32_t* pinned_host_out, int32_t reset, void* stream) {
    GNB_REQUIRE(pinned_host_out != nullptr, "gnb_f16_overflow_fetch_async: null pointer");
    uint32_t* flag = gnb::f16_flag_ptr();
    GNB_REQUIRE(flag != nullptr, "gnb_f16_overflow_fetch_async: flag allocation failed");
    // a kernel store into the (device-addressable) pinned word, not a cudaMemcpyAsync: see copy_words_to_host_kernel
    cudaPointerAttributes attr;
    GNB_CUDA(cudaPointerGetAttributes(&attr, pinned_host_out));
    GNB_REQUIRE(attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr,
                "gnb_f16_overflow_fetch_async: destination is not pinned (device-addressable) host memory");
    gnb::f16_flag_to_host_kernel<<<1, 1, 0, gnb::as_stream(stream)>>>(flag, static_cast<uint32_t*>(attr.devicePointer), reset);
    return gnb::check_launch("gnb_f16_overflow_fetch_async");
}

int32_t gnb_version(void) { return 100; /* 0.1.0 */ }

const char* gnb_last_error(void) { return gnb::g_err; }

int32_t gnb_device_sm_count(void) {
    int dev = 0, n = 0;
    GNB_CUDA(cudaGetDevice(&dev));
    GNB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    return n;
}

}  // extern "C"
