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

ConstBankGuard::ConstBankGuard(ConstBank bank, cudaStream_t st) : bank_(bank), dev_(0), st_(st) {
    BankState& b = g_banks[bank_];
    b.mu.lock();
    if (cudaGetDevice(&dev_) != cudaSuccess || dev_ < 0 || dev_ >= 64) dev_ = 0;
    if (b.done[dev_] == nullptr) cudaEventCreateWithFlags(&b.done[dev_], cudaEventDisableTiming);
    if (b.used[dev_] && b.last[dev_] != st_) cudaStreamWaitEvent(st_, b.done[dev_], 0);
}

ConstBankGuard::~ConstBankGuard() {
    BankState& b = g_banks[bank_];
    if (b.done[dev_] != nullptr) cudaEventRecord(b.done[dev_], st_);
    b.last[dev_] = st_;
    b.used[dev_] = true;
    b.mu.unlock();
}

}  // namespace gnb

extern "C" {

int32_t gnb_version(void) { return 100; /* 0.1.0 */ }

const char* gnb_last_error(void) { return gnb::g_err; }

int32_t gnb_device_sm_count(void) {
    int dev = 0, n = 0;
    GNB_CUDA(cudaGetDevice(&dev));
    GNB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    return n;
}

}  // extern "C"
