// Library plumbing: ABI version, thread-local error string, device check, per-device
// cache of small constant tables.
#include "common.cuh"

#include <map>
#include <mutex>
#include <set>
#include <vector>

namespace advmix {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

static std::mutex g_mu;
static std::map<int, int> g_sm_count;
static std::map<std::pair<int, std::string>, void*> g_tables;

static std::set<std::pair<int, const void*>> g_first_use;

bool first_use_on_device(const void* tag) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return true;
    std::lock_guard<std::mutex> lk(g_mu);
    return g_first_use.insert(std::make_pair(dev, tag)).second;
}

int sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_sm_count.find(dev);
    if (it != g_sm_count.end()) return it->second;
    int n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    g_sm_count[dev] = n;
    return n;
}

const void* cached_table(const std::string& key, const void* host, size_t bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    auto k = std::make_pair(dev, key);
    auto it = g_tables.find(k);
    if (it != g_tables.end()) return it->second;
    void* d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess) {
        set_error("cached_table(%s): cudaMalloc(%zu) failed", key.c_str(), bytes);
        return nullptr;
    }
    // Synchronous copy from pageable memory: completes before return, so later
    // launches on any stream see the table.
    if (cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("cached_table(%s): cudaMemcpy failed", key.c_str());
        cudaFree(d);
        return nullptr;
    }
    g_tables[k] = d;
    return d;
}

// zero-initialised device scratch, allocated once per (device, key)
void* cached_buffer(const std::string& key, size_t bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    auto k = std::make_pair(dev, key);
    auto it = g_tables.find(k);
    if (it != g_tables.end()) return it->second;
    if (bytes == 0) return nullptr;                  // peek only (e.g. while a stream capture forbids cudaMalloc)
    void* d = nullptr;
    if (cudaMalloc(&d, bytes) != cudaSuccess || cudaMemset(d, 0, bytes) != cudaSuccess) {
        set_error("cached_buffer(%s): cudaMalloc / cudaMemset(%zu) failed", key.c_str(), bytes);
        cudaGetLastError();
        return nullptr;
    }
    g_tables[k] = d;
    return d;
}

}  // namespace advmix

extern "C" {

int advmix_abi_version(void) { return ADVMIX_ABI_VERSION; }

const char* advmix_last_error(void) { return advmix::g_last_error.c_str(); }

int advmix_device_check(int device) {
    cudaDeviceProp p;
    ADVMIX_CUDA_OK(cudaGetDeviceProperties(&p, device));
    if (p.major != 10)
        return advmix::fail(ADVMIX_ERR_UNSUPPORTED,
                            "device %d is sm_%d%d; libadvmix_b200 carries sm_100a code only", device,
                            p.major, p.minor);
    return ADVMIX_OK;
}

}  // extern "C"
