// common.cuh -- shared host/device helpers for libmimeo_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace mb2 {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define MB2_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            throw ::mb2::Error(-100, std::string(#expr) + ": " + cudaGetErrorString(_e) +      \
                                         " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

#define MB2_REQUIRE(cond, code, msg)                       \
    do {                                                   \
        if (!(cond)) throw ::mb2::Error((code), (msg));    \
    } while (0)

// One context per process (one process per GPU). Owns the stream every kernel of the
// library is launched on, the stream-ordered memory pool and a launch counter.
struct Ctx {
    int device = -1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool ready = false;
    unsigned long long launches = 0;   // kernels of THIS library launched so far
    // optional per-kernel timing with CUDA events on the library stream (bench.py's roofline leg)
    bool debug_sync = false;
    bool prof = false;
    struct ProfRec { std::string tag; cudaEvent_t a, b; };
    std::vector<ProfRec> prof_recs;
    std::map<std::string, std::pair<double, unsigned long long>> prof_acc;   // tag -> (ms, count)
};
Ctx& ctx();
void ensure_init();

// Device scratch allocator of the library: a host-side best-fit allocator over a few large slabs obtained with cudaMalloc
// and kept until mb2_shutdown. Every user of the memory is enqueued on the ONE library stream, so a freed block may be
// handed out again at once (stream order protects it), and the steady state of a pipeline makes NO driver calls: driver
// allocation calls contend with other driver clients (an nvidia-smi poller stalled them for tens of milliseconds per step).
void* scratch_alloc(size_t bytes);
void scratch_free(void* p);
void scratch_release_all();        // mb2_shutdown
size_t scratch_reserved_bytes();

// Stream-ordered scratch buffer on the library stream.
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) p = static_cast<T*>(scratch_alloc(count * sizeof(T)));
    }
    void release() {
        if (p) scratch_free(p);
        p = nullptr; n = 0;
    }
    T* get() const { return p; }
};

inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

// Launch on the library stream, count it, and surface launch errors immediately.
template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
    kernel<<<grid, block, smem, ctx().stream>>>(static_cast<KArgs>(args)...);
    ctx().launches++;
    MB2_CUDA(cudaGetLastError());
    if (ctx().debug_sync) {   // MB2_DEBUG_SYNC=1: localise asynchronous faults to the launch that caused them
        cudaError_t e = cudaStreamSynchronize(ctx().stream);
        if (e != cudaSuccess)
            throw Error(-100, std::string("kernel fault after launch #") + std::to_string(ctx().launches) + " grid " +
                                  std::to_string(grid.x) + " block " + std::to_string(block.x) + ": " + cudaGetErrorString(e));
    }
}

// Times everything enqueued on the library stream during its lifetime (only when profiling is on).
struct ProfScope {
    bool on;
    cudaEvent_t a{}, b{};
    const char* tag;
    explicit ProfScope(const char* t) : on(ctx().prof), tag(t) {
        if (on) {
            cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a, ctx().stream);
        }
    }
    ~ProfScope() {
        if (on) {
            cudaEventRecord(b, ctx().stream);
            ctx().prof_recs.push_back({tag, a, b});
        }
    }
};

}  // namespace mb2
