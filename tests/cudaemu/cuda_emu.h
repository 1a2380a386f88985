// cuda_emu.h — TEST INFRASTRUCTURE: runs CUDA kernels on the CPU, one OS thread per CUDA thread of a block, blocks one after another,
// one kernel at a time.
//
// rnaseqc_b200/csrc/rsq_kernels.cu, rsq_covstats.cu and rsq_capi.cu are compiled UNCHANGED with RSQ_CUDA_EMU defined (emu_*.cpp) into
// oracle/_build/librsq_emu.so — the product's whole C-ABI on the CPU — so that the CPU test suite exercises the very code the GPU runs,
// barriers, warp collectives and atomics included:
//   __syncthreads()            a barrier over the block's live threads (threads that returned have left it, as on the device)
//   __shfl_*_sync, __ballot_sync, __any_sync, __match_any_sync, __reduce_*_sync, __syncwarp
//                              exchange through a per-warp slot array between two warp barriers; every live lane of the warp takes part.
//                              A collective over a ONE-lane mask is local, and __activemask() answers "just me": the hot kernel's
//                              opportunistic aggregation (match.any over whatever lanes happen to be converged) accepts any subset
//   atomicAdd / Max / Min / CAS / Or / Sub   real atomics (the threads of a block run concurrently)
//   __shared__                 static storage (one block, one kernel at a time), dynamic shared memory through RSQ_DYNAMIC_SMEM
//   cudaMalloc, cudaMemcpyAsync, streams, events, ...   synchronous stand-ins on host memory ("device" pointers are host pointers)
// The hot kernel's inline PTX (TMA bulk copy + mbarrier, cp.async, ld.global.nc, prefetch) sits behind six helpers in rsq_kernels.cu
// that have an emulation body (memcpy by the issuing lane + warp barrier, plain loads, nothing).
// What this cannot show: races between CTAs (blocks run one after another), anything that depends on the device's memory model
// (everything here is sequentially consistent), performance.
#pragma once
#include <cuda_runtime.h>      // vector types, cudaStream_t (host-side declarations only; nothing of the runtime is called)

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __shared__
#undef __forceinline__
#undef __launch_bounds__
#undef __align__
#undef __noinline__
#define __global__
#define __device__
#define __host__
#define __shared__ static
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __noinline__

namespace cuda_emu {

// a barrier whose party count can shrink (threads that return from the kernel leave it)
class Barrier {
public:
    void reset(unsigned parties) { parties_ = parties; waiting_ = 0; generation_ = 0; }
    void arrive_and_wait() {
        std::unique_lock<std::mutex> lk(mu_);
        const unsigned long gen = generation_;
        if (++waiting_ >= parties_) { waiting_ = 0; ++generation_; cv_.notify_all(); return; }
        cv_.wait(lk, [&] { return gen != generation_; });
    }
    void drop() {
        std::lock_guard<std::mutex> lk(mu_);
        if (parties_) --parties_;
        if (parties_ && waiting_ >= parties_) { waiting_ = 0; ++generation_; cv_.notify_all(); }
    }
private:
    std::mutex mu_;
    std::condition_variable cv_;
    unsigned parties_ = 0, waiting_ = 0;
    unsigned long generation_ = 0;
};

struct Warp {
    Barrier bar;
    unsigned long long slot[32];
    std::atomic<unsigned> live{0};      // mask of the lanes still running
};

struct Block {
    Barrier bar;
    std::vector<Warp> warps;
    std::vector<unsigned char> dyn_smem;
};

struct Tls {
    uint3 tid, bid;
    dim3 bdim, gdim;
    Block* block = nullptr;
    Warp* warp = nullptr;
    unsigned lane = 0;
};
inline Tls& tls() { static thread_local Tls t; return t; }
inline void* dynamic_smem() { return tls().block->dyn_smem.data(); }

inline unsigned to_dim(unsigned v) { return v; }
inline unsigned to_dim(int v) { return (unsigned)v; }
inline unsigned to_dim(unsigned long v) { return (unsigned)v; }
inline unsigned to_dim(unsigned long long v) { return (unsigned)v; }
inline unsigned to_dim(long v) { return (unsigned)v; }

inline std::mutex& kernel_mutex() { static std::mutex m; return m; }

template <class G, class B>
void launch(G grid, B block_threads, size_t smem, const std::function<void()>& body) {
    const unsigned n_blocks = to_dim(grid), n_threads = to_dim(block_threads);
    if (!n_blocks || !n_threads) return;
    // one kernel at a time in the whole process: __shared__ variables are statics here, so launches from several host threads
    // (one context per GPU in the multi-lane CLI) must not overlap
    std::lock_guard<std::mutex> only(kernel_mutex());
    Block blk;
    blk.warps = std::vector<Warp>((n_threads + 31) / 32);
    blk.dyn_smem.assign(smem + 16, 0);
    for (unsigned b = 0; b < n_blocks; ++b) {
        blk.bar.reset(n_threads);
        for (unsigned w = 0; w < blk.warps.size(); ++w) {
            const unsigned lanes = std::min(32u, n_threads - 32 * w);
            blk.warps[w].bar.reset(lanes);
            blk.warps[w].live = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1u);
        }
        std::vector<std::thread> ts;
        ts.reserve(n_threads);
        for (unsigned t = 0; t < n_threads; ++t)
            ts.emplace_back([&, t, b] {
                Tls& T = tls();
                T.tid = make_uint3(t, 0, 0); T.bid = make_uint3(b, 0, 0);
                T.bdim = dim3(n_threads, 1, 1); T.gdim = dim3(n_blocks, 1, 1);
                T.block = &blk; T.warp = &blk.warps[t / 32]; T.lane = t % 32;
                body();
                T.warp->live.fetch_and(~(1u << T.lane));
                T.warp->bar.drop();
                blk.bar.drop();
            });
        for (auto& th : ts) th.join();
    }
}

}  // namespace cuda_emu

#define threadIdx (::cuda_emu::tls().tid)
#define blockIdx (::cuda_emu::tls().bid)
#define blockDim (::cuda_emu::tls().bdim)
#define gridDim (::cuda_emu::tls().gdim)

// ---- synchronisation and warp collectives ---------------------------------------------------------------------------------------
// A collective over a mask of ONE lane is local (the hot kernel's opportunistic aggregation asks __activemask() for the lanes that
// happen to be converged: here that is always just the caller — any subset is a correct answer).  Every other mask must be the
// warp's live lanes, all of which take part.
inline void __syncthreads() { ::cuda_emu::tls().block->bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { ::cuda_emu::tls().warp->bar.arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline unsigned __activemask() { return 1u << ::cuda_emu::tls().lane; }

namespace cuda_emu {
template <class T> inline unsigned long long to_bits(T v) { unsigned long long b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
inline bool solo(unsigned mask) { return (mask & (mask - 1u)) == 0u; }
// every live lane publishes `v`; f(slots, live mask) computes the caller's result between the two warp barriers
template <class T, class F> inline auto collective(T v, F f) {
    Tls& t = tls();
    t.warp->slot[t.lane] = to_bits(v);
    t.warp->bar.arrive_and_wait();
    auto r = f(t.warp->slot, t.warp->live.load(), t.lane);
    t.warp->bar.arrive_and_wait();
    return r;
}
template <class T> inline T exchange(T v, int src) {
    return collective(v, [&](const unsigned long long* slot, unsigned live, unsigned) {
        return (src >= 0 && src < 32 && (live >> src & 1u)) ? from_bits<T>(slot[src]) : v; });
}
}  // namespace cuda_emu
template <class T> inline T __shfl_sync(unsigned mask, T v, int src) { return ::cuda_emu::solo(mask) ? v : ::cuda_emu::exchange(v, src & 31); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int d) { return ::cuda_emu::exchange(v, (int)(::cuda_emu::tls().lane ^ (unsigned)d)); }
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    const int lane = (int)::cuda_emu::tls().lane;
    return ::cuda_emu::exchange(v, lane - (int)d >= 0 ? lane - (int)d : lane);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    if (::cuda_emu::solo(mask)) return pred ? mask : 0u;
    return ::cuda_emu::collective(pred ? 1ull : 0ull, [](const unsigned long long* slot, unsigned live, unsigned) {
        unsigned m = 0;
        for (int l = 0; l < 32; ++l) if ((live >> l & 1u) && slot[l]) m |= 1u << l;
        return m; });
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
template <class T> inline unsigned __match_any_sync(unsigned mask, T v) {
    if (::cuda_emu::solo(mask)) return mask;
    return ::cuda_emu::collective(v, [](const unsigned long long* slot, unsigned live, unsigned lane) {
        unsigned m = 0;
        for (int l = 0; l < 32; ++l) if ((live >> l & 1u) && slot[l] == slot[lane]) m |= 1u << l;
        return m; });
}
namespace cuda_emu {
template <class T, class Op> inline T reduce(unsigned mask, T v, Op op) {
    if (solo(mask)) return v;
    return collective(v, [&](const unsigned long long* slot, unsigned live, unsigned lane) {
        T r = from_bits<T>(slot[lane]);
        bool first = true;
        for (int l = 0; l < 32; ++l) if (live >> l & 1u) { const T x = from_bits<T>(slot[l]); r = first ? x : op(r, x); first = false; }
        return r; });
}
}  // namespace cuda_emu
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return ::cuda_emu::reduce(mask, v, [](unsigned a, unsigned b) { return a + b; }); }
inline int __reduce_add_sync(unsigned mask, int v) { return ::cuda_emu::reduce(mask, v, [](int a, int b) { return (int)((unsigned)a + (unsigned)b); }); }
inline int __reduce_min_sync(unsigned mask, int v) { return ::cuda_emu::reduce(mask, v, [](int a, int b) { return a < b ? a : b; }); }
inline int __reduce_max_sync(unsigned mask, int v) { return ::cuda_emu::reduce(mask, v, [](int a, int b) { return a > b ? a : b; }); }

// ---- atomics ----------------------------------------------------------------------------------------------------------------------
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned int atomicAdd(unsigned int* p, unsigned int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned int atomicSub(unsigned int* p, unsigned int v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
inline unsigned int atomicOr(unsigned int* p, unsigned int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
    unsigned long long cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (cur < v && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return cur;
}
inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
    unsigned long long cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (cur > v && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return cur;
}

// ---- loads, bit tricks, maths ----------------------------------------------------------------------------------------------------
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fadd_rn(float a, float b) { return a + b; }
using std::isinf;
using std::isnan;
using std::sqrt;
template <class T> inline T min(T a, T b) { return b < a ? b : a; }
template <class T> inline T max(T a, T b) { return a < b ? b : a; }
inline unsigned long long min(unsigned long long a, unsigned int b) { return a < b ? a : b; }
inline unsigned int min(unsigned int a, int b) { return a < (unsigned)b ? a : (unsigned)b; }

// ---- the CUDA runtime calls of the library, synchronous on host memory --------------------------------------------------------------
namespace cuda_emu {
template <class T> inline cudaError_t Malloc(T** p, size_t n) { *p = static_cast<T*>(malloc(n ? n : 1)); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t Free(void* p) { free(p); return cudaSuccess; }
inline cudaError_t Memset(void* p, int v, size_t n) { if (n) memset(p, v, n); return cudaSuccess; }
inline cudaError_t Memcpy(void* d, const void* s, size_t n) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t Ok() { return cudaSuccess; }
inline cudaError_t StreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
inline cudaError_t EventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t ElapsedTime(float* ms) { *ms = 0.001f; return cudaSuccess; }
inline cudaError_t DeviceCount(int* n) { *n = 64; return cudaSuccess; }
inline cudaError_t One(int* v) { *v = 1; return cudaSuccess; }
inline cudaError_t Zero(int* v) { *v = 0; return cudaSuccess; }
}  // namespace cuda_emu
#define cudaMalloc(p, n) ::cuda_emu::Malloc((p), (n))
#define cudaMallocHost(p, n) ::cuda_emu::Malloc((p), (n))
#define cudaFree(p) ::cuda_emu::Free((void*)(p))
#define cudaFreeHost(p) ::cuda_emu::Free((void*)(p))
#define cudaMemsetAsync(p, v, n, s) ::cuda_emu::Memset((p), (v), (n))
#define cudaMemcpyAsync(d, s, n, kind, st) ::cuda_emu::Memcpy((d), (s), (n))
#define cudaMemcpy(d, s, n, kind) ::cuda_emu::Memcpy((d), (s), (n))
#define cudaStreamSynchronize(s) ::cuda_emu::Ok()
#define cudaDeviceSynchronize() ::cuda_emu::Ok()
#define cudaSetDevice(d) ::cuda_emu::Ok()
#define cudaGetLastError() ::cuda_emu::Ok()
#define cudaGetDeviceCount(n) ::cuda_emu::DeviceCount(n)
#define cudaGetDevice(d) ::cuda_emu::Zero(d)
#define cudaDeviceGetAttribute(v, attr, dev) ::cuda_emu::One(v)
#define cudaOccupancyMaxActiveBlocksPerMultiprocessor(v, k, t, s) ::cuda_emu::One(v)
#define cudaStreamCreateWithFlags(s, f) ::cuda_emu::StreamCreate(s)
#define cudaStreamDestroy(s) ::cuda_emu::Ok()
#define cudaStreamWaitEvent(s, e, f) ::cuda_emu::Ok()
#define cudaEventCreate(e) ::cuda_emu::EventCreate(e)
#define cudaEventCreateWithFlags(e, f) ::cuda_emu::EventCreate(e)
#define cudaEventDestroy(e) ::cuda_emu::Ok()
#define cudaEventRecord(e, s) ::cuda_emu::Ok()
#define cudaEventSynchronize(e) ::cuda_emu::Ok()
#define cudaEventElapsedTime(ms, a, b) ::cuda_emu::ElapsedTime(ms)
#define cudaFuncSetAttribute(...) ::cuda_emu::Ok()
#define cudaGetErrorString(e) "emulated runtime"
