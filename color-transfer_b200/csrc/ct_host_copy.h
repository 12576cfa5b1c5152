// Copies between PAGEABLE caller memory (a numpy array) and the device, for the single-pair
// drop-in calls (methods.* on one [H,W,3] pair).  cudaMemcpy on pageable memory moves
// ~17 GB/s (the driver stages it through one thread); a 0964-size float64 call spent 3.9 of its
// 4.2 ms there.  Here a small persistent pool of host threads copies 8 MB chunks between the caller's
// memory and two pinned bounce buffers while the DMA engine moves the previous chunk at PCIe speed;
// on the way out the same threads also take the first-touch page faults of the fresh output array
// in parallel.  Pinned / registered caller memory and small images keep the plain cudaMemcpyAsync.
// Measured on the GPU box (CT_PROFILE_HOST=1): H2D of 2 x 22 MB 2.6 -> 1.9 ms, D2H of 22 MB into a
// fresh array 1.3 -> 1.6 ms incl. the kernels; the host memory system (~27 GB/s of memcpy across
// four threads) is the limit, so the whole call goes from 4.2-4.6 to 3.3-4.5 ms on idle cores -
// and to 6 ms when other threads of the process keep the cores busy (BLAS workers spinning after a
// numpy call).  Hence OPT-IN (CT_STAGED_COPY=1, see run_host_pipeline); the default is the driver's
// pageable cudaMemcpyAsync.
#pragma once

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "ct_context.h"

namespace ct {

class CopyPool {
public:
    static CopyPool &get() {
        static CopyPool *pool = new CopyPool();  // leaked on purpose: workers are detached
        return *pool;
    }
    // memcpy(dst, src, n) split over the pool and the calling thread; returns when done
    void copy(void *dst, const void *src, size_t n) {
        if (workers_ == 0 || n < (1u << 20)) {
            memcpy(dst, src, n);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            dst_ = static_cast<char *>(dst);
            src_ = static_cast<const char *>(src);
            n_ = n;
            piece_ = (((n + (size_t)(workers_ + 1) * 2 - 1) / ((size_t)(workers_ + 1) * 2)) + 4095) & ~(size_t)4095;
            npieces_ = (n + piece_ - 1) / piece_;
            next_.store(0);
            active_ = workers_;
            ++gen_;
        }
        cv_.notify_all();
        run();
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return active_ == 0; });
    }

private:
    CopyPool() {
        unsigned hw = std::thread::hardware_concurrency();
        int want = hw > 2 ? (int)(hw / 2) - 1 : 0;
        if (want > 3) want = 3;  // four copying threads saturate the host memory system (~27 GB/s measured)
        if (const char *e = getenv("CT_COPY_THREADS")) want = atoi(e) - 1;
        if (want < 0) want = 0;
        for (int i = 0; i < want; ++i) {
            try {
                std::thread(&CopyPool::worker, this).detach();
                ++workers_;
            } catch (...) {
                break;
            }
        }
    }
    void run() {
        for (;;) {
            const size_t i = next_.fetch_add(1);
            if (i >= npieces_) break;
            const size_t off = i * piece_;
            memcpy(dst_ + off, src_ + off, n_ - off < piece_ ? n_ - off : piece_);
        }
    }
    void worker() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
            }
            run();
            std::lock_guard<std::mutex> lk(m_);
            if (--active_ == 0) done_.notify_one();
        }
    }
    std::mutex m_;
    std::condition_variable cv_, done_;
    char *dst_ = nullptr;
    const char *src_ = nullptr;
    size_t n_ = 0, piece_ = 0, npieces_ = 0;
    std::atomic<size_t> next_{0};
    int active_ = 0, workers_ = 0;
    uint64_t gen_ = 0;
};

constexpr size_t kBounceChunk = 8u << 20;

inline bool is_pageable(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

inline int ensure_bounce(ct_context *h) {
    if (h->bounce[0][0]) return CT_OK;
    for (int d = 0; d < 2; ++d)
        for (int s = 0; s < 2; ++s) {
            CT_CUDA(h, cudaHostAlloc(reinterpret_cast<void **>(&h->bounce[d][s]), kBounceChunk, cudaHostAllocDefault));
            CT_CUDA(h, cudaEventCreateWithFlags(&h->bounce_done[d][s], cudaEventDisableTiming));
        }
    return CT_OK;
}

// host (pageable) -> device, ordered on `stream`.  Blocks the calling thread while it copies.
inline int staged_h2d(ct_context *h, void *dev, const void *host, size_t bytes, cudaStream_t stream) {
    CT_TRY(ensure_bounce(h));
    CopyPool &pool = CopyPool::get();
    int k = 0;
    for (size_t off = 0; off < bytes; off += kBounceChunk, ++k) {
        const int s = k & 1;
        const size_t n = bytes - off < kBounceChunk ? bytes - off : kBounceChunk;
        CT_CUDA(h, cudaEventSynchronize(h->bounce_done[0][s]));  // the DMA that last read this buffer
        pool.copy(h->bounce[0][s], static_cast<const char *>(host) + off, n);
        CT_CUDA(h, cudaMemcpyAsync(static_cast<char *>(dev) + off, h->bounce[0][s], n, cudaMemcpyHostToDevice, stream));
        CT_CUDA(h, cudaEventRecord(h->bounce_done[0][s], stream));
    }
    return CT_OK;
}

// device -> host (pageable), after everything queued on `stream`.  Returns when the data is there.
inline int staged_d2h(ct_context *h, void *host, const void *dev, size_t bytes, cudaStream_t stream) {
    CT_TRY(ensure_bounce(h));
    CopyPool &pool = CopyPool::get();
    const size_t nchunks = (bytes + kBounceChunk - 1) / kBounceChunk;
    auto issue = [&](size_t c) -> int {
        const int s = (int)(c & 1);
        const size_t off = c * kBounceChunk, n = bytes - off < kBounceChunk ? bytes - off : kBounceChunk;
        CT_CUDA(h, cudaMemcpyAsync(h->bounce[1][s], static_cast<const char *>(dev) + off, n, cudaMemcpyDeviceToHost, stream));
        CT_CUDA(h, cudaEventRecord(h->bounce_done[1][s], stream));
        return CT_OK;
    };
    for (size_t c = 0; c < nchunks && c < 2; ++c) CT_TRY(issue(c));
    for (size_t c = 0; c < nchunks; ++c) {
        const int s = (int)(c & 1);
        const size_t off = c * kBounceChunk, n = bytes - off < kBounceChunk ? bytes - off : kBounceChunk;
        CT_CUDA(h, cudaEventSynchronize(h->bounce_done[1][s]));
        pool.copy(static_cast<char *>(host) + off, h->bounce[1][s], n);
        if (c + 2 < nchunks) CT_TRY(issue(c + 2));
    }
    return CT_OK;
}

}  // namespace ct
