// stager.hpp -- caller-owned (pageable) host memory -> device memory at PCIe speed.
//
// Replay mode (mcts_replay) receives the reference's own trade sequences as one fp64 array in the
// caller's memory: 8 B per (path, trade), by far the largest transfer of the library.  A plain
// cudaMemcpyAsync from pageable memory is staged by the driver through one small bounce buffer on one
// thread (~11 GB/s measured on the B200 box).  Here W worker threads each copy slices of the array into
// their own pinned double buffers and queue the DMA on their own streams, so the host-side memcpy of
// one slice overlaps the DMA of the previous ones and the copy runs at min(host memcpy rate, PCIe).
// Memory that is already pinned (cudaHostAlloc / cudaHostRegister) is copied directly.
#pragma once
#include <cuda_runtime.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

namespace mcts {

class Stager {
  public:
    static constexpr size_t kSlice = (size_t)8 << 20;  // bytes per pinned buffer

    cudaError_t init(int device)
    {
        device_ = device;
        unsigned hw = std::thread::hardware_concurrency();
        const int n = (int)std::max(1u, std::min(8u, hw ? hw : 4u));
        w_.resize((size_t)n);
        for (auto &w : w_) {
            cudaError_t e;
            if ((e = cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
            for (int b = 0; b < 2; b++) {
                if ((e = cudaHostAlloc(&w.buf[b], kSlice, cudaHostAllocDefault)) != cudaSuccess) return e;
                if ((e = cudaEventCreateWithFlags(&w.free_ev[b], cudaEventDisableTiming)) != cudaSuccess) return e;
            }
            if ((e = cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        ready_ = true;
        return cudaSuccess;
    }

    void destroy()
    {
        for (auto &w : w_) {
            if (w.stream) cudaStreamSynchronize(w.stream);
            for (int b = 0; b < 2; b++) {
                if (w.buf[b]) cudaFreeHost(w.buf[b]);
                if (w.free_ev[b]) cudaEventDestroy(w.free_ev[b]);
            }
            if (w.done) cudaEventDestroy(w.done);
            if (w.stream) cudaStreamDestroy(w.stream);
        }
        w_.clear();
        ready_ = false;
    }

    bool ready() const { return ready_; }
    int workers() const { return (int)w_.size(); }

    // Queue dst[0, bytes) <- src[0, bytes).  The DMA starts after `dst_free` (an event of the stream that
    // last read `dst`; may be null) and `consumer` waits for all of it.  Returns when every slice has left
    // the caller's memory (the caller may reuse `src`), not when the DMA has finished.
    cudaError_t copy(void *dst, const void *src, size_t bytes, cudaEvent_t dst_free, cudaStream_t consumer)
    {
        if (bytes == 0) return cudaSuccess;
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost) {  // already pinned
            if (dst_free) cudaStreamWaitEvent(consumer, dst_free, 0);
            return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, consumer);
        }
        cudaGetLastError();  // unregistered pageable memory may report an error on older drivers
        const size_t n_slices = (bytes + kSlice - 1) / kSlice;
        const int n = (int)std::min<size_t>(w_.size(), n_slices);
        std::vector<cudaError_t> err((size_t)n, cudaSuccess);
        auto work = [&](int i) {
            Worker &w = w_[(size_t)i];
            cudaError_t e = cudaSetDevice(device_);
            if (e == cudaSuccess && dst_free) e = cudaStreamWaitEvent(w.stream, dst_free, 0);
            int b = 0;
            for (size_t s = (size_t)i; s < n_slices && e == cudaSuccess; s += (size_t)n, b ^= 1) {
                const size_t off = s * kSlice, len = std::min(kSlice, bytes - off);
                if (w.used[b]) e = cudaEventSynchronize(w.free_ev[b]);  // the DMA that last read this buffer
                if (e != cudaSuccess) break;
                memcpy(w.buf[b], (const char *)src + off, len);
                e = cudaMemcpyAsync((char *)dst + off, w.buf[b], len, cudaMemcpyHostToDevice, w.stream);
                if (e == cudaSuccess) e = cudaEventRecord(w.free_ev[b], w.stream);
                w.used[b] = true;
            }
            if (e == cudaSuccess) e = cudaEventRecord(w.done, w.stream);
            err[(size_t)i] = e;
        };
        std::vector<std::thread> th;
        th.reserve((size_t)n);
        for (int i = 1; i < n; i++) {
            try {
                th.emplace_back(work, i);
            } catch (...) {  // no more threads: this thread does the stripe itself (no C++ exception may cross the C ABI)
                work(i);
            }
        }
        work(0);
        for (auto &t : th) t.join();
        for (int i = 0; i < n; i++) {
            if (err[(size_t)i] != cudaSuccess) return err[(size_t)i];
            cudaError_t e = cudaStreamWaitEvent(consumer, w_[(size_t)i].done, 0);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }

  private:
    struct Worker {
        cudaStream_t stream = nullptr;
        void *buf[2] = {nullptr, nullptr};
        cudaEvent_t free_ev[2] = {nullptr, nullptr};
        bool used[2] = {false, false};
        cudaEvent_t done = nullptr;
    };
    std::vector<Worker> w_;
    int device_ = 0;
    bool ready_ = false;
};

}  // namespace mcts
