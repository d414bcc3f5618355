// Host <-> device staging (see fc_hostpipe.h).
#include "fc_hostpipe.h"

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "fc_common.h"

namespace fcb200 {

// ------------------------------------------------------------------------------------------------
// copy pool
// ------------------------------------------------------------------------------------------------
CopyPool& CopyPool::instance()
{
    static CopyPool pool;
    return pool;
}

CopyPool::CopyPool()
{
    int n = 0;
    if (const char* e = std::getenv("FCB200_COPY_THREADS")) n = std::atoi(e);
    if (n <= 0) n = (int)std::thread::hardware_concurrency() / 2;
    n = std::max(1, std::min(n, 16));
    for (int i = 0; i < n; ++i) workers_.emplace_back([this] { worker(); });
}

CopyPool::~CopyPool()
{
    {
        std::lock_guard<std::mutex> lock(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    for (std::thread& t : workers_) t.join();
}

void CopyPool::worker()
{
    for (;;) {
        std::function<void()> task;
        {
            std::unique_lock<std::mutex> lock(mu_);
            cv_.wait(lock, [this] { return stop_ || !tasks_.empty(); });
            if (stop_ && tasks_.empty()) return;
            task = std::move(tasks_.front());
            tasks_.pop_front();
        }
        task();
    }
}

void CopyPool::copy(void* dst, const void* src, size_t bytes)
{
    const size_t kMinPart = 1u << 20;
    const int parts = (int)std::max<size_t>(1, std::min<size_t>(workers_.size(), bytes / kMinPart));
    if (parts == 1) {
        std::memcpy(dst, src, bytes);
        return;
    }
    struct Join {
        std::mutex mu;
        std::condition_variable cv;
        int left;
    } join;
    join.left = parts - 1;
    const size_t part = ((bytes / parts) + 63) & ~(size_t)63;
    {
        std::lock_guard<std::mutex> lock(mu_);
        for (int i = 1; i < parts; ++i) {
            const size_t off = (size_t)i * part;
            const size_t len = std::min(part, bytes - std::min(bytes, off));
            tasks_.emplace_back([=, &join] {
                if (len) std::memcpy((char*)dst + off, (const char*)src + off, len);
                std::lock_guard<std::mutex> l(join.mu);
                if (--join.left == 0) join.cv.notify_one();
            });
        }
    }
    cv_.notify_all();
    std::memcpy(dst, src, std::min(part, bytes));   // the calling thread takes the first part
    std::unique_lock<std::mutex> l(join.mu);
    join.cv.wait(l, [&] { return join.left == 0; });
}

HostMem classify_pointer(const void* p, int dev)
{
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return HostMem::Pageable;
    }
    if (attr.type == cudaMemoryTypeDevice) {
        if (attr.device != dev) throw std::runtime_error("fcb200: device pointer belongs to another device");
        return HostMem::Device;
    }
    if (attr.type == cudaMemoryTypeManaged) return HostMem::Device;
    if (attr.type == cudaMemoryTypeHost) return HostMem::Pinned;
    return HostMem::Pageable;
}

// ------------------------------------------------------------------------------------------------
// stager
// ------------------------------------------------------------------------------------------------
size_t HostStager::chunk_bytes()
{
    static const size_t v = [] {
        const char* e = std::getenv("FCB200_STAGE_CHUNK_MB");
        const int mb = e ? std::atoi(e) : 16;
        return (size_t)std::max(1, std::min(mb, 256)) << 20;
    }();
    return v;
}

void HostStager::ensure()
{
    const size_t kChunk = chunk_bytes();
    if (in_[0]) return;
    for (int i = 0; i < 2; ++i) {
        FC_CUDA(cudaHostAlloc((void**)&in_[i], kChunk, cudaHostAllocDefault));
        FC_CUDA(cudaHostAlloc((void**)&out_[i], kChunk, cudaHostAllocDefault));
        FC_CUDA(cudaEventCreateWithFlags(&ev_in_[i], cudaEventDisableTiming));
        FC_CUDA(cudaEventCreateWithFlags(&ev_out_[i], cudaEventDisableTiming));
    }
}

HostStager::~HostStager()
{
    for (int i = 0; i < 2; ++i) {
        if (in_[i]) cudaFreeHost(in_[i]);
        if (out_[i]) cudaFreeHost(out_[i]);
        if (ev_in_[i]) cudaEventDestroy(ev_in_[i]);
        if (ev_out_[i]) cudaEventDestroy(ev_out_[i]);
    }
}

void HostStager::upload(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st)
{
    ensure();
    const size_t kChunk = chunk_bytes();
    CopyPool& pool = CopyPool::instance();
    int c = 0;
    for (size_t off = 0; off < bytes; off += kChunk, ++c) {
        const size_t len = std::min(kChunk, bytes - off);
        const int s = c & 1;
        FC_CUDA(cudaEventSynchronize(ev_in_[s]));   // the DMA that last read this slot has finished
        pool.copy(in_[s], (const char*)h_src + off, len);
        FC_CUDA(cudaMemcpyAsync((char*)d_dst + off, in_[s], len, cudaMemcpyHostToDevice, st));
        FC_CUDA(cudaEventRecord(ev_in_[s], st));
    }
}

void HostStager::download(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st)
{
    ensure();
    const size_t kChunk = chunk_bytes();
    CopyPool& pool = CopyPool::instance();
    const int nchunks = (int)((bytes + kChunk - 1) / kChunk);
    auto drain = [&](int c) {   // slot of chunk c -> user memory
        const size_t off = (size_t)c * kChunk;
        const size_t len = std::min(kChunk, bytes - off);
        FC_CUDA(cudaEventSynchronize(ev_out_[c & 1]));
        pool.copy((char*)h_dst + off, out_[c & 1], len);
    };
    for (int c = 0; c < nchunks; ++c) {
        const size_t off = (size_t)c * kChunk;
        const size_t len = std::min(kChunk, bytes - off);
        if (c >= 2) drain(c - 2);   // frees the slot chunk c is about to use
        FC_CUDA(cudaMemcpyAsync(out_[c & 1], (const char*)d_src + off, len, cudaMemcpyDeviceToHost, st));
        FC_CUDA(cudaEventRecord(ev_out_[c & 1], st));
    }
    for (int c = std::max(0, nchunks - 2); c < nchunks; ++c) drain(c);
}

}  // namespace fcb200
