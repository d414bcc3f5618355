// Host <-> device staging for the host-pointer ABI (host C++).
//
// The reference moves every volume with blocking cudaMemcpy from / to pageable memory
// (/root/reference/src/convolution3Dfft.cu:512-515, :551-554).  Callers such as Fiji/JNA hand over pageable
// arrays, which the driver copies through its own small bounce buffer, single-threaded.  Here pageable
// buffers are staged through pinned slots by a few copy threads, chunk by chunk, so the host memcpy of chunk
// c+1 overlaps the DMA of chunk c; pinned (or registered) buffers are DMA'd directly.
#pragma once
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstddef>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace fcb200 {

// Small persistent pool used for parallel memcpy.  Safe to use from several host threads at once.
class CopyPool {
public:
    static CopyPool& instance();
    int threads() const { return (int)workers_.size(); }
    // dst <- src, split over the pool; returns when every part has been copied
    void copy(void* dst, const void* src, size_t bytes);
    ~CopyPool();

private:
    CopyPool();
    void worker();
    std::vector<std::thread> workers_;
    std::deque<std::function<void()>> tasks_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool stop_ = false;
};

enum class HostMem { Pageable, Pinned, Device };
HostMem classify_pointer(const void* p, int dev);

// Two pinned slots per direction; one stager per plan (calls on a plan are serialised by the plan's mutex,
// except that one upload and one download may run concurrently from two threads).
class HostStager {
public:
    static size_t chunk_bytes();   // FCB200_STAGE_CHUNK_MB, default 16 (measured: profiles/r01_host_path.jsonl)
    HostStager() = default;
    ~HostStager();
    HostStager(const HostStager&) = delete;
    HostStager& operator=(const HostStager&) = delete;

    void prepare() { ensure(); }   // allocate the slots (call before using the stager from two threads)
    // pageable host -> device.  Returns when every chunk has been staged and its H2D copy enqueued on `st`.
    void upload(void* d_dst, const void* h_src, size_t bytes, cudaStream_t st);
    // device -> pageable host.  Returns when the data is in h_dst.  The copies are ordered after the work
    // already enqueued on `st`.
    void download(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st);

private:
    void ensure();
    char* in_[2] = {nullptr, nullptr};
    char* out_[2] = {nullptr, nullptr};
    cudaEvent_t ev_in_[2] = {nullptr, nullptr};
    cudaEvent_t ev_out_[2] = {nullptr, nullptr};
};

}  // namespace fcb200
