// In-library multi-GPU orchestration: ONE process drives several devices with peer access enabled (SURVEY section 5 /
// section 7 step 9; the reference's only multi-GPU hook is cudaSetDevice(devCUDA), /root/reference/src/convolution3Dfft.cu:409).
//
//  * slab_convolve: one volume cut in z slabs, one rank per device, one persistent worker thread per rank.
//      phase A  x + y forward on the rank's planes; the last stage of the y pass stores every ky row straight into
//               the owning rank's y-slab buffer over NVLink (RowsSplit, fft_static.cuh) -- meanwhile the rank's slab of
//               the PSF spectrum is built on a side stream;
//      phase B  forward z . x H x 1/N . inverse z on the rank's ky rows; the last inverse stage stores every plane
//               straight into its owner's receive buffer;
//      phase C  y + x inverse on the rank's planes.
//    The phases are separated by CUDA events only: every rank records an event after its phase and every other
//    rank's stream waits for it (cudaStreamWaitEvent across devices) -- no NCCL, no IPC, no host synchronisation
//    besides the thread barrier that orders "all records issued" before "all waits issued".
//  * batch_multi: independent blocks, one pipelined batch (batch_core, fc_api.cu) per device, all fed from one
//    shared counter.
#include "fc_multi.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <map>
#include <thread>

#include "fc_api_util.h"

namespace fcb200 {

namespace {

// All workers of a context meet here between phases (spin + yield: the waits are microseconds long).
class ThreadBarrier {
public:
    explicit ThreadBarrier(int n) : n_(n) {}
    void wait()
    {
        const int gen = gen_.load(std::memory_order_acquire);
        if (count_.fetch_add(1, std::memory_order_acq_rel) + 1 == n_) {
            count_.store(0, std::memory_order_relaxed);
            gen_.fetch_add(1, std::memory_order_release);
            return;
        }
        int spins = 0;
        while (gen_.load(std::memory_order_acquire) == gen)
            if (++spins > 64) std::this_thread::yield();
    }

private:
    const int n_;
    std::atomic<int> count_{0};
    std::atomic<int> gen_{0};
};

struct SlabRank {
    int dev = 0, rank = 0;
    int nzl = 0, ny_here = 0;
    std::shared_ptr<ConvPlan> plan;   // tables only; the slab-sized buffers live here
    cudaStream_t st = nullptr, s_psf = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    cudaStream_t s_x[4] = {nullptr, nullptr, nullptr, nullptr};   // exchange copies (several copy engines at once)
    float* real = nullptr;            // host-pointer calls: this rank's z slab
    float2 *zslab = nullptr, *recv = nullptr, *yslab = nullptr, *H = nullptr, *scratch = nullptr;
    float2* send = nullptr;           // pull exchange: exchange-layout output of my y pass, read by the peers' fused z pass
    float2** d_peer_send = nullptr;   // entry q: rank q's send buffer at MY block
    size_t scratch_cap = 0;
    float* d_kernel = nullptr;
    size_t kernel_cap = 0;
    float2** d_peer_yslab = nullptr;  // device arrays of P pointers
    float2** d_peer_recv = nullptr;
    cudaEvent_t ev_fwd = nullptr, ev_z = nullptr, ev_done = nullptr, ev_psf = nullptr;
    cudaEvent_t ev_t[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_chunk[16] = {};
    cudaEvent_t ev_d2h = nullptr;
    cudaEvent_t ev_x[4] = {nullptr, nullptr, nullptr, nullptr};
    std::unique_ptr<HostStager> stager;   // pageable host volumes (per rank: emulated ranks share a plan)
    bool timed = false;
};

struct SlabCall {
    float* im = nullptr;
    float* const* slabs = nullptr;
    const float* kernel = nullptr;
    HostMem im_kind = HostMem::Device;
    bool psf_cached = false;
    int pdims[6] = {0, 0, 0, 0, 0, 0};
    int exchange = 1;   // forward exchange: 0 peer stores from the y pass, 1 copy engines, 2 pulled by the fused z pass
};

struct SlabContext {
    std::vector<int> devs;
    int dims[3] = {0, 0, 0};
    int P = 0, nzp = 0, nyl = 0, xcp = 0;
    std::vector<SlabRank> ranks;
    std::unique_ptr<ThreadBarrier> barrier;
    // worker pool
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    unsigned long long job_gen = 0;
    int done = 0;
    bool stop = false;
    SlabCall call;
    std::atomic<bool> failed{false};
    std::vector<std::string> errors;
    std::mutex call_mu;   // one call at a time per context
    // PSF-spectrum cache across calls (host-pointer kernels; FCB200_PSF_CACHE)
    bool h_valid = false;
    int h_dims[6] = {0, 0, 0, 0, 0, 0};
    std::vector<float> h_taps;
    unsigned long long last_use = 0;
    int scratch_key[6] = {0, 0, 0, 0, 0, 0};   // PSF scratch size of these placement dims (host-side row scan, cached)
    size_t scratch_need = 0;

    size_t slab_spec_elems() const { return (size_t)P * nzp * nyl * xcp; }
    void worker(int r);
    void run_call(int r);
    void phase(int r, const std::function<void()>& fn);
    ~SlabContext();
};

void SlabContext::phase(int r, const std::function<void()>& fn)
{
    if (failed.load()) return;
    try {
        fn();
    } catch (const std::exception& e) {
        errors[(size_t)r] = e.what();
        failed.store(true);
        cudaGetLastError();
    }
}

// One call, rank r (runs on the rank's worker thread; its device is current).
void SlabContext::run_call(int r)
{
    SlabRank& k = ranks[(size_t)r];
    ConvPlan& p = *k.plan;
    const SlabCall& c = call;
    const size_t rplane = (size_t)dims[1] * dims[0];
    const size_t slab_bytes = (size_t)k.nzl * rplane * sizeof(float);
    const size_t z_first = (size_t)r * nzp;
    float* real = c.slabs ? c.slabs[r] : k.real;
    float* h_slab = c.im ? c.im + z_first * rplane : nullptr;
    static const bool staging_on = env_flag("FCB200_STAGING", true);
    const bool pinned = c.im && (c.im_kind == HostMem::Pinned || !staging_on);
    // Forward exchange.  Peer stores from the y pass are 128-byte segments (64-byte for L = 2048, whose tile only fits
    // shared memory with half rows): config 5 on 8 GPUs reached 300 GB/s per GPU that way and the forward phase took
    // 6.3 ms for 2.5 ms of local work.  With copy_exchange the y pass writes the exchange layout LOCALLY ([P][nzp][nyl][xcp],
    // into the receive buffer, which is idle until the fused z pass of the peers) and the copy engines move block q to
    // rank q in one contiguous transfer per chunk of planes, while the next chunk is being transformed.
    const bool pull_x = c.exchange == 2 && P > 1 && k.send != nullptr;
    const bool copy_x = (c.exchange == 1 || (c.exchange == 2 && !pull_x)) && P > 1;
    static const int xchunks = [] {
        const char* e = std::getenv("FCB200_SLAB_XCHUNKS");
        return std::max(1, std::min(8, e ? std::atoi(e) : 4));
    }();
    static const int nxs = [] {   // exchange streams (measured on 8 GPUs, config 5: 1 stream 11.3 ms, 4 streams 11.6 ms)
        const char* e = std::getenv("FCB200_SLAB_XSTREAMS");
        return std::max(1, std::min(4, e ? std::atoi(e) : 1));
    }();
    const int nch = c.im ? (int)std::max<long long>(1, std::min<long long>(8, std::min<long long>(k.nzl, (long long)(slab_bytes >> 25))))
                         : (copy_x ? (int)std::max(1, std::min(xchunks, k.nzl / 8)) : 1);
    const int per = (k.nzl + nch - 1) / nch;
    auto z0_of = [&](int ch) { return std::min(k.nzl, ch * per); };

    // ---- phase A: PSF slab on the side stream, upload, x + y forward with peer stores
    phase(r, [&] {
        // every rank has finished the previous call (it read my buffers, I am about to overwrite its)
        for (SlabRank& q : ranks) FC_CUDA(cudaStreamWaitEvent(k.st, q.ev_done, 0));
        if (!c.psf_cached) {
            std::lock_guard<std::mutex> lock(p.mu);
            FC_CUDA(cudaStreamWaitEvent(k.s_psf, k.ev_done, 0));   // my previous fused z pass has read H
            const size_t ktaps = (size_t)c.pdims[0] * c.pdims[1] * c.pdims[2];
            FC_CUDA(cudaMemcpyAsync(k.d_kernel, c.kernel, ktaps * sizeof(float), cudaMemcpyDefault, k.s_psf));
            run_slab_psf(p, k.d_kernel, c.pdims, r * nyl, nyl, k.H, k.scratch, k.s_psf);
            FC_CUDA(cudaEventRecord(k.ev_psf, k.s_psf));
        }
        if (c.im && !pinned) k.stager->upload(real, h_slab, slab_bytes, k.st);
        FC_CUDA(cudaEventRecord(k.ev_t[0], k.st));
        for (int ch = 0; ch < nch; ++ch) {
            const int z0 = z0_of(ch), n = z0_of(ch + 1) - z0;
            if (n <= 0) continue;
            if (pinned) {
                if (ch == 0) FC_CUDA(cudaStreamWaitEvent(k.s_h2d, k.ev_done, 0));
                FC_CUDA(cudaMemcpyAsync(real + z0 * rplane, h_slab + z0 * rplane, n * rplane * sizeof(float),
                                        cudaMemcpyHostToDevice, k.s_h2d));
                FC_CUDA(cudaEventRecord(k.ev_chunk[ch], k.s_h2d));
                FC_CUDA(cudaStreamWaitEvent(k.st, k.ev_chunk[ch], 0));
            }
            if (pull_x) {   // the y pass only writes the exchange layout; the peers' fused z pass fetches it
                run_slab_xy_forward(p, real, k.zslab, k.send, k.nzl, nyl, k.st, nullptr, r, nzp, z0, n);
                continue;
            }
            if (!copy_x) {
                run_slab_xy_forward(p, real, k.zslab, nullptr, k.nzl, nyl, k.st, k.d_peer_yslab, r, nzp, z0, n);
                continue;
            }
            float2* send = k.recv;
            run_slab_xy_forward(p, real, k.zslab, send, k.nzl, nyl, k.st, nullptr, r, nzp, z0, n);
            FC_CUDA(cudaEventRecord(k.ev_chunk[8 + ch], k.st));
            for (int xs = 0; xs < nxs; ++xs) FC_CUDA(cudaStreamWaitEvent(k.s_x[xs], k.ev_chunk[8 + ch], 0));
            const size_t rowblk = (size_t)nyl * xcp;
            static const int debug_nocopy = [] {   // timing experiments only (results are wrong): skip the exchange copies
                const char* e = std::getenv("FCB200_SLAB_DEBUG_NOCOPY");
                return e ? std::atoi(e) : 0;
            }();
            for (int i = 1; i <= P && !debug_nocopy; ++i) {   // start with the next rank so that the peers are hit evenly
                const int q = (r + i) % P;
                FC_CUDA(cudaMemcpyAsync(ranks[(size_t)q].yslab + ((size_t)r * nzp + z0) * rowblk,
                                        send + ((size_t)q * nzp + z0) * rowblk, (size_t)n * rowblk * sizeof(float2),
                                        cudaMemcpyDefault, k.s_x[i % nxs]));
            }
        }
        if (copy_x) {
            for (int xs = 0; xs < nxs; ++xs) {
                FC_CUDA(cudaEventRecord(k.ev_x[xs], k.s_x[xs]));
                FC_CUDA(cudaStreamWaitEvent(k.st, k.ev_x[xs], 0));
            }
        }
        FC_CUDA(cudaEventRecord(k.ev_fwd, k.st));
        FC_CUDA(cudaEventRecord(k.ev_t[1], k.st));
    });
    barrier->wait();   // every rank's ev_fwd has been recorded

    // ---- phase B: fused z pass on my ky rows, planes stored straight into their owners' receive buffers
    phase(r, [&] {
        for (SlabRank& q : ranks) FC_CUDA(cudaStreamWaitEvent(k.st, q.ev_fwd, 0));
        if (!c.psf_cached) FC_CUDA(cudaStreamWaitEvent(k.st, k.ev_psf, 0));
        run_slab_z_fused(p, k.yslab, k.H, nyl, k.st, k.d_peer_recv, r, nzp, pull_x ? k.d_peer_send : nullptr);
        FC_CUDA(cudaEventRecord(k.ev_z, k.st));
        FC_CUDA(cudaEventRecord(k.ev_t[2], k.st));
    });
    barrier->wait();

    // ---- phase C: y + x inverse on my planes, download
    phase(r, [&] {
        for (SlabRank& q : ranks) FC_CUDA(cudaStreamWaitEvent(k.st, q.ev_z, 0));
        for (int ch = 0; ch < nch; ++ch) {
            const int z0 = z0_of(ch), n = z0_of(ch + 1) - z0;
            if (n <= 0) continue;
            run_slab_yx_inverse(p, k.recv, k.zslab, real, k.nzl, nyl, k.st, nzp, z0, n);
            if (pinned) {
                FC_CUDA(cudaEventRecord(k.ev_chunk[8 + ch], k.st));
                FC_CUDA(cudaStreamWaitEvent(k.s_d2h, k.ev_chunk[8 + ch], 0));
                FC_CUDA(cudaMemcpyAsync(h_slab + z0 * rplane, real + z0 * rplane, n * rplane * sizeof(float),
                                        cudaMemcpyDeviceToHost, k.s_d2h));
            }
        }
        FC_CUDA(cudaEventRecord(k.ev_t[3], k.st));
        if (c.im && !pinned) k.stager->download(h_slab, real, slab_bytes, k.st);
        if (pinned) {
            FC_CUDA(cudaEventRecord(k.ev_d2h, k.s_d2h));
            FC_CUDA(cudaStreamWaitEvent(k.st, k.ev_d2h, 0));
        }
        FC_CUDA(cudaEventRecord(k.ev_done, k.st));
        k.timed = true;
    });
    // the call is synchronous: the caller reads its buffers (and may free the kernel) on return
    cudaStreamSynchronize(k.s_psf);
    if (cudaStreamSynchronize(k.st) != cudaSuccess && !failed.load()) {
        errors[(size_t)r] = std::string("fcb200: slab rank failed: ") + cudaGetErrorString(cudaGetLastError());
        failed.store(true);
    }
}

void SlabContext::worker(int r)
{
    cudaSetDevice(ranks[(size_t)r].dev);
    unsigned long long seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lock(mu);
            cv_job.wait(lock, [&] { return stop || job_gen != seen; });
            if (stop) return;
            seen = job_gen;
        }
        run_call(r);
        {
            std::lock_guard<std::mutex> lock(mu);
            ++done;
        }
        cv_done.notify_one();
    }
}

SlabContext::~SlabContext()
{
    {
        std::lock_guard<std::mutex> lock(mu);
        stop = true;
    }
    cv_job.notify_all();
    for (std::thread& t : workers) t.join();
    int prev = -1;
    cudaGetDevice(&prev);
    for (SlabRank& k : ranks) {
        cudaSetDevice(k.dev);
        if (k.st) cudaStreamSynchronize(k.st);
        cudaFree(k.real);
        cudaFree(k.zslab);
        cudaFree(k.recv);
        cudaFree(k.yslab);
        cudaFree(k.H);
        cudaFree(k.scratch);
        cudaFree(k.d_kernel);
        cudaFree(k.d_peer_yslab);
        cudaFree(k.d_peer_recv);
        cudaFree(k.send);
        cudaFree(k.d_peer_send);
        k.stager.reset();
        for (cudaEvent_t e : {k.ev_fwd, k.ev_z, k.ev_done, k.ev_psf, k.ev_d2h, k.ev_x[0], k.ev_x[1], k.ev_x[2], k.ev_x[3], k.ev_t[0],
                              k.ev_t[1], k.ev_t[2], k.ev_t[3]})
            if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : k.ev_chunk)
            if (e) cudaEventDestroy(e);
        for (cudaStream_t s : {k.st, k.s_psf, k.s_h2d, k.s_d2h, k.s_x[0], k.s_x[1], k.s_x[2], k.s_x[3]})
            if (s) cudaStreamDestroy(s);
    }
    if (prev >= 0) cudaSetDevice(prev);
}

void enable_peer_access(const std::vector<int>& devs)
{
    std::vector<int> uniq(devs);
    std::sort(uniq.begin(), uniq.end());
    uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
    for (int a : uniq) {
        FC_CUDA(cudaSetDevice(a));
        for (int b : uniq) {
            if (a == b) continue;
            int can = 0;
            FC_CUDA(cudaDeviceCanAccessPeer(&can, a, b));
            if (!can) throw std::runtime_error("fcb200: slab mode needs peer access between all devices of the call");
            const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else FC_CUDA(e);
        }
    }
}

struct SlabKey {
    std::vector<int> v;   // d0, d1, d2, devs...
    bool operator<(const SlabKey& o) const { return v < o.v; }
};
std::mutex g_slab_mu;
std::map<SlabKey, std::shared_ptr<SlabContext>> g_slab_cache;
unsigned long long g_slab_tick = 0;

std::shared_ptr<SlabContext> build_context(const int* imDim, const int* devs, int ndev)
{
    auto ctx = std::make_shared<SlabContext>();
    SlabContext& c = *ctx;
    c.devs.assign(devs, devs + ndev);
    std::memcpy(c.dims, imDim, sizeof(int) * 3);
    c.P = ndev;
    c.nzp = (imDim[2] + ndev - 1) / ndev;
    c.nyl = (imDim[1] + ndev - 1) / ndev;
    if ((long long)(ndev - 1) * c.nzp >= imDim[2] || (long long)(ndev - 1) * c.nyl >= imDim[1])
        throw std::runtime_error("fcb200: slab mode: every rank must own at least one z plane and one ky row");
    c.xcp = make_geometry(imDim[0], 1, 1).xcp;
    c.errors.assign((size_t)ndev, std::string());
    c.barrier.reset(new ThreadBarrier(ndev));
    enable_peer_access(c.devs);
    c.ranks.resize((size_t)ndev);
    const size_t spec_bytes = c.slab_spec_elems() * sizeof(float2);
    for (int r = 0; r < ndev; ++r) {
        SlabRank& k = c.ranks[(size_t)r];
        k.dev = devs[r];
        k.rank = r;
        k.nzl = std::min(c.nzp, imDim[2] - r * c.nzp);
        k.ny_here = std::min(c.nyl, imDim[1] - r * c.nyl);
        FC_CUDA(cudaSetDevice(k.dev));
        k.plan = get_plan(k.dev, imDim[0], imDim[1], imDim[2], false);
        for (cudaStream_t* s : {&k.st, &k.s_psf, &k.s_h2d, &k.s_d2h, &k.s_x[0], &k.s_x[1], &k.s_x[2], &k.s_x[3]})
            FC_CUDA(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking));
        for (float2** b : {&k.zslab, &k.recv, &k.yslab, &k.H}) {
            FC_CUDA(cudaMalloc(b, spec_bytes));
            FC_CUDA(cudaMemset(*b, 0, spec_bytes));   // ragged slabs leave the pad rows / planes of a block unwritten
        }
        for (cudaEvent_t* e : {&k.ev_fwd, &k.ev_z, &k.ev_done, &k.ev_psf, &k.ev_d2h, &k.ev_x[0], &k.ev_x[1], &k.ev_x[2], &k.ev_x[3]})
            FC_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        k.stager.reset(new HostStager());
        for (cudaEvent_t& e : k.ev_t) FC_CUDA(cudaEventCreate(&e));
        for (cudaEvent_t& e : k.ev_chunk) FC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        FC_CUDA(cudaEventRecord(k.ev_done, k.st));   // "previous call finished" for the first call
    }
    // pointer tables: entry q = rank q's buffer (peer-mapped through unified addressing)
    std::vector<float2*> ys((size_t)ndev), rv((size_t)ndev);
    for (int q = 0; q < ndev; ++q) {
        ys[(size_t)q] = c.ranks[(size_t)q].yslab;
        rv[(size_t)q] = c.ranks[(size_t)q].recv;
    }
    for (SlabRank& k : c.ranks) {
        FC_CUDA(cudaSetDevice(k.dev));
        FC_CUDA(cudaMalloc(&k.d_peer_yslab, sizeof(float2*) * ndev));
        FC_CUDA(cudaMalloc(&k.d_peer_recv, sizeof(float2*) * ndev));
        FC_CUDA(cudaMemcpy(k.d_peer_yslab, ys.data(), sizeof(float2*) * ndev, cudaMemcpyHostToDevice));
        FC_CUDA(cudaMemcpy(k.d_peer_recv, rv.data(), sizeof(float2*) * ndev, cudaMemcpyHostToDevice));
        FC_CUDA(cudaDeviceSynchronize());
    }
    for (int r = 0; r < ndev; ++r) c.workers.emplace_back([ctx_raw = ctx.get(), r] { ctx_raw->worker(r); });
    return ctx;
}

std::shared_ptr<SlabContext> get_context(const int* imDim, const int* devs, int ndev, bool create)
{
    std::lock_guard<std::mutex> lock(g_slab_mu);
    SlabKey key;
    key.v.assign(imDim, imDim + 3);
    key.v.insert(key.v.end(), devs, devs + ndev);
    auto it = g_slab_cache.find(key);
    if (it != g_slab_cache.end()) {
        it->second->last_use = ++g_slab_tick;
        return it->second;
    }
    if (!create) return nullptr;
    // slab contexts hold several volume-sized buffers per device: keep one shape at a time
    for (auto i = g_slab_cache.begin(); i != g_slab_cache.end();)
        i = (i->second.use_count() == 1) ? g_slab_cache.erase(i) : std::next(i);
    auto ctx = build_context(imDim, devs, ndev);
    ctx->last_use = ++g_slab_tick;
    g_slab_cache[key] = ctx;
    return ctx;
}

}  // namespace

void slab_convolve(float* im, float* const* slabs, const int* imDim, const float* kernel, const int* kernelDim,
                   const int* devs, int ndev)
{
    if (ndev < 1 || !devs) throw std::runtime_error("fcb200: slab mode needs a device list");
    if ((im == nullptr) == (slabs == nullptr)) throw std::runtime_error("fcb200: slab mode takes one volume or one slab per device");
    for (int i = 0; i < 3; ++i)
        if (kernelDim[i] > imDim[i]) throw std::runtime_error("fcb200: kernel larger than image");
    struct Restore {
        int dev;
        ~Restore() { cudaSetDevice(dev); }
    } restore{devs[0]};   // like the reference, the (first) device of the call stays current
    auto ctx = get_context(imDim, devs, ndev, true);
    SlabContext& c = *ctx;
    std::lock_guard<std::mutex> call_lock(c.call_mu);

    SlabCall call;
    call.im = im;
    call.slabs = slabs;
    call.kernel = kernel;
    const int pdims[6] = {kernelDim[0], kernelDim[1], kernelDim[2], imDim[0], imDim[1], imDim[2]};
    std::memcpy(call.pdims, pdims, sizeof(pdims));
    const size_t ktaps = (size_t)kernelDim[0] * kernelDim[1] * kernelDim[2];
    const size_t rplane = (size_t)imDim[1] * imDim[0];
    if (im) {
        call.im_kind = classify_pointer(im, devs[0]);
        if (call.im_kind == HostMem::Device)
            throw std::runtime_error("fcb200: slab mode takes a HOST volume, or one device slab per rank (fcb200_convolve_slab_device)");
    }
    // kernel kind and the PSF-spectrum cache (host-pointer kernels only, like the single-device path)
    cudaPointerAttributes attr{};
    bool k_dev = false;
    if (cudaPointerGetAttributes(&attr, kernel) == cudaSuccess) k_dev = attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
    else cudaGetLastError();
    static const bool cache_on = env_flag("FCB200_PSF_CACHE", true);
    {
        // forward exchange: 1 (default) copy engines, 0 peer stores from the y pass, 2 pulled by the fused z pass
        const char* e = std::getenv("FCB200_SLAB_EXCHANGE");
        call.exchange = e ? std::atoi(e) : 1;
        if (call.exchange == 2 && ndev > 1) {   // the pull form needs one more exchange-sized buffer per rank and the peer tables
            std::vector<float2*> blocks((size_t)ndev);
            const size_t spec_bytes = c.slab_spec_elems() * sizeof(float2);
            for (SlabRank& k : c.ranks)
                if (!k.send) {
                    FC_CUDA(cudaSetDevice(k.dev));
                    FC_CUDA(cudaMalloc(&k.send, spec_bytes));
                    FC_CUDA(cudaMemset(k.send, 0, spec_bytes));
                }
            for (SlabRank& k : c.ranks)
                if (!k.d_peer_send) {
                    FC_CUDA(cudaSetDevice(k.dev));
                    for (int q = 0; q < ndev; ++q)
                        blocks[(size_t)q] = c.ranks[(size_t)q].send + (size_t)k.rank * c.nzp * c.nyl * c.xcp;
                    FC_CUDA(cudaMalloc(&k.d_peer_send, sizeof(float2*) * ndev));
                    FC_CUDA(cudaMemcpy(k.d_peer_send, blocks.data(), sizeof(float2*) * ndev, cudaMemcpyHostToDevice));
                }
        }
    }
    call.psf_cached = !k_dev && cache_on && c.h_valid && std::memcmp(c.h_dims, pdims, sizeof(pdims)) == 0 &&
                      c.h_taps.size() == ktaps && std::memcmp(c.h_taps.data(), kernel, ktaps * sizeof(float)) == 0;
    c.h_valid = false;

    // per-rank buffers that depend on the call (kernel taps, PSF scratch, the slab itself for host volumes)
    for (SlabRank& k : c.ranks) {
        FC_CUDA(cudaSetDevice(k.dev));
        if (slabs) {
            if (classify_pointer(slabs[k.rank], k.dev) != HostMem::Device)
                throw std::runtime_error("fcb200: slabs[r] must be a device pointer on devs[r]");
            check_real_alignment(slabs[k.rank], imDim[0]);
        } else {
            if (!k.real) FC_CUDA(cudaMalloc(&k.real, (size_t)k.nzl * rplane * sizeof(float)));
            if (call.im_kind == HostMem::Pageable) k.stager->prepare();
        }
        if (!call.psf_cached) {
            if (ktaps > k.kernel_cap) {
                float* nk = nullptr;
                FC_CUDA(cudaMalloc(&nk, ktaps * sizeof(float)));
                FC_CUDA(cudaStreamSynchronize(k.s_psf));
                cudaFree(k.d_kernel);
                k.d_kernel = nk;
                k.kernel_cap = ktaps;
            }
            if (std::memcmp(c.scratch_key, pdims, sizeof(pdims)) != 0) {
                c.scratch_need = std::max<size_t>(1, psf_slab_scratch_elems(*k.plan, pdims));
                std::memcpy(c.scratch_key, pdims, sizeof(pdims));
            }
            const size_t need = c.scratch_need;
            if (need > k.scratch_cap) {
                float2* ns = nullptr;
                FC_CUDA(cudaMalloc(&ns, need * sizeof(float2)));
                FC_CUDA(cudaStreamSynchronize(k.s_psf));
                cudaFree(k.scratch);
                k.scratch = ns;
                k.scratch_cap = need;
            }
        }
    }

    // hand the call to the workers and wait for all of them
    c.failed.store(false);
    for (std::string& e : c.errors) e.clear();
    {
        std::lock_guard<std::mutex> lock(c.mu);
        c.call = call;
        c.done = 0;
        ++c.job_gen;
    }
    c.cv_job.notify_all();
    {
        std::unique_lock<std::mutex> lock(c.mu);
        c.cv_done.wait(lock, [&] { return c.done == c.P; });
    }
    if (c.failed.load()) {
        for (const std::string& e : c.errors)
            if (!e.empty()) throw std::runtime_error(e);
        throw std::runtime_error("fcb200: slab call failed");
    }
    if (!k_dev && cache_on) {
        c.h_taps.assign(kernel, kernel + ktaps);
        std::memcpy(c.h_dims, pdims, sizeof(pdims));
        c.h_valid = true;
    }
}

int slab_last_timing(const int* imDim, const int* devs, int ndev, float* ms, int cap)
{
    auto ctx = get_context(imDim, devs, ndev, false);
    if (!ctx) throw std::runtime_error("fcb200: no slab call has run for this shape and device list");
    std::lock_guard<std::mutex> call_lock(ctx->call_mu);
    int n = 0;
    for (SlabRank& k : ctx->ranks) {
        if (4 * (n + 1) > cap) break;
        float* o = ms + 4 * n;
        o[0] = o[1] = o[2] = o[3] = 0.f;
        if (k.timed) {
            FC_CUDA(cudaSetDevice(k.dev));
            FC_CUDA(cudaEventSynchronize(k.ev_t[3]));
            FC_CUDA(cudaEventElapsedTime(&o[0], k.ev_t[0], k.ev_t[1]));
            FC_CUDA(cudaEventElapsedTime(&o[1], k.ev_t[1], k.ev_t[2]));
            FC_CUDA(cudaEventElapsedTime(&o[2], k.ev_t[2], k.ev_t[3]));
            FC_CUDA(cudaEventElapsedTime(&o[3], k.ev_t[0], k.ev_t[3]));
        }
        ++n;
    }
    if (ndev > 0) cudaSetDevice(devs[0]);
    return n;
}

void batch_multi(float* const* ims, int n, const int* imDim, const float* kernel, const int* kernelDim, const int* devs,
                 int ndev, int* blocks_per_dev)
{
    if (ndev < 1 || !devs) throw std::runtime_error("fcb200: batch needs a device list");
    if (blocks_per_dev) std::fill(blocks_per_dev, blocks_per_dev + ndev, 0);
    if (n <= 0) return;
    const int pdims[6] = {kernelDim[0], kernelDim[1], kernelDim[2], imDim[0], imDim[1], imDim[2]};
    const BatchKinds kinds = classify_batch(ims, n, devs[0]);
    if (kinds.any_device && ndev > 1) throw std::runtime_error("fcb200: a multi-device batch takes host blocks");
    std::atomic<int> counter{0};
    std::vector<std::string> errors((size_t)ndev);
    std::vector<int> taken((size_t)ndev, 0);
    std::atomic<bool> failed{false};
    auto run = [&](int r) {
        try {
            batch_core(ims,
                       [&, r] {
                           if (failed.load()) return -1;
                           const int b = counter.fetch_add(1);
                           if (b >= n) return -1;
                           ++taken[(size_t)r];
                           return b;
                       },
                       kinds, imDim[0], imDim[1], imDim[2], kernel, pdims, devs[r], false, nullptr);
        } catch (const std::exception& e) {
            errors[(size_t)r] = e.what();
            failed.store(true);
            cudaGetLastError();
        }
    };
    std::vector<std::thread> threads;
    for (int r = 1; r < ndev; ++r) threads.emplace_back(run, r);
    run(0);   // the calling thread drives the first device (which stays current, like devCUDA in the reference)
    for (std::thread& t : threads) t.join();
    cudaSetDevice(devs[0]);
    if (blocks_per_dev) std::copy(taken.begin(), taken.end(), blocks_per_dev);
    for (const std::string& e : errors)
        if (!e.empty()) throw std::runtime_error(e);
}

std::vector<int> slab_devices_for(const int* imDim, int devCUDA, bool host_pointer)
{
    std::vector<int> out;
    if (!host_pointer) return out;
    const char* mode = std::getenv("FCB200_SLAB");   // 0: never, 1: whenever >= 2 devices can be used, unset: when needed
    if (mode && std::atoi(mode) == 0) return out;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return out;
    }
    const char* list = std::getenv("FCB200_SLAB_DEVICES");   // explicit ranks (a device may repeat: emulated ranks)
    if (!list && ndev < 2) return out;
    if (!(mode && std::atoi(mode) == 1)) {
        // does the single-device path fit?  real volume + image spectrum + PSF spectrum (or its window) + slack
        const Geometry g = make_geometry(imDim[0], imDim[1], imDim[2]);
        const double need = 4.0 * g.nx * g.ny * g.nz + 2.0 * 8.0 * g.xcp * g.ny * g.nz + 512e6;
        size_t free_b = 0, total_b = 0;
        cudaSetDevice(devCUDA);
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
            cudaGetLastError();
            return out;
        }
        if (need <= (double)free_b) return out;
    }
    std::vector<int> cand;
    if (list) {
        for (const char* s = list; *s;) {
            char* end = nullptr;
            const long v = std::strtol(s, &end, 10);
            if (end == s) break;
            if (v >= 0 && v < ndev) cand.push_back((int)v);
            s = (*end == ',') ? end + 1 : end;
        }
    } else {
        cand.push_back(devCUDA);
        for (int d = 0; d < ndev; ++d) {
            if (d == devCUDA) continue;
            int ab = 0, ba = 0;
            cudaDeviceCanAccessPeer(&ab, devCUDA, d);
            cudaDeviceCanAccessPeer(&ba, d, devCUDA);
            if (ab && ba) cand.push_back(d);
        }
    }
    // every rank must own at least one z plane and one ky row
    int P = (int)cand.size();
    while (P >= 2) {
        const int nzp = (imDim[2] + P - 1) / P, nyl = (imDim[1] + P - 1) / P;
        if ((long long)(P - 1) * nzp < imDim[2] && (long long)(P - 1) * nyl < imDim[1]) break;
        --P;
    }
    if (P >= 2) out.assign(cand.begin(), cand.begin() + P);
    return out;
}

void release_multi()
{
    std::lock_guard<std::mutex> lock(g_slab_mu);
    g_slab_cache.clear();
}

}  // namespace fcb200
