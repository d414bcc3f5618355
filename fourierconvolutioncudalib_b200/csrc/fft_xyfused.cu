// x+y passes of one z plane fused in ONE kernel through distributed shared memory (thread-block clusters).
//
// A cluster of 8 CTAs owns a plane of 512 x 512 voxels.  Forward:
//   phase A  CTA c runs the register-resident row-wise R2C transform (fft_xrow.cuh) on its 64 rows and keeps
//            the 64 x 260 spectrum rows in ITS shared memory (133 KB) instead of writing them to HBM;
//   cluster barrier;
//   phase B  the y pass on column tiles of 16 kx bins: the first radix-8 stage needs rows j + 64 k, k = 0..7,
//            i.e. exactly ONE row from each CTA of the cluster -- read straight from the peers' shared memory
//            (DSMEM) into registers; the remaining stages and the store to HBM are the ordinary static stages.
// The inverse mirrors it: y inverse from HBM, last stage scattered into the peers' slabs, cluster barrier,
// row-wise C2R from the local slab.  The intermediate spectrum (8 Nc bytes written + 8 Nc read per direction)
// never touches HBM: 16 of the 44.5 bytes per voxel of the image path disappear.
// Shapes: nx = 512, ny = 512 (the bench workload and every 512 x 512 x nz volume); anything else uses the
// separate passes.
//
// STATUS (round 1): correct (parity tests pass with FCB200_XY_CLUSTER=1) but NOT faster, hence opt-in:
// C3 forward 0.375 ms / inverse 0.434 ms against 0.222 ms for the two separate passes.  The 133 KB slab leaves
// room for one 256-thread CTA per SM, so nothing hides the HBM / DSMEM latencies of the strictly sequential
// phases (about 23 us per plane and cluster where ~9 us would be needed).  See profiles/r01_notes.md.
#include <cooperative_groups.h>

#include "fft_static_plans.h"
#include "fft_xyfused.h"

namespace cg = cooperative_groups;

namespace fcb200 {

namespace {

constexpr int kR = 16, kM = 256, kPadM = kM + kR;    // x transform: M = 16 * 16 complex points
constexpr int kNX = 512, kNY = 512, kXCP = 260;      // plane geometry
constexpr int kRows = 64;                            // rows per CTA (8 CTAs per plane)
constexpr int kL = 512;                              // y transform (8, 8, 8)
constexpr int kThreads = 256;
constexpr int kNW = kThreads / 8;
constexpr int kGroups = kThreads / kR;               // row pairs in flight per CTA in the x phase

constexpr size_t kSlabF4 = (size_t)kRows * kXCP / 2;            // float4 per slab
constexpr size_t kWorkF4 = (size_t)kGroups * kPadM;             // >= kL * 8 (y tile)
static_assert(kWorkF4 >= (size_t)kL * 8, "work buffer must hold one y tile");
constexpr size_t kSmemBytes = (kSlabF4 + kR * kR + kL + kWorkF4) * sizeof(float4);

__device__ __forceinline__ int xp(int p) { return p + p / kR; }

// ---- x forward on one row pair (rows a, b): registers -> slab rows --------------------------------------
__device__ __forceinline__ void xrow_fwd_pair(const float2* __restrict__ srcA, const float2* __restrict__ srcB,
                                              float2* __restrict__ dstA, float2* __restrict__ dstB,
                                              const float4* __restrict__ twt, float4* __restrict__ x,
                                              const float2* __restrict__ twx, int j, bool odd)
{
    constexpr int R = kR, M = kM;
    p2 r[R], i[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const float2 ua = __ldg(srcA + j + R * k);
        const float2 ub = __ldg(srcB + j + R * k);
        r[k] = make_float2(ua.x, ub.x);
        i[k] = make_float2(ua.y, ub.y);
    }
    Dft<R>::run(r, i);
#pragma unroll
    for (int m = 1; m < R; ++m) cmul(r[m], i[m], twt[m * R + j]);
#pragma unroll
    for (int m = 0; m < R; ++m) x[j + (R + 1) * m] = make_float4(r[m].x, r[m].y, i[m].x, i[m].y);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const float4 v = x[(R + 1) * j + k];
        r[k] = make_float2(v.x, v.y);
        i[k] = make_float2(v.z, v.w);
    }
    Dft<R>::run(r, i);   // Z[j + R*m]
    const p2 nyq = psub(r[0], i[0]);
    __syncwarp();
#pragma unroll
    for (int m = 0; m < R; ++m) x[j + (R + 1) * m] = make_float4(r[m].x, r[m].y, i[m].x, i[m].y);
    __syncwarp();
#pragma unroll
    for (int m = 0; m < R; ++m) {
        const int k = j + R * m;
        const float4 pb = x[xp((M - k) & (M - 1))];
        const p2 pr = make_float2(pb.x, pb.y), pi = make_float2(pb.z, pb.w);
        const float2 tk = __ldg(twx + k);
        const p2 er = pmuls(padd(r[m], pr), 0.5f), ei = pmuls(psub(i[m], pi), 0.5f);
        const p2 orr = pmuls(padd(i[m], pi), 0.5f), oi = pmuls(psub(r[m], pr), -0.5f);
        r[m] = padd(er, pfmas(orr, tk.x, pmuls(oi, -tk.y)));
        i[m] = padd(ei, pfmas(oi, tk.x, pmuls(orr, tk.y)));
    }
#pragma unroll
    for (int m = 0; m < R; ++m) {
        const p2 send = odd ? r[m] : i[m];
        p2 got;
        got.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
        got.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
        dstA[j + R * m] = odd ? make_float2(got.x, i[m].x) : make_float2(r[m].x, got.x);
        dstB[j + R * m] = odd ? make_float2(got.y, i[m].y) : make_float2(r[m].y, got.y);
    }
    if (j == 0) {
        dstA[M] = make_float2(nyq.x, 0.f);
        dstB[M] = make_float2(nyq.y, 0.f);
#pragma unroll
        for (int k = M + 1; k < kXCP; ++k) {
            dstA[k] = make_float2(0.f, 0.f);
            dstB[k] = make_float2(0.f, 0.f);
        }
    }
    __syncwarp();   // the exchange buffer is reused by the next row pair of this group
}

// ---- x inverse on one row pair: slab rows -> registers -> real rows ---------------------------------------
__device__ __forceinline__ void xrow_inv_pair(const float2* __restrict__ srcA, const float2* __restrict__ srcB,
                                              float2* __restrict__ dstA, float2* __restrict__ dstB,
                                              const float4* __restrict__ twt, float4* __restrict__ x,
                                              const float2* __restrict__ twx, int j, bool odd)
{
    constexpr int R = kR, M = kM;
    p2 r[R], i[R];
#pragma unroll
    for (int m = 0; m < R; ++m) {
        const float2 va = srcA[j + R * m];
        const float2 vb = srcB[j + R * m];
        const p2 send = odd ? make_float2(va.x, vb.x) : make_float2(va.y, vb.y);
        p2 got;
        got.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
        got.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
        r[m] = odd ? got : make_float2(va.x, vb.x);
        i[m] = odd ? make_float2(va.y, vb.y) : got;
    }
#pragma unroll
    for (int m = 0; m < R; ++m) x[j + (R + 1) * m] = make_float4(r[m].x, r[m].y, i[m].x, i[m].y);
    __syncwarp();
    const p2 x0r = r[0];
#pragma unroll
    for (int m = 0; m < R; ++m) {
        const int k = j + R * m;
        const float4 pb = x[xp((M - k) & (M - 1))];
        const p2 pr = make_float2(pb.x, pb.y), pi = make_float2(pb.z, pb.w);
        const float2 tk = __ldg(twx + k);
        const p2 sr = padd(r[m], pr), si = psub(i[m], pi);
        const p2 Dr = psub(r[m], pr), Di = padd(i[m], pi);
        const p2 dr = pfmas(Dr, tk.x, pmuls(Di, tk.y));
        const p2 di = pfmas(Di, tk.x, pmuls(Dr, -tk.y));
        r[m] = psub(sr, di);
        i[m] = padd(si, dr);
    }
    __syncwarp();
    if (j == 0) {
        const p2 xm = make_float2(srcA[M].x, srcB[M].x);
        r[0] = padd(x0r, xm);
        i[0] = psub(x0r, xm);
    }
    Dft<R>::run(i, r);
#pragma unroll
    for (int k = 0; k < R; ++k) x[(R + 1) * j + k] = make_float4(r[k].x, r[k].y, i[k].x, i[k].y);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const float4 v = x[j + (R + 1) * k];
        r[k] = make_float2(v.x, v.y);
        i[k] = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int k = 1; k < R; ++k) cmulc(r[k], i[k], twt[k * R + j]);
    Dft<R>::run(i, r);
#pragma unroll
    for (int m = 0; m < R; ++m) {
        dstA[j + R * m] = make_float2(r[m].x, i[m].x);
        dstB[j + R * m] = make_float2(r[m].y, i[m].y);
    }
    __syncwarp();
}

struct Smem {
    float2* slab;
    float4* twt;
    float4* twy;
    float4* work;
};

__device__ __forceinline__ Smem carve(float4* smem)
{
    Smem s;
    s.slab = reinterpret_cast<float2*>(smem);
    s.twt = smem + kSlabF4;
    s.twy = s.twt + kR * kR;
    s.work = s.twy + kL;
    return s;
}

__device__ __forceinline__ void load_tables(const Smem& s, const XYArgs& a)
{
    for (int idx = threadIdx.x; idx < kR * kR; idx += kThreads) {
        const int m = idx / kR, jj = idx % kR;
        const float2 w = __ldg(a.Px.tw + jj * m);
        s.twt[idx] = make_float4(w.x, w.x, w.y, w.y);
    }
    load_twiddles(s.twy, a.Py.tw, kL);
}

// ------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(kThreads, 1) xy_fwd_cluster_kernel(XYArgs a)
{
    extern __shared__ float4 smem[];
    const Smem s = carve(smem);
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const int cid = blockIdx.x / 8, ncl = gridDim.x / 8;
    const int t = threadIdx.x;
    load_tables(s, a);
    const float2* peer[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) peer[k] = cluster.map_shared_rank(s.slab, k);
    __syncthreads();

    const int grp = t / kR, j = t % kR, cp = t & 7, w = t >> 3;
    const bool odd = t & 1;
    for (int plane = cid; plane < a.nplanes; plane += ncl) {
        // ---- phase A: 64 rows, row-wise R2C into the local slab
        const float* rbase = a.in_real + ((size_t)plane * kNY + (size_t)c * kRows) * kNX;
#pragma unroll 1
        for (int it = 0; it < kRows / (2 * kGroups); ++it) {
            const int rl = 2 * (it * kGroups + grp);
            xrow_fwd_pair(reinterpret_cast<const float2*>(rbase + (size_t)rl * kNX),
                          reinterpret_cast<const float2*>(rbase + (size_t)(rl + 1) * kNX), s.slab + (size_t)rl * kXCP,
                          s.slab + (size_t)(rl + 1) * kXCP, s.twt, s.work + (size_t)grp * kPadM, a.twx, j, odd);
        }
        cluster.sync();   // every slab of the plane is complete (and every CTA has left the x exchange buffers)

        // ---- phase B: y pass on my column tiles, first stage straight from the peers' slabs
        float2* gplane = a.spec + (size_t)plane * kNY * kXCP;
#pragma unroll 1
        for (int tile = c; tile * 16 < kXCP; tile += 8) {
            const int col0 = tile * 16;
            const bool active = cp < min(8, (kXCP - col0) >> 1);
            if (active) {
                constexpr int S = kL / 8;
                float4 v[2][8];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int jb = w + u * kNW;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        v[u][k] = *reinterpret_cast<const float4*>(peer[k] + (size_t)jb * kXCP + col0 + 2 * cp);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int jb = w + u * kNW;
                    p2 r[8], i[8];
                    ssplit<8>(v[u], r, i);
                    Dft<8>::run(r, i);
#pragma unroll
                    for (int m = 1; m < 8; ++m) cmul(r[m], i[m], s.twy[jb * m]);
                    store_pairs<8>(s.work, jb * 8 + cp, S * 8, r, i);
                }
            }
            __syncthreads();
            if (active) sstage<8, kL, kL / 8, kNW, false, 8>(s.work, s.twy, cp, w);
            __syncthreads();
            if (active) slast_fwd<8, kL, kNW, 8>(gplane + col0 + 2 * cp, (size_t)kXCP, s.work, a.Py.rev, cp, w);
            __syncthreads();
        }
        cluster.sync();   // nobody reads my slab any more: the next plane may overwrite it
    }
}

__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(kThreads, 1) yx_inv_cluster_kernel(XYArgs a)
{
    extern __shared__ float4 smem[];
    const Smem s = carve(smem);
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const int cid = blockIdx.x / 8, ncl = gridDim.x / 8;
    const int t = threadIdx.x;
    load_tables(s, a);
    float2* peer[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) peer[k] = cluster.map_shared_rank(s.slab, k);
    __syncthreads();

    const int grp = t / kR, j = t % kR, cp = t & 7, w = t >> 3;
    const bool odd = t & 1;
    for (int plane = cid; plane < a.nplanes; plane += ncl) {
        // ---- phase B': y inverse on my column tiles (forward structure with re/im exchanged); the last stage
        // ---- scatters row r0 + 64 m into the slab of CTA m
        const float2* gplane = a.spec + (size_t)plane * kNY * kXCP;
#pragma unroll 1
        for (int tile = c; tile * 16 < kXCP; tile += 8) {
            const int col0 = tile * 16;
            const bool active = cp < min(8, (kXCP - col0) >> 1);
            if (active)
                sfirst_fwd<8, kL, kNW, 2, false, 8, true>(gplane + col0 + 2 * cp, (size_t)kXCP, s.work, s.twy, cp, w, nullptr);
            __syncthreads();
            if (active) sstage<8, kL, kL / 8, kNW, false, 8>(s.work, s.twy, cp, w);
            __syncthreads();
            if (active) {
                constexpr int nb = kL / 8;
#pragma unroll
                for (int it = 0; it < nb / kNW; ++it) {
                    const int b = w + it * kNW;
                    p2 r[8], i[8];
                    load_pairs<8>(s.work, b * 8 * 8 + cp, 8, r, i);
                    const int r0 = __ldg(a.Py.rev + b * 8);   // < 64: local row in the owner's slab
                    Dft<8>::run(r, i);
#pragma unroll
                    for (int m = 0; m < 8; ++m)
                        *reinterpret_cast<float4*>(peer[m] + (size_t)r0 * kXCP + col0 + 2 * cp) =
                            make_float4(i[m].x, i[m].y, r[m].x, r[m].y);
                }
            }
            __syncthreads();
        }
        cluster.sync();   // all rows of the plane have landed in the slabs

        // ---- phase A': row-wise C2R from the local slab
        float* rbase = a.out_real + ((size_t)plane * kNY + (size_t)c * kRows) * kNX;
#pragma unroll 1
        for (int it = 0; it < kRows / (2 * kGroups); ++it) {
            const int rl = 2 * (it * kGroups + grp);
            xrow_inv_pair(s.slab + (size_t)rl * kXCP, s.slab + (size_t)(rl + 1) * kXCP,
                          reinterpret_cast<float2*>(rbase + (size_t)rl * kNX),
                          reinterpret_cast<float2*>(rbase + (size_t)(rl + 1) * kNX), s.twt, s.work + (size_t)grp * kPadM,
                          a.twx, j, odd);
        }
        cluster.sync();   // my slab has been consumed: peers may write the next plane into it
    }
}

template <int WHICH, typename K>
int cluster_grid(K kernel, int nplanes)
{
    static int max_clusters = [&] {
        FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(8 * 64, 1, 1);
        cfg.blockDim = dim3(kThreads, 1, 1);
        cfg.dynamicSmemBytes = kSmemBytes;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 8;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) {
            cudaGetLastError();
            n = 0;
        }
        return n;
    }();
    return std::min(max_clusters, nplanes);
}

bool shape_ok(const XYArgs& a)
{
    static const bool on = env_int("FCB200_XY_CLUSTER", 0) != 0;
    return on && static_enabled() && a.g.nx == kNX && a.g.ny == kNY && a.g.xcp == kXCP && plan_matches<P256>(a.Px) &&
           plan_matches<P512>(a.Py) && a.nplanes > 0;
}

}  // namespace

bool launch_xy_fwd_cluster(const XYArgs& a, cudaStream_t st)
{
    if (!shape_ok(a)) return false;
    const int ncl = cluster_grid<0>(xy_fwd_cluster_kernel, a.nplanes);
    if (ncl <= 0) return false;
    xy_fwd_cluster_kernel<<<ncl * 8, kThreads, kSmemBytes, st>>>(a);
    FC_CUDA_KERNEL();
    return true;
}

bool launch_yx_inv_cluster(const XYArgs& a, cudaStream_t st)
{
    if (!shape_ok(a)) return false;
    const int ncl = cluster_grid<1>(yx_inv_cluster_kernel, a.nplanes);
    if (ncl <= 0) return false;
    yx_inv_cluster_kernel<<<ncl * 8, kThreads, kSmemBytes, st>>>(a);
    FC_CUDA_KERNEL();
    return true;
}

}  // namespace fcb200
