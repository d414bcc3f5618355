// Host-side planner and plan cache of the B200-native FFT convolution library.
#pragma once
#include "fc_hostpipe.h"
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "fc_common.h"
#include "fft_kernels.h"

namespace fcb200 {

// ---- pure host planning (no CUDA calls; unit-tested on the CPU) -----------------------------------
std::vector<int> factorize(int L, bool* generic, int style = 0);
void build_tables(int L, const std::vector<int>& radix, std::vector<int>& rev, std::vector<int>& pos,
                  std::vector<float2>& tw);
Geometry make_geometry(int nx, int ny, int nz);
// rows (z*ny + y) of the padded PSF volume that hold >= 1 tap under the reference placement
std::vector<int> psf_active_rows(const int* dims /*d0,d1,d2 as passed to fftShiftKernel*/, const int* kdims,
                                 int nx /*fastest extent of the consumed volume*/);

// ---- device-side plan -------------------------------------------------------------------------
struct AxisPlan {
    int L = 0;
    std::vector<int> radix;
    bool generic = false;
    float2* d_tw = nullptr;
    int* d_rev = nullptr;
    int* d_pos = nullptr;
    // Rader tables (device), when the last radix is a large prime whose p - 1 is smooth
    float2 *d_rtw = nullptr, *d_rbf = nullptr, *d_rbi = nullptr;
    int *d_rperm = nullptr, *d_riperm = nullptr;
    AxisPlanDev dev{};
};

// Host-side Rader tables for a prime p (pure host code; unit-tested on the CPU through fcb200_plan_rader).
struct RaderTables {
    int p = 0, n = 0;
    std::vector<int> radix;            // stages of the n-point transform (register radices only)
    std::vector<int> perm, iperm;      // g^m mod p, g^(-q) mod p
    std::vector<float2> tw, bf, bi;    // n roots; spectra of b (forward / inverse), / n, in position order of the n-point DIF
};
// false when p - 1 needs a prime factor above 16 other than 13 / 11 (then the direct sum stays)
bool build_rader(int p, RaderTables& r);

struct ConvPlan {
    int device = 0;
    Geometry g{};
    AxisPlan px, py, pz;
    float2* d_twx = nullptr;     // exp(-2*pi*i*k/nx), k = 0..nx/2
    int txp_y = 8, txp_z = 8;
    // workspace
    float2* d_spec = nullptr;    // image spectrum   [nz][ny][xcp]
    float2* d_H = nullptr;       // PSF spectrum     [nz][ny][xcp]
    float* d_real = nullptr;     // staging for host-pointer calls [nz][ny][nx]
    float* d_kernel = nullptr;   // PSF taps staging
    size_t kernel_cap = 0;
    float* d_unpadded = nullptr; // padded entry points with host pointers: the caller's (unpadded) volume
    size_t unpadded_cap = 0;     // floats
    // PSF pruning lists, cached per kernel shape / placement dims
    int psf_key[6] = {0, 0, 0, 0, 0, 0};
    int win_key[6] = {0, 0, 0, 0, 0, 0};
    int* d_rows = nullptr;       // rows (z*ny+y) the PSF x pass processes: the rows of the listed planes that hold a tap
    int* d_rows_c = nullptr;     // the same rows as indices into a compact buffer of the listed planes: (list plane)*ny + y
    long long n_rows = 0;
    size_t rows_cap = 0;
    int* d_planes = nullptr;     // z planes that hold >= 1 tap
    int n_planes = 0;
    std::vector<int> h_planes;   // host copy of d_planes
    float2* d_Hwin = nullptr;    // on-the-fly path: (x,y)-transformed PSF planes of the window [psf_window_planes][ny][xcp]
    size_t hwin_cap = 0;         // float2 elements
    int* d_win_slot = nullptr;   // [16] compact plane of window position n, -1 = none (16-plane windows, register kernel)
    int psf_window_z0 = -1;      // >= 0: every plane that holds a tap lies in [z0, z0 + psf_window_planes) mod nz
    int psf_window_planes = 0;   // 16, 32 or 64; the plane lists are then in WINDOW order (list plane n = z0 + n)
    // the window buffer holds the planes of exactly these host taps (PSF cache of the on-the-fly path)
    bool hwin_valid = false;
    int hwin_dims[6] = {0, 0, 0, 0, 0, 0};
    std::vector<float> hwin_taps;
    int* d_tap_start = nullptr;  // CSR tap lists over d_rows (see XArgs)
    int* d_tap_x = nullptr;
    int* d_tap_idx = nullptr;
    size_t taps_cap = 0;
    unsigned char* d_plane_mask = nullptr;   // [nz] 1 = plane active
    // "workspace busy": recorded on the stream of every call when its last kernel has been enqueued; the next call --
    // possibly on another stream -- makes its stream wait for it before it touches d_spec / d_H / the PSF buffers
    // (stream-ordered async calls of one shape on different streams would otherwise race on the shared workspace)
    cudaEvent_t ev_busy = nullptr;
    cudaStream_t stream = nullptr;   // used for host-pointer calls
    // host-pointer pipeline: pinned staging for pageable buffers, ring of device image buffers + copy
    // streams for the batch entry point (ring[0] aliases d_real)
    HostStager stager;
    float* d_ring[3] = {nullptr, nullptr, nullptr};
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_up[3] = {nullptr, nullptr, nullptr}, ev_comp[3] = {nullptr, nullptr, nullptr},
                ev_down[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_chunk[16] = {};
    // the PSF passes run on their own stream next to the image's x/y passes and join before the fused z pass
    cudaStream_t s_psf = nullptr;
    cudaEvent_t ev_psf_fork = nullptr, ev_psf_done = nullptr;   // single pinned call: per-chunk upload / compute events (8 + 8)
    // PSF-spectrum cache across calls: d_H holds the spectrum of exactly these taps
    bool h_valid = false;
    int h_dims[6] = {0, 0, 0, 0, 0, 0};
    std::vector<float> h_taps;
    std::mutex mu;
    unsigned long long last_use = 0;
    size_t spec_bytes() const { return (size_t)g.nz * g.ny * g.xcp * sizeof(float2); }
    size_t real_bytes() const { return (size_t)g.nz * g.ny * g.nx * sizeof(float); }
    ~ConvPlan();
};

// Returns the cached plan for (device, nx, ny, nz), creating it (tables + workspace) if needed.
// The caller must hold plan->mu while using the workspace.
// workspace = false: tables only (slab mode, where the caller owns the slab-sized buffers).
std::shared_ptr<ConvPlan> get_plan(int device, int nx, int ny, int nz, bool workspace = true);
// Orders a call on stream `st` after the previous user of the plan's workspace / marks the end of this call's use.
void workspace_acquire(ConvPlan& p, cudaStream_t st);
void workspace_release(ConvPlan& p, cudaStream_t st);
void release_all_plans();
// cudaMalloc for the lazily allocated, volume-sized buffers of a plan (PSF spectrum, staging, ring, window): when the
// device is out of memory every IDLE cached plan is dropped and the allocation retried once -- the reference frees
// everything after each call, so a call with a new large shape must not fail because idle plans still hold memory.
cudaError_t device_alloc_retry(void** p, size_t bytes);
template <typename T>
inline cudaError_t device_alloc_retry(T** p, size_t bytes) { return device_alloc_retry(reinterpret_cast<void**>(p), bytes); }
long long launch_count();
void count_launches(int n);

// pass ids for the optional CUDA-event profile (fcb200_profile_*)
enum PassId { kPassPsfClear = 0, kPassPsfX, kPassPsfY, kPassPsfZ, kPassXFwd, kPassYFwd, kPassZFused, kPassYInv,
              kPassXInv, kNumPassIds };
void profile_enable(int on);
bool profile_enabled();
int profile_read(float* ms_sum, long long* counts, int n);

// ---- pipeline pieces (all enqueue on `st`) -------------------------------------------------------
// PSF spectrum into plan.d_H.  pdims = the six ints handed to fftShiftKernel by the reference
// (k0,k1,k2,d0,d1,d2); d_kernel = taps on the device.
void run_psf_spectrum(ConvPlan& p, const float* d_kernel, const int* pdims, cudaStream_t st);
// On-the-fly path: only the z planes of the PSF window (16..64 consecutive planes mod nz that hold every tap) are
// (x,y)-transformed, into a compact buffer; the fused z kernel derives the PSF spectrum from them tile by tile.
// psf_window_applies: can this plan / PSF placement take that path (no device work is enqueued)?
// run_psf_window: returns false (nothing done) when it does not apply: use the functions above.
bool psf_window_applies(ConvPlan& p, const int* pdims, cudaStream_t st);
bool run_psf_window(ConvPlan& p, const float* d_kernel, const int* pdims, cudaStream_t st);
void run_convolve_window(ConvPlan& p, float* d_real, cudaStream_t st);
void ensure_full_workspace(ConvPlan& p);   // allocates the image-sized PSF spectrum buffer on first use
// Spectrum of a dense image-sized volume into dst (used for the image and for the legacy
// convolution3DfftCUDA_test whose kernel is already image-sized).
void run_forward(ConvPlan& p, const float* d_real, float2* dst, int passes, cudaStream_t st);
// Image path: d_real (device, dense) is convolved in place with the PSF spectrum in plan.d_H.
void run_convolve(ConvPlan& p, float* d_real, cudaStream_t st);
// the same in three pieces (x+y forward / y+x inverse on z planes [z0, z0+n) of the volume at d_real)
// pad != nullptr (in-library padding, fused): the volume of the plan is the PADDED grid, d_real is the caller's
// UNPADDED volume; the x pass builds the padded rows of planes [z0, z0+n) while loading / writes only the
// interior of the rows back (no padded real volume exists in memory)
void run_xy_forward_planes(ConvPlan& p, const float* d_real, int z0, int n, cudaStream_t st,
                           const PadGeom* pad = nullptr);
void run_z_fused(ConvPlan& p, bool window, cudaStream_t st);
void run_yx_inverse_planes(ConvPlan& p, float* d_real, int z0, int n, cudaStream_t st, const PadGeom* pad = nullptr);
void run_inverse(ConvPlan& p, float2* spec, float* d_real, cudaStream_t st);

// ---- in-library padding (fc_pad.cu) ---------------------------------------------------------------
// Source volume [sz][sy][sx] embedded at offsets (ox,oy,oz) in the padded volume [pz][py][px] (x fastest).
// mode 0: zeros outside (reference tests/padd_utils.h:157-171); mode 1: mirror (numpy "reflect").
// padded planes [pz0, pz0+pn) of d_pad <- d_src (whole source volume on the device)
void run_pad_embed(const float* d_src, float* d_pad, const PadGeom& g, int pz0, int pn, cudaStream_t st);
// source planes [z0, z0+n) of d_dst <- interior of d_pad
void run_pad_crop(const float* d_pad, float* d_dst, const PadGeom& g, int z0, int n, cudaStream_t st);
// policy 0: image + 2*(kernel/2) per axis (the reference's zero_padd); 1: rounded up to a 7-smooth size
void padded_extents(const int* imDim, const int* kernelDim, int policy, int* padDim);

// ---- slab-decomposed single volume (multi-GPU): pass-level pieces on caller-owned device buffers ----
// A rank owns nzl consecutive z planes of the real volume and, after the exchange, nyl consecutive ky
// rows of every plane.  Buffers:  real slab [nzl][ny][nx];  z-slab spectrum [nzl][ny][xcp];
// exchange buffer [P][nzl][nyl][xcp] (P = ny / nyl blocks);  y-slab spectrum [nz][nyl][xcp].
// x + y forward on the slab; the y pass writes the exchange (send) buffer directly
// peers != nullptr: the y pass stores straight into the peers' y-slab buffers (NVLink), `send` is unused
// nzp: planes per exchange block (>= nzl; the same on every rank -- ragged slabs: nzp = ceil(nz / P) while the
// last rank owns fewer planes).  nyl likewise is the ROW PITCH of a ky block, ceil(ny / P).
void run_slab_xy_forward(ConvPlan& p, const float* d_real, float2* zslab, float2* send, int nzl, int nyl,
                         cudaStream_t st, float2* const* peers, int rank, int nzp, int z0 = 0, int nz_run = -1);
// fused z pass on the y-slab spectrum (in place) with the y-slab of the PSF spectrum
// peers != nullptr: the last inverse stage stores each output plane straight into its owner's receive buffer
// in_peers != nullptr (with peers): pull exchange -- input plane z is read from in_peers[z / nzl] + (z % nzl) * nyl * xcp
void run_slab_z_fused(ConvPlan& p, float2* yslab, const float2* Hslab, int nyl, cudaStream_t st,
                      float2* const* peers = nullptr, int rank = 0, int nzl = 0, float2* const* in_peers = nullptr);
// y + x inverse; the y pass reads the exchange (receive) buffer directly
void run_slab_yx_inverse(ConvPlan& p, const float2* recv, float2* zslab, float* d_real, int nzl, int nyl,
                         cudaStream_t st, int nzp, int z0 = 0, int nz_run = -1);
// PSF spectrum of the rank's y-slab: x and y passes on the planes that hold taps (into scratch, compact),
// rows [y0, y0+nyl) copied into Hslab, z pass with the plane mask.  Returns nothing; scratch must hold
// psf_slab_scratch_elems() float2.
size_t psf_slab_scratch_elems(ConvPlan& p, const int* pdims);
void run_slab_psf(ConvPlan& p, const float* d_kernel, const int* pdims, int y0, int nyl, float2* Hslab,
                  float2* scratch, cudaStream_t st);

}  // namespace fcb200
