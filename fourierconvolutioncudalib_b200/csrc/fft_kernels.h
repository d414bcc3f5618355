// Kernel argument blocks and launchers (see fft_kernels.cu).
#pragma once
#include "fc_common.h"

namespace fcb200 {

struct XArgs {
    const float* in_real;   // forward, image loader: [rows][nx]
    float* out_real;        // inverse: [rows][nx]
    float2* spec;           // [rows][xcp]
    Geometry g;
    AxisPlanDev P;          // complex transform of length g.M
    const float2* twx;      // exp(-2*pi*i*k/nx), k = 0..nx/2 (even nx only)
    long long nrows;        // rows to process (length of rowList when given, else ny*nz)
    const int* rowList;     // optional: global row index of each processed row (PSF pruning)
    int txp;                // run-time-plan kernels: row pairs per CTA (8, or fewer when the row is very long)
    int compactOut;         // with rowList: write spectrum row i of the list to spec row i (not to its global row)
    PsfGather psf;          // PSF loader only
    // PSF loader, optional CSR tap lists (one entry per processed row): taps tapStart[i] .. tapStart[i+1]-1 of
    // list row i sit at x = tapX[t] with value kernel[tapIdx[t]]; when given, no per-voxel index math is done
    const int* tapStart;
    const int* tapX;
    const int* tapIdx;
    // In-library padding fused into the image x passes: the rows of this launch are rows of the PADDED grid
    // (g = padded geometry, row 0 = padded plane padZ0); in_real / out_real point at the caller's UNPADDED
    // volume.  Forward: each padded row is assembled from the source row (zeros / mirror) while loading.
    // Inverse: only the interior part of interior rows is stored.
    int padOn;
    int padZ0;
    PadGeom pad;
    int pdl;                // launch with programmatic stream serialization (fc_common.h: launch_pdl)
};

struct ColArgs {
    float2* data;           // transformed in place
    const float2* H;        // fused mode: PSF spectrum, same layout as data
    AxisPlanDev P;
    long long stride;       // float2 elements between consecutive transform positions
    long long groupStride;  // float2 elements between tile groups (z planes for the y pass)
    int tilesPerGroup;
    int rowLen;             // float2 elements per group row (xcp for y, ny*xcp for z); even
    int txp;                // column pairs per tile
    float scale;            // fused mode: 1/N
    const int* groupList;   // optional: group index of each launched group (PSF pruning: active planes)
    const unsigned char* rowMask;  // optional, forward only: rowMask[r]==0 => input row r is all zero, not read
    // Split layout (slab decomposition, multi-GPU): the transform axis is cut in blocks of splitRows rows,
    // one block per peer.  Forward (mode 0) WRITES and inverse (mode 1) READS row r at
    //   split + (r / splitRows) * splitBlock + group * splitGroup + (r % splitRows) * stride + column
    // so the y pass produces / consumes the all-to-all send / receive buffer directly (no pack pass).
    const int* winSlot;     // on-the-fly PSF path: compact plane slot of window position n (16 ints, -1 = no taps)
    int winPlanes;          // on-the-fly PSF path: planes of the window buffer a.H (16, 32 or 64), in window order
    float2* split;
    int splitRows;
    long long splitBlock;
    long long splitGroup;
    // Peer form of the split layout (fused compute + exchange over NVLink): block b of the transform axis lives
    // in ANOTHER GPU's buffer, splitPeers[b] (peer-mapped device pointers), at element offset splitPeerOffset.
    // Mode 0 and mode 2 (fused z) write their output rows there; mode 1 reads from the plain split buffer.
    float2* const* splitPeers;
    long long splitPeerOffset;
    // Pull form of the forward exchange (mode 2 only): INPUT row r of the transform axis is read from
    //   splitInPeers[r / splitRows] + (r % splitRows) * stride + column
    // i.e. the fused z pass fetches the planes straight out of the GPUs that produced them (128-byte loads over NVLink).
    float2* const* splitInPeers;
    int pdl;                // launch with programmatic stream serialization (fc_common.h: launch_pdl)
};

bool x_pass_supported(const Geometry& g, const AxisPlanDev& P);
size_t x_smem_bytes(const Geometry& g, const AxisPlanDev& P, int txp = 8);
int x_pick_txp(const Geometry& g, const AxisPlanDev& P);   // largest of 8,4,2,1 that fits; 0 if none
int col_pick_txp(const AxisPlanDev& P);
void launch_x_fwd(const XArgs& a, bool psf, cudaStream_t st);
void launch_x_inv(const XArgs& a, cudaStream_t st);
void launch_col(const ColArgs& a, int mode, long long ngroups, cudaStream_t st);
// PSF z pass when all non-zero input planes lie in a window of `planes` = 16, 32 or 64 consecutive planes starting at z0
// (mod L) and L % 16 == 0: each output residue is one radix-16 butterfly of the 16 twiddled inputs
// (input-pruned FFT); in place on `data` with the geometry of a z pass.  Returns false if not applicable.
bool launch_psf_z_pruned(const ColArgs& a, int z0, int planes, cudaStream_t st);
// fast path (fft_col_fast.cu): first/last stage fused with the global loads/stores
bool col_fast_supported(const AxisPlanDev& P);
void launch_col_fast(const ColArgs& a, int mode, long long ngroups, cudaStream_t st);
// TMA-staged persistent kernels (fft_col_tma.cu): tensor-map loads and stores, digit reversal done by the store's
// tensor map; return false when no configuration matches (plan, layout, mode)
bool launch_col_tma(const ColArgs& a, int mode, long long ngroups, cudaStream_t st);
// compile-time specialised kernels (fft_static.cu); return false when none matches the plan
bool col_static_has_plan(const AxisPlanDev& P);
bool launch_col_static(const ColArgs& a, int mode, long long ngroups, cudaStream_t st);
// fused z pass that derives the PSF-spectrum tile on the fly from the <=16 window planes starting at z0;
// a.H = buffer holding those planes (after the x and y passes); probe = only test applicability
bool launch_col_otf(const ColArgs& a, long long ngroups, int z0, cudaStream_t st, bool probe);
// the same on the TMA pipeline (fft_col_tma.cu); windows of 16..64 planes
bool launch_col_otf_tma(const ColArgs& a, long long ngroups, int z0, cudaStream_t st, bool probe);
bool launch_x_fwd_static(const XArgs& a, bool psf, cudaStream_t st);
bool launch_x_inv_static(const XArgs& a, cudaStream_t st);

}  // namespace fcb200
