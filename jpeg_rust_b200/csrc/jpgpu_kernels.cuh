// jpgpu_kernels.cuh — launch wrappers of the sm_100a kernels (jpgpu_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include "jpgpu_core.h"

namespace jpgpu {

struct SeqDesc {         // one warp job of the sync / write kernels: 32 consecutive subsequences of one image
    uint32_t img;        // image index, 0xffffffff = padding (a CTA's jobs must share their Huffman tables)
    uint32_t first_sub;  // first subsequence of the job within the image
};
constexpr uint32_t kNoImage = 0xffffffffu;

constexpr int kNumKinds = 6;

struct BatchDev {        // device pointers of one planned batch
    const ImgDev* imgs;
    ImgDyn* dyn;
    const SeqDesc* seqs;
    const HuffLut* luts;
    const uint32_t* mlut;     // multi-symbol tables of the synchronisation pass (jpgpu_core.h), one per entry of luts
    const uint32_t* mlut_off; // where each begins in mlut (words)
    uint32_t max_mlut_words;  // most words the tables of one image's slots take together (shared memory of sync_kernel)
    uint32_t sync_multi;      // 1: the synchronisation pass decodes through the multi-symbol tables
    uint32_t verify_multi;    // 1: so do the repair walks (their tables cost shared memory: only while all images' CTAs stay co-resident)
    const float* qt;     // pre-scaled dequantisation multipliers, 64 per table, column-major
    const uint8_t* raw;
    uint32_t* stream;
    uint32_t* segtab;
    SubInfo* subs;
    int16_t* coefs;
    uint8_t* rgb;
    uint32_t out_planar;     // 0: interleaved RGB triples (the reference's Vec<(u8,u8,u8)>), 1: three W x H u8 planes per image,
                             // 2: three W x H f32 planes holding u8 * out_scale[c] + out_bias[c] (4 bytes per sample: image i at rgb + 4 * rgb_off)
    float out_scale[3], out_bias[3];
    uint32_t n_images;   // images this launch covers, starting at img0 (a whole batch or one group of it)
    uint32_t n_seqs;     // warp jobs this launch covers, starting at job0
    uint32_t img0, job0;
    uint32_t nsync;      // images of this launch range that need sync_kernel / verify_scan_kernel (not interval_mode)
    uint32_t sub_bits;   // bits per subsequence for this batch
    uint32_t lw;         // log2(sub_bits / 32): words per subsequence
    uint32_t lookback_bits;
    uint32_t max_slots;
    uint32_t max_chunks;      // most 4 KiB raw chunks any image has
    uint2* chunk_counts;      // per chunk: kept bytes / RSTn markers, then (after the scan) those before the chunk
    uint32_t seg_bits;        // checkpoint distance inside a subsequence (divides sub_bits)
    uint32_t wp_shift;        // log2(units of the write pass per subsequence); a unit is a whole number of segments
    SegRec* segs;             // sub_bits / seg_bits records per subsequence
    // images grouped by colour-kernel variant (ImgKind)
    const uint32_t* kind_imgs[kNumKinds];
    uint32_t kind_count[kNumKinds];
    uint32_t kind_max_tiles[kNumKinds];
    // gather path (images of kind kKindGeneric)
    const uint32_t* gmap;    // placement maps: (block index << 6 | sample) per pixel and component
    float* samples;          // per-block IDCT output, [block][row*8+col]
    uint32_t gather_max_blocks;
    uint32_t gather_max_quads;  // most ceil(W*H/4) of any gather image
    // compose path (jpgpu_core.h): frames whose pixels are put together plane by plane from per-block samples
    const FrameDev* frames;
    uint32_t n_frames, frame_max_quads;
};

cudaError_t init_constants();

// Scan bytes that already lie in device memory (image i at base + dev_offs[i], indexed like BatchDev::imgs) -> raw arena.
void launch_gather_scans(const BatchDev& b, const void* base, const uint64_t* dev_offs, cudaStream_t s);
// The alignment gaps between the images' outputs in the RGB arena get defined contents (once per plan / arena / format).
void launch_zero_output_pads(const BatchDev& b, cudaStream_t s);
// Stage 1a: byte-unstuffing + RSTn detection: count per 4 KiB chunk, scan per image, compact per chunk.
void launch_prepass(const BatchDev& b, cudaStream_t s);
void launch_prepass_step(const BatchDev& b, cudaStream_t s, int step);  // 0 count, 1 scan, 2 write (profiling)
// Stage 1b: look-back synchronisation, one thread per subsequence.
cudaError_t launch_sync(const BatchDev& b, cudaStream_t s);
// Stage 1c: chain verification, repair of the links the look-back did not synchronise, prefix scan; one CTA per image.
cudaError_t launch_verify_scan(const BatchDev& b, cudaStream_t s);
// Stage 1d: final decode, whole coefficient blocks written to HBM.
cudaError_t launch_decode_write(const BatchDev& b, cudaStream_t s);
// Stage 2+3: dequantise, IDCT, upsample, YCbCr->RGB, interleaved store (SPEC geometry).
// Images whose requested layout the fused kernel cannot produce (REF placement of 4:2:0 / ragged widths,
// decoder.rs:259-312 + 347-379; generic sampling factors) go through block_idct_kernel +
// gather_colour_kernel and a host-built placement map.  Returns the number of kernels launched.
int launch_idct_colour(const BatchDev& b, cudaStream_t s);

}  // namespace jpgpu
