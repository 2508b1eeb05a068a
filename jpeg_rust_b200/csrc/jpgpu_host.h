// jpgpu_host.h — host-side planning shared by the C ABI (jpgpu_api.cu) and the
// CPU simulation harness used by tests.  No CUDA calls in here.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/jpgpu.h"
#include "jpgpu_kernels.cuh"

namespace jpgpu {

struct Geometry {
    uint32_t units = 0;           // MCUs decode() reads
    uint32_t blocks_per_mcu = 0;
    uint32_t mcux = 0, mcuy = 0;  // SPEC MCU grid
    uint8_t h[4] = {0, 0, 0, 0}, v[4] = {0, 0, 0, 0};
    uint8_t hmax = 1, vmax = 1;
    uint8_t kind = kKindGeneric;
    bool fused_ok = false;        // the fused SPEC-geometry kernel reproduces the requested layout
    bool compose = false;         // compose path (block IDCT + plane-wise up-sampling / colour): fancy up-sampling
    uint32_t nblocks[4] = {0, 0, 0, 0};
};

// decoder.rs:164-192 (REF) / T.81 A.2 (SPEC). Returns JPGPU_* status.
int compute_geometry(const jpgpu_image_desc& d, Geometry& g);

// Builds the device Huffman table from DHT BITS/HUFFVAL (huffman.rs:37-58, 80-98).
int build_huff_lut(const uint8_t bits[16], const uint8_t* vals, int nvals, bool is_dc, HuffLut& out);

// Multi-symbol table of the synchronisation pass for one DHT table (format: jpgpu_core.h, multi_entry): appends
// 2^kMultiBitsDc / 2^kMultiBitsAc entries to `out`.
void build_multi_lut(const uint8_t bits[16], const uint8_t* vals, bool is_dc, std::vector<uint32_t>& out);

// 64 multipliers (column-major) = q * aan[u] * aan[v] / 8 from a zigzag-order DQT table.
void build_qt_multipliers(const uint16_t qt_zigzag[64], float out[64]);

// A contiguous range of images whose kernels are launched together; the groups of a batch run on alternating
// streams so that the tail of one kernel overlaps the next group's work (jpgpu_batch_decode).
struct GroupPlan {
    uint32_t img0 = 0, nimg = 0, job0 = 0, njobs = 0, max_chunks = 0;
    uint32_t kind_lo[kNumKinds] = {0, 0, 0, 0, 0, 0}, kind_hi[kNumKinds] = {0, 0, 0, 0, 0, 0};
    uint32_t kind_max_tiles[kNumKinds] = {0, 0, 0, 0, 0, 0};
    uint32_t gather_max_blocks = 0, gather_max_quads = 0;
    uint32_t nsync = 0;          // images of the group that need the synchronisation pass (not interval_mode)
    uint32_t frame_lo = 0, frame_hi = 0, frame_max_quads = 0;   // compose path: the group's frames
};

struct HostPlan {
    std::vector<GroupPlan> groups;
    uint32_t sub_bits = kMinSubseqBits;
    uint32_t lw = 5;             // log2(words per subsequence)
    uint32_t lookback_bits = kDefaultLookbackBits;
    uint32_t seg_bits = kMinSegBits;  // checkpoint distance inside a subsequence
    uint32_t wp_shift = 0;       // log2(write-pass units per subsequence); a unit is a whole number of segments
    uint32_t max_slots = 1;      // most Huffman LUT slots any image references
    uint32_t nsync = 0;          // images that need the synchronisation pass
    std::vector<ImgDev> imgs;
    std::vector<int32_t> status;  // per image: JPGPU_OK or why it is skipped
    std::vector<uint32_t> out_w, out_h;   // per image: size of the output it owns (a frame's first scan: the frame; further scans: 0 x 0)
    std::vector<uint8_t> frame_part;      // per image: jpgpu_image_desc::frame_part
    std::vector<SeqDesc> seqs;
    std::vector<HuffLut> luts;
    std::vector<uint32_t> mluts;      // multi-symbol tables of the synchronisation pass, one per entry of `luts`
    std::vector<uint32_t> mlut_off;   // where each begins in `mluts` (words)
    uint32_t max_mlut_words = 0;      // most words the tables of one image's slots take together
    std::vector<float> qt;
    std::vector<uint32_t> kind_imgs[kNumKinds];
    uint32_t kind_max_tiles[kNumKinds] = {0, 0, 0, 0, 0, 0};
    uint64_t raw_bytes = 0;      // arena sizes
    uint64_t stream_words = 0;
    uint64_t seg_entries = 0;
    uint64_t sub_entries = 0;
    uint64_t chunk_entries = 0;  // pre-pass chunk table
    uint32_t max_chunks = 0;
    uint64_t coef_elems = 0;
    uint64_t rgb_bytes = 0;
    std::vector<FrameDev> frames;   // compose path: output frames (fancy up-sampling)
    uint32_t frame_max_quads = 0;
    std::vector<uint32_t> gmap;  // placement maps of the gather path, one per distinct shape
    uint64_t sample_floats = 0;  // per-block IDCT samples of the gather-path images
    uint32_t gather_max_blocks = 0, gather_max_quads = 0;
    // algorithmic totals over the valid images
    uint64_t tot_scan_bytes = 0, tot_blocks = 0, tot_pixels = 0, tot_rgb_bytes = 0;
};

// sub_bits: a power of two in [1024, 8192], or 0 = choose from the batch size (env JPGPU_SUBSEQ_BITS overrides).
uint32_t choose_subseq_bits(uint64_t total_scan_bytes);
int build_plan(const jpgpu_image_desc* descs, size_t n, HostPlan& plan, uint32_t sub_bits = 0);

constexpr uint64_t kMaxGatherPixels = 1ull << 26;   // largest image the gather path (REF placement / generic sampling) takes

// Placement map of the gather path: for every component plane and pixel the (arena block index << 6 | sample index)
// whose value the layout puts there, kMapNone where nothing is written.  REF replays decoder.rs:290-312 + 347-379 on
// indices (last writer wins, spill past the right edge included); SPEC is T.81 A.2.3 with box replication.  Returns
// JPGPU_OK or the panic the reference's placement would raise (index out of bounds, decoder.rs:295 / 372).
int build_gather_map(const jpgpu_image_desc& d, const Geometry& g, uint32_t plane_entries, std::vector<uint32_t>& out);

// Reorders one image's coefficient arena ([mcu][block][column-major]) into the
// reference arrangement: per component, decode order, zigzag (decoder.rs:208-212).
void export_reference_order(const ImgDev& im, const int16_t* arena, int16_t* out, uint32_t nblocks[4]);

}  // namespace jpgpu
