// device_types.h -- PODs shared by the host runtime (tmc2gpu.cu) and the sm_100a kernels (kernels.cu).
//
// HBM layout of one batch (a GOF, or the slice of a GOF given to one GPU); F frames, M = 2 maps:
//   occ    [F][occ_h][occ_pitch]       u8   low-resolution occupancy video     (reference atlas.occ_frames)
//   geo    [F][2][H][geo_pitch]        u16  geometry video, channel 0 only     (reference atlas.geo_frames[0], frame f*2+m)
//   attr_y [F][2][H][attr_pitch_y]     u16  attribute video, channel 0         (reference atlas.attr_frames[0])
//   attr_u [F][2][H/2][attr_pitch_c]   u16  channel 1 (4:2:0)      attr_v likewise
//   patches[total]  DevPatch ; slot_patch[n_tiles*kWarpsPerTile] ; tile_frame[n_tiles] ; frame_tile_begin[F+1]
//   block_to_patch [F][bw*bh] u32
// Pitches are multiples of 64 elements so every 16x16 canvas block row starts on a 32-byte boundary.
// Outputs are per-frame slabs of `cap` points:  pos [F][cap][3] u16, rgb [F][cap][3] u8, and (debug / stage API only)
// yuv [F][cap][3] u16, partition [F][cap] u16, pixel [F][cap] u32 (x | y<<15 | map<<30), btype [F][cap] u8; count [F] u32.
// Smoothing state: per frame-in-group sparse voxel-cell tables (GeoCell / ColCell), touched-slot lists, and per frame a
// compact list of type-1 boundary points (BoundaryEntry).
#pragma once
#include <cstdint>

namespace tmc2 {

#ifndef TMC2_WARPS_PER_TILE
#define TMC2_WARPS_PER_TILE 8
#endif
#ifndef TMC2_MIN_CTAS
#define TMC2_MIN_CTAS 3
#endif
constexpr int kWarpsPerTile = TMC2_WARPS_PER_TILE;   // one warp per 16x16 patch block ("slot"); one CTA per tile of slots
constexpr uint32_t kNoPatch = 0xFFFFFFFFu;  // padding slot
constexpr int kSlotPoints = 512;            // max points of a 16x16 block (2 maps)

struct alignas(16) DevPatch { // reference Patch (src/decoder.rs:711-783), pre-digested on the host (64 B)
  int32_t  x0, y0;           // uv0 * occupancy_resolution  (pixels)
  uint32_t u0, v0;           // uv0 (blocks)
  uint32_t size_u0, size_v0; // size_uv0 (blocks)
  uint32_t u1, v1, d1;       // 3D shifts
  uint16_t lod_x, lod_y;
  uint8_t  normal, tangent, bitangent, mode;
  uint8_t  orient, _pad[3];
  uint32_t slot_base;        // index of this patch's first slot in slot_patch[]
  uint32_t local_index;      // patch index inside its frame (partition value; block_to_patch holds local_index+1)
  uint32_t frame;            // frame inside the batch
};

struct Planes {
  const uint8_t*  occ;
  const uint16_t* geo;
  const uint16_t* attr_y;
  const uint16_t* attr_u;
  const uint16_t* attr_v;
  uint32_t occ_pitch, occ_w, occ_h;
  uint32_t geo_pitch, attr_pitch_y, attr_pitch_c;
  uint64_t occ_frame_stride;     // elements between frames
  uint64_t geo_map_stride;       // elements between maps (frame stride = 2x)
  uint64_t attr_y_map_stride;
  uint64_t attr_c_map_stride;
};

struct Outputs {                // any pointer may be null = stream not wanted
  uint16_t* pos;
  uint8_t*  rgb;
  uint16_t* yuv;
  uint16_t* part;
  uint32_t* pix;
  uint8_t*  btype;
  uint64_t  cap;                // points per frame slab
};

// ---- sparse voxel-cell tables of the grid smoothing stages (own spec) ---------------------------------
constexpr uint32_t kCellEmpty = 0xFFFFFFFFu;
// A cell is "multi-patch" (the smoothing trigger) when points of two different patches fell into it: the first
// toucher CASes pfirst from 0 to patch+1, anybody who finds a different value there sets bit 31 of `count`.
constexpr uint32_t kCellMulti = 0x80000000u;
struct GeoCell {     // 32 B = one DRAM sector
  uint32_t key;      // cx | cy<<10 | cz<<20 ; kCellEmpty = free (hashed tables only; dense tables ignore it)
  uint32_t pfirst;   // patch index + 1 of the first point, 0 = untouched
  uint32_t count;    // points in the cell | kCellMulti
  uint32_t sx, sy, sz;          // sums of (coordinate - cell origin)  (< grid size each)
  uint32_t _pad[2];
};
struct ColCell {     // 32 B
  uint32_t key, pfirst, count;
  uint32_t sy, su, sv;             // sums of Y, U, V (exact while count <= 65536; the filter checks)
  unsigned long long sy2;          // sum of Y*Y
};
struct alignas(16) BoundaryEntry {  // one type-1 boundary point (16 B)
  uint32_t idx;         // point index inside its frame
  uint16_t pos[3];      // reconstructed (pre-smoothing) position
  uint16_t yuv[3];      // 16-bit colour
};

struct GridDesc {               // geometry of one voxel grid
  uint32_t on;
  uint32_t g, w, disth, th;     // cell edge, cells per axis, border margin, g*w
  uint32_t magic;               // ceil(2^32 / g): x / g == umulhi(x, magic) for x < 65536
  uint32_t identity;            // 1: slot = dense cell index (table covers the whole grid)
  uint64_t slots;               // table slots per frame-in-group (power of two unless identity)
  void*    table;               // GeoCell / ColCell [frames_in_group][slots]
  uint32_t* touched;            // [frames_in_group][touched_cap]
  uint32_t* touched_count;      // [frames_in_group]
};

struct SmoothArgs {
  GridDesc geo, col;
  uint64_t touched_cap;
  BoundaryEntry* blist;         // [F][blist_cap]
  uint32_t* blist_count;        // [F]
  uint64_t blist_cap;
  uint32_t group_first_frame;   // tables are indexed by (frame - group_first_frame)
  uint32_t group_frames;
  uint32_t thr_geo;             // threshold_smoothing
  uint32_t thr_col_smooth, thr_col_diff, thr_col_var;   // colour thresholds, already scaled to the sample bit depth
  unsigned long long* changed;  // [2][F] moved / recoloured points
};

struct UnpackArgs {
  Planes   in;
  Outputs  out;
  uint32_t W, H, res, prec;
  int32_t  prec_shift;          // log2(prec) or -1 when prec is not a power of two
  uint32_t bw, bh;              // block grid (W/res, H/res)
  uint32_t n_frames, n_tiles;
  uint8_t  absolute_d1, spec_orientation, has_attr, want_btype;
  const DevPatch* patches;
  const uint32_t* slot_patch;
  const uint32_t* tile_frame;
  const uint32_t* frame_tile_begin;   // [F+1]
  const uint32_t* block_to_patch;     // [F][bw*bh]
  uint64_t*       tile_status;        // chained-scan state, one word per tile
  uint32_t        epoch;              // launch tag inside the status words (no memset between launches)
  uint32_t*       tile_total;         // two-pass mode: per-tile totals (count kernel) / exclusive bases (emit kernel)
  uint32_t*       frame_count;        // [F] points per frame
  int*            err;                // device error flag (0 ok)
  // byte offsets of each staged stream inside a warp's shared-memory region, and the region size
  uint32_t off_scan, off_pos, off_rgb, off_yuv, off_part, off_pix, off_bt, warp_bytes;
  SmoothArgs sm;                      // used by the smoothing instantiation only
};

// launch wrappers (kernels.cu); every one enqueues on `stream` and returns the cudaGetLastError() code
int launch_block_to_patch(const UnpackArgs& a, uint32_t n_slots, void* stream);
// mode: 0 fused single pass, 1 count, 2 emit.  Tiles [tile_begin, tile_end).  smooth: accumulate cell tables + boundary list
int launch_unpack(const UnpackArgs& a, int mode, bool smooth, uint32_t tile_begin, uint32_t tile_end, void* stream);
int launch_tile_scan(const UnpackArgs& a, void* stream);
int launch_upsample(const UnpackArgs& a, uint8_t* occ_full /*[F][H][W]*/, void* stream);
int launch_smooth_filter(const UnpackArgs& a, void* stream);   // boundary points of the current frame group
int launch_smooth_clear(const UnpackArgs& a, void* stream);    // reset touched cells + list counters of the group
int launch_yuv_to_rgb_flat(const uint16_t* yuv, uint8_t* rgb, uint64_t n, void* stream);
int launch_table_init(void* table, uint64_t total_slots, int is_color, void* stream);
size_t unpack_smem_bytes(const UnpackArgs& a);
int kernel_launch_count_reset();   // returns launches since the last reset

}  // namespace tmc2
