// kernels.cu -- hand-written sm_100a kernels of the V-PCC rec0 reconstruction path.
//
//   K2  block_to_patch_kernel     src/codec.rs:205-250   (atomicMax == "later patch overwrites", :242-244)
//   K1  upsample_kernel           src/codec.rs:288-300   (materialised only for the stage API; fused otherwise)
//   K3+K4 unpack_kernel           src/codec.rs:352-480 (unpack loop, order, dedup), :517-565 (generate_points),
//                                 :569-658 (attribute fetch), :661-687 (YUV->RGB), src/decoder.rs:827-888 (patch maths)
//       + K5 boundary type per point, K6/K7 cell statistics (own spec) in the smoothing instantiation
//   K6/K7 smooth_filter_kernel    grid geometry + colour smoothing of the boundary points (own integer spec, DESIGN.md;
//                                 the reference has only stubs: decoder.rs:291-299)
//
// Ordering.  The reference emits points in (patch, v0, u0, v1, u1, map) order.  A 16x16 patch block ("slot") that
// owns its canvas block emits one contiguous run, so the output position of a run is an exclusive prefix sum of
// per-slot counts in slot order.  unpack_kernel is ONE pass: a warp owns a slot, a CTA owns a tile of 8 consecutive
// slots, and tiles publish / look back their prefix through `tile_status` (single-pass chained scan with decoupled
// look-back, tile id == blockIdx.x, one scan domain per frame).  Nothing is read twice from HBM.
//
// All arithmetic on the bit-exact path is integer.  The colour conversion of the reference is IEEE f64; it is evaluated
// here in 32.32 fixed point with a proven error margin, and re-done with the literal f64 sequence (__dmul_rn/__dadd_rn/
// __ddiv_rn, every operation rounded on its own like rustc's code) whenever the fixed-point value is within that margin
// of an integer boundary, so the result is bit-exact for every u16 input.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "device_types.h"

namespace tmc2 {

static int g_launches = 0;
int kernel_launch_count_reset() { int n = g_launches; g_launches = 0; return n; }

// ----------------------------------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ uint32_t div_prec(uint32_t x, uint32_t prec, int shift) {
  return shift >= 0 ? (x >> shift) : (x / prec);
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {   // streaming 16-byte load, no L1 allocation
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_nc_v2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint32_t u16_of(const uint4& v, int j) {   // j-th u16 of a 16-byte vector (j constant)
  const uint32_t w = (j >> 1) == 0 ? v.x : (j >> 1) == 1 ? v.y : (j >> 1) == 2 ? v.z : v.w;
  return (j & 1) ? (w >> 16) : (w & 0xFFFFu);
}
__device__ __forceinline__ uint32_t u16_of(const uint2& v, int j) {
  const uint32_t w = (j >> 1) == 0 ? v.x : v.y;
  return (j & 1) ? (w >> 16) : (w & 0xFFFFu);
}

// ---- Patch maths, src/decoder.rs:853-888 --------------------------------------------------------------------------
// Forward map patch (u,v) -> canvas (x,y), decoder.rs:853-867.  `sscale` multiplies size_uv0: 1 reproduces the
// reference (sizes stay in blocks even at pixel level), `res` is the spec-correct form.  Signed 64-bit arithmetic: the
// host has verified that every pixel of the patch lands inside the canvas, where wrapping usize and i64 agree.
__device__ __forceinline__ void patch_to_canvas(const DevPatch& P, int64_t u, int64_t v, int64_t res, int64_t sscale,
                                                int64_t& x, int64_t& y) {
  const int64_t u0 = (int64_t)P.u0 * res, v0 = (int64_t)P.v0 * res;
  const int64_t su = (int64_t)P.size_u0 * sscale, sv = (int64_t)P.size_v0 * sscale;
  switch (P.orient) {
    case 0:  x = u + u0;           y = v + v0;           break;   // Default
    case 2:  x = sv - 1 - v + u0;  y = u + v0;           break;   // Rot90
    case 3:  x = su - 1 - u + u0;  y = sv - 1 - v + v0;  break;   // Rot180
    case 4:  x = v + u0;           y = su - 1 - u + v0;  break;   // Rot270
    case 5:  x = su - 1 - u + u0;  y = v + v0;           break;   // Mirror
    case 6:  x = sv - 1 - v + u0;  y = su - 1 - u + v0;  break;   // MRot90
    case 7:  x = u + u0;           y = sv - 1 - v + v0;  break;   // MRot180
    default: x = v + u0;           y = u + v0;           break;   // Swap (1) and MRot270 (8)
  }
}
// Inverse map canvas (x,y) -> patch (u,v) for block-aligned orientations (sizes scaled by res).
__device__ __forceinline__ void canvas_to_patch(const DevPatch& P, int32_t x, int32_t y, int32_t res, int32_t& u,
                                                int32_t& v) {
  const int32_t dx = x - P.x0, dy = y - P.y0;
  const int32_t su = (int32_t)P.size_u0 * res, sv = (int32_t)P.size_v0 * res;
  switch (P.orient) {
    case 0:  u = dx;          v = dy;          break;
    case 2:  u = dy;          v = sv - 1 - dx; break;
    case 3:  u = su - 1 - dx; v = sv - 1 - dy; break;
    case 4:  u = su - 1 - dy; v = dx;          break;
    case 5:  u = su - 1 - dx; v = dy;          break;
    case 6:  u = su - 1 - dy; v = sv - 1 - dx; break;
    case 7:  u = dx;          v = sv - 1 - dy; break;
    default: u = dy;          v = dx;          break;
  }
}

// generate_normal_coordinate (decoder.rs:881-888), truncated to u16 like the `as u16` cast at :874
__device__ __forceinline__ uint32_t normal_coord(const DevPatch& P, uint32_t depth) {
  const uint32_t n = P.mode == 0 ? depth + P.d1 : (P.d1 > depth ? P.d1 : depth) - depth;
  return n & 0xFFFFu;
}
// generate_point (decoder.rs:871-878) writes point[normal], point[tangent], point[bitangent] in that order, so a later
// axis overwrites an earlier one if they coincide.  axis_source() returns, for output axis `a`, which value lands there:
// 0 nothing, 1 normal, 2 tangent, 3 bitangent.
__device__ __forceinline__ uint32_t axis_source(const DevPatch& P, uint32_t a) {
  return P.bitangent == a ? 3u : P.tangent == a ? 2u : P.normal == a ? 1u : 0u;
}
__device__ __forceinline__ uint32_t pick(uint32_t src, uint32_t n, uint32_t t, uint32_t b) {
  return src == 3u ? b : src == 2u ? t : src == 1u ? n : 0u;
}

// ---- convert_yuv10_to_rgb8, src/codec.rs:661-687 ---------------------------------------------------------------------
// literal f64 sequence: one channel = clamp(floor(c / 1023 * 255))
__device__ __forceinline__ uint32_t quant_channel_f64(double c) {
  const double q = floor(__dmul_rn(__ddiv_rn(c, 1023.0), 255.0));
  if (q < 0.0) return 0u;
  if (q > 255.0) return 255u;
  return (uint32_t)q;
}
__device__ __noinline__ uint32_t yuv_to_rgb_f64(uint32_t Y, uint32_t U, uint32_t V) {
  const double y = (double)Y, u = __dsub_rn((double)U, 512.0), v = __dsub_rn((double)V, 512.0);
  const double r = __dadd_rn(y, __dmul_rn(1.57480, v));
  const double g = __dsub_rn(__dsub_rn(y, __dmul_rn(0.18733, u)), __dmul_rn(0.46813, v));
  const double b = __dadd_rn(y, __dmul_rn(1.85563, u));
  return quant_channel_f64(r) | (quant_channel_f64(g) << 8) | (quant_channel_f64(b) << 16);
}
// Exact integer evaluation.  With d = chroma - 512 and t = k*d, a channel is floor(T), T = 255*(Y + t)/1023.
// Write 255*t = I + f with I integer and 0 <= f < 1: then floor((255*Y + I + f)/1023) == floor((255*Y + I)/1023), because
// (255*Y + I)/1023 has a fractional part <= 1022/1023 and f/1023 < 1/1023.  So the chroma-dependent work is done ONCE per
// chroma sample (ChromaTerm: three integers), and a point costs three 32-bit multiply-shift divisions.
// Exactness versus the reference's f64 chain: the chain deviates from the real value by < 2e-11, and I, f are computed
// from the f64 constants rounded to 32 fractional bits (error <= 255*|d|/2^33), so floor() can only differ when f is
// within `eps` = 255*|d|/2^33 + 2^-24 of 0 or 1 AND 255*Y + I is congruent to 0 / 1022 mod 1023.  Such chroma samples are
// flagged (a few per million for 10-bit content) and their points take the literal f64 path.
constexpr long long kKr = 6763714498LL;   // 1.57480 * 2^32
constexpr long long kKgu = 804576224LL;   // 0.18733 * 2^32
constexpr long long kKgv = 2010603040LL;  // 0.46813 * 2^32
constexpr long long kKb = 7969870163LL;   // 1.85563 * 2^32
struct ChromaTerm { int32_t ir, ig, ib; uint32_t flagged; };   // flagged: some channel's f is within eps of 0 or 1
__device__ __forceinline__ int32_t chroma_floor(long long p32 /* 255*k*d in 32.32 */, uint32_t eps, uint32_t& flagged) {
  const uint32_t f = (uint32_t)p32;                         // fractional part, units of 2^-32
  flagged |= (f + eps <= 2u * eps) ? 1u : 0u;               // f < eps or f > 2^32 - 1 - eps (unsigned wrap)
  return (int32_t)(p32 >> 32);                              // floor (arithmetic shift)
}
__device__ __forceinline__ ChromaTerm chroma_term(uint32_t U, uint32_t V) {
  const int32_t du = (int32_t)U - 512, dv = (int32_t)V - 512;
  const uint32_t adu = (uint32_t)abs(du), adv = (uint32_t)abs(dv);
  ChromaTerm c;
  c.flagged = 0;
  c.ir = chroma_floor(255LL * kKr * dv, 128u * adv + 256u, c.flagged);
  c.ig = chroma_floor(-(255LL * kKgu) * du - (255LL * kKgv) * dv, 128u * (adu + adv) + 256u, c.flagged);
  c.ib = chroma_floor(255LL * kKb * du, 128u * adu + 256u, c.flagged);
  return c;
}
// clamp(floor(m / 1023), 0, 255) for m < 7.1e7: floor(m / 1023) == umulhi(m, ceil(2^36 / 1023)) >> 4
__device__ __forceinline__ uint32_t quant_int(int32_t m) {
  const uint32_t q = __umulhi((uint32_t)max(m, 0), 67174465u) >> 4;
  return min(q, 255u);
}
// A flagged chroma sample only matters when 255*Y + I sits right at a multiple of 1023: then the f64 chain decides.
__device__ __noinline__ uint32_t yuv_to_rgb_flagged(uint32_t Y, uint32_t U, uint32_t V, int32_t ir, int32_t ig, int32_t ib) {
  const int32_t y255 = (int32_t)(255u * Y);
  const int32_t m[3] = {y255 + ir, y255 + ig, y255 + ib};
  bool unc = false;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (m[i] <= 0) continue;
    const uint32_t rem = (uint32_t)m[i] % 1023u;
    unc |= rem == 0u || rem == 1022u;
  }
  if (unc) return yuv_to_rgb_f64(Y, U, V);
  return quant_int(m[0]) | (quant_int(m[1]) << 8) | (quant_int(m[2]) << 16);
}
__device__ __forceinline__ uint32_t yuv_to_rgb_term(uint32_t Y, uint32_t U, uint32_t V, const ChromaTerm& c) {
  if (c.flagged) return yuv_to_rgb_flagged(Y, U, V, c.ir, c.ig, c.ib);   // rare unless the chroma is exactly neutral
  const int32_t y255 = (int32_t)(255u * Y);
  return quant_int(y255 + c.ir) | (quant_int(y255 + c.ig) << 8) | (quant_int(y255 + c.ib) << 16);
}
__device__ __forceinline__ uint32_t yuv_to_rgb_packed(uint32_t Y, uint32_t U, uint32_t V) {
  return yuv_to_rgb_term(Y, U, V, chroma_term(U, V));
}

// occupancy of the full-resolution pixel (x,y) straight from the low-resolution video (codec.rs:294-298)
__device__ __forceinline__ uint32_t occ_at(const UnpackArgs& a, const uint8_t* occ_f, uint32_t x, uint32_t y) {
  return occ_f[(uint64_t)div_prec(y, a.prec, a.prec_shift) * a.in.occ_pitch + div_prec(x, a.prec, a.prec_shift)];
}

// K5: boundary type of an occupied pixel (own spec): 1 = image border or an unoccupied 4-neighbour, 2 = an unoccupied
// pixel inside the 5x5 window (clipped to the image), 0 = interior.  Evaluated per low-resolution cell.
__device__ uint32_t boundary_type(const UnpackArgs& a, const uint8_t* occ_f, int32_t x, int32_t y) {
  const int32_t W = (int32_t)a.W, H = (int32_t)a.H;
  if (x == 0 || y == 0 || x == W - 1 || y == H - 1) return 1;
  if (!occ_at(a, occ_f, x - 1, y) || !occ_at(a, occ_f, x + 1, y) || !occ_at(a, occ_f, x, y - 1) ||
      !occ_at(a, occ_f, x, y + 1))
    return 1;
  const int32_t xa = max(x - 2, 0), xb = min(x + 2, W - 1), ya = max(y - 2, 0), yb = min(y + 2, H - 1);
  const uint32_t cxa = div_prec(xa, a.prec, a.prec_shift), cxb = div_prec(xb, a.prec, a.prec_shift);
  const uint32_t cya = div_prec(ya, a.prec, a.prec_shift), cyb = div_prec(yb, a.prec, a.prec_shift);
  for (uint32_t cy = cya; cy <= cyb; ++cy)
    for (uint32_t cx = cxa; cx <= cxb; ++cx)
      if (occ_f[(uint64_t)cy * a.in.occ_pitch + cx] == 0) return 2;
  return 0;
}

// ----------------------------------------------------------------------------------------------------------------
// sparse voxel-cell tables (own spec; see DESIGN.md "Smoothing specification")
// ----------------------------------------------------------------------------------------------------------------
// x / G.g for x < 65536 by multiply-high (magic = ceil(2^32 / g); exact while x * g < 2^32)
__device__ __forceinline__ uint32_t cell_div(uint32_t x, const GridDesc& G) { return G.g == 1u ? x : __umulhi(x, G.magic); }

__device__ __forceinline__ uint64_t cell_slot0(uint32_t key, const GridDesc& G) {
  const uint32_t cx = key & 1023u, cy = (key >> 10) & 1023u, cz = key >> 20;
  if (G.identity) return cx + (uint64_t)G.w * (cy + (uint64_t)G.w * cz);
  // 2x2x2 neighbouring cells share one 8-slot group: the filter's neighbourhood lookups stay within a few lines
  const uint32_t grp = (cx >> 1) | ((cy >> 1) << 9) | ((cz >> 1) << 18);
  const uint64_t h = ((uint64_t)grp * 0x9E3779B97F4A7C15ull) >> 24;
  return ((h << 3) | ((cx & 1u) | ((cy & 1u) << 1) | ((cz & 1u) << 2))) & (G.slots - 1);
}

// slot of `key` in the table of frame-in-group `fig`: direct index for dense tables, find-or-claim for hashed ones
template <typename Cell>
__device__ __forceinline__ Cell* cell_slot(const GridDesc& G, uint32_t fig, uint32_t key, int* err) {
  Cell* tab = reinterpret_cast<Cell*>(G.table) + (uint64_t)fig * G.slots;
  uint64_t i = cell_slot0(key, G);
  if (G.identity) return &tab[i];
  for (uint64_t probe = 0; probe < G.slots; ++probe) {
    uint32_t cur = *reinterpret_cast<volatile uint32_t*>(&tab[i].key);
    if (cur == kCellEmpty) cur = atomicCAS(&tab[i].key, kCellEmpty, key);
    if (cur == kCellEmpty || cur == key) return &tab[i];
    i = (i + 1) & (G.slots - 1);
  }
  atomicExch(err, 11);
  return nullptr;
}
template <typename Cell>
__device__ __forceinline__ const Cell* cell_find(const GridDesc& G, uint32_t fig, uint32_t key) {
  const Cell* tab = reinterpret_cast<const Cell*>(G.table) + (uint64_t)fig * G.slots;
  uint64_t i = cell_slot0(key, G);
  if (G.identity) return tab[i].pfirst ? &tab[i] : nullptr;
  for (uint64_t probe = 0; probe < G.slots; ++probe) {
    const uint32_t k = tab[i].key;
    if (k == key) return &tab[i];
    if (k == kCellEmpty) return nullptr;
    i = (i + 1) & (G.slots - 1);
  }
  return nullptr;
}
// first-toucher / multi-patch bookkeeping; the only operation of a flush that needs an answer from L2
template <typename Cell>
__device__ __forceinline__ void cell_claim(const GridDesc& G, uint32_t fig, Cell* c, uint32_t patch, uint64_t touched_cap, int* err) {
  const uint32_t old = atomicCAS(&c->pfirst, 0u, patch + 1u);
  if (old == 0u) {
    const uint32_t t = atomicAdd(&G.touched_count[fig], 1u);
    if (t < touched_cap) G.touched[(uint64_t)fig * touched_cap + t] =
        (uint32_t)(c - (reinterpret_cast<Cell*>(G.table) + (uint64_t)fig * G.slots));
    else atomicExch(err, 11);
  } else if (old != patch + 1u) {
    atomicOr(&c->count, kCellMulti);
  }
}

struct GeoRun { uint32_t key, cnt, sx, sy, sz; };
struct ColRun { uint32_t key, cnt, sy, su, sv; unsigned long long sy2; };

__device__ __forceinline__ void flush_geo(const UnpackArgs& a, uint32_t fig, const GeoRun& r, uint32_t patch) {
  if (r.cnt == 0) return;
  GeoCell* c = cell_slot<GeoCell>(a.sm.geo, fig, r.key, a.err);
  if (!c) return;
  atomicAdd(&c->count, r.cnt); atomicAdd(&c->sx, r.sx); atomicAdd(&c->sy, r.sy); atomicAdd(&c->sz, r.sz);   // REDs
  cell_claim(a.sm.geo, fig, c, patch, a.sm.touched_cap, a.err);
}
__device__ __forceinline__ void flush_col(const UnpackArgs& a, uint32_t fig, const ColRun& r, uint32_t patch) {
  if (r.cnt == 0) return;
  ColCell* c = cell_slot<ColCell>(a.sm.col, fig, r.key, a.err);
  if (!c) return;
  atomicAdd(&c->count, r.cnt); atomicAdd(&c->sy, r.sy); atomicAdd(&c->su, r.su); atomicAdd(&c->sv, r.sv);
  atomicAdd(&c->sy2, r.sy2);
  cell_claim(a.sm.col, fig, c, patch, a.sm.touched_cap, a.err);
}

// ----------------------------------------------------------------------------------------------------------------
// K2: block-to-patch map.  One warp per slot (= one 16x16 block of one patch).
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) block_to_patch_kernel(const UnpackArgs a, uint32_t n_slots,
                                                             uint32_t* __restrict__ b2p) {
  const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slot >= n_slots) return;
  const uint32_t pid = a.slot_patch[slot];
  if (pid == kNoPatch) return;
  const DevPatch P = a.patches[pid];
  const uint32_t s = slot - P.slot_base;
  const uint32_t v0 = s / P.size_u0, u0 = s - v0 * P.size_u0;
  int64_t bx, by;
  patch_to_canvas(P, u0, v0, 1, 1, bx, by);                       // codec.rs:220-225 (block variant, resolution 1)
  const uint8_t* occ_f = a.in.occ + (uint64_t)P.frame * a.in.occ_frame_stride;
  const uint32_t lane = lane_id();
  const uint32_t res = a.res;
  bool nz = false;
  const bool aligned = a.spec_orientation || P.orient <= 1 || P.orient == 8;
  if (aligned) {
    // the 16x16 patch pixels are exactly the canvas block: test the low-resolution samples that cover it
    const uint32_t x0 = (uint32_t)bx * res, y0 = (uint32_t)by * res;
    const uint32_t cxa = div_prec(x0, a.prec, a.prec_shift), cxb = div_prec(x0 + res - 1, a.prec, a.prec_shift);
    const uint32_t cya = div_prec(y0, a.prec, a.prec_shift), cyb = div_prec(y0 + res - 1, a.prec, a.prec_shift);
    const uint32_t nx = cxb - cxa + 1, n = nx * (cyb - cya + 1);
    for (uint32_t i = lane; i < n; i += 32) {
      const uint32_t cy = cya + i / nx, cx = cxa + i % nx;
      nz |= occ_f[(uint64_t)cy * a.in.occ_pitch + cx] != 0;
    }
  } else {
    // reference-literal pixel mapping for the rotated / mirrored orientations (codec.rs:227-241)
    const int64_t sscale = a.spec_orientation ? res : 1;
    for (uint32_t i = lane; i < res * res; i += 32) {
      const uint32_t v1 = i / res, u1 = i - v1 * res;
      int64_t x, y;
      patch_to_canvas(P, (int64_t)u0 * res + u1, (int64_t)v0 * res + v1, res, sscale, x, y);
      nz |= occ_at(a, occ_f, (uint32_t)x, (uint32_t)y) != 0;
    }
  }
  if (__any_sync(0xFFFFFFFFu, nz) && lane == 0)
    atomicMax(&b2p[(uint64_t)P.frame * a.bw * a.bh + (uint64_t)by * a.bw + (uint64_t)bx], P.local_index + 1);
}

// ----------------------------------------------------------------------------------------------------------------
// K1: occupancy upsample (codec.rs:288-300), materialised only when the caller asks for tile.occupancy_map.
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample_kernel(const UnpackArgs a, uint8_t* __restrict__ occ_full) {
  const uint64_t total = (uint64_t)a.n_frames * a.W * a.H;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t f = i / ((uint64_t)a.W * a.H);
    const uint32_t r = (uint32_t)(i - f * a.W * a.H);
    const uint32_t y = r / a.W, x = r - y * a.W;
    occ_full[i] = (uint8_t)occ_at(a, a.in.occ + f * a.in.occ_frame_stride, x, y);
  }
}

// ----------------------------------------------------------------------------------------------------------------
// K3+K4(+K5, + K6/K7 statistics): fused unpack.
// ----------------------------------------------------------------------------------------------------------------
constexpr unsigned long long kFlagAggregate = 1ull, kFlagInclusive = 2ull;
__device__ __forceinline__ unsigned long long pack_status(uint32_t epoch, unsigned long long flag, uint32_t value) {
  return ((unsigned long long)epoch << 34) | (flag << 32) | value;
}

// Copy `nbytes` staged at shared `sm` (16-byte aligned, staged before the global position was known) to global `g`
// (any alignment).  Global stores are aligned 16-byte vectors; the source words are re-aligned with a funnel shift.
__device__ __forceinline__ void warp_copy_out(uint8_t* __restrict__ g, const uint8_t* sm, uint32_t nbytes, uint32_t lane) {
  const uint32_t head = min(nbytes, (uint32_t)((16u - (uint32_t)((uintptr_t)g & 15u)) & 15u));
  if (lane < head) g[lane] = sm[lane];
  const uint32_t nvec = (nbytes - head) >> 4;
  uint4* g4 = reinterpret_cast<uint4*>(g + head);
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(sm) + (head >> 2);   // first source word (4-byte aligned)
  const uint32_t sh = (head & 3u) * 8u;                                          // byte phase inside a word, as bits
  for (uint32_t i = lane; i < nvec; i += 32) {
    const uint32_t* w = sw + i * 4;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = sh ? w[4] : 0u;
    uint4 v;
    v.x = __funnelshift_r(w0, w1, sh); v.y = __funnelshift_r(w1, w2, sh);
    v.z = __funnelshift_r(w2, w3, sh); v.w = __funnelshift_r(w3, w4, sh);
    g4[i] = v;
  }
  const uint32_t done = head + (nvec << 4);
  if (done + lane < nbytes) g[done + lane] = sm[done + lane];
}

__device__ __forceinline__ int64_t ceil_div_pos(int64_t n, int64_t d) { return n <= 0 ? 0 : (n + d - 1) / d; }

// normal coordinates (n0 | n1 << 16) of one pixel from its two geometry samples (codec.rs:534-558)
__device__ __forceinline__ uint32_t normals_of(const DevPatch& P, uint32_t s0, uint32_t s1, bool absolute_d1) {
  const uint32_t d0 = s0 >> 2, d1 = s1 >> 2;                                     // depth = sample / 4
  const uint32_t n0 = normal_coord(P, d0);
  const uint32_t n1 = absolute_d1 ? normal_coord(P, d1) : ((P.mode == 0 ? n0 + d1 : n0 - d1) & 0xFFFFu);
  return n0 | (n1 << 16);
}

// Generic slot path (any occupancy resolution; reference-literal rotated / mirrored orientations): lane = pixel, 32 at
// a time in patch raster order, everything straight to global memory.  Rare, kept out of line to keep the fast path lean.
template <bool kSmooth, bool kDebug>
__device__ __noinline__ void generic_slot_emit(const UnpackArgs& a, const DevPatch& P, uint32_t frame, uint32_t fig,
                                               uint32_t u0b, uint32_t v0b, uint64_t gidx) {
  const uint32_t lane = lane_id(), res = a.res;
  const uint8_t* occ_f = a.in.occ + (uint64_t)frame * a.in.occ_frame_stride;
  const uint16_t* geo0 = a.in.geo + (uint64_t)frame * 2 * a.in.geo_map_stride;
  const uint16_t* geo1 = geo0 + a.in.geo_map_stride;
  const int64_t sscale = a.spec_orientation ? res : 1;
  const uint32_t srcx = axis_source(P, 0), srcy = axis_source(P, 1), srcz = axis_source(P, 2);
  uint64_t run = gidx;
  for (uint32_t base = 0; base < res * res; base += 32) {
    const uint32_t i = base + lane;
    uint32_t c = 0, n0 = 0, n1 = 0, t = 0, b = 0;
    int64_t x = 0, y = 0;
    if (i < res * res) {
      const uint32_t v1 = i / res, u1 = i - v1 * res;
      const uint32_t u = u0b * res + u1, v = v0b * res + v1;
      patch_to_canvas(P, u, v, res, sscale, x, y);
      if (occ_at(a, occ_f, (uint32_t)x, (uint32_t)y)) {
        const uint64_t off = (uint64_t)y * a.in.geo_pitch + (uint64_t)x;
        const uint32_t nn = normals_of(P, geo0[off], geo1[off], a.absolute_d1);
        n0 = nn & 0xFFFFu; n1 = nn >> 16;
        c = n1 != n0 ? 2u : 1u;
        t = (u * P.lod_x + P.u1) & 0xFFFFu;
        b = (v * P.lod_y + P.v1) & 0xFFFFu;
      }
    }
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t tt = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= (uint32_t)d) incl += tt;
    }
    const uint32_t chunk_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    uint64_t k = run + (incl - c);
    uint32_t bt = 0;
    if (c && ((kDebug && a.out.btype) || kSmooth)) bt = boundary_type(a, occ_f, (int32_t)x, (int32_t)y);
    for (uint32_t m = 0; m < c; ++m, ++k) {
      const uint32_t n = m == 0 ? n0 : n1;
      const uint32_t X = pick(srcx, n, t, b), Yc = pick(srcy, n, t, b), Z = pick(srcz, n, t, b);
      if (a.out.pos) { uint16_t* d = a.out.pos + k * 3; d[0] = (uint16_t)X; d[1] = (uint16_t)Yc; d[2] = (uint16_t)Z; }
      uint32_t Y = 0, U = 0, V = 0;
      if (a.has_attr) {
        const uint64_t fm = (uint64_t)frame * 2 + m;
        Y = a.in.attr_y[fm * a.in.attr_y_map_stride + (uint64_t)y * a.in.attr_pitch_y + (uint64_t)x];
        const uint64_t co = fm * a.in.attr_c_map_stride + (uint64_t)(y >> 1) * a.in.attr_pitch_c + (uint64_t)(x >> 1);
        U = a.in.attr_u[co]; V = a.in.attr_v[co];
        if (kDebug && a.out.yuv) { uint16_t* d = a.out.yuv + k * 3; d[0] = (uint16_t)Y; d[1] = (uint16_t)U; d[2] = (uint16_t)V; }
        if (a.out.rgb) {
          const uint32_t cc = yuv_to_rgb_packed(Y, U, V);
          uint8_t* d = a.out.rgb + k * 3; d[0] = (uint8_t)cc; d[1] = (uint8_t)(cc >> 8); d[2] = (uint8_t)(cc >> 16);
        }
      }
      if (kDebug && a.out.part) a.out.part[k] = (uint16_t)P.local_index;
      if (kDebug && a.out.pix) a.out.pix[k] = (uint32_t)x | ((uint32_t)y << 15) | (m << 30);
      if (kDebug && a.out.btype) a.out.btype[k] = (uint8_t)bt;
      if (kSmooth) {
        if (a.sm.geo.on && X < a.sm.geo.th && Yc < a.sm.geo.th && Z < a.sm.geo.th) {
          const uint32_t g = a.sm.geo.g, cx = cell_div(X, a.sm.geo), cy = cell_div(Yc, a.sm.geo), cz = cell_div(Z, a.sm.geo);
          const GeoRun r = {cx | (cy << 10) | (cz << 20), 1, X - cx * g, Yc - cy * g, Z - cz * g};
          flush_geo(a, fig, r, P.local_index);
        }
        if (a.sm.col.on && a.has_attr && X < a.sm.col.th && Yc < a.sm.col.th && Z < a.sm.col.th) {
          const ColRun r = {cell_div(X, a.sm.col) | (cell_div(Yc, a.sm.col) << 10) | (cell_div(Z, a.sm.col) << 20), 1, Y, U, V,
                            (unsigned long long)Y * Y};
          flush_col(a, fig, r, P.local_index);
        }
        if (bt == 1) {
          const uint32_t li = atomicAdd(&a.sm.blist_count[frame], 1u);
          if (li < a.sm.blist_cap) {
            BoundaryEntry e;
            e.idx = (uint32_t)(k - (uint64_t)frame * a.out.cap);
            e.pos[0] = (uint16_t)X; e.pos[1] = (uint16_t)Yc; e.pos[2] = (uint16_t)Z;
            e.yuv[0] = (uint16_t)Y; e.yuv[1] = (uint16_t)U; e.yuv[2] = (uint16_t)V;
            a.sm.blist[(uint64_t)frame * a.sm.blist_cap + li] = e;
          } else atomicExch(a.err, 7);
        }
      }
    }
    run += chunk_total;
  }
}

// Cell statistics of one staged run (smoothing instantiation).  Work is split by cell-aligned squares of the colour
// grid in patch space (<= 25 per block for a cell edge of 4) so that a lane's points mostly share a cell; keys always
// come from the staged positions, so the split is only a grouping heuristic and stays exact under u16 wrap-around.
__device__ __noinline__ void accumulate_cells(const UnpackArgs& a, const DevPatch& P, uint32_t fig, uint32_t u0b, uint32_t v0b,
                                              const uint8_t* cnt_sm, const uint16_t* pre_sm, const uint8_t* s_pos,
                                              const uint8_t* s_yuv) {
  const uint32_t lane = lane_id();
  const bool do_geo = a.sm.geo.on != 0, do_col = a.sm.col.on != 0 && a.has_attr;
  const int64_t cg = do_col ? a.sm.col.g : a.sm.geo.g;
  const int64_t ulo = (int64_t)u0b * 16, vlo = (int64_t)v0b * 16;
  const int64_t lx = P.lod_x, ly = P.lod_y;
  const int64_t tc0 = (ulo * lx + P.u1) / cg, tc1 = ((ulo + 15) * lx + P.u1) / cg;
  const int64_t bc0 = (vlo * ly + P.v1) / cg, bc1 = ((vlo + 15) * ly + P.v1) / cg;
  const uint32_t nt = (uint32_t)(tc1 - tc0 + 1), nb = (uint32_t)(bc1 - bc0 + 1);
  GeoRun gr0 = {kCellEmpty, 0, 0, 0, 0}, gr1 = gr0;
  ColRun cr0 = {kCellEmpty, 0, 0, 0, 0, 0ull}, cr1 = cr0;
  const uint32_t patch = P.local_index;
  for (uint32_t pair = lane; pair < nt * nb; pair += 32) {
    const int64_t tc = tc0 + pair % nt, bc = bc0 + pair / nt;
    const int32_t ua_ = (int32_t)(lx ? min(max(ceil_div_pos(tc * cg - P.u1, lx), ulo), ulo + 16) : ulo);
    const int32_t ub_ = (int32_t)(lx ? min(max(ceil_div_pos((tc + 1) * cg - P.u1, lx), ulo), ulo + 16) : ulo + 16);
    const int32_t va_ = (int32_t)(ly ? min(max(ceil_div_pos(bc * cg - P.v1, ly), vlo), vlo + 16) : vlo);
    const int32_t vb_ = (int32_t)(ly ? min(max(ceil_div_pos((bc + 1) * cg - P.v1, ly), vlo), vlo + 16) : vlo + 16);
    for (int32_t vv = va_; vv < vb_; ++vv) {
      for (int32_t uu = ua_; uu < ub_; ++uu) {
        const uint32_t rank = (uint32_t)((vv & 15) * 16 + (uu & 15));
        const uint32_t cnt = cnt_sm[rank];
        const uint32_t k0 = pre_sm[rank];
        for (uint32_t i = 0; i < cnt; ++i) {
          const uint16_t* p = reinterpret_cast<const uint16_t*>(s_pos + (k0 + i) * 6);
          const uint32_t x = p[0], y = p[1], z = p[2];
          if (do_geo && x < a.sm.geo.th && y < a.sm.geo.th && z < a.sm.geo.th) {
            const uint32_t g = a.sm.geo.g;
            const uint32_t cx = cell_div(x, a.sm.geo), cy = cell_div(y, a.sm.geo), cz = cell_div(z, a.sm.geo);
            const uint32_t key = cx | (cy << 10) | (cz << 20);
            if (key != gr0.key) {
              if (key == gr1.key) { const GeoRun t = gr0; gr0 = gr1; gr1 = t; }
              else { flush_geo(a, fig, gr1, patch); gr1 = gr0; gr0 = {key, 0, 0, 0, 0}; }
            }
            gr0.cnt += 1; gr0.sx += x - cx * g; gr0.sy += y - cy * g; gr0.sz += z - cz * g;
          }
          if (do_col && x < a.sm.col.th && y < a.sm.col.th && z < a.sm.col.th) {
            const uint32_t key = cell_div(x, a.sm.col) | (cell_div(y, a.sm.col) << 10) | (cell_div(z, a.sm.col) << 20);
            if (key != cr0.key) {
              if (key == cr1.key) { const ColRun t = cr0; cr0 = cr1; cr1 = t; }
              else { flush_col(a, fig, cr1, patch); cr1 = cr0; cr0 = {key, 0, 0, 0, 0, 0ull}; }
            }
            const uint16_t* c = reinterpret_cast<const uint16_t*>(s_yuv + (k0 + i) * 6);
            const uint32_t Y = c[0];
            cr0.cnt += 1; cr0.sy += Y; cr0.su += c[1]; cr0.sv += c[2]; cr0.sy2 += (unsigned long long)Y * Y;
          }
        }
      }
    }
  }
  if (do_geo) { flush_geo(a, fig, gr0, patch); flush_geo(a, fig, gr1, patch); }
  if (do_col) { flush_col(a, fig, cr0, patch); flush_col(a, fig, cr1, patch); }
}

// kMode 0: fused single pass (chained scan) ; 1: count only ; 2: emit with tile bases.  kDebug adds the streams the
// reference materialises but nobody downstream needs (colors16bit, partition, point_to_pixel, boundary types).
// Exclusive prefix of this tile inside its frame by decoupled look-back over the tile status words (one warp).
__device__ __forceinline__ uint32_t tile_lookback(const UnpackArgs& a, uint32_t tile, uint32_t first_tile, uint32_t tile_sum,
                                                  uint32_t lane) {
  unsigned long long* status = reinterpret_cast<unsigned long long*>(a.tile_status);
  if (tile == first_tile) {
    if (lane == 0) st_relaxed_u64(status + tile, pack_status(a.epoch, kFlagInclusive, tile_sum));
    return 0;
  }
  if (lane == 0) st_relaxed_u64(status + tile, pack_status(a.epoch, kFlagAggregate, tile_sum));
  uint32_t excl = 0, spins = 0;
  int64_t look = (int64_t)tile - 1;
  while (true) {
    const int64_t idx = look - (int64_t)lane;
    const bool valid = idx >= (int64_t)first_tile;
    unsigned long long st = pack_status(a.epoch, kFlagInclusive, 0);       // before the frame: prefix 0
    if (valid) st = ld_relaxed_u64(status + idx);
    const bool ready = (st >> 34) == a.epoch && ((st >> 32) & 3ull) != 0;
    const uint32_t ready_mask = __ballot_sync(0xFFFFFFFFu, ready);
    const uint32_t incl_mask = __ballot_sync(0xFFFFFFFFu, ready && ((st >> 32) & 3ull) == kFlagInclusive);
    // usable as soon as every predecessor up to the nearest inclusive one (or the whole window) has published
    const uint32_t upto = incl_mask ? (uint32_t)(__ffs(incl_mask) - 1) : 31u;
    const uint32_t need = upto == 31u ? 0xFFFFFFFFu : ((2u << upto) - 1u);
    if ((ready_mask & need) != need) {
      if (++spins > (1u << 22)) { if (lane == 0) atomicExch(a.err, 11); break; }   // watchdog: never hang the GPU
      __nanosleep(20);
      continue;
    }
    const uint32_t v = lane <= upto ? (uint32_t)(st & 0xFFFFFFFFull) : 0u;
    excl += __reduce_add_sync(0xFFFFFFFFu, v);
    if (incl_mask) break;
    look -= 32;
  }
  if (lane == 0) st_relaxed_u64(status + tile, pack_status(a.epoch, kFlagInclusive, excl + tile_sum));
  return excl;
}

// kMode 0: fused single pass (chained scan) ; 1: count only ; 2: emit with tile bases.  kDebug adds the streams the
// reference materialises but nobody downstream needs (colors16bit, partition, point_to_pixel, boundary types).
//
// CTA protocol (no __syncthreads after the prologue): every warp posts its slot's point count, then stages its output
// in shared memory BEFORE the tile's base is known; the first warp that gets that far claims the look-back, publishes the
// base, and everybody copies out.  Look-back latency therefore overlaps the other warps' staging work.
template <int kMode, bool kSmooth, bool kDebug>
__global__ void __launch_bounds__(kWarpsPerTile * 32, (kSmooth || kDebug) ? (TMC2_MIN_CTAS * 8 / kWarpsPerTile * 2 / 3) : (TMC2_MIN_CTAS * 8 / kWarpsPerTile))
unpack_kernel(const __grid_constant__ UnpackArgs a, uint32_t tile_offset) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ uint32_t s_tot[kWarpsPerTile];
  __shared__ uint32_t s_posted, s_claim, s_ready, s_base;

  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t tile = blockIdx.x + tile_offset;
  const uint32_t slot = tile * kWarpsPerTile + warp;
  if (threadIdx.x == 0) { s_posted = 0; s_claim = 0; s_ready = 0; }
  const uint32_t frame = a.tile_frame[tile];
  const uint32_t pid = a.slot_patch[slot];
  const uint32_t res = a.res;
  __syncthreads();                                               // the only block-wide barrier

  DevPatch P;
  bool owned = false, cand = false;
  int32_t bx = 0, by = 0;
  uint32_t u0b = 0, v0b = 0;
  const uint32_t* b2p_ptr = nullptr;
  if (pid != kNoPatch) {
    P = a.patches[pid];
    const uint32_t s = slot - P.slot_base;
    v0b = s / P.size_u0; u0b = s - v0b * P.size_u0;
    int64_t bxx, byy;
    patch_to_canvas(P, u0b, v0b, 1, 1, bxx, byy);                 // codec.rs:373-378
    bx = (int32_t)bxx; by = (int32_t)byy;
    b2p_ptr = a.block_to_patch + (uint64_t)frame * a.bw * a.bh + (uint64_t)by * a.bw + bx;
    cand = res == 16 && (a.spec_orientation || P.orient <= 1 || P.orient == 8);
  }
  const uint8_t* occ_f = a.in.occ + (uint64_t)frame * a.in.occ_frame_stride;

  // ---- phase 1: load the block, decide which pixels emit 1 or 2 points ------------------------------------------
  uint32_t nn[8];                                          // fast path: n0 | n1 << 16 of this lane's 8 pixels
  uint4 ya = {0, 0, 0, 0}, yb = {0, 0, 0, 0};
  uint2 ua = {0, 0}, va = {0, 0}, ub = {0, 0}, vb = {0, 0};
  uint32_t m1 = 0, m2 = 0;     // fast path: bit j set = pixel j of this lane emits >=1 / 2 points
  uint32_t total = 0;
  const int32_t px = bx * 16 + (int32_t)(lane & 1) * 8;   // fast path: this lane's 8 canvas pixels (px..px+7, py)
  const int32_t py = by * 16 + (int32_t)(lane >> 1);
  bool fast = false;

  if (cand) {
    // plane loads are issued before the ownership answer arrives (one less dependent round trip); an unowned block
    // (a later patch took the canvas block, codec.rs:379) just drops them
    const uint32_t owner = __ldg(b2p_ptr);
    const uint16_t* geo0 = a.in.geo + (uint64_t)frame * 2 * a.in.geo_map_stride;
    const uint64_t goff = (uint64_t)py * a.in.geo_pitch + px;
    const uint4 g0 = ldg_nc_v4(geo0 + goff);
    const uint4 g1 = ldg_nc_v4(geo0 + a.in.geo_map_stride + goff);
    if (kMode != 1 && a.has_attr) {
      const uint16_t* ay0 = a.in.attr_y + (uint64_t)frame * 2 * a.in.attr_y_map_stride;
      const uint64_t yoff = (uint64_t)py * a.in.attr_pitch_y + px;
      ya = ldg_nc_v4(ay0 + yoff);
      yb = ldg_nc_v4(ay0 + a.in.attr_y_map_stride + yoff);
      const uint64_t coff = (uint64_t)(py >> 1) * a.in.attr_pitch_c + (px >> 1);
      const uint64_t cf = (uint64_t)frame * 2 * a.in.attr_c_map_stride;
      ua = ldg_nc_v2(a.in.attr_u + cf + coff);
      va = ldg_nc_v2(a.in.attr_v + cf + coff);
      ub = ldg_nc_v2(a.in.attr_u + cf + a.in.attr_c_map_stride + coff);
      vb = ldg_nc_v2(a.in.attr_v + cf + a.in.attr_c_map_stride + coff);
    }
    // occupancy bits of the 8 pixels (codec.rs:393-396: any non-zero sample counts)
    uint32_t occ_bits = 0;
    {
      const uint8_t* row = occ_f + (uint64_t)div_prec((uint32_t)py, a.prec, a.prec_shift) * a.in.occ_pitch;
      uint32_t prev_c = 0xFFFFFFFFu, prev_v = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint32_t c = div_prec((uint32_t)px + j, a.prec, a.prec_shift);
        if (c != prev_c) { prev_c = c; prev_v = row[c]; }
        occ_bits |= (prev_v != 0 ? 1u : 0u) << j;
      }
    }
    owned = owner == P.local_index + 1;                           // codec.rs:379
    fast = owned;
    if (owned) {
      m1 = occ_bits;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        nn[j] = normals_of(P, u16_of(g0, j), u16_of(g1, j), a.absolute_d1);
        if ((nn[j] >> 16) != (nn[j] & 0xFFFFu)) m2 |= (occ_bits & (1u << j));   // codec.rs:422-428 duplicate skip
      }
    }
    total = __reduce_add_sync(0xFFFFFFFFu, __popc(m1) + __popc(m2));
  } else if (pid != kNoPatch) {
    owned = __ldg(b2p_ptr) == P.local_index + 1;
    if (owned) {
      // generic path (any resolution, reference-literal rotated orientations): lane = pixel, 32 at a time
      const uint16_t* geo0 = a.in.geo + (uint64_t)frame * 2 * a.in.geo_map_stride;
      const uint16_t* geo1 = geo0 + a.in.geo_map_stride;
      const int64_t sscale = a.spec_orientation ? res : 1;
      for (uint32_t base = 0; base < res * res; base += 32) {
        const uint32_t i = base + lane;
        uint32_t c = 0;
        if (i < res * res) {
          const uint32_t v1 = i / res, u1 = i - v1 * res;
          int64_t x, y;
          patch_to_canvas(P, (int64_t)u0b * res + u1, (int64_t)v0b * res + v1, res, sscale, x, y);
          if (occ_at(a, occ_f, (uint32_t)x, (uint32_t)y)) {
            const uint64_t off = (uint64_t)y * a.in.geo_pitch + (uint64_t)x;
            const uint32_t n = normals_of(P, geo0[off], geo1[off], a.absolute_d1);
            c = (n >> 16) != (n & 0xFFFFu) ? 2u : 1u;
          }
        }
        total += __reduce_add_sync(0xFFFFFFFFu, c);
      }
    }
  }

  // post this slot's count
  if (lane == 0) {
    s_tot[warp] = total;
    __threadfence_block();
    const uint32_t prev = atomicAdd(&s_posted, 1u);
    if (kMode == 1 && prev == kWarpsPerTile - 1) {               // count-only launch: the last poster sums the tile
      uint32_t t = 0;
#pragma unroll
      for (int w = 0; w < kWarpsPerTile; ++w) t += reinterpret_cast<volatile uint32_t*>(s_tot)[w];
      a.tile_total[tile] = t;
    }
  }
  if (kMode == 1) return;

  // ---- phase 2a: stage the run in shared memory (needs only warp-local offsets) -----------------------------------
  uint8_t* wsm = smem + (size_t)warp * a.warp_bytes;
  uint8_t* cnt_sm = wsm + a.off_scan;                             // [256] counts by rank
  uint16_t* pre_sm = reinterpret_cast<uint16_t*>(cnt_sm + 256);   // [256] exclusive prefix by rank
  uint8_t* s_pos = wsm + a.off_pos;
  uint8_t* s_rgb = wsm + a.off_rgb;
  uint8_t* s_yuv = wsm + a.off_yuv;      // smoothing: staged even when not written out
  uint8_t* s_part = wsm + a.off_part;
  uint8_t* s_pix = wsm + a.off_pix;
  uint8_t* s_bt = wsm + a.off_bt;        // smoothing: staged even when not written out
  const bool w_rgb = a.out.rgb != nullptr;
  const bool w_yuv = kDebug && a.out.yuv != nullptr, w_part = kDebug && a.out.part != nullptr;
  const bool w_pix = kDebug && a.out.pix != nullptr, w_bt = kDebug && a.out.btype != nullptr;
  uint32_t n_boundary = 0;
  if (fast && total) {
    // exclusive prefix of the per-pixel counts in PATCH-LOCAL raster order (v1 major, u1 minor; codec.rs:382-385)
    int32_t pu, pv, pu1, pv1;
    canvas_to_patch(P, px, py, 16, pu, pv);
    canvas_to_patch(P, px + 1, py, 16, pu1, pv1);
    const int32_t su = pu1 - pu, sv = pv1 - pv;                     // patch-space step per canvas pixel
    const uint32_t rank0 = (uint32_t)((pv & 15) * 16 + (pu & 15));
    const int32_t rstep = sv * 16 + su;                             // rank step per canvas pixel (no wrap inside a block)
#pragma unroll
    for (int j = 0; j < 8; ++j)
      cnt_sm[rank0 + j * rstep] = (uint8_t)(((m1 >> j) & 1u) + ((m2 >> j) & 1u));
    __syncwarp();
    {
      const uint2 c8 = *reinterpret_cast<const uint2*>(cnt_sm + lane * 8);
      uint32_t c[8] = {c8.x & 255u, (c8.x >> 8) & 255u, (c8.x >> 16) & 255u, c8.x >> 24,
                       c8.y & 255u, (c8.y >> 8) & 255u, (c8.y >> 16) & 255u, c8.y >> 24};
      uint32_t sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += c[j];
      uint32_t incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= (uint32_t)d) incl += t;
      }
      uint32_t run = incl - sum;
      uint32_t o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { o[j] = run; run += c[j]; }
      uint4 packed;
      packed.x = o[0] | (o[1] << 16); packed.y = o[2] | (o[3] << 16);
      packed.z = o[4] | (o[5] << 16); packed.w = o[6] | (o[7] << 16);
      *reinterpret_cast<uint4*>(pre_sm + lane * 8) = packed;
    }
    __syncwarp();

    const bool st_yuv = a.has_attr && (w_yuv || kSmooth);
    const bool st_bt = w_bt || kSmooth;
    // generate_point (decoder.rs:871-878) stores normal, tangent, bitangent in that order: byte offsets inside a point.
    // Axes that are not a permutation leave a coordinate at 0 (and let later stores overwrite earlier ones).
    const uint32_t o_n = 2u * P.normal, o_t = 2u * P.tangent, o_b = 2u * P.bitangent;
    const bool perm = ((1u << P.normal) | (1u << P.tangent) | (1u << P.bitangent)) == 7u;

#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {                                            // chroma column: pixels 2cc, 2cc+1
      if (!((m1 >> (2 * cc)) & 3u)) continue;
      ChromaTerm ta, tb;
      uint32_t Ua = 0, Va = 0, Ub = 0, Vb = 0;
      if (a.has_attr) {
        Ua = u16_of(ua, cc); Va = u16_of(va, cc); Ub = u16_of(ub, cc); Vb = u16_of(vb, cc);   // decoder.rs:976-977
        if (w_rgb) { ta = chroma_term(Ua, Va); tb = chroma_term(Ub, Vb); }
      }
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = 2 * cc + jj;
        if (!((m1 >> j) & 1u)) continue;
        const int32_t u = pu + j * su, v = pv + j * sv;
        const uint32_t k0 = pre_sm[rank0 + j * rstep];
        const uint32_t t = (uint32_t)u * P.lod_x + P.u1;                          // decoder.rs:875 (stored as u16)
        const uint32_t b = (uint32_t)v * P.lod_y + P.v1;                          // decoder.rs:876
        const bool two = (m2 >> j) & 1u;
        uint32_t bt = 0;
        if (st_bt) { bt = boundary_type(a, occ_f, px + j, py); n_boundary += bt == 1u ? (two ? 2u : 1u) : 0u; }
        {                                                                         // map 0 (codec.rs:421, i == 0)
          uint8_t* d = s_pos + k0 * 6;
          if (!perm) { reinterpret_cast<uint16_t*>(d)[0] = 0; reinterpret_cast<uint16_t*>(d)[1] = 0; reinterpret_cast<uint16_t*>(d)[2] = 0; }
          *reinterpret_cast<uint16_t*>(d + o_n) = (uint16_t)nn[j];
          *reinterpret_cast<uint16_t*>(d + o_t) = (uint16_t)t;
          *reinterpret_cast<uint16_t*>(d + o_b) = (uint16_t)b;
          if (a.has_attr) {
            const uint32_t Y = u16_of(ya, j);                                     // codec.rs:637-640
            if (st_yuv) { uint16_t* q = reinterpret_cast<uint16_t*>(s_yuv + k0 * 6); q[0] = (uint16_t)Y; q[1] = (uint16_t)Ua; q[2] = (uint16_t)Va; }
            if (w_rgb) {
              const uint32_t c = yuv_to_rgb_term(Y, Ua, Va, ta);
              uint8_t* q = s_rgb + k0 * 3; q[0] = (uint8_t)c; q[1] = (uint8_t)(c >> 8); q[2] = (uint8_t)(c >> 16);
            }
          }
          if (w_part) *reinterpret_cast<uint16_t*>(s_part + k0 * 2) = (uint16_t)P.local_index;   // codec.rs:452
          if (w_pix) *reinterpret_cast<uint32_t*>(s_pix + k0 * 4) = (uint32_t)(px + j) | ((uint32_t)py << 15);
          if (st_bt) s_bt[k0] = (uint8_t)bt;
        }
        if (two) {                                                                // map 1 unless it duplicates map 0
          const uint32_t k1 = k0 + 1;
          uint8_t* d = s_pos + k1 * 6;
          if (!perm) { reinterpret_cast<uint16_t*>(d)[0] = 0; reinterpret_cast<uint16_t*>(d)[1] = 0; reinterpret_cast<uint16_t*>(d)[2] = 0; }
          *reinterpret_cast<uint16_t*>(d + o_n) = (uint16_t)(nn[j] >> 16);
          *reinterpret_cast<uint16_t*>(d + o_t) = (uint16_t)t;
          *reinterpret_cast<uint16_t*>(d + o_b) = (uint16_t)b;
          if (a.has_attr) {
            const uint32_t Y = u16_of(yb, j);
            if (st_yuv) { uint16_t* q = reinterpret_cast<uint16_t*>(s_yuv + k1 * 6); q[0] = (uint16_t)Y; q[1] = (uint16_t)Ub; q[2] = (uint16_t)Vb; }
            if (w_rgb) {
              const uint32_t c = yuv_to_rgb_term(Y, Ub, Vb, tb);
              uint8_t* q = s_rgb + k1 * 3; q[0] = (uint8_t)c; q[1] = (uint8_t)(c >> 8); q[2] = (uint8_t)(c >> 16);
            }
          }
          if (w_part) *reinterpret_cast<uint16_t*>(s_part + k1 * 2) = (uint16_t)P.local_index;
          if (w_pix) *reinterpret_cast<uint32_t*>(s_pix + k1 * 4) = (uint32_t)(px + j) | ((uint32_t)py << 15) | (1u << 30);
          if (st_bt) s_bt[k1] = (uint8_t)bt;
        }
      }
    }
    __syncwarp();
  }

  // ---- tile base: the first warp to get here does the look-back for the whole tile --------------------------------
  uint32_t claim = 0;
  if (lane == 0) claim = atomicAdd(&s_claim, 1u);
  claim = __shfl_sync(0xFFFFFFFFu, claim, 0);
  if (claim == 0) {
    while (*reinterpret_cast<volatile uint32_t*>(&s_posted) < (uint32_t)kWarpsPerTile) { }
    __threadfence_block();
    uint32_t tile_sum = 0;
#pragma unroll
    for (int w = 0; w < kWarpsPerTile; ++w) tile_sum += reinterpret_cast<volatile uint32_t*>(s_tot)[w];
    const uint32_t first_tile = a.frame_tile_begin[frame];
    const uint32_t excl = kMode == 2 ? a.tile_total[tile] : tile_lookback(a, tile, first_tile, tile_sum, lane);
    if (lane == 0) {
      if (tile + 1 == a.frame_tile_begin[frame + 1]) a.frame_count[frame] = excl + tile_sum;  // codec.rs:482
      s_base = excl;
      __threadfence_block();
      *reinterpret_cast<volatile uint32_t*>(&s_ready) = 1u;
    }
  }
  if (!owned || total == 0) return;
  while (*reinterpret_cast<volatile uint32_t*>(&s_ready) == 0u) { }
  __threadfence_block();
  uint32_t run_base = *reinterpret_cast<volatile uint32_t*>(&s_base);
  for (uint32_t w = 0; w < warp; ++w) run_base += reinterpret_cast<volatile uint32_t*>(s_tot)[w];
  if ((uint64_t)run_base + total > a.out.cap) {                   // cannot happen for footprints inside the canvas
    if (lane == 0) atomicExch(a.err, 7);
    return;
  }
  const uint64_t gidx = (uint64_t)frame * a.out.cap + run_base;   // first point of this run
  const uint32_t fig = kSmooth ? frame - a.sm.group_first_frame : 0u;   // frame inside the smoothing group

  // ---- phase 2b: copy-out ---------------------------------------------------------------------------------------------
  if (!fast) {
    generic_slot_emit<kSmooth, kDebug>(a, P, frame, fig, u0b, v0b, gidx);
    return;
  }
  warp_copy_out(reinterpret_cast<uint8_t*>(a.out.pos) + gidx * 6, s_pos, total * 6, lane);
  if (w_rgb) warp_copy_out(a.out.rgb + gidx * 3, s_rgb, total * 3, lane);
  if (w_yuv) warp_copy_out(reinterpret_cast<uint8_t*>(a.out.yuv) + gidx * 6, s_yuv, total * 6, lane);
  if (w_part) warp_copy_out(reinterpret_cast<uint8_t*>(a.out.part) + gidx * 2, s_part, total * 2, lane);
  if (w_pix) warp_copy_out(reinterpret_cast<uint8_t*>(a.out.pix) + gidx * 4, s_pix, total * 4, lane);
  if (w_bt) warp_copy_out(a.out.btype + gidx, s_bt, total, lane);

  if (kSmooth) {
    // compact list of the type-1 boundary points of this run (order inside the list is irrelevant)
    n_boundary = __reduce_add_sync(0xFFFFFFFFu, n_boundary);
    if (n_boundary) {
      uint32_t lbase = 0;
      if (lane == 0) lbase = atomicAdd(&a.sm.blist_count[frame], n_boundary);
      lbase = __shfl_sync(0xFFFFFFFFu, lbase, 0);
      if ((uint64_t)lbase + n_boundary > a.sm.blist_cap) {
        if (lane == 0) atomicExch(a.err, 7);
      } else {
        BoundaryEntry* L = a.sm.blist + (uint64_t)frame * a.sm.blist_cap + lbase;
        uint32_t done = 0;
        for (uint32_t kb = 0; kb < total; kb += 32) {
          const uint32_t k = kb + lane;
          const bool isb = k < total && s_bt[k] == 1;
          const uint32_t mask = __ballot_sync(0xFFFFFFFFu, isb);
          if (isb) {
            const uint16_t* p = reinterpret_cast<const uint16_t*>(s_pos + k * 6);
            uint4 e;
            e.x = run_base + k;
            e.y = p[0] | ((uint32_t)p[1] << 16);
            e.z = p[2];
            e.w = 0;
            if (a.has_attr) {
              const uint16_t* c = reinterpret_cast<const uint16_t*>(s_yuv + k * 6);
              e.z |= (uint32_t)c[0] << 16;
              e.w = c[1] | ((uint32_t)c[2] << 16);
            }
            reinterpret_cast<uint4*>(L)[done + __popc(mask & ((1u << lane) - 1u))] = e;
          }
          done += __popc(mask);
        }
      }
    }
    // cell statistics of the staged points
    accumulate_cells(a, P, fig, u0b, v0b, cnt_sm, pre_sm, s_pos, s_yuv);
  }
}

// two-pass mode: exclusive scan of tile totals inside each frame (one CTA per frame)
__global__ void __launch_bounds__(256) tile_scan_kernel(const UnpackArgs a) {
  const uint32_t f = blockIdx.x;
  const uint32_t t0 = a.frame_tile_begin[f], t1 = a.frame_tile_begin[f + 1];
  __shared__ uint32_t s_w[8];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t b = t0; b < t1; b += 256) {
    const uint32_t i = b + threadIdx.x;
    const uint32_t v = i < t1 ? a.tile_total[i] : 0;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane_id() >= (uint32_t)d) incl += t;
    }
    if (lane_id() == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t wbase = s_carry;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) wbase += s_w[w];
    if (i < t1) a.tile_total[i] = wbase + incl - v;
    __syncthreads();
    if (threadIdx.x == 255) s_carry = wbase + incl;
    __syncthreads();
  }
}

// ----------------------------------------------------------------------------------------------------------------
// yuv -> rgb over a flat colour array (codec.rs:88-94)
// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) yuv_to_rgb_flat_kernel(const uint16_t* __restrict__ yuv, uint8_t* __restrict__ rgb,
                                                              uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t c = yuv_to_rgb_packed(yuv[3 * i], yuv[3 * i + 1], yuv[3 * i + 2]);
    rgb[3 * i] = (uint8_t)c; rgb[3 * i + 1] = (uint8_t)(c >> 8); rgb[3 * i + 2] = (uint8_t)(c >> 16);
  }
}

// ----------------------------------------------------------------------------------------------------------------
// K6 / K7: filter the type-1 boundary points against the trilinear blend of the 8 surrounding cell means
// ----------------------------------------------------------------------------------------------------------------
struct Nbhd { uint32_t key[8]; unsigned long long wgt[8]; unsigned long long w3; };
__device__ __forceinline__ bool neighbourhood(const GridDesc& G, const uint32_t p[3], Nbhd& N) {
  if (!(p[0] < G.th && p[1] < G.th && p[2] < G.th)) return false;
#pragma unroll
  for (int a = 0; a < 3; ++a)
    if (p[a] < G.disth || p[a] + G.disth >= G.th) return false;
  int32_t s[3]; unsigned long long wa[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const uint32_t c = cell_div(p[a], G), rem = p[a] - c * G.g;
    s[a] = (int32_t)c + (rem < G.g / 2 ? -1 : 0);
    wa[a] = 2ull * (unsigned long long)((long long)p[a] - (long long)s[a] * (long long)G.g - (long long)(G.g / 2)) + 1ull;
  }
  const unsigned long long g2 = 2ull * G.g;
  N.w3 = g2 * g2 * g2;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int dx = k & 1, dy = (k >> 1) & 1, dz = (k >> 2) & 1;
    const int32_t cx = s[0] + dx, cy = s[1] + dy, cz = s[2] + dz;
    const bool valid = cx >= 0 && cy >= 0 && cz >= 0 && (uint32_t)cx < G.w && (uint32_t)cy < G.w && (uint32_t)cz < G.w;
    N.key[k] = valid ? ((uint32_t)cx | ((uint32_t)cy << 10) | ((uint32_t)cz << 20)) : kCellEmpty;
    N.wgt[k] = (dx ? wa[0] : g2 - wa[0]) * (dy ? wa[1] : g2 - wa[1]) * (dz ? wa[2] : g2 - wa[2]);
  }
  return true;
}

__global__ void __launch_bounds__(256) smooth_filter_kernel(const __grid_constant__ UnpackArgs a) {
  const uint32_t fig = blockIdx.y;
  const uint32_t f = a.sm.group_first_frame + fig;
  const uint32_t n = min((uint64_t)a.sm.blist_count[f], a.sm.blist_cap);
  const BoundaryEntry* L = a.sm.blist + (uint64_t)f * a.sm.blist_cap;
  uint32_t moved = 0, recol = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint4 raw = *reinterpret_cast<const uint4*>(&L[i]);
    const BoundaryEntry e = *reinterpret_cast<const BoundaryEntry*>(&raw);
    const uint32_t p[3] = {e.pos[0], e.pos[1], e.pos[2]};
    const uint64_t gi = (uint64_t)f * a.out.cap + e.idx;
    Nbhd N;
    // ---- geometry (K6) ----
    if (a.sm.geo.on && neighbourhood(a.sm.geo, p, N)) {
      const GridDesc& G = a.sm.geo;
      unsigned long long C[3] = {0, 0, 0}, cntw = 0;
      bool other = false;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const GeoCell* c = N.key[j] != kCellEmpty ? cell_find<GeoCell>(G, fig, N.key[j]) : nullptr;
        uint32_t cnt = 0, s[3] = {0, 0, 0}, o[3] = {0, 0, 0};
        if (c) {
          const uint32_t cw = c->count;
          cnt = cw & ~kCellMulti; s[0] = c->sx; s[1] = c->sy; s[2] = c->sz;
          if (cnt > 0 && (cw & kCellMulti)) other = true;
          o[0] = (N.key[j] & 1023u) * G.g; o[1] = ((N.key[j] >> 10) & 1023u) * G.g; o[2] = (N.key[j] >> 20) * G.g;
        }
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
          unsigned long long m = 256ull * p[ax];
          if (cnt > 0) m = 256ull * o[ax] + (256ull * s[ax] + cnt / 2) / cnt;      // cell mean, Q8
          C[ax] += N.wgt[j] * m;
        }
        cntw += N.wgt[j] * cnt;
      }
      const unsigned long long count = cntw / N.w3;
      if (other && count > 0) {
        unsigned long long c4[3], D2 = 0;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
          c4[ax] = (C[ax] + N.w3 / 2) / N.w3;
          const long long d = (long long)(256ull * p[ax]) - (long long)c4[ax];
          D2 += (unsigned long long)(d * d);
        }
        const unsigned long long m = a.sm.thr_geo > count ? a.sm.thr_geo : count;
        const unsigned __int128 lhs = (unsigned __int128)2 * count * D2 + 65536u;
        const unsigned __int128 rhs = (unsigned __int128)262144u * m;
        if (lhs >= rhs) {
          bool changed = false;
          uint16_t q[3];
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) {
            unsigned long long r = (c4[ax] + 128) >> 8;
            if (r > 65535) r = 65535;
            q[ax] = (uint16_t)r;
            changed |= q[ax] != p[ax];
          }
          if (changed) {
            uint16_t* d = a.out.pos + gi * 3;
            d[0] = q[0]; d[1] = q[1]; d[2] = q[2];
            moved += 1;
          }
        }
      }
    }
    // ---- colour (K7), on the same pre-smoothing position ----
    if (a.sm.col.on && a.has_attr && neighbourhood(a.sm.col, p, N)) {
      const GridDesc& G = a.sm.col;
      const uint32_t col[3] = {e.yuv[0], e.yuv[1], e.yuv[2]};
      unsigned long long C[3] = {0, 0, 0};
      bool other = false;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const ColCell* c = N.key[j] != kCellEmpty ? cell_find<ColCell>(G, fig, N.key[j]) : nullptr;
        bool usable = false;
        unsigned long long mean[3] = {0, 0, 0};
        if (c && (c->count & ~kCellMulti) > 0) {
          const unsigned long long cnt = c->count & ~kCellMulti;
          if (c->count & kCellMulti) other = true;
          if (cnt > 65536ull) atomicExch(a.err, 6);     // u32 colour sums are only exact up to 65536 points per cell
          usable = true;
          const unsigned long long s[3] = {c->sy, c->su, c->sv};
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) mean[ax] = (256ull * s[ax] + cnt / 2) / cnt;
          const unsigned __int128 num = (unsigned __int128)cnt * c->sy2 - (unsigned __int128)s[0] * s[0];
          const unsigned long long tv = (unsigned long long)a.sm.thr_col_var * cnt;
          const unsigned __int128 lim = (unsigned __int128)tv * tv;
          if (num > lim) usable = false;
          const long long dy = (long long)mean[0] - (long long)(256ull * col[0]);
          if ((unsigned long long)(dy < 0 ? -dy : dy) > 256ull * a.sm.thr_col_diff) usable = false;
        }
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) C[ax] += N.wgt[j] * (usable ? mean[ax] : 256ull * col[ax]);
      }
      if (other) {
        uint32_t q[3]; unsigned long long dist = 0;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
          const unsigned long long c4 = (C[ax] + N.w3 / 2) / N.w3;
          unsigned long long r = (c4 + 128) >> 8;
          if (r > 65535) r = 65535;
          q[ax] = (uint32_t)r;
          const long long d = (long long)q[ax] - (long long)col[ax];
          dist += (unsigned long long)(d < 0 ? -d : d) * (ax == 0 ? 10u : 1u);
        }
        if (dist >= a.sm.thr_col_smooth && dist > 0) {
          const uint32_t c = yuv_to_rgb_packed(q[0], q[1], q[2]);
          uint8_t* d = a.out.rgb + gi * 3;
          d[0] = (uint8_t)c; d[1] = (uint8_t)(c >> 8); d[2] = (uint8_t)(c >> 16);
          if (a.out.yuv) { uint16_t* y = a.out.yuv + gi * 3; y[0] = (uint16_t)q[0]; y[1] = (uint16_t)q[1]; y[2] = (uint16_t)q[2]; }
          recol += 1;
        }
      }
    }
  }
  moved = __reduce_add_sync(0xFFFFFFFFu, moved);
  recol = __reduce_add_sync(0xFFFFFFFFu, recol);
  if (lane_id() == 0) {
    if (moved) atomicAdd(&a.sm.changed[f], (unsigned long long)moved);
    if (recol) atomicAdd(&a.sm.changed[a.n_frames + f], (unsigned long long)recol);
  }
}

template <typename Cell>
__device__ __forceinline__ void clear_cells(const GridDesc& G, uint32_t fig, uint64_t touched_cap) {
  if (!G.on) return;
  const uint32_t n = min((uint64_t)G.touched_count[fig], touched_cap);
  Cell* tab = reinterpret_cast<Cell*>(G.table) + (uint64_t)fig * G.slots;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    Cell z;
    memset(&z, 0, sizeof z);
    z.key = kCellEmpty;
    tab[G.touched[(uint64_t)fig * touched_cap + i]] = z;
  }
}
__global__ void __launch_bounds__(256) smooth_clear_kernel(const UnpackArgs a) {
  const uint32_t fig = blockIdx.y;
  clear_cells<GeoCell>(a.sm.geo, fig, a.sm.touched_cap);
  clear_cells<ColCell>(a.sm.col, fig, a.sm.touched_cap);
}
__global__ void smooth_reset_kernel(const UnpackArgs a) {   // after the clear: counters back to zero for the next group
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.sm.group_frames) {
    if (a.sm.geo.on) a.sm.geo.touched_count[i] = 0;
    if (a.sm.col.on) a.sm.col.touched_count[i] = 0;
    a.sm.blist_count[a.sm.group_first_frame + i] = 0;
  }
}
template <typename Cell>
__global__ void __launch_bounds__(256) table_init_kernel(Cell* tab, uint64_t n) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    Cell z;
    memset(&z, 0, sizeof z);
    z.key = kCellEmpty;
    tab[i] = z;
  }
}

// ----------------------------------------------------------------------------------------------------------------
// launch wrappers
// ----------------------------------------------------------------------------------------------------------------
static inline int after_launch() { ++g_launches; return (int)cudaGetLastError(); }

size_t unpack_smem_bytes(const UnpackArgs& a) { return (size_t)a.warp_bytes * kWarpsPerTile; }

int launch_block_to_patch(const UnpackArgs& a, uint32_t n_slots, void* stream) {
  if (n_slots == 0) return 0;
  const uint32_t blocks = (n_slots * 32 + 255) / 256;
  block_to_patch_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, n_slots, const_cast<uint32_t*>(a.block_to_patch));
  return after_launch();
}

template <int kMode, bool kSmooth, bool kDebug>
static int launch_unpack_t(const UnpackArgs& a, uint32_t tile_begin, uint32_t tile_end, size_t smem, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute((const void*)unpack_kernel<kMode, kSmooth, kDebug>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  unpack_kernel<kMode, kSmooth, kDebug><<<tile_end - tile_begin, kWarpsPerTile * 32, smem, s>>>(a, tile_begin);
  return after_launch();
}
template <int kMode>
static int launch_unpack_m(const UnpackArgs& a, bool smooth, bool debug, uint32_t t0, uint32_t t1, size_t smem, cudaStream_t s) {
  if (smooth) return debug ? launch_unpack_t<kMode, true, true>(a, t0, t1, smem, s) : launch_unpack_t<kMode, true, false>(a, t0, t1, smem, s);
  return debug ? launch_unpack_t<kMode, false, true>(a, t0, t1, smem, s) : launch_unpack_t<kMode, false, false>(a, t0, t1, smem, s);
}

int launch_unpack(const UnpackArgs& a, int mode, bool smooth, uint32_t tile_begin, uint32_t tile_end, void* stream) {
  if (tile_end <= tile_begin) return 0;
  const size_t smem = mode == 1 ? 0 : unpack_smem_bytes(a);
  const cudaStream_t s = (cudaStream_t)stream;
  const bool debug = a.out.yuv || a.out.part || a.out.pix || a.out.btype;
  if (mode == 1) return launch_unpack_t<1, false, false>(a, tile_begin, tile_end, smem, s);
  return mode == 0 ? launch_unpack_m<0>(a, smooth, debug, tile_begin, tile_end, smem, s)
                   : launch_unpack_m<2>(a, smooth, debug, tile_begin, tile_end, smem, s);
}

int launch_tile_scan(const UnpackArgs& a, void* stream) {
  if (a.n_frames == 0) return 0;
  tile_scan_kernel<<<a.n_frames, 256, 0, (cudaStream_t)stream>>>(a);
  return after_launch();
}

int launch_upsample(const UnpackArgs& a, uint8_t* occ_full, void* stream) {
  upsample_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(a, occ_full);
  return after_launch();
}

int launch_smooth_filter(const UnpackArgs& a, void* stream) {
  if (a.sm.group_frames == 0) return 0;
  const unsigned bx = (148u * 8u + a.sm.group_frames - 1) / a.sm.group_frames;
  smooth_filter_kernel<<<dim3(bx, a.sm.group_frames), 256, 0, (cudaStream_t)stream>>>(a);
  return after_launch();
}

int launch_smooth_clear(const UnpackArgs& a, void* stream) {
  if (a.sm.group_frames == 0) return 0;
  const unsigned bx = (148u * 4u + a.sm.group_frames - 1) / a.sm.group_frames;
  smooth_clear_kernel<<<dim3(bx, a.sm.group_frames), 256, 0, (cudaStream_t)stream>>>(a);
  int e = after_launch();
  if (e) return e;
  smooth_reset_kernel<<<(a.sm.group_frames + 63) / 64, 64, 0, (cudaStream_t)stream>>>(a);
  return after_launch();
}

int launch_yuv_to_rgb_flat(const uint16_t* yuv, uint8_t* rgb, uint64_t n, void* stream) {
  if (n == 0) return 0;
  uint64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  yuv_to_rgb_flat_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(yuv, rgb, n);
  return after_launch();
}
int launch_table_init(void* table, uint64_t total_slots, int is_color, void* stream) {
  if (total_slots == 0) return 0;
  if (is_color) table_init_kernel<ColCell><<<148 * 8, 256, 0, (cudaStream_t)stream>>>((ColCell*)table, total_slots);
  else table_init_kernel<GeoCell><<<148 * 8, 256, 0, (cudaStream_t)stream>>>((GeoCell*)table, total_slots);
  return after_launch();
}

}  // namespace tmc2
