// kernels.cu -- hand-written sm_100a kernels of the V-PCC rec0 reconstruction path.
//
//   K2  block_to_patch_kernel     src/codec.rs:205-250   (atomicMax == "later patch overwrites", :242-244)
//   K1  upsample_kernel           src/codec.rs:288-300   (materialised only for the stage API; fused otherwise)
//   K3+K4 unpack_kernel           src/codec.rs:352-480 (unpack loop, order, dedup), :517-565 (generate_points),
//                                 :569-658 (attribute fetch), :661-687 (YUV->RGB), src/decoder.rs:827-888 (patch maths)
//       + K5 boundary type per pixel, K6/K7 cell statistics (own spec) in the smoothing instantiation
//   K6/K7 smooth_finalize_kernel / smooth_filter_kernel / smooth_clear_kernel
//                                 grid geometry + colour smoothing of the boundary points (own integer spec, DESIGN.md;
//                                 the reference has only stubs: decoder.rs:291-299)
//
// Ordering.  The reference emits points in (patch, v0, u0, v1, u1, map) order.  A 16x16 patch block ("slot") that
// owns its canvas block emits one contiguous run, so the output position of a run is an exclusive prefix sum of
// per-slot counts in slot order.  unpack_kernel is ONE pass: a warp owns a slot, a CTA owns a tile of 8 consecutive
// slots, and tiles publish / look back their prefix through `tile_status` (single-pass chained scan with decoupled
// look-back, tile id == blockIdx.x, one scan domain per frame).  Nothing is read twice from HBM.
//
// Lane layout of a slot.  Lane l owns the 8 pixels of patch-local ranks 8l .. 8l+7 (row v1 = l/2, columns
// u1 = 8(l&1) .. +7), whatever the patch orientation: the orientation only changes WHERE those pixels are loaded from
// (an affine map with steps in {-1,0,+1}; a 16-byte vector per plane for Default, eight 2-byte loads per plane for the
// transposed / mirrored cases).  A lane therefore emits one contiguous piece of the output run and the ordered
// compaction is a plain warp scan of lane totals.
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

constexpr uint32_t kFull = 0xFFFFFFFFu;

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
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg_nc_u16(const uint16_t* p) {
  uint16_t r;
  asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(r) : "l"(p));
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
__device__ __forceinline__ void stg_cs_v4(void* p, const uint4& v) {   // streaming store: output is never re-read here
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t word_of(const uint4& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
__device__ __forceinline__ uint32_t word_of(const uint2& v, int i) { return i == 0 ? v.x : v.y; }

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
// byte-permute selectors that build a staged position from A = n | t << 16 and Bv = b (upper half zero):
// word 0 = pos[0] | pos[1] << 16, word 1 = pos[2].   source nibbles: n -> bytes 0,1 ; t -> 2,3 ; b -> 4,5 ; none -> 6,7
__device__ __forceinline__ uint32_t sel_of_src(uint32_t src) { return src == 1u ? 0x10u : src == 2u ? 0x32u : src == 3u ? 0x54u : 0x76u; }

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
// chroma sample (ChromaTerm: three integers), and a point costs three 32-bit multiply-high divisions.
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
// clamp(floor(m / 1023), 0, 255): clamp m to [0, 255*1023 + 1022] first, then floor(m / 1023) == umulhi(m, ceil(2^32/1023))
// (exact while m * 1019 < 2^32, i.e. m < 4.2e6)
__device__ __forceinline__ uint32_t quant_int(int32_t m) {
  return __umulhi((uint32_t)min(max(m, 0), 261887), 4198405u);
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
// pixel inside the 5x5 window (clipped to the image), 0 = interior.  Per-pixel form (generic slots only).
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
// voxel-cell tables (own spec; see DESIGN.md "Smoothing specification")
// ----------------------------------------------------------------------------------------------------------------
// x / G.g for x < 65536 by multiply-high (magic = ceil(2^32 / g); exact while x * g < 2^32)
__device__ __forceinline__ uint32_t cell_div(uint32_t x, const GridDesc& G) {
  return G.g_shift >= 0 ? (x >> G.g_shift) : __umulhi(x, G.magic);
}

__device__ __forceinline__ uint64_t hash_slot0(uint32_t key, const GridDesc& G) {
  const uint32_t cx = key & 1023u, cy = (key >> 10) & 1023u, cz = key >> 20;
  // 2x2x2 neighbouring cells share one 8-slot group: the filter's neighbourhood lookups stay within a few lines
  const uint32_t grp = (cx >> 1) | ((cy >> 1) << 9) | ((cz >> 1) << 18);
  const uint64_t h = ((uint64_t)grp * 0x9E3779B97F4A7C15ull) >> 24;
  return ((h << 3) | ((cx & 1u) | ((cy & 1u) << 1) | ((cz & 1u) << 2))) & (G.slots - 1);
}
// table slot of cell `key` (cx | cy<<10 | cz<<20) in the table of frame-in-group `fig`: the dense index, or find-or-claim
// in the key array of a hashed table.  kCellEmpty on failure (table full: cannot happen, slots >= 2 * points).
__device__ __forceinline__ uint32_t cell_slot(const GridDesc& G, uint32_t fig, uint32_t key, int* err) {
  if (G.identity) return (key & 1023u) + G.w * (((key >> 10) & 1023u) + G.w * (key >> 20));
  uint32_t* keys = G.keys + (uint64_t)fig * G.slots;
  uint64_t i = hash_slot0(key, G);
  for (uint64_t probe = 0; probe < G.slots; ++probe) {
    uint32_t cur = *reinterpret_cast<volatile uint32_t*>(&keys[i]);
    if (cur == kCellEmpty) cur = atomicCAS(&keys[i], kCellEmpty, key);
    if (cur == kCellEmpty || cur == key) return (uint32_t)i;
    i = (i + 1) & (G.slots - 1);
  }
  atomicExch(err, 11);
  return kCellEmpty;
}
__device__ __forceinline__ uint32_t cell_find(const GridDesc& G, uint32_t fig, uint32_t key) {
  if (G.identity) return (key & 1023u) + G.w * (((key >> 10) & 1023u) + G.w * (key >> 20));
  const uint32_t* keys = G.keys + (uint64_t)fig * G.slots;
  uint64_t i = hash_slot0(key, G);
  for (uint64_t probe = 0; probe < G.slots; ++probe) {
    const uint32_t k = keys[i];
    if (k == key) return (uint32_t)i;
    if (k == kCellEmpty) return kCellEmpty;
    i = (i + 1) & (G.slots - 1);
  }
  return kCellEmpty;
}

// fire-and-forget accumulation into a cell (REDs: nothing is read back)
__device__ __forceinline__ void geo_cell_add(const GridDesc& G, uint32_t fig, uint32_t slot, uint32_t patch, uint32_t cnt,
                                             uint32_t sx, uint32_t sy, uint32_t sz) {
  GeoCell* c = reinterpret_cast<GeoCell*>(G.table) + (uint64_t)fig * G.slots + slot;
  atomicMax(&c->pmax1, patch + 1u);
  atomicMax(&c->pminc, ~patch);
  atomicAdd(&c->cnt_sx, (unsigned long long)cnt | ((unsigned long long)sx << 32));
  atomicAdd(&c->sy_sz, (unsigned long long)sy | ((unsigned long long)sz << 32));
}
__device__ __forceinline__ void col_cell_add(const GridDesc& G, uint32_t fig, uint32_t slot, uint32_t patch, uint32_t cnt,
                                             uint32_t sy, uint32_t su, uint32_t sv, unsigned long long sy2) {
  ColCell* c = reinterpret_cast<ColCell*>(G.table) + (uint64_t)fig * G.slots + slot;
  atomicMax(&c->pmax1, patch + 1u);
  atomicMax(&c->pminc, ~patch);
  atomicAdd(&c->cnt_sy, (unsigned long long)cnt | ((unsigned long long)sy << 24));
  atomicAdd(&c->su_sv, (unsigned long long)su | ((unsigned long long)sv << 32));
  atomicAdd(&c->sy2, sy2);
}

// ----------------------------------------------------------------------------------------------------------------
// K2: block-to-patch map.  One thread per slot (= one 16x16 block of one patch).
// ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ SlotRec load_slot_rec(const SlotRec* p) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
  SlotRec r;
  r.pid = v.x; r.u0b = (uint16_t)v.y; r.v0b = (uint16_t)(v.y >> 16); r.bx = (uint16_t)v.z; r.by = (uint16_t)(v.z >> 16);
  r.ax = (int8_t)v.w; r.ay = (int8_t)(v.w >> 8); r.rx = (int8_t)(v.w >> 16); r.ry = (int8_t)(v.w >> 24);
  return r;
}

__global__ void __launch_bounds__(256) block_to_patch_kernel(const UnpackArgs a, uint32_t n_slots,
                                                             uint32_t* __restrict__ b2p) {
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_slots) return;
  const SlotRec R = load_slot_rec(a.slot_rec + slot);
  if (R.pid == kNoPatch) return;
  const uint32_t frame = a.tile_frame[slot / kWarpsPerTile];
  const uint8_t* occ_f = a.in.occ + (uint64_t)frame * a.in.occ_frame_stride;
  const uint32_t res = a.res;
  const uint32_t packed = __ldg(reinterpret_cast<const uint32_t*>(&a.patches[R.pid].orient));   // orient | aligned << 8 | ...
  bool nz = false;
  if ((packed >> 8) & 0xFFu) {
    // the res x res patch pixels are exactly the canvas block: test the low-resolution samples that cover it
    const uint32_t x0 = (uint32_t)R.bx * res, y0 = (uint32_t)R.by * res;
    const uint32_t cxa = div_prec(x0, a.prec, a.prec_shift), cxb = div_prec(x0 + res - 1, a.prec, a.prec_shift);
    const uint32_t cya = div_prec(y0, a.prec, a.prec_shift), cyb = div_prec(y0 + res - 1, a.prec, a.prec_shift);
    for (uint32_t cy = cya; cy <= cyb && !nz; ++cy) {
      const uint8_t* row = occ_f + (uint64_t)cy * a.in.occ_pitch;
      if (((cxa | (cxb + 1)) & 3u) == 0) {                           // whole aligned words (precision 4, resolution 16)
        for (uint32_t cx = cxa; cx <= cxb; cx += 4) nz |= *reinterpret_cast<const uint32_t*>(row + cx) != 0u;
      } else {
        for (uint32_t cx = cxa; cx <= cxb; ++cx) nz |= row[cx] != 0;
      }
    }
  } else {
    // reference-literal pixel mapping for the rotated / mirrored orientations (codec.rs:227-241)
    const DevPatch P = a.patches[R.pid];
    for (uint32_t i = 0; i < res * res && !nz; ++i) {
      const uint32_t v1 = i / res, u1 = i - v1 * res;
      int64_t x, y;
      patch_to_canvas(P, (int64_t)R.u0b * res + u1, (int64_t)R.v0b * res + v1, res, 1, x, y);
      nz |= occ_at(a, occ_f, (uint32_t)x, (uint32_t)y) != 0;
    }
  }
  if (nz) {
    const uint32_t local_index = __ldg(&a.patches[R.pid].local_index);
    atomicMax(&b2p[(uint64_t)frame * a.bw * a.bh + (uint64_t)R.by * a.bw + (uint64_t)R.bx], local_index + 1);
  }
}

// After K2: the slots that own their canvas block (codec.rs:379), compacted per frame in slot order.  One CTA per frame.
// The unpack kernel walks this list, so every warp of a tile has work and no plane is touched for a block that is
// skipped.
__global__ void __launch_bounds__(1024) compact_owned_kernel(const UnpackArgs a) {
  const uint32_t f = blockIdx.x;
  const uint32_t s0 = a.frame_tile_begin[f] * kWarpsPerTile, s1 = a.frame_tile_begin[f + 1] * kWarpsPerTile;
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  for (uint32_t b = s0; b < s1; b += blockDim.x) {
    const uint32_t slot = b + threadIdx.x;
    bool own = false;
    if (slot < s1) {
      const SlotRec R = load_slot_rec(a.slot_rec + slot);
      if (R.pid != kNoPatch) {
        const uint32_t local_index = __ldg(&a.patches[R.pid].local_index);
        own = a.block_to_patch[(uint64_t)f * a.bw * a.bh + (uint64_t)R.by * a.bw + R.bx] == local_index + 1;
      }
    }
    const uint32_t m = __ballot_sync(kFull, own);
    if (lane == 0) s_w[warp] = __popc(m);
    __syncthreads();
    uint32_t base = s_carry;
    for (uint32_t w = 0; w < warp; ++w) base += s_w[w];
    if (own) a.owned[s0 + base + __popc(m & ((1u << lane) - 1u))] = slot;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = base + __popc(m);
    __syncthreads();
  }
  if (threadIdx.x == 0) a.owned_count[f] = s_carry;
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
// K3+K4(+K5, + K6/K7 statistics): unpack.  Three launches:
//   count_kernel      per owned slot: how many points it emits (reads occupancy + the two geometry planes only)
//   slot_scan_kernel  per frame: exclusive prefix of those counts in slot order == where every run starts (codec.rs:482)
//   emit_kernel       per owned slot: reload, build positions + colours, write the run at its final place
// A warp owns a slot; warps never talk to each other, so there is no barrier, flag or look-back anywhere.
// ----------------------------------------------------------------------------------------------------------------

// normal coordinates (n0 | n1 << 16) of one pixel from its two geometry samples (codec.rs:534-558)
__device__ __forceinline__ uint32_t normals_of(const DevPatch& P, uint32_t s0, uint32_t s1, bool absolute_d1) {
  const uint32_t d0 = s0 >> 2, d1 = s1 >> 2;                                     // depth = sample / 4
  const uint32_t n0 = normal_coord(P, d0);
  const uint32_t n1 = absolute_d1 ? normal_coord(P, d1) : ((P.mode == 0 ? n0 + d1 : n0 - d1) & 0xFFFFu);
  return n0 | (n1 << 16);
}

// staged point k of a chunk: positions 8 B apart with one pad slot every 8 points, colours 4 B apart with one pad slot
// every 16 points, so that the copy-out (one lane per group of 8 / 16 points) is free of bank conflicts
__device__ __forceinline__ uint32_t spos_off(uint32_t k) { return (k + (k >> 3)) * 8u; }
__device__ __forceinline__ uint32_t srgb_off(uint32_t k) { return (k + (k >> 4)) * 4u; }

__device__ __forceinline__ uint32_t cell_key_of(const GridDesc& G, uint32_t x, uint32_t y, uint32_t z) {
  if (!(x < G.th && y < G.th && z < G.th)) return kCellEmpty;
  return cell_div(x, G) | (cell_div(y, G) << 10) | (cell_div(z, G) << 20);
}

// Is the block-aligned lane layout usable for this slot?  (16x16 blocks, power-of-two precision, affine steps present:
// the host zeroes the steps of reference-literal rotated patches, which take the generic per-pixel path.)
__device__ __forceinline__ bool slot_is_fast(const UnpackArgs& a, const SlotRec& R) {
  return a.res == 16 && a.prec_shift >= 0 && (R.ax != 0 || R.ay != 0);
}

// Lane layout of a slot.  Lane l owns the 8 pixels of patch-local ranks 8l .. 8l+7 (row v1 = l/2, columns
// u1 = 8(l&1) .. +7), whatever the patch orientation: the orientation only changes WHERE those pixels are loaded from.
struct LaneBlock {
  uint32_t nn[8];            // n0 | n1 << 16 per pixel
  uint32_t yy[8];            // attribute Y of map 0 | map 1 << 16 per pixel
  uint32_t cA[4], cB[4];     // chroma of map 0 / map 1 per pixel pair: U | V << 16
  uint32_t m1, m2;           // bit j set = pixel j of this lane emits >= 1 / 2 points
  int32_t xs, ys;            // canvas position of pixel 0; pixel j is (xs + ax*j, ys + ay*j)
};

template <bool kAttr>
__device__ __forceinline__ void load_lane_block(const UnpackArgs& a, const SlotRec& R, uint32_t frame, uint32_t lane,
                                                DevPatch& P, LaneBlock& L) {
  const int32_t h = (int32_t)(lane & 1u), r = (int32_t)(lane >> 1);
  const int32_t ax = R.ax, ay = R.ay, rx = R.rx, ry = R.ry;
  const int32_t cx0 = (int32_t)R.bx * 16 + ((ax < 0 || rx < 0) ? 15 : 0);
  const int32_t cy0 = (int32_t)R.by * 16 + ((ay < 0 || ry < 0) ? 15 : 0);
  const int32_t xs = cx0 + ax * 8 * h + rx * r, ys = cy0 + ay * 8 * h + ry * r;
  L.xs = xs; L.ys = ys;
  const uint16_t* geo0 = a.in.geo + (uint64_t)frame * 2 * a.in.geo_map_stride;
  const uint16_t* geo1 = geo0 + a.in.geo_map_stride;
  const bool attr = kAttr && a.has_attr;
  const uint16_t* ay0 = a.in.attr_y + (uint64_t)frame * 2 * a.in.attr_y_map_stride;
  const uint16_t* ay1 = ay0 + a.in.attr_y_map_stride;
  const uint64_t cf = (uint64_t)frame * 2 * a.in.attr_c_map_stride;
  const uint16_t* au0 = a.in.attr_u + cf; const uint16_t* au1 = au0 + a.in.attr_c_map_stride;
  const uint16_t* av0 = a.in.attr_v + cf; const uint16_t* av1 = av0 + a.in.attr_c_map_stride;
  if (ax == 1) {
    // Default-like rows: the 8 pixels are 16 contiguous bytes of every plane
    const uint32_t goff = (uint32_t)ys * a.in.geo_pitch + (uint32_t)xs;
    const uint4 g0 = ldg_nc_v4(geo0 + goff);
    const uint4 g1 = ldg_nc_v4(geo1 + goff);
    uint4 ya = {0, 0, 0, 0}, yb = {0, 0, 0, 0};
    uint2 ua = {0, 0}, va = {0, 0}, ub = {0, 0}, vb = {0, 0};
    if (attr) {
      const uint32_t yoff = (uint32_t)ys * a.in.attr_pitch_y + (uint32_t)xs;
      ya = ldg_nc_v4(ay0 + yoff); yb = ldg_nc_v4(ay1 + yoff);
      const uint32_t coff = (uint32_t)(ys >> 1) * a.in.attr_pitch_c + (uint32_t)(xs >> 1);
      ua = ldg_nc_v2(au0 + coff); va = ldg_nc_v2(av0 + coff);
      ub = ldg_nc_v2(au1 + coff); vb = ldg_nc_v2(av1 + coff);
    }
    P = a.patches[R.pid];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t sel = (j & 1) ? 0x7632u : 0x5410u;
      L.nn[j] = __byte_perm(word_of(g0, j >> 1), word_of(g1, j >> 1), sel);      // raw samples, converted below
      if (kAttr) L.yy[j] = __byte_perm(word_of(ya, j >> 1), word_of(yb, j >> 1), sel);
    }
    if (kAttr) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint32_t sel = (c & 1) ? 0x7632u : 0x5410u;
        L.cA[c] = __byte_perm(word_of(ua, c >> 1), word_of(va, c >> 1), sel);
        L.cB[c] = __byte_perm(word_of(ub, c >> 1), word_of(vb, c >> 1), sel);
      }
    }
  } else {
    // transposed / mirrored: eight 2-byte loads per plane; across the warp every load still covers whole 32-byte sectors
    const int32_t dstep = ay * (int32_t)a.in.geo_pitch + ax;
    const int32_t goff = ys * (int32_t)a.in.geo_pitch + xs;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int32_t o = goff + j * dstep;
      L.nn[j] = ldg_nc_u16(geo0 + o) | (ldg_nc_u16(geo1 + o) << 16);
    }
    if (kAttr) {
      if (attr) {
        const int32_t ystep = ay * (int32_t)a.in.attr_pitch_y + ax;
        const int32_t yoff = ys * (int32_t)a.in.attr_pitch_y + xs;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int32_t o = yoff + j * ystep;
          L.yy[j] = ldg_nc_u16(ay0 + o) | (ldg_nc_u16(ay1 + o) << 16);
        }
        const int32_t cstep = ay * (int32_t)a.in.attr_pitch_c + ax;
        const int32_t coff = (ys >> 1) * (int32_t)a.in.attr_pitch_c + (xs >> 1);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int32_t o = coff + c * cstep;
          L.cA[c] = ldg_nc_u16(au0 + o) | (ldg_nc_u16(av0 + o) << 16);
          L.cB[c] = ldg_nc_u16(au1 + o) | (ldg_nc_u16(av1 + o) << 16);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) L.yy[j] = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) { L.cA[c] = 0; L.cB[c] = 0; }
      }
    }
    P = a.patches[R.pid];
  }
  // occupancy of the 8 pixels (codec.rs:393-396: any non-zero sample counts).  Pixels j*p .. j*p+p-1 share a sample.
  uint32_t m1 = 0, m2 = 0;
  {
    const uint8_t* occ_f = a.in.occ + (uint64_t)frame * a.in.occ_frame_stride;
    const int32_t lp = a.prec_shift;
    const int32_t p = 1 << lp;
    const int32_t nb = lp >= 3 ? 1 : (8 >> lp);
    const uint32_t ones = lp >= 3 ? 0xFFu : ((1u << p) - 1u);
    for (int32_t k = 0; k < nb; ++k) {
      const int32_t j = k << lp;
      const uint32_t ox = (uint32_t)(xs + ax * j) >> lp, oy = (uint32_t)(ys + ay * j) >> lp;
      if (occ_f[(uint64_t)oy * a.in.occ_pitch + ox] != 0) m1 |= ones << j;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    L.nn[j] = normals_of(P, L.nn[j] & 0xFFFFu, L.nn[j] >> 16, a.absolute_d1);
    if ((L.nn[j] >> 16) != (L.nn[j] & 0xFFFFu)) m2 |= (m1 & (1u << j));            // codec.rs:422-428 duplicate skip
  }
  L.m1 = m1; L.m2 = m2;
}

// points of a generic slot (any resolution / precision, reference-literal rotated orientations): lane = pixel
__device__ __noinline__ uint32_t generic_slot_count(const UnpackArgs& a, const DevPatch& P, uint32_t frame, uint32_t u0b,
                                                    uint32_t v0b) {
  const uint32_t lane = lane_id(), res = a.res;
  const uint8_t* occ_f = a.in.occ + (uint64_t)frame * a.in.occ_frame_stride;
  const uint16_t* geo0 = a.in.geo + (uint64_t)frame * 2 * a.in.geo_map_stride;
  const uint16_t* geo1 = geo0 + a.in.geo_map_stride;
  const int64_t sscale = a.spec_orientation ? res : 1;
  uint32_t total = 0;
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
    total += __reduce_add_sync(kFull, c);
  }
  return total;
}

// Generic slot path: lane = pixel, 32 at a time in patch raster order, everything straight to global memory.  Rare.
template <bool kSmooth, bool kDebug>
__device__ __noinline__ void generic_slot_emit(const UnpackArgs& a, const DevPatch& P, uint32_t frame, uint32_t fig,
                                               uint32_t u0b, uint32_t v0b, uint64_t gidx, uint32_t* log_geo, uint32_t* log_col,
                                               uint32_t* n_log /* shared: [0] geo, [1] col */) {
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
      const uint32_t tt = __shfl_up_sync(kFull, incl, d);
      if (lane >= (uint32_t)d) incl += tt;
    }
    const uint32_t chunk_total = __shfl_sync(kFull, incl, 31);
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
        if (a.sm.geo.on) {
          const uint32_t key = cell_key_of(a.sm.geo, X, Yc, Z);
          if (key != kCellEmpty) {
            const uint32_t cs = cell_slot(a.sm.geo, fig, key, a.err);
            if (cs != kCellEmpty) {
              const uint32_t g = a.sm.geo.g;
              geo_cell_add(a.sm.geo, fig, cs, P.local_index, 1, X - (key & 1023u) * g, Yc - ((key >> 10) & 1023u) * g, Z - (key >> 20) * g);
              log_geo[atomicAdd(&n_log[0], 1u)] = cs;
            }
          }
        }
        if (a.sm.col.on && a.has_attr && bt == 2) {
          const uint32_t key = cell_key_of(a.sm.col, X, Yc, Z);
          if (key != kCellEmpty) {
            const uint32_t cs = cell_slot(a.sm.col, fig, key, a.err);
            if (cs != kCellEmpty) {
              col_cell_add(a.sm.col, fig, cs, P.local_index, 1, Y, U, V, (unsigned long long)Y * Y);
              log_col[atomicAdd(&n_log[1], 1u)] = cs;
            }
          }
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

// ---- pass 1: points per owned slot ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerTile * 32) count_kernel(const __grid_constant__ UnpackArgs a, uint32_t tile_offset) {
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t tile = blockIdx.x + tile_offset;
  const uint32_t frame = a.tile_frame[tile];
  const uint32_t first_tile = a.frame_tile_begin[frame];
  const uint32_t pos_in_frame = (tile - first_tile) * kWarpsPerTile + warp;
  if (pos_in_frame >= a.owned_count[frame]) return;
  const uint32_t lpos = tile * kWarpsPerTile + warp;
  const SlotRec R = load_slot_rec(a.slot_rec + a.owned[lpos]);
  uint32_t total;
  DevPatch P;
  if (slot_is_fast(a, R)) {
    LaneBlock L;
    load_lane_block<false>(a, R, frame, lane, P, L);
    total = __reduce_add_sync(kFull, __popc(L.m1) + __popc(L.m2));
  } else {
    P = a.patches[R.pid];
    total = generic_slot_count(a, P, frame, R.u0b, R.v0b);
  }
  if (lane == 0) a.slot_total[lpos] = total;
}

// ---- pass 2: where every run starts.  One CTA per frame; the scan domain is the frame's owned-slot list ---------------------
__global__ void __launch_bounds__(1024) slot_scan_kernel(const UnpackArgs a) {
  const uint32_t f = blockIdx.x;
  const uint32_t s0 = a.frame_tile_begin[f] * kWarpsPerTile, n = a.owned_count[f];
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  for (uint32_t b = 0; b < n; b += blockDim.x) {
    const uint32_t i = b + threadIdx.x;
    const uint32_t v = i < n ? a.slot_total[s0 + i] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, incl, d);
      if (lane >= (uint32_t)d) incl += t;
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t wbase = s_carry;
    for (uint32_t w = 0; w < warp; ++w) wbase += s_w[w];
    if (i < n) a.slot_base[s0 + i] = wbase + incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = wbase + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) a.frame_count[f] = s_carry;              // tile.total_number_of_regular_points, codec.rs:482
}

// ---- copy-out of one staged chunk ----------------------------------------------------------------------------------------
// The chunk starts at point `run_base` of its frame slab (slabs are 16-byte aligned and hold a multiple of 16 points).
// Packed output is written in aligned groups: 8 points = 48 B = three 16-byte vectors for positions (a group starts at a
// point index that is a multiple of 8), 16 points = 48 B for colours (multiple of 16).  One lane assembles one group
// from the padded staging with byte permutes; the few points before the first / after the last full group go out as
// 2-byte / 1-byte stores.
__device__ __forceinline__ void copy_out_pos(uint16_t* __restrict__ gpos /* frame slab */, uint32_t run_base, uint32_t total,
                                             const uint8_t* s_pos, uint32_t lane) {
  const uint32_t head = min(total, (0u - run_base) & 7u);
  const uint32_t groups = (total - head) >> 3;
  const uint32_t tail0 = head + (groups << 3);                  // first point of the tail
  for (uint32_t g = lane; g < groups; g += 32) {
    const uint32_t k0 = head + (g << 3);
    // staged points k0 .. k0+7: slots are consecutive except for one pad slot after every 8th staged point
    const uint32_t base = spos_off(k0);
    const uint32_t brk = 8u - (k0 & 7u);                          // points at local index >= brk sit one slot further
    uint2 p[8];
#pragma unroll
    for (uint32_t i = 0; i < 8; ++i)
      p[i] = *reinterpret_cast<const uint2*>(s_pos + base + i * 8u + (i >= brk ? 8u : 0u));
    // stream words: A(P) = x|y<<16 ; B(P,Pn) = z | Pn.x << 16 ; C(P) = y | z << 16   (two points = three words)
    uint4 v0, v1, v2;
    v0.x = p[0].x;                                    v0.y = __byte_perm(p[0].y, p[1].x, 0x5410);
    v0.z = __byte_perm(p[1].x, p[1].y, 0x5432);      v0.w = p[2].x;
    v1.x = __byte_perm(p[2].y, p[3].x, 0x5410);      v1.y = __byte_perm(p[3].x, p[3].y, 0x5432);
    v1.z = p[4].x;                                    v1.w = __byte_perm(p[4].y, p[5].x, 0x5410);
    v2.x = __byte_perm(p[5].x, p[5].y, 0x5432);      v2.y = p[6].x;
    v2.z = __byte_perm(p[6].y, p[7].x, 0x5410);      v2.w = __byte_perm(p[7].x, p[7].y, 0x5432);
    uint4* dst = reinterpret_cast<uint4*>(gpos + (uint64_t)(run_base + k0) * 3);
    stg_cs_v4(dst, v0); stg_cs_v4(dst + 1, v1); stg_cs_v4(dst + 2, v2);
  }
  const uint32_t n16 = 3u * (head + (total - tail0));           // u16 elements outside full groups (<= 42)
  for (uint32_t i = lane; i < n16; i += 32) {
    const uint32_t pt = i / 3u, c = i - pt * 3u;
    const uint32_t k = pt < head ? pt : tail0 + (pt - head);
    gpos[(uint64_t)(run_base + k) * 3 + c] = *reinterpret_cast<const uint16_t*>(s_pos + spos_off(k) + c * 2u);
  }
}
__device__ __forceinline__ void copy_out_rgb(uint8_t* __restrict__ grgb /* frame slab */, uint32_t run_base, uint32_t total,
                                             const uint8_t* s_rgb, uint32_t lane) {
  const uint32_t head = min(total, (0u - run_base) & 15u);
  const uint32_t groups = (total - head) >> 4;
  const uint32_t tail0 = head + (groups << 4);
  for (uint32_t g = lane; g < groups; g += 32) {
    const uint32_t k0 = head + (g << 4);
    const uint32_t base = srgb_off(k0);
    const uint32_t brk = 16u - (k0 & 15u);
    uint32_t p[16];
#pragma unroll
    for (uint32_t i = 0; i < 16; ++i)
      p[i] = *reinterpret_cast<const uint32_t*>(s_rgb + base + i * 4u + (i >= brk ? 4u : 0u));
    // four points (r,g,b,0 each) = three stream words
    uint32_t w[12];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      w[3 * q + 0] = __byte_perm(p[4 * q + 0], p[4 * q + 1], 0x4210);
      w[3 * q + 1] = __byte_perm(p[4 * q + 1], p[4 * q + 2], 0x5421);
      w[3 * q + 2] = __byte_perm(p[4 * q + 2], p[4 * q + 3], 0x6542);
    }
    uint4* dst = reinterpret_cast<uint4*>(grgb + (uint64_t)(run_base + k0) * 3);
    stg_cs_v4(dst, make_uint4(w[0], w[1], w[2], w[3]));
    stg_cs_v4(dst + 1, make_uint4(w[4], w[5], w[6], w[7]));
    stg_cs_v4(dst + 2, make_uint4(w[8], w[9], w[10], w[11]));
  }
  const uint32_t n8 = 3u * (head + (total - tail0));            // bytes outside full groups (<= 90)
  for (uint32_t i = lane; i < n8; i += 32) {
    const uint32_t pt = i / 3u, c = i - pt * 3u;
    const uint32_t k = pt < head ? pt : tail0 + (pt - head);
    grgb[(uint64_t)(run_base + k) * 3 + c] = s_rgb[srgb_off(k) + c];
  }
}

// one row of the 20x20 occupancy bitmap (block + 2-pixel margin, patch-local axes): bit cc = pixel (cc-2, rr) of the block
// is occupied, or lies outside the image (the 5x5 test ignores those).  Walks the row one occupancy cell at a time.
__device__ __forceinline__ uint32_t bitmap_row(const UnpackArgs& a, const uint8_t* occ_f, int32_t x0, int32_t y0, int32_t sx,
                                               int32_t sy) {
  // pixel cc of the row is (x0 + sx*cc, y0 + sy*cc); exactly one of sx, sy is non-zero
  const int32_t W = (int32_t)a.W, H = (int32_t)a.H, lp = a.prec_shift;
  const bool along_x = sx != 0;
  const int32_t fixed = along_x ? y0 : x0, flim = along_x ? H : W;
  if (fixed < 0 || fixed >= flim) return 0xFFFFFu;
  const int32_t m0 = along_x ? x0 : y0, st = along_x ? sx : sy, mlim = along_x ? W : H;
  const uint8_t* rowp = along_x ? occ_f + (uint64_t)(fixed >> lp) * a.in.occ_pitch : occ_f + (fixed >> lp);
  const uint32_t cell_stride = along_x ? 1u : a.in.occ_pitch;
  const int32_t pm = (1 << lp) - 1;
  uint32_t bits = 0;
  int32_t cc = 0;
  while (cc < 20) {
    const int32_t m = m0 + st * cc;
    if (m < 0 || m >= mlim) { bits |= 1u << cc; ++cc; continue; }
    const int32_t run = min(20 - cc, st > 0 ? (pm + 1) - (m & pm) : (m & pm) + 1);     // pixels left inside this cell
    if (rowp[(uint64_t)(m >> lp) * cell_stride] != 0) bits |= ((1u << run) - 1u) << cc;
    cc += run;
  }
  // a cell run may have crossed the image edge (image size not a multiple of the precision): those pixels count as set
  return bits | (st > 0 ? (m0 + 19 >= mlim ? (0xFFFFFu << max(mlim - m0, 0)) & 0xFFFFFu : 0u)
                        : (m0 - 19 < 0 ? (0xFFFFFu << max(m0 + 1, 0)) & 0xFFFFFu : 0u));
}

// ---- pass 3: emit ------------------------------------------------------------------------------------------------------------
// Per warp (slot): (1) load the block in the lane layout, (2) spill it into per-pixel tables in shared memory and build
// the list "output point k <- (pixel rank, map)", (3) POINT-parallel loop: lane = output point, dense, rolled (small code):
// position, colour, and in the smoothing instantiation boundary class, cell statistics and the boundary list, staged in
// chunks of <= 256 points, (4) aligned copy-out of each chunk.
template <bool kSmooth, bool kDebug>
__global__ void __launch_bounds__(kWarpsPerTile * 32, 3) emit_kernel(const __grid_constant__ UnpackArgs a, uint32_t tile_offset) {
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t tile = blockIdx.x + tile_offset;
  const uint32_t frame = a.tile_frame[tile];
  const uint32_t first_tile = a.frame_tile_begin[frame];
  const uint32_t pos_in_frame = (tile - first_tile) * kWarpsPerTile + warp;   // index into the frame's owned-slot list
  const uint32_t lpos = tile * kWarpsPerTile + warp;                          // global position (per-slot arrays, logs)
  uint32_t total = 0;
  const bool active = pos_in_frame < a.owned_count[frame];
  if (active) total = a.slot_total[lpos];
  uint32_t n_log_geo = 0, n_log_col = 0;
  if (total != 0) {
    const uint32_t run_base = a.slot_base[lpos];
    const SlotRec R = load_slot_rec(a.slot_rec + a.owned[lpos]);
    const uint32_t fig = kSmooth ? frame - a.sm.group_first_frame : 0u;   // frame inside the smoothing group
    uint32_t* log_geo = nullptr; uint32_t* log_col = nullptr;
    if (kSmooth) {
      const uint64_t ls = (uint64_t)(lpos - a.sm.group_first_slot) * a.sm.log_stride;
      if (a.sm.geo.on) log_geo = a.sm.geo.log + ls;
      if (a.sm.col.on) log_col = a.sm.col.log + ls;
    }
    uint8_t* wsm = smem + (size_t)warp * kWarpSmemBytes;
    uint32_t* s_nn = reinterpret_cast<uint32_t*>(wsm + kOffNn);
    uint32_t* s_yy = reinterpret_cast<uint32_t*>(wsm + kOffYy);
    uint4* s_term = reinterpret_cast<uint4*>(wsm + kOffTerm);
    uint16_t* s_src = reinterpret_cast<uint16_t*>(wsm + kOffSrc);
    uint32_t* s_bt = reinterpret_cast<uint32_t*>(wsm + kOffBt);
    uint32_t* s_bmp = reinterpret_cast<uint32_t*>(wsm + kOffBmp);
    uint8_t* s_pos = wsm + kOffPos;
    uint8_t* s_rgb = wsm + kOffRgb;
    DevPatch P;

    if ((uint64_t)run_base + total > a.out.cap) {                   // cannot happen for footprints inside the canvas
      if (lane == 0) atomicExch(a.err, 7);
    } else if (!slot_is_fast(a, R)) {
      P = a.patches[R.pid];
      if (lane == 0) { s_bmp[0] = 0; s_bmp[1] = 0; }
      __syncwarp();
      generic_slot_emit<kSmooth, kDebug>(a, P, frame, fig, R.u0b, R.v0b, (uint64_t)frame * a.out.cap + run_base, log_geo,
                                         log_col, s_bmp);
      __syncwarp();
      n_log_geo = reinterpret_cast<volatile uint32_t*>(s_bmp)[0];
      n_log_col = reinterpret_cast<volatile uint32_t*>(s_bmp)[1];
    } else {
      const int32_t h = (int32_t)(lane & 1u), r = (int32_t)(lane >> 1);
      const int32_t ax = R.ax, ay = R.ay, rx = R.rx, ry = R.ry;
      const bool w_rgb = a.out.rgb != nullptr;
      const bool want_bt = kSmooth || (kDebug && a.out.btype != nullptr);
      int32_t cx0, cy0;                                              // canvas pixel of patch-local (0,0) of the block
      {
        // ---- (1) + (2): lane layout -> tables ----------------------------------------------------------------------
        LaneBlock L;
        load_lane_block<true>(a, R, frame, lane, P, L);
        cx0 = L.xs - ax * 8 * h - rx * r; cy0 = L.ys - ay * 8 * h - ry * r;
        *reinterpret_cast<uint4*>(s_nn + 8 * lane) = make_uint4(L.nn[0], L.nn[1], L.nn[2], L.nn[3]);
        *reinterpret_cast<uint4*>(s_nn + 8 * lane + 4) = make_uint4(L.nn[4], L.nn[5], L.nn[6], L.nn[7]);
        if (a.has_attr) {
          *reinterpret_cast<uint4*>(s_yy + 8 * lane) = make_uint4(L.yy[0], L.yy[1], L.yy[2], L.yy[3]);
          *reinterpret_cast<uint4*>(s_yy + 8 * lane + 4) = make_uint4(L.yy[4], L.yy[5], L.yy[6], L.yy[7]);
          // chroma terms, once per (chroma sample, map): the two lanes of a row pair read the same 4 samples; the even
          // row evaluates map 0, the odd row map 1.  Entry = {ir, ig, ib << 1 | flagged, U | V << 16}.
          const bool odd = (lane & 2u) != 0;
          const uint32_t e0 = (((uint32_t)r >> 1) * 8u + 4u * (uint32_t)h) * 2u + (odd ? 1u : 0u);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const uint32_t uv = odd ? L.cB[cc] : L.cA[cc];                            // decoder.rs:976-977
            const ChromaTerm t = chroma_term(uv & 0xFFFFu, uv >> 16);
            s_term[e0 + 2 * cc] = make_uint4((uint32_t)t.ir, (uint32_t)t.ig, ((uint32_t)t.ib << 1) | t.flagged, uv);
          }
        }
        // output point k of the slot <- (pixel rank, map); this lane's points are consecutive
        const uint32_t c = __popc(L.m1) + __popc(L.m2);
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t t = __shfl_up_sync(kFull, incl, d);
          if (lane >= (uint32_t)d) incl += t;
        }
        uint32_t k = incl - c;
        const uint32_t m1 = L.m1, m2 = L.m2;
#pragma unroll 1
        for (uint32_t j = 0; j < 8; ++j) {
          if ((m1 >> j) & 1u) {
            const uint32_t e = (8u * lane + j) << 1;
            s_src[k++] = (uint16_t)e;
            if ((m2 >> j) & 1u) s_src[k++] = (uint16_t)(e | 1u);
          }
        }
        // ---- K5: boundary classes from a 20x20 occupancy bitmap of the block and its 2-pixel margin ---------------
        if (want_bt) {
          const int32_t W = (int32_t)a.W, H = (int32_t)a.H;
          const uint8_t* occ_f = a.in.occ + (uint64_t)frame * a.in.occ_frame_stride;
          if (lane < 20) {
            const int32_t rr = (int32_t)lane - 2;
            s_bmp[lane] = bitmap_row(a, occ_f, cx0 - 2 * ax + rx * rr, cy0 - 2 * ay + ry * rr, ax, ay);
          }
          __syncwarp();
          const uint32_t r0 = s_bmp[r], r1 = s_bmp[r + 1], r2 = s_bmp[r + 2], r3 = s_bmp[r + 3], r4 = s_bmp[r + 4];
          const uint32_t cross = r1 & r3 & (r2 >> 1) & (r2 << 1);    // bit c: the four neighbours of column c are occupied
          const uint32_t all5 = r0 & r1 & r2 & r3 & r4;
          const uint32_t full = all5 & (all5 >> 1) & (all5 >> 2) & (all5 << 1) & (all5 << 2);
          const uint32_t sh = 8u * (uint32_t)h + 2u;
          uint32_t border = 0;
          if (R.bx == 0 || R.by == 0 || ((int32_t)R.bx + 1) * 16 >= W || ((int32_t)R.by + 1) * 16 >= H) {
#pragma unroll 1
            for (int j = 0; j < 8; ++j) {
              const int32_t x = L.xs + ax * j, y = L.ys + ay * j;
              if (x == 0 || y == 0 || x == W - 1 || y == H - 1) border |= 1u << j;
            }
          }
          const uint32_t bt1 = ((~(cross >> sh)) & 0xFFu) | border;
          const uint32_t bt2 = (~(full >> sh)) & 0xFFu & ~bt1;
          s_bt[lane] = bt1 | (bt2 << 8);     // bit j: pixel j is type 1 ; bit 8+j: type 2 (meaningful where occupied)
        }
      }
      __syncwarp();

      // ---- (3): point-parallel ------------------------------------------------------------------------------------------
      // generate_point (decoder.rs:871-878) stores normal, tangent, bitangent in that order; the selectors reproduce
      // "later stores overwrite earlier ones" and leave unset coordinates at 0
      const uint32_t s0 = sel_of_src(axis_source(P, 0)), s1 = sel_of_src(axis_source(P, 1)), s2 = sel_of_src(axis_source(P, 2));
      const uint32_t selA = s0 | (s1 << 8), selB = s2 | 0x7600u;
      const uint32_t T00 = (uint32_t)R.u0b * 16u * P.lod_x + P.u1, B00 = (uint32_t)R.v0b * 16u * P.lod_y + P.v1;   // decoder.rs:875-876
      const uint32_t lodx = P.lod_x, lody = P.lod_y, patch = P.local_index;
      uint16_t* gpos = a.out.pos + (uint64_t)frame * a.out.cap * 3;
      uint8_t* grgb = w_rgb ? a.out.rgb + (uint64_t)frame * a.out.cap * 3 : nullptr;
      const uint64_t gidx = (uint64_t)frame * a.out.cap + run_base;

      for (uint32_t c0 = 0; c0 < total;) {
        // a chunk ends on a point whose index in the frame is a multiple of 16 (so the next one starts group-aligned)
        uint32_t c1 = c0 + kChunkPoints - ((run_base + c0 + kChunkPoints) & 15u);
        if (c1 > total) c1 = total;
#pragma unroll 1
        for (uint32_t kb = c0; kb < c1; kb += 32) {
          const uint32_t k = kb + lane;
          const bool valid = k < c1;
          uint32_t w0 = 0, w1 = 0, Y = 0, uv = 0, rank = 0, map = 0;
          if (valid) {
            const uint32_t src = s_src[k];
            rank = src >> 1; map = src & 1u;
            const uint32_t n = (s_nn[rank] >> (16u * map)) & 0xFFFFu;
            const uint32_t u1 = rank & 15u, v1 = rank >> 4;
            const uint32_t t = (T00 + u1 * lodx) & 0xFFFFu, b = (B00 + v1 * lody) & 0xFFFFu;
            const uint32_t A = n | (t << 16);
            w0 = __byte_perm(A, b, selA); w1 = __byte_perm(A, b, selB);
            *reinterpret_cast<uint2*>(s_pos + spos_off(k - c0)) = make_uint2(w0, w1);
            if (a.has_attr) {
              Y = (s_yy[rank] >> (16u * map)) & 0xFFFFu;                                     // codec.rs:637-640
              const uint4 te = s_term[(((v1 >> 1) * 8u + (u1 >> 1)) << 1) | map];
              uv = te.w;
              if (w_rgb) {
                ChromaTerm ct;
                ct.ir = (int32_t)te.x; ct.ig = (int32_t)te.y; ct.ib = (int32_t)te.z >> 1; ct.flagged = te.z & 1u;
                *reinterpret_cast<uint32_t*>(s_rgb + srgb_off(k - c0)) = yuv_to_rgb_term(Y, uv & 0xFFFFu, uv >> 16, ct);
              }
            }
          }
          uint32_t bt = 0;
          if (kSmooth || kDebug) {
            if (valid && want_bt) {
              const uint32_t bw = s_bt[rank >> 3], j = rank & 7u;
              bt = ((bw >> j) & 1u) ? 1u : ((bw >> (8u + j)) & 1u) ? 2u : 0u;
            }
          }
          if (kDebug && valid) {                                       // streams only the stage API / tests ask for
            const uint64_t gk = gidx + k;
            if (a.out.yuv && a.has_attr) {
              uint16_t* q = a.out.yuv + gk * 3;
              q[0] = (uint16_t)Y; q[1] = (uint16_t)(uv & 0xFFFFu); q[2] = (uint16_t)(uv >> 16);
            }
            if (a.out.part) a.out.part[gk] = (uint16_t)patch;                          // codec.rs:452
            if (a.out.pix) {                                                            // codec.rs:463-472
              const int32_t u1 = (int32_t)(rank & 15u), v1 = (int32_t)(rank >> 4);
              a.out.pix[gk] = (uint32_t)(cx0 + ax * u1 + rx * v1) | ((uint32_t)(cy0 + ay * u1 + ry * v1) << 15) | (map << 30);
            }
            if (a.out.btype) a.out.btype[gk] = (uint8_t)bt;
          }
          if (kSmooth) {
            const uint32_t X = w0 & 0xFFFFu, Yc = w0 >> 16, Z = w1 & 0xFFFFu;
            // K6 statistics: geometry cells over ALL points.  32 consecutive points hold a few runs of equal cell key
            // (a row of the block crosses a cell every g pixels): segmented scan, the last lane of a run flushes it.
            if (a.sm.geo.on) {
              const GridDesc& G = a.sm.geo;
              const uint32_t key = valid ? cell_key_of(G, X, Yc, Z) : kCellEmpty;
              uint32_t v0 = 0, v1 = 0;
              if (key != kCellEmpty) {
                const uint32_t g = G.g;
                v0 = 1u | ((X - (key & 1023u) * g) << 16);                              // count | sum rel x
                v1 = (Yc - ((key >> 10) & 1023u) * g) | ((Z - (key >> 20) * g) << 16);  // sum rel y | sum rel z
              }
              const uint32_t kprev = __shfl_up_sync(kFull, key, 1), knext = __shfl_down_sync(kFull, key, 1);
              uint32_t fl = (lane == 0 || kprev != key) ? 1u : 0u;
#pragma unroll
              for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o0 = __shfl_up_sync(kFull, v0, d), o1 = __shfl_up_sync(kFull, v1, d);
                const uint32_t of = __shfl_up_sync(kFull, fl, d);
                if (lane >= (uint32_t)d && !fl) { v0 += o0; v1 += o1; fl = of; }
              }
              const bool tail = key != kCellEmpty && (lane == 31 || knext != key);
              uint32_t cs = kCellEmpty;
              if (tail) {
                cs = cell_slot(G, fig, key, a.err);
                if (cs != kCellEmpty) geo_cell_add(G, fig, cs, patch, v0 & 0xFFFFu, v0 >> 16, v1 & 0xFFFFu, v1 >> 16);
              }
              const uint32_t fm = __ballot_sync(kFull, cs != kCellEmpty);
              if (cs != kCellEmpty) log_geo[n_log_geo + __popc(fm & ((1u << lane) - 1u))] = cs;
              n_log_geo += __popc(fm);
            }
            // K7 statistics: colour cells over the type-2 (second ring) points
            if (a.sm.col.on && a.has_attr && __any_sync(kFull, bt == 2u)) {
              const GridDesc& G = a.sm.col;
              uint32_t cs = kCellEmpty;
              if (bt == 2u) {
                const uint32_t key = cell_key_of(G, X, Yc, Z);
                if (key != kCellEmpty) {
                  cs = cell_slot(G, fig, key, a.err);
                  if (cs != kCellEmpty) col_cell_add(G, fig, cs, patch, 1, Y, uv & 0xFFFFu, uv >> 16, (unsigned long long)Y * Y);
                }
              }
              const uint32_t fm = __ballot_sync(kFull, cs != kCellEmpty);
              if (cs != kCellEmpty) log_col[n_log_col + __popc(fm & ((1u << lane) - 1u))] = cs;
              n_log_col += __popc(fm);
            }
            // compact list of the type-1 boundary points (order inside the list is irrelevant)
            const uint32_t bm = __ballot_sync(kFull, bt == 1u);
            if (bm) {
              uint32_t lbase = 0;
              if (lane == 0) lbase = atomicAdd(&a.sm.blist_count[frame], (uint32_t)__popc(bm));
              lbase = __shfl_sync(kFull, lbase, 0);
              if ((uint64_t)lbase + __popc(bm) > a.sm.blist_cap) {
                if (lane == 0) atomicExch(a.err, 7);
              } else if (bt == 1u) {
                uint4 e;
                e.x = run_base + k;
                e.y = w0;                                  // pos[0] | pos[1] << 16
                e.z = (w1 & 0xFFFFu) | (Y << 16);          // pos[2] | Y << 16
                e.w = uv;                                  // U | V << 16
                reinterpret_cast<uint4*>(a.sm.blist + (uint64_t)frame * a.sm.blist_cap)[lbase + __popc(bm & ((1u << lane) - 1u))] = e;
              }
            }
          }
        }
        __syncwarp();
        // ---- (4): copy-out of the chunk ----------------------------------------------------------------------------------
        copy_out_pos(gpos, run_base + c0, c1 - c0, s_pos, lane);
        if (w_rgb) copy_out_rgb(grgb, run_base + c0, c1 - c0, s_rgb, lane);
        __syncwarp();
        c0 = c1;
      }
    }
  }
  if (kSmooth && lane == 0) {                                      // every position of the group reports its log sizes
    if (a.sm.geo.on) a.sm.geo.log_count[lpos - a.sm.group_first_slot] = n_log_geo;
    if (a.sm.col.on) a.sm.col.log_count[lpos - a.sm.group_first_slot] = n_log_col;
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
// K6 / K7 finalize: once per touched cell, sums -> Q8 means (and the colour cell's luminance-variance verdict), so that
// the filter (8 cells per boundary point per grid) does no division.  One warp per unpack slot of the group walks that
// slot's log; a cell logged by several slots is finalized more than once (geometry: idempotent, means live in their own
// field; colour: claimed with an atomic flag because the means replace the sums).
// ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mean_q8_u32(uint32_t s, uint32_t cnt) {       // (256*s + cnt/2) / cnt
  const unsigned long long num = 256ull * s + (cnt >> 1);
  return num < (1ull << 32) ? (uint32_t)num / cnt : (uint32_t)(num / cnt);
}
__global__ void __launch_bounds__(256) smooth_finalize_kernel(const __grid_constant__ UnpackArgs a) {
  const uint32_t ls = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;       // slot inside the group
  if (ls >= a.sm.group_slots) return;
  const uint32_t lane = lane_id();
  const uint32_t frame = a.tile_frame[(a.sm.group_first_slot + ls) / kWarpsPerTile];
  const uint32_t fig = frame - a.sm.group_first_frame;
  if (a.sm.geo.on) {
    const GridDesc& G = a.sm.geo;
    const uint32_t n = min(G.log_count[ls], a.sm.log_stride);
    const uint32_t* log = G.log + (uint64_t)ls * a.sm.log_stride;
    GeoCell* tab = reinterpret_cast<GeoCell*>(G.table) + (uint64_t)fig * G.slots;
    for (uint32_t i = lane; i < n; i += 32) {
      GeoCell* c = tab + log[i];
      const unsigned long long w0 = c->cnt_sx, w1 = c->sy_sz;
      const uint32_t cnt = (uint32_t)w0;
      if (cnt == 0) continue;
      const unsigned long long mx = mean_q8_u32((uint32_t)(w0 >> 32), cnt), my = mean_q8_u32((uint32_t)w1, cnt),
                               mz = mean_q8_u32((uint32_t)(w1 >> 32), cnt);
      c->mean = mx | (my << 16) | (mz << 32);
    }
  }
  if (a.sm.col.on) {
    const GridDesc& G = a.sm.col;
    const uint32_t n = min(G.log_count[ls], a.sm.log_stride);
    const uint32_t* log = G.log + (uint64_t)ls * a.sm.log_stride;
    ColCell* tab = reinterpret_cast<ColCell*>(G.table) + (uint64_t)fig * G.slots;
    for (uint32_t i = lane; i < n; i += 32) {
      ColCell* c = tab + log[i];
      if (atomicOr(&c->pmax1, kCellFinal) & kCellFinal) continue;          // somebody else finalizes / finalized it
      const unsigned long long w0 = c->cnt_sy, w1 = c->su_sv, sy2 = c->sy2;
      const unsigned long long cnt = w0 & 0xFFFFFFull, sy = w0 >> 24, su = w1 & 0xFFFFFFFFull, sv = w1 >> 32;
      if (cnt > 65536ull) atomicExch(a.err, 6);     // the packed U / V sums are only exact up to 65536 points per cell
      const unsigned long long my = (256ull * sy + cnt / 2) / cnt, mu = (256ull * su + cnt / 2) / cnt,
                               mv = (256ull * sv + cnt / 2) / cnt;
      // luminance variation: var(Y) = (cnt*sumY2 - sumY^2)/cnt^2 must not exceed t_var^2
      const unsigned __int128 num = (unsigned __int128)cnt * sy2 - (unsigned __int128)sy * sy;
      const unsigned long long tv = (unsigned long long)a.sm.thr_col_var * cnt;
      const unsigned __int128 lim = (unsigned __int128)tv * tv;
      c->cnt_sy = cnt | (my << 32);
      c->su_sv = mu | (mv << 32);
      c->sy2 = num > lim ? 0ull : 1ull;
    }
  }
}

// ----------------------------------------------------------------------------------------------------------------
// K6 / K7: filter the type-1 boundary points against the trilinear blend of the 8 surrounding cell means
// ----------------------------------------------------------------------------------------------------------------
struct Nbhd { uint32_t key[8]; uint32_t wgt[8]; uint32_t w3; };
// weights fit 32 bits: (2g)^3 <= 2^27 for g <= 256
__device__ __forceinline__ bool neighbourhood(const GridDesc& G, const uint32_t p[3], Nbhd& N) {
  if (!(p[0] < G.th && p[1] < G.th && p[2] < G.th)) return false;
#pragma unroll
  for (int a = 0; a < 3; ++a)
    if (p[a] < G.disth || p[a] + G.disth >= G.th) return false;
  int32_t s[3]; uint32_t wa[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const uint32_t c = cell_div(p[a], G), rem = p[a] - c * G.g;
    s[a] = (int32_t)c + (rem < G.g / 2 ? -1 : 0);
    wa[a] = 2u * (uint32_t)((int32_t)p[a] - s[a] * (int32_t)G.g - (int32_t)(G.g / 2)) + 1u;
  }
  const uint32_t g2 = 2u * G.g;
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
// x / w3 where w3 = (2g)^3: a shift when g is a power of two
__device__ __forceinline__ unsigned long long div_w3(unsigned long long x, const GridDesc& G, uint32_t w3) {
  return G.g_shift >= 0 ? (x >> (3 * (G.g_shift + 1))) : (x / w3);
}

__global__ void __launch_bounds__(256) smooth_filter_kernel(const __grid_constant__ UnpackArgs a) {
  const uint32_t fig = blockIdx.y;
  const uint32_t f = a.sm.group_first_frame + fig;
  const uint32_t n = min((uint64_t)a.sm.blist_count[f], a.sm.blist_cap);
  const BoundaryEntry* L = a.sm.blist + (uint64_t)f * a.sm.blist_cap;
  uint32_t moved = 0, recol = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint4 raw = *reinterpret_cast<const uint4*>(&L[i]);
    const uint32_t p[3] = {raw.y & 0xFFFFu, raw.y >> 16, raw.z & 0xFFFFu};
    const uint32_t col[3] = {raw.z >> 16, raw.w & 0xFFFFu, raw.w >> 16};
    const uint64_t gi = (uint64_t)f * a.out.cap + raw.x;
    Nbhd N;
    // ---- geometry (K6) ----
    if (a.sm.geo.on && neighbourhood(a.sm.geo, p, N)) {
      const GridDesc& G = a.sm.geo;
      const GeoCell* tab = reinterpret_cast<const GeoCell*>(G.table) + (uint64_t)fig * G.slots;
      uint4 c0[8]; uint2 cm[8];
      bool other = false;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        c0[j] = make_uint4(0, 0, 0, 0); cm[j] = make_uint2(0, 0);
        const uint32_t cs = N.key[j] != kCellEmpty ? cell_find(G, fig, N.key[j]) : kCellEmpty;
        if (cs != kCellEmpty) {
          c0[j] = *reinterpret_cast<const uint4*>(tab + cs);                       // pmax1, pminc, count, sx
          if (c0[j].z != 0) {
            cm[j] = *reinterpret_cast<const uint2*>(&(tab + cs)->mean);
            other |= (c0[j].x - 1u) != ~c0[j].y;
          }
        }
      }
      if (other) {
        unsigned long long C[3] = {0, 0, 0}, cntw = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t cnt = c0[j].z;
          uint32_t m[3] = {256u * p[0], 256u * p[1], 256u * p[2]};                  // empty cell -> the point itself
          if (cnt > 0) {
            m[0] = 256u * ((N.key[j] & 1023u) * G.g) + (cm[j].x & 0xFFFFu);
            m[1] = 256u * (((N.key[j] >> 10) & 1023u) * G.g) + (cm[j].x >> 16);
            m[2] = 256u * ((N.key[j] >> 20) * G.g) + (cm[j].y & 0xFFFFu);
          }
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) C[ax] += (unsigned long long)N.wgt[j] * m[ax];
          cntw += (unsigned long long)N.wgt[j] * cnt;
        }
        const unsigned long long count = div_w3(cntw, G, N.w3);
        if (count > 0) {
          unsigned long long c4[3], D2 = 0;
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) {
            c4[ax] = div_w3(C[ax] + N.w3 / 2, G, N.w3);
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
              unsigned long long rr = (c4[ax] + 128) >> 8;
              if (rr > 65535) rr = 65535;
              q[ax] = (uint16_t)rr;
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
    }
    // ---- colour (K7), on the same pre-smoothing position ----
    if (a.sm.col.on && a.has_attr && neighbourhood(a.sm.col, p, N)) {
      const GridDesc& G = a.sm.col;
      const ColCell* tab = reinterpret_cast<const ColCell*>(G.table) + (uint64_t)fig * G.slots;
      uint4 c0[8]; uint4 c1[8];
      bool other = false;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        c0[j] = make_uint4(0, 0, 0, 0); c1[j] = make_uint4(0, 0, 0, 0);
        const uint32_t cs = N.key[j] != kCellEmpty ? cell_find(G, fig, N.key[j]) : kCellEmpty;
        if (cs != kCellEmpty) {
          c0[j] = *reinterpret_cast<const uint4*>(tab + cs);                       // pmax1 | final, pminc, count, meanY
          if (c0[j].z != 0) {
            c1[j] = *(reinterpret_cast<const uint4*>(tab + cs) + 1);               // meanU, meanV, variance verdict
            other |= ((c0[j].x & ~kCellFinal) - 1u) != ~c0[j].y;
          }
        }
      }
      if (other) {
        unsigned long long C[3] = {0, 0, 0};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          bool usable = c0[j].z != 0 && c1[j].z != 0;
          const uint32_t mean[3] = {c0[j].w, c1[j].x, c1[j].y};
          if (usable) {
            const long long dy = (long long)mean[0] - (long long)(256u * col[0]);
            if ((unsigned long long)(dy < 0 ? -dy : dy) > 256ull * a.sm.thr_col_diff) usable = false;
          }
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) C[ax] += (unsigned long long)N.wgt[j] * (usable ? mean[ax] : 256u * col[ax]);
        }
        uint32_t q[3]; unsigned long long dist = 0;
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
          const unsigned long long c4 = div_w3(C[ax] + N.w3 / 2, G, N.w3);
          unsigned long long rr = (c4 + 128) >> 8;
          if (rr > 65535) rr = 65535;
          q[ax] = (uint32_t)rr;
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
  moved = __reduce_add_sync(kFull, moved);
  recol = __reduce_add_sync(kFull, recol);
  if (lane_id() == 0) {
    if (moved) atomicAdd(&a.sm.changed[f], (unsigned long long)moved);
    if (recol) atomicAdd(&a.sm.changed[a.n_frames + f], (unsigned long long)recol);
  }
}

// back to all-zero cells (and free keys) for the next group: walk the same per-slot logs
__global__ void __launch_bounds__(256) smooth_clear_kernel(const __grid_constant__ UnpackArgs a) {
  const uint32_t ls = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ls >= a.sm.group_slots) return;
  const uint32_t lane = lane_id();
  const uint32_t frame = a.tile_frame[(a.sm.group_first_slot + ls) / kWarpsPerTile];
  const uint32_t fig = frame - a.sm.group_first_frame;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const GridDesc& G = which ? a.sm.col : a.sm.geo;
    if (!G.on) continue;
    const uint32_t n = min(G.log_count[ls], a.sm.log_stride);
    const uint32_t* log = G.log + (uint64_t)ls * a.sm.log_stride;
    uint4* tab = reinterpret_cast<uint4*>(G.table) + ((uint64_t)fig * G.slots) * 2;      // 32-byte cells
    for (uint32_t i = lane; i < n; i += 32) {
      const uint32_t cs = log[i];
      tab[(uint64_t)cs * 2] = make_uint4(0, 0, 0, 0);
      tab[(uint64_t)cs * 2 + 1] = make_uint4(0, 0, 0, 0);
      if (!G.identity) G.keys[(uint64_t)fig * G.slots + cs] = kCellEmpty;
    }
  }
}

__global__ void __launch_bounds__(256) fill_u32_kernel(uint32_t* p, uint64_t n, uint32_t v) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}

// ----------------------------------------------------------------------------------------------------------------
// launch wrappers
// ----------------------------------------------------------------------------------------------------------------
static inline int after_launch() { ++g_launches; return (int)cudaGetLastError(); }

int launch_block_to_patch(const UnpackArgs& a, uint32_t n_slots, void* stream) {
  if (n_slots == 0) return 0;
  const uint32_t blocks = (n_slots + 255) / 256;
  block_to_patch_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, n_slots, const_cast<uint32_t*>(a.block_to_patch));
  return after_launch();
}

int launch_compact_owned(const UnpackArgs& a, void* stream) {
  if (a.n_frames == 0) return 0;
  compact_owned_kernel<<<a.n_frames, 1024, 0, (cudaStream_t)stream>>>(a);
  return after_launch();
}

template <bool kSmooth, bool kDebug>
static int launch_emit_t(const UnpackArgs& a, uint32_t tile_begin, uint32_t tile_end, cudaStream_t s) {
  const size_t smem = (size_t)kWarpSmemBytes * kWarpsPerTile;
  cudaError_t e = cudaFuncSetAttribute((const void*)emit_kernel<kSmooth, kDebug>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  emit_kernel<kSmooth, kDebug><<<tile_end - tile_begin, kWarpsPerTile * 32, smem, s>>>(a, tile_begin);
  return after_launch();
}

int launch_count(const UnpackArgs& a, uint32_t tile_begin, uint32_t tile_end, void* stream) {
  if (tile_end <= tile_begin) return 0;
  count_kernel<<<tile_end - tile_begin, kWarpsPerTile * 32, 0, (cudaStream_t)stream>>>(a, tile_begin);
  return after_launch();
}

int launch_slot_scan(const UnpackArgs& a, void* stream) {
  if (a.n_frames == 0) return 0;
  slot_scan_kernel<<<a.n_frames, 1024, 0, (cudaStream_t)stream>>>(a);
  return after_launch();
}

int launch_emit(const UnpackArgs& a, bool smooth, uint32_t tile_begin, uint32_t tile_end, void* stream) {
  if (tile_end <= tile_begin) return 0;
  const cudaStream_t s = (cudaStream_t)stream;
  const bool debug = a.out.yuv || a.out.part || a.out.pix || a.out.btype;
  if (smooth) return debug ? launch_emit_t<true, true>(a, tile_begin, tile_end, s) : launch_emit_t<true, false>(a, tile_begin, tile_end, s);
  return debug ? launch_emit_t<false, true>(a, tile_begin, tile_end, s) : launch_emit_t<false, false>(a, tile_begin, tile_end, s);
}

int launch_upsample(const UnpackArgs& a, uint8_t* occ_full, void* stream) {
  upsample_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(a, occ_full);
  return after_launch();
}

int launch_smooth_finalize(const UnpackArgs& a, void* stream) {
  if (a.sm.group_slots == 0) return 0;
  smooth_finalize_kernel<<<(a.sm.group_slots * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
  return after_launch();
}

int launch_smooth_filter(const UnpackArgs& a, void* stream) {
  if (a.sm.group_frames == 0) return 0;
  const unsigned bx = (148u * 8u + a.sm.group_frames - 1) / a.sm.group_frames;
  smooth_filter_kernel<<<dim3(bx, a.sm.group_frames), 256, 0, (cudaStream_t)stream>>>(a);
  return after_launch();
}

int launch_smooth_clear(const UnpackArgs& a, void* stream) {
  if (a.sm.group_slots == 0) return 0;
  smooth_clear_kernel<<<(a.sm.group_slots * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
  return after_launch();
}

int launch_yuv_to_rgb_flat(const uint16_t* yuv, uint8_t* rgb, uint64_t n, void* stream) {
  if (n == 0) return 0;
  uint64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  yuv_to_rgb_flat_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(yuv, rgb, n);
  return after_launch();
}
int launch_fill_u32(uint32_t* p, uint64_t n, uint32_t v, void* stream) {
  if (n == 0) return 0;
  fill_u32_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(p, n, v);
  return after_launch();
}

}  // namespace tmc2
