// ply.cu -- a reconstructed frame formatted as a PLY file ON THE DEVICE.
//
// The reference's consumer of a PointSet3 is PlyWriter (src/writer.rs:15-75): a header (write_header :31-60) and one line
// "x y z r g b\n" per point in decimal (write_body :62-75; "x y z\n" without colours).  A frame that already sits in HBM is
// turned into exactly those bytes here, so the host receives a finished file instead of formatting 0.8 M lines itself:
//   measure  : decimal length of every point's line, summed per block of 1024 points
//   scan     : exclusive prefix over the block sums (one CTA) -> byte offset of every block, total body length
//   write    : every block formats its 1024 lines into shared memory (<= 30 bytes per line) and copies the segment out with
//              16-byte stores
// The binary_little_endian variant the reference lists but leaves commented out (:10-11, :41-46) has fixed records (uint x y z,
// uchar red green blue: 15 bytes, 12 without colours) and needs the write pass only.
#include <cstdint>
#include <cuda_runtime.h>

#include "device_types.h"

namespace tmc2 {
namespace {

constexpr uint32_t kPlyThreads = 256, kPlyPerThread = 4, kPlyPerBlock = kPlyThreads * kPlyPerThread;
constexpr uint32_t kPlyMaxLine = 30;   // "65535 65535 65535 255 255 255\n"

__device__ __forceinline__ uint32_t dec_digits(uint32_t v) {   // v < 100 000
  return 1u + (v >= 10u) + (v >= 100u) + (v >= 1000u) + (v >= 10000u);
}

struct PlyPoint { uint32_t x, y, z, r, g, b; };

__device__ __forceinline__ PlyPoint load_point(const uint16_t* __restrict__ pos, const uint8_t* __restrict__ rgb, uint64_t i) {
  PlyPoint p;
  p.x = pos[i * 3]; p.y = pos[i * 3 + 1]; p.z = pos[i * 3 + 2];
  p.r = p.g = p.b = 0;
  if (rgb) { p.r = rgb[i * 3]; p.g = rgb[i * 3 + 1]; p.b = rgb[i * 3 + 2]; }
  return p;
}
__device__ __forceinline__ uint32_t line_length(const PlyPoint& p, bool colours) {
  uint32_t n = dec_digits(p.x) + dec_digits(p.y) + dec_digits(p.z) + 3u;              // two blanks + '\n'
  if (colours) n += dec_digits(p.r) + dec_digits(p.g) + dec_digits(p.b) + 3u;         // three more blanks
  return n;
}
__device__ __forceinline__ uint32_t put_dec(uint8_t* s, uint32_t v, uint8_t after) {  // returns bytes written
  const uint32_t d = dec_digits(v);
  for (uint32_t k = d; k-- > 0;) { s[k] = (uint8_t)('0' + v % 10u); v /= 10u; }
  s[d] = after;
  return d + 1u;
}

// sum over the block; valid in every thread
__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* s_w) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  if ((threadIdx.x & 31u) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  uint32_t t = 0;
  for (uint32_t w = 0; w < kPlyThreads / 32; ++w) t += s_w[w];
  __syncthreads();
  return t;
}
// exclusive prefix over the block (thread order)
__device__ __forceinline__ uint32_t block_exclusive(uint32_t v, uint32_t* s_w, uint32_t& total) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t inc = v;
  for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= (uint32_t)o) inc += u; }
  if (lane == 31u) s_w[warp] = inc;
  __syncthreads();
  uint32_t before = 0; total = 0;
  for (uint32_t w = 0; w < kPlyThreads / 32; ++w) { const uint32_t t = s_w[w]; if (w < warp) before += t; total += t; }
  __syncthreads();
  return before + inc - v;
}

__global__ void __launch_bounds__(kPlyThreads) ply_measure_kernel(const uint16_t* __restrict__ pos, const uint8_t* __restrict__ rgb,
                                                                  uint64_t n, uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t s_w[kPlyThreads / 32];
  const uint64_t i0 = (uint64_t)blockIdx.x * kPlyPerBlock + (uint64_t)threadIdx.x * kPlyPerThread;
  uint32_t len = 0;
#pragma unroll
  for (uint32_t k = 0; k < kPlyPerThread; ++k)
    if (i0 + k < n) len += line_length(load_point(pos, rgb, i0 + k), rgb != nullptr);
  const uint32_t t = block_sum(len, s_w);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = t;
}

__global__ void __launch_bounds__(1024) ply_scan_kernel(const uint32_t* __restrict__ block_sums, uint32_t n_blocks,
                                                        unsigned long long* __restrict__ block_offs, unsigned long long* __restrict__ total) {
  __shared__ unsigned long long s_w[32];
  __shared__ unsigned long long s_carry;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < n_blocks; b0 += 1024) {
    const uint32_t i = b0 + threadIdx.x;
    const unsigned long long v = i < n_blocks ? block_sums[i] : 0ull;
    unsigned long long inc = v;
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, inc, o); if (lane >= (uint32_t)o) inc += u; }
    if (lane == 31u) s_w[warp] = inc;
    __syncthreads();
    unsigned long long before = s_carry, all = 0;
    for (uint32_t w = 0; w < 32; ++w) { const unsigned long long t = s_w[w]; if (w < warp) before += t; all += t; }
    if (i < n_blocks) block_offs[i] = before + inc - v;
    __syncthreads();
    if (threadIdx.x == 0) s_carry += all;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

// kAscii: block_offs from the scan; otherwise fixed records.  `out` is the first byte of the BODY.
template <bool kAscii>
__global__ void __launch_bounds__(kPlyThreads) ply_write_kernel(const uint16_t* __restrict__ pos, const uint8_t* __restrict__ rgb, uint64_t n,
                                                                const unsigned long long* __restrict__ block_offs, uint8_t* __restrict__ out) {
  __shared__ __align__(16) uint8_t s_stage[kPlyPerBlock * kPlyMaxLine + 16];
  __shared__ uint32_t s_w[kPlyThreads / 32];
  const bool colours = rgb != nullptr;
  const uint32_t rec = colours ? 15u : 12u;
  const uint64_t p0 = (uint64_t)blockIdx.x * kPlyPerBlock, i0 = p0 + (uint64_t)threadIdx.x * kPlyPerThread;
  PlyPoint pt[kPlyPerThread];
  uint32_t len = 0;
#pragma unroll
  for (uint32_t k = 0; k < kPlyPerThread; ++k) {
    if (i0 + k < n) {
      pt[k] = load_point(pos, rgb, i0 + k);
      len += kAscii ? line_length(pt[k], colours) : rec;
    }
  }
  uint32_t seg_len;
  const uint32_t off = block_exclusive(len, s_w, seg_len);
  uint8_t* const g0 = out + (kAscii ? block_offs[blockIdx.x] : p0 * rec);   // first byte of the block's segment
  const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(g0) & 15u);  // stage byte i <-> global byte (g0 - shift + i)
  uint8_t* s = s_stage + shift + off;
#pragma unroll
  for (uint32_t k = 0; k < kPlyPerThread; ++k) {
    if (i0 + k >= n) break;
    const PlyPoint& p = pt[k];
    if (kAscii) {
      s += put_dec(s, p.x, ' ');
      s += put_dec(s, p.y, ' ');
      s += put_dec(s, p.z, colours ? ' ' : '\n');
      if (colours) {
        s += put_dec(s, p.r, ' ');
        s += put_dec(s, p.g, ' ');
        s += put_dec(s, p.b, '\n');
      }
    } else {
      const uint32_t w[3] = {p.x, p.y, p.z};                                 // property uint x / y / z, little endian
#pragma unroll
      for (int c = 0; c < 3; ++c) { s[4 * c] = (uint8_t)w[c]; s[4 * c + 1] = (uint8_t)(w[c] >> 8); s[4 * c + 2] = 0; s[4 * c + 3] = 0; }
      if (colours) { s[12] = (uint8_t)p.r; s[13] = (uint8_t)p.g; s[14] = (uint8_t)p.b; }
      s += rec;
    }
  }
  __syncthreads();
  // copy out: 16-byte chunks of the aligned window that lie wholly inside the segment as one store, the ragged ends bytewise
  uint8_t* const ga = g0 - shift;
  const uint32_t end = shift + seg_len, chunks = (end + 15u) >> 4;
  for (uint32_t c = threadIdx.x; c < chunks; c += kPlyThreads) {
    const uint32_t b0 = c << 4;
    if (b0 >= shift && b0 + 16u <= end) {
      *reinterpret_cast<uint4*>(ga + b0) = *reinterpret_cast<const uint4*>(s_stage + b0);
    } else {
      const uint32_t lo = b0 < shift ? shift : b0, hi = b0 + 16u < end ? b0 + 16u : end;
      for (uint32_t i = lo; i < hi; ++i) ga[i] = s_stage[i];
    }
  }
}

}  // namespace

int launch_ply_measure(const uint16_t* pos, const uint8_t* rgb, uint64_t n, uint32_t* block_sums, void* stream) {
  if (n == 0) return 0;
  ply_measure_kernel<<<(uint32_t)ply_blocks(n), kPlyThreads, 0, (cudaStream_t)stream>>>(pos, rgb, n, block_sums);
  return (int)cudaGetLastError();
}
int launch_ply_scan(const uint32_t* block_sums, uint64_t n, unsigned long long* block_offs, unsigned long long* total, void* stream) {
  ply_scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(block_sums, (uint32_t)ply_blocks(n), block_offs, total);
  return (int)cudaGetLastError();
}
int launch_ply_write(const uint16_t* pos, const uint8_t* rgb, uint64_t n, const unsigned long long* block_offs, uint8_t* body,
                     bool ascii, void* stream) {
  if (n == 0) return 0;
  if (ascii) ply_write_kernel<true><<<(uint32_t)ply_blocks(n), kPlyThreads, 0, (cudaStream_t)stream>>>(pos, rgb, n, block_offs, body);
  else ply_write_kernel<false><<<(uint32_t)ply_blocks(n), kPlyThreads, 0, (cudaStream_t)stream>>>(pos, rgb, n, block_offs, body);
  return (int)cudaGetLastError();
}

}  // namespace tmc2
