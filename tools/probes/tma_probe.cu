// tma_probe.cu -- stand-alone check of the TMA / mbarrier / cp.async building blocks emit_kernel uses (B200, sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

struct alignas(64) Maps { unsigned long long geo[16], occ[16]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int kWhat>
__global__ void probe(const __grid_constant__ Maps tm, const uint16_t* geo, uint32_t* out, int x0, int y0, int f, int ox, int oy) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4096);
  const uint32_t lane = threadIdx.x;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (kWhat == 0) {          // cp.async only
    if (lane < 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smem_u32(smem + 16 * lane)), "l"(geo + 8 * lane) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    out[lane] = reinterpret_cast<uint16_t*>(smem)[lane & 15];
    return;
  }
  if (lane == 0) {
    const uint32_t bytes = kWhat == 1 ? 1024u : kWhat == 2 ? 128u : 1152u;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
    if (kWhat == 1 || kWhat == 3)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   :: "r"(smem_u32(smem)), "l"(tm.geo), "r"(smem_u32(bar)), "r"(x0), "r"(y0), "r"(0), "r"(f) : "memory");
    if (kWhat == 2 || kWhat == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   :: "r"(smem_u32(smem + 2048)), "l"(tm.occ), "r"(smem_u32(bar)), "r"(ox), "r"(oy), "r"(f) : "memory");
  }
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
      :: "r"(smem_u32(bar)), "r"(0) : "memory");
  // checksum of what landed
  uint32_t s = 0;
  if (kWhat == 1 || kWhat == 3) for (int i = lane; i < 512; i += 32) s += reinterpret_cast<uint16_t*>(smem)[i] * (i + 1);
  if (kWhat == 2 || kWhat == 3) for (int i = lane; i < 128; i += 32) s += smem[2048 + i] * (i + 7);
  out[lane] = s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const uint32_t W = 256, H = 256, F = 3, pitch = 256, ow = 64, oh = 64, opitch = 64;
  std::vector<uint16_t> geo((size_t)F * 2 * H * pitch);
  std::vector<uint8_t> occ((size_t)F * oh * opitch);
  for (size_t i = 0; i < geo.size(); ++i) geo[i] = (uint16_t)(i * 2654435761u >> 13);
  for (size_t i = 0; i < occ.size(); ++i) occ[i] = (uint8_t)(i * 40503u >> 7);
  uint16_t* dgeo; uint8_t* docc; uint32_t* dout;
  cudaMalloc(&dgeo, geo.size() * 2); cudaMalloc(&docc, occ.size()); cudaMalloc(&dout, 128);
  cudaMemcpy(dgeo, geo.data(), geo.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(docc, occ.data(), occ.size(), cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("no entry point\n"); return 2; }
  EncodeTiledFn enc = (EncodeTiledFn)p;
  Maps tm; memset(&tm, 0, sizeof tm);
  {
    cuuint64_t d[4] = {W, H, 2, F}, st[3] = {pitch * 2, (cuuint64_t)H * pitch * 2, (cuuint64_t)2 * H * pitch * 2};
    cuuint32_t b[4] = {16, 16, 2, 1}, es[4] = {1, 1, 1, 1};
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, dgeo, d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode geo: %d\n", (int)r); memcpy(tm.geo, &m, 128);
  }
  {
    cuuint64_t d[3] = {ow, oh, F}, st[2] = {opitch, (cuuint64_t)oh * opitch};
    cuuint32_t b[3] = {16, 8, 1}, es[3] = {1, 1, 1};
    CUtensorMap m;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, docc, d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode occ: %d\n", (int)r); memcpy(tm.occ, &m, 128);
  }
  auto run = [&](int what, int x0, int y0, int f, int ox, int oy) {
    cudaMemset(dout, 0, 128);
    if (what == 0) probe<0><<<1, 32, 8192>>>(tm, dgeo, dout, x0, y0, f, ox, oy);
    if (what == 1) probe<1><<<1, 32, 8192>>>(tm, dgeo, dout, x0, y0, f, ox, oy);
    if (what == 2) probe<2><<<1, 32, 8192>>>(tm, dgeo, dout, x0, y0, f, ox, oy);
    if (what == 3) probe<3><<<1, 32, 8192>>>(tm, dgeo, dout, x0, y0, f, ox, oy);
    cudaError_t e = cudaDeviceSynchronize();
    uint32_t h[32] = {};
    if (e == cudaSuccess) cudaMemcpy(h, dout, 128, cudaMemcpyDeviceToHost);
    uint32_t s = 0; for (int i = 0; i < 32; ++i) s += h[i];
    // host checksum
    uint32_t want = 0;
    if (what == 0) { for (int l = 0; l < 32; ++l) want += geo[l & 15]; }
    if (what == 1 || what == 3) for (int i = 0; i < 512; ++i) { int m = i / 256, r = (i / 16) % 16, c = i % 16; want += geo[(((size_t)f * 2 + m) * H + y0 + r) * pitch + x0 + c] * (i + 1); }
    if (what == 2 || what == 3) for (int i = 0; i < 128; ++i) { int r = i / 16, c = i % 16, x = ox + c, y = oy + r; uint8_t v = (x < 0 || y < 0 || x >= (int)ow || y >= (int)oh) ? 0 : occ[((size_t)f * oh + y) * opitch + x]; want += v * (i + 7); }
    printf("probe %d (%d,%d,%d | %d,%d): %s  sum %u want %u %s\n", what, x0, y0, f, ox, oy, cudaGetErrorString(e), s, want, s == want ? "OK" : "MISMATCH");
    return e;
  };
  if (run(0, 0, 0, 0, 0, 0) != cudaSuccess) return 1;
  if (run(1, 32, 48, 1, 0, 0) != cudaSuccess) return 1;
  if (run(2, 0, 0, 2, 4, 6) != cudaSuccess) return 1;
  if (run(2, 0, 0, 0, -4, -2) != cudaSuccess) return 1;
  if (run(2, 0, 0, 2, 60, 58) != cudaSuccess) return 1;
  if (run(3, 240, 240, 2, 56, 58) != cudaSuccess) return 1;
  printf("all probes done\n");
  return 0;
}
