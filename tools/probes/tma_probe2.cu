// tma_probe2.cu -- which box shapes does cp.async.bulk.tensor accept on sm_100a?  (inner box extents of 16 vs 32 bytes, u8 vs u16)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
struct alignas(64) Map1 { unsigned long long m[16]; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int kRank>
__global__ void probe(const __grid_constant__ Map1 tm, uint32_t* out, uint32_t bytes, int c0, int c1, int c2, int c3) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4096);
  const uint32_t lane = threadIdx.x;
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (lane == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
    if (kRank == 4)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   :: "r"(smem_u32(smem)), "l"(tm.m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   :: "r"(smem_u32(smem)), "l"(tm.m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
               :: "r"(smem_u32(bar)), "r"(0) : "memory");
  uint32_t s = 0;
  for (uint32_t i = lane; i < bytes; i += 32) s += smem[i] * (i + 7);
  out[lane] = s;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#include <cstdlib>
int main(int argc, char** argv) {
  const uint32_t F = 3;
  std::vector<uint8_t> buf((size_t)F * 2 * 128 * 256);
  for (size_t i = 0; i < buf.size(); ++i) buf[i] = (uint8_t)(i * 40503u >> 7);
  uint8_t* d; uint32_t* dout;
  cudaMalloc(&d, buf.size()); cudaMalloc(&dout, 128);
  cudaMemcpy(d, buf.data(), buf.size(), cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  auto test = [&](const char* name, int esz, int rank, uint32_t bx, uint32_t by, uint32_t bz, int c0, int c1) {
    Map1 tm; memset(&tm, 0, sizeof tm);
    CUtensorMap m;
    CUresult r;
    const uint64_t pitch = 256;          // bytes per row
    if (rank == 4) {
      cuuint64_t dm[4] = {pitch / esz, 128, 2, F}, st[3] = {pitch, 128 * pitch, 2 * 128 * pitch};
      cuuint32_t b[4] = {bx, by, bz, 1}, es[4] = {1, 1, 1, 1};
      r = enc(&m, esz == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, d, dm, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      cuuint64_t dm[3] = {pitch / esz, 256, F}, st[2] = {pitch, 256 * pitch};
      cuuint32_t b[3] = {bx, by, 1}, es[3] = {1, 1, 1};
      r = enc(&m, esz == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, d, dm, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    memcpy(tm.m, &m, 128);
    const uint32_t bytes = bx * by * (rank == 4 ? bz : 1) * esz;
    cudaMemset(dout, 0, 128);
    if (rank == 4) probe<4><<<1, 32, 8192>>>(tm, dout, bytes, c0, c1, 0, 1);
    else probe<3><<<1, 32, 8192>>>(tm, dout, bytes, c0, c1, 1, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%-40s encode %d  run: %s\n", name, (int)r, cudaGetErrorString(e));
    return e == cudaSuccess;
  };
  bool ok = true;
  // each case in its own process invocation (argv[1] selects): an error is sticky
  const int which = argc > 1 ? atoi(argv[1]) : 0;
  if (which == 0) ok = test("u16 4d box 16x16x2 (32 B rows)", 2, 4, 16, 16, 2, 32, 16);
  if (which == 1) ok = test("u16 3d box 16x16", 2, 3, 16, 16, 1, 32, 16);
  if (which == 2) ok = test("u8 4d box 32x8x2", 1, 4, 32, 8, 2, 4, 6);
  if (which == 3) ok = test("u8 4d box 32x8x1", 1, 4, 32, 8, 1, 4, 6);
  if (which == 4) ok = test("u16 4d box 8x8x2 (16 B rows)", 2, 4, 8, 8, 2, 8, 8);
  if (which == 5) ok = test("u8 3d box 64x8", 1, 3, 64, 8, 1, 0, 0);
  if (which == 6) ok = test("u16 3d box 8x8 (16 B rows)", 2, 3, 8, 8, 1, 8, 8);
  if (which == 7) ok = test("u8 4d box 16x8x1 (16 B rows)", 1, 4, 16, 8, 1, -4, -2);
  printf(ok ? "all ok\n" : "stopped at the first failure (sticky error)\n");
  return 0;
}
