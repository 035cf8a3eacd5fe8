// stream_decode.cpp -- a plain C++ consumer of the C ABI (include/tmc2gpu.h), shaped like the reference's runtime:
// a worker thread reconstructs GOFs and hands frames to the consumer through a bounded(1) channel, in order
// (tmc2-rs src/lib.rs:81,113-137: `bounded(1)` crossbeam channel between the decoder thread and the iterator).
//
//   stream_decode <input.gof> <output.bin> [repeat]
//
// input.gof (written by tests/test_cpp_driver.py): u32 magic 'TMC2', u32 W, H, occW, occH, F, then sizeof(tmc2_params)
// raw bytes, then per frame: u32 patch_count + patches (tmc2_patch raw), occ plane, geo[2], attrY[2], attrU[2], attrV[2]
// (tight planes).  output.bin: per frame u64 point_count, positions (u16 x 3n), colors (u8 x 3n).
// Nothing here touches CUDA headers or the oracle: it only links libtmc2gpu.so.
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <optional>
#include <thread>
#include <vector>

#include "../include/tmc2gpu.h"

struct PointSet3 {                       // reference src/codec.rs:20-36
  std::vector<uint16_t> positions;       // [n][3]
  std::vector<uint8_t> colors;           // [n][3]
  bool with_colors = false;
};

template <typename T> class Bounded1 {   // crossbeam bounded(1)
  std::mutex mu; std::condition_variable cv; std::optional<T> slot; bool closed = false;
 public:
  void send(T v) { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return !slot.has_value(); }); slot = std::move(v); cv.notify_all(); }
  void close() { std::lock_guard<std::mutex> lk(mu); closed = true; cv.notify_all(); }
  bool recv(T& out) {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return slot.has_value() || closed; });
    if (!slot.has_value()) return false;
    out = std::move(*slot); slot.reset(); cv.notify_all();
    return true;
  }
};

static bool read_exact(FILE* f, void* p, size_t n) { return n == 0 || fread(p, 1, n, f) == n; }

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s <input.gof> <output.bin> [repeat]\n", argv[0]); return 2; }
  const int repeat = argc > 3 ? atoi(argv[3]) : 1;
  FILE* in = fopen(argv[1], "rb");
  if (!in) { perror("input"); return 2; }
  uint32_t hdr[6];
  if (!read_exact(in, hdr, sizeof hdr) || hdr[0] != 0x32434D54u) { fprintf(stderr, "bad header\n"); return 2; }
  const uint32_t W = hdr[1], H = hdr[2], occW = hdr[3], occH = hdr[4], F = hdr[5];
  tmc2_params params;
  if (!read_exact(in, &params, sizeof params)) return 2;

  // planes live in pinned memory from the library (zero staging)
  const size_t n_occ = (size_t)occW * occH, n_y = (size_t)W * H, n_c = (size_t)(W / 2) * (H / 2);
  const size_t per_frame = n_occ + 2 * (2 * n_y) * 2 + 4 * n_c * 2;      // bytes: occ + geo[2] + attrY[2] (u16) + U,V [2] (u16)
  uint8_t* pinned = static_cast<uint8_t*>(tmc2gpu_alloc_pinned(per_frame * F + 64));
  if (!pinned) { fprintf(stderr, "tmc2gpu_alloc_pinned failed (no CUDA device?)\n"); return 3; }
  std::vector<std::vector<tmc2_patch>> patches(F);
  std::vector<tmc2_frame> frames(F);
  uint8_t* p = pinned;
  for (uint32_t f = 0; f < F; ++f) {
    uint32_t pc;
    if (!read_exact(in, &pc, 4)) return 2;
    patches[f].resize(pc);
    if (!read_exact(in, patches[f].data(), pc * sizeof(tmc2_patch))) return 2;
    tmc2_frame fr{};
    auto take = [&](size_t bytes) { uint8_t* q = p; if (!read_exact(in, q, bytes)) exit(2); p += bytes; return q; };
    fr.occ = take(n_occ);
    if (reinterpret_cast<uintptr_t>(p) & 1) ++p;                          // u16 planes 2-byte aligned
    for (int m = 0; m < 2; ++m) fr.geo[m] = reinterpret_cast<const uint16_t*>(take(n_y * 2));
    for (int m = 0; m < 2; ++m) fr.attr_y[m] = reinterpret_cast<const uint16_t*>(take(n_y * 2));
    for (int m = 0; m < 2; ++m) fr.attr_u[m] = reinterpret_cast<const uint16_t*>(take(n_c * 2));
    for (int m = 0; m < 2; ++m) fr.attr_v[m] = reinterpret_cast<const uint16_t*>(take(n_c * 2));
    fr.patches = patches[f].data(); fr.patch_count = pc;
    fr.occ_stride = occW; fr.geo_stride = W; fr.attr_stride_y = W; fr.attr_stride_c = W / 2;
    frames[f] = fr;
  }
  fclose(in);
  tmc2_gof gof{};
  gof.width = W; gof.height = H; gof.occ_width = occW; gof.occ_height = occH; gof.frame_count = F;
  gof.geo_video_frames = 2 * F; gof.attr_video_frames = 2 * F; gof.frames = frames.data(); gof.params = params;

  tmc2gpu_ctx* ctx = nullptr;
  tmc2_limits lim{}; lim.gofs_in_flight = 2;
  tmc2_status st = tmc2gpu_create(nullptr, 0, &lim, &ctx);
  if (st != TMC2_OK) { fprintf(stderr, "tmc2gpu_create: %s\n", tmc2gpu_status_string(st)); return 3; }

  Bounded1<PointSet3> chan;
  int worker_status = 0;
  std::thread worker([&] {                                                // the reference's decoder thread, src/lib.rs:113
    for (int r = 0; r < repeat && !worker_status; ++r) {
      st = tmc2gpu_submit_gof(ctx, &gof);                                 // replaces src/decoder.rs:188-305
      if (st != TMC2_OK) { fprintf(stderr, "submit_gof: %s: %s\n", tmc2gpu_status_string(st), tmc2gpu_last_error(ctx)); worker_status = 4; break; }
      for (uint32_t f = 0; f < F; ++f) {
        tmc2_frame_out out;
        st = tmc2gpu_next_frame(ctx, &out);
        if (st != TMC2_OK) { fprintf(stderr, "next_frame: %s: %s\n", tmc2gpu_status_string(st), tmc2gpu_last_error(ctx)); worker_status = 4; break; }
        PointSet3 ps;
        ps.positions.assign(out.positions, out.positions + 3 * out.point_count);
        ps.with_colors = out.with_colors != 0;
        if (ps.with_colors) ps.colors.assign(out.colors, out.colors + 3 * out.point_count);
        tmc2gpu_release_frame(ctx, &out);
        chan.send(std::move(ps));                                         // tx.send(reconstruct), src/decoder.rs:311
      }
    }
    chan.close();
  });

  FILE* outf = fopen(argv[2], "wb");
  PointSet3 ps;
  uint64_t frames_out = 0, points = 0;
  while (chan.recv(ps)) {                                                 // the library user's `for frame in decoder`
    const uint64_t n = ps.positions.size() / 3;
    if (frames_out < F) {                                                 // keep the first pass only
      fwrite(&n, 8, 1, outf);
      fwrite(ps.positions.data(), 2, ps.positions.size(), outf);
      fwrite(ps.colors.data(), 1, ps.colors.size(), outf);
    }
    ++frames_out; points += n;
  }
  worker.join();
  fclose(outf);
  tmc2gpu_destroy(ctx);
  tmc2gpu_free_pinned(pinned);
  printf("frames %llu points %llu status %d\n", (unsigned long long)frames_out, (unsigned long long)points, worker_status);
  return worker_status;
}
