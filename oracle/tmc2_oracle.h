/*
 * tmc2_oracle.h -- CPU oracle for the V-PCC rec0 reconstruction hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded restatement of the algorithm of tmc2-rs (benclmnt/tmc2-rs):
 * src/codec.rs:205-250, :256-514, :517-565, :569-658, :661-687, :88-94 and the Patch / Image helpers
 * src/decoder.rs:827-888, :973-1020, driven like the frame loop src/decoder.rs:188-314.
 * Same loop nest, same push order, same per-point growing vectors as the reference.
 *
 * PARITY PINNING: the reference cannot be built in this image (no cargo/rustc, no ffmpeg) and holds
 * no golden vectors, known-answer tests or fixtures for this path (its only tests are five bit-reader
 * tests, src/bitstream.rs:345-438).  This oracle is therefore pinned by (1) the source text it cites,
 * (2) the hand-derived known-answer vector of SURVEY.md Appendix C and the colour KATs derived from
 * src/codec.rs:661-687 (tests/test_oracle_kat.py), (3) an independent numpy model of the same lines
 * (tests/refmodel.py).  For the post-processing stages (boundary detection, grid geometry smoothing,
 * grid colour smoothing) the reference has only `unimplemented!()` stubs: "parity unpinned" -- the
 * arithmetic below is this repository's own frozen integer specification.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
 * this library.  The product path (tmc2-rs_b200/csrc) never links, loads or calls it.
 */
#ifndef TMC2_ORACLE_H
#define TMC2_ORACLE_H

#include <stddef.h>
#include <stdint.h>
#include "../include/tmc2gpu.h" /* input PODs (tmc2_patch, tmc2_params, tmc2_frame, tmc2_gof) + status codes */

#ifdef __cplusplus
extern "C" {
#endif

/* One reconstructed frame: reference PointSet3 (codec.rs:20-36) plus everything generate_point_cloud
 * leaves behind in the TileContext (context.rs:395-439) and returns (partition, point_to_pixel).
 * `usize` is uint64_t. */
typedef struct orc_frame {
  uint64_t  point_count;          /* tile.total_number_of_regular_points (codec.rs:482)            */
  uint16_t* positions;            /* [n][3] PointSet3.positions  (after smoothing if enabled)      */
  uint8_t*  colors;               /* [n][3] PointSet3.colors                                       */
  uint16_t* colors16bit;          /* [n][3] PointSet3.colors16bit                                  */
  uint64_t* point_patch_indexes;  /* [n][2] (tile_index, patch_index)                              */
  uint64_t* partition;            /* [n]                                                           */
  uint64_t* point_to_pixel;       /* [n][3] (x, y, map)                                            */
  uint8_t*  occupancy_map;        /* [H*W]  tile.occupancy_map                                     */
  uint64_t* block_to_patch;       /* [blocks] tile.block_to_patch                                  */
  uint64_t  block_count;
  uint8_t   with_colors;
  /* post-processing (own spec) */
  uint8_t*  boundary_type;            /* [n]                                                       */
  uint16_t* positions_presmooth;      /* [n][3]                                                    */
  uint16_t* colors16bit_presmooth;    /* [n][3]                                                    */
  uint64_t  smoothed_positions;
  uint64_t  smoothed_colors;
} orc_frame;

/* src/codec.rs:205-250.  block_to_patch has (width/res)*(height/res) entries.  Returns tmc2_status. */
int orc_generate_block_to_patch_from_occupancy_map_video(const tmc2_gof* gof, uint32_t frame_index,
                                                         uint64_t* block_to_patch);

/* The per-frame body of src/decoder.rs:188-311: block_to_patch, generate_point_cloud, append_point_set,
 * [geometry smoothing], [colour smoothing], convert_yuv16_to_rgb8.  Returns tmc2_status; *out is owned
 * by the caller (orc_frame_free). */
int  orc_reconstruct_frame(const tmc2_gof* gof, uint32_t frame_index, orc_frame** out);
void orc_frame_free(orc_frame* f);

/* src/codec.rs:661-687 */
void orc_convert_yuv10_to_rgb8(const uint16_t yuv[3], uint8_t rgb[3]);

/* src/decoder.rs:853-867 (+ the asserts of :835 / :848).  resolution==1 and is_block!=0 give the block
 * variant.  Returns TMC2_OK or TMC2_ERR_PATCH_OUT_OF_CANVAS. */
int orc_patch_to_canvas(const tmc2_patch* p, uint32_t occupancy_resolution, int orientation_mode, int is_block,
                        uint64_t u, uint64_t v, uint64_t canvas_stride, uint64_t canvas_height,
                        uint64_t* x, uint64_t* y);

/* src/decoder.rs:871-888 */
void orc_patch_generate_point(const tmc2_patch* p, uint64_t u, uint64_t v, uint16_t depth, uint16_t out[3]);

/* Timing helper for bench.py's cpu_baseline: reconstruct frames [first, first+count) of the GOF one after
 * another on the calling thread, discard the results, return total points (or -status on error). */
int64_t orc_time_frames(const tmc2_gof* gof, uint32_t first, uint32_t count);

#ifdef __cplusplus
}
#endif
#endif
