// eloc.cuh -- what the kernels of the one-pass local energy share (eloc_scan.cu, eloc_block.cu):
// the hit lists the scan kernels hand to the evaluation kernel, the device counters of one call and the
// test / experiment knobs (set through pynqs_set_tuning, never read from the environment).
#pragma once
#include "gindex.cuh"

namespace pynqs {

// A sample's hits are stored as up to `run_stride` runs in the global hit buffer; run_cnt[s] says how many.
struct HitRun {
  u32 off, cnt;  // cnt & kOverflow: a queue or the buffer overflowed -> the sample takes the full route
};
constexpr u32 kOverflow = 0x80000000u;
constexpr u32 kNoSelf = 0xffffffffu;

// hit word: position in the grouped copy | kHitA (alpha-grouped copy) | kHitOwn (found in the scan of one of
// the sample's own strings, folded route only -- the eval kernel checks the class of the key accordingly)
constexpr u32 kHitA = 0x80000000u, kHitOwn = 0x40000000u, kHitPos = 0x3fffffffu;

// device counters of one call (zeroed by the launcher)
struct ElocCounters {
  u32 hit_cursor;    // next free slot of the hit buffer
  u32 tile_count;    // block route: tiles made by the grouping pass
  u32 tile_next;     // ... and handed out so far
  u32 slot_front;    // block route: samples placed in tiles
  u32 n_single;      // samples left to the per-sample kernel (stored at the back of the slot array)
  u32 single_next;
  u32 n_heavy;       // evaluation: samples with long hit lists, handed from the tile kernel to the warp-per-sample kernel
  u32 heavy_next;    // ... and taken so far
  u32 pad[56];
};
static_assert(sizeof(ElocCounters) == 256, "counters are 256 bytes");

// knobs for tests and experiments (defaults are the production values)
struct ElocTuning {
  int scan_threads = 0;        // 0: by the number of groups; else 64 / 128 / 256
  int search_factor = 64;      // a bucket this many times larger than what it could hold is searched, not walked
  int full_keys = 0;           // 1: one-word ONVs take the full-key route of multi-word ONVs
  int block_min_samples = 4096;  // calls with fewer samples go to the per-sample kernel only
  int block_min_group = 8;     // samples sharing a beta string needed for a tile of the block kernel
  int block_enable = 1;
  int eval_tiles = 1;          // 1: large calls evaluate 32 samples per warp; 2: every call does; 0: one warp per sample
  int block_parts = 0;         // block kernel: parts a tile's alpha-beta groups are split into (0: by the number of tiles; 1, 2, 4)
  int lut_pipeline = 1;        // lut_indexed_kernel: next query and its directory slot fetched ahead (0: one query at a time)
};
ElocTuning &eloc_tuning();

// alpha / beta strings of a one-word ONV from their 32-bit folds (inverse of fold_alpha / fold_beta)
__device__ __forceinline__ u64 unfold_alpha_word(u32 f) { return (u64)(f & 0x55555555u) | ((u64)(f & 0xAAAAAAAAu) << 31); }
__device__ __forceinline__ u64 unfold_beta_word(u32 f) { return ((u64)(f & 0x55555555u) << 1) | ((u64)(f & 0xAAAAAAAAu) << 32); }

}  // namespace pynqs
