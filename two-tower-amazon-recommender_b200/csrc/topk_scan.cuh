// Plan of the threshold-scan form of the bf16 brute-force top-k (topk_scan.cu), shared with the workspace sizing.
#pragma once
#include "common.cuh"

namespace tt {

struct ScanPlan {
  bool use;                 // false: shape outside the scan path (d > 128, too few candidates for a sample, huge k)
  int nqt;                  // query tiles (128 rows) per CTA: 2 when more than one tile exists
  int q_groups, nq_pad;
  int total_tiles;          // 128-candidate tiles
  int n_samp, G, P;         // sampled tiles, group maxima per row (4 per tile), sort width of the threshold kernel
  int cap, kp;              // survivors per row the buffers hold; pool width
  int n_slots, seg_cap;     // the row's buffer = one segment per candidate split of either phase
  int samp_splits, samp_tps, scan_splits, scan_tps;
  bool two_phase;           // scan 1/8 of the tiles, refine the threshold on their survivors, scan the rest
  int first_splits, first_tps;
  int64_t off_samp, off_tau, off_cnt, off_flag, off_buf, bytes;
};

ScanPlan tc_topk_scan_plan(int64_t nq, int64_t nc, int64_t d, int kp);
int tc_topk_scan(const ScanPlan& p, const void* queries, const void* candidates, int64_t nq, int64_t nc, int64_t d,
                 float* pool_s, int64_t* pool_i, void* ws, int** flag_out, cudaStream_t st);

}  // namespace tt
