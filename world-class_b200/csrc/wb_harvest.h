// Internal interface of the Harvest stage (wb_harvest.cu + wb_harvest_tail.cu).
#pragma once
#include <vector>

#include "wb_internal.h"

struct WbHarvestOptionInternal {
  double f0_floor, f0_ceil, frame_period, target_fs, channels_in_octave;
};

struct WbHarvestPlan {
  int fs;
  WbHarvestOptionInternal opt;
  int decimation_ratio;
  double actual_fs;
  int nch;
  std::vector<double> boundary_f0;  // harvest.cpp:1393-1396
  std::vector<int> half_len;        // filter_length_half per channel, harvest.cpp:1264
  int h_max;
  int max_candidates;               // harvest.cpp:1418-1419
  int NB, V;                        // overlap-save block size and hop
  double decim_coef[5];             // a[3], b[2]
  bool filters_ready;
  // frame-time table of the interval stage (valid for this length / period / buffer)
  int ttab_len = 0, ttab_period = 0;
  const void *ttab_ptr = nullptr;
  // auxiliary stream: table clears that nothing before the candidate stage depends on run beside the chain
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
  WbHarvestPlan() = default;
  WbHarvestPlan(const WbHarvestPlan &) = delete;
  WbHarvestPlan &operator=(const WbHarvestPlan &) = delete;
  ~WbHarvestPlan() {
    if (aux_fork) cudaEventDestroy(aux_fork);
    if (aux_join) cudaEventDestroy(aux_join);
    if (aux_stream) cudaStreamDestroy(aux_stream);
  }
};

int wb_harvest_plan_init(WbHarvestPlan *pl, int fs, const WbHarvestOptionInternal &opt);

// x on the device -> f0 on a `frame_period` ms grid (the reference always calls this with 1 ms)
int wb_harvest_run_basic(WbHarvestPlan *pl, WbWorkspace *ws, const double *d_x, int x_length, int frame_period,
                         double *d_f0_basic, int *f0_length_out, cudaStream_t stream);

// contour fixing + smoothing on the pruned candidate table (wb_harvest_tail.cu)
int wb_harvest_tail(WbWorkspace *ws, const double *d_cand, const double *d_score, const int *d_nc, int f0_length,
                    int max_candidates, double *d_f0_out, cudaStream_t stream);

// f0[i] = basic_f0[min(basic_len - 1, round(i * frame_period))] and the temporal positions (harvest.cpp:199-204)
int wb_harvest_pick(const double *d_basic_f0, int basic_len, double frame_period, int f0_length, double *d_tpos,
                    double *d_f0, cudaStream_t stream);
