// pgn_memchain_types.cuh — plain-data types of the memory-resident scan kernel shared with the host code.
#pragma once
#include "pgn_kernels.cuh"

namespace pgn {

struct MemRec {   // per local chain; lives in HBM between a warp's visits
  unsigned long long ctr;
  int replica_index, rt_state;
  double e0, e1;
  MeanAcc expl_acc, am, rev, swap_acc;
  LogSumAcc ls_fwd, ls_bwd;
  long long n_steps, n_points, n_ref, n_restarts, n_trips;
  long long on_n;
  double lr, u;
  int pad_;
};

struct MemParams {
  Params base;
  MemRec* rec;
  double *VP, *VG0, *VSX, *VSP, *VSG, *VTX, *VTP, *VTG, *VFX, *VFG;   // [n_local][d_pad]
  int nslots;
};

}  // namespace pgn
