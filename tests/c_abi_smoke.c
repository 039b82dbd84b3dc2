/* c_abi_smoke.c — the engine driven from plain C through include/pigeons_b200.h only (no Python, no C++):
 * toy_mvn_target(2), 10 chains, SliceSampler, rounds 1..6 on the default schedule (BASELINE config 1 without adaptation).
 * Prints the per-pair swap acceptance of the last round and the stepping-stone estimate as JSON; tests/test_gpu_parity.py
 * runs it on the GPU box and compares the numbers with the same calls made through the ctypes binding.
 * Built by __graft_entry__.build():  gcc -std=c11 -I include tests/c_abi_smoke.c -o tests/c_abi_smoke.bin -ldl -lm */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pigeons_b200.h"

#define N_CHAINS 10
#define DIM 2

typedef int (*create_fn)(const pgn_config*, pgn_handle**, char**);
typedef int (*h_err_fn)(pgn_handle*, char**);
typedef int (*sched_fn)(pgn_handle*, const double*, int32_t, char**);
typedef int (*expl_fn)(pgn_handle*, const pgn_explorer_params*, char**);
typedef int (*round_fn)(pgn_handle*, int64_t, pgn_round_out*, char**);
typedef int (*destroy_fn)(pgn_handle*);
typedef void (*free_fn)(char*);

static void* must(void* lib, const char* name) {
  void* p = dlsym(lib, name);
  if (!p) { fprintf(stderr, "missing symbol %s\n", name); exit(2); }
  return p;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s path/to/libpigeons_b200.so\n", argv[0]); return 2; }
  void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
  create_fn pgn_create_p = (create_fn)must(lib, "pgn_create");
  h_err_fn pgn_init_p = (h_err_fn)must(lib, "pgn_init_replicas");
  sched_fn pgn_sched_p = (sched_fn)must(lib, "pgn_set_schedule");
  expl_fn pgn_expl_p = (expl_fn)must(lib, "pgn_set_explorer");
  round_fn pgn_round_p = (round_fn)must(lib, "pgn_run_round");
  destroy_fn pgn_destroy_p = (destroy_fn)must(lib, "pgn_destroy");
  free_fn pgn_free_p = (free_fn)must(lib, "pgn_free_string");

  pgn_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.abi_version = PGN_ABI_VERSION;
  cfg.target_kind = PGN_TARGET_TOY_MVN;
  cfg.dim = DIM; cfg.n_chains = N_CHAINS; cfg.seed = 1; cfg.rank = 0; cfg.world_size = 1; cfg.device = 0;
  cfg.p[0] = 1.0; cfg.p[1] = 10.0;                       /* precision0, precision1 (toy_mvn_target.jl:8) */
  cfg.recorder_order = PGN_RECORDERS_PER_REPLICA;
  char* err = NULL;
  pgn_handle* h = NULL;
  int rc = pgn_create_p(&cfg, &h, &err);
  if (rc != PGN_OK) { printf("{\"rc\": %d, \"error\": \"%s\"}\n", rc, err ? err : ""); if (err) pgn_free_p(err); return rc == PGN_ERR_NO_DEVICE ? 0 : 1; }
  pgn_explorer_params ep;
  memset(&ep, 0, sizeof(ep));
  ep.kind = PGN_EXPLORER_SLICE; ep.slice_w = 10.0; ep.slice_p = 20; ep.slice_n_passes = 3; ep.slice_max_iter = 1024;
  double beta[N_CHAINS];
  for (int i = 0; i < N_CHAINS; ++i) beta[i] = (double)i / (N_CHAINS - 1);
  if ((rc = pgn_init_p(h, &err)) || (rc = pgn_sched_p(h, beta, N_CHAINS, &err)) || (rc = pgn_expl_p(h, &ep, &err))) {
    fprintf(stderr, "rc=%d %s\n", rc, err ? err : ""); return 1;
  }
  int64_t swap_n[N_CHAINS], i64[5][N_CHAINS];
  double swap_mean[N_CHAINS], ls_f[N_CHAINS], ls_b[N_CHAINS], f64[4][N_CHAINS], on_mean[DIM], on_var[DIM];
  pgn_round_out out;
  for (int round = 1; round <= 6; ++round) {
    memset(&out, 0, sizeof(out));
    out.swap_n = swap_n; out.swap_mean = swap_mean; out.logsum_fwd = ls_f; out.logsum_bwd = ls_b;
    out.expl_acc_n = i64[0]; out.expl_acc_mean = f64[0]; out.expl_n_steps = i64[1];
    out.am_n = i64[2]; out.am_mean = f64[1]; out.rev_n = i64[3]; out.rev_mean = f64[2];
    out.online_mean = on_mean; out.online_var = on_var;
    rc = pgn_round_p(h, (int64_t)1 << round, &out, &err);
    if (rc != PGN_OK) { fprintf(stderr, "pgn_run_round rc=%d %s\n", rc, err ? err : ""); if (err) pgn_free_p(err); return 1; }
  }
  double e1 = 0.0, e2 = 0.0;                              /* stepping_stone_pair (src/evidence/stepping_stone.jl:28-43) */
  for (int i = 0; i < N_CHAINS - 1; ++i) { e1 += ls_f[i] - log((double)swap_n[i]); e2 += ls_b[i] - log((double)swap_n[i]); }
  printf("{\"rc\": 0, \"n_round_trips\": %lld, \"stepping_stone\": %.17g, \"swap_mean\": [", (long long)out.n_round_trips, 0.5 * (e1 - e2));
  for (int i = 0; i < N_CHAINS - 1; ++i) printf("%s%.17g", i ? ", " : "", swap_mean[i]);
  printf("], \"online_mean\": [%.17g, %.17g], \"kernel_ms\": %.4f}\n", on_mean[0], on_mean[1], out.kernel_ms);
  pgn_destroy_p(h);
  dlclose(lib);
  return 0;
}
