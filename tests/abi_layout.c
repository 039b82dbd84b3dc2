/* abi_layout.c — prints sizeof / offsetof of every struct of include/pigeons_b200.h as JSON.
 * tests/test_abi_layout.py compares the output with the ctypes mirror (pigeons.jl_b200/_capi.py) and with the
 * field lists of the Julia mirror (julia/PigeonsB200.jl), so the three descriptions of the ABI cannot drift apart.
 * Built by __graft_entry__.build() with plain gcc: the header is C. */
#include <stddef.h>
#include <stdio.h>

#include "pigeons_b200.h"

#define FIELD(S, f) printf("%s[\"%s\", %zu, %zu]", first ? "" : ", ", #f, offsetof(S, f), sizeof(((S*)0)->f)), first = 0
#define BEGIN(S) printf("%s\"%s\": {\"size\": %zu, \"fields\": [", first_struct ? "" : ", ", #S, sizeof(S)), first = 1, first_struct = 0
#define END() printf("]}")

int main(void) {
  int first = 1, first_struct = 1;
  printf("{\"abi_version\": %d, \"structs\": {", PGN_ABI_VERSION);
  BEGIN(pgn_config);
  FIELD(pgn_config, abi_version); FIELD(pgn_config, target_kind); FIELD(pgn_config, dim); FIELD(pgn_config, n_chains);
  FIELD(pgn_config, seed); FIELD(pgn_config, rank); FIELD(pgn_config, world_size); FIELD(pgn_config, device);
  FIELD(pgn_config, n_modes); FIELD(pgn_config, p); FIELD(pgn_config, means); FIELD(pgn_config, log_weights);
  FIELD(pgn_config, data_x); FIELD(pgn_config, data_y); FIELD(pgn_config, recorder_order); FIELD(pgn_config, n_chains_variational);
  END();
  BEGIN(pgn_explorer_params);
  FIELD(pgn_explorer_params, kind); FIELD(pgn_explorer_params, slice_w); FIELD(pgn_explorer_params, slice_p);
  FIELD(pgn_explorer_params, slice_n_passes); FIELD(pgn_explorer_params, slice_max_iter); FIELD(pgn_explorer_params, n_refresh);
  FIELD(pgn_explorer_params, step_size); FIELD(pgn_explorer_params, precond_kind); FIELD(pgn_explorer_params, mix_p0);
  FIELD(pgn_explorer_params, mix_p01); FIELD(pgn_explorer_params, std_devs); FIELD(pgn_explorer_params, ising_n_steps);
  FIELD(pgn_explorer_params, n_steps); FIELD(pgn_explorer_params, step_kind);
  FIELD(pgn_explorer_params, n_mix); FIELD(pgn_explorer_params, mix_n_refresh); FIELD(pgn_explorer_params, mix_precond_kind);
  FIELD(pgn_explorer_params, mix_step_size); FIELD(pgn_explorer_params, mix_variant_p0); FIELD(pgn_explorer_params, mix_variant_p01);
  END();
  BEGIN(pgn_round_out);
  FIELD(pgn_round_out, swap_n); FIELD(pgn_round_out, swap_mean); FIELD(pgn_round_out, logsum_fwd); FIELD(pgn_round_out, logsum_bwd);
  FIELD(pgn_round_out, expl_acc_n); FIELD(pgn_round_out, expl_acc_mean); FIELD(pgn_round_out, expl_n_steps);
  FIELD(pgn_round_out, am_n); FIELD(pgn_round_out, am_mean); FIELD(pgn_round_out, rev_n); FIELD(pgn_round_out, rev_mean);
  FIELD(pgn_round_out, n_tempered_restarts); FIELD(pgn_round_out, n_round_trips); FIELD(pgn_round_out, online_n);
  FIELD(pgn_round_out, online_mean); FIELD(pgn_round_out, online_var); FIELD(pgn_round_out, index_process);
  FIELD(pgn_round_out, swap_lr); FIELD(pgn_round_out, swap_u); FIELD(pgn_round_out, swap_accept); FIELD(pgn_round_out, target_trace);
  FIELD(pgn_round_out, n_density_points); FIELD(pgn_round_out, n_ref_equiv_evals); FIELD(pgn_round_out, kernel_ms);
  FIELD(pgn_round_out, gemm_ms); FIELD(pgn_round_out, batch_steps); FIELD(pgn_round_out, n_launches);
  FIELD(pgn_round_out, active_columns); FIELD(pgn_round_out, gemm_columns);
  END();
  BEGIN(pgn_replica_state);
  FIELD(pgn_replica_state, x); FIELD(pgn_replica_state, replica_index); FIELD(pgn_replica_state, rng_counter);
  FIELD(pgn_replica_state, round_trip_state);
  END();
  BEGIN(pgn_device_info_t);
  FIELD(pgn_device_info_t, sm_major); FIELD(pgn_device_info_t, sm_minor); FIELD(pgn_device_info_t, n_sms);
  FIELD(pgn_device_info_t, global_mem_bytes); FIELD(pgn_device_info_t, max_resident_chains); FIELD(pgn_device_info_t, name);
  END();
  printf("}}\n");
  return 0;
}
