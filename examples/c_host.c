/* Minimal C host of the C ABI (include/mmr_b200.h): plans the buffers of one route-fusion call and checks the struct
 * layout against the library, without touching a GPU.  Shows that the boundary is plain C99 -- no C++, no torch:
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude examples/c_host.c -Lmultimodalrouting_b200/csrc -lmmr_b200 \
 *       -Wl,-rpath,$PWD/multimodalrouting_b200/csrc -o /tmp/c_host && /tmp/c_host
 *
 * A real host would now cudaMalloc the four buffers, fill mmr_fusion_dims / the parameter pointer table (state_dict
 * order) and call mmr_route_fusion_fwd / mmr_route_fusion_bwd_ex on its stream. */
#include <stdio.h>

#include "mmr_b200.h"

int main(void) {
  mmr_fusion_dims d;
  size_t packed = 0, saved = 0, scratch_fwd = 0, scratch_bwd = 0, sz[7];
  int n;
  d.B = 512; d.TL = 48; d.TN = 16; d.TI = 49;      /* BASELINE configs[1] */
  d.dL = 256; d.dN = 256; d.dI = 256; d.layers = 4;
  d.dtype = MMR_DTYPE_BF16; d.gemm_engine = MMR_GEMM_AUTO;
  if (mmr_version() < 100) { fprintf(stderr, "unexpected version\n"); return 1; }
  if (mmr_fusion_sizes(&d, &packed, &saved, &scratch_fwd, &scratch_bwd) != MMR_OK) {
    fprintf(stderr, "mmr_fusion_sizes: %s\n", mmr_last_error_string());
    return 1;
  }
  n = mmr_abi_struct_sizes(sz, 7);
  if (n != 7 || sz[0] != sizeof(mmr_fusion_dims) || sz[1] != sizeof(mmr_routing_dims) ||
      sz[2] != sizeof(mmr_routing_params) || sz[3] != sizeof(mmr_routing_grads) || sz[4] != sizeof(mmr_opt_tensor) ||
      sz[5] != sizeof(mmr_opt_hyper) || sz[6] != sizeof(mmr_opt_state)) {
    fprintf(stderr, "struct layout differs between this translation unit and the library\n");
    return 1;
  }
  d.B = 0;                                           /* invalid: must be rejected with a message, not crash */
  if (mmr_fusion_sizes(&d, &packed, &saved, &scratch_fwd, &scratch_bwd) == MMR_OK || !mmr_last_error_string()[0]) return 1;
  printf("params=%d packed=%lu saved=%lu scratch_fwd=%lu scratch_bwd=%lu\n", (d.B = 512, mmr_fusion_num_params(&d)),
         (unsigned long)packed, (unsigned long)saved, (unsigned long)scratch_fwd, (unsigned long)scratch_bwd);
  return 0;
}
