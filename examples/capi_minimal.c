/* The C-ABI from plain C (C99): what any foreign-function binding sees.  One D3Q19 BGK step on a 16^3 box, observables
 * read back.  Without a CUDA device mlbm_create fails loudly (there is no CPU fallback) and the program says so.
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -I include examples/capi_minimal.c -L metalbm_b200 -lmetalbm_b200 \
 *       -Wl,-rpath,$PWD/metalbm_b200 -o capi_minimal
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "metalbm_b200.h"

int main(void) {
  mlbm_config config;
  memset(&config, 0, sizeof(config));
  config.abi_version = MLBM_ABI_VERSION;
  config.lattice = MLBM_D3Q19;
  config.collision = MLBM_BGK;
  config.equilibrium = MLBM_TRUNCATION_MA3;
  config.forcing_scheme = MLBM_GUO;
  config.force = MLBM_FORCE_KOLMOGOROV;
  config.dtype = MLBM_F64;
  config.overlap = MLBM_OVERLAP_OFF;
  config.global_length[0] = config.global_length[1] = config.global_length[2] = 16;
  config.rank = 0;
  config.nranks = 1;
  config.device = -1;
  config.tau = 0.6;
  config.force_amplitude[0] = 1e-5;
  config.force_wavelength[0] = config.force_wavelength[1] = config.force_wavelength[2] = 8.0;

  if (mlbm_abi_version() != MLBM_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 2; }
  mlbm_ctx* ctx = NULL;
  int status = mlbm_create(&config, &ctx);
  if (status != MLBM_OK) {
    printf("mlbm_create: status %d: %s\n", status, mlbm_last_error());
    return status == MLBM_ERR_CUDA ? 0 : 1;   /* no device: the expected, loud failure */
  }
  if (mlbm_init_synthetic(ctx, 0.05, 0.05) != MLBM_OK || mlbm_step(ctx, 1, 1) != MLBM_OK) {
    fprintf(stderr, "%s\n", mlbm_last_error());
    return 1;
  }
  double observables[4];
  if (mlbm_observables(ctx, observables) != MLBM_OK) { fprintf(stderr, "%s\n", mlbm_last_error()); return 1; }
  printf("energy %.12e enstrophy %.12e mach %.6f mass %.6f\n", observables[0], observables[1], observables[2], observables[3]);
  return mlbm_destroy(ctx) == MLBM_OK ? 0 : 1;
}
