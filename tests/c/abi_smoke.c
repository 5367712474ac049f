/* Plain-C consumer of include/caco_b200.h: proves the header compiles as C and the library links and answers without
 * Python or torch.  Only host-side entry points are called (no GPU needed): version, architecture, the HTK filterbank
 * (torchaudio.functional.melscale_fbanks as called at src/eval/eval_caco_torch.py:94-101) and argument validation.
 * Build: gcc -std=c99 -Iinclude tests/c/abi_smoke.c -Lcacophony_b200 -lcaco_b200 -Wl,-rpath,cacophony_b200 -o abi_smoke */
#include <stdio.h>
#include <stdlib.h>

#include "caco_b200.h"

int main(void) {
  float* fb = (float*)calloc(257 * 128, sizeof(float));
  int nz = 0, bin0 = 0, k, m;
  if (!fb) return 2;
  if (caco_version() != 100 || caco_built_arch() != 100) { printf("bad version\n"); return 1; }
  if (caco_mel_filterbank(fb) != 0) { printf("filterbank failed\n"); return 1; }
  for (k = 0; k < 257; ++k)
    for (m = 0; m < 128; ++m) {
      if (fb[k * 128 + m] != 0.0f) ++nz;
      if (m == 0 && fb[k * 128] != 0.0f) ++bin0;
    }
  if (nz != 505 || bin0 != 0) { printf("filterbank: %d non-zeros, %d in mel bin 0\n", nz, bin0); return 1; }
  /* NULL arguments are rejected with an error code, never a crash or an exception across the boundary */
  if (caco_frontend(NULL, 1, 16000, 100, NULL, NULL, NULL, NULL, NULL, NULL, NULL) == 0) return 1;
  if (caco_topk_rows(NULL, 1, 1, 1, 1, NULL, NULL, NULL) == 0) return 1;
  if (caco_model_audio_embedding(NULL, NULL, NULL, NULL, NULL, 1, 1, 0, NULL, NULL, NULL) == 0) return 1;
  if (caco_model_decode_cache_bytes(NULL, 1, 500, 32) != 0) return 1;
  if (caco_model_decode_begin(NULL, NULL, 0, NULL, NULL, 1, 500, 32, NULL) == 0) return 1;
  if (caco_model_decode_step(NULL, NULL, NULL, NULL, 1, 500, 32, NULL, NULL, NULL) == 0) return 1;
  printf("abi ok: version %d, %d filterbank non-zeros, launches so far %lld\n", caco_version(), nz, (long long)caco_launch_count());
  free(fb);
  return 0;
}
