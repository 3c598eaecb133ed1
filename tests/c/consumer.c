/* A consumer of the C ABI written in plain C99 (TEST ONLY): no Python, no torch, no CUDA headers.  It binds
 * libqvc_b200.so the way a maintainer of a non-Python host would (dlopen + the prototypes of include/qvc_b200.h) and
 * exercises the host-only entry points: ABI version, size queries, the weight fold's argument checking and its error
 * string.  Compute entry points need a B200 and are covered by the `-m gpu` tests through the same ABI. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qvc_b200.h"

#define BIND(name)                                                        \
  do {                                                                    \
    *(void**)(&p_##name) = dlsym(lib, #name);                             \
    if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; } \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: consumer <libqvc_b200.so>\n"); return 2; }
  void* lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }

  int (*p_qvc_abi_version)(void);
  const char* (*p_qvc_last_error)(void);
  size_t (*p_qvc_prepared_bytes)(int);
  int (*p_qvc_fold_host)(const qvc_state_entry*, int, int, int, void*, size_t, float*, qvc_model*);
  BIND(qvc_abi_version);
  BIND(qvc_last_error);
  BIND(qvc_prepared_bytes);
  BIND(qvc_fold_host);

  if (p_qvc_abi_version() != QVC_ABI_VERSION) {
    fprintf(stderr, "ABI version %d, header says %d\n", p_qvc_abi_version(), QVC_ABI_VERSION);
    return 1;
  }
  const size_t n32 = p_qvc_prepared_bytes(QVC_OPF_TF32), n16 = p_qvc_prepared_bytes(QVC_OPF_BF16);
  if (n32 == 0 || n16 == 0 || n16 >= n32 || p_qvc_prepared_bytes(99) != 0) {
    fprintf(stderr, "qvc_prepared_bytes: tf32 %zu, bf16 %zu, bad format %zu\n", n32, n16, p_qvc_prepared_bytes(99));
    return 1;
  }

  /* a state_dict with one entry of the wrong size: the fold must refuse it and say which key */
  static float w[16];
  qvc_state_entry e;
  e.name = "enc_p.pre.weight";
  e.data = w;
  e.numel = 16;
  void* block = malloc(n32);
  qvc_model model;
  memset(&model, 0, sizeof model);
  const int rc = p_qvc_fold_host(&e, 1, QVC_OPF_TF32, QVC_BACKEND_TCGEN05, block, n32, NULL, &model);
  const char* msg = p_qvc_last_error();
  free(block);
  if (rc != QVC_ERR_ARG || msg == NULL || strstr(msg, "enc_p.") == NULL) {
    fprintf(stderr, "qvc_fold_host: rc %d, error '%s'\n", rc, msg ? msg : "(null)");
    return 1;
  }
  /* what the C compiler makes of the header's declarations: a binding in another language checks its own mirror against this */
  printf("sizeof qvc_tensor=%zu qvc_epi_segment=%zu qvc_conv_args=%zu qvc_spk_weights=%zu qvc_mel_weights=%zu "
         "qvc_tail_weights=%zu qvc_layer=%zu qvc_model=%zu qvc_state_entry=%zu qvc_taps=%zu\n",
         sizeof(qvc_tensor), sizeof(qvc_epi_segment), sizeof(qvc_conv_args), sizeof(qvc_spk_weights), sizeof(qvc_mel_weights),
         sizeof(qvc_tail_weights), sizeof(qvc_layer), sizeof(qvc_model), sizeof(qvc_state_entry), sizeof(qvc_taps));
  printf("ok abi=%d prepared_bytes tf32=%zu bf16=%zu fold_error='%s'\n", p_qvc_abi_version(), n32, n16, msg);
  dlclose(lib);
  return 0;
}
