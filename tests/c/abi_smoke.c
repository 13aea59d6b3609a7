/* A plain-C consumer of include/pepflow_b200.h: proves the header is valid C99 (no C++-isms, no torch types) and that
 * libpepflow_b200.so can be bound with nothing but dlopen - what a cgo / JNI / ctypes binding does.  No GPU needed:
 * only the argument-checking paths are exercised.   gcc -std=c99 -Wall -Werror -I include tests/c/abi_smoke.c -ldl */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "pepflow_b200.h"

#define BIND(name) \
  *(void**)(&p_##name) = dlsym(h, #name); \
  if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

int main(int argc, char** argv) {
  if (argc < 2) return 64;
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "%s\n", dlerror()); return 1; }
  int (*p_pf_version)(void);
  const char* (*p_pf_strerror)(int);
  int (*p_pf_check_config)(int, int, int, int, int, int, int, int);
  int (*p_pf_so3_log)(const float*, float*, int, void*);
  int (*p_pf_full_atom_reconstruction)(const float*, const float*, const float*, const int64_t*, const float*,
                                       const float*, const int32_t*, const float*, const uint8_t*, float*, float*,
                                       float*, uint8_t*, long long, void*);
  int (*p_pf_torsion_angles)(const float*, const int64_t*, const int32_t*, float*, uint8_t*, long long, int, void*);
  size_t (*p_pf_ga_encoder_workspace_bytes)(int, int);
  BIND(pf_version) BIND(pf_strerror) BIND(pf_check_config) BIND(pf_so3_log) BIND(pf_full_atom_reconstruction)
  BIND(pf_torsion_angles) BIND(pf_ga_encoder_workspace_bytes)
  if (p_pf_version() != 4) return 3;
  if (strcmp(p_pf_strerror(PF_OK), "ok") != 0 || !strlen(p_pf_strerror(PF_ERR_NULL_POINTER))) return 4;
  if (p_pf_check_config(128, 64, 128, 8, 8, 12, 4, 2) != PF_OK) return 5;
  if (p_pf_check_config(256, 64, 128, 8, 8, 12, 4, 2) != PF_ERR_BAD_CONFIG) return 6;
  if (p_pf_so3_log(NULL, NULL, 4, NULL) != PF_ERR_NULL_POINTER) return 7;
  if (p_pf_full_atom_reconstruction(NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 8,
                                    NULL) != PF_ERR_NULL_POINTER) return 8;
  if (p_pf_torsion_angles(NULL, NULL, NULL, NULL, NULL, 8, 13, NULL) != PF_ERR_BAD_SHAPE) return 9;
  if (p_pf_ga_encoder_workspace_bytes(64, 271) < ((size_t)64 * 271 * 271 * 64 * 4)) return 10;
  printf("abi ok: version %d, workspace(64,271) = %zu bytes\n", p_pf_version(), p_pf_ga_encoder_workspace_bytes(64, 271));
  dlclose(h);
  return 0;
}
