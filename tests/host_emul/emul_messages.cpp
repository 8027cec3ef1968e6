// CPU run of the perpetual-message packing code of the CUDA kernel (csrc/messages.cuh), compiled with g++.
// stdin: lines "kind felt_hex... int_dec..." (as many felts / integers as the kind takes).
// stdout: "status elem_hex ..." (the elements of the Pedersen chain, canonical).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../stark_perpetual_b200/csrc/messages.cuh"

static void parse_hex(const char* s, uint64_t w[4]) {
  memset(w, 0, 32);
  size_t n = strlen(s);
  for (size_t i = 0; i < n && i < 64; i++) {
    char c = s[n - 1 - i];
    uint64_t d = (c >= '0' && c <= '9') ? c - '0' : (c >= 'a' && c <= 'f') ? c - 'a' + 10 : c - 'A' + 10;
    w[i / 16] |= d << (4 * (i % 16));
  }
}
int main() {
  int kind;
  char tok[128];
  while (scanf("%d", &kind) == 1) {
    const int nf = spg_msg_n_felts(kind), ni = spg_msg_n_ints(kind), len = spg_msg_chain_len(kind);
    if (!len) return 1;
    uint64_t felts[SPG_MSG_MAX_FELTS][4], ints[SPG_MSG_MAX_INTS], elems[6 * 4];
    const uint64_t* fp[SPG_MSG_MAX_FELTS] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < nf; k++) { if (scanf("%100s", tok) != 1) return 1; parse_hex(tok, felts[k]); fp[k] = felts[k]; }
    for (int k = 0; k < SPG_MSG_MAX_INTS; k++) ints[k] = 0;
    for (int k = 0; k < ni; k++) { if (scanf("%100s", tok) != 1) return 1; ints[k] = strtoull(tok, nullptr, 10); }
    const int st = spg_pack_message(kind, fp, ints, elems);
    printf("%d", st);
    for (int e = 0; e < len; e++)
      printf(" %016llx%016llx%016llx%016llx", (unsigned long long)elems[4 * e + 3], (unsigned long long)elems[4 * e + 2],
             (unsigned long long)elems[4 * e + 1], (unsigned long long)elems[4 * e]);
    printf("\n");
  }
  return 0;
}
