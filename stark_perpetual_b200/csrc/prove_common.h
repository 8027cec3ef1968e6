// Host-side pieces shared by the single-GPU prover (prove.cu) and the sharded prover (sharded.cu): serialisation,
// the Fiat-Shamir channel, a bump allocator.  Protocol: DESIGN.md "Protocol"; CPU restatement: oracle/stark.py.
#pragma once
#include <string.h>

#include <vector>

#include "blake2s.cuh"
#include "stark_kernels.cuh"
#include "ecdsa_air_point.cuh"

// ------------------------------------------------------------------ host helpers
static void ser_fp(const Fp& a /*canonical representative of the Montgomery form*/, uint8_t* out) {
  for (int k = 0; k < 8; k++) {
    const uint32_t w = a.v[7 - k];
    out[4 * k] = (uint8_t)(w >> 24); out[4 * k + 1] = (uint8_t)(w >> 16); out[4 * k + 2] = (uint8_t)(w >> 8); out[4 * k + 3] = (uint8_t)w;
  }
}
static void put_u32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; i++) v.push_back((uint8_t)(x >> (8 * i))); }
static void put_fp(std::vector<uint8_t>& v, const Fp& a) { uint8_t b[32]; ser_fp(a, b); v.insert(v.end(), b, b + 32); }
static void put_bytes(std::vector<uint8_t>& v, const uint8_t* p, size_t n) { v.insert(v.end(), p, p + n); }

struct Channel {
  uint8_t state[32];
  uint64_t counter = 0;
  explicit Channel(const std::vector<uint8_t>& seed) {
    std::vector<uint8_t> b;
    const char* tag = "spg-stark-v1";
    b.insert(b.end(), tag, tag + 12);
    b.insert(b.end(), seed.begin(), seed.end());
    b2s_hash_bytes(b.data(), b.size(), state);
  }
  void absorb(const uint8_t* data, size_t len) {
    std::vector<uint8_t> b(state, state + 32);
    b.insert(b.end(), data, data + len);
    b2s_hash_bytes(b.data(), b.size(), state);
    counter = 0;
  }
  void draw(uint8_t out[32]) {
    uint8_t b[40];
    memcpy(b, state, 32);
    for (int i = 0; i < 8; i++) b[32 + i] = (uint8_t)(counter >> (8 * i));
    b2s_hash_bytes(b, 40, out);
    counter++;
  }
  Fp draw_felt() {   // Montgomery form of the drawn value (251 bits, always < p)
    uint8_t d[32];
    draw(d);
    uint64_t w[4];
    for (int k = 0; k < 4; k++) { w[k] = 0; for (int b = 7; b >= 0; b--) w[k] = (w[k] << 8) | d[8 * k + b]; }
    w[3] &= (1ull << 59) - 1;
    return spg_host_from_u64(w);
  }
  uint64_t draw_index(uint64_t n) {
    uint8_t d[32];
    draw(d);
    uint64_t v = 0;
    for (int b = 7; b >= 0; b--) v = (v << 8) | d[b];
    return v % n;
  }
};

// which AIR a proof is for: the protocol around it is the same (25 columns, mask {x, x w_N}, 4 composition chunks)
struct AirSpec {
  int kind = 1;                         // proof header VERSION: 1 = Pedersen hash chain, 2 = ECDSA builtin
  unsigned chain_log = 0;               // kind 1
  const uint64_t* x0_canon = nullptr;   // kind 1: the 5 lane seeds
  const uint64_t* msgs_canon = nullptr; // kind 2: the public input, [N/256] message hashes ...
  const uint64_t* keys_canon = nullptr; //         ... and [N/256] keys' x (canonical, host)
};
#define SPG_MAX_ALPHA (SPG_AIR_LANES * SPG_AIR_NCONSTR)
static_assert(SPG_EAIR_NALPHA <= SPG_MAX_ALPHA && SPG_EAIR_COLS == SPG_AIR_COLS, "both AIRs share the protocol's shape");

// the statement of a kind-2 proof into the channel seed and the proof header (oracle/stark_ecdsa.py EcdsaAir.seed / .header)
static void put_ecdsa_statement(const AirSpec& air, unsigned log_n, unsigned n_queries, std::vector<uint8_t>& seed,
                                std::vector<uint8_t>& proof) {
  const char* tag = "ecdsa-builtin";
  seed.insert(seed.end(), tag, tag + 13);
  put_u32(seed, log_n); put_u32(seed, n_queries);
  for (size_t b = 0; b < (((size_t)1 << log_n) >> 8); b++) {
    const Fp m = spg_host_from_u64(air.msgs_canon + 4 * b), k = spg_host_from_u64(air.keys_canon + 4 * b);
    put_fp(seed, m); put_fp(seed, k); put_fp(proof, m); put_fp(proof, k);
  }
}

// bump allocator over one cached device block
struct Arena {
  char* base; size_t cap, off;
  template <class T> T* get(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    if (off + bytes > cap) return nullptr;
    T* p = (T*)(base + off);
    off += bytes;
    return p;
  }
};

enum { ST_LDE = 0, ST_MERKLE_T, ST_AIR, ST_HLDE, ST_MERKLE_H, ST_OODS, ST_DEEP, ST_FRI, ST_QUERY, ST_H2D, ST_COMM };
