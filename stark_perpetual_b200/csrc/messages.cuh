// Packing of the perpetual messages other than limit orders -- transfer, conditional transfer, withdrawal to address,
// oracle price -- into the elements of their Pedersen chains; per-message code shared by the CUDA kernel (orders.cu)
// and the host emulation (tests/host_emul/emul_messages.cpp).
//
// Reference restated: src/services/perpetual/public/perpetual_messages.py
//   :24-94   get_conditional_transfer_msg   H(H(H(H(H(asset, asset_fee), receiver_key), condition), w0), w1)
//   :97-162  get_transfer_msg               H(H(H(H(asset, asset_fee), receiver_key), w0), w1)
//            w0 = sender_position | receiver_position | src_fee_position | nonce                (64, 64, 64, 32 bits)
//            w1 = type | amount | max_amount_fee | expiration | 81 zero bits                    (type 4 / 5; 64, 64, 32)
//   :165-209 get_withdrawal_to_address_msg  H(H(asset_collateral, eth_address), w)
//            w  = 7 | position | nonce | amount | expiration | 49 zero bits                     (64, 32, 64, 32)
//   :311-326 get_price_msg                  H(asset_pair << 40 | oracle_name, price << 32 | timestamp)
// (Cairo twins: src/services/exchange/cairo/signature_message_hashes.cairo:106-170,
//  src/services/perpetual/cairo/oracle/oracle_price.cairo:96-108.)  The bounds the reference asserts are reported as
// status 1 instead of raised.
#pragma once
#include <stdint.h>

#include "fp.cuh"

#define SPG_MSG_TRANSFER 4
#define SPG_MSG_CONDITIONAL_TRANSFER 5
#define SPG_MSG_WITHDRAWAL_TO_ADDRESS 7
#define SPG_MSG_PRICE 100
#define SPG_MSG_MAX_FELTS 4
#define SPG_MSG_MAX_INTS 7

// chain length (number of hashed elements) of a message kind; 0 = unknown kind
SPG_HD int spg_msg_chain_len(int kind) {
  return kind == SPG_MSG_TRANSFER ? 5 : kind == SPG_MSG_CONDITIONAL_TRANSFER ? 6 : kind == SPG_MSG_WITHDRAWAL_TO_ADDRESS ? 3
         : kind == SPG_MSG_PRICE ? 2 : 0;
}
SPG_HD int spg_msg_n_felts(int kind) {
  return kind == SPG_MSG_TRANSFER ? 3 : kind == SPG_MSG_CONDITIONAL_TRANSFER ? 4 : kind == SPG_MSG_PRICE ? 2 : 2;
}
SPG_HD int spg_msg_n_ints(int kind) {
  return (kind == SPG_MSG_TRANSFER || kind == SPG_MSG_CONDITIONAL_TRANSFER) ? 7 : kind == SPG_MSG_WITHDRAWAL_TO_ADDRESS ? 4 : 2;
}

// OR a 64-bit value into a 256-bit little-endian word at bit offset `off`
SPG_HD void spg_put_bits(uint64_t (&w)[4], uint64_t v, int off) {
  const int k = off >> 6, sh = off & 63;
  w[k] |= v << sh;
  if (sh && k + 1 < 4) w[k + 1] |= v >> (64 - sh);
}
// value (4 x u64 little-endian) < 2^bits ?
SPG_HD bool spg_below_pow2(const uint64_t* v, int bits) {
  for (int k = 3; k >= 0; k--) {
    const int lo = 64 * k;
    if (bits <= lo) { if (v[k]) return false; }
    else if (bits < lo + 64) { if (v[k] >> (bits - lo)) return false; }
  }
  return true;
}

// felts: the message's full-width arguments in the reference's order (each 4 x u64); ints: its narrow arguments in the
// order listed in spg.h.  elems: chain_len x 4 words.  Returns 0, or 1 when a bound of the reference is violated.
SPG_HD int spg_pack_message(int kind, const uint64_t* const* felts, const uint64_t* ints, uint64_t* elems) {
  int st = 0;
  uint64_t w0[4] = {0, 0, 0, 0}, w1[4] = {0, 0, 0, 0};
  if (kind == SPG_MSG_TRANSFER || kind == SPG_MSG_CONDITIONAL_TRANSFER) {
    const int nf = kind == SPG_MSG_TRANSFER ? 3 : 4;
    // :38-48 / :110-119: asset ids below 2^250, receiver key and condition below 2^251, nonce and expiration below 2^32
    if (!spg_below_pow2(felts[0], 250) || !spg_below_pow2(felts[1], 250) || !spg_below_pow2(felts[2], 251)) st = 1;
    if (nf == 4 && !spg_below_pow2(felts[3], 251)) st = 1;
    // ints: sender_position_id, receiver_position_id, src_fee_position_id, nonce, amount, max_amount_fee, expiration
    if ((ints[3] >> 32) | (ints[6] >> 32)) st = 1;
    spg_put_bits(w0, ints[3], 0); spg_put_bits(w0, ints[2], 32); spg_put_bits(w0, ints[1], 96); spg_put_bits(w0, ints[0], 160);
    spg_put_bits(w1, ints[6], 81); spg_put_bits(w1, ints[5], 113); spg_put_bits(w1, ints[4], 177);
    spg_put_bits(w1, (uint64_t)kind, 241);
    for (int f = 0; f < nf; f++)
      for (int k = 0; k < 4; k++) elems[4 * f + k] = felts[f][k];
    for (int k = 0; k < 4; k++) { elems[4 * nf + k] = w0[k]; elems[4 * (nf + 1) + k] = w1[k]; }
  } else if (kind == SPG_MSG_WITHDRAWAL_TO_ADDRESS) {
    // :174-179: collateral id below 2^250, eth address below 2^160; ints: position_id, nonce, amount, expiration
    if (!spg_below_pow2(felts[0], 250) || !spg_below_pow2(felts[1], 160)) st = 1;
    if ((ints[1] >> 32) | (ints[3] >> 32)) st = 1;
    spg_put_bits(w0, ints[3], 49); spg_put_bits(w0, ints[2], 81); spg_put_bits(w0, ints[1], 145); spg_put_bits(w0, ints[0], 177);
    spg_put_bits(w0, (uint64_t)SPG_MSG_WITHDRAWAL_TO_ADDRESS, 241);
    for (int k = 0; k < 4; k++) { elems[k] = felts[0][k]; elems[4 + k] = felts[1][k]; elems[8 + k] = w0[k]; }
  } else {
    // :314-317: felts = asset_pair (< 2^128), price (< 2^120); ints = oracle_name (< 2^40), timestamp (< 2^32)
    if (!spg_below_pow2(felts[0], 128) || !spg_below_pow2(felts[1], 120)) st = 1;
    if ((ints[0] >> 40) | (ints[1] >> 32)) st = 1;
    spg_put_bits(w0, ints[0], 0); spg_put_bits(w0, felts[0][0], 40); spg_put_bits(w0, felts[0][1], 104);
    spg_put_bits(w1, ints[1], 0); spg_put_bits(w1, felts[1][0], 32); spg_put_bits(w1, felts[1][1], 96);
    for (int k = 0; k < 4; k++) { elems[k] = w0[k]; elems[4 + k] = w1[k]; }
  }
  return st;
}
