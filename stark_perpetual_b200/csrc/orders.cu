// Batched perpetual limit orders on the device: field packing + the 4-deep Pedersen chain of the order message
// (+ optionally the STARK-curve ECDSA check of the order signature), with no host round trip in between.
// SURVEY.md section 8 rows a12 / f-1; BASELINE.json configs[4].
//
// Restates src/services/perpetual/public/perpetual_messages.py:212-286 (get_limit_order_msg and
// get_limit_order_msg_without_bounds; Cairo twin src/services/exchange/cairo/signature_message_hashes.cairo:56-91):
//   (sell, buy) = is_buying_synthetic ? (collateral, synthetic) : (synthetic, collateral)
//   msg = H(H(H(H(asset_sell, asset_buy), asset_fee), amount_sell | amount_buy | max_fee | nonce),
//           3 | position | position | position | expiration | 17 zero bits)
// with the bounds asserted at :226-236 reported per element instead of raised.
#include <string.h>

#include "common.h"
#include "messages.cuh"
#include "../../include/spg.h"

int spg_pedersen_chain_device(spg_ctx* ctx, const uint64_t* elems, int chain_len, uint64_t* out, uint8_t* status, size_t n,
                              uint64_t* out_y = nullptr, const uint64_t* second = nullptr);
int spg_ecdsa_verify_device(spg_ctx* ctx, const uint64_t* msg, const uint64_t* r, const uint64_t* s, const uint64_t* px,
                            const uint64_t* py, uint8_t* status, size_t n);

#define SPG_LIMIT_ORDER_WITH_FEES 3ull   // perpetual_messages.py:8

#define put_bits spg_put_bits

__global__ void __launch_bounds__(128) k_pack_limit_orders(spg_limit_orders o, uint64_t* __restrict__ elems,
                                                           uint8_t* __restrict__ status, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t *syn = o.asset_id_synthetic + 4 * i, *col = o.asset_id_collateral + 4 * i, *fee = o.asset_id_fee + 4 * i;
  uint8_t st = 0;
  // perpetual_messages.py:226-230: synthetic id below 2^128, collateral and fee ids below 2^250 (the 64- and 32-bit
  // fields cannot leave their ranges by construction of this ABI)
  if (syn[2] | syn[3]) st = 1;
  if ((col[3] >> 58) | (fee[3] >> 58)) st = 1;
  const bool buying = o.is_buying_synthetic[i] != 0;
  const uint64_t* sell = buying ? col : syn;
  const uint64_t* buy = buying ? syn : col;
  const uint64_t a_sell = buying ? o.amount_collateral[i] : o.amount_synthetic[i];
  const uint64_t a_buy = buying ? o.amount_synthetic[i] : o.amount_collateral[i];
  uint64_t* e = elems + i * 20;
#pragma unroll
  for (int k = 0; k < 4; k++) { e[k] = sell[k]; e[4 + k] = buy[k]; e[8 + k] = fee[k]; }
  uint64_t p0[4] = {0, 0, 0, 0}, p1[4] = {0, 0, 0, 0};
  put_bits(p0, (uint64_t)o.nonce[i], 0);
  put_bits(p0, o.max_amount_fee[i], 32);
  put_bits(p0, a_buy, 96);
  put_bits(p0, a_sell, 160);
  const uint64_t pos = o.position_id[i];
  put_bits(p1, (uint64_t)o.expiration_timestamp[i], 17);
  put_bits(p1, pos, 49);
  put_bits(p1, pos, 113);
  put_bits(p1, pos, 177);
  put_bits(p1, SPG_LIMIT_ORDER_WITH_FEES, 241);
#pragma unroll
  for (int k = 0; k < 4; k++) { e[12 + k] = p0[k]; e[16 + k] = p1[k]; }
  status[i] = st;
}

// merge: pack status (1 = the reference raises) and chain status (1 / 2 = the reference raises) into
// verify status 2; otherwise keep the ECDSA status
__global__ void k_merge_order_status(const uint8_t* __restrict__ pack_st, const uint8_t* __restrict__ chain_st,
                                     uint8_t* __restrict__ st, size_t n, int verify) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t p = pack_st[i], c = chain_st[i];
  if (verify) { if (p | c) st[i] = 2; }
  else st[i] = p ? 1 : c;
}

struct OrderStage {
  DevBuf b[10], elems, pack_st, chain_st, msg;
  spg_limit_orders d;
};

static int stage_orders(spg_ctx* ctx, const spg_limit_orders* o, size_t n, int flags, OrderStage& S) {
  SPG_ARG(o->asset_id_synthetic && o->asset_id_collateral && o->asset_id_fee && o->is_buying_synthetic && o->amount_synthetic &&
          o->amount_collateral && o->max_amount_fee && o->position_id && o->nonce && o->expiration_timestamp,
          "spg_limit_orders: null field array");
  S.d = *o;
  if (!(flags & SPG_DEVICE_PTRS)) {
    const void* src[10] = {o->asset_id_synthetic, o->asset_id_collateral, o->asset_id_fee, o->is_buying_synthetic,
                           o->amount_synthetic, o->amount_collateral, o->max_amount_fee, o->position_id, o->nonce,
                           o->expiration_timestamp};
    const size_t width[10] = {32, 32, 32, 1, 8, 8, 8, 8, 4, 4};
    for (int k = 0; k < 10; k++) {
      SPG_CUDA(S.b[k].alloc(ctx, n * width[k]));
      SPG_CUDA(cudaMemcpyAsync(S.b[k].p, src[k], n * width[k], cudaMemcpyHostToDevice, ctx->stream));
    }
    S.d.asset_id_synthetic = S.b[0].as<uint64_t>(); S.d.asset_id_collateral = S.b[1].as<uint64_t>();
    S.d.asset_id_fee = S.b[2].as<uint64_t>(); S.d.is_buying_synthetic = S.b[3].as<uint8_t>();
    S.d.amount_synthetic = S.b[4].as<uint64_t>(); S.d.amount_collateral = S.b[5].as<uint64_t>();
    S.d.max_amount_fee = S.b[6].as<uint64_t>(); S.d.position_id = S.b[7].as<uint64_t>();
    S.d.nonce = S.b[8].as<uint32_t>(); S.d.expiration_timestamp = S.b[9].as<uint32_t>();
  }
  SPG_CUDA(S.elems.alloc(ctx, n * 5 * 32)); SPG_CUDA(S.pack_st.alloc(ctx, n)); SPG_CUDA(S.chain_st.alloc(ctx, n));
  return SPG_OK;
}

extern "C" int spg_limit_order_msg_batch(spg_ctx* ctx, const spg_limit_orders* orders, uint64_t* msg_out, uint8_t* status,
                                         size_t n, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && orders && msg_out && status, "spg_limit_order_msg_batch: null");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return SPG_OK;
  OrderStage S;
  int rc = stage_orders(ctx, orders, n, flags, S);
  if (rc) return rc;
  uint64_t* dmsg = msg_out; uint8_t* dst = status;
  DevBuf bs;
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(S.msg.alloc(ctx, n * 32)); SPG_CUDA(bs.alloc(ctx, n));
    dmsg = S.msg.as<uint64_t>(); dst = bs.as<uint8_t>();
  }
  const unsigned blocks = (unsigned)((n + 127) / 128);
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  k_pack_limit_orders<<<blocks, 128, 0, ctx->stream>>>(S.d, S.elems.as<uint64_t>(), S.pack_st.as<uint8_t>(), n);
  SPG_LAUNCH_CHECK();
  if ((rc = spg_pedersen_chain_device(ctx, S.elems.as<uint64_t>(), 5, dmsg, S.chain_st.as<uint8_t>(), n))) return rc;
  k_merge_order_status<<<blocks, 128, 0, ctx->stream>>>(S.pack_st.as<uint8_t>(), S.chain_st.as<uint8_t>(), dst, n, 0);
  SPG_LAUNCH_CHECK();
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(cudaMemcpyAsync(msg_out, dmsg, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(status, dst, n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}

extern "C" int spg_limit_order_verify_batch(spg_ctx* ctx, const spg_limit_orders* orders, const uint64_t* r, const uint64_t* s,
                                            const uint64_t* pub_x, uint8_t* status, size_t n, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && orders && r && s && pub_x && status, "spg_limit_order_verify_batch: null");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return SPG_OK;
  OrderStage S;
  int rc = stage_orders(ctx, orders, n, flags, S);
  if (rc) return rc;
  SPG_CUDA(S.msg.alloc(ctx, n * 32));
  const uint64_t *dr = r, *ds = s, *dx = pub_x;
  uint8_t* dst = status;
  DevBuf br, bss, bx, bst;
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(br.alloc(ctx, n * 32)); SPG_CUDA(bss.alloc(ctx, n * 32)); SPG_CUDA(bx.alloc(ctx, n * 32)); SPG_CUDA(bst.alloc(ctx, n));
    SPG_CUDA(cudaMemcpyAsync(br.p, r, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(bss.p, s, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(bx.p, pub_x, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    dr = br.as<uint64_t>(); ds = bss.as<uint64_t>(); dx = bx.as<uint64_t>(); dst = bst.as<uint8_t>();
  }
  const unsigned blocks = (unsigned)((n + 127) / 128);
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  k_pack_limit_orders<<<blocks, 128, 0, ctx->stream>>>(S.d, S.elems.as<uint64_t>(), S.pack_st.as<uint8_t>(), n);
  SPG_LAUNCH_CHECK();
  if ((rc = spg_pedersen_chain_device(ctx, S.elems.as<uint64_t>(), 5, S.msg.as<uint64_t>(), S.chain_st.as<uint8_t>(), n))) return rc;
  if ((rc = spg_ecdsa_verify_device(ctx, S.msg.as<uint64_t>(), dr, ds, dx, nullptr, dst, n))) return rc;
  k_merge_order_status<<<blocks, 128, 0, ctx->stream>>>(S.pack_st.as<uint8_t>(), S.chain_st.as<uint8_t>(), dst, n, 1);
  SPG_LAUNCH_CHECK();
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!(flags & SPG_DEVICE_PTRS)) SPG_CUDA(cudaMemcpyAsync(status, dst, n, cudaMemcpyDeviceToHost, ctx->stream));
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}

// ------------------------------------------------------------------ the other perpetual messages (messages.cuh)
struct MsgFieldsDev {
  const uint64_t* felts[SPG_MSG_MAX_FELTS];
  const uint64_t* ints[SPG_MSG_MAX_INTS];
};

__global__ void __launch_bounds__(128) k_pack_messages(int kind, MsgFieldsDev f, uint64_t* __restrict__ elems,
                                                       uint8_t* __restrict__ status, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int nf = spg_msg_n_felts(kind), ni = spg_msg_n_ints(kind), len = spg_msg_chain_len(kind);
  const uint64_t* fp[SPG_MSG_MAX_FELTS];
  uint64_t iv[SPG_MSG_MAX_INTS];
#pragma unroll
  for (int k = 0; k < SPG_MSG_MAX_FELTS; k++) fp[k] = k < nf ? f.felts[k] + 4 * i : nullptr;
#pragma unroll
  for (int k = 0; k < SPG_MSG_MAX_INTS; k++) iv[k] = k < ni ? f.ints[k][i] : 0;
  status[i] = (uint8_t)spg_pack_message(kind, fp, iv, elems + i * (size_t)len * 4);
}

extern "C" int spg_message_hash_batch(spg_ctx* ctx, int kind, const spg_message_fields* fields, uint64_t* msg_out,
                                      uint8_t* status, size_t n, int flags) {
  SPG_LOCK(ctx);
  SPG_ARG(ctx && fields && msg_out && status, "spg_message_hash_batch: null");
  const int len = spg_msg_chain_len(kind);
  SPG_ARG(len != 0, "spg_message_hash_batch: unknown message kind");
  const int nf = spg_msg_n_felts(kind), ni = spg_msg_n_ints(kind);
  for (int k = 0; k < nf; k++) SPG_ARG(fields->felts[k], "spg_message_hash_batch: null felt array");
  for (int k = 0; k < ni; k++) SPG_ARG(fields->ints[k], "spg_message_hash_batch: null integer array");
  SPG_CUDA(cudaSetDevice(ctx->device));
  if (n == 0) return SPG_OK;
  MsgFieldsDev d;
  for (int k = 0; k < SPG_MSG_MAX_FELTS; k++) d.felts[k] = k < nf ? fields->felts[k] : nullptr;
  for (int k = 0; k < SPG_MSG_MAX_INTS; k++) d.ints[k] = k < ni ? fields->ints[k] : nullptr;
  DevBuf bf[SPG_MSG_MAX_FELTS], bi[SPG_MSG_MAX_INTS], elems, pack_st, chain_st, bmsg, bst;
  uint64_t* dmsg = msg_out; uint8_t* dst = status;
  if (!(flags & SPG_DEVICE_PTRS)) {
    for (int k = 0; k < nf; k++) {
      SPG_CUDA(bf[k].alloc(ctx, n * 32));
      SPG_CUDA(cudaMemcpyAsync(bf[k].p, fields->felts[k], n * 32, cudaMemcpyHostToDevice, ctx->stream));
      d.felts[k] = bf[k].as<uint64_t>();
    }
    for (int k = 0; k < ni; k++) {
      SPG_CUDA(bi[k].alloc(ctx, n * 8));
      SPG_CUDA(cudaMemcpyAsync(bi[k].p, fields->ints[k], n * 8, cudaMemcpyHostToDevice, ctx->stream));
      d.ints[k] = bi[k].as<uint64_t>();
    }
    SPG_CUDA(bmsg.alloc(ctx, n * 32)); SPG_CUDA(bst.alloc(ctx, n));
    dmsg = bmsg.as<uint64_t>(); dst = bst.as<uint8_t>();
  }
  SPG_CUDA(elems.alloc(ctx, n * (size_t)len * 32)); SPG_CUDA(pack_st.alloc(ctx, n)); SPG_CUDA(chain_st.alloc(ctx, n));
  const unsigned blocks = (unsigned)((n + 127) / 128);
  SPG_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  k_pack_messages<<<blocks, 128, 0, ctx->stream>>>(kind, d, elems.as<uint64_t>(), pack_st.as<uint8_t>(), n);
  SPG_LAUNCH_CHECK();
  int rc;
  if ((rc = spg_pedersen_chain_device(ctx, elems.as<uint64_t>(), len, dmsg, chain_st.as<uint8_t>(), n))) return rc;
  k_merge_order_status<<<blocks, 128, 0, ctx->stream>>>(pack_st.as<uint8_t>(), chain_st.as<uint8_t>(), dst, n, 0);
  SPG_LAUNCH_CHECK();
  SPG_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  if (!(flags & SPG_DEVICE_PTRS)) {
    SPG_CUDA(cudaMemcpyAsync(msg_out, dmsg, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    SPG_CUDA(cudaMemcpyAsync(status, dst, n, cudaMemcpyDeviceToHost, ctx->stream));
  }
  SPG_CUDA(cudaStreamSynchronize(ctx->stream));
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1); ctx->last_ms = ms;
  return SPG_OK;
}
