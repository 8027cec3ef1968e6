/* libspg -- C-ABI of the B200-native hot path behind StarkEx Perpetual.
 *
 * Every entry point names the reference interface it stands in for (paths relative to the
 * reference repo starkware-libs/stark-perpetual).  Conventions:
 *   - field elements ("felts") cross the boundary as 4 x uint64 little-endian limbs holding the
 *     CANONICAL value in [0, p), p = 2^251 + 17*2^192 + 1; the *_be32 variants use the
 *     reference's own byte ABI, 32-byte big-endian (fast_pedersen_hash.py:47-52,
 *     starkware/python/utils.py:414-451);
 *   - buffers are caller-owned and contiguous; they are host memory unless SPG_DEVICE_PTRS is set
 *     in `flags`, in which case they are device pointers on the context's GPU;
 *   - every function returns 0 or a negative SPG_E* code and never throws; spg_last_error()
 *     gives the message.  There is NO CPU fallback: without a CUDA device spg_create fails.
 *   - threading: every entry point locks its context for the duration of the call, so a context may be shared by
 *     threads (calls queue up); distinct contexts run concurrently.  spg_last_error / spg_last_kernel_ms /
 *     spg_stage_ms report the most recent call on the context, whichever thread made it.
 *   - per-element outcomes of batched calls are reported in `status` arrays (one byte each) so
 *     the Python layer can raise exactly the exception the reference raises.
 */
#ifndef SPG_H
#define SPG_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct spg_ctx spg_ctx;

#define SPG_OK 0
#define SPG_E_CUDA (-1)      /* CUDA runtime error, see spg_last_error */
#define SPG_E_ARG (-2)       /* bad argument */
#define SPG_E_NOMEM (-3)
#define SPG_E_NODEVICE (-4)
#define SPG_E_PROOF (-5)     /* prover could not build a proof (e.g. trace does not satisfy the AIR) */

#define SPG_DEVICE_PTRS 1    /* flags: buffers are device pointers */
#define SPG_MONT_OUT 4       /* flags (spg_lde_coeffs): additionally multiply by R = 2^256, so that a canonical trace
                                yields Montgomery-form coefficients / evaluations (what the prover stages consume) */
#define SPG_NO_SYNC 2        /* flags (with SPG_DEVICE_PTRS): enqueue on the context stream and return; the
                                caller synchronises (spg_synchronize / its own events) */

/* NTT orderings */
#define SPG_NTT_NAT_TO_REV 0 /* natural order in, bit-reversed out (DIF) */
#define SPG_NTT_REV_TO_NAT 1 /* bit-reversed in, natural out (DIT) */
#define SPG_NTT_NAT_TO_NAT 2 /* natural in, natural out (DIF + permutation) */

int spg_create(int device_ordinal, spg_ctx** out);
void spg_destroy(spg_ctx* ctx);
const char* spg_last_error(spg_ctx* ctx);
int spg_device_count(void);
/* device milliseconds (CUDA events on the context stream) spent in kernels by the last call */
double spg_last_kernel_ms(spg_ctx* ctx);
/* number of kernel launches issued by this context since creation */
uint64_t spg_launch_count(spg_ctx* ctx);
int spg_synchronize(spg_ctx* ctx);
/* run all subsequent work of this context on an existing CUDA stream (a cudaStream_t, e.g. the host
 * framework's current stream) instead of the context's own one */
int spg_set_stream(spg_ctx* ctx, void* cuda_stream);
/* device milliseconds of stage `stage` of the last pipeline call (stage ids are listed per entry point) */
double spg_stage_ms(spg_ctx* ctx, int stage);

/* ---- field layer (a1/a2 of SURVEY section 8; signature.py:41-42, math_utils.py:50-56) ----------------- */
/* out[i] = op(a[i], b[i]); op: 0 mul, 1 add, 2 sub, 3 inverse of a (b ignored), 4 a^b[0..3] */
int spg_field_op(spg_ctx* ctx, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n,
                 int flags);
/* throughput probe: every thread runs `iters` dependent Montgomery multiplications on `chains`
 * independent accumulators; returns field multiplications per second in *mul_per_s, and the issue rate of
 * IMAD.WIDE (carry-chained rows, the multiplication's own pattern) in *imad_wide_per_s. */
int spg_bench_field_mul(spg_ctx* ctx, int iters, int chains, double* mul_per_s, double* imad_wide_per_s);

/* ---- Pedersen hash (BASELINE.json configs[0]) ------------------------------------
 * Replaces signature.py:296-318 pedersen_hash(*elements) and fast_pedersen_hash.py:34-52.
 * status[i]: 0 ok; 1 an input is >= p (the reference raises AssertionError, signature.py:307);
 *            2 "Unhashable input." (signature.py:313).  out[i] = 0 when status[i] != 0. */
int spg_pedersen_hash2_batch(spg_ctx* ctx, const uint64_t* x, const uint64_t* y, uint64_t* out, uint8_t* status,
                             size_t n, int flags);
/* the reference's byte ABI: 32-byte big-endian in and out (fast_pedersen_hash.py:47-52); host pointers */
int spg_pedersen_hash2_batch_be32(spg_ctx* ctx, const uint8_t* x, const uint8_t* y, uint8_t* out, uint8_t* status,
                                  size_t n);
/* n hash chains of chain_len elements each, elems[n][chain_len]:  h = H(e0, e1); h = H(h, e2); ...
 * (chain_len = 1 hashes a single element, like pedersen_hash(x)).  This is the shape of the message hashes
 * of perpetual_messages.py:253-286 and of the program-hash chain (program_hash_test_utils.py:7-21). */
int spg_pedersen_chain_batch(spg_ctx* ctx, const uint64_t* elems, size_t chain_len, uint64_t* out, uint8_t* status,
                             size_t n, int flags);

/* signature.py:300-318 pedersen_hash_as_point: both coordinates of the hash point.  elems: [n][n_elems] canonical felts,
 * n_elems = 1 or 2 (host pointers); status as spg_pedersen_hash2_batch. */
int spg_pedersen_hash_point_batch(spg_ctx* ctx, const uint64_t* elems, size_t n_elems, uint64_t* out_x, uint64_t* out_y,
                                  uint8_t* status, size_t n, int flags);

/* Merkle tree whose node function is pedersen_hash(left, right): the StarkEx state trees (positions / orders) that
 * src/services/perpetual/cairo/state/state.cairo:155-173 updates with merkle_multi_update, and for which
 * src/starkware/python/merkle_tree.py:4-44 builds the update hints.  leaves: [n_leaves][4] canonical felts, n_leaves a
 * power of two >= 2.  root_out: [4]; nodes_out (optional): the n_leaves - 1 internal nodes, level by level from the
 * bottom (n/2, n/4, ..., 1).  *status_out (host): 0 ok, 1 a leaf >= p, 2 "Unhashable input." somewhere in the tree. */
int spg_pedersen_merkle_tree(spg_ctx* ctx, const uint64_t* leaves, size_t n_leaves, uint64_t* root_out, uint64_t* nodes_out,
                             uint8_t* status_out, int flags);

/* ---- STARK-curve ECDSA (BASELINE.json configs[4]) ---------------------------------
 * Replaces signature.py:217-260 verify(msg_hash, r, s, public_key).  All operands are 256-bit values as 4 x u64
 * LE limbs.  pub_y_or_null = NULL: x-only keys (signature.py:229-238: y = smaller root, then -y).
 * status[i]: 1 valid; 0 invalid (the reference returns False, incl. every assertion inside the three
 * mimic_ec_mult_air calls, signature.py:251-257, and keys that are not on the curve as x-only keys);
 *            2 a precondition is violated and the reference RAISES (signature.py:219, :225-227, :241), or a key
 *              coordinate is not a field element. */
int spg_ecdsa_verify_batch(spg_ctx* ctx, const uint64_t* msg, const uint64_t* r, const uint64_t* s,
                           const uint64_t* pub_x, const uint64_t* pub_y_or_null, uint8_t* status, size_t n, int flags);
/* signature.py:99-110 private_key_to_ec_point_on_stark_curve / private_to_stark_key: (pub_x[i], pub_y[i]) =
 * priv[i] * G (pub_y_or_null may be NULL); status 1 if priv is outside (0, n). */
int spg_private_to_stark_key_batch(spg_ctx* ctx, const uint64_t* priv, uint64_t* pub_x, uint64_t* pub_y_or_null,
                                   uint8_t* status, size_t n, int flags);
/* signature.py:137-173 sign(msg_hash, priv_key, seed): deterministic signatures, nonce by RFC 6979 (HMAC-SHA256) exactly
 * as signature.py:117-134 derives it through python-ecdsa 0.17.0, the three rejection rules (:158-170) retried on the
 * device with the bumped seed.  seed_or_null[i] is the reference's `seed` argument (NULL or 0 = None: both give no
 * extra entropy on the first attempt and 1 on the second).  (r[i], s[i]) canonical.
 * status[i]: 0 ok; 1 "Message not signable" (msg >= 2^251, the reference's AssertionError :141); 2 private key outside
 * [1, n) (outside the reference's domain, is_valid_stark_private_key); 3 no signature after 128 attempts (never). */
int spg_sign_batch(spg_ctx* ctx, const uint64_t* msg, const uint64_t* priv, const uint64_t* seed_or_null, uint64_t* r,
                   uint64_t* s, uint8_t* status, size_t n, int flags);

/* signature.py:84-96 get_y_coordinate: y[i] = the smaller square root of x^3 + x + beta (math_utils.py:43-47).
 * status[i]: 0 ok; 1 InvalidPublicKeyError (no point with this x; is_valid_stark_key, signature.py:204-214, is
 * status == 0); 2 x is not a field element. */
int spg_get_y_coordinate_batch(spg_ctx* ctx, const uint64_t* x, uint64_t* y, uint8_t* status, size_t n, int flags);
/* signature.py:176-190 mimic_ec_mult_air: out = m * point + shift_point computed with the AIR's steps.  m: [n][4];
 * point_xy, shift_xy, out_xy: [n][8] (x then y, canonical).  status[i]: 0 ok; 1 the reference raises AssertionError
 * (m outside (0, 2^251), partial_sum.x == point.x on some step, doubling of a point with y = 0); 2 a coordinate is not
 * a field element. */
int spg_mimic_ec_mult_air_batch(spg_ctx* ctx, const uint64_t* m, const uint64_t* point_xy, const uint64_t* shift_xy,
                                uint64_t* out_xy, uint8_t* status, size_t n, int flags);

/* ---- math_utils.py as batched device operations -----------------------------------
 * Replaces src/starkware/crypto/signature/math_utils.py:59-68 ec_add (op 0), :79-88 ec_double (op 1, alpha = 1) and
 * :91-100 ec_mult (op 2) over the STARK prime.  a_xy, out_xy: [n][8] canonical (x, y); b: [n][8] second point (op 0),
 * unused (op 1), [n][4] scalar m (op 2).  ec_mult follows the reference's recursion: every doubling 2^k P up to the top
 * bit of m first, then the additions from the highest set bit down, so the inputs on which an assertion fires are the
 * reference's own.  status[i]: 0 ok; 1 the reference raises AssertionError (x1 == x2 in ec_add, :64; y == 0 in
 * ec_double, :84); 2 a coordinate is not a field element; 3 m == 0 (the reference recurses forever).  Host pointers. */
int spg_ec_op_batch(spg_ctx* ctx, int op, const uint64_t* a_xy, const uint64_t* b, uint64_t* out_xy, uint8_t* status,
                    size_t n, int flags);
/* math_utils.py:36-47 is_quad_residue / sqrt_mod over the STARK prime: y[i] = the SMALLER square root of a[i].
 * status[i]: 0 ok (a is a quadratic residue, 0 included), 1 non-residue, 2 a >= p. */
int spg_field_sqrt_batch(spg_ctx* ctx, const uint64_t* a, uint64_t* y, uint8_t* status, size_t n, int flags);

/* ---- perpetual limit orders (BASELINE.json configs[4]) ------------------------------
 * Replaces src/services/perpetual/public/perpetual_messages.py:212-286 get_limit_order_msg: field packing and the 4-deep
 * Pedersen chain run on the device (Cairo twin: src/services/exchange/cairo/signature_message_hashes.cairo:56-91).
 * Struct of arrays, n entries each; asset ids are canonical felts ([n][4] u64), the other fields have exactly the
 * widths the reference asserts (perpetual_messages.py:231-236), so only the three id bounds (:226-230) can fail. */
typedef struct spg_limit_orders {
  const uint64_t* asset_id_synthetic;    /* [n][4], must be < 2^128 */
  const uint64_t* asset_id_collateral;   /* [n][4], must be < 2^250 */
  const uint64_t* asset_id_fee;          /* [n][4], must be < 2^250 */
  const uint8_t* is_buying_synthetic;    /* [n] */
  const uint64_t* amount_synthetic;      /* [n] */
  const uint64_t* amount_collateral;     /* [n] */
  const uint64_t* max_amount_fee;        /* [n] */
  const uint64_t* position_id;           /* [n] */
  const uint32_t* nonce;                 /* [n] */
  const uint32_t* expiration_timestamp;  /* [n] */
} spg_limit_orders;
/* msg_out[i] = get_limit_order_msg(order i).  status[i]: 0 ok; 1 a bound of perpetual_messages.py:226-230 is violated
 * (the reference raises AssertionError); 2 "Unhashable input." (signature.py:313). */
int spg_limit_order_msg_batch(spg_ctx* ctx, const spg_limit_orders* orders, uint64_t* msg_out, uint8_t* status, size_t n,
                              int flags);
/* verify(get_limit_order_msg(order i), r[i], s[i], pub_x[i]) with x-only keys, the check
 * src/services/perpetual/cairo/order/order.cairo:132-166 performs per order.  status as spg_ecdsa_verify_batch
 * (1 valid, 0 invalid, 2 the reference raises -- including a failed order bound or an unhashable message). */
int spg_limit_order_verify_batch(spg_ctx* ctx, const spg_limit_orders* orders, const uint64_t* r, const uint64_t* s,
                                 const uint64_t* pub_x, uint8_t* status, size_t n, int flags);

/* The other perpetual messages: field packing and the Pedersen chain on the device.
 *   kind 4  get_transfer_msg              (perpetual_messages.py:97-162)
 *           felts: asset_id, asset_id_fee, receiver_public_key
 *           ints:  sender_position_id, receiver_position_id, src_fee_position_id, nonce, amount, max_amount_fee,
 *                  expiration_timestamp
 *   kind 5  get_conditional_transfer_msg  (:24-94)   felts: asset_id, asset_id_fee, receiver_public_key, condition;
 *           ints as kind 4
 *   kind 7  get_withdrawal_to_address_msg (:165-209) felts: asset_id_collateral, eth_address (as an integer);
 *           ints:  position_id, nonce, amount, expiration_timestamp
 *   kind 100 get_price_msg                (:311-326) felts: asset_pair, price;  ints: oracle_name, timestamp
 * felts[k]: [n][4] u64 canonical values; ints[k]: [n] u64 (narrower reference fields are range-checked).
 * msg_out[i]: the message hash.  status[i]: 0 ok; 1 a bound the reference asserts (:38-48, :110-119, :174-179, :314-317)
 * is violated; 2 "Unhashable input." (signature.py:313). */
#define SPG_MSG_TRANSFER 4
#define SPG_MSG_CONDITIONAL_TRANSFER 5
#define SPG_MSG_WITHDRAWAL_TO_ADDRESS 7
#define SPG_MSG_PRICE 100
typedef struct spg_message_fields {
  const uint64_t* felts[4];
  const uint64_t* ints[7];
} spg_message_fields;
int spg_message_hash_batch(spg_ctx* ctx, int kind, const spg_message_fields* fields, uint64_t* msg_out, uint8_t* status,
                           size_t n, int flags);

/* ---- state-tree work of a perpetual batch --------------------------------------------------
 * Position hashing: replaces src/services/perpetual/cairo/position/hash.cairo:22-74 (position_hash_assets, position_hash),
 * what state.cairo:143-150 (hash_position_updates) does for every touched position before the Merkle update.
 *   h = 0;  for each asset: h = H(h, (asset_id * 2^64 + (cached_funding_index + 2^63)) * 2^64 + (balance + 2^63));
 *   h = H(h, public_key);   h = H(h, (collateral_balance + 2^63) * 2^16 + n_assets)          (constants.cairo:10-38)
 * Positions in CSR form: the assets of position i are entries [asset_offsets[i], asset_offsets[i+1]) of the asset arrays.
 * status[i]: 0 ok; 1 a bound is violated (asset_id >= 2^120, n_assets >= 2^16, public_key >= p); 2 "Unhashable input." */
typedef struct spg_positions {
  const uint64_t* public_key;            /* [n][4] canonical felt */
  const int64_t* collateral_balance;     /* [n] */
  const uint64_t* asset_offsets;         /* [n + 1], non-decreasing, asset_offsets[0] = 0 */
  const uint64_t* asset_id;              /* [total][2] little-endian 128-bit words, value < 2^120 */
  const int64_t* balance;                /* [total] */
  const int64_t* cached_funding_index;   /* [total] */
} spg_positions;
int spg_position_hash_batch(spg_ctx* ctx, const spg_positions* positions, uint64_t* hash_out, uint8_t* status, size_t n, int flags);

/* Sparse Merkle multi-update with pedersen_hash nodes: the computation merkle_multi_update performs for the positions and
 * orders trees (state.cairo:151-173), height <= 64.  keys: n strictly increasing leaf indices (a squashed dict);
 * prev_leaves / new_leaves: [n][4].  The update tree is the one src/starkware/python/merkle_tree.py:4-29 build_update_tree
 * describes; wherever a node of it has only one child in the tree, the other child is a SIBLING -- the hash of an untouched
 * subtree, which the Cairo hints read from the preimage dictionary and the caller supplies here, in this order: level by
 * level from the leaves up (level l = children at height l above the leaves), ascending node index within a level.
 * spg_merkle_update_siblings lists (level, node index) of every sibling in that order (null outputs: count only).
 * Outputs: prev_root (to be compared with the state's current root, as merkle_multi_update asserts) and new_root;
 * nodes_out (optional): every node of the update tree above the leaves, level by level: for a level with m nodes, m
 * previous values then m new values ([2][m][4]; per-level counts from spg_merkle_update_node_count, [height] entries).
 * *status_out: 0 ok, 1 an input is >= p, 2 "Unhashable input." somewhere.  Host pointers. */
int spg_merkle_update_siblings(spg_ctx* ctx, unsigned height, const uint64_t* keys, size_t n, uint8_t* level_out,
                               uint64_t* index_out, size_t cap, size_t* count_out);
int spg_merkle_update_node_count(spg_ctx* ctx, unsigned height, const uint64_t* keys, size_t n, size_t* level_counts_out);
int spg_merkle_multi_update(spg_ctx* ctx, unsigned height, const uint64_t* keys, const uint64_t* prev_leaves,
                            const uint64_t* new_leaves, size_t n, const uint64_t* siblings, size_t n_siblings,
                            uint64_t* prev_root_out, uint64_t* new_root_out, uint64_t* nodes_out, uint8_t* status_out, int flags);

/* Right-folded Pedersen hash chain  h(d[0], h(d[1], ... h(d[len-2], d[len-1])))  of n chains of `len` canonical felts each
 * (data: [n][len][4]): compute_hash_chain of cairo-lang (un-vendored dependency; the reference reaches it through
 * compute_program_hash_chain at src/starkware/cairo/bootloaders/program_hash_test_utils.py:7-9, golden value
 * src/services/perpetual/cairo/program_hash.json:2).  len = 1 returns the element.  status[i]: 0 ok, 1 an element >= p,
 * 2 "Unhashable input.".  Sequential by construction: one device thread per chain.  Host pointers. */
int spg_hash_chain_rfold_batch(spg_ctx* ctx, const uint64_t* data, size_t len, uint64_t* out, uint8_t* status, size_t n, int flags);

/* ---- NTT over the STARK prime (no reference symbol, field from signature.py:41-42) */
/* In-place transform of `batch` vectors of 2^log_n felts stored back to back.  omega = 3^((p-1)/2^log_n).
 * inverse != 0 uses omega^-1 and scales by 2^-log_n. */
int spg_ntt(spg_ctx* ctx, uint64_t* data, unsigned log_n, size_t batch, int inverse, int order, int flags);

/* ---- LDE (BASELINE.json configs[2]; no reference symbol) ------------------------------
 * trace: n_cols columns of 2^log_n felts, column-major [n_cols][N], values on <w_N> in natural order.
 * out:   [B][n_cols][N], B = 2^log_blowup:  out[j][c][i] = f_c(g * w_{BN}^j * w_N^i), f_c the degree < N
 *        interpolant of column c, g = *coset_offset (canonical felt, HOST pointer; NULL = 3 = FIELD_GEN,
 *        signature.py:42).  Stages for spg_stage_ms: 0 interpolation (inverse NTT), 1 coset evaluations. */
int spg_lde(spg_ctx* ctx, const uint64_t* trace, unsigned log_n, size_t n_cols, unsigned log_blowup,
            const uint64_t* coset_offset, uint64_t* out, int flags);
/* The two phases of spg_lde separately, for the multi-GPU split (device pointers only):
 *   spg_lde_coeffs: columns -> coefficient columns scaled by g^k, bit-reversed order (column-sharded phase);
 *   spg_lde_cosets: (all-gathered) coefficient columns -> cosets [coset_begin, coset_begin + coset_count),
 *                   out[j - coset_begin][c][i]  (coset-sharded phase). */
int spg_lde_coeffs(spg_ctx* ctx, const uint64_t* trace, unsigned log_n, size_t n_cols,
                   const uint64_t* coset_offset, uint64_t* coeffs, int flags);
int spg_lde_cosets(spg_ctx* ctx, const uint64_t* coeffs, unsigned log_n, size_t n_cols, unsigned log_blowup,
                   size_t coset_begin, size_t coset_count, uint64_t* out, int flags);

/* ---- Merkle commitment, BLAKE2s (no reference symbol) -----------------------------------
 * table: [8][n_cols][rows] 256-bit values (hashed as given, 32 bytes big-endian each).  Leaf j*rows/8 + i'
 * = BLAKE2s of the rows i' + k*rows/8 (k = 0..7) of coset j, all columns of a row in order; node =
 * BLAKE2s(left || right).  root32 receives the root; tree_out (optional) the whole tree: `rows` leaf
 * digests, then rows/2, ..., 1 node digests (2*rows - 1 digests of 32 bytes). */
int spg_merkle_commit(spg_ctx* ctx, const uint64_t* table, size_t n_cols, size_t rows, uint8_t* root32,
                      uint8_t* tree_out, int flags);

/* ---- Pedersen hash-chain AIR: witness, constraint evaluation, proof -----------
 * The statement (DESIGN.md "AIR"): 5 lanes of Pedersen hashes (signature.py:296-318), 512 trace rows per
 * hash, chained in segments of 2^chain_log hashes starting from the public seed x0[lane]; public output =
 * result of the last hash of every lane.  Trace = [25][2^log_n] canonical felts, columns (X, Y, S, M, I)
 * per lane.  spg_pedersen_chain_trace is the witness generator (the role cairo-run plays for the real
 * program, cairo_cmake_rules.cmake:94-110): x0 [5] and ys [5][2^log_n / 512] canonical felts. */
int spg_pedersen_chain_trace(spg_ctx* ctx, unsigned log_n, unsigned chain_log, const uint64_t* x0,
                             const uint64_t* ys, uint64_t* trace_out, int flags);
/* composition polynomial sum_k alpha^k C_k(x) / Z_k(x) of the trace on the LDE cosets j = 0, 2, 4, 6:
 * cp_out [4][2^log_n] canonical; x0, outs [5], alpha canonical (host pointers).  Stage 2 of spg_stage_ms. */
int spg_air_eval(spg_ctx* ctx, const uint64_t* trace, unsigned log_n, unsigned chain_log, const uint64_t* x0,
                 const uint64_t* outs, const uint64_t* alpha, uint64_t* cp_out, int flags);
/* Full proof.  trace: [25][2^log_n] canonical felts (host, or device with SPG_DEVICE_PTRS); x0 [5] canonical
 * (host).  Writes the proof bytes (format: DESIGN.md "Proof") to proof_out (capacity proof_cap) and their
 * number to *proof_len; with proof_out = NULL only the length is returned.  SPG_E_PROOF if the trace does
 * not satisfy the AIR.  Stages for spg_stage_ms: 0 trace LDE, 1 trace Merkle, 2 AIR/composition + chunk
 * split, 3 chunk LDE, 4 chunk Merkle, 5 out-of-domain evaluation, 6 DEEP quotient, 7 FRI, 8 query openings. */
int spg_prove(spg_ctx* ctx, const uint64_t* trace, unsigned log_n, unsigned chain_log, const uint64_t* x0,
              unsigned n_queries, uint8_t* proof_out, size_t proof_cap, size_t* proof_len, int flags);

/* ---- second AIR: the ECDSA builtin (csrc/air_ecdsa.cu) -------------------------------------------
 * The reference has no AIR; what it pins is the computation each 256-row block of this trace encodes, one `verify` call
 * (src/starkware/crypto/signature/signature.py:243-260: zG, rQ, w(zG + rQ) by mimic_ec_mult_air, :176-190, then
 * r == ec_add(wB, -shift).x), with every assertion of those loops held by an inverse cell.
 * Public input of the statement: the message hash and the key's x of every signature (the two cells the Cairo ECDSA builtin
 * exposes); r and w are witness.
 * spg_ecdsa_air_trace: msg, r, w, key_x, key_y = [2^log_n / 256] canonical felts each (w = s^-1 mod the curve order, as
 *   verify computes it at :219; the key as a curve point) -> trace_out [25][2^log_n] canonical.  SPG_E_ARG where the
 *   reference asserts (scalar outside [1, 2^251), key off the curve, x collision) or verify() is False.
 * spg_air_eval_ecdsa: composition polynomial of such a trace on the cosets 0, 2, 4, 6 -> cp_out [4][2^log_n] canonical
 *   (parity entry point; msgs, key_x = the public input, [2^log_n / 256] canonical each; alpha canonical; host pointers).
 * spg_prove_ecdsa: the protocol of spg_prove over this AIR; proof header VERSION = 2 followed by the public input
 *   (msg, key x per signature).  msgs / key_x are host arrays whatever `flags` says about the trace. */
int spg_ecdsa_air_trace(spg_ctx* ctx, unsigned log_n, const uint64_t* msg, const uint64_t* r, const uint64_t* w,
                        const uint64_t* key_x, const uint64_t* key_y, uint64_t* trace_out, int flags);
int spg_air_eval_ecdsa(spg_ctx* ctx, const uint64_t* trace, unsigned log_n, const uint64_t* msgs, const uint64_t* key_x,
                       const uint64_t* alpha, uint64_t* cp_out, int flags);
int spg_prove_ecdsa(spg_ctx* ctx, const uint64_t* trace, unsigned log_n, const uint64_t* msgs, const uint64_t* key_x,
                    unsigned n_queries, uint8_t* proof_out, size_t proof_cap, size_t* proof_len, int flags);

/* ---- multi-GPU prover: one process per GPU, NCCL called from inside libspg (DESIGN.md "Multi-GPU") ----------------------
 * spg_comm_unique_id: rank 0 obtains a 128-byte NCCL id and hands it to the other ranks by any means (the Python host side
 * broadcasts it over torch.distributed); spg_comm_init joins the communicator on the context's GPU (world 1, 2, 4 or 8;
 * world = 1 needs no id and no NCCL).  NCCL is bound at run time (the libnccl already in the process, or SPG_NCCL_LIB).
 * spg_prove_sharded: cols_local = THIS rank's trace columns, dealt cyclically -- columns rank, rank + world, ... of the 25 --
 * as [my_cols][2^log_n] canonical felts (host, or device with SPG_DEVICE_PTRS); x0, outs [5] canonical (host; outs = the
 * public outputs, column 5 l at the last row).  Collective: every rank must call it with the same arguments.  Every
 * rank receives the proof, byte-identical to spg_prove's.  Stages for spg_stage_ms as spg_prove (stage 0 includes the
 * column exchange). */
int spg_comm_unique_id(spg_ctx* ctx, uint8_t* id_out /*[128]*/);
int spg_comm_init(spg_ctx* ctx, int rank, int world, const uint8_t* id /*[128], may be NULL when world == 1*/);
int spg_prove_sharded(spg_ctx* ctx, const uint64_t* cols_local, unsigned log_n, unsigned chain_log, const uint64_t* x0,
                      const uint64_t* outs, unsigned n_queries, uint8_t* proof_out, size_t proof_cap, size_t* proof_len, int flags);
/* the same collective over the ECDSA-builtin AIR (cols_local: this rank's columns of the trace spg_ecdsa_air_trace writes;
 * msgs, key_x: the whole public input on every rank, host).  Byte-identical to spg_prove_ecdsa. */
int spg_prove_ecdsa_sharded(spg_ctx* ctx, const uint64_t* cols_local, unsigned log_n, const uint64_t* msgs, const uint64_t* key_x,
                            unsigned n_queries, uint8_t* proof_out, size_t proof_cap, size_t* proof_len, int flags);

/* ---- stage-level entry points for the multi-GPU host driver (device pointers; DESIGN.md "Multi-GPU") ----------
 * A GPU owns n_cosets consecutive cosets starting at first_coset; its tables hold exactly those cosets,
 * [n_cosets][n_cols][rows].  Scalars (alpha, z, gamma, beta, x0, outs, oods) are canonical felts in host memory.
 * Same kernels as spg_prove; with first_coset = 0, n_cosets = 8 they reproduce its stages one by one. */
int spg_stage_merkle(spg_ctx* ctx, const uint64_t* table, size_t n_cols, size_t rows, int n_cosets, uint8_t* tree_out);
int spg_stage_air(spg_ctx* ctx, unsigned log_n, unsigned chain_log, const uint64_t* t_lde, int first_coset, int jj0,
                  int n_even, const uint64_t* x0, const uint64_t* outs, const uint64_t* alpha, uint64_t* cp_out);
int spg_stage_cp_split(spg_ctx* ctx, unsigned log_n, const uint64_t* cp, int jj0, int n_even, uint64_t* hev);
int spg_stage_poly_eval(spg_ctx* ctx, unsigned log_n, const uint64_t* const* cols, const int* pt_idx, int n_items,
                        const uint64_t* pts, int n_pts, uint64_t* out);
int spg_stage_deep(spg_ctx* ctx, unsigned log_n, const uint64_t* t_lde, const uint64_t* h_lde, int first_coset,
                   int n_cosets, const uint64_t* z, const uint64_t* gamma, const uint64_t* oods, uint64_t* inv_scratch,
                   uint64_t* out);
int spg_stage_fri_fold(spg_ctx* ctx, const uint64_t* in, unsigned log_rows, int first_coset, int n_cosets,
                       const uint64_t* beta, int layer_index, uint64_t* out);
int spg_stage_open(spg_ctx* ctx, const uint64_t* table, size_t n_cols, size_t rows, int n_cosets, const uint8_t* tree,
                   const uint32_t* idx, int count, uint8_t* leaves_out, uint8_t* paths_out);
/* host-side checks the driver runs between stages (the same routines spg_prove uses): the composition identity at the
 * out-of-domain point (oods: [54][4] canonical), and the interpolation + low-degree check of the last FRI layer
 * (vals: [8][2^log_rows_last][4] raw Montgomery limbs as stored on the device, coset-major; coeffs_out: 32 bytes per
 * coefficient in the proof's serialisation).  Both return SPG_E_PROOF when the trace does not satisfy the AIR. */
int spg_stage_check_oods(spg_ctx* ctx, unsigned log_n, unsigned chain_log, const uint64_t* x0, const uint64_t* outs,
                         const uint64_t* alpha, const uint64_t* z, const uint64_t* oods);
int spg_stage_last_layer(spg_ctx* ctx, const uint64_t* vals, unsigned log_rows_last, int n_folds, uint8_t* coeffs_out);

#ifdef __cplusplus
}
#endif
#endif
