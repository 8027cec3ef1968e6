"""BASELINE.json configs[0], [1] and [4] measured on the GPU for the bench line's `aux` object (bench.py imports this after
its timed region; `python tools/aux_bench.py` prints the object on its own).

Per config: kernel milliseconds (CUDA events inside libspg), end-to-end milliseconds of the host-buffer C-ABI call,
algorithmic multiplications per unit (the kernels' own formulas, below), 252-bit multiplications per second and that as
a fraction of the integer-pipe ceiling (148 SMs x 32 IMAD.WIDE lanes/clk / 64 IMAD.WIDE per multiplication), a parity
check of a sample against the oracle, and -- for the two rows the reference implements (Pedersen hash, order
verification) -- the REFERENCE's own Python timed on the host cores (`cpu_reference.kind = "reference"`, one
oracle/ref_worker.py process per core, bounded sample; oracle/refenv.py).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

P = 2**251 + 17 * 2**192 + 1
EC_ORDER = 0x800000000000010FFFFFFFFFFFFFFFFB781126DCAE7B2321E66A241ADC64D2F

# ---- algorithmic multiplication counts (per thread, from the kernels' code: csrc/ec.cuh, ecdsa.cuh) ------------------
FERMAT_INV = 264                      # fp_inv_chain: 251 squarings + 13 multiplications
MADD = 8 + 3 + 2                      # ec_madd_nocheck (8M + 3S) + the cached Z^2, Z^3
JADD, JDBL = 12 + 4, 9                # ec_jadd_nocheck with the two Z^2 and u1, u2; ec_jdouble_nocheck
SQRT_MIN = 3 + 55 + 2 + sum(184 - 8 * i for i in range(24)) + 48    # fp_sqrt_min: a^q, 24 windows of squarings, table products


def pedersen_muls(popcount):
    """pedersen_hash2: 504 collision checks (one multiplication each), a mixed addition per set bit, one inversion"""
    return 504 + MADD * popcount + FERMAT_INV + 3


def verify_muls(pop_msg, pop_r, pop_w, roots_tried):
    """ecdsa_verify_one: s^-1 mod n (252 squarings + ~126 products mod n), the square root, and per candidate root
    mimic_ec_mult_air x3 (251 steps each) + 2 additions + the final comparison"""
    mimic_gen = 251 + MADD * pop_msg
    mimic_var = lambda pop: 251 * (4 + JDBL) + (JADD - 4) * pop     # noqa: E731
    point = mimic_gen + mimic_var(pop_r) + mimic_var(pop_w) + 2 * JADD + 3
    return 378 + 6 + SQRT_MIN + roots_tried * point


def int_ceiling(sm_mhz):
    return 148 * 32 * sm_mhz * 1e6 / 64.0


def _popcounts(limbs):
    a = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 4)
    return np.unpackbits(a.view(np.uint8), axis=1).sum(axis=1)


def _timed(fn, reps=5):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return out, best * 1e3


# ---- reference-timed CPU legs: one oracle/ref_worker.py process per host core, each importing the reference ------------
def reference_rates(pairs, order_cases, cores=None, timeout=600):
    """Times the reference's pedersen_hash on `pairs` and get_limit_order_msg + verify on `order_cases`, spread over all
    host cores (one process each; the rate is units / the slowest worker's loop time).  Returns (hash row, orders row,
    hashes, verdicts) or None when the reference is not available."""
    import subprocess
    from oracle import refenv
    if refenv.ref_src() is None:
        return None
    cores = max(1, min(cores or os.cpu_count() or 1, max(len(pairs), len(order_cases))))
    jobs = [{"pairs": [[hex(a), hex(b)] for a, b in pairs[w::cores]],
             "orders": [[{k: hex(v) for k, v in od.items()}, hex(r), hex(s), hex(pub)] for od, r, s, pub in order_cases[w::cores]]}
            for w in range(cores)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "oracle", "ref_worker.py")], stdin=subprocess.PIPE,
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env) for _ in range(cores)]
    outs = []
    deadline = time.time() + timeout
    for p, job in zip(procs, jobs):
        try:
            o, e = p.communicate(json.dumps(job), timeout=max(1.0, deadline - time.time()))
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise RuntimeError("reference worker timed out")
        if p.returncode != 0:
            for q in procs:
                q.kill()
            raise RuntimeError("reference worker failed: " + e[-300:])
        outs.append(json.loads(o))
    hashes, verdicts = [None] * len(pairs), [None] * len(order_cases)
    for w, o in enumerate(outs):
        hashes[w::cores] = [int(h, 16) for h in o["hashes"]]
        verdicts[w::cores] = o["verdicts"]
    t_hash, t_ord = max(o["t_hash"] for o in outs), max(o["t_orders"] for o in outs)
    src = "staged copy oracle/_ref" if "_ref" in outs[0]["src"] else outs[0]["src"]
    return ({"value": len(pairs) / t_hash, "unit": "hash/s", "cores": cores, "kind": "reference",
             "sample": "signature.pedersen_hash (signature.py:296-318) on %d of the pairs, one process per core (%d), %.1f s; %s"
                       % (len(pairs), cores, t_hash, src)},
            {"value": len(order_cases) / t_ord, "unit": "orders/s", "cores": cores, "kind": "reference",
             "sample": "get_limit_order_msg + verify (perpetual_messages.py:212-286, signature.py:217-260) on %d orders, "
                       "one process per core (%d), %.1f s; %s" % (len(order_cases), cores, t_ord, src)},
            hashes, verdicts)


# ---- the GPU side ------------------------------------------------------------------------------------------------------
def synthetic_orders(n, seed):
    from conftest import rand_felts
    g = np.random.Generator(np.random.PCG64(seed))
    o = {"asset_id_synthetic": rand_felts(n, seed + 1), "asset_id_collateral": rand_felts(n, seed + 2),
         "asset_id_fee": rand_felts(n, seed + 3), "is_buying_synthetic": g.integers(0, 2, n, dtype=np.uint8)}
    o["asset_id_synthetic"][:, 2:] = 0
    for f in ("asset_id_collateral", "asset_id_fee"):
        o[f][:, 3] &= np.uint64((1 << 58) - 1)
    for f in ("amount_synthetic", "amount_collateral", "max_amount_fee", "position_id"):
        o[f] = g.integers(0, 2**64, n, dtype=np.uint64)
    for f in ("nonce", "expiration_timestamp"):
        o[f] = g.integers(0, 2**32, n, dtype=np.uint32)
    return o


def order_dict(o, i):
    from stark_perpetual_b200._lib import limbs_to_ints
    d = {}
    for f in ("asset_id_synthetic", "asset_id_collateral", "asset_id_fee"):
        d[f] = limbs_to_ints(o[f][i:i + 1])[0]
    for f in ("is_buying_synthetic", "amount_synthetic", "amount_collateral", "max_amount_fee", "position_id", "nonce",
              "expiration_timestamp"):
        d[f] = int(o[f][i])
    return d


def signed_orders(ctx, n, seed, n_keys=1024):
    """n synthetic limit orders, signed on the device (spg_limit_order_msg_batch + spg_sign_batch) with n_keys distinct keys,
    1 % of them corrupted afterwards (a flipped bit of r, s, the key or an order field).  Returns (orders, r, s, pub_x,
    expected statuses, corrupted indices, message hashes, n_keys, bad-status count of the signing calls)."""
    from conftest import rand_felts
    orders = synthetic_orders(n, seed)
    privs = rand_felts(n_keys, seed + 72)
    privs[:, 3] &= np.uint64(0x03ffffffffffffff)
    privs[:, 0] |= np.uint64(1)
    pubs, _ = ctx.private_to_stark_key(privs)
    kidx = np.arange(n) % n_keys
    msgs, mst = ctx.limit_order_msg(orders)
    r, s, sst = ctx.sign(msgs, privs[kidx])
    px = pubs[kidx].copy()
    g = np.random.Generator(np.random.PCG64(seed + 94))
    bad = g.choice(n, size=max(1, n // 100), replace=False)
    expect = np.ones(n, dtype=np.uint8)
    for k, i in enumerate(bad):
        which = k % 4
        bit = np.uint64(1) << np.uint64(int(g.integers(0, 60)))
        if which == 0:
            r[i, 0] ^= bit
        elif which == 1:
            s[i, 0] ^= bit
        elif which == 2:
            px[i, 0] ^= bit
        else:
            orders["amount_collateral"][i] ^= bit
        expect[i] = 0
    return orders, r, s, px, expect, bad, msgs, n_keys, int((sst != 0).sum()) + int((mst != 0).sum())


def ecdsa_air(ctx, log_n=20, n_queries=30, verify=True):
    """The second AIR at the headline trace size: 2^log_n / 256 signatures made on the device, their `verify` walks as a
    25-column trace (spg_ecdsa_air_trace), one proof (spg_prove_ecdsa), checked by the oracle's verifier."""
    import random
    import torch
    from stark_perpetual_b200._lib import ints_to_limbs, limbs_to_ints
    from stark_perpetual_b200.ecdsa_air import air_inputs
    count = (1 << log_n) >> 8
    rng = random.Random(2026)
    privs = [rng.randrange(1, 1 << 250) for _ in range(count)]
    msgs = [rng.randrange(1, 1 << 251) for _ in range(count)]
    kx, ky, st = ctx.private_to_stark_key(ints_to_limbs(privs), want_y=True)
    r, s_, st2 = ctx.sign(ints_to_limbs(msgs), ints_to_limbs(privs))
    assert not st.any() and not st2.any()
    ri, si = limbs_to_ints(r), limbs_to_ints(s_)
    keys = list(zip(limbs_to_ints(kx), limbs_to_ints(ky)))
    m_, r_, w_, kx_, ky_ = air_inputs(msgs, ri, si, keys)
    trace_ms = 1e30
    for _ in range(2):                      # the first call also pays the kernels' lazy module load
        trace = ctx.ecdsa_air_trace(log_n, m_, r_, w_, kx_, ky_)
        trace_ms = min(trace_ms, ctx.last_kernel_ms)
    d_trace = torch.from_numpy(trace.view(np.int64)).cuda()
    best, stages = 1e30, None
    for _ in range(4):
        proof = ctx.prove_ecdsa(None, log_n, m_, kx_, n_queries, device_ptr=d_trace.data_ptr())
        if ctx.last_kernel_ms < best:
            best, stages = ctx.last_kernel_ms, [round(ctx.stage_ms(k), 3) for k in range(9)]
    pinned = torch.from_numpy(trace.view(np.int64)).pin_memory()          # the upload overlaps the LDE only from pinned memory
    host_trace = pinned.numpy().view(np.uint64)
    e2e_ms = 1e30
    for _ in range(2):
        t0 = time.perf_counter()
        proof_h = ctx.prove_ecdsa(host_trace, log_n, m_, kx_, n_queries)
        e2e_ms = min(e2e_ms, (time.perf_counter() - t0) * 1e3)
    row = {"log_n": log_n, "signatures": count, "trace_ms": trace_ms, "proof_ms": best, "e2e_ms": e2e_ms,
           "signatures_per_s": count / (best * 1e-3), "proof_bytes": len(proof), "same_proof_from_host_trace": proof_h == proof,
           "stage_ms": dict(zip(["lde_trace", "merkle_trace", "air_composition", "lde_chunks", "merkle_chunks", "oods_eval",
                                 "deep_quotient", "fri", "queries"], stages))}
    if verify:
        from oracle import stark
        st = stark.verify(proof)
        row["verified_by_oracle"] = st["air"] == "ecdsa" and st["msgs"] == msgs and st["keys"] == [k[0] for k in keys]
    del d_trace
    return row


def measure(ctx, sm_mhz=1965.0, with_reference=True, n_orders=65536, hbm_gbs=6452.8):
    from conftest import rand_felts
    from oracle.pedersen import pedersen_hash as o_pedersen
    from stark_perpetual_b200._lib import ints_to_limbs, limbs_to_ints
    ceil = int_ceiling(sm_mhz)
    aux = {"int_ceiling_field_mul_per_s": ceil, "sm_mhz": sm_mhz}

    # ---- configs[0]: Pedersen hash of 1024 pairs (seed 1001), and the same kernel at 2^20 pairs
    for n, tag in ((1024, "cfg0_pedersen_1024"), (1 << 20, "pedersen_2^20")):
        x, y = rand_felts(n, 1001), rand_felts(n, 1002)
        (out, st), e2e_ms = _timed(lambda: ctx.pedersen_hash2(x, y))
        k_ms = ctx.last_kernel_ms
        muls = float(np.mean([pedersen_muls(int(p)) for p in (_popcounts(x[:4096]) + _popcounts(y[:4096]))]))
        row = {"n": n, "ms": k_ms, "e2e_ms": e2e_ms, "hash_per_s": n / (k_ms * 1e-3), "e2e_hash_per_s": n / (e2e_ms * 1e-3),
               "muls_per_hash": muls, "field_mul_per_s": n * muls / (k_ms * 1e-3),
               "frac_of_int_ceiling": n * muls / (k_ms * 1e-3) / ceil, "bytes_per_hash": 96, "bad_status": int((st != 0).sum())}
        xs, ys, os_ = limbs_to_ints(x[:8]), limbs_to_ints(y[:8]), limbs_to_ints(out[:8])
        row["oracle_sample_ok"] = all(o_pedersen(a, b) == c for a, b, c in zip(xs, ys, os_))
        aux[tag] = row
    # ---- configs[1]: NTT 2^18, one vector and a batch of 64 (device-resident, natural in / bit-reversed out)
    import torch
    for batch, tag in ((1, "cfg1_ntt_2^18_single"), (64, "cfg1_ntt_2^18_batch64")):
        v = torch.from_numpy(rand_felts(batch << 18, 1002).view(np.int64)).cuda()
        best = 1e30
        for _ in range(8):
            ctx.ntt_device(v.data_ptr(), 18, batch)
            best = min(best, ctx.last_kernel_ms)
        muls = batch * (1 << 17) * 18
        byts = batch * 64.0 * (1 << 18)
        aux[tag] = {"batch": batch, "ms": best, "us_per_vector": 1e3 * best / batch, "algorithmic_muls": muls,
                    "field_mul_per_s": muls / (best * 1e-3), "frac_of_int_ceiling": muls / (best * 1e-3) / ceil,
                    "algorithmic_bytes": byts, "hbm_gbs": byts / (best * 1e-3) / 1e9, "hbm_frac": byts / (best * 1e-3) / 1e9 / hbm_gbs}
        del v
    # ---- configs[4]: n_orders limit orders; (a) validly signed on the device with 1 % corrupted, (b) random signatures
    n = n_orders
    orders, r, s, px, expect, bad, msgs, n_keys, sign_bad = signed_orders(ctx, n, 1005)
    st, e2e_ms = _timed(lambda: ctx.limit_order_verify(orders, r, s, px), reps=3)
    k_ms = ctx.last_kernel_ms
    pm, pr = float(_popcounts(msgs[:2048]).mean()), float(_popcounts(r[:2048]).mean())
    muls_valid = 4 * pedersen_muls(252) + verify_muls(pm, pr, 125.5, 1)
    row = {"n": n, "mix": "valid signatures made by spg_sign_batch over %d keys, %d corrupted (1 %%)" % (n_keys, len(bad)),
           "ms": k_ms, "e2e_ms": e2e_ms, "orders_per_s": n / (k_ms * 1e-3), "e2e_orders_per_s": n / (e2e_ms * 1e-3),
           "muls_per_order": muls_valid, "field_mul_per_s": n * muls_valid / (k_ms * 1e-3),
           "frac_of_int_ceiling": n * muls_valid / (k_ms * 1e-3) / ceil,
           "status_counts": {int(a): int(b) for a, b in zip(*np.unique(st, return_counts=True))},
           "statuses_as_expected": bool(np.array_equal(st, expect)), "sign_bad_status": sign_bad}
    aux["cfg4_orders_valid_mix"] = row
    rr, ss = rand_felts(n, 32), rand_felts(n, 33)
    for a in (rr, ss):
        a[:, 3] &= np.uint64(0x07ffffffffffffff)
    st2, e2e2 = _timed(lambda: ctx.limit_order_verify(orders, rr, ss, px), reps=3)
    k2 = ctx.last_kernel_ms
    muls_inv = 4 * pedersen_muls(252) + verify_muls(pm, 125.5, 125.5, 2)
    aux["cfg4_orders_invalid_mix"] = {"n": n, "mix": "random (r, s): every order fails after both square roots are tried", "ms": k2,
                                      "e2e_ms": e2e2, "orders_per_s": n / (k2 * 1e-3), "e2e_orders_per_s": n / (e2e2 * 1e-3),
                                      "muls_per_order": muls_inv, "field_mul_per_s": n * muls_inv / (k2 * 1e-3),
                                      "frac_of_int_ceiling": n * muls_inv / (k2 * 1e-3) / ceil,
                                      "status_counts": {int(a): int(b) for a, b in zip(*np.unique(st2, return_counts=True))}}
    # ---- the reference on the host cores, on a bounded sample of the same inputs; its verdicts double as a parity check
    if with_reference:
        cores = os.cpu_count() or 1
        x, y = rand_felts(1024, 1001), rand_felts(1024, 1002)
        n_pairs = min(1024, 16 * cores)
        pairs = list(zip(limbs_to_ints(x[:n_pairs]), limbs_to_ints(y[:n_pairs])))
        sample = sorted(set(list(range(0, n, max(1, n // (2 * cores)))) [:2 * cores] + [int(i) for i in bad[:cores]]))
        rs_, ss_, ks_ = limbs_to_ints(r[sample]), limbs_to_ints(s[sample]), limbs_to_ints(px[sample])
        cases = [(order_dict(orders, i), rs_[k], ss_[k], ks_[k]) for k, i in enumerate(sample)]
        try:
            res = reference_rates(pairs, cases, cores)
        except Exception as e:       # the reference leg must never take the bench line down
            res, aux["cpu_reference_error"] = None, repr(e)[:300]
        if res:
            h, o, hashes, verdicts = res
            gpu_h = limbs_to_ints(ctx.pedersen_hash2(x[:n_pairs], y[:n_pairs])[0])
            h["gpu_equals_reference"] = bool(hashes == gpu_h)
            o["gpu_equals_reference"] = bool([int(st[i]) for i in sample] == list(verdicts))
            aux["cfg0_pedersen_1024"]["cpu_reference"] = h
            aux["cfg4_orders_valid_mix"]["cpu_reference"] = o
        else:
            aux.setdefault("cpu_reference_error", "reference sources not available (oracle/_ref not staged)")
    return aux


if __name__ == "__main__":
    import stark_perpetual_b200 as spg
    print(json.dumps(measure(spg.get_context(0)), indent=1))
