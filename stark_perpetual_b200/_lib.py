"""ctypes binding of libspg.so (include/spg.h).  Fails loudly when the library or a GPU is missing."""
import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LOCK = threading.Lock()

FIELD_PRIME = 2**251 + 17 * 2**192 + 1

SPG_DEVICE_PTRS = 1
SPG_NO_SYNC = 2
NTT_NAT_TO_REV, NTT_REV_TO_NAT, NTT_NAT_TO_NAT = 0, 1, 2


class SpgError(RuntimeError):
    pass


def lib_path():
    # SPG_LIB: an alternative build of the same CUDA library (kernel-variant experiments); never a fallback
    return os.environ.get("SPG_LIB") or os.path.join(_HERE, "libspg.so")


def _load():
    global _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = lib_path()
        if not os.path.exists(path):
            raise SpgError(
                "libspg.so not built (%s missing): run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C stark_perpetual_b200/csrc`. There is no CPU fallback." % path)
        lib = C.CDLL(path)
        u64p, u8p, vp = C.POINTER(C.c_uint64), C.POINTER(C.c_uint8), C.c_void_p
        sig = {
            "spg_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
            "spg_destroy": (None, [vp]),
            "spg_last_error": (C.c_char_p, [vp]),
            "spg_device_count": (C.c_int, []),
            "spg_last_kernel_ms": (C.c_double, [vp]),
            "spg_launch_count": (C.c_uint64, [vp]),
            "spg_synchronize": (C.c_int, [vp]),
            "spg_field_op": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_size_t, C.c_int]),
            "spg_bench_field_mul": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
            "spg_ntt": (C.c_int, [vp, vp, C.c_uint, C.c_size_t, C.c_int, C.c_int, C.c_int]),
            "spg_pedersen_hash2_batch": (C.c_int, [vp, vp, vp, vp, vp, C.c_size_t, C.c_int]),
            "spg_pedersen_hash2_batch_be32": (C.c_int, [vp, vp, vp, vp, vp, C.c_size_t]),
            "spg_pedersen_chain_batch": (C.c_int, [vp, vp, C.c_size_t, vp, vp, C.c_size_t, C.c_int]),
            "spg_merkle_commit": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t, vp, vp, C.c_int]),
            "spg_pedersen_chain_trace": (C.c_int, [vp, C.c_uint, C.c_uint, vp, vp, vp, C.c_int]),
            "spg_air_eval": (C.c_int, [vp, vp, C.c_uint, C.c_uint, vp, vp, vp, vp, C.c_int]),
            "spg_prove": (C.c_int, [vp, vp, C.c_uint, C.c_uint, vp, C.c_uint, vp, C.c_size_t, C.POINTER(C.c_size_t), C.c_int]),
            "spg_ecdsa_air_trace": (C.c_int, [vp, C.c_uint, vp, vp, vp, vp, vp, vp, C.c_int]),
            "spg_air_eval_ecdsa": (C.c_int, [vp, vp, C.c_uint, vp, vp, vp, vp, C.c_int]),
            "spg_prove_ecdsa": (C.c_int, [vp, vp, C.c_uint, vp, vp, C.c_uint, vp, C.c_size_t, C.POINTER(C.c_size_t), C.c_int]),
            "spg_ecdsa_verify_batch": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_size_t, C.c_int]),
            "spg_private_to_stark_key_batch": (C.c_int, [vp, vp, vp, vp, vp, C.c_size_t, C.c_int]),
            "spg_sign_batch": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_size_t, C.c_int]),
            "spg_pedersen_hash_point_batch": (C.c_int, [vp, vp, C.c_size_t, vp, vp, vp, C.c_size_t, C.c_int]),
            "spg_get_y_coordinate_batch": (C.c_int, [vp, vp, vp, vp, C.c_size_t, C.c_int]),
            "spg_mimic_ec_mult_air_batch": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_size_t, C.c_int]),
            "spg_pedersen_merkle_tree": (C.c_int, [vp, vp, C.c_size_t, vp, vp, vp, C.c_int]),
            "spg_ec_op_batch": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, C.c_size_t, C.c_int]),
            "spg_hash_chain_rfold_batch": (C.c_int, [vp, vp, C.c_size_t, vp, vp, C.c_size_t, C.c_int]),
            "spg_position_hash_batch": (C.c_int, [vp, vp, vp, vp, C.c_size_t, C.c_int]),
            "spg_merkle_update_siblings": (C.c_int, [vp, C.c_uint, vp, C.c_size_t, vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]),
            "spg_merkle_update_node_count": (C.c_int, [vp, C.c_uint, vp, C.c_size_t, vp]),
            "spg_merkle_multi_update": (C.c_int, [vp, C.c_uint, vp, vp, vp, C.c_size_t, vp, C.c_size_t, vp, vp, vp, vp, C.c_int]),
            "spg_field_sqrt_batch": (C.c_int, [vp, vp, vp, vp, C.c_size_t, C.c_int]),
            "spg_limit_order_msg_batch": (C.c_int, [vp, vp, vp, vp, C.c_size_t, C.c_int]),
            "spg_message_hash_batch": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_size_t, C.c_int]),
            "spg_limit_order_verify_batch": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_size_t, C.c_int]),
            "spg_set_stream": (C.c_int, [vp, vp]),
            "spg_comm_unique_id": (C.c_int, [vp, vp]),
            "spg_comm_init": (C.c_int, [vp, C.c_int, C.c_int, vp]),
            "spg_prove_sharded": (C.c_int, [vp, vp, C.c_uint, C.c_uint, vp, vp, C.c_uint, vp, C.c_size_t, C.POINTER(C.c_size_t), C.c_int]),
            "spg_prove_ecdsa_sharded": (C.c_int, [vp, vp, C.c_uint, vp, vp, C.c_uint, vp, C.c_size_t, C.POINTER(C.c_size_t), C.c_int]),
            "spg_stage_ms": (C.c_double, [vp, C.c_int]),
            "spg_lde": (C.c_int, [vp, vp, C.c_uint, C.c_size_t, C.c_uint, vp, vp, C.c_int]),
            "spg_lde_coeffs": (C.c_int, [vp, vp, C.c_uint, C.c_size_t, vp, vp, C.c_int]),
            # stage-level entry points driven by prover.py (multi-GPU)
            "spg_stage_merkle": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t, C.c_int, vp]),
            "spg_stage_air": (C.c_int, [vp, C.c_uint, C.c_uint, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]),
            "spg_stage_cp_split": (C.c_int, [vp, C.c_uint, vp, C.c_int, C.c_int, vp]),
            "spg_stage_poly_eval": (C.c_int, [vp, C.c_uint, vp, vp, C.c_int, vp, C.c_int, vp]),
            "spg_stage_deep": (C.c_int, [vp, C.c_uint, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]),
            "spg_stage_fri_fold": (C.c_int, [vp, vp, C.c_uint, C.c_int, C.c_int, vp, C.c_int, vp]),
            "spg_stage_open": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t, C.c_int, vp, vp, C.c_int, vp, vp]),
            "spg_stage_check_oods": (C.c_int, [vp, C.c_uint, C.c_uint, vp, vp, vp, vp, vp]),
            "spg_stage_last_layer": (C.c_int, [vp, vp, C.c_uint, C.c_int, vp]),
            "spg_lde_cosets": (C.c_int, [vp, vp, C.c_uint, C.c_size_t, C.c_uint, C.c_size_t, C.c_size_t, vp, C.c_int]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)     # AttributeError if the symbol is missing: loud by design
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
        return lib


# ---- felt <-> numpy helpers -------------------------------------------------------------------
def ints_to_limbs(values):
    """list of Python ints in [0, 2^256) -> (n, 4) uint64 little-endian limbs."""
    buf = b"".join(int(v).to_bytes(32, "little") for v in values)
    return np.frombuffer(buf, dtype="<u8").reshape(-1, 4).copy()


def limbs_to_ints(arr):
    a = np.ascontiguousarray(arr, dtype="<u8").reshape(-1, 4)
    raw = a.tobytes()
    return [int.from_bytes(raw[32 * i:32 * i + 32], "little") for i in range(a.shape[0])]


def _ptr(arr):
    return arr.ctypes.data_as(C.c_void_p)


class Context:
    """One spg_ctx (one GPU, one stream)."""

    def __init__(self, device=0):
        self._lib = _load()
        h = C.c_void_p()
        rc = self._lib.spg_create(int(device), C.byref(h))
        if rc != 0:
            msg = self._lib.spg_last_error(h).decode() if h else ""
            if h:
                self._lib.spg_destroy(h)
            raise SpgError("spg_create(device=%d) failed with %d %s -- a CUDA device is required, "
                           "there is no CPU fallback" % (device, rc, msg))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.spg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise SpgError("libspg error %d: %s" % (rc, self._lib.spg_last_error(self._h).decode()))

    @property
    def last_kernel_ms(self):
        return self._lib.spg_last_kernel_ms(self._h)

    @property
    def launch_count(self):
        return int(self._lib.spg_launch_count(self._h))

    def synchronize(self):
        self._check(self._lib.spg_synchronize(self._h))

    def set_stream(self, cuda_stream_handle):
        """Run this context's work on an existing CUDA stream (e.g. torch.cuda.current_stream().cuda_stream)."""
        self._check(self._lib.spg_set_stream(self._h, C.c_void_p(cuda_stream_handle)))

    def stage_ms(self, stage):
        return self._lib.spg_stage_ms(self._h, int(stage))

    # ---- field layer ----
    def field_op(self, op, a, b=None):
        """a, b: (n,4) uint64 canonical felts. op: 'mul','add','sub','inv','pow'."""
        code = {"mul": 0, "add": 1, "sub": 2, "inv": 3, "pow": 4, "rawmul": 5, "widelo": 6, "widehi": 7, "reduce": 8, "sublazy2": 9,
                "partial": 10, "reducefull": 11, "addraw": 12, "rawsqr": 13, "sqrwidelo": 14, "sqrwidehi": 15}[op]
        a = np.ascontiguousarray(a, dtype=np.uint64)
        out = np.empty_like(a)
        bp = None
        if b is not None:
            b = np.ascontiguousarray(b, dtype=np.uint64)
            assert b.shape == a.shape
            bp = _ptr(b)
        self._check(self._lib.spg_field_op(self._h, code, _ptr(a), bp, _ptr(out), a.shape[0], 0))
        return out

    def bench_field_mul(self, iters=2000, chains=4):
        m, w = C.c_double(), C.c_double()
        self._check(self._lib.spg_bench_field_mul(self._h, iters, chains, C.byref(m), C.byref(w)))
        return m.value, w.value

    # ---- Pedersen ----
    def pedersen_hash2(self, x, y):
        """x, y: (n, 4) uint64 canonical felts -> (out (n, 4), status (n,) uint8)."""
        x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4)
        y = np.ascontiguousarray(y, dtype=np.uint64).reshape(-1, 4)
        assert x.shape == y.shape
        out = np.empty_like(x)
        st = np.empty(x.shape[0], dtype=np.uint8)
        self._check(self._lib.spg_pedersen_hash2_batch(self._h, _ptr(x), _ptr(y), _ptr(out), _ptr(st), x.shape[0], 0))
        return out, st

    def pedersen_hash2_be32(self, x, y):
        """x, y: (n, 32) uint8 big-endian -> (out (n, 32) uint8, status)."""
        x = np.ascontiguousarray(x, dtype=np.uint8).reshape(-1, 32)
        y = np.ascontiguousarray(y, dtype=np.uint8).reshape(-1, 32)
        out = np.empty_like(x)
        st = np.empty(x.shape[0], dtype=np.uint8)
        self._check(self._lib.spg_pedersen_hash2_batch_be32(self._h, _ptr(x), _ptr(y), _ptr(out), _ptr(st), x.shape[0]))
        return out, st

    def pedersen_chain(self, elems, chain_len):
        """elems: (n * chain_len, 4) canonical felts, row-major [n][chain_len] -> (out (n, 4), status)."""
        e = np.ascontiguousarray(elems, dtype=np.uint64).reshape(-1, 4)
        assert e.shape[0] % chain_len == 0
        n = e.shape[0] // chain_len
        out = np.empty((n, 4), dtype=np.uint64)
        st = np.empty(n, dtype=np.uint8)
        self._check(self._lib.spg_pedersen_chain_batch(self._h, _ptr(e), chain_len, _ptr(out), _ptr(st), n, 0))
        return out, st

    def pedersen_hash_point(self, elems, n_elems):
        """elems: (n * n_elems, 4) canonical felts -> (x (n, 4), y (n, 4), status)."""
        e = np.ascontiguousarray(elems, dtype=np.uint64).reshape(-1, 4)
        n = e.shape[0] // n_elems
        x, y, st = np.empty((n, 4), np.uint64), np.empty((n, 4), np.uint64), np.empty(n, np.uint8)
        self._check(self._lib.spg_pedersen_hash_point_batch(self._h, _ptr(e), n_elems, _ptr(x), _ptr(y), _ptr(st), n, 0))
        return x, y, st

    def get_y_coordinate(self, xs):
        x = np.ascontiguousarray(xs, dtype=np.uint64).reshape(-1, 4)
        y, st = np.empty_like(x), np.empty(x.shape[0], np.uint8)
        self._check(self._lib.spg_get_y_coordinate_batch(self._h, _ptr(x), _ptr(y), _ptr(st), x.shape[0], 0))
        return y, st

    def mimic_ec_mult_air(self, m, point_xy, shift_xy):
        """m: (n, 4); point_xy, shift_xy: (n, 8) -> (out_xy (n, 8), status)."""
        m = np.ascontiguousarray(m, dtype=np.uint64).reshape(-1, 4)
        p = np.ascontiguousarray(point_xy, dtype=np.uint64).reshape(-1, 8)
        s = np.ascontiguousarray(shift_xy, dtype=np.uint64).reshape(-1, 8)
        out, st = np.empty_like(p), np.empty(m.shape[0], np.uint8)
        self._check(self._lib.spg_mimic_ec_mult_air_batch(self._h, _ptr(m), _ptr(p), _ptr(s), _ptr(out), _ptr(st), m.shape[0], 0))
        return out, st

    def ec_op(self, op, a_xy, b=None):
        """math_utils.py:59-100 on the device.  op 0: ec_add(a, b) with b (n, 8); op 1: ec_double(a); op 2: ec_mult(m, a) with
        b = m (n, 4).  a_xy (n, 8) canonical (x, y).  Returns (out_xy (n, 8), status): 0 ok, 1 the reference's assertion
        fails (equal x / y == 0), 2 a coordinate >= p, 3 m == 0."""
        a = np.ascontiguousarray(a_xy, dtype=np.uint64).reshape(-1, 8)
        bb = None if b is None else np.ascontiguousarray(b, dtype=np.uint64).reshape(a.shape[0], -1)
        out, st = np.empty_like(a), np.empty(a.shape[0], np.uint8)
        self._check(self._lib.spg_ec_op_batch(self._h, op, _ptr(a), _ptr(bb) if bb is not None else None, _ptr(out), _ptr(st),
                                              a.shape[0], 0))
        return out, st

    def field_sqrt(self, a):
        """math_utils.py:36-47: (smaller square root (n, 4), status) with status 0 ok, 1 non-residue, 2 a >= p."""
        x = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        y, st = np.empty_like(x), np.empty(x.shape[0], np.uint8)
        self._check(self._lib.spg_field_sqrt_batch(self._h, _ptr(x), _ptr(y), _ptr(st), x.shape[0], 0))
        return y, st

    def hash_chain_rfold(self, data, length):
        """n right-folded chains h(d0, h(d1, ... h(d[len-2], d[len-1]))): data (n * length, 4) -> (out (n, 4), status (n,))."""
        d = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 4)
        assert length >= 1 and d.shape[0] % length == 0
        n = d.shape[0] // length
        out, st = np.empty((n, 4), dtype=np.uint64), np.empty(n, dtype=np.uint8)
        self._check(self._lib.spg_hash_chain_rfold_batch(self._h, _ptr(d), length, _ptr(out), _ptr(st), n, 0))
        return out, st

    # ---- state trees (f-4) ----
    def position_hash(self, public_keys, collateral, offsets, asset_ids, balances, funding):
        """position_hash (hash.cairo:22-74) of n positions in CSR form: public_keys (n, 4) felts, collateral (n,) int64,
        offsets (n + 1,) uint64, asset_ids (total, 2) uint64 (128-bit little-endian), balances / funding (total,) int64.
        Returns (hashes (n, 4), status (n,))."""
        pk = np.ascontiguousarray(public_keys, dtype=np.uint64).reshape(-1, 4)
        n = pk.shape[0]
        col = np.ascontiguousarray(collateral, dtype=np.int64).reshape(n)
        off = np.ascontiguousarray(offsets, dtype=np.uint64).reshape(n + 1)
        total = int(off[-1])
        aid = np.ascontiguousarray(asset_ids, dtype=np.uint64).reshape(-1, 2)
        bal = np.ascontiguousarray(balances, dtype=np.int64).reshape(-1)
        fi = np.ascontiguousarray(funding, dtype=np.int64).reshape(-1)
        assert off[0] == 0 and aid.shape[0] == total and bal.shape[0] == total and fi.shape[0] == total
        struct = (C.c_void_p * 6)(pk.ctypes.data, col.ctypes.data, off.ctypes.data, aid.ctypes.data if total else None,
                                  bal.ctypes.data if total else None, fi.ctypes.data if total else None)
        out, st = np.empty((n, 4), dtype=np.uint64), np.empty(n, dtype=np.uint8)
        self._check(self._lib.spg_position_hash_batch(self._h, struct, _ptr(out), _ptr(st), n, 0))
        return out, st

    def merkle_update_siblings(self, height, keys):
        """[(level, node index)] of the sibling hashes spg_merkle_multi_update needs, in its order."""
        k = np.ascontiguousarray(keys, dtype=np.uint64).reshape(-1)
        cnt = C.c_size_t(0)
        self._check(self._lib.spg_merkle_update_siblings(self._h, height, _ptr(k), k.shape[0], None, None, 0, C.byref(cnt)))
        lv, ix = np.empty(cnt.value, dtype=np.uint8), np.empty(cnt.value, dtype=np.uint64)
        if cnt.value:
            self._check(self._lib.spg_merkle_update_siblings(self._h, height, _ptr(k), k.shape[0], _ptr(lv), _ptr(ix), cnt.value,
                                                             C.byref(cnt)))
        return list(zip(lv.tolist(), ix.tolist()))

    def merkle_multi_update(self, height, keys, prev_leaves, new_leaves, siblings, want_nodes=False):
        """Sparse Merkle multi-update (state.cairo:151-173): keys strictly increasing; prev / new leaves (n, 4); siblings
        (n_siblings, 4) in the order of merkle_update_siblings.  Returns (prev_root (4,), new_root (4,), status, nodes)
        with nodes = list per level of an array (2, m, 4) (previous values, new values) or None."""
        k = np.ascontiguousarray(keys, dtype=np.uint64).reshape(-1)
        n = k.shape[0]
        pl = np.ascontiguousarray(prev_leaves, dtype=np.uint64).reshape(n, 4)
        nl = np.ascontiguousarray(new_leaves, dtype=np.uint64).reshape(n, 4)
        sb = np.ascontiguousarray(siblings, dtype=np.uint64).reshape(-1, 4)
        pr, nr, st = np.empty(4, dtype=np.uint64), np.empty(4, dtype=np.uint64), np.empty(1, dtype=np.uint8)
        nodes, counts = None, None
        if want_nodes:
            counts = (C.c_size_t * height)()
            self._check(self._lib.spg_merkle_update_node_count(self._h, height, _ptr(k), n, counts))
            nodes = np.empty((2 * sum(counts), 4), dtype=np.uint64)
        self._check(self._lib.spg_merkle_multi_update(self._h, height, _ptr(k), _ptr(pl), _ptr(nl), n, _ptr(sb) if sb.shape[0] else None,
                                                      sb.shape[0], _ptr(pr), _ptr(nr), _ptr(nodes) if nodes is not None else None,
                                                      _ptr(st), 0))
        per_level = None
        if want_nodes:
            per_level, off = [], 0
            for m in counts:
                per_level.append(nodes[off:off + 2 * m].reshape(2, m, 4))
                off += 2 * m
        return pr, nr, int(st[0]), per_level

    def pedersen_merkle_tree(self, leaves, want_nodes=False):
        """leaves: (n, 4) canonical felts, n a power of two -> (root (4,), nodes (n - 1, 4) or None, status)."""
        lv = np.ascontiguousarray(leaves, dtype=np.uint64).reshape(-1, 4)
        n = lv.shape[0]
        root = np.empty(4, dtype=np.uint64)
        nodes = np.empty((n - 1, 4), dtype=np.uint64) if want_nodes else None
        st = np.zeros(1, dtype=np.uint8)
        self._check(self._lib.spg_pedersen_merkle_tree(self._h, _ptr(lv), n, _ptr(root), _ptr(nodes) if want_nodes else None,
                                                       _ptr(st), 0))
        return root, nodes, int(st[0])

    # ---- ECDSA ----
    def ecdsa_verify(self, msg, r, s, pub_x, pub_y=None):
        """All arguments (n, 4) uint64.  Returns status (n,) uint8: 1 valid, 0 invalid, 2 the reference raises."""
        arrs = [np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4) for a in (msg, r, s, pub_x)]
        n = arrs[0].shape[0]
        assert all(a.shape[0] == n for a in arrs)
        yp = None
        if pub_y is not None:
            pub_y = np.ascontiguousarray(pub_y, dtype=np.uint64).reshape(-1, 4)
            assert pub_y.shape[0] == n
            yp = _ptr(pub_y)
        st = np.empty(n, dtype=np.uint8)
        self._check(self._lib.spg_ecdsa_verify_batch(self._h, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), _ptr(arrs[3]),
                                                     yp, _ptr(st), n, 0))
        return st

    def private_to_stark_key(self, priv, want_y=False):
        p = np.ascontiguousarray(priv, dtype=np.uint64).reshape(-1, 4)
        out = np.empty_like(p)
        outy = np.empty_like(p) if want_y else None
        st = np.empty(p.shape[0], dtype=np.uint8)
        self._check(self._lib.spg_private_to_stark_key_batch(self._h, _ptr(p), _ptr(out), _ptr(outy) if want_y else None,
                                                             _ptr(st), p.shape[0], 0))
        return (out, outy, st) if want_y else (out, st)

    def sign(self, msg, priv, seeds=None):
        """msg, priv: (n, 4) canonical; seeds: (n,) uint64 or None -> (r (n, 4), s (n, 4), status)."""
        m = np.ascontiguousarray(msg, dtype=np.uint64).reshape(-1, 4)
        d = np.ascontiguousarray(priv, dtype=np.uint64).reshape(-1, 4)
        assert m.shape == d.shape
        sd = None if seeds is None else np.ascontiguousarray(seeds, dtype=np.uint64).reshape(-1)
        assert sd is None or sd.shape[0] == m.shape[0]
        r, s, st = np.empty_like(m), np.empty_like(m), np.empty(m.shape[0], dtype=np.uint8)
        self._check(self._lib.spg_sign_batch(self._h, _ptr(m), _ptr(d), _ptr(sd) if sd is not None else None, _ptr(r), _ptr(s),
                                             _ptr(st), m.shape[0], 0))
        return r, s, st

    MSG_KINDS = {"transfer": (4, 3, 7), "conditional_transfer": (5, 4, 7), "withdrawal_to_address": (7, 2, 4),
                 "price": (100, 2, 2)}

    def message_hash(self, kind, felts, ints):
        """kind: a key of MSG_KINDS; felts: list of (n, 4) uint64 arrays, ints: list of (n,) uint64 arrays, in the
        order include/spg.h lists for the kind -> (msg (n, 4), status (n,))."""
        code, nf, ni = self.MSG_KINDS[kind]
        assert len(felts) == nf and len(ints) == ni
        fa = [np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4) for a in felts]
        ia = [np.ascontiguousarray(a, dtype=np.uint64).reshape(-1) for a in ints]
        n = fa[0].shape[0]
        assert all(a.shape[0] == n for a in fa + ia)
        ptrs = (C.c_void_p * 11)()
        for k, a in enumerate(fa):
            ptrs[k] = a.ctypes.data
        for k, a in enumerate(ia):
            ptrs[4 + k] = a.ctypes.data
        out, st = np.empty((n, 4), dtype=np.uint64), np.empty(n, dtype=np.uint8)
        self._check(self._lib.spg_message_hash_batch(self._h, code, C.cast(ptrs, C.c_void_p), _ptr(out), _ptr(st), n, 0))
        return out, st

    # ---- perpetual limit orders ----
    _ORDER_LAYOUT = (("asset_id_synthetic", np.uint64, 4), ("asset_id_collateral", np.uint64, 4), ("asset_id_fee", np.uint64, 4),
                     ("is_buying_synthetic", np.uint8, 1), ("amount_synthetic", np.uint64, 1), ("amount_collateral", np.uint64, 1),
                     ("max_amount_fee", np.uint64, 1), ("position_id", np.uint64, 1), ("nonce", np.uint32, 1),
                     ("expiration_timestamp", np.uint32, 1))

    def _orders_struct(self, arr):
        """dict of numpy arrays -> (spg_limit_orders as a ctypes pointer array, n, keep-alive list)"""
        keep, n = [], None
        for name, dt, width in self._ORDER_LAYOUT:
            a = np.ascontiguousarray(arr[name], dtype=dt).reshape(-1, width)
            n = a.shape[0] if n is None else n
            assert a.shape[0] == n, name
            keep.append(a)
        ptrs = (C.c_void_p * len(keep))(*[a.ctypes.data for a in keep])
        return ptrs, n, keep

    def limit_order_msg(self, arr):
        """arr: dict of numpy arrays in the spg_limit_orders layout -> (msg (n, 4) uint64, status (n,) uint8)."""
        ptrs, n, keep = self._orders_struct(arr)
        out = np.empty((n, 4), dtype=np.uint64)
        st = np.empty(n, dtype=np.uint8)
        self._check(self._lib.spg_limit_order_msg_batch(self._h, ptrs, _ptr(out), _ptr(st), n, 0))
        return out, st

    def limit_order_verify(self, arr, r, s, pub_x):
        ptrs, n, keep = self._orders_struct(arr)
        sig = [np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4) for a in (r, s, pub_x)]
        assert all(a.shape[0] == n for a in sig)
        st = np.empty(n, dtype=np.uint8)
        self._check(self._lib.spg_limit_order_verify_batch(self._h, ptrs, _ptr(sig[0]), _ptr(sig[1]), _ptr(sig[2]), _ptr(st), n, 0))
        return st

    # ---- NTT ----
    def ntt(self, data, log_n, inverse=False, order=NTT_NAT_TO_NAT):
        """data: (batch * 2^log_n, 4) uint64 canonical felts; returns a new array."""
        arr = np.array(data, dtype=np.uint64, order="C", copy=True).reshape(-1, 4)
        n = 1 << log_n
        assert arr.shape[0] % n == 0
        self._check(self._lib.spg_ntt(self._h, _ptr(arr), log_n, arr.shape[0] // n, int(bool(inverse)), order, 0))
        return arr

    def ntt_device(self, dev_ptr, log_n, batch, inverse=False, order=NTT_NAT_TO_REV):
        """In-place transform of device-resident data (pointer from torch .data_ptr())."""
        self._check(self._lib.spg_ntt(self._h, C.c_void_p(dev_ptr), log_n, batch, int(bool(inverse)), order,
                                      SPG_DEVICE_PTRS))


    # ---- LDE ----
    def _offset(self, offset):
        if offset is None:
            return None, None
        arr = ints_to_limbs([offset])
        return arr, _ptr(arr)

    def lde(self, trace, log_n, n_cols, log_blowup, offset=None):
        """trace: (n_cols * 2^log_n, 4) canonical felts, column-major; returns ([B * n_cols * N], 4),
        layout [B][n_cols][N] (include/spg.h spg_lde)."""
        tr = np.ascontiguousarray(trace, dtype=np.uint64).reshape(-1, 4)
        assert tr.shape[0] == n_cols << log_n
        out = np.empty((tr.shape[0] << log_blowup, 4), dtype=np.uint64)
        keep, op = self._offset(offset)
        self._check(self._lib.spg_lde(self._h, _ptr(tr), log_n, n_cols, log_blowup, op, _ptr(out), 0))
        return out

    def lde_device(self, trace_ptr, log_n, n_cols, log_blowup, out_ptr, offset=None, sync=True):
        keep, op = self._offset(offset)
        flags = SPG_DEVICE_PTRS | (0 if sync else SPG_NO_SYNC)
        self._check(self._lib.spg_lde(self._h, C.c_void_p(trace_ptr), log_n, n_cols, log_blowup, op,
                                      C.c_void_p(out_ptr), flags))

    def lde_coeffs_device(self, trace_ptr, log_n, n_cols, coeffs_ptr, offset=None, sync=True):
        keep, op = self._offset(offset)
        flags = SPG_DEVICE_PTRS | (0 if sync else SPG_NO_SYNC)
        self._check(self._lib.spg_lde_coeffs(self._h, C.c_void_p(trace_ptr), log_n, n_cols, op,
                                             C.c_void_p(coeffs_ptr), flags))

    def lde_cosets_device(self, coeffs_ptr, log_n, n_cols, log_blowup, coset_begin, coset_count, out_ptr, sync=True):
        flags = SPG_DEVICE_PTRS | (0 if sync else SPG_NO_SYNC)
        self._check(self._lib.spg_lde_cosets(self._h, C.c_void_p(coeffs_ptr), log_n, n_cols, log_blowup,
                                             coset_begin, coset_count, C.c_void_p(out_ptr), flags))

    # ---- Merkle / AIR / proof ----
    def merkle_commit(self, table, n_cols, rows, want_tree=False):
        """table: (8 * n_cols * rows, 4) uint64, layout [8][n_cols][rows] -> root bytes (and the tree)."""
        t = np.ascontiguousarray(table, dtype=np.uint64).reshape(-1, 4)
        assert t.shape[0] == 8 * n_cols * rows
        root = np.empty(32, dtype=np.uint8)
        tree = np.empty((2 * rows - 1, 32), dtype=np.uint8) if want_tree else None
        self._check(self._lib.spg_merkle_commit(self._h, _ptr(t), n_cols, rows, _ptr(root),
                                                _ptr(tree) if want_tree else None, 0))
        return (root.tobytes(), tree) if want_tree else root.tobytes()

    def pedersen_chain_trace(self, log_n, chain_log, x0, ys):
        """x0: 5 ints; ys: (5 * 2^log_n / 512, 4) uint64 canonical, lane-major -> trace (25 * 2^log_n, 4)."""
        x0a = ints_to_limbs(x0)
        ysa = np.ascontiguousarray(ys, dtype=np.uint64).reshape(-1, 4)
        assert x0a.shape[0] == 5 and ysa.shape[0] == 5 * ((1 << log_n) >> 9)
        out = np.empty((25 << log_n, 4), dtype=np.uint64)
        self._check(self._lib.spg_pedersen_chain_trace(self._h, log_n, chain_log, _ptr(x0a), _ptr(ysa), _ptr(out), 0))
        return out

    def air_eval(self, trace, log_n, chain_log, x0, outs, alpha):
        tr = np.ascontiguousarray(trace, dtype=np.uint64).reshape(-1, 4)
        assert tr.shape[0] == 25 << log_n
        x0a, oa, aa = ints_to_limbs(x0), ints_to_limbs(outs), ints_to_limbs([alpha])
        cp = np.empty((4 << log_n, 4), dtype=np.uint64)
        self._check(self._lib.spg_air_eval(self._h, _ptr(tr), log_n, chain_log, _ptr(x0a), _ptr(oa), _ptr(aa), _ptr(cp), 0))
        return cp

    def prove(self, trace, log_n, chain_log, x0, n_queries=30, device_ptr=None):
        """trace: (25 * 2^log_n, 4) uint64 canonical (host), or device_ptr = address of the same on the GPU.
        Returns the proof bytes."""
        x0a = ints_to_limbs(x0)
        if device_ptr is None:
            tr = np.ascontiguousarray(trace, dtype=np.uint64).reshape(-1, 4)
            assert tr.shape[0] == 25 << log_n
            tp, flags = _ptr(tr), 0
        else:
            tp, flags = C.c_void_p(device_ptr), SPG_DEVICE_PTRS
        cap = 64 + 64 * 32 + 64 * 32 + 128 * 32 + n_queries * (8 * 29 + 8 * 8 + 10 * 24) * 32 + 65536
        buf = np.empty(cap, dtype=np.uint8)
        ln = C.c_size_t(0)
        self._check(self._lib.spg_prove(self._h, tp, log_n, chain_log, _ptr(x0a), n_queries, _ptr(buf), cap,
                                        C.byref(ln), flags))
        return buf[:ln.value].tobytes()

    # ---- second AIR: the ECDSA builtin (csrc/air_ecdsa.cu; CPU twin oracle/stark_ecdsa.py) ----
    def ecdsa_air_trace(self, log_n, msgs, r, w, key_x, key_y):
        """One `verify` call per 256-row block (signature.py:243-260): msgs / r / w / key_x / key_y are (2^log_n / 256, 4)
        uint64 canonical arrays (w = s^-1 mod the curve order, key point on the curve).  -> trace (25 * 2^log_n, 4).
        SpgError where the reference asserts (scalar range, x collision) or returns False."""
        arrs = [np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4) for a in (msgs, r, w, key_x, key_y)]
        nb = (1 << log_n) >> 8
        assert all(a.shape[0] == nb for a in arrs)
        out = np.empty((25 << log_n, 4), dtype=np.uint64)
        self._check(self._lib.spg_ecdsa_air_trace(self._h, log_n, *[_ptr(a) for a in arrs], _ptr(out), 0))
        return out

    def air_eval_ecdsa(self, trace, log_n, msgs, key_x, alpha):
        tr = np.ascontiguousarray(trace, dtype=np.uint64).reshape(-1, 4)
        assert tr.shape[0] == 25 << log_n
        ma = np.ascontiguousarray(msgs, dtype=np.uint64).reshape(-1, 4)
        ka = np.ascontiguousarray(key_x, dtype=np.uint64).reshape(-1, 4)
        assert ma.shape[0] == ka.shape[0] == (1 << log_n) >> 8
        aa = ints_to_limbs([alpha])
        cp = np.empty((4 << log_n, 4), dtype=np.uint64)
        self._check(self._lib.spg_air_eval_ecdsa(self._h, _ptr(tr), log_n, _ptr(ma), _ptr(ka), _ptr(aa), _ptr(cp), 0))
        return cp

    def prove_ecdsa(self, trace, log_n, msgs, key_x, n_queries=30, device_ptr=None):
        """Proof that block b of `trace` is a verifying signature on msgs[b] under the key with x = key_x[b], for every b;
        msgs, key_x: (2^log_n / 256, 4) uint64 canonical (host) -- the public input, carried in the proof header."""
        ma = np.ascontiguousarray(msgs, dtype=np.uint64).reshape(-1, 4)
        ka = np.ascontiguousarray(key_x, dtype=np.uint64).reshape(-1, 4)
        assert ma.shape[0] == ka.shape[0] == (1 << log_n) >> 8
        if device_ptr is None:
            tr = np.ascontiguousarray(trace, dtype=np.uint64).reshape(-1, 4)
            assert tr.shape[0] == 25 << log_n
            tp, flags = _ptr(tr), 0
        else:
            tp, flags = C.c_void_p(device_ptr), SPG_DEVICE_PTRS
        cap = 64 + 64 * 32 + 64 * 32 + 128 * 32 + n_queries * (8 * 29 + 8 * 8 + 10 * 24) * 32 + 65536 + 64 * ma.shape[0]
        buf = np.empty(cap, dtype=np.uint8)
        ln = C.c_size_t(0)
        self._check(self._lib.spg_prove_ecdsa(self._h, tp, log_n, _ptr(ma), _ptr(ka), n_queries, _ptr(buf), cap, C.byref(ln), flags))
        return buf[:ln.value].tobytes()

    # ---- multi-GPU (one process per GPU) ----
    def comm_unique_id(self):
        buf = np.zeros(128, dtype=np.uint8)
        self._check(self._lib.spg_comm_unique_id(self._h, _ptr(buf)))
        return buf.tobytes()

    def comm_init(self, rank, world, unique_id=None):
        idb = np.frombuffer(unique_id, dtype=np.uint8).copy() if unique_id is not None else None
        self._check(self._lib.spg_comm_init(self._h, rank, world, _ptr(idb) if idb is not None else None))

    def prove_ecdsa_sharded(self, cols_local, log_n, msgs, key_x, n_queries=30, device_ptr=None):
        """spg_prove_ecdsa over the communicator: cols_local = this rank's columns (cyclic deal) of an ECDSA-AIR trace; msgs,
        key_x = the whole public input.  Collective; every rank gets the proof, byte-identical to prove_ecdsa's."""
        ma = np.ascontiguousarray(msgs, dtype=np.uint64).reshape(-1, 4)
        ka = np.ascontiguousarray(key_x, dtype=np.uint64).reshape(-1, 4)
        assert ma.shape[0] == ka.shape[0] == (1 << log_n) >> 8
        if device_ptr is None:
            tr = np.ascontiguousarray(cols_local, dtype=np.uint64).reshape(-1, 4)
            tp, flags = (_ptr(tr) if tr.shape[0] else None), 0
        else:
            tp, flags = C.c_void_p(device_ptr), SPG_DEVICE_PTRS
        cap = 64 + 64 * 32 + 64 * 32 + 128 * 32 + n_queries * (8 * 29 + 8 * 8 + 10 * 24) * 32 + 65536 + 64 * ma.shape[0]
        buf = np.empty(cap, dtype=np.uint8)
        ln = C.c_size_t(0)
        self._check(self._lib.spg_prove_ecdsa_sharded(self._h, tp, log_n, _ptr(ma), _ptr(ka), n_queries, _ptr(buf), cap,
                                                      C.byref(ln), flags))
        return buf[:ln.value].tobytes()

    def prove_sharded(self, cols_local, log_n, chain_log, x0, outs, n_queries=30, device_ptr=None):
        """cols_local: this rank's columns (cyclic deal) as a (my_cols * 2^log_n, 4) uint64 host array, or device_ptr = their
        device address.  Collective call; returns the proof bytes on every rank."""
        x0a, oa = ints_to_limbs(x0), ints_to_limbs(outs)
        if device_ptr is None:
            tr = np.ascontiguousarray(cols_local, dtype=np.uint64).reshape(-1, 4)
            tp, flags = (_ptr(tr) if tr.shape[0] else None), 0
        else:
            tp, flags = C.c_void_p(device_ptr), SPG_DEVICE_PTRS
        cap = 64 + 64 * 32 + 64 * 32 + 128 * 32 + n_queries * (8 * 29 + 8 * 8 + 10 * 24) * 32 + 65536
        buf = np.empty(cap, dtype=np.uint8)
        ln = C.c_size_t(0)
        self._check(self._lib.spg_prove_sharded(self._h, tp, log_n, chain_log, _ptr(x0a), _ptr(oa), n_queries, _ptr(buf), cap,
                                                C.byref(ln), flags))
        return buf[:ln.value].tobytes()


_CTX = {}
_CTX_LOCK = threading.Lock()


def get_context(device=0):
    """Process-wide cached Context per device.  Safe to call and to use from several threads: creation is guarded here,
    and every libspg entry point serialises on the context's own mutex (csrc/common.h SPG_LOCK), so concurrent
    pedersen_hash / verify calls on the shared context queue up instead of racing on its stream and buffer pool.  For
    parallel streams of work create one Context per thread."""
    with _CTX_LOCK:
        if device not in _CTX:
            _CTX[device] = Context(device)
        return _CTX[device]
