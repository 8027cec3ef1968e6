#!/usr/bin/env python3
"""One worker of the reference-timed CPU baseline (TEST / BENCH INFRASTRUCTURE): imports the REFERENCE's own Python
through oracle/refenv.py, reads {"pairs": [[x, y], ...], "orders": [[order_dict, r, s, pub], ...]} (hex strings) on
stdin, runs signature.pedersen_hash on the pairs and get_limit_order_msg + verify on the orders, and prints the results
with the seconds each loop took (imports and the first warm-up hash are outside the timed loops)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.setrecursionlimit(5000)


def main():
    from oracle import refenv
    sig, _mu, pm = refenv.import_reference()
    job = json.load(sys.stdin)
    pairs = [(int(a, 16), int(b, 16)) for a, b in job.get("pairs", [])]
    orders = job.get("orders", [])
    sig.pedersen_hash(1, 2)
    t0 = time.perf_counter()
    hashes = [hex(sig.pedersen_hash(a, b)) for a, b in pairs]
    t_hash = time.perf_counter() - t0
    verdicts = []
    t0 = time.perf_counter()
    for od, r, s, pub in orders:
        msg = pm.get_limit_order_msg(**{k: int(v, 16) for k, v in od.items()})
        try:
            verdicts.append(1 if sig.verify(msg, int(r, 16), int(s, 16), int(pub, 16)) else 0)
        except AssertionError:
            verdicts.append(2)
    t_orders = time.perf_counter() - t0
    json.dump({"hashes": hashes, "t_hash": t_hash, "verdicts": verdicts, "t_orders": t_orders, "src": refenv.ref_src()}, sys.stdout)


if __name__ == "__main__":
    main()
