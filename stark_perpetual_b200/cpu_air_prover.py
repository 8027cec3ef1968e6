#!/usr/bin/env python3
"""File-level prover entry point in the shape of the prover CLI that follows `cairo-run --proof_mode` in the reference's
build (src/starkware/cairo/lang/cairo_cmake_rules.cmake:72-110 names the runner's artefacts; the prover it feeds is
Stone's `cpu_air_prover --out_file --private_input_file --public_input_file --prover_config_file --parameter_file`, which
the reference never vendors -- SURVEY.md section 8(b) item 4):

    python -m stark_perpetual_b200.cpu_air_prover --out_file proof.bin --private_input_file private.json \\
        --public_input_file public.json [--parameter_file params.json] [--prover_config_file cfg.json]

for the AIR this repo implements (the Pedersen hash chain):
  private input  {"trace_path": "<file>"}   raw trace, 25 columns x 2^log_n rows x 32 bytes (4 x u64 little-endian limbs,
                 canonical), column-major -- or {"ys_path": "<file>"} (5 x 2^log_n/512 felts): the witness is then generated
                 on the device (spg_pedersen_chain_trace), the role cairo-run plays for a Cairo program;
  public input   {"log_n": .., "chain_log": .., "x0": ["0x..", x5]}
  parameters     {"n_queries": 30}   (optional)
  prover config  {"device": 0}       (optional)
For the second AIR (the ECDSA builtin) the public input names it:
  public input   {"air": "ecdsa", "instances": [{"msg": "0x..", "key": "0x.."}, ...]}   -- a power of two (>= 2) of them
  private input  {"signatures": [{"r": "0x..", "s": "0x..", "key_y": "0x.."}, ...]}      -- one per instance, in order
Output: the proof bytes (format: DESIGN.md section 4) and, next to it, `<out_file>.public.json` with the public outputs.
Traces of the Cairo layouts themselves are not provable here.
"""
import argparse
import json
import sys

import numpy as np


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--out_file", required=True)
    ap.add_argument("--private_input_file", required=True)
    ap.add_argument("--public_input_file", required=True)
    ap.add_argument("--parameter_file")
    ap.add_argument("--prover_config_file")
    args = ap.parse_args(argv)
    from . import Context
    from ._lib import limbs_to_ints
    pub = json.load(open(args.public_input_file))
    prv = json.load(open(args.private_input_file))
    params = json.load(open(args.parameter_file)) if args.parameter_file else {}
    cfg = json.load(open(args.prover_config_file)) if args.prover_config_file else {}
    if pub.get("air") == "ecdsa":
        return _main_ecdsa(args, pub, prv, params, cfg)
    log_n, chain_log = int(pub["log_n"]), int(pub.get("chain_log", 0))
    x0 = [int(v, 16) if isinstance(v, str) else int(v) for v in pub["x0"]]
    if len(x0) != 5:
        raise SystemExit("public input: x0 must hold 5 seeds")
    n = 1 << log_n
    ctx = Context(int(cfg.get("device", 0)))
    if "trace_path" in prv:
        trace = np.fromfile(prv["trace_path"], dtype="<u8")
        if trace.size != 25 * n * 4:
            raise SystemExit("trace file holds %d words, expected 25 x 2^%d x 4" % (trace.size, log_n))
        trace = trace.reshape(25 * n, 4)
    elif "ys_path" in prv:
        ys = np.fromfile(prv["ys_path"], dtype="<u8")
        if ys.size != 5 * (n >> 9) * 4:
            raise SystemExit("ys file holds %d words, expected 5 x 2^%d / 512 x 4" % (ys.size, log_n))
        trace = ctx.pedersen_chain_trace(log_n, chain_log, x0, ys.reshape(-1, 4))
    else:
        raise SystemExit("private input needs trace_path or ys_path")
    proof = ctx.prove(trace, log_n, chain_log, x0, int(params.get("n_queries", 30)))
    with open(args.out_file, "wb") as f:
        f.write(proof)
    outs = limbs_to_ints(trace.reshape(25, n, 4)[[5 * l for l in range(5)], n - 1])
    with open(args.out_file + ".public.json", "w") as f:
        json.dump({"log_n": log_n, "chain_log": chain_log, "x0": [hex(v) for v in x0], "outs": [hex(v) for v in outs],
                   "proof_bytes": len(proof)}, f)
    return 0


def _main_ecdsa(args, pub, prv, params, cfg):
    from . import Context
    from ._lib import SpgError
    from .ecdsa_air import prove_signatures

    def num(v):
        return int(v, 16) if isinstance(v, str) else int(v)
    inst, sigs = pub["instances"], prv["signatures"]
    if len(inst) != len(sigs):
        raise SystemExit("private input must carry one signature per instance")
    ctx = Context(int(cfg.get("device", 0)))
    try:
        proof, log_n = prove_signatures([num(i["msg"]) for i in inst], [num(s["r"]) for s in sigs], [num(s["s"]) for s in sigs],
                                        [(num(i["key"]), num(s["key_y"])) for i, s in zip(inst, sigs)],
                                        int(params.get("n_queries", 30)), ctx=ctx)
    except (ValueError, SpgError) as e:
        raise SystemExit("ecdsa air: %s" % e)
    with open(args.out_file, "wb") as f:
        f.write(proof)
    with open(args.out_file + ".public.json", "w") as f:
        json.dump({"air": "ecdsa", "log_n": log_n, "instances": len(inst), "proof_bytes": len(proof)}, f)
    return 0


if __name__ == "__main__":
    sys.exit(main())
