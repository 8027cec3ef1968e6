"""compute_program_hash_chain of cairo-lang (starkware/cairo/bootloaders/hash_program.py, un-vendored; imported by the
reference at src/starkware/cairo/bootloaders/program_hash_test_utils.py:3), restated from its published definition:

    builtin_list   = [int.from_bytes(name.encode("ascii"), "big") for name in program.builtins]
    program_header = [bootloader_version, program.main, len(program.builtins)] + builtin_list
    data_chain     = program_header + program.data
    hash           = compute_hash_chain([len(data_chain)] + data_chain)

PARITY UNPINNED: the golden values (src/services/perpetual/cairo/program_hash.json:2, src/starkware/cairo/dex/
program_hash.json:2) need the compiled programs, i.e. the Cairo compiler, which is not in the image; the chain
primitive itself is pinned through pedersen_hash.  `program` is either an object with .builtins / .main / .data (a
cairo-lang Program) or the dict of a compiled-program JSON ("builtins", "data" as hex strings, the entry point under
identifiers["__main__.main"]["pc"]).
"""
from starkware.cairo.common.hash_chain import compute_hash_chain


def _fields(program):
    if isinstance(program, dict):
        data = [int(v, 16) if isinstance(v, str) else int(v) for v in program["data"]]
        main = program.get("main")
        if main is None:
            main = program["identifiers"]["__main__.main"]["pc"]
        return list(program["builtins"]), int(main), data
    return list(program.builtins), int(program.main), [int(v) for v in program.data]


def compute_program_hash_chain(program, bootloader_version: int = 0, hash_func=None):
    builtins, main, data = _fields(program)
    builtin_list = [int.from_bytes(b.encode("ascii"), "big") for b in builtins]
    program_header = [bootloader_version, main, len(builtins)] + builtin_list
    data_chain = program_header + data
    return compute_hash_chain([len(data_chain)] + data_chain, hash_func)
