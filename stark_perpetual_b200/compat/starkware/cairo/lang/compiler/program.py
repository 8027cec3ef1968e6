"""The sliver of cairo-lang's `starkware.cairo.lang.compiler.program` that the reference's program-hash test uses
(src/starkware/cairo/bootloaders/program_hash_test_utils.py:4,8: `Program.Schema().load(json.load(open(path)))`), for
images without cairo-lang: a compiled-program JSON becomes an object with the three fields compute_program_hash_chain
reads -- builtins, main, data.  EXTERNAL interface (cairo-lang is an un-vendored dependency of the reference)."""


class Program:
    def __init__(self, builtins, main, data, prime=None):
        self.builtins, self.main, self.data, self.prime = list(builtins), main, list(data), prime

    class Schema:
        def load(self, obj):
            data = [int(v, 16) if isinstance(v, str) else int(v) for v in obj["data"]]
            main = obj.get("main")
            if main is None:
                main = obj["identifiers"]["__main__.main"]["pc"]
            prime = obj.get("prime")
            return Program(obj["builtins"], int(main), data, int(prime, 16) if isinstance(prime, str) else prime)
