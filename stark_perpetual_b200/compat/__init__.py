"""Reference-compatible import tree: put this directory on sys.path (before the reference's `src`) and
`from starkware.crypto.signature.signature import pedersen_hash, verify, sign, ...` resolves to the GPU-backed
implementations, with the reference's names, argument meaning and error behaviour
(src/starkware/crypto/signature/signature.py)."""


def extend_package_path(path, name):
    """Called by every package of this tree: lets modules that the compat tree does NOT provide resolve in the same package
    of any later sys.path entry (the reference's `src`), so that e.g. `starkware.python.utils` or the reference's own
    `starkware.cairo.bootloaders.program_hash_test_utils` stay importable next to the GPU-backed modules.  The compat
    directory stays first: a module present here always wins."""
    import os
    import sys
    rel = os.path.join(*name.split("."))
    for p in sys.path:
        d = os.path.join(p, rel)
        if os.path.isdir(d) and d not in path:
            path.append(d)
