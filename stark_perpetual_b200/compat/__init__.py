"""Reference-compatible import tree: put this directory on sys.path (before the reference's `src`) and
`from starkware.crypto.signature.signature import pedersen_hash, verify, sign, ...` resolves to the GPU-backed
implementations, with the reference's names, argument meaning and error behaviour
(src/starkware/crypto/signature/signature.py)."""
