import os, sys, random
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import stark_perpetual_b200 as spg
from stark_perpetual_b200._lib import ints_to_limbs, limbs_to_ints
P = 2**251 + 17 * 2**192 + 1
R = 2**256
ctx = spg.get_context(0)
rng = random.Random(3)
a = [0, 1, 2**32, 2**255, 2**256 - 1, 3, P - 1] + [rng.randrange(2**254) for _ in range(9)]
b = [5, 1, 2**32, 2**255, 2**256 - 1, 7, P - 1] + [rng.randrange(2**254) for _ in range(9)]
A, B = ints_to_limbs(a), ints_to_limbs(b)
lo = limbs_to_ints(ctx.field_op("widelo", A, B)); hi = limbs_to_ints(ctx.field_op("widehi", A, B))
for i, (x, y) in enumerate(zip(a, b)):
    got = lo[i] + (hi[i] << 256)
    print("wide", i, "OK" if got == x * y else "BAD got=%x want=%x" % (got, x * y))
rm = limbs_to_ints(ctx.field_op("rawmul", A, B))
Rinv = pow(R, -1, P)
for i, (x, y) in enumerate(zip(a, b)):
    ok = rm[i] % P == x * y * Rinv % P and rm[i] < P + (x * y >> 256) + 1
    print("redc", i, "OK" if ok else "BAD got=%x want=%x" % (rm[i], x * y * Rinv % P))
rd = limbs_to_ints(ctx.field_op("reduce", A, B))
for i, x in enumerate(a):
    if x < 4 * P:
        print("reduce", i, "OK" if rd[i] == x % P else "BAD got=%x want=%x" % (rd[i], x % P))
dbg = ctx.field_op("redcdbg", A, B)
for i in (0, 2, 12, 1):
    print("dbg", i, [hex(int(v)) for v in dbg[i].view('<u4')], "T=%x" % (a[i] * b[i]))
