"""Bit-level emulator of the row-chain layout of fe_mul_inl / fe_sq_wide (fe25519.cuh) with pruned carry captures.

A row chain adds n 64-bit products into consecutive 64-bit slots of an accumulator with one carry chain and normally ends
with `addc top, top, 0` to capture the carry out of its last word.  When that last word was still zero before the row, the
capture can never fire: hi(a*b) <= 2^32 - 2, so hi + 0 + carry_in <= 2^32 - 1.  This script derives which rows need a
capture (top word possibly non-zero), checks on random and extreme operands that no carry is ever lost, and prints the
capture flags that fe25519.cuh hard-codes.  Run: python tools/check_fe_rows.py
"""
import random
M32 = 0xffffffff


def mul_rows():
    """(acc, start_word, [(a_index, b_index) ...]) in the order of fe_mul_inl"""
    rows = []
    for i in range(0, 8, 2):
        rows.append(("ev", i, [(0, i), (2, i), (4, i), (6, i)]))
        rows.append(("od", i, [(1, i), (3, i), (5, i), (7, i)]))
        rows.append(("od", i, [(0, i + 1), (2, i + 1), (4, i + 1), (6, i + 1)]))
        rows.append(("ev", i + 2, [(1, i + 1), (3, i + 1), (5, i + 1), (7, i + 1)]))
    return rows


def sq_rows():
    out = []
    for i in range(7):
        odd = [(i, j) for j in range(i + 1, 8) if (i + j) % 2 == 1]
        even = [(i, j) for j in range(i + 1, 8) if (i + j) % 2 == 0]
        if odd:
            out.append(("od", odd[0][0] + odd[0][1] - 1, odd))
        if even:
            out.append(("ev", even[0][0] + even[0][1], even))
    return out


def capture_flags(rows):
    """a row needs its capture iff its last word may be non-zero before the row runs"""
    written = {"ev": set(), "od": set()}
    flags = []
    for acc, start, prods in rows:
        top = start + 2 * len(prods) - 1
        need = top in written[acc]
        flags.append(need)
        written[acc].update(range(start, top + 1))
        if need:
            written[acc].add(top + 1)
    return flags


def run(rows, flags, a, b):
    acc = {"ev": [0] * 20, "od": [0] * 20}
    for (name, start, prods), need in zip(rows, flags):
        r = acc[name]; carry = 0; w = start
        for (i, j) in prods:
            p = a[i] * b[j]
            t = r[w] + (p & M32) + carry; r[w] = t & M32; carry = t >> 32
            t = r[w + 1] + (p >> 32) + carry; r[w + 1] = t & M32; carry = t >> 32
            w += 2
        if need:
            t = r[w] + carry; r[w] = t & M32
            assert t >> 32 == 0
        else:
            assert carry == 0, "a pruned capture would have fired"
    return sum(x << (32 * k) for k, x in enumerate(acc["ev"])) + (sum(x << (32 * k) for k, x in enumerate(acc["od"])) << 32)


def operands(rnd, it):
    if it % 7 == 0:
        return [M32] * 8
    if it % 7 == 1:
        return [rnd.choice([0, 1, M32, M32 - 1, 0x80000000]) for _ in range(8)]
    return [rnd.getrandbits(32) for _ in range(8)]


if __name__ == "__main__":
    rnd = random.Random(7)
    mr, sr = mul_rows(), sq_rows()
    mf, sf = capture_flags(mr), capture_flags(sr)
    for it in range(20000):
        a, b = operands(rnd, it), operands(rnd, it * 3 + 1)
        A = sum(x << (32 * i) for i, x in enumerate(a)); B = sum(x << (32 * i) for i, x in enumerate(b))
        assert run(mr, mf, a, b) == A * B
        off = run(sr, sf, a, a)     # off-diagonal half of the square
        assert 2 * off + sum((a[i] * a[i]) << (64 * i) for i in range(8)) == A * A
    print("mul rows:", [(r[0], r[1], int(f)) for r, f in zip(mr, mf)], "captures", sum(mf), "of", len(mf))
    print("sq  rows:", [(r[0], r[1], len(r[2]), int(f)) for r, f in zip(sr, sf)], "captures", sum(sf), "of", len(sf))
