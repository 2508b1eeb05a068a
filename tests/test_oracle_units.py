"""Stage-level checks of the oracle against first-principles restatements of each reference function."""
import numpy as np

import oracle_ffi as O
import ctypes as C


def direct_idct_f64(F):
    """transform.rs:66-84 in float64."""
    out = np.zeros((8, 8))
    for y in range(8):
        for x in range(8):
            s = 0.0
            for v in range(8):
                for u in range(8):
                    au = 1 / np.sqrt(2) if u == 0 else 1.0
                    av = 1 / np.sqrt(2) if v == 0 else 1.0
                    s += au * av * F[v, u] * np.cos((2 * x + 1) * u * np.pi / 16) * np.cos((2 * y + 1) * v * np.pi / 16)
            out[y, x] = s / 4
    return out


def test_idct_matches_the_formula():
    rng = np.random.default_rng(1)
    for _ in range(20):
        F = np.zeros((8, 8), np.float32)
        k = rng.integers(1, 30)
        F.flat[rng.choice(64, k, replace=False)] = rng.integers(-500, 500, k)
        got = O.idct_8x8(F)
        assert np.abs(got - direct_idct_f64(F.astype(np.float64))).max() < 2e-3
        assert np.array_equal(got, O.idct_8x8(F, O.COS_CALL))


def test_dc_only_block_is_flat():
    F = np.zeros((8, 8), np.float32)
    F[0, 0] = 1016
    out = O.idct_8x8(F)
    assert np.abs(out - 127.0).max() < 1e-4          # alpha(0)^2 = 0.49999997 in f32
    assert O.lib().oracle_f32_to_u8(float(np.float32(out[0, 0]) + np.float32(128.0))) in (254, 255)


def test_value_correction_is_extend():
    L = O.lib()
    for size in range(1, 12):
        for v in range(1 << size):
            want = v if v >= (1 << (size - 1)) else v - (1 << size) + 1     # T.81 F.2.2.1
            assert L.oracle_value_correction(v, size) == want
    assert L.oracle_value_correction(0, 0) == 0


def test_f32_to_u8_truncates_and_clamps():
    L = O.lib()
    for x, want in [(-3.5, 0), (-0.0, 0), (0.99, 0), (1.0, 1), (127.999, 127), (254.99998, 254), (255.0, 255),
                    (255.5, 255), (1e9, 255)]:
        assert L.oracle_f32_to_u8(x) == want


def test_ycbcr_to_rgb_formula():
    L = O.lib()
    out = (C.c_uint8 * 3)()
    rng = np.random.default_rng(2)
    for _ in range(200):
        y, cb, cr = [float(np.float32(v)) for v in rng.uniform(-128, 127, 3)]
        L.oracle_ycbcr_to_rgb(y, cb, cr, out)
        r = cr * (2 - 2 * 0.299) + y
        b = cb * (2 - 2 * 0.114) + y
        g = (y - 0.114 * b - 0.299 * r) / 0.587
        want = [int(np.clip(v + 128, 0, 255)) for v in (r, g, b)]
        assert all(abs(int(a) - w) <= 1 for a, w in zip(out, want))


def test_zigzag_is_a_permutation_and_matches_t81():
    z = [O.lib().oracle_zigzag_indices()[k] for k in range(64)]
    assert sorted(z) == list(range(64))
    # T.81 Figure A.6: walk the anti-diagonals
    want, (r, c), up = [], (0, 0), True
    for _ in range(64):
        want.append(r * 8 + c)
        if up:
            if c == 7: r, up = r + 1, False
            elif r == 0: c, up = c + 1, False
            else: r, c = r - 1, c + 1
        else:
            if r == 7: c, up = c + 1, True
            elif c == 0: r, up = r + 1, True
            else: r, c = r + 1, c - 1
    assert z == want


def test_unstuff():
    L = O.lib()
    src = bytes([1, 0xff, 0x00, 2, 0xff, 0x00, 0x00, 0xff, 0xd9])
    out = C.create_string_buffer(len(src))
    n = L.oracle_unstuff(src, len(src), out)
    assert out.raw[:n] == bytes([1, 0xff, 2, 0xff, 0x00, 0xff, 0xd9])
    assert L.oracle_unstuff(bytes([1, 0xff]), 2, out) == C.c_size_t(-1).value     # mod.rs:378 index out of bounds


def test_code_table_is_canonical():
    bits = bytes([0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d])
    vals = bytes(range(162))
    ln = (C.c_uint8 * 256)(); code = (C.c_uint16 * 256)(); val = (C.c_uint8 * 256)()
    n = O.lib().oracle_build_codes(bits, vals, 162, ln, code, val)
    assert n == 162
    c, k = 0, 0
    for l in range(1, 17):                       # T.81 Figure C.2
        for _ in range(bits[l - 1]):
            assert (ln[k], code[k], val[k]) == (l, c, vals[k])
            c += 1; k += 1
        c <<= 1
