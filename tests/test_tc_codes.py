"""CPU: the arithmetic the tensor-core Hamming engine rests on, checked on the code functions the kernels run.

d(q, r) = popc((q ^ r) & m) = popc(q & m) + sum_s a_s * r_s with a_s = m_s (1 - 2 q_s).  The kernels turn packed bits
into narrow-float operand codes (csrc/hamming_tc.cu: expand_panel_word_fp4/fp8, expand_query_chunk) such that every
product is exactly -1, 0 or +1.  `snv_debug_tc_codes` runs those functions on the host; this test decodes the codes by
the E2M1 / E4M3 format definitions and checks the contraction word by word, plus the documented site -> K-position
maps (the permutation must be the same on both operands, or the dot product pairs the wrong sites)."""
import numpy as np
import pytest


def e2m1(nib):
    s, e, m = (nib >> 3) & 1, (nib >> 1) & 3, nib & 1
    v = 0.5 * m if e == 0 else (1 + 0.5 * m) * 2.0 ** (e - 1)
    return -v if s else v


def e4m3(byte):
    s, e, m = (byte >> 7) & 1, (byte >> 3) & 15, byte & 7
    v = (m / 8.0) * 2.0 ** -6 if e == 0 else (1 + m / 8.0) * 2.0 ** (e - 7)
    return -v if s else v


def decode(codes, fp4):
    if fp4:
        out = []
        for b in codes.tolist():
            out += [e2m1(b & 15), e2m1(b >> 4)]
        return np.array(out)
    return np.array([e4m3(b) for b in codes.tolist()])


@pytest.fixture(scope="module")
def L():
    from rag_snvbert_b200 import _lib

    _lib.lib()
    return _lib


@pytest.mark.parametrize("fp4", [True, False])
def test_word_contraction_is_masked_hamming_minus_bias(L, fp4):
    rng = np.random.default_rng(4 if fp4 else 8)
    words = [(0, 0, 0), (0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF, 0), (0, 0xFFFFFFFF, 0xFFFFFFFF),
             (0xFFFFFFFF, 0, 0xFFFFFFFF), (0xAAAAAAAA, 0xFFFFFFFF, 0x55555555), (0x80000001, 0xFFFFFFFF, 0x80000000)]
    words += [tuple(int(x) for x in rng.integers(0, 1 << 32, 3)) for _ in range(400)]
    for q, m, r in words:
        a, b = L.debug_tc_codes(fp4, q, m, r)
        pa, pb = decode(a, fp4), decode(b, fp4)
        prod = pa * pb
        assert set(np.unique(prod)).issubset({-1.0, 0.0, 1.0}), (hex(q), hex(m), hex(r))
        want = bin((q ^ r) & m).count("1") - bin(q & m).count("1")
        assert prod.sum() == want, (hex(q), hex(m), hex(r))
        # each operand alone is exact too: |query code| * |panel code| = 1 wherever both are non-zero
        assert (np.abs(prod[(pa != 0) & (pb != 0)]) == 1).all()


@pytest.mark.parametrize("fp4", [True, False])
def test_site_to_k_position_map(L, fp4):
    """one set panel bit -> exactly one non-zero code, at the position hamming_tc.cu documents, and the query side uses the
    same position for that site"""
    for bit in range(32):
        a, b = L.debug_tc_codes(fp4, 1 << bit, 0xFFFFFFFF, 1 << bit)
        pa, pb = decode(a, fp4), decode(b, fp4)
        nz = np.nonzero(pb)[0]
        assert len(nz) == 1
        if fp4:
            # nibble n of byte 4 j + n / 2  <->  bit 4 n + j   (element index = 2 * byte + (n & 1))
            n, j = bit // 4, bit % 4
            pos = 2 * (4 * j + n // 2) + (n & 1)
        else:
            # byte 4 j + b (chunk pair of 32 bytes)  <->  bit 8 b + j
            bb, j = bit // 8, bit % 8
            pos = 4 * j + bb
        assert nz[0] == pos, (bit, nz[0], pos)
        # the query carries allele 1 (negative) at the same position, magnitude the inverse of the panel's
        assert pa[pos] * pb[pos] == -1.0
        # every other query position is an observed allele 0: positive, never zero
        assert (np.delete(pa, pos) > 0).all()
    # unobserved sites contribute nothing whatever the alleles
    a, b = L.debug_tc_codes(fp4, 0xFFFFFFFF, 0, 0xFFFFFFFF)
    assert (decode(a, fp4) == 0).all()
