"""CPU model of the shared-memory tile layout of the small-matrix kernels (csrc/gpr_small_v4.cu, csrc/chol.cu diagonal-block
kernel): 8x8 fp64 tiles whose 16-byte row chunks are XOR-swizzled by (row & 2).  The three access patterns a DMMA
m8n8k4 fragment needs must be free of bank conflicts (32 banks x 4 bytes; a 64-bit access is served per half-warp, a
128-bit access per quarter-warp), and the swizzle must be a bijection on the tile."""
import numpy as np


def tile_off(r, c):  # element offset (doubles) inside a 64-double tile -- same expression as the CUDA source
    return r * 8 + ((((c >> 1) ^ (r & 2))) << 1) + (c & 1)


def conflict_free(addresses, nbytes, lanes_per_wavefront):
    """addresses[lane] = byte address of an nbytes-wide access.  Within one wavefront every bank may serve one 4-byte word
    (lanes reading the same word share it)."""
    for w0 in range(0, 32, lanes_per_wavefront):
        word_of_bank = {}
        for lane in range(w0, w0 + lanes_per_wavefront):
            for b in range(0, nbytes, 4):
                word = (addresses[lane] + b) // 4
                if word_of_bank.setdefault(word % 32, word) != word:
                    return False
    return True


def test_swizzle_is_a_bijection():
    offs = sorted(tile_off(r, c) for r in range(8) for c in range(8))
    assert offs == list(range(64))
    # a row stays inside its own 64-byte line and 16-byte chunks stay intact (needed by the 128-bit C-fragment accesses)
    for r in range(8):
        for c in range(0, 8, 2):
            assert tile_off(r, c) // 8 == r and tile_off(r, c + 1) == tile_off(r, c) + 1 and tile_off(r, c) % 2 == 0


def test_fragment_access_patterns_have_no_bank_conflicts():
    lanes = [(lane >> 2, lane & 3) for lane in range(32)]  # (g, t)
    for s in (0, 1):
        k_major = [8 * tile_off(g, t + 4 * s) for g, t in lanes]   # A / B^T fragment: (row g, col t + 4s), LDS.64
        m_major = [8 * tile_off(t + 4 * s, g) for g, t in lanes]   # transposed fragment: (row t + 4s, col g), LDS.64
        assert conflict_free(k_major, 8, 16), ("K-major", s)
        assert conflict_free(m_major, 8, 16), ("M-major", s)
    c_frag = [8 * tile_off(g, 2 * t) for g, t in lanes]            # accumulator fragment: (row g, cols 2t, 2t+1), LDS/STS.128
    assert conflict_free(c_frag, 16, 8)
    # without the swizzle the M-major pattern WOULD conflict (that is what the XOR is for)
    plain = [8 * ((t + 0) * 8 + g) for g, t in lanes]
    assert not conflict_free(plain, 8, 16)


def test_column_packed_tile_slots():
    """cslot(i, j) = j*NT - j(j-1)/2 + (i - j): a bijection onto 0..NT(NT+1)/2-1 with consecutive rows of a column adjacent."""
    for NT in range(1, 9):
        cs = lambda i, j: j * NT - j * (j - 1) // 2 + (i - j)
        slots = sorted(cs(i, j) for j in range(NT) for i in range(j, NT))
        assert slots == list(range(NT * (NT + 1) // 2))
        for j in range(NT):
            for i in range(j, NT - 1):
                assert cs(i + 1, j) == cs(i, j) + 1
        # left-looking walk of gpr_small_v4: p_k = tile(kb, k); p_{k+1} = p_k + (NT - k - 1)
        for kb in range(NT):
            p = cs(kb, 0)
            for k in range(kb):
                assert p == cs(kb, k)
                p += NT - k - 1
            assert p == cs(kb, kb)
