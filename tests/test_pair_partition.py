"""The half-ring partition of the unordered pairs used by the potential-energy kernel (csrc/energy.cu `pair_weight`,
`potential_kernel`): a CPU restatement of the rule, checked exhaustively for small chunk counts — every unordered
chunk pair is counted exactly once, every target chunk gets the same work, and the ring range a target block streams
contains every chunk it has a non-zero weight for."""
from fractions import Fraction

import pytest


def pair_weight(own: int, q: int, ring: int) -> Fraction:
    if ring <= 0:
        return Fraction(1, 2)
    d = (q - own) % ring
    if d == 0 or 2 * d == ring:
        return Fraction(1, 2)
    return Fraction(1) if 2 * d < ring else Fraction(0)


@pytest.mark.parametrize("C", range(1, 41))
def test_every_unordered_chunk_pair_is_counted_once(C):
    for p in range(C):
        assert pair_weight(p, p, C) == Fraction(1, 2)          # own square: both orders, halved
        for q in range(p + 1, C):
            assert pair_weight(p, q, C) + pair_weight(q, p, C) == 1
    work = [sum(pair_weight(p, q, C) > 0 for q in range(C)) for p in range(C)]
    assert max(work) == min(work)                              # equal work for every target chunk => for every rank


@pytest.mark.parametrize("C", range(1, 41))
@pytest.mark.parametrize("own_chunks", [1, 2, 4])
def test_streamed_ring_range_covers_every_weighted_chunk(C, own_chunks):
    for g0 in range(0, C, own_chunks):
        own = min(own_chunks, C - g0)                          # a partial last block
        span = min(C, own + C // 2)                            # potential_kernel: ring offsets this block streams
        streamed = [(g0 + off) % C for off in range(span)]
        assert len(set(streamed)) == len(streamed)             # no chunk twice
        for p in range(g0, g0 + own):
            for q in range(C):
                if pair_weight(p, q, C) > 0:
                    assert q in streamed
