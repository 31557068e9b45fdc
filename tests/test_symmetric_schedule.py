"""Host-side restatement of the symmetric scan's tile schedule (csrc/scan_mma.cu, scanMmaSymKernel): the CTA that owns
super block A visits the column super blocks (A - d) mod S for the offsets d = 0 .. S/2; the diagonal (d = 0) and, for
even S, the offset S/2 are one-directional (both owners visit them), every other tile is two-directional.  The
property the kernel relies on: every ORDERED pair of super blocks (X gets candidates from Y) is produced exactly once.
No GPU needed."""
from collections import Counter

import pytest


def directed_contributions(S):
    """(receiver, source) super-block pairs produced by the schedule."""
    out = Counter()
    offsets = S // 2 + 1
    half = S // 2 if S % 2 == 0 else 0
    for A in range(S):
        for d in range(offsets):
            C = (A + S - d) % S
            out[(A, C)] += 1                      # row direction: rows of A receive columns of C
            if d != 0 and d != half:
                out[(C, A)] += 1                  # column direction: columns of C receive rows of A
    return out


@pytest.mark.parametrize("S", list(range(1, 40)) + [391, 392, 781])
def test_every_ordered_super_block_pair_exactly_once(S):
    got = directed_contributions(S)
    assert len(got) == S * S
    assert set(got.values()) == {1}


@pytest.mark.parametrize("S", [2, 3, 7, 8, 391, 392])
def test_tiles_executed_are_about_half(S):
    offsets = S // 2 + 1
    tiles = S * offsets
    assert tiles <= S * S // 2 + S + (S if S % 2 == 0 else 0)
