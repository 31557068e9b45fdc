"""Host-side restatement of the symmetric scan's tile schedule (csrc/scan_mma.cu, runSymmetric + scanMmaSymKernel).

The owner of row super block A visits
  * the near window: the column super blocks (A - d) mod S for d = -w .. w, ROW DIRECTION ONLY (both owners of such a
    pair of super blocks visit it), with w = max(16, ceil(N / 32 / 512)) capped at (S - 1) / 2;
  * the far sweep: the offsets d = w + 1 .. S / 2, both directions -- except the offset S / 2 of an even S, which both
    owners visit and which is therefore row direction only.
The property the kernel relies on: every ORDERED pair of super blocks (X gets candidates from Y) is produced exactly
once, whatever subset of row super blocks each GPU owns.  No GPU needed."""
from collections import Counter

import pytest


def near_half_width(S, N=None, option=0):
    N = S * 256 if N is None else N
    w = max(16, (N // 32 + 511) // 512)
    if option:
        w = option
    if 2 * w + 1 > S:
        w = (S - 1) // 2
    return w


def directed_contributions(S, w, owners=None):
    """(receiver, source) super-block pairs produced by the schedule; owners: iterable of the row super blocks visited."""
    out = Counter()
    half = S // 2 if S % 2 == 0 else 0
    for A in (range(S) if owners is None else owners):
        for d in range(-w, w + 1):
            out[(A, (A - d) % S)] += 1            # near window: rows of A receive columns of C
        for d in range(w + 1, S // 2 + 1):
            C = (A - d) % S
            out[(A, C)] += 1                      # row direction
            if d != half:
                out[(C, A)] += 1                  # column direction: columns of C receive rows of A
    return out


@pytest.mark.parametrize("S", list(range(1, 80)) + [391, 392, 781])
def test_every_ordered_super_block_pair_exactly_once(S):
    for option in (0, 1, 3):
        w = near_half_width(S, option=option)
        assert 2 * w + 1 <= S
        got = directed_contributions(S, w)
        assert len(got) == S * S
        assert set(got.values()) == {1}


@pytest.mark.parametrize("S,P", [(40, 2), (41, 3), (392, 8), (97, 4)])
def test_ranks_partition_the_contributions(S, P):
    """Row super blocks are dealt to the GPUs in contiguous ranges; together the ranks still produce every ordered pair once."""
    w = near_half_width(S)
    per = (S + P - 1) // P
    total = Counter()
    for r in range(P):
        total.update(directed_contributions(S, w, owners=range(min(S, r * per), min(S, (r + 1) * per))))
    assert len(total) == S * S and set(total.values()) == {1}


@pytest.mark.parametrize("S", [64, 391, 392, 3907])
def test_tiles_executed_are_about_half(S):
    w = near_half_width(S)
    tiles = S * (2 * w + 1 + max(0, S // 2 - w))
    assert tiles <= S * S // 2 + S * (w + 2)
