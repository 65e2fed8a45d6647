"""Index arithmetic of the walker kernel's small-ensemble instantiation (csrc/mw2.cuh, ILP >= 2), modelled on the host:
the restructured loops must visit exactly what the plain loops of the full-GPU instantiation visit, in the same order --
that is why they leave every bit of the results where it was.  (The kernels themselves are held to the oracle on the
GPU, tests/test_gpu_*.py; this pins the invariants the restructuring relies on.)"""
import random


def test_item_passes_in_flight_cover_the_same_passes():
    """one / two / three passes per turn as the items left need == one pass per turn (local_energies, item loop)"""
    for first in range(0, 32):
        for nitems in range(first, 97):
            seq, t0 = [], first
            while t0 < nitems:
                left = nitems - t0
                if left > 64:
                    seq += [t0, t0 + 32, t0 + 64]; t0 += 96
                elif left > 32:
                    seq += [t0, t0 + 32]; t0 += 64
                else:
                    seq += [t0]; t0 += 32
            assert seq == list(range(first, nitems, 32))


def test_two_ended_expansion_writes_the_same_table():
    """lowest slot from the front, highest from the back per turn == ascending enumeration of a centre's mask"""
    rnd = random.Random(7)
    for _ in range(20000):
        m = rnd.getrandbits(32) if rnd.random() < 0.9 else 0
        want = [i for i in range(32) if (m >> i) & 1]
        tab = [None] * len(want)
        it, ie, b = 0, len(want) - 1, m
        while b:
            lo, hi = (b & -b).bit_length() - 1, b.bit_length() - 1
            tab[it] = lo; it += 1
            tab[ie] = hi; ie -= 1                   # the odd one out is written twice, with the same value
            b &= b - 1
            b &= ~(1 << hi)
        assert tab == want


def test_paired_rotation_steps_keep_the_order_of_additions():
    """two rotation steps d per turn + remainder == the plain loop over d (pair sums centred on the moved molecule)"""
    for stride in (1, 2):
        for d0 in ((1,) if stride == 1 else (1, 2)):
            for maxd in range(0, 17):
                new, d = [], d0
                while d + stride <= maxd:
                    new += [d, d + stride]; d += 2 * stride
                if d <= maxd:
                    new.append(d)
                assert new == list(range(d0, maxd + 1, stride))
