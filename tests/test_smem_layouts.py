"""CPU model of the shared-memory layouts of the fused kernels: every LDS.128 /
STS.128 pattern the tile kernels issue is enumerated lane by lane and must be free of bank conflicts.

A 128-bit shared access is served per quarter-warp (8 lanes x 16 B = one 128-byte wavefront); two lanes of a
quarter conflict when their double2 slots differ but fall into the same 16-byte bank group (slot mod 8).
The index formulas restate jwave_b200/csrc: pad2 / padr (jwc_fused.cuh, jwc_wpt_fwd.cu), lay and
store_group (jwc_fwt_rev.cu), rl and the TMA-store images of jwc_wpt_rev.cu / jwc_wpt_fwd.cu."""
import itertools


def wavefronts(slots):
    """slots: the double2 index each of the 32 lanes touches -> wavefronts the warp's access needs (>= 4)."""
    total = 0
    for q in range(4):
        groups = {}
        for s in slots[8 * q:8 * q + 8]:
            groups.setdefault(s % 8, set()).add(s)
        total += max(len(v) for v in groups.values())
    return total


def pad2(k):
    return k + (k >> 2)


def padr(k, r):
    return k + k // r


def lay(k):
    return k ^ ((k >> 3) & 1)


def test_forward_windows_and_stores_pad2():
    # k_fwt_fwd, R = 4: lane g reads double2 4g + q of the window (q = 0 .. L/2 + 2)
    for q in range(0, 23):
        assert wavefronts([pad2(4 * g) + q + (q >> 2) for g in range(32)]) == 4
    # Cost of the padded layout (long filters keep it; it was the forward kernel's residual bank conflicts
    # in the ncu captures): the cp.async staging (consecutive slots) and the a_k stores (2 slots per lane)
    # are 2-way.  Filters up to 24 taps use k ^ ((k >> 3) & 3) (fl, jwc_fwt_fwd.cu): all three patterns
    # conflict-free.
    assert wavefronts([pad2(k) for k in range(32)]) == 8
    assert all(wavefronts([pad2(2 * g + e) for g in range(32)]) == 8 for e in range(2))
    xor = lambda k: k ^ ((k >> 3) & 3)
    for q in range(0, 23):
        assert wavefronts([xor(4 * g + q) for g in range(32)]) == 4
    assert wavefronts([xor(k) for k in range(32)]) == 4
    assert all(wavefronts([xor(2 * g + e) for g in range(32)]) == 4 for e in range(2))


def test_reverse_fwt_xor_layout():
    # k_fwt_rev, kRS = 4: lane g reads slot c - w with c = 4 g0 + 2 g + 1, any g0, every window step w
    for g0, w in itertools.product(range(8), range(12)):
        assert wavefronts([lay(4 * g0 + 2 * g + 1 - w + 64) for g in range(32)]) == 4
    # the padded forward layout would conflict here (what the reverse kernel used before)
    assert wavefronts([pad2(2 * g + 1) for g in range(32)]) > 4
    # store_group<4>: lanes 4-7 of a phase store their pairs in the order 2, 3, 0, 1, index ^ bit 1 of g
    for e in range(4):
        slots = []
        for g in range(32):
            rot = (g >> 2) & 1
            slots.append((4 * g + (e ^ (2 if rot else 0))) ^ ((g >> 1) & 1))
        assert wavefronts(slots) == 4
        assert sorted(slots) == sorted(lay(4 * g + ee) for g in range(32) for ee in [e ^ (2 if (g >> 2) & 1 else 0)])
    # cp.async staging: consecutive slots
    assert wavefronts([lay(k) for k in range(32)]) == 4


def test_wpt_layouts():
    # k_wpt_fwd_tile, R = 8: windows at stride 9 padded slots, stores at stride 4 + (g >> 1)
    for q in range(0, 15):
        assert wavefronts([9 * g + q + q // 8 for g in range(32)]) == 4
    for e in range(4):
        assert wavefronts([4 * g + (g >> 1) + e for g in range(32)]) == 4
    # k_wpt_rev_tile, kRS = 8, -DJWC_WPT_REV_PAD4=1 (the A/B alternative): windows at stride 5 (pad2 of 4 g' + 3 - w),
    # rotated stores at stride 10
    for w in range(8):
        assert wavefronts([5 * g + (3 - w) + ((3 - w) >> 2) + 64 for g in range(32)]) == 4
    for e in range(8):
        plain = [10 * g + e + (e >> 2) for g in range(32)]
        assert wavefronts(plain) == 8  # groups g and g + 4 share a bank group: 2-way
        rotated = []
        for g in range(32):
            rot = (g >> 2) & 1
            ee = e ^ 4 if rot else e
            rotated.append(10 * g + ee + (ee >> 2))
        assert wavefronts(rotated) == 4


def rl(k):
    return k + (k >> 3)


def test_wpt_reverse_pad_per_8_and_tma_images():
    """k_wpt_rev_tile as shipped: one pad slot per 8 (rl).  A thread's window is the 4-slot blocks G and G - 1 with
    G = g0 + group number, g0 + gl == 2 at every level of the C3 geometry: the loads of block G (phases start at an
    even block) are conflict-free, the loads of block G - 1 collide in one lane pair per phase (2 passes - what the ncu
    source page shows for 6 of the 16 window loads); the 8-slot result runs are conflict-free without a rotated order."""
    for w in range(4):
        assert wavefronts([rl(4 * (2 + g)) + 3 - w for g in range(32)]) == 4
        assert wavefronts([rl(4 * (1 + g)) + 3 - w for g in range(32)]) == 8
    for e in range(8):
        assert wavefronts([rl(8 * g) + e for g in range(32)]) == 4
    # images of the TMA stores (SWIZZLE_128B: 16-byte chunk c of 128-byte row r sits at c ^ (r & 7)):
    # reverse - thread g owns row g; forward - thread g owns half of row g >> 1 of two leaf segments
    for e in range(8):
        assert wavefronts([8 * g + (e ^ (g & 7)) for g in range(32)]) == 4
    for e in range(4):
        assert wavefronts([8 * (g >> 1) + ((4 * (g & 1) + e) ^ ((g >> 1) & 7)) for g in range(32)]) == 4
