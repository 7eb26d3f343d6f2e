"""Host restatement of the reduction the tower convolution's epilogue uses for the GroupNorm statistics (csrc/tower.cu,
conv3x3_kernel, `gn_partial`): every lane of a warp holds 8 values (4 groups x {sum, sum of squares} of its pixel); a halving
butterfly (xor 16 -> keep 4, xor 8 -> keep 2, xor 4 -> keep 1, then xor 2, xor 1) leaves value `idx(lane)` summed over all 32
lanes in every lane, 9 shuffles instead of 40.  This pins the lane -> value mapping the kernel's store uses
(`partial[... + c * 8 + idx]` from the lanes with lane % 4 == 0) and the fixed summation order."""
import numpy as np


def _shfl_xor(vals, mask):
    return vals[np.arange(32) ^ mask]


def butterfly(st):
    """st [32 lanes, 8] float32 -> (value held by each lane, index of that value), the kernel's operations in the kernel's order."""
    lane = np.arange(32)
    u16, u8, u4 = (lane & 16) != 0, (lane & 8) != 0, (lane & 4) != 0
    r4 = np.empty((32, 4), np.float32)
    for i in range(4):
        keep = np.where(u16, st[:, i + 4], st[:, i])
        send = np.where(u16, st[:, i], st[:, i + 4])
        r4[:, i] = keep + _shfl_xor(send, 16)
    r2 = np.empty((32, 2), np.float32)
    for i in range(2):
        keep = np.where(u8, r4[:, i + 2], r4[:, i])
        send = np.where(u8, r4[:, i], r4[:, i + 2])
        r2[:, i] = keep + _shfl_xor(send, 8)
    keep = np.where(u4, r2[:, 1], r2[:, 0])
    send = np.where(u4, r2[:, 0], r2[:, 1])
    r1 = keep + _shfl_xor(send, 4)
    r1 = r1 + _shfl_xor(r1, 2)
    r1 = r1 + _shfl_xor(r1, 1)
    idx = np.where(u16, 4, 0) + np.where(u8, 2, 0) + np.where(u4, 1, 0)
    return r1.astype(np.float32), idx


def test_halving_butterfly_sums_every_value_over_the_warp():
    rs = np.random.RandomState(5)
    st = rs.standard_normal((32, 8)).astype(np.float32)
    got, idx = butterfly(st)
    want = st.astype(np.float64).sum(0)
    for lane in range(32):
        assert abs(float(got[lane]) - want[idx[lane]]) <= 1e-5 * np.abs(st[:, idx[lane]]).sum()
    # the four lanes of a quad hold the same value; the storing lanes (lane % 4 == 0) cover each of the 8 values exactly once
    assert all(got[l] == got[l & ~3] for l in range(32))
    assert sorted(int(idx[l]) for l in range(0, 32, 4)) == list(range(8))
    # value index -> (group within the 32-column chunk, {sum, sum of squares}): the layout conv_gn_finalize_kernel reads as float2
    assert [(int(i) >> 1, int(i) & 1) for i in sorted(set(idx))] == [(g, k) for g in range(4) for k in range(2)]


def test_masked_lanes_contribute_exact_zeros():
    """Pixels outside the image are SELECTED to zero before the butterfly (a NaN in an unused accumulator row must not leak)."""
    rs = np.random.RandomState(6)
    st = rs.standard_normal((32, 8)).astype(np.float32)
    act = rs.rand(32) > 0.4
    raw = st.copy()
    raw[~act] = np.nan
    masked = np.where(act[:, None], raw, np.float32(0))
    got, idx = butterfly(masked)
    want = st[act].astype(np.float64).sum(0)
    assert np.isfinite(got).all()
    for lane in range(32):
        assert abs(float(got[lane]) - want[idx[lane]]) <= 1e-5 * max(np.abs(st[act][:, idx[lane]]).sum(), 1e-6)
