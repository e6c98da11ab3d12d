"""CPU: sanity properties of the Atari preprocessing restatement (oracle/atari_oracle.py).  The resize follows the
published algorithm of image 0.23.14 (not vendored in the reference: parity unpinned, see the oracle header); what can be
checked without the crate are the algorithm's invariants and the parts the reference spells out itself."""
import numpy as np

from oracle import atari_oracle as ato


def test_tap_tables_are_normalised_and_cover_the_support():
    for n_in in (210, 160, 84, 100):
        for left, ws in ato._taps(n_in, 84):
            assert 0 <= left < n_in and left + len(ws) <= n_in
            assert abs(float(np.sum(np.array(ws, np.float64))) - 1.0) < 1e-6
            assert all(w >= 0 for w in ws)
    # 210 -> 84: ratio 2.5, support 2.5 => at most 6 taps; 84 -> 84: the identity (one tap of weight 1 around the centre)
    assert max(len(ws) for _, ws in ato._taps(210, 84)) <= 6
    for o, (left, ws) in enumerate(ato._taps(84, 84)):
        w = dict(zip(range(left, left + len(ws)), ws))
        assert w[o] == np.float32(1.0) and all(v == 0 for k, v in w.items() if k != o)


def test_resize_of_constant_and_identity_images():
    img = np.full((210, 160, 3), 137, np.uint8)
    assert np.all(ato.resize_triangle_u8(img) == 137)
    rng = np.random.default_rng(0)
    same = rng.integers(0, 256, (84, 84, 3), dtype=np.uint8)
    assert np.array_equal(ato.resize_triangle_u8(same), same)


def test_grey_weights_follow_the_reference_channel_order():
    # env.rs:168-176: (b, g, r) name bytes 0, 1, 2, and the 0.299 weight multiplies `r` = byte 2
    img = np.zeros((84, 84, 3), np.uint8)
    img[..., 2] = 200
    assert np.all(ato.warp_and_grayscale(img) == int(np.float32(0.299) * np.float32(200)))
    img[:] = 0
    img[..., 0] = 200
    assert np.all(ato.warp_and_grayscale(img) == int(np.float32(0.114) * np.float32(200)))


def test_stack_order_reset_and_reward_clip():
    rng = np.random.default_rng(1)
    f = rng.integers(0, 256, (3, 210, 160, 3), dtype=np.uint8)
    st = ato.FrameStack()
    o0 = st.reset(f[0])
    assert all(np.array_equal(o0[k], o0[0]) for k in range(4))          # env.rs:288-297
    o1 = st.step(f[1], f[2])
    assert np.array_equal(o1[1:], o0[:3])                                # stack_frame: older frames shift back
    assert np.array_equal(o1[0], ato.warp_and_grayscale(np.maximum(f[1], f[2])))
    assert [ato.clip_reward(r, True) for r in (-4.0, 0.0, 0.25)] == [-1.0, 0.0, 1.0]
    assert ato.clip_reward(-4.0, False) == -4.0
