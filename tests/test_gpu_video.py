"""decode -> transfer -> encode: the chunked, threaded video driver gives the frames a sequential
loop over the uint8 API gives, in order, for every method; and an OpenCV file round trip runs end to end."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _frames(n, h=48, w=80, seed=3):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    out = []
    for k in range(n):
        t = np.clip(base.astype(np.int16) + rng.integers(-20, 20, base.shape) + k, 0, 255).astype(np.uint8)
        r = np.clip(base[::-1].astype(np.int16) * 0.8 + 30 + rng.integers(-10, 10, base.shape), 0, 255).astype(np.uint8)
        out.append((t, r))
    return out


@pytest.mark.parametrize("method", ["mkl", "reinhard", "idt"])
@pytest.mark.parametrize("rgb", [False, True])
def test_chunked_driver_equals_sequential_calls(method, rgb):
    import color_transfer_b200  # noqa: F401
    from color_transfer_b200 import batch, video
    frames = _frames(11)
    got = {}
    np.random.seed(7)
    n = video.transfer_frames(frames, lambda i, f: got.__setitem__(i, f.copy()), method=method, chunk=4, rgb=rgb)
    assert n == 11 and sorted(got) == list(range(11))
    flip = (lambda x: x[..., ::-1]) if rgb else (lambda x: x)
    np.random.seed(7)
    for i, (t, r) in enumerate(frames):
        t1, r1 = np.ascontiguousarray(flip(t))[None], np.ascontiguousarray(flip(r))[None]
        want = batch.idt_frames_u8(t1, r1) if method == "idt" else batch.linear_transfer_frames_u8(method, t1, r1)
        np.testing.assert_array_equal(got[i], flip(want[0]), err_msg=f"frame {i}")


def test_opencv_file_round_trip(tmp_path):
    cv2 = pytest.importorskip("cv2")
    import color_transfer_b200  # noqa: F401
    from color_transfer_b200 import video
    frames = _frames(9, 64, 96)
    paths = [tmp_path / "left.mp4", tmp_path / "right.mp4"]
    for p, idx in zip(paths, (0, 1)):
        w = cv2.VideoWriter(str(p), cv2.VideoWriter_fourcc(*"mp4v"), 25.0, (96, 64))
        if not w.isOpened():
            pytest.skip("this OpenCV build cannot write mp4v")
        for f in frames:
            w.write(f[idx])
        w.release()
    out = tmp_path / "corrected.mp4"
    n = video.transfer_stereo_video(paths[0], paths[1], out, method="mkl", chunk=4)
    assert n == 9
    cap = cv2.VideoCapture(str(out))
    count = 0
    while True:
        ok, f = cap.read()
        if not ok:
            break
        assert f.shape == (64, 96, 3)
        count += 1
    assert count == 9
    with pytest.raises(IOError):
        video.transfer_stereo_video(tmp_path / "missing.mp4", paths[1], out)
