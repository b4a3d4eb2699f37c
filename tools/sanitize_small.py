"""Small renderings through every N = 2048 instantiation for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from jadespectrogram_b200 import Engine
rng = np.random.default_rng(0)
for ch, mix, hop, n in [(2, "absmean", 512, 512 * 40), (1, "absmean", 256, 256 * 90), (4, "absmean", 512, 512 * 30),
                        (2, "absmean", 510, 510 * 30), (2, "absmean", 205, 205 * 40), (2, "max", 512, 512 * 20)]:
    eng = Engine(0, sample_rate=48000.0, fft_size=2048, hop=hop, channels=ch, window="hann", mix_mode=mix, max_push=512)
    x = (rng.random((3, ch, n), dtype=np.float32) - 0.5)
    pix, db = eng.render_batch(x, want_db=True)
    pix2, _ = eng.render_batch(x)
    assert np.array_equal(pix, pix2)
    print(ch, mix, hop, eng.kernel_name, pix.shape, int(pix.sum() & 0xffff))
