"""Small renderings through the fast-path instantiations for compute-sanitizer (memcheck / racecheck / synccheck).
Run with JADE_MAX_GRID=2 JADE_RUN_MIN=1 so that the long-run (tensor-memory ring) instantiations are taken and every warp
walks a run of several columns."""
import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from jadespectrogram_b200 import Engine
rng = np.random.default_rng(0)
CASES = [(2048, 2, "absmean", 512, 512 * 60, {}), (2048, 1, "absmean", 256, 256 * 120, {}), (2048, 1, "absmean", 512, 512 * 60, {}),
         (2048, 4, "absmean", 512, 512 * 30, {}), (2048, 2, "absmean", 510, 510 * 30, {}), (2048, 2, "max", 512, 512 * 20, {}),
         (1024, 1, "absmean", 512, 512 * 60, {}), (512, 2, "absmean", 256, 256 * 60, {}),
         (16384, 1, "absmean", 4096, 4096 * 14, {}),
         (65536, 1, "absmean", 1024, 65536 + 1024 * 6, dict(row_map="log_maxpool", rows=216, fmin=20.0, fmax=20000.0))]
only = sys.argv[1:]
for N, ch, mix, hop, n, extra in CASES:
    if only and str(N) not in only:
        continue
    eng = Engine(0, sample_rate=48000.0, fft_size=N, hop=hop, channels=ch, window="hann", mix_mode=mix, max_push=512, **extra)
    x = (rng.random((3 if N <= 2048 else 1, ch, n), dtype=np.float32) - 0.5)
    pix, db = eng.render_batch(x, want_db=True)
    pix2, _ = eng.render_batch(x)
    assert np.array_equal(pix, pix2)
    print(N, ch, mix, hop, eng.kernel_name, pix.shape, int(pix.sum() & 0xffff), flush=True)
    eng.close()
