"""e2e leg of bench.py alone (pinned host buffers through jade_render_batch); JADE_CHUNK_MB selects the chunk size."""
import sys, pathlib, time, os
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
from jadespectrogram_b200 import Engine, host_alloc, host_free
FS, N, HOP, CH, S = 48000.0, 2048, 512, 2, 64
nsamp = int(FS * 20)
eng = Engine(0, sample_rate=FS, fft_size=N, hop=HOP, channels=CH, window="hann", mix_mode="absmean", max_push=512)
ncols = eng.columns_for(nsamp)
h_in = host_alloc((S, CH, nsamp), np.float32)
h_pix = host_alloc((S, ncols, N // 2 + 1), np.uint32)
h_in[:] = (np.random.default_rng(1).random((S, CH, nsamp), dtype=np.float32) - 0.5)
for _ in range(2):
    eng.render_batch(h_in, out_pix=h_pix)
t0 = time.perf_counter()
steps = 6
for _ in range(steps):
    eng.render_batch(h_in, out_pix=h_pix)
dt = time.perf_counter() - t0
print(f"JADE_CHUNK_MB={os.environ.get('JADE_CHUNK_MB','default')}: {S*ncols*steps/dt/1e6:.2f} M frames/s e2e, "
      f"{(h_in.nbytes+h_pix.nbytes)*steps/dt/1e9:.1f} GB/s PCIe both directions")
